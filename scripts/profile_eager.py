"""Host-side profile (cProfile) of the eager drop-in calls: SelsaRoIHead.simple_test uncached and with the reference-frame cache.
    python scripts/profile_eager.py"""
import cProfile
import os
import pstats
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

args = types.SimpleNamespace(steps=6, warmup=1, no_overlap=False, clip_len=40)
ctx = bench.Ctx(args)
cfg = bench.CONFIGS['cfg3']
run = bench.SelsaRunner(ctx, cfg, n_sets=4, pinned=False)
T = cfg['T']
metas = [dict(img_shape=bench.IMG_SHAPE, scale_factor=(1., 1., 1., 1.))]
memo_metas = [dict(video_id=0, frame_id=-(t + 1), img_shape=bench.IMG_SHAPE, scale_factor=(1., 1., 1., 1.)) for t in range(T - 1)]
ref_x0, props0 = run.dev_sets[0]


def cached(k):
    for i in range(k):
        key_x, key_props = run.dev_sets[i % 4]
        km = dict(video_id=0, frame_id=i, img_shape=bench.IMG_SHAPE, scale_factor=(1., 1., 1., 1.))
        ref_x = torch.cat((ref_x0[:T - 1], key_x[T - 1:T]), 0)
        props = [props0[t] for t in range(T - 1)] + [key_props[T]]
        run.head.simple_test((key_x[T - 1:T],), (ref_x,), [key_props[T]], props, [km], rescale=False, ref_img_metas=memo_metas + [km])


with torch.no_grad(), bench.library_math(True):
    for fn, name in ((lambda k: [bench.run_step(run.head, *run.dev_sets[i % 4], metas) for i in range(k)], 'uncached'), (cached, 'cached')):
        fn(3)
        torch.cuda.synchronize()
        pr = cProfile.Profile()
        pr.enable()
        fn(20)
        pr.disable()
        torch.cuda.synchronize()
        print('=====', name, 'eager call, 20 frames: host-side cumulative times')
        st = pstats.Stats(pr)
        st.sort_stats('tottime').print_stats(18)
