"""Clip sharding + detection gather on a world_size-2 gloo group (CPU)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lowlightenvironmentvideoobjectdetection_b200 import parallel


def test_shard_clips_contiguous_and_complete():
    lengths = [30, 12, 45, 8, 19, 27, 33]
    for world in (1, 2, 3, 4, 7):
        owned = [parallel.shard_clips(lengths, world, r) for r in range(world)]
        flat = [c for o in owned for c in o]
        assert flat == list(range(len(lengths)))                      # every clip exactly once, order kept
        assert max(len(o) for o in owned) - min(len(o) for o in owned) <= 1
        for o in owned:
            assert o == list(range(o[0], o[0] + len(o)))              # contiguous: whole clips, never split
    try:
        parallel.shard_clips([3, 4], 3, 0)
        assert False
    except ValueError:
        pass


def test_pack_detections():
    d = [torch.rand(3, 5), torch.rand(0, 5), torch.rand(150, 5)]
    l = [torch.tensor([1, 2, 3]), torch.zeros(0, dtype=torch.long), torch.arange(150)]
    p, c = parallel.pack_detections(d, l, 100)
    assert p.shape == (3, 100, 6) and c.tolist() == [3, 0, 100]
    assert torch.equal(p[0, :3, :5], d[0]) and p[0, :3, 5].tolist() == [1., 2., 3.]
    assert torch.equal(p[2, :, :5], d[2][:100])
    assert parallel.gather_detections(p, c)[0][0] is p                 # no process group: identity


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    clips = parallel.shard_clips([4, 2, 3], world, rank)               # rank 0: clips 0,1 ; rank 1: clip 2
    frames = sum([4, 2, 3][c] for c in clips)
    g = torch.Generator().manual_seed(100 + rank)
    dets = [torch.rand(5 + rank, 5, generator=g) for _ in range(frames)]
    labels = [torch.full((5 + rank,), rank, dtype=torch.long) for _ in range(frames)]
    packed, counts = parallel.pack_detections(dets, labels, 10)
    gathered = parallel.gather_detections(packed, counts)
    ok = len(gathered) == world
    ok &= [g_[0].shape[0] for g_ in gathered] == [6, 3]
    ok &= torch.equal(gathered[rank][0], packed) and torch.equal(gathered[rank][1], counts)
    other = 1 - rank
    ok &= bool((gathered[other][1] == 5 + other).all()) and bool((gathered[other][0][:, :5 + other, 5] == other).all())
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_gather_detections_gloo_world2():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]
