"""oracle/make_ref.py -- TEST INFRASTRUCTURE ONLY.

    python -m oracle.make_ref          (also run by __graft_entry__.build() when /root/reference is present)

Stages the reference's OWN hot-path Python files, byte for byte, from /root/reference into oracle/_ref/ (same relative paths).
oracle/_ref/ is git-ignored -- reference sources never enter this repository's history -- but it is not gpurun-ignored, so the
staged copy travels to the GPU box with the snapshot, exactly like a compiled oracle/_ref/*.so would for a C reference.  There
``oracle/ref_shim.py`` loads the files unmodified (under its mmcv stand-ins) and ``oracle/ref_step.py`` runs the reference's own
step on the box's host cores (``bench.py --impl reference``, ``cpu_baseline.kind = "reference"``) and on the B200 in
torch-eager + torchvision CUDA ops (``eager_cuda_reference``): the comparators SURVEY 8(d) / BASELINE.md section 4 name.

The list below is SURVEY Appendix C's six files plus the callers the end-to-end step needs (bbox_nms, the bbox coder, RPN head).
"""
import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, '_ref')
SRC = os.environ.get('VOD_REFERENCE_SRC', '/root/reference')

FILES = [
    'mmdetection/mmdet/models/roi_heads/roi_extractors/base_roi_extractor.py',
    'mmdetection/mmdet/models/roi_heads/roi_extractors/single_level_roi_extractor.py',
    'mmdetection/mmdet/core/post_processing/bbox_nms.py',
    'mmdetection/mmdet/core/bbox/coder/base_bbox_coder.py',
    'mmdetection/mmdet/core/bbox/coder/delta_xywh_bbox_coder.py',
    'mmdetection/mmdet/models/dense_heads/rpn_head.py',
    'mmtracking/mmtrack/models/aggregators/selsa_aggregator.py',
    'mmtracking/mmtrack/models/aggregators/embed_aggregator.py',
    'mmtracking/mmtrack/models/roi_heads/roi_extractors/temporal_roi_align.py',
    'mmtracking/mmtrack/models/roi_heads/roi_extractors/single_level_roi_extractor.py',
    'mmtracking/mmtrack/core/motion/flow.py',
    'mmtracking/mmtrack/models/motion/flownet_simple.py',
    'mmtracking/mmtrack/models/aggregators/denoising2_aggregator.py',
]


def stage(verbose=True):
    """Copies FILES from the reference tree into oracle/_ref/.  Returns the number of files staged (0 when the reference tree is
    absent, e.g. on the GPU box, where the already staged copy is used)."""
    if not os.path.isdir(os.path.join(SRC, 'mmtracking', 'mmtrack')):
        return 0
    n = 0
    for rel in FILES:
        src, dst = os.path.join(SRC, rel), os.path.join(DEST, rel)
        if not os.path.exists(src):
            raise FileNotFoundError(src)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not (os.path.exists(dst) and filecmp.cmp(src, dst, shallow=False)):
            shutil.copyfile(src, dst)
        n += 1
    with open(os.path.join(DEST, 'README'), 'w') as f:
        f.write('Unmodified files of the reference, staged by oracle/make_ref.py from %s (git-ignored; test infrastructure).\n' % SRC)
    if verbose:
        print('staged %d reference files into %s' % (n, DEST))
    return n


if __name__ == '__main__':
    sys.exit(0 if stage() or os.path.isdir(DEST) else 1)
