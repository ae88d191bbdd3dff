"""Clip sharding across the GPUs of one box + the end-of-run gather of detections.

The hot path has no data-path collective: whole video clips are independent (the per-video frame memory
forbids splitting a clip), so rank r owns a contiguous range of clips exactly as the reference's
``DistributedVideoSampler`` does (mmtracking/mmtrack/datasets/samplers/distributed_video_sampler.py:24-45).
The single exchange is the result gather, which the reference performs through pickle files in a shared
tmpdir + broadcast + barrier (mmtracking/mmtrack/apis/test.py:125-173); here it is one all_gather of
fixed-shape detection tensors over NCCL/NVLink (gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def shard_clips(clip_lengths, world_size, rank):
    """Contiguous chunk of whole clips for ``rank`` (distributed_video_sampler.py:24-40: the list of clips is
    split into ``world_size`` chunks of near-equal clip count; a clip is never split).

    Returns the list of clip indices owned by ``rank``."""
    n = len(clip_lengths)
    if world_size > n:
        raise ValueError('only %d clips for %d ranks: every rank must own at least one whole clip' % (n, world_size))
    base, rem = divmod(n, world_size)
    start = rank * base + min(rank, rem)
    return list(range(start, start + base + (1 if rank < rem else 0)))


def pack_detections(det_bboxes, det_labels, max_per_img=100):
    """Per-frame (dets [k,5], labels [k]) lists -> fixed-shape [frames, max_per_img, 6] + counts [frames]."""
    frames = len(det_bboxes)
    dev = det_bboxes[0].device if frames else torch.device('cpu')
    packed = torch.zeros((frames, max_per_img, 6), dtype=torch.float32, device=dev)
    counts = torch.zeros((frames,), dtype=torch.int32, device=dev)
    for i, (d, l) in enumerate(zip(det_bboxes, det_labels)):
        k = min(d.shape[0], max_per_img)
        packed[i, :k, :5] = d[:k]
        packed[i, :k, 5] = l[:k].to(torch.float32)
        counts[i] = k
    return packed, counts


def gather_detections(packed, counts, frames_per_rank=None):
    """all_gather of every rank's packed detections.  Ranks may own different frame counts: tensors are padded
    to the maximum and trimmed after the exchange.  Returns a list (one entry per rank) of (packed, counts).
    Without an initialised process group (single GPU) it returns [(packed, counts)]."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [(packed, counts)]
    world = dist.get_world_size()
    n_local = torch.tensor([packed.shape[0]], dtype=torch.int64, device=packed.device)
    if frames_per_rank is None:
        sizes = [torch.zeros_like(n_local) for _ in range(world)]
        dist.all_gather(sizes, n_local)
        frames_per_rank = [int(s.item()) for s in sizes]
    n_max = max(frames_per_rank)
    pad_p = packed.new_zeros((n_max,) + tuple(packed.shape[1:]))
    pad_c = counts.new_zeros((n_max,))
    pad_p[:packed.shape[0]] = packed
    pad_c[:counts.shape[0]] = counts
    out_p = [torch.empty_like(pad_p) for _ in range(world)]
    out_c = [torch.empty_like(pad_c) for _ in range(world)]
    dist.all_gather(out_p, pad_p)
    dist.all_gather(out_c, pad_c)
    return [(out_p[r][:frames_per_rank[r]], out_c[r][:frames_per_rank[r]]) for r in range(world)]
