"""ctypes binding of libvodagg.so (the C ABI declared in include/vodagg.h).

The product path has NO fallback: if the library is missing or a call fails the
wrappers raise.  Nothing here imports the oracle.
"""
import ctypes
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libvodagg.so')

VOD_DTYPE_F32 = 0
VOD_DTYPE_BF16 = 1

_c = ctypes
_P = _c.c_void_p
_I = _c.c_int
_F = _c.c_float
_SZ = _c.c_size_t

# symbol -> (restype, argtypes); must list every symbol include/vodagg.h declares
SIGNATURES = {
    'vod_version': (_I, []),
    'vod_last_error': (_c.c_char_p, []),
    'vod_device_is_sm100': (_I, []),
    'vod_kernel_launch_count': (_c.c_longlong, []),
    'vod_nchw_to_nhwc': (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _P]),
    'vod_rows_l2norm': (_I, [_P, _P, _P, _I, _I, _P]),
    'vod_roi_align_fwd': (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _F, _I, _I, _I, _P]),
    'vod_flow_warp': (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    'vod_flow_warp_shared': (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P]),
    'vod_flow_warp_lowres': (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _c.c_double, _F, _F, _P]),
    'vod_embed_weighted_sum': (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _P, _SZ, _P]),
    'vod_fgfa_warp_weighted_sum': (_I, [_P, _P, _P, _P, _P, _I, _P, _I, _I, _I, _I, _I, _I, _I, _P, _SZ, _P]),
    'vod_selsa_attn_workspace_bytes': (_SZ, [_I, _I, _I, _I]),
    'vod_selsa_attn': (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _F, _I, _I, _I, _I, _P, _SZ, _P]),
    'vod_selsa_residual_relu': (_I, [_P, _P, _P, _I, _I, _P, _c.c_long, _P]),
    'vod_msra_workspace_bytes': (_SZ, [_I, _I, _I, _I, _I]),
    'vod_msra_topk_sample': (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P, _SZ, _P]),
    'vod_msra_overflow_counter_offset': (_SZ, [_I, _I, _I, _I]),
    'vod_msra_gemm_candidates': (_I, [_P, _P, _P, _I, _I, _I, _I, _P]),
    'vod_tafa_weighted_sum': (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    'vod_tafa_keyproj_chunk': (_I, [_I, _I, _I, _I]),
    'vod_tafa_keyproj_logits': (_I, [_P, _P, _I, _P, _I, _I, _I, _I, _I, _I, _I, _P]),
    'vod_tafa_weighted_sum_logits': (_I, [_P, _P, _I, _P, _I, _I, _I, _I, _I, _I, _P]),
    'vod_mdcn_im2col': (_I, [_P, _P, _P, _P] + [_I] * 12 + [_P]),
    'vod_temporal_softmax_fuse': (_I, [_P, _P, _P, _I, _I, _c.c_long, _P]),
    'vod_nms_workspace_bytes': (_SZ, [_I, _I]),
    'vod_batched_nms': (_I, [_P, _P, _P, _I, _c.POINTER(_I), _I, _F, _I, _I, _P, _P, _P, _SZ, _P]),
    'vod_batched_nms_ex': (_I, [_P, _P, _P, _I, _c.POINTER(_I), _I, _F, _I, _I, _P, _I, _P, _P, _P, _P, _P, _SZ, _P]),
    'vod_bbox_decode_candidates': (_I, [_P, _P, _P, _I, _I, _I, _c.POINTER(_F), _c.POINTER(_F), _F, _F, _F,
                                        _c.POINTER(_F), _F, _P, _P, _P, _P, _P]),
    'vod_rpn_decode_topk': (_I, [_P, _P, _P, _P, _I, _I, _I, _F, _F, _F, _P]),
}

_lib = None
_selftest = None
_lock = threading.Lock()
launch_count = 0  # number of kernel-launching C-ABI calls made through this module (bench bookkeeping)


# include/vodagg_selftest.h (libvodagg_selftest.so: test-only kernels, never on an operator's path)
SELFTEST_LIB_PATH = os.path.join(_HERE, 'libvodagg_selftest.so')
SELFTEST_SIGNATURES = {
    'vod_last_error': (_c.c_char_p, []),
    'vod_test_gemm_nt': (_I, [_P, _P, _P, _I, _I, _I, _I, _P]),
}


class VodError(RuntimeError):
    pass


def load():
    """Loads libvodagg.so; raises if it has not been built (no fallback path exists)."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise VodError(
                        'libvodagg.so not found at %s: build it with '
                        '`python -m lowlightenvironmentvideoobjectdetection_b200.build` '
                        '(there is no CPU / PyTorch fallback for these ops)' % LIB_PATH)
                lib = ctypes.CDLL(LIB_PATH)
                for name, (res, args) in SIGNATURES.items():
                    fn = getattr(lib, name)
                    fn.restype = res
                    fn.argtypes = args
                _lib = lib
    return _lib


def load_selftest():
    """Loads libvodagg_selftest.so (the tcgen05 / TMA building-block test kernels of include/vodagg_selftest.h)."""
    global _selftest
    if _selftest is None:
        with _lock:
            if _selftest is None:
                if not os.path.exists(SELFTEST_LIB_PATH):
                    raise VodError('libvodagg_selftest.so not found at %s: build it with '
                                   '`python -m lowlightenvironmentvideoobjectdetection_b200.build`' % SELFTEST_LIB_PATH)
                lib = ctypes.CDLL(SELFTEST_LIB_PATH)
                for name, (res, args) in SELFTEST_SIGNATURES.items():
                    fn = getattr(lib, name)
                    fn.restype = res
                    fn.argtypes = args
                _selftest = lib
    return _selftest


def call_selftest(name, *args):
    lib = load_selftest()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        msg = lib.vod_last_error()
        raise VodError('%s failed (%d): %s' % (name, rc, msg.decode() if msg else ''))


def check(rc, what):
    if rc != 0:
        msg = load().vod_last_error()
        raise VodError('%s failed (%d): %s' % (what, rc, msg.decode() if msg else ''))


def ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def stream_ptr(device=None):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def call(name, *args):
    """Invoke a kernel-launching entry point and raise on a non-zero status."""
    global launch_count
    launch_count += 1
    check(getattr(load(), name)(*args), name)


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise VodError('vodagg ops run on CUDA tensors only (got a %s tensor); '
                           'there is no CPU fallback' % t.device.type)


class Workspace:
    """Scratch memory for the kernels, 1024-byte aligned (TMA-swizzled tiles need it; torch's caching allocator only
    guarantees 512).  One grow-only buffer per (device, stream): callers on different streams never share scratch
    (SURVEY 8b threading contract).  While a CUDA graph is being captured nothing is cached: every request is a fresh
    allocation from the graph's private pool, so a replay can never touch a buffer that was later outgrown and freed."""

    def __init__(self):
        self._buf = {}

    @staticmethod
    def _aligned(nbytes, device):
        raw = torch.empty(max(int(nbytes), 1024) + 1024, dtype=torch.uint8, device=device)
        off = (-raw.data_ptr()) % 1024
        return raw[off:off + max(int(nbytes), 1024)]

    def get(self, nbytes, device):
        if device.type == 'cuda' and torch.cuda.is_current_stream_capturing():
            return self._aligned(nbytes, device)
        stream = torch.cuda.current_stream(device).cuda_stream if device.type == 'cuda' else 0
        key = (device.type, device.index, stream)
        buf = self._buf.get(key)
        if buf is None or buf.numel() < nbytes:
            buf = self._aligned(nbytes, device)
            self._buf[key] = buf
        return buf
