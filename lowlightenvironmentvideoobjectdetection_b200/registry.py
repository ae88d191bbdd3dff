"""Minimal stand-ins for the mmcv plumbing the hot-path modules sit behind.

The reference builds its modules through registries (``AGGREGATORS`` at
mmtracking/mmtrack/models/builder.py:9,53-55; ``ROI_EXTRACTORS`` at
mmdetection/mmdet/models/builder.py:8,47-49) and ``build_from_cfg``.  mmcv is not
installable in this image, so the same observable contract is carried here;
when the real mmtrack / mmdet packages ARE importable, ``register_into_openmmlab``
registers the B200 modules into their registries (``force=True``) so existing
configs pick them up unchanged.
"""
import functools

import torch
import torch.nn as nn


class Registry:
    """name -> class table with mmcv.utils.Registry's register_module/get/build contract."""

    def __init__(self, name):
        self._name = name
        self._module_dict = {}

    @property
    def name(self):
        return self._name

    @property
    def module_dict(self):
        return self._module_dict

    def __len__(self):
        return len(self._module_dict)

    def __contains__(self, key):
        return key in self._module_dict

    def get(self, key):
        return self._module_dict.get(key)

    def _register(self, cls, name=None, force=False):
        key = name or cls.__name__
        if not force and key in self._module_dict:
            raise KeyError('%s is already registered in %s' % (key, self._name))
        self._module_dict[key] = cls

    def register_module(self, name=None, force=False, module=None):
        if module is not None:
            self._register(module, name, force)
            return module

        def _deco(cls):
            self._register(cls, name, force)
            return cls
        return _deco

    def build(self, cfg, default_args=None):
        return build_from_cfg(cfg, self, default_args)


def build_from_cfg(cfg, registry, default_args=None):
    """mmcv.utils.build_from_cfg: ``cls(**cfg_without_type)``."""
    if not isinstance(cfg, dict):
        raise TypeError('cfg must be a dict, but got %s' % type(cfg))
    if 'type' not in cfg and not (default_args and 'type' in default_args):
        raise KeyError('`cfg` or `default_args` must contain the key "type"')
    args = dict(cfg)
    if default_args is not None:
        for k, v in default_args.items():
            args.setdefault(k, v)
    obj_type = args.pop('type')
    if isinstance(obj_type, str):
        obj_cls = registry.get(obj_type)
        if obj_cls is None:
            raise KeyError('%s is not in the %s registry' % (obj_type, registry.name))
    elif isinstance(obj_type, type):
        obj_cls = obj_type
    else:
        raise TypeError('type must be a str or valid type, but got %s' % type(obj_type))
    return obj_cls(**args)


AGGREGATORS = Registry('aggregator')
ROI_EXTRACTORS = Registry('roi_extractor')
HEADS = Registry('head')
MOTION = Registry('motion')           # mmtracking/mmtrack/models/builder.py:8 (FlowNetSimple)


def build_aggregator(cfg):
    """mmtracking/mmtrack/models/builder.py:53-55."""
    return build_from_cfg(cfg, AGGREGATORS)


def build_roi_extractor(cfg):
    """mmdetection/mmdet/models/builder.py:47-49."""
    return build_from_cfg(cfg, ROI_EXTRACTORS)


def build_head(cfg):
    return build_from_cfg(cfg, HEADS)


def build_motion(cfg):
    """mmtracking/mmtrack/models/builder.py:43-45."""
    return build_from_cfg(cfg, MOTION)


class ConvModule(nn.Module):
    """The slice of mmcv.cnn.ConvModule the hot-path modules use: conv or transposed conv (``conv_cfg=dict(type='deconv')``)
    + optional ReLU / LeakyReLU, sub-modules named ``conv`` / ``activate`` so reference checkpoints load (SURVEY section 5,
    checkpoint keys).  Weights use mmcv's default Kaiming-normal init."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 bias='auto', conv_cfg=None, norm_cfg=None, act_cfg=dict(type='ReLU'), inplace=True, **kwargs):
        super().__init__()
        if norm_cfg is not None:
            raise NotImplementedError('norm layers are not part of the hot path (reference configs use norm_cfg=None)')
        conv_type = 'Conv2d' if conv_cfg is None else conv_cfg.get('type', 'Conv2d')
        if conv_type not in ('Conv2d', 'Conv', 'deconv'):
            raise NotImplementedError('conv type %r is not supported' % conv_type)
        self.with_activation = act_cfg is not None
        self.with_norm = False
        with_bias = (bias == 'auto' or bool(bias))
        if conv_type == 'deconv':
            self.conv = nn.ConvTranspose2d(in_channels, out_channels, kernel_size, stride=stride, padding=padding,
                                           dilation=dilation, groups=groups, bias=with_bias)
        else:
            self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, stride=stride, padding=padding,
                                  dilation=dilation, groups=groups, bias=with_bias)
        slope = 0.0
        if self.with_activation:
            act_type = act_cfg.get('type', 'ReLU')
            if act_type == 'ReLU':
                self.activate = nn.ReLU(inplace=inplace)
            elif act_type == 'LeakyReLU':
                slope = act_cfg.get('negative_slope', 0.01)
                self.activate = nn.LeakyReLU(negative_slope=slope, inplace=inplace)
            else:
                raise NotImplementedError('only ReLU / LeakyReLU activations are supported')
        nn.init.kaiming_normal_(self.conv.weight, a=slope, mode='fan_out', nonlinearity='leaky_relu' if slope else 'relu')
        if self.conv.bias is not None:
            nn.init.constant_(self.conv.bias, 0)

    def forward(self, x):
        x = self.conv(x)
        if self.with_activation:
            x = self.activate(x)
        return x


def force_fp32(apply_to=None, out_fp16=False):
    """mmcv.runner.force_fp32: active only when the module sets ``fp16_enabled``; then the named tensor
    (or tuple-of-tensor) arguments are cast to fp32 and, with ``out_fp16``, the result back to fp16."""

    def deco(fn):
        @functools.wraps(fn)
        def wrapper(self, *args, **kwargs):
            if not getattr(self, 'fp16_enabled', False):
                return fn(self, *args, **kwargs)
            import inspect
            names = list(inspect.signature(fn).parameters)[1:]

            def cast(v):
                if torch.is_tensor(v) and v.dtype in (torch.float16, torch.bfloat16):
                    return v.float()
                if isinstance(v, (list, tuple)):
                    return type(v)(cast(u) for u in v)
                return v
            args = list(args)
            for i, a in enumerate(args):
                if apply_to is None or (i < len(names) and names[i] in apply_to):
                    args[i] = cast(a)
            for k in kwargs:
                if apply_to is None or k in apply_to:
                    kwargs[k] = cast(kwargs[k])
            out = fn(self, *args, **kwargs)
            if out_fp16 and torch.is_tensor(out):
                out = out.half()
            return out
        return wrapper
    return deco


def register_into_openmmlab(strict=False):
    """If the real mmtrack / mmdet packages are importable, register the B200 modules into their
    registries (same names, force=True).

    Returns ``(touched, failed)``: the registries that now hold the B200 modules, and a dict
    ``registry name -> reason`` for every registry that does not (package absent -> ``'not installed'``; anything else is
    the exception's text, and is also logged as a warning -- a caller can tell that the reference modules are still in
    use).  ``strict=True`` raises on any failure other than the package being absent.

    The B200 modules are inference-only (no backward): their ``forward`` raises while autograd is recording in training
    mode instead of silently returning tensors without a graph."""
    import warnings
    touched, failed = [], {}
    from . import aggregators, roi_extractors

    def attempt(name, importer, entries):
        try:
            reg = importer()
        except ImportError:
            failed[name] = 'not installed'
            return
        try:
            for mod_name, mod in entries:
                reg.register_module(name=mod_name, force=True, module=mod)
            touched.append(name)
        except Exception as e:   # registration itself failed: report, do not hide
            failed[name] = '%s: %s' % (type(e).__name__, e)
            warnings.warn('vodagg: could not register into %s (%s); the reference modules stay in use' % (name, failed[name]))
            if strict:
                raise

    def mm_agg():
        from mmtrack.models.builder import AGGREGATORS as reg
        return reg

    def mm_roi():
        from mmdet.models.builder import ROI_EXTRACTORS as reg
        return reg

    attempt('mmtrack.AGGREGATORS', mm_agg, [('SelsaAggregator', aggregators.SelsaAggregator),
                                             ('EmbedAggregator', aggregators.EmbedAggregator)])
    attempt('mmdet.ROI_EXTRACTORS', mm_roi, [('SingleRoIExtractor', roi_extractors.SingleRoIExtractor),
                                             ('TemporalRoIAlign', roi_extractors.TemporalRoIAlign)])
    return touched, failed


def inference_only(fn):
    """Decorator for the ``forward`` of the B200 drop-ins: the kernels have no backward, so a call that the reference
    would have differentiated through (module in training mode, autograd recording, an input that requires grad -- i.e. a
    real training step, where the inputs come from the backbone) raises instead of silently cutting the graph.  Everything
    else runs under ``torch.no_grad()``."""

    def needs_grad(v):
        if torch.is_tensor(v):
            return v.requires_grad
        if isinstance(v, (list, tuple)):
            return any(needs_grad(u) for u in v)
        return False

    @functools.wraps(fn)
    def wrapper(self, *args, **kwargs):
        if self.training and torch.is_grad_enabled() and (needs_grad(args) or needs_grad(tuple(kwargs.values()))):
            raise RuntimeError('%s (vodagg B200 drop-in) is inference-only: call model.eval() or run under torch.no_grad(); '
                               'a training run must keep the reference module' % type(self).__name__)
        with torch.no_grad():
            return fn(self, *args, **kwargs)
    return wrapper
