"""Stand-in for the handful of mmcv / mmdet names HiP-AD's model code imports (mmcv, mmdet and mmdet3d are absent
from this image and from the GPU box).  TEST / BENCH INFRASTRUCTURE ONLY: it exists so that the UNMODIFIED reference
decoder (vendored under baseline/_ref/hipad by harness/vendor.py) can be constructed and run with different
``projects.mmdet3d_plugin.ops`` packages plugged in.  Nothing here computes anything on the hot
path: registries, builders, init helpers, pass-through decorators, and the layers mmcv would build from config dicts
(``nn.Linear``, ``nn.LayerNorm``, ``nn.ReLU``, ``nn.Dropout``).
"""
import functools
import inspect
import sys
import types

import torch
import torch.nn as nn


class Registry:
    def __init__(self, name):
        self.name, self.module_dict = name, {}

    def register_module(self, name=None, force=False, module=None):
        def deco(cls):
            self.module_dict[name or cls.__name__] = cls
            return cls
        return deco(module) if module is not None else deco

    def get(self, key):
        return self.module_dict.get(key)

    def build(self, cfg, **kw):
        return build_from_cfg(cfg, self, kw or None)

    def __contains__(self, key):
        return key in self.module_dict


def build_from_cfg(cfg, registry, default_args=None):
    if cfg is None:
        return None
    cfg = dict(cfg)
    if default_args:
        for k, v in default_args.items():
            cfg.setdefault(k, v)
    kind = cfg.pop("type")
    cls = registry.get(kind) if isinstance(kind, str) else kind
    if cls is None:
        raise KeyError("%s is not registered in %s" % (kind, registry.name))
    return cls(**cfg)


class BaseModule(nn.Module):
    def __init__(self, init_cfg=None):
        super().__init__()
        self.init_cfg = init_cfg

    def init_weights(self):
        for m in self.children():
            if hasattr(m, "init_weights"):
                m.init_weights()


class Sequential(BaseModule, nn.Sequential):
    def __init__(self, *args, init_cfg=None):
        BaseModule.__init__(self, init_cfg)
        nn.Sequential.__init__(self, *args)


class Scale(nn.Module):
    def __init__(self, scale=1.0):
        super().__init__()
        self.scale = nn.Parameter(torch.tensor(scale, dtype=torch.float))

    def forward(self, x):
        return x * self.scale


def bias_init_with_prob(prior_prob):
    import math
    return float(-math.log((1 - prior_prob) / prior_prob))


def xavier_init(module, gain=1, bias=0, distribution="normal"):
    if getattr(module, "weight", None) is not None:
        (nn.init.xavier_uniform_ if distribution == "uniform" else nn.init.xavier_normal_)(module.weight, gain=gain)
    if getattr(module, "bias", None) is not None:
        nn.init.constant_(module.bias, bias)


def constant_init(module, val, bias=0):
    if getattr(module, "weight", None) is not None:
        nn.init.constant_(module.weight, val)
    if getattr(module, "bias", None) is not None:
        nn.init.constant_(module.bias, bias)


def build_norm_layer(cfg, num_features, postfix=""):
    kind = cfg.get("type", "LN")
    if kind != "LN":
        raise KeyError("norm layer %s not provided by the harness stand-in" % kind)
    return "ln" + str(postfix), nn.LayerNorm(num_features, eps=cfg.get("eps", 1e-5))


def build_activation_layer(cfg):
    kind = cfg.get("type", "ReLU")
    if kind == "ReLU":
        return nn.ReLU(inplace=cfg.get("inplace", False))
    if kind == "GELU":
        return nn.GELU()
    raise KeyError("activation %s not provided by the harness stand-in" % kind)


def build_dropout(cfg, default_args=None):
    cfg = dict(cfg or {})
    kind = cfg.get("type", "Dropout")
    if kind == "Dropout":
        return nn.Dropout(cfg.get("drop_prob", cfg.get("p", 0.5)))
    raise KeyError("dropout %s not provided by the harness stand-in" % kind)


def _passthrough_decorator(*dargs, **dkwargs):
    """@deco or @deco(...): returns the function unchanged."""
    if len(dargs) == 1 and callable(dargs[0]) and not dkwargs:
        return dargs[0]
    return lambda fn: fn


def auto_fp16(apply_to=None, out_fp32=False, supported_types=(nn.Module,)):
    """mmcv.runner.auto_fp16: when the owning module has ``fp16_enabled`` set, cast the named tensor arguments to
    half and (out_fp32) the outputs back to float.  The reference relies on this for flash-attn (attention.py:52)."""
    def wrapper(fn):
        names = list(inspect.signature(fn).parameters)

        @functools.wraps(fn)
        def new_fn(*args, **kwargs):
            self = args[0]
            if not getattr(self, "fp16_enabled", False):
                return fn(*args, **kwargs)
            want = set(names if apply_to is None else apply_to)

            def cast(x, src, dst):
                if torch.is_tensor(x):
                    return x.to(dst) if x.dtype == src else x
                if isinstance(x, (list, tuple)):
                    return type(x)(cast(v, src, dst) for v in x)
                if isinstance(x, dict):
                    return {k: cast(v, src, dst) for k, v in x.items()}
                return x
            new_args = [a if names[i] not in want else cast(a, torch.float32, torch.float16)
                        for i, a in enumerate(args)]
            new_kwargs = {k: (cast(v, torch.float32, torch.float16) if k in want else v) for k, v in kwargs.items()}
            out = fn(*new_args, **new_kwargs)
            return cast(out, torch.float16, torch.float32) if out_fp32 else out
        return new_fn
    return wrapper


def force_fp32(apply_to=None, out_fp16=False):
    return lambda fn: fn        # the harness runs the head in fp32 already


def reduce_mean(t):
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return t
    t = t.clone()
    dist.all_reduce(t.div_(dist.get_world_size()), op=dist.ReduceOp.SUM)
    return t


def weighted_loss(loss_func):
    @functools.wraps(loss_func)
    def wrapper(pred, target, weight=None, reduction="mean", avg_factor=None, **kwargs):
        loss = loss_func(pred, target, **kwargs)
        if weight is not None:
            loss = loss * weight
        if avg_factor is None:
            return loss.mean() if reduction == "mean" else loss.sum() if reduction == "sum" else loss
        if reduction == "mean":
            return loss.sum() / (avg_factor + torch.finfo(torch.float32).eps)
        return loss
    return wrapper


@weighted_loss
def l1_loss(pred, target):
    return torch.abs(pred - target)


@weighted_loss
def smooth_l1_loss(pred, target, beta=1.0):
    diff = torch.abs(pred - target)
    return torch.where(diff < beta, 0.5 * diff * diff / beta, diff - 0.5 * beta)


class _Loss(nn.Module):
    """Placeholder for mmdet loss classes: constructible from the config, never called by the harness
    (losses / targets are out of scope)."""
    def __init__(self, **kwargs):
        super().__init__()
        self.cfg = kwargs
        self.loss_weight = kwargs.get("loss_weight", 1.0)

    def forward(self, *a, **k):
        raise NotImplementedError("mmdet losses are not part of the decoder harness")


class _Plain:
    def __init__(self, *a, **k):
        pass


def install():
    """Register the stand-in modules in sys.modules (idempotent) and return the registries."""
    if "mmcv" in sys.modules and getattr(sys.modules["mmcv"], "_hipad_harness", False):
        return sys.modules["mmcv"]._registries

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    names = ("ATTENTION", "PLUGIN_LAYERS", "POSITIONAL_ENCODING", "FEEDFORWARD_NETWORK", "NORM_LAYERS",
             "BBOX_SAMPLERS", "BBOX_CODERS", "BBOX_ASSIGNERS", "MATCH_COST", "HEADS", "LOSSES", "DETECTORS",
             "BACKBONES", "NECKS")
    regs = {n: Registry(n) for n in names}
    regs["NORM_LAYERS"].register_module("LN", module=nn.LayerNorm)
    for loss in ("FocalLoss", "L1Loss", "CrossEntropyLoss", "GaussianFocalLoss", "SmoothL1Loss"):
        regs["LOSSES"].register_module(loss, module=type(loss, (_Loss,), {}))
    regs["MATCH_COST"].register_module("FocalLossCost", module=type("FocalLossCost", (_Plain,), {}))

    class FFN(BaseModule):      # imported by blocks.py, never instantiated by the shipped configs
        def __init__(self, *a, **k):
            raise NotImplementedError("mmcv FFN is not used by the HiP-AD configs (AsymmetricFFN is)")

    mmcv = mod("mmcv", jit=_passthrough_decorator, _hipad_harness=True, _registries=regs)
    mod("mmcv.cnn", Linear=nn.Linear, Scale=Scale, bias_init_with_prob=bias_init_with_prob, xavier_init=xavier_init,
        constant_init=constant_init, build_norm_layer=build_norm_layer, build_activation_layer=build_activation_layer)
    mod("mmcv.cnn.bricks")
    mod("mmcv.cnn.bricks.registry", **{k: regs[k] for k in ("ATTENTION", "PLUGIN_LAYERS", "POSITIONAL_ENCODING",
                                                              "FEEDFORWARD_NETWORK", "NORM_LAYERS")})
    mod("mmcv.cnn.bricks.transformer", FFN=FFN)
    mod("mmcv.cnn.bricks.drop", build_dropout=build_dropout)
    mod("mmcv.runner", BaseModule=BaseModule, force_fp32=force_fp32, auto_fp16=auto_fp16)
    mod("mmcv.runner.base_module", BaseModule=BaseModule, Sequential=Sequential)
    mod("mmcv.utils", build_from_cfg=build_from_cfg, deprecated_api_warning=_passthrough_decorator, Registry=Registry)

    def builder(reg):
        return lambda cfg, **kw: build_from_cfg(cfg, regs[reg], kw or None)

    class BaseDetector(BaseModule):
        pass

    class AssignResult(_Plain):
        pass

    class BaseAssigner(_Plain):
        pass

    mod("mmdet")
    mod("mmdet.core", reduce_mean=reduce_mean, build_assigner=builder("BBOX_ASSIGNERS"),
        build_sampler=builder("BBOX_SAMPLERS"))
    mod("mmdet.core.bbox")
    mod("mmdet.core.bbox.builder", BBOX_SAMPLERS=regs["BBOX_SAMPLERS"], BBOX_CODERS=regs["BBOX_CODERS"],
        BBOX_ASSIGNERS=regs["BBOX_ASSIGNERS"])
    mod("mmdet.core.bbox.assigners", AssignResult=AssignResult, BaseAssigner=BaseAssigner)
    mod("mmdet.core.bbox.match_costs", build_match_cost=builder("MATCH_COST"))
    mod("mmdet.core.bbox.match_costs.builder", MATCH_COST=regs["MATCH_COST"])
    mod("mmdet.models", HEADS=regs["HEADS"], LOSSES=regs["LOSSES"], DETECTORS=regs["DETECTORS"],
        BaseDetector=BaseDetector, build_backbone=builder("BACKBONES"), build_neck=builder("NECKS"),
        build_head=builder("HEADS"), weighted_loss=weighted_loss)
    mod("mmdet.models.builder", LOSSES=regs["LOSSES"], HEADS=regs["HEADS"], DETECTORS=regs["DETECTORS"])
    mod("mmdet.models.losses", l1_loss=l1_loss, smooth_l1_loss=smooth_l1_loss)
    return regs
