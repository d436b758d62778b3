"""TEST INFRASTRUCTURE ONLY -- import shims that let the UNMODIFIED reference
(`/root/reference`, read-only) be imported and executed on CPU in the build
container, so that `oracle/make_golden.py` can pin the oracle restatement
(`oracle/gdrn_oracle.py`) against the live reference.

Nothing in the product path (`gdr_net_b200/`) may import this module.
`/root/reference` does not exist on the GPU box; only `make_golden.py` (run here,
outputs committed under `tests/golden/`) uses this file.

The reference hot path needs detectron2 / mmcv / fvcore / transforms3d /
termcolor ... none of which are installed (no network).  We install tiny stub
modules that provide exactly the names the reference imports on this path
(SURVEY.md section 8c / Appendix B).  No reference source is copied.
"""
from __future__ import annotations

import importlib
import importlib.abc
import importlib.machinery
import math
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

REFERENCE_ROOT = os.environ.get("GDRN_REFERENCE_ROOT", "/root/reference")


def _mod(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__path__ = []  # behave like a package so sub-imports resolve
    sys.modules[name] = m
    parent, _, leaf = name.rpartition(".")
    if parent and parent in sys.modules:
        setattr(sys.modules[parent], leaf, m)
    return m


class _Registry(dict):
    """Minimal stand-in for mmcv's OPTIMIZERS registry."""

    def register_module(self, name=None, force=False, module=None):
        def deco(cls):
            self[name or cls.__name__] = cls
            return cls

        if module is not None:
            return deco(module)
        return deco

    def get(self, key, default=None):
        return dict.get(self, key, default)


def _axangle2mat(axis, angle, is_normalized=False):
    """Rodrigues rotation (public formula; transforms3d.axangles.axangle2mat)."""
    x, y, z = [float(v) for v in axis]
    if not is_normalized:
        n = math.sqrt(x * x + y * y + z * z)
        x, y, z = x / n, y / n, z / n
    c, s = math.cos(angle), math.sin(angle)
    C = 1 - c
    xs, ys, zs = x * s, y * s, z * s
    xC, yC, zC = x * C, y * C, z * C
    xyC, yzC, zxC = x * yC, y * zC, z * xC
    return np.array(
        [[x * xC + c, xyC - zs, zxC + ys], [xyC + zs, y * yC + c, yzC - xs], [zxC - ys, yzC + xs, z * zC + c]]
    )


class _EventStorage:
    """detectron2.utils.events.EventStorage stand-in: records the last scalars."""

    def __init__(self):
        self.scalars = {}

    def put_scalar(self, k, v, **kw):
        self.scalars[k] = v

    def put_scalars(self, **kw):
        self.scalars.update(kw)


_STORAGE = _EventStorage()


class _LeafFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    """Returns empty placeholder modules for unused third-party leaf imports."""

    PREFIXES = (
        "imageio", "png", "chardet", "plyfile", "pycocotools", "matplotlib", "pyassimp", "OpenGL", "glumpy",
        "vispy", "imgaug", "termcolor", "six", "transforms3d", "fvcore", "detectron2", "mmcv", "scipy.misc",
        "tensorboardX", "pytorch_lightning", "setproctitle", "ruamel", "pyrender", "trimesh", "open3d", "skimage",
        "fairscale", "horovod", "apex", "timm", "egl", "pypng", "yaml_include", "ujson", "simplejson",
    )

    def find_spec(self, fullname, path, target=None):
        if fullname in sys.modules:
            return None
        if any(fullname == p or fullname.startswith(p + ".") for p in self.PREFIXES):
            try:  # prefer a real module when one exists
                for f in sys.meta_path:
                    if f is self:
                        continue
                    spec = f.find_spec(fullname, path, target) if hasattr(f, "find_spec") else None
                    if spec is not None:
                        return None
            except Exception:
                pass
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = types.ModuleType(spec.name)
        m.__path__ = []

        def _getattr(name, _n=spec.name):
            if name.startswith("__"):
                raise AttributeError(name)
            return _Placeholder(f"{_n}.{name}")

        m.__getattr__ = _getattr
        return m

    def exec_module(self, module):
        pass


class _Placeholder:
    def __init__(self, name):
        self._name = name

    def __call__(self, *a, **k):
        # used as decorator (e.g. numba-like) or constructor: return first callable arg or self
        if len(a) == 1 and callable(a[0]) and not k:
            return a[0]
        return self

    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        return _Placeholder(f"{self._name}.{k}")

    def __mro_entries__(self, bases):
        return (object,)


_INSTALLED = False


def install():
    """Install stubs + put the reference on sys.path.  Idempotent."""
    global _INSTALLED
    if _INSTALLED:
        return
    if not os.path.isdir(REFERENCE_ROOT):
        raise RuntimeError(f"reference not found at {REFERENCE_ROOT} (only available in the build container)")

    # numpy-2 compat; do NOT touch np.bool
    for k, v in (("float", float), ("int", int)):
        if not hasattr(np, k):
            setattr(np, k, v)
    if not hasattr(np, "maximum_sctype"):
        np.maximum_sctype = lambda t: np.float64

    # ---- mmcv -----------------------------------------------------------------
    def normal_init(module, mean=0, std=1, bias=0):
        if hasattr(module, "weight") and module.weight is not None:
            nn.init.normal_(module.weight, mean, std)
        if hasattr(module, "bias") and module.bias is not None:
            nn.init.constant_(module.bias, bias)

    def constant_init(module, val, bias=0):
        if hasattr(module, "weight") and module.weight is not None:
            nn.init.constant_(module.weight, val)
        if hasattr(module, "bias") and module.bias is not None:
            nn.init.constant_(module.bias, bias)

    def kaiming_init(module, a=0, mode="fan_out", nonlinearity="relu", bias=0, distribution="normal"):
        nn.init.kaiming_normal_(module.weight, a=a, mode=mode, nonlinearity=nonlinearity)
        if hasattr(module, "bias") and module.bias is not None:
            nn.init.constant_(module.bias, bias)

    optimizers = _Registry()
    for _n in ("SGD", "Adam", "AdamW", "RMSprop"):
        optimizers[_n] = getattr(torch.optim, _n)

    def build_from_cfg(cfg, registry, default_args=None):
        args = dict(cfg)
        if default_args:
            for k, v in default_args.items():
                args.setdefault(k, v)
        typ = args.pop("type")
        cls = registry[typ] if isinstance(typ, str) else typ
        return cls(**args)

    _noop = lambda *a, **k: None
    _mod("mmcv", Config=object, mkdir_or_exist=lambda p: os.makedirs(p, exist_ok=True), load=_noop, dump=_noop)
    _mod("mmcv.cnn", normal_init=normal_init, constant_init=constant_init, kaiming_init=kaiming_init)
    _mod("mmcv.runner", load_checkpoint=_noop, _load_checkpoint=_noop, obj_from_dict=_noop,
         load_state_dict=_noop)
    _mod("mmcv.runner.optimizer", OPTIMIZERS=optimizers, DefaultOptimizerConstructor=object,
         build_optimizer=_noop)
    _mod("mmcv.runner.checkpoint", _load_checkpoint=_noop, load_state_dict=_noop)
    _mod("mmcv.utils", build_from_cfg=build_from_cfg, Registry=_Registry)

    # ---- detectron2 -------------------------------------------------------------
    class FrozenBatchNorm2d(nn.BatchNorm2d):
        pass

    _mod("detectron2")
    _mod("detectron2.utils")
    _mod("detectron2.utils.events", get_event_storage=lambda: _STORAGE, EventStorage=_EventStorage)
    _mod("detectron2.layers", cat=lambda ts, dim=0: torch.cat(ts, dim=dim))
    _mod("detectron2.layers.batch_norm", BatchNorm2d=nn.BatchNorm2d, FrozenBatchNorm2d=FrozenBatchNorm2d,
         NaiveSyncBatchNorm=nn.BatchNorm2d)
    _mod("detectron2.utils.comm", get_world_size=lambda: 1, get_rank=lambda: 0, is_main_process=lambda: True,
         synchronize=_noop)
    _mod("detectron2.utils.env", TORCH_VERSION=(2, 11))
    _mod("detectron2.utils.logger", log_first_n=_noop, setup_logger=_noop)
    _mod("detectron2.config", CfgNode=dict)
    _mod("detectron2.solver", WarmupCosineLR=object, WarmupMultiStepLR=object)

    # ---- misc -------------------------------------------------------------------
    def smooth_l1_loss(input, target, beta, reduction="none"):
        return torch.nn.functional.smooth_l1_loss(input, target, beta=beta, reduction=reduction)

    _mod("fvcore")
    _mod("fvcore.nn", smooth_l1_loss=smooth_l1_loss)
    _mod("termcolor", colored=lambda s, *a, **k: s)
    _mod("transforms3d")
    _ph = _Placeholder
    _mod("transforms3d.axangles", axangle2mat=_axangle2mat, mat2axangle=_ph("mat2axangle"))
    _mod("transforms3d.quaternions", quat2mat=_ph("quat2mat"), mat2quat=_ph("mat2quat"),
         axangle2quat=_ph("axangle2quat"), qmult=_ph("qmult"), quat2axangle=_ph("quat2axangle"),
         qinverse=_ph("qinverse"))
    _mod("transforms3d.euler", _AXES2TUPLE={}, _NEXT_AXIS=[1, 2, 0, 1], _TUPLE2AXES={}, euler2mat=_ph("euler2mat"),
         euler2quat=_ph("euler2quat"), mat2euler=_ph("mat2euler"), quat2euler=_ph("quat2euler"))

    sys.meta_path.append(_LeafFinder())
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _INSTALLED = True


def event_storage() -> _EventStorage:
    return _STORAGE


def import_reference_gdrn():
    """Returns the reference module `core.gdrn_modeling.models.GDRN` (unmodified)."""
    install()
    return importlib.import_module("core.gdrn_modeling.models.GDRN")
