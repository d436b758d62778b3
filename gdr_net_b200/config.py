"""mmcv-style python config loading (the reference passes an `mmcv.Config`, see
reference `core/gdrn_modeling/main_gdrn.py:37-41`).  mmcv is not part of this
package's dependencies, so this is a small independent implementation of the
semantics the hot path relies on:

  * a config file is a python file whose top-level names are the config keys,
  * `_base_ = "rel/path.py"` (or a list) is loaded first and dict-merged,
  * a dict carrying `_delete_=True` replaces the inherited dict instead of merging,
  * attribute access (`cfg.MODEL.CDPN.NAME`), `.get`, `.pop`, item access.

`a6_config()` builds the one network shape used by all 41 shipped experiment
configs (reference `configs/_base_/gdrn_base.py:5-143` +
`configs/gdrn/lm/a6_cPnP_lm13.py:41-66`) without needing the reference tree.
"""
from __future__ import annotations

import copy
import os
from typing import Any

__all__ = ["ConfigDict", "Config", "a6_config"]


class ConfigDict(dict):
    """dict with attribute access, nested dicts converted recursively."""

    def __init__(self, *args, **kwargs):
        super().__init__()
        for k, v in dict(*args, **kwargs).items():
            self[k] = v

    @staticmethod
    def _wrap(v):
        if isinstance(v, dict) and not isinstance(v, ConfigDict):
            return ConfigDict(v)
        if isinstance(v, (list, tuple)):
            return type(v)(ConfigDict._wrap(x) for x in v)
        return v

    def __setitem__(self, k, v):
        super().__setitem__(k, ConfigDict._wrap(v))

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def __delattr__(self, k):
        del self[k]

    def update(self, *args, **kwargs):
        for k, v in dict(*args, **kwargs).items():
            self[k] = v

    def __deepcopy__(self, memo):
        return ConfigDict({k: copy.deepcopy(v, memo) for k, v in self.items()})

    def to_dict(self):
        out = {}
        for k, v in self.items():
            out[k] = v.to_dict() if isinstance(v, ConfigDict) else v
        return out


def _merge(base: dict, child: dict) -> dict:
    out = copy.deepcopy(base)
    for k, v in child.items():
        if isinstance(v, dict):
            v = dict(v)
            delete = v.pop("_delete_", False)
            if not delete and isinstance(out.get(k), dict):
                out[k] = _merge(out[k], v)
            else:
                out[k] = _merge({}, v)
        else:
            out[k] = copy.deepcopy(v)
    return out


def _load_file(path: str) -> dict:
    path = os.path.abspath(path)
    ns: dict[str, Any] = {"__file__": path}
    with open(path, "r") as f:
        code = compile(f.read(), path, "exec")
    exec(code, ns)  # config files are python by contract
    cfg = {k: v for k, v in ns.items() if not k.startswith("__") and not callable(v) and not _is_module(v)}
    bases = cfg.pop("_base_", None)
    if bases is None:
        return cfg
    if isinstance(bases, str):
        bases = [bases]
    merged: dict = {}
    for b in bases:
        merged = _merge(merged, _load_file(os.path.join(os.path.dirname(path), b)))
    return _merge(merged, cfg)


def _is_module(v):
    import types

    return isinstance(v, types.ModuleType)


class Config(ConfigDict):
    @staticmethod
    def fromfile(path: str) -> "Config":
        return Config(_load_file(path))

    def merge_from_dict(self, options: dict):
        """`--opts A.B.C=v` style overrides (reference `default_args_setup.py:65-67`)."""
        for key, v in options.items():
            d = self
            parts = key.split(".")
            for p in parts[:-1]:
                d = d.setdefault(p, ConfigDict())
            d[parts[-1]] = v


def postprocess_like_main_gdrn(cfg: ConfigDict, device: str = "cuda") -> ConfigDict:
    """The part of reference `main_gdrn.py:setup` (:63-73) the model builder depends on."""
    opt = cfg.SOLVER.get("OPTIMIZER_CFG", "")
    if isinstance(opt, str) and opt != "":
        opt = eval(opt)  # reference does the same for string-typed optimizer cfgs
        cfg.SOLVER.OPTIMIZER_CFG = opt
    if opt != "":
        if "lr" in opt:
            cfg.SOLVER.BASE_LR = opt["lr"]
        if "weight_decay" in opt:
            cfg.SOLVER.WEIGHT_DECAY = opt["weight_decay"]
    cfg.MODEL.DEVICE = device
    return cfg


def a6_config(num_regions: int = 64, pm_loss_sym: bool = False, device: str = "cuda", use_pnp_test: bool = False,
              optimizer: dict | None = None, with_2d_coord: bool = True) -> Config:
    """The `a6_cPnP` network/loss configuration (keys the hot path reads)."""
    cfg = Config(
        MODEL=dict(
            DEVICE=device,
            WEIGHTS="",
            CDPN=dict(
                NAME="GDRN",
                TASK="rot",
                USE_MTL=False,
                BACKBONE=dict(PRETRAINED="", ARCH="resnet", NUM_LAYERS=34, INPUT_CHANNEL=3, INPUT_RES=256,
                              OUTPUT_RES=64, FREEZE=False),
                ROT_HEAD=dict(
                    FREEZE=False, ROT_CONCAT=False, XYZ_BIN=64, NUM_LAYERS=3, NUM_FILTERS=256, CONV_KERNEL_SIZE=3,
                    NORM="BN", NUM_GN_GROUPS=32, OUT_CONV_KERNEL_SIZE=1, NUM_CLASSES=13, ROT_CLASS_AWARE=False,
                    XYZ_LOSS_TYPE="L1", XYZ_LOSS_MASK_GT="visib", XYZ_LW=1.0, MASK_CLASS_AWARE=False,
                    MASK_LOSS_TYPE="L1", MASK_LOSS_GT="trunc", MASK_LW=1.0, MASK_THR_TEST=0.5,
                    NUM_REGIONS=num_regions, REGION_CLASS_AWARE=False, REGION_LOSS_TYPE="CE",
                    REGION_LOSS_MASK_GT="visib", REGION_LW=1.0,
                ),
                PNP_NET=dict(
                    FREEZE=False, R_ONLY=False, LR_MULT=1.0,
                    PNP_HEAD_CFG=dict(type="ConvPnPNet", norm="GN", num_gn_groups=32, drop_prob=0.0),
                    WITH_2D_COORD=with_2d_coord, REGION_ATTENTION=True, MASK_ATTENTION="none", TRANS_WITH_BOX_INFO="none",
                    ROT_TYPE="allo_rot6d", TRANS_TYPE="centroid_z", Z_TYPE="REL",
                    NUM_PM_POINTS=3000, PM_LOSS_TYPE="L1", PM_SMOOTH_L1_BETA=1.0, PM_LOSS_SYM=pm_loss_sym,
                    PM_NORM_BY_EXTENT=True, PM_R_ONLY=True, PM_DISENTANGLE_T=False, PM_DISENTANGLE_Z=False,
                    PM_T_USE_POINTS=False, PM_LW=1.0, ROT_LOSS_TYPE="angular", ROT_LW=0.0,
                    CENTROID_LOSS_TYPE="L1", CENTROID_LW=1.0, Z_LOSS_TYPE="L1", Z_LW=1.0,
                    TRANS_LOSS_TYPE="L1", TRANS_LOSS_DISENTANGLE=True, TRANS_LW=0.0,
                    BIND_LOSS_TYPE="L1", BIND_LW=0.0,
                ),
                TRANS_HEAD=dict(ENABLED=False, FREEZE=True, LR_MULT=1.0, NUM_LAYERS=3, NUM_FILTERS=256, NORM="BN",
                                NUM_GN_GROUPS=32, CONV_KERNEL_SIZE=3, OUT_CHANNEL=3, TRANS_TYPE="centroid_z",
                                Z_TYPE="REL", CENTROID_LOSS_TYPE="L1", CENTROID_LW=0.0, Z_LOSS_TYPE="L1", Z_LW=0.0,
                                TRANS_LOSS_TYPE="L1", TRANS_LW=0.0),
            ),
        ),
        SOLVER=dict(
            IMS_PER_BATCH=24, BASE_LR=1e-4, WEIGHT_DECAY=0.0,
            OPTIMIZER_CFG=optimizer if optimizer is not None else dict(type="Ranger", lr=1e-4, weight_decay=0),
            AMP=dict(ENABLED=False),
        ),
        TEST=dict(USE_PNP=use_pnp_test, AMP_TEST=False),
    )
    return cfg
