"""Train-step harness without Lightning / detectron2: the pieces of the reference's `do_train` loop
(core/gdrn_modeling/engine.py:228-283) that touch the hot path, with the host synchronisations removed.

  * `batch_data(cfg, data, device, phase)` -- the reference's collate (core/gdrn_modeling/engine_utils.py:6-60): a list of
    per-sample dicts -> device tensors under the same keys (`roi_img`, `roi_cls`, `roi_cam`, `roi_center`, `roi_wh`,
    `resize_ratio`, `roi_extent`, `roi_trans_ratio`, `roi_xyz`, masks, `roi_region` (long), `ego_rot`, `trans`, `roi_points`,
    `sym_info` list).  One pinned staging buffer per key and non-blocking copies instead of per-sample `.to(device)`.
  * `TrainStep(model, optimizer, world)` -- forward, `sum(loss_dict.values())`, backward, optimizer step:
      - `model.use_cuda_graphs`: forward and backward replay CUDA graphs (incl. the data-parallel all-reduces),
      - the fused Ranger step is one launch per param group,
      - the loss is NOT `.item()`-ed per iteration (the reference does `{k: v.item() ...}` 8 times + 18 `vis/*` scalars per
        step): `losses_async()` returns the values of the PREVIOUS step from a pinned buffer (one async D2H copy per step,
        consumed one step later); `model.last_vis_dict` fetches the logging scalars lazily.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

_FLOAT_KEYS = ["roi_xyz", "roi_xyz_bin", "roi_mask_trunc", "roi_mask_visib", "roi_mask_obj", "ego_quat", "allo_quat", "ego_rot6d",
               "allo_rot6d", "ego_rot", "trans", "roi_points"]


def _stack(vals, dtype=None):
    t = torch.stack([torch.as_tensor(v) for v in vals], dim=0)
    return t if dtype is None else t.to(dtype)


def batch_data(cfg, data: List[dict], device="cuda", phase: str = "train") -> dict:
    """reference engine_utils.py:6-60 (train) / :63-110 (test: the subset of keys the model's inference call reads)."""
    dev = torch.device(device)
    host: Dict[str, torch.Tensor] = {}
    host["roi_img"] = _stack([d["roi_img"] for d in data], torch.float32)
    host["roi_cls"] = torch.tensor([int(d["roi_cls"]) for d in data], dtype=torch.long)
    if "roi_coord_2d" in data[0]:
        host["roi_coord_2d"] = _stack([d["roi_coord_2d"] for d in data], torch.float32)
    host["roi_cam"] = _stack([d["cam"] for d in data], torch.float32)
    host["roi_center"] = _stack([d["bbox_center"] for d in data], torch.float32)
    host["roi_wh"] = _stack([d["roi_wh"] for d in data], torch.float32)
    host["resize_ratio"] = torch.tensor([float(d["resize_ratio"]) for d in data], dtype=torch.float32)
    if "roi_extent" in data[0]:
        host["roi_extent"] = _stack([d["roi_extent"] for d in data], torch.float32)
    if phase == "train":
        host["roi_trans_ratio"] = _stack([d["trans_ratio"] for d in data], torch.float32)
        for key in _FLOAT_KEYS:
            if key in data[0]:
                host[key] = _stack([d[key] for d in data], torch.float32)
        if "roi_region" in data[0]:
            host["roi_region"] = _stack([d["roi_region"] for d in data], torch.long)
    batch = {}
    for k, t in host.items():
        if dev.type == "cuda":
            t = t.pin_memory()
        batch[k] = t.to(dev, non_blocking=True)
    if phase == "train" and "sym_info" in data[0]:
        batch["sym_info"] = [d["sym_info"] for d in data]
    return batch


def forward_kwargs(batch: dict, train: bool = True) -> dict:
    """The keyword arguments of `model(batch["roi_img"], ...)` exactly as the reference passes them (engine.py:244-269,
    gdrn_evaluator.py:569-578)."""
    kw = dict(roi_classes=batch["roi_cls"], roi_cams=batch["roi_cam"], roi_whs=batch["roi_wh"], roi_centers=batch["roi_center"],
              resize_ratios=batch["resize_ratio"], roi_coord_2d=batch.get("roi_coord_2d"), roi_extents=batch.get("roi_extent"))
    if train:
        kw.update(gt_xyz=batch.get("roi_xyz"), gt_xyz_bin=batch.get("roi_xyz_bin"), gt_mask_trunc=batch["roi_mask_trunc"],
                  gt_mask_visib=batch["roi_mask_visib"], gt_mask_obj=batch.get("roi_mask_obj"), gt_region=batch.get("roi_region"),
                  gt_allo_quat=batch.get("allo_quat"), gt_ego_quat=batch.get("ego_quat"), gt_allo_rot6d=batch.get("allo_rot6d"),
                  gt_ego_rot6d=batch.get("ego_rot6d"), gt_ego_rot=batch.get("ego_rot"), gt_trans=batch.get("trans"),
                  gt_trans_ratio=batch["roi_trans_ratio"], gt_points=batch.get("roi_points"), sym_infos=batch.get("sym_info"),
                  do_loss=True)
    return kw


class TrainStep:
    """One optimisation step of the reference loop (engine.py:240-278) without per-iteration host synchronisation."""

    def __init__(self, model, optimizer, scheduler=None, use_cuda_graphs: bool = True):
        self.model, self.optimizer, self.scheduler = model, optimizer, scheduler
        model.use_cuda_graphs = use_cuda_graphs
        self._names: Optional[List[str]] = None
        self._host = [None, None]
        self._events = [None, None]
        self._i = 0

    def __call__(self, batch: dict) -> dict:
        out_dict, loss_dict = self.model(batch["roi_img"], **forward_kwargs(batch, train=True))
        losses = sum(loss_dict.values())
        self.optimizer.zero_grad(set_to_none=True)
        losses.backward()
        self.optimizer.step()
        if self.scheduler is not None:
            self.scheduler.step()
        # queue ONE async copy of this step's 8 loss values; they are read back one step later (no stall)
        if self._names is None:
            self._names = sorted(loss_dict)
            self._host = [torch.empty(len(self._names) + 1, dtype=torch.float32).pin_memory() for _ in range(2)]
            self._events = [torch.cuda.Event(), torch.cuda.Event()]
        slot = self._i & 1
        vals = torch.stack([loss_dict[k].detach() for k in self._names] + [losses.detach()])
        self._host[slot].copy_(vals, non_blocking=True)
        self._events[slot].record()
        self._i += 1
        return loss_dict

    def losses_async(self, wait_current: bool = False) -> Optional[dict]:
        """Loss values of the previous step (or, with wait_current, of the step just issued)."""
        if self._i == 0 or (self._i == 1 and not wait_current):
            return None
        slot = (self._i - 1) & 1 if wait_current else (self._i - 2) & 1
        self._events[slot].synchronize()
        v = self._host[slot].tolist()
        d = dict(zip(self._names, v[:-1]))
        d["total_loss"] = v[-1]
        return d
