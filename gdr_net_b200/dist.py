"""Data-parallel gradient exchange (reference: plain DDP through LightningLite, main_gdrn.py:113,134-142).

The path shards over crops only (per-GPU BatchNorm statistics): one exchange per step -- the all-reduce
(mean) of 35.05 M fp32 gradients.  The engine writes every weight gradient straight into ONE flat fp32 buffer in
parameter order, so a "bucket" is just a contiguous slice.  `GradAllReducer` is called by the engine as soon as a
sub-network's gradients are final (Patch-PnP -> head -> layer4 ... layer1/stem, i.e. reverse forward order) and
launches NCCL on a side stream, overlapping the exchange with the remaining backward kernels.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch
import torch.distributed as dist


def segment_bounds(named_params, boundaries: List[str]) -> Dict[str, Tuple[int, int]]:
    """Contiguous [start, end) element ranges of the flat gradient buffer for each name prefix in `boundaries`
    (a parameter belongs to the first prefix it matches)."""
    out: Dict[str, List[int]] = {}
    off = 0
    for name, p in named_params:
        n = p.numel()
        for b in boundaries:
            if name.startswith(b):
                if b not in out:
                    out[b] = [off, off + n]
                else:
                    assert out[b][1] == off, f"segment {b} is not contiguous at {name}"
                    out[b][1] = off + n
                break
        else:
            raise KeyError(f"parameter {name} matches no segment prefix")
        off += n
    return {k: (v[0], v[1]) for k, v in out.items()}


SEGMENTS = ["backbone.conv1", "backbone.bn1", "backbone.layer1", "backbone.layer2", "backbone.layer3", "backbone.layer4",
            "rot_head_net", "pnp_net"]


class GradAllReducer:
    """Bucketed, overlapped all-reduce(mean) over a flat gradient buffer."""

    def __init__(self, flat_grad: torch.Tensor, named_params, process_group=None, use_side_stream: bool = True):
        self.flat = flat_grad
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.bounds = segment_bounds(named_params, SEGMENTS)
        self.cuda = flat_grad.is_cuda
        self.stream = torch.cuda.Stream() if (self.cuda and use_side_stream) else None
        self.pending = []
        self.bytes_reduced = 0

    # engine.grad_hook signature
    def __call__(self, engine, stage: str):
        if stage == "backbone.stem":
            for seg in ("backbone.layer1", "backbone.bn1", "backbone.conv1"):
                self.reduce_segment(seg)
        else:
            self.reduce_segment(stage)

    def reduce_segment(self, seg: str):
        if self.world == 1:
            return
        lo, hi = self.bounds[seg]
        chunk = self.flat[lo:hi]
        self.bytes_reduced += chunk.numel() * chunk.element_size()
        if self.cuda:
            # NCCL averages inside the collective (ReduceOp.AVG): no separate scaling launch per bucket
            if self.stream is not None:
                self.stream.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(self.stream):
                    w = dist.all_reduce(chunk, op=dist.ReduceOp.AVG, group=self.pg, async_op=True)
            else:
                w = dist.all_reduce(chunk, op=dist.ReduceOp.AVG, group=self.pg, async_op=True)
        else:  # gloo (CPU tests) has no AVG
            chunk.div_(self.world)
            w = dist.all_reduce(chunk, op=dist.ReduceOp.SUM, group=self.pg, async_op=True)
        self.pending.append(w)

    def finish(self):
        """Make the reduced gradients visible to the compute stream (call before the optimizer step)."""
        for w in self.pending:
            w.wait()
        self.pending.clear()
        if self.stream is not None:
            torch.cuda.current_stream().wait_stream(self.stream)
