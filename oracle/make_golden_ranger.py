"""TEST INFRASTRUCTURE ONLY -- golden vectors for the optimizer step, produced by the UNMODIFIED reference
`lib.torch_utils.solver.ranger.Ranger` (ranger.py:100-200: gradient centralisation + RAdam + Lookahead, k = 6) imported from
/root/reference: 14 steps (two Lookahead syncs, the RAdam rectification switching on) on seeded parameters / gradients, with and
without weight decay.  Output: tests/golden/ranger_14steps.npz.  Usage: python -m oracle.make_golden_ranger"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402

SHAPES = [(8, 4, 3, 3), (16, 8), (5,)]
STEPS = 14


def make_inputs(seed=0):
    g = torch.Generator().manual_seed(seed)
    params = [torch.randn(*s, generator=g) for s in SHAPES]
    grads = [[torch.randn(*s, generator=g) for s in SHAPES] for _ in range(STEPS)]
    return params, grads


def main():
    ref_shim.install()
    from lib.torch_utils.solver.ranger import Ranger

    params, grads = make_inputs()
    out = {}
    for i, p in enumerate(params):
        out[f"p0_{i}"] = p.numpy()
    for t in range(STEPS):
        for i, gr in enumerate(grads[t]):
            out[f"g{t}_{i}"] = gr.numpy()
    for tag, wd in (("wd0", 0.0), ("wd1e-2", 1e-2)):
        ps = [p.clone().requires_grad_(True) for p in params]
        opt = Ranger(ps, lr=1e-2, weight_decay=wd)
        for t in range(STEPS):
            for p, gr in zip(ps, grads[t]):
                p.grad = gr.clone()
            opt.step()
        for i, p in enumerate(ps):
            out[f"{tag}_p{i}"] = p.detach().numpy()
    path = os.path.join(ROOT, "tests", "golden", "ranger_14steps.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
