"""TEST INFRASTRUCTURE ONLY -- golden vectors for the evaluator's pose-error metrics, produced by the UNMODIFIED
`lib.pysixd.pose_error.{add, adi, re, te}` (pose_error.py:297-337, 400-436) imported from /root/reference, on seeded poses and
point clouds.  Output: tests/golden/pose_errors_b6.npz.  Usage: python -m oracle.make_golden_pose_errors"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from gdr_net_b200 import synth  # noqa: E402
from oracle import ref_shim  # noqa: E402


def make_inputs(B=6, n=1500):
    g = synth._gen(9, "pose_errors_golden")
    pts = (torch.rand(B, n, 3, generator=g) - 0.5) * 0.2
    R_gt = synth.random_rotations(B, g)
    R_est = torch.linalg.qr(R_gt + 0.05 * torch.randn(B, 3, 3, generator=g))[0]
    R_est = R_est * torch.sign(torch.linalg.det(R_est))[:, None, None]
    t_gt = torch.stack([torch.rand(B, generator=g) * 0.4 - 0.2, torch.rand(B, generator=g) * 0.4 - 0.2, 0.4 + torch.rand(B, generator=g)], 1)
    t_est = t_gt + 0.01 * torch.randn(B, 3, generator=g)
    return dict(pts=pts.numpy(), R_gt=R_gt.numpy(), R_est=R_est.numpy(), t_gt=t_gt.numpy(), t_est=t_est.numpy())


def main():
    ref_shim.install()
    import lib.pysixd.pose_error as pe

    d = make_inputs()
    B = d["pts"].shape[0]
    out = dict(d)
    out["add"] = np.array([pe.add(d["R_est"][i], d["t_est"][i].reshape(3, 1), d["R_gt"][i], d["t_gt"][i].reshape(3, 1), d["pts"][i]) for i in range(B)])
    out["adi"] = np.array([pe.adi(d["R_est"][i], d["t_est"][i].reshape(3, 1), d["R_gt"][i], d["t_gt"][i].reshape(3, 1), d["pts"][i]) for i in range(B)])
    out["re"] = np.array([pe.re(d["R_est"][i], d["R_gt"][i]) for i in range(B)])
    out["te"] = np.array([pe.te(d["t_est"][i], d["t_gt"][i]) for i in range(B)])
    print(out["add"], out["adi"], out["re"], out["te"])
    path = os.path.join(ROOT, "tests", "golden", "pose_errors_b6.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
