"""Evaluator metrics, CPU part: the oracle's restatement of lib/pysixd/pose_error.py (add / adi / re / te) against the
outputs of the UNMODIFIED reference functions stored in tests/golden/pose_errors_b6.npz (oracle/make_golden_pose_errors.py).  The
GPU test (tests/test_ops_gpu.py::test_pose_errors_match_pysixd_restatement) compares the kernel with this oracle."""
import os

import numpy as np

from oracle import gdrn_oracle as O


def test_pose_error_oracle_matches_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "pose_errors_b6.npz"))
    for i in range(g["pts"].shape[0]):
        args = (g["R_est"][i], g["t_est"][i], g["R_gt"][i], g["t_gt"][i], g["pts"][i])
        assert abs(O.add_metric(*args) - g["add"][i]) <= 1e-6 * g["add"][i]
        assert abs(O.adi_metric(*args) - g["adi"][i]) <= 1e-6 * g["adi"][i]
        assert abs(O._re_deg(g["R_est"][i], g["R_gt"][i]) - g["re"][i]) < 1e-3  # float32 trace near 180 degrees
        assert abs(float(np.linalg.norm(g["t_gt"][i] - g["t_est"][i])) - g["te"][i]) < 1e-7
