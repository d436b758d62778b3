"""GPU crop / target generation, CPU part: gdr_net_b200/csrc/roi.cu rebuilt for the host (one thread per CTA,
tests/emu/roi_emu.cpp) and compared with the reference's own procedure run with cv2 / scipy (oracle/roi_oracle.py restates
data_loader.py:487-560 with the real cv2.warpAffine) -- the same assertions as tests/test_ops_gpu.py::test_roi_targets_match_reference_procedure."""
import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")

from oracle import roi_oracle as RO  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    if shutil.which("g++") is None or not os.path.exists("/usr/local/cuda/include/cuda_runtime.h"):
        pytest.skip("g++ / CUDA headers needed for the host emulation build")
    # -Bsymbolic: other CPU tests load the real libgdrn_b200.so with RTLD_GLOBAL; without it this library's calls to its own
    # (host-compiled) kernels would bind to the real library's CUDA launch stubs of the same name
    out = tmp_path_factory.mktemp("emu") / "libroi_emu.so"
    subprocess.check_call(["g++", "-O1", "-ffp-contract=off", "-shared", "-fPIC", "-Wl,-Bsymbolic", "-I/usr/local/cuda/include",
                           os.path.join(ROOT, "tests", "emu", "roi_emu.cpp"), "-o", str(out)])
    return ctypes.CDLL(str(out))


def test_roi_kernels_emulated_match_cv2_procedure(emu):
    rng = np.random.default_rng(3)
    B, H, W, F_, R, Ro = 4, 240, 320, 64, 256, 64
    img = rng.integers(0, 256, (B, H, W, 3), dtype=np.uint8)
    xyz = np.zeros((B, H, W, 3), np.float32)
    seg = np.zeros((B, H, W), np.float32)
    trunc = (rng.random((B, H, W)) > 0.2).astype(np.float32)
    centers = np.stack([rng.uniform(100, 220, B), rng.uniform(75, 165, B)], 1)  # float64, like the reference's aug_bbox output
    centers[3] = [15.3, 10.7]  # crop hanging over the image border
    scales = rng.uniform(45, 150, B)
    ext = rng.uniform(0.05, 0.3, (B, 3)).astype(np.float32)
    fps = ((rng.random((B, F_, 3)) - 0.5) * ext[:, None, :]).astype(np.float32)
    yy, xx = np.mgrid[0:H, 0:W]
    for b in range(B):
        blob = ((xx - centers[b, 0]) ** 2 + (yy - centers[b, 1]) ** 2) < (0.35 * scales[b]) ** 2
        xyz[b][blob] = ((rng.random((int(blob.sum()), 3)) - 0.5) * ext[b]).astype(np.float32)
        seg[b] = (blob & (rng.random((H, W)) > 0.1)).astype(np.float32)
    out = dict(roi_img=np.zeros((B, 3, R, R), np.float32), roi_xyz=np.zeros((B, 3, Ro, Ro), np.float32),
               roi_mask_trunc=np.zeros((B, Ro, Ro), np.float32), roi_mask_visib=np.zeros((B, Ro, Ro), np.float32),
               roi_mask_obj=np.zeros((B, Ro, Ro), np.float32), roi_region=np.zeros((B, Ro, Ro), np.int64),
               roi_coord_2d=np.zeros((B, 2, Ro, Ro), np.float32))
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
    centers_c, scales_c = np.ascontiguousarray(centers, np.float64), np.ascontiguousarray(scales, np.float64)
    assert emu.gdrn_roi_crop_image(P(img), P(centers_c), P(scales_c), P(out["roi_img"]), B, H, W, R, ctypes.c_float(255.0), None) == 0
    assert emu.gdrn_roi_targets(P(xyz), P(seg), P(trunc), P(centers_c), P(scales_c), P(ext), P(fps), F_, P(out["roi_xyz"]),
                                P(out["roi_mask_trunc"]), P(out["roi_mask_visib"]), P(out["roi_mask_obj"]), P(out["roi_region"]),
                                P(out["roi_coord_2d"]), B, H, W, Ro, None) == 0
    for b in range(B):
        ref = RO.roi_instance(img[b], xyz[b], seg[b], trunc[b], centers[b], float(scales[b]), ext[b], fps[b])
        for k in ("roi_mask_trunc", "roi_mask_visib", "roi_mask_obj"):
            assert (out[k][b] != ref[k]).mean() < 5e-4, (b, k)
        assert (out["roi_region"][b] != ref["roi_region"]).mean() < 5e-4, b
        dx = np.abs(out["roi_xyz"][b] - ref["roi_xyz"])
        assert (dx > 1e-6).mean() < 5e-4, (b, (dx > 1e-6).mean())
        dc = np.abs(out["roi_coord_2d"][b] - ref["roi_coord_2d"])
        assert dc.max() < 1.01 / (32.0 * (W - 1)) + 1e-6 and (dc > 1e-6).mean() < 5e-4, (b, dc.max())
        d = np.abs(out["roi_img"][b] - ref["roi_img"]) * 255.0
        assert d.max() <= 1.0 + 1e-3 and (d > 0.5).mean() < 0.02, (b, d.max(), (d > 0.5).mean())


def test_roi_oracle_matches_reference_golden(golden_dir):
    """tests/golden/roi_targets_b3.npz: outputs of the UNMODIFIED reference helpers (oracle/make_golden_roi.py imports
    crop_resize_by_warp_affine / get_2d_coord_np / xyz_to_region from /root/reference).  The oracle's restatement of those helpers
    must reproduce them bit for bit (same cv2 build; a different build may differ in warpAffine's last bits)."""
    g = np.load(os.path.join(golden_dir, "roi_targets_b3.npz"))
    if str(g["cv2_version"]) != cv2.__version__:
        pytest.skip(f"golden made with cv2 {g['cv2_version']}, running {cv2.__version__}")
    B, H, W = g["image"].shape[:3]
    in_res, out_res = int(g["in_res"]), int(g["out_res"])
    x = np.linspace(0, 1, W, dtype=np.float32)
    y = np.linspace(0, 1, H, dtype=np.float32)
    coord_2d = np.asarray(np.meshgrid(x, y)).transpose(1, 2, 0)
    for b in range(B):
        c, s = g["centers"][b], float(g["scales"][b])
        xyz = g["xyz"][b]
        mask_obj = ((xyz[:, :, 0] != 0) | (xyz[:, :, 1] != 0) | (xyz[:, :, 2] != 0)).astype(np.float32)
        assert np.array_equal(RO.crop_resize_by_warp_affine(g["image"][b], c, s, in_res, cv2.INTER_LINEAR), g[f"img_{b}"])
        assert np.array_equal(RO.crop_resize_by_warp_affine(coord_2d, c, s, out_res, cv2.INTER_LINEAR), g[f"coord_{b}"])
        assert np.array_equal(RO.crop_resize_by_warp_affine(mask_obj[:, :, None], c, s, out_res, cv2.INTER_NEAREST), g[f"obj_{b}"])
        roi_xyz = RO.crop_resize_by_warp_affine(xyz, c, s, out_res, cv2.INTER_NEAREST)
        assert np.array_equal(roi_xyz, g[f"xyz_{b}"])
        assert np.array_equal(RO.xyz_to_region(roi_xyz, g["fps"][b]), g[f"region_{b}"])


def test_roi_kernels_emulated_match_reference_golden(emu, golden_dir):
    """The host-emulated kernels against the reference helpers' stored outputs (nearest-sampled targets and labels exact up to a
    stray pixel, coordinates within one 1/32-pixel step, image within one grey level)."""
    g = np.load(os.path.join(golden_dir, "roi_targets_b3.npz"))
    B, H, W = g["image"].shape[:3]
    R, Ro, F_ = int(g["in_res"]), int(g["out_res"]), g["fps"].shape[1]
    img, xyz = np.ascontiguousarray(g["image"]), np.ascontiguousarray(g["xyz"])
    seg, trunc = np.ascontiguousarray(g["seg"]), np.ascontiguousarray(g["trunc"])
    centers, scales = np.ascontiguousarray(g["centers"], np.float64), np.ascontiguousarray(g["scales"], np.float64)
    ext, fps = np.ascontiguousarray(g["extents"]), np.ascontiguousarray(g["fps"])
    out = dict(roi_img=np.zeros((B, 3, R, R), np.float32), roi_xyz=np.zeros((B, 3, Ro, Ro), np.float32),
               roi_mask_trunc=np.zeros((B, Ro, Ro), np.float32), roi_mask_visib=np.zeros((B, Ro, Ro), np.float32),
               roi_mask_obj=np.zeros((B, Ro, Ro), np.float32), roi_region=np.zeros((B, Ro, Ro), np.int64),
               roi_coord_2d=np.zeros((B, 2, Ro, Ro), np.float32))
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
    assert emu.gdrn_roi_crop_image(P(img), P(centers), P(scales), P(out["roi_img"]), B, H, W, R, ctypes.c_float(255.0), None) == 0
    assert emu.gdrn_roi_targets(P(xyz), P(seg), P(trunc), P(centers), P(scales), P(ext), P(fps), F_, P(out["roi_xyz"]),
                                P(out["roi_mask_trunc"]), P(out["roi_mask_visib"]), P(out["roi_mask_obj"]), P(out["roi_region"]),
                                P(out["roi_coord_2d"]), B, H, W, Ro, None) == 0
    for b in range(B):
        assert (out["roi_mask_obj"][b] != g[f"obj_{b}"]).mean() < 5e-4, b
        assert (out["roi_region"][b] != g[f"region_{b}"]).mean() < 5e-4, b
        want_xyz = g[f"xyz_{b}"].transpose(2, 0, 1) / g["extents"][b][:, None, None] + 0.5  # data_loader.py:541-545
        assert (np.abs(out["roi_xyz"][b] - want_xyz) > 1e-6).mean() < 5e-4, b
        dc = np.abs(out["roi_coord_2d"][b] - g[f"coord_{b}"].transpose(2, 0, 1))
        assert dc.max() < 1.01 / (32.0 * (W - 1)) + 1e-6 and (dc > 1e-6).mean() < 5e-4, (b, dc.max())
        d = np.abs(out["roi_img"][b] * 255.0 - g[f"img_{b}"].transpose(2, 0, 1).astype(np.float32))
        assert d.max() <= 1.0 + 1e-3 and (d > 0.5).mean() < 0.02, (b, d.max(), (d > 0.5).mean())
