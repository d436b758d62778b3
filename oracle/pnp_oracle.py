"""CPU oracle for the test-time RANSAC-PnP of the evaluator -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module; the product path
(gdr_net_b200/pnp_ransac.py -> csrc/pnp_ransac.cu) never does.

What the reference does (core/gdrn_modeling/gdrn_evaluator.py:316-436):
  * `get_img_model_points_with_coords2d` (gdrn_evaluator.py:89-126): model points = (xyz - 0.5) * extent, image points =
    coord2d * (im_W, im_H), selected where mask > thr and |xyz_c| > 1e-4 * extent_c, in row-major pixel order;
  * `misc.pnp_v2(..., method=cv2.SOLVEPNP_EPNP, ransac=True, ransac_reprojErr=3, ransac_iter=100)` (lib/pysixd/misc.py:145-194)
    = `cv2.solvePnPRansac` + `cv2.Rodrigues`;
  * translation replaced by the network's (gdrn_evaluator.py:398-421) -- that part is the decode already on the device.

The arithmetic lives in a third-party dependency that is NOT under /root/reference: OpenCV (`opencv-python`, requirements.txt,
unpinned; 4.13.0 in this image).  Two oracles are kept here:
  * `pnp_ransac_cv2`     -- the reference's own call, executed by the installed cv2 (the real thing);
  * `pnp_ransac_restated`-- a numpy restatement of the published algorithm OpenCV runs for this call (calib3d solvePnPRansac:
    RNG-driven 5-point subsets, EPnP [Lepetit, Moreno-Noguer, Fua, IJCV 2009] as the minimal and as the final solver, squared
    reprojection error threshold, adaptive iteration count), PINNED against cv2 by tests/test_pnp_ransac_cpu.py (same inlier
    sets, poses equal to ~1e-8) -- the CUDA kernel follows this restatement line by line.

Pinning against the reference itself: oracle/make_golden_pnp.py imports the UNMODIFIED `get_out_mask`,
`GDRN_Evaluator.get_img_model_points_with_coords2d` and `misc.pnp_v2` from /root/reference and stores their outputs for a
seeded batch in tests/golden/pnp_ransac_b4.npz; `select_points` reproduces those point lists bit for bit and `pnp_ransac_cv2`
the poses (tests/test_pnp_ransac_cpu.py::test_oracle_matches_reference_golden).
"""
from __future__ import annotations

import numpy as np

# ------------------------------------------------------------------------------------------------ selection


def select_points(mask, xyz, coord2d, extent, im_W, im_H, mask_thr=0.5):
    """gdrn_evaluator.py:89-126 (max_num_points = -1).  mask [H,W] f32, xyz [3,H,W] f32 in [0,1], coord2d [2,H,W] f32 in [0,1],
    extent [3] f32.  Returns (image_points [n,2] f32, model_points [n,3] f32) in row-major pixel order, float32 like the reference."""
    xyz = np.asarray(xyz, np.float32).transpose(1, 2, 0).copy()
    c2d = np.asarray(coord2d, np.float32).transpose(1, 2, 0).copy()
    extent = np.asarray(extent, np.float32)
    for c in range(3):
        xyz[:, :, c] = (xyz[:, :, c] - np.float32(0.5)) * extent[c]
    c2d[:, :, 0] = c2d[:, :, 0] * im_W
    c2d[:, :, 1] = c2d[:, :, 1] * im_H
    sel = np.asarray(mask, np.float32) > mask_thr
    for c in range(3):
        sel &= np.abs(xyz[:, :, c]) > 0.0001 * extent[c]
    return c2d[sel].reshape(-1, 2), xyz[sel].reshape(-1, 3)


def pnp_ransac_cv2(model_points, image_points, K, reproj_err=3.0, iters=100):
    """misc.pnp_v2:145-194 with method=EPNP, ransac=True.  Returns ([3,4] pose, inlier index array)."""
    import cv2

    p3 = np.ascontiguousarray(np.expand_dims(model_points, 0).astype(np.float64))
    p2 = np.ascontiguousarray(np.expand_dims(image_points, 0).astype(np.float64))
    ok, rvec, tvec, inliers = cv2.solvePnPRansac(p3, p2, np.asarray(K, np.float64), np.zeros((8, 1)), flags=cv2.SOLVEPNP_EPNP,
                                                 reprojectionError=reproj_err, iterationsCount=iters)
    R, _ = cv2.Rodrigues(rvec)
    inl = np.zeros(0, np.int64) if inliers is None else np.asarray(inliers).reshape(-1).astype(np.int64)
    return np.concatenate([R, np.asarray(tvec).reshape(3, 1)], axis=-1), inl


# ------------------------------------------------------------------------------------------------ EPnP (restated)


def jacobi_svd_cv(A):
    """cv::SVD::compute for a square double matrix (OpenCV core/src/lapack.cpp, JacobiSVDImpl_: one-sided Jacobi on the columns of
    A, V accumulated from the identity, singular values sorted descending by selection).  Returns (w, Ut, Vt), the rows of Ut / Vt
    being the left / right singular vectors.  The SIGNS matter: EPnP's control points are c0 +- k u_i and the pose for noisy data
    depends on that sign at the 1e-4 level, so the restatement has to reproduce them -- checked against cv2.SVDecomp (1e-16)."""
    A = np.asarray(A, np.float64)
    n = A.shape[0]
    At = A.T.copy()
    Vt = np.eye(n)
    W = (At * At).sum(1)
    eps = np.finfo(np.float64).eps * 10
    for _ in range(max(n, 30)):
        changed = False
        for i in range(n - 1):
            for j in range(i + 1, n):
                a, b = W[i], W[j]
                p = float(At[i] @ At[j])
                if abs(p) <= eps * np.sqrt(a * b):
                    continue
                p *= 2
                beta = a - b
                gamma = np.hypot(p, beta)
                if beta < 0:
                    sn = np.sqrt((gamma - beta) * 0.5 / gamma)
                    cs = p / (gamma * sn * 2)
                else:
                    cs = np.sqrt((gamma + beta) / (gamma * 2))
                    sn = p / (gamma * cs * 2)
                t0, t1 = cs * At[i] + sn * At[j], -sn * At[i] + cs * At[j]
                At[i], At[j] = t0, t1
                W[i], W[j] = t0 @ t0, t1 @ t1
                changed = True
                v0, v1 = cs * Vt[i] + sn * Vt[j], -sn * Vt[i] + cs * Vt[j]
                Vt[i], Vt[j] = v0, v1
        if not changed:
            break
    W = np.sqrt((At * At).sum(1))
    for i in range(n - 1):
        j = i
        for k in range(i + 1, n):
            if W[j] < W[k]:
                j = k
        if i != j:
            W[[i, j]] = W[[j, i]]
            At[[i, j]] = At[[j, i]]
            Vt[[i, j]] = Vt[[j, i]]
    Ut = At / np.where(W > np.finfo(np.float64).tiny, W, 1.0)[:, None]
    return W, Ut, Vt


def _sym_eig_desc(A):
    """cvSVD(MODIFY_A | U_T) of a symmetric PSD matrix: singular values descending, rows of Ut = left singular vectors."""
    w, ut, _ = jacobi_svd_cv(A)
    return w, ut


def _choose_control_points(pws):
    n = len(pws)
    cws = np.zeros((4, 3))
    cws[0] = pws.sum(0) / n
    pw0 = pws - cws[0]
    dc, uct = _sym_eig_desc(pw0.T @ pw0)
    for i in range(1, 4):
        k = np.sqrt(max(dc[i - 1], 0.0) / n)
        cws[i] = cws[0] + k * uct[i - 1]
    return cws


def _barycentric(pws, cws):
    cc = (cws[1:] - cws[0]).T  # columns = c_j - c_0
    ci = np.linalg.inv(cc)
    a123 = (pws - cws[0]) @ ci.T
    a0 = 1.0 - a123.sum(1)
    return np.concatenate([a0[:, None], a123], axis=1)


def _fill_M(alphas, us, fu, fv, uc, vc):
    n = len(us)
    M = np.zeros((2 * n, 12))
    for j in range(4):
        M[0::2, 3 * j] = alphas[:, j] * fu
        M[0::2, 3 * j + 2] = alphas[:, j] * (uc - us[:, 0])
        M[1::2, 3 * j + 1] = alphas[:, j] * fv
        M[1::2, 3 * j + 2] = alphas[:, j] * (vc - us[:, 1])
    return M


_PAIRS = [(0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3)]


def _L_6x10(ut):
    v = [ut[11], ut[10], ut[9], ut[8]]
    dv = np.zeros((4, 6, 3))
    for i in range(4):
        for k, (a, b) in enumerate(_PAIRS):
            dv[i, k] = v[i][3 * a:3 * a + 3] - v[i][3 * b:3 * b + 3]
    L = np.zeros((6, 10))
    d = lambda i, j, k: float(dv[i, k] @ dv[j, k])  # noqa: E731
    for k in range(6):
        L[k] = [d(0, 0, k), 2 * d(0, 1, k), d(1, 1, k), 2 * d(0, 2, k), 2 * d(1, 2, k), d(2, 2, k), 2 * d(0, 3, k), 2 * d(1, 3, k),
                2 * d(2, 3, k), d(3, 3, k)]
    return L


def _rho(cws):
    return np.array([((cws[a] - cws[b]) ** 2).sum() for a, b in _PAIRS])


def _lstsq(A, b):
    return np.linalg.lstsq(A, b, rcond=None)[0]


def _betas_approx_1(L, rho):  # betas10 = [B11 B12 B22 B13 B23 B33 B14 B24 B34 B44];  approx_1 = [B11 B12 B13 B14]
    b4 = _lstsq(L[:, [0, 1, 3, 6]], rho)
    betas = np.zeros(4)
    if b4[0] < 0:
        betas[0] = np.sqrt(-b4[0])
        betas[1:] = -b4[1:] / betas[0]
    else:
        betas[0] = np.sqrt(b4[0])
        betas[1:] = b4[1:] / betas[0]
    return betas


def _betas_approx_2(L, rho):  # approx_2 = [B11 B12 B22]
    b3 = _lstsq(L[:, [0, 1, 2]], rho)
    betas = np.zeros(4)
    if b3[0] < 0:
        betas[0] = np.sqrt(-b3[0])
        betas[1] = np.sqrt(-b3[2]) if b3[2] < 0 else 0.0
    else:
        betas[0] = np.sqrt(b3[0])
        betas[1] = np.sqrt(b3[2]) if b3[2] > 0 else 0.0
    if b3[1] < 0:
        betas[0] = -betas[0]
    return betas


def _betas_approx_3(L, rho):  # approx_3 = [B11 B12 B22 B13 B23]
    b5 = _lstsq(L[:, [0, 1, 2, 3, 4]], rho)
    betas = np.zeros(4)
    if b5[0] < 0:
        betas[0] = np.sqrt(-b5[0])
        betas[1] = np.sqrt(-b5[2]) if b5[2] < 0 else 0.0
    else:
        betas[0] = np.sqrt(b5[0])
        betas[1] = np.sqrt(b5[2]) if b5[2] > 0 else 0.0
    if b5[1] < 0:
        betas[0] = -betas[0]
    betas[2] = b5[3] / betas[0]
    return betas


def _gauss_newton(L, rho, betas, iters=5):
    b = betas.copy()
    for _ in range(iters):
        A = np.zeros((6, 4))
        r = np.zeros(6)
        for i in range(6):
            l = L[i]
            A[i, 0] = 2 * l[0] * b[0] + l[1] * b[1] + l[3] * b[2] + l[6] * b[3]
            A[i, 1] = l[1] * b[0] + 2 * l[2] * b[1] + l[4] * b[2] + l[7] * b[3]
            A[i, 2] = l[3] * b[0] + l[4] * b[1] + 2 * l[5] * b[2] + l[8] * b[3]
            A[i, 3] = l[6] * b[0] + l[7] * b[1] + l[8] * b[2] + 2 * l[9] * b[3]
            r[i] = rho[i] - (l[0] * b[0] * b[0] + l[1] * b[0] * b[1] + l[2] * b[1] * b[1] + l[3] * b[0] * b[2] + l[4] * b[1] * b[2]
                             + l[5] * b[2] * b[2] + l[6] * b[0] * b[3] + l[7] * b[1] * b[3] + l[8] * b[2] * b[3] + l[9] * b[3] * b[3])
        b = b + _lstsq(A, r)
    return b


def _R_t_from_betas(ut, betas, alphas, pws, us, fu, fv, uc, vc):
    ccs = np.zeros((4, 3))
    for i in range(4):
        ccs += betas[i] * ut[11 - i].reshape(4, 3)
    pcs = alphas @ ccs
    if pcs[0, 2] < 0.0:  # solve_for_sign
        ccs, pcs = -ccs, -pcs
    pc0, pw0 = pcs.mean(0), pws.mean(0)
    abt = (pcs - pc0).T @ (pws - pw0)
    _, ut_, vt = jacobi_svd_cv(abt)
    R = ut_.T @ vt
    if np.linalg.det(R) < 0:
        R[2] = -R[2]
    t = pc0 - R @ pw0
    pc = pws @ R.T + t
    ue = uc + fu * pc[:, 0] / pc[:, 2]
    ve = vc + fv * pc[:, 1] / pc[:, 2]
    err = np.sqrt((us[:, 0] - ue) ** 2 + (us[:, 1] - ve) ** 2).sum() / len(pws)
    return R, t, err


def epnp(pws, us, fu, fv, uc, vc):
    """EPnP as OpenCV runs it for SOLVEPNP_EPNP: solvePnP normalises the image points (undistortPoints, zero distortion) and the
    epnp object maps them back to pixels with the camera matrix (`us[i] = x_n * fu + uc`), so the linear system is built in
    pixel units.  Returns (R [3,3], t [3])."""
    pws, us = np.asarray(pws, np.float64), np.asarray(us, np.float64)
    cws = _choose_control_points(pws)
    alphas = _barycentric(pws, cws)
    M = _fill_M(alphas, us, fu, fv, uc, vc)
    _, ut = _sym_eig_desc(M.T @ M)
    L, rho = _L_6x10(ut), _rho(cws)
    best = None
    for approx in (_betas_approx_1, _betas_approx_2, _betas_approx_3):
        betas = _gauss_newton(L, rho, approx(L, rho))
        R, t, err = _R_t_from_betas(ut, betas, alphas, pws, us, fu, fv, uc, vc)
        if best is None or err < best[2]:
            best = (R, t, err)
    return best[0], best[1]


# ------------------------------------------------------------------------------------------------ RANSAC (restated)


class CvRNG:
    """cv::RNG (multiply-with-carry), `RNG rng((uint64)-1)` as in RANSACPointSetRegistrator::run."""

    def __init__(self, state=0xFFFFFFFFFFFFFFFF):
        self.state = state

    def next(self):
        self.state = ((self.state & 0xFFFFFFFF) * 4164903690 + (self.state >> 32)) & 0xFFFFFFFFFFFFFFFF
        return self.state & 0xFFFFFFFF

    def uniform(self, a, b):
        return a if a == b else int(self.next() % (b - a) + a)


def _rodrigues_roundtrip(R):
    """The RANSAC model is stored as (rvec, tvec): R -> rvec -> R.  Done with the closed forms cv::Rodrigues uses."""
    import cv2

    rvec, _ = cv2.Rodrigues(R)
    R2, _ = cv2.Rodrigues(rvec)
    return R2


def ransac_update_num_iters(p, ep, model_points, max_iters):
    p = min(max(p, 0.0), 1.0)
    ep = min(max(ep, 0.0), 1.0)
    num = max(1.0 - p, np.finfo(np.float64).tiny)
    denom = 1.0 - (1.0 - ep) ** model_points
    if denom < np.finfo(np.float64).tiny:
        return 0
    num, denom = np.log(num), np.log(denom)
    return max_iters if (denom >= 0 or -num >= max_iters * (-denom)) else int(round(num / denom))


def pnp_ransac_restated(model_points, image_points, K, reproj_err=3.0, iters=100, confidence=0.99, exact_rodrigues=False):
    """Restatement of cv2.solvePnPRansac(flags=EPNP) for n > 4 points.  Returns ([3,4] pose, inlier indices, iterations run)."""
    pws = np.asarray(model_points, np.float64)
    ipts = np.asarray(image_points, np.float64)
    K = np.asarray(K, np.float64)
    n = len(pws)
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    # undistortPoints (zero distortion) then epnp::init_points: pixel -> normalised -> pixel
    us = np.stack([(ipts[:, 0] - cx) * (1.0 / fx) * fx + cx, (ipts[:, 1] - cy) * (1.0 / fy) * fy + cy], axis=1)
    rng = CvRNG()
    thr = reproj_err * reproj_err
    niters, best_count, best_mask = iters, 0, None
    it = 0
    model_n = 5
    while it < niters:
        idx = []
        for _ in range(model_n):  # getSubset: distinct indices, redraw on repetition
            i = rng.uniform(0, n)
            while i in idx:
                i = rng.uniform(0, n)
            idx.append(i)
        R, t = epnp(pws[idx], us[idx], fx, fy, cx, cy)
        if exact_rodrigues:
            R = _rodrigues_roundtrip(R)
        pc = pws @ R.T + t
        u = fx * pc[:, 0] / pc[:, 2] + cx
        v = fy * pc[:, 1] / pc[:, 2] + cy
        err = ((ipts[:, 0] - u) ** 2 + (ipts[:, 1] - v) ** 2).astype(np.float32)
        mask = err <= thr
        count = int(mask.sum())
        if count > max(best_count, model_n - 1):
            best_count, best_mask = count, mask
            niters = ransac_update_num_iters(confidence, (n - count) / n, model_n, niters)
        it += 1
    if best_mask is None:
        return np.concatenate([np.eye(3), np.zeros((3, 1))], axis=1), np.zeros(0, np.int64), it
    inl = np.nonzero(best_mask)[0]
    R, t = epnp(pws[inl], us[inl], fx, fy, cx, cy)
    return np.concatenate([R, t.reshape(3, 1)], axis=1), inl, it
