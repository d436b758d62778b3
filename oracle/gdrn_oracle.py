"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the GDR-Net hot path.

A functional restatement (plain torch ops on CPU, fp32 or fp64) of the reference
algorithm.  It is pinned against the live, unmodified reference by
`oracle/make_golden.py` (run in the build container, where `/root/reference`
exists); the resulting fixtures live in `tests/golden/` and
`tests/test_oracle_golden.py` re-checks this file against them everywhere.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import this module.  The product path
(`gdr_net_b200/`) never does, and fails loudly when its CUDA library is missing.

Reference citations are relative to `/root/reference/`.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor

RESNET34_LAYERS = [3, 4, 6, 3]  # core/gdrn_modeling/models/resnet_backbone.py:10
RESNET34_PLANES = [64, 128, 256, 512]


# --------------------------------------------------------------------------------------
# normalisation helpers
# --------------------------------------------------------------------------------------
BN_MOMENTUM = [0.1]  # nn.BatchNorm2d default; fixtures.calibrated_state_dict temporarily sets 1.0


def _bn(x: Tensor, sd: Dict[str, Tensor], prefix: str, train: bool, update_stats: bool) -> Tensor:
    """nn.BatchNorm2d(eps=1e-5, momentum=0.1) (detectron2 BatchNorm2d == torch's; layer_utils.py:17-39)."""
    w, b = sd[prefix + ".weight"], sd[prefix + ".bias"]
    rm, rv = sd[prefix + ".running_mean"], sd[prefix + ".running_var"]
    if train:
        if update_stats:
            y = F.batch_norm(x, rm, rv, w, b, True, BN_MOMENTUM[0], 1e-5)
            key = prefix + ".num_batches_tracked"
            if key in sd:
                sd[key] += 1
            return y
        return F.batch_norm(x, None, None, w, b, True, 0.1, 1e-5)
    return F.batch_norm(x, rm.to(x.dtype), rv.to(x.dtype), w, b, False, 0.1, 1e-5)


# --------------------------------------------------------------------------------------
# a1: backbone  (resnet_backbone.py:53-80 + torchvision BasicBlock.forward)
# --------------------------------------------------------------------------------------
def backbone_forward(x: Tensor, sd: Dict[str, Tensor], train: bool, update_stats: bool = False) -> Tensor:
    p = "backbone."
    x = F.conv2d(x, sd[p + "conv1.weight"], None, stride=2, padding=3)  # :69
    x = F.relu(_bn(x, sd, p + "bn1", train, update_stats))  # :70-71
    x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)  # :72
    for li, (nblk, planes) in enumerate(zip(RESNET34_LAYERS, RESNET34_PLANES), start=1):
        for bi in range(nblk):
            q = f"{p}layer{li}.{bi}."
            stride = 2 if (bi == 0 and li > 1) else 1
            identity = x
            out = F.conv2d(x, sd[q + "conv1.weight"], None, stride=stride, padding=1)
            out = F.relu(_bn(out, sd, q + "bn1", train, update_stats))
            out = F.conv2d(out, sd[q + "conv2.weight"], None, stride=1, padding=1)
            out = _bn(out, sd, q + "bn2", train, update_stats)
            if (q + "downsample.0.weight") in sd:  # resnet_backbone.py:38-45
                identity = F.conv2d(x, sd[q + "downsample.0.weight"], None, stride=stride)
                identity = _bn(identity, sd, q + "downsample.1", train, update_stats)
            x = F.relu(out + identity)
    return x  # [B,512,8,8]


# --------------------------------------------------------------------------------------
# a2: geometry head (cdpn_rot_head_region.py:80-135 build, :183-193 forward)
# --------------------------------------------------------------------------------------
def head_forward(feat: Tensor, sd: Dict[str, Tensor], train: bool, update_stats: bool = False) -> Tensor:
    p = "rot_head_net.features."
    x = F.conv_transpose2d(feat, sd[p + "0.weight"], None, stride=2, padding=1, output_padding=1)  # :82-91
    x = F.relu(_bn(x, sd, p + "1", train, update_stats))
    conv_idx = [(3, 4), (6, 7), (10, 11), (13, 14), (17, 18), (20, 21)]
    for ci, bi in conv_idx:
        if ci in (10, 17):  # features[9], features[16]: UpsamplingBilinear2d => align_corners=True (:102)
            x = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=True)
        x = F.conv2d(x, sd[f"{p}{ci}.weight"], None, stride=1, padding=1)
        x = F.relu(_bn(x, sd, f"{p}{bi}", train, update_stats))
    x = F.conv2d(x, sd[p + "23.weight"], sd[p + "23.bias"])  # :127-135, 1x1, bias=True
    return x  # [B, 1+3+65, 64, 64]: ch0 mask | 1..3 xyz | 4..68 region (4 = bg)


# --------------------------------------------------------------------------------------
# a4: Patch-PnP (conv_pnp_net.py:111-157)
# --------------------------------------------------------------------------------------
def pnp_forward(coor_feat: Tensor, region: Tensor, extents: Tensor, sd: Dict[str, Tensor]):
    p = "pnp_net."
    bs, in_c = coor_feat.shape[:2]
    if in_c in (3, 5):  # :120-122 (in-place in the reference; functional here, same values)
        xyz = (coor_feat[:, :3] - 0.5) * extents.view(bs, 3, 1, 1)
        coor_feat = torch.cat([xyz, coor_feat[:, 3:]], dim=1)
    x = torch.cat([coor_feat, region], dim=1)  # :124-125
    for ci, gi in [(0, 1), (3, 4), (6, 7)]:  # :76-80,142-143
        x = F.conv2d(x, sd[f"{p}features.{ci}.weight"], None, stride=2, padding=1)
        x = F.relu(F.group_norm(x, 32, sd[f"{p}features.{gi}.weight"], sd[f"{p}features.{gi}.bias"], 1e-5))
    x = x.reshape(-1, 128 * 8 * 8)  # :145
    x = F.leaky_relu(F.linear(x, sd[p + "fc1.weight"], sd[p + "fc1.bias"]), 0.1)  # :152
    x = F.leaky_relu(F.linear(x, sd[p + "fc2.weight"], sd[p + "fc2.bias"]), 0.1)  # :153
    rot = F.linear(x, sd[p + "fc_r.weight"], sd[p + "fc_r.bias"])  # :155
    t = F.linear(x, sd[p + "fc_t.weight"], sd[p + "fc_t.bias"])  # :156
    return rot, t


# --------------------------------------------------------------------------------------
# a5: rot6d -> R  (core/utils/rot_reps.py:9-49)
# --------------------------------------------------------------------------------------
def _cross(u: Tensor, v: Tensor) -> Tensor:
    return torch.stack(
        [u[:, 1] * v[:, 2] - u[:, 2] * v[:, 1], u[:, 2] * v[:, 0] - u[:, 0] * v[:, 2], u[:, 0] * v[:, 1] - u[:, 1] * v[:, 0]],
        dim=1,
    )


def ortho6d_to_mat(poses: Tensor) -> Tensor:
    x = F.normalize(poses[:, 0:3], p=2, dim=1)  # eps 1e-12
    z = F.normalize(_cross(x, poses[:, 3:6]), p=2, dim=1)
    y = _cross(z, x)
    return torch.stack([x, y, z], dim=2)  # columns x | y | z


# --------------------------------------------------------------------------------------
# a6: pose decode (pose_from_pred_centroid_z.py:144-227, utils.py:208-236, pose_utils.py:323-370)
# --------------------------------------------------------------------------------------
def quat2mat(quat: Tensor) -> Tensor:
    q = quat / quat.norm(p=2, dim=1, keepdim=True)  # eps = 0 at this call site
    qw, qx, qy, qz = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    X, Y, Z = 2 * qx, 2 * qy, 2 * qz
    wX, wY, wZ = qw * X, qw * Y, qw * Z
    xX, xY, xZ = qx * X, qx * Y, qx * Z
    yY, yZ, zZ = qy * Y, qy * Z, qz * Z
    return torch.stack(
        [1 - (yY + zZ), xY - wZ, xZ + wY, xY + wZ, 1 - (xX + zZ), yZ - wX, xZ - wY, yZ + wX, 1 - (xX + yY)], dim=1
    ).reshape(-1, 3, 3)


def translation_from_site(pred_centroids, pred_z_vals, roi_cams, roi_centers, resize_ratios, roi_whs) -> Tensor:
    """z_type == "REL" branch (:176-214)."""
    cx = pred_centroids[:, 0:1] * roi_whs[:, 0:1] + roi_centers[:, 0:1]
    cy = pred_centroids[:, 1:2] * roi_whs[:, 1:2] + roi_centers[:, 1:2]
    z = pred_z_vals * resize_ratios.view(-1, 1)
    return torch.cat(
        [z * (cx - roi_cams[:, 0:1, 2]) / roi_cams[:, 0:1, 0], z * (cy - roi_cams[:, 1:2, 2]) / roi_cams[:, 1:2, 1], z],
        dim=1,
    )


def allo_to_ego_mat(translation: Tensor, rot_allo: Tensor, eps: float = 1e-4) -> Tensor:
    obj_ray = translation / (translation.norm(dim=1, keepdim=True) + eps)
    angle = obj_ray[:, 2:3].acos()
    # cross((0,0,1), obj_ray) = (-ry, rx, 0)
    axis = torch.stack([-obj_ray[:, 1], obj_ray[:, 0], torch.zeros_like(obj_ray[:, 0])], dim=1)
    axis = axis / (axis.norm(dim=1, keepdim=True) + eps)
    q = torch.cat([torch.cos(angle / 2.0), axis * torch.sin(angle / 2.0)], dim=1)
    return quat2mat(q) @ rot_allo


def pose_decode_train(rot_allo, pred_t_, roi_cams, roi_centers, resize_ratios, roi_whs):
    trans = translation_from_site(pred_t_[:, :2], pred_t_[:, 2:3], roi_cams, roi_centers, resize_ratios, roi_whs)
    return allo_to_ego_mat(trans, rot_allo, eps=1e-4), trans


def _axangle2mat_np(axis, angle):
    x, y, z = [float(v) for v in axis]
    n = math.sqrt(x * x + y * y + z * z)
    x, y, z = x / n, y / n, z / n
    c, s = math.cos(angle), math.sin(angle)
    C = 1 - c
    return np.array(
        [
            [x * x * C + c, x * y * C - z * s, z * x * C + y * s],
            [x * y * C + z * s, y * y * C + c, y * z * C - x * s],
            [z * x * C - y * s, y * z * C + x * s, z * z * C + c],
        ]
    )


def pose_decode_test(rot_allo, pred_t_, roi_cams, roi_centers, resize_ratios, roi_whs):
    """Test-mode branch: numpy per-sample loop (pose_from_pred_centroid_z.py:52-141, utils.py:39-94)."""
    trans = translation_from_site(pred_t_[:, :2], pred_t_[:, 2:3], roi_cams, roi_centers, resize_ratios, roi_whs)
    R = rot_allo.detach().cpu().numpy()
    out = np.zeros_like(R)
    for i in range(R.shape[0]):
        t = trans[i].detach().cpu().numpy()
        obj_ray = t / np.linalg.norm(t)
        angle = math.acos(float(obj_ray[2]))
        if angle > 0:
            axis = np.cross(np.array([0, 0, 1.0]), obj_ray)
            out[i] = _axangle2mat_np(axis, angle).dot(R[i])
        else:
            out[i] = R[i]
    return torch.from_numpy(out), trans


# --------------------------------------------------------------------------------------
# a9: closest symmetric GT rotation (pose_utils.py:430-482, pose_error.py:400-415)
# --------------------------------------------------------------------------------------
def _re_deg(R_est: np.ndarray, R_gt: np.ndarray) -> float:
    trace = float(np.trace(R_est.dot(R_gt.T)))
    trace = trace if trace <= 3 else 3
    return float(np.rad2deg(np.arccos(min(1.0, max(-1.0, 0.5 * (trace - 1.0))))))


def closest_rot_batch(pred_rots: Tensor, gt_rots: Tensor, sym_infos: List[Optional[Tensor]]) -> Tensor:
    out = gt_rots.clone().cpu().numpy()
    for i in range(pred_rots.shape[0]):
        sym = sym_infos[i]
        if sym is None:
            continue
        sym = sym.cpu().numpy() if isinstance(sym, torch.Tensor) else np.asarray(sym)
        sym = sym.reshape(-1, 3, 3)
        est = pred_rots[i].detach().cpu().numpy()
        gt = gt_rots[i].cpu().numpy()
        best, best_R = _re_deg(est, gt), gt
        for k in range(sym.shape[0]):
            cand = gt.dot(sym[k])
            e = _re_deg(est, cand)
            if e < best:
                best, best_R = e, cand
        out[i] = best_R
    return torch.tensor(out, dtype=gt_rots.dtype, device=gt_rots.device)  # pose_utils.py:481


# --------------------------------------------------------------------------------------
# a7/a8/a10: losses (GDRN.py:341-471, pm_loss.py:82-114, misc.py:930-949)
# --------------------------------------------------------------------------------------
def gdrn_losses(head_out: Tensor, pred_ego_rot: Tensor, pred_t_: Tensor, batch: dict, pm_sym: bool = False):
    out_mask = head_out[:, 0:1]
    out_x, out_y, out_z = head_out[:, 1:2], head_out[:, 2:3], head_out[:, 3:4]
    out_region = head_out[:, 4:]
    dt = head_out.dtype
    gt_xyz = batch["roi_xyz"].to(dt)
    m_vis = batch["roi_mask_visib"].to(dt)
    m_trunc = batch["roi_mask_trunc"].to(dt)
    L = {}
    denom = m_vis.sum().float().clamp(min=1.0)
    for name, o, c in (("x", out_x, 0), ("y", out_y, 1), ("z", out_z, 2)):  # :345-355
        L[f"loss_coor_{name}"] = F.l1_loss(o * m_vis[:, None], gt_xyz[:, c : c + 1] * m_vis[:, None], reduction="sum") / denom
    L["loss_mask"] = F.l1_loss(out_mask[:, 0], m_trunc, reduction="mean")  # :378-379
    gt_region = batch["roi_region"].long()
    L["loss_region"] = (  # :392-397 (logits AND labels are multiplied by the mask)
        F.cross_entropy(out_region * m_vis[:, None], gt_region * m_vis.long(), reduction="sum") / denom
    )
    # PM loss, r_only, L1-mean, norm_by_extent (pm_loss.py:82-114)
    gt_rot = batch["ego_rot"].to(dt)
    if pm_sym:
        gt_rot = closest_rot_batch(pred_ego_rot, gt_rot, batch["sym_info"]).to(dt)
    pts = batch["roi_points"].to(dt)
    w = (1.0 / batch["roi_extent"].to(dt).max(1, keepdim=True)[0]).view(-1, 1, 1)
    est = (pred_ego_rot[:, None] @ pts[..., None]).squeeze(-1)
    tgt = (gt_rot[:, None] @ pts[..., None]).squeeze(-1)
    L["loss_PM_R"] = 3 * F.l1_loss(w * est, w * tgt, reduction="mean")
    ratio = batch["roi_trans_ratio"].to(dt)
    L["loss_centroid"] = F.l1_loss(pred_t_[:, :2], ratio[:, :2], reduction="mean")  # :439-452
    L["loss_z"] = F.l1_loss(pred_t_[:, 2], ratio[:, 2], reduction="mean")  # :455-471
    return L


def mean_re_te(pred_trans, pred_rot, gt_trans, gt_rot):
    """model_utils.py:40-52 (float32 accumulation like the reference)."""
    pt, pr = pred_trans.detach().cpu().numpy(), pred_rot.detach().cpu().numpy()
    gt, gr = gt_trans.detach().cpu().numpy(), gt_rot.detach().cpu().numpy()
    bs = pr.shape[0]
    R_errs = np.zeros((bs,), dtype=np.float32)
    T_errs = np.zeros((bs,), dtype=np.float32)
    for i in range(bs):
        R_errs[i] = _re_deg(pr[i], gr[i])
        T_errs[i] = np.linalg.norm(gt[i].flatten() - pt[i].flatten())
    return R_errs.mean(), T_errs.mean()


# --------------------------------------------------------------------------------------
# GDRN.forward (GDRN.py:83-306), a6 configuration
# --------------------------------------------------------------------------------------
def gdrn_forward(sd: Dict[str, Tensor], batch: dict, train: bool, do_loss: bool, pm_sym: bool = False,
                 update_stats: bool = False) -> dict:
    """`train` selects BN batch statistics (module.train()); `do_loss` mirrors the kwarg."""
    x = batch["roi_img"]
    feat = backbone_forward(x, sd, train, update_stats)  # :121
    head = head_forward(feat, sd, train, update_stats)  # :123
    coor_feat = head[:, 1:4]
    if sd["pnp_net.features.0.weight"].shape[1] == 69:  # cfg PNP_NET.WITH_2D_COORD (:171-173); 67 input channels without
        coor_feat = torch.cat([coor_feat, batch["roi_coord_2d"].to(head.dtype)], dim=1)  # :162-166
    region_softmax = F.softmax(head[:, 5:], dim=1)  # :169 (region[:,1:])
    rot6d, pred_t_ = pnp_forward(coor_feat, region_softmax, batch["roi_extent"].to(head.dtype), sd)  # :179-181
    rot_m = ortho6d_to_mat(rot6d)  # :193-194
    args = (batch["roi_cam"].to(head.dtype), batch["roi_center"].to(head.dtype), batch["resize_ratio"].to(head.dtype),
            batch["roi_wh"].to(head.dtype))
    res = dict(head=head, rot6d=rot6d, pred_t_=pred_t_, rot_allo=rot_m, feat=feat)
    if not do_loss:
        rot, trans = pose_decode_test(rot_m, pred_t_, *args)  # :198-212 with is_train=False
        res.update(rot=rot, trans=trans)
        return res
    rot, trans = pose_decode_train(rot_m, pred_t_, *args)
    res.update(rot=rot, trans=trans)
    res["losses"] = gdrn_losses(head, rot, pred_t_, batch, pm_sym=pm_sym)
    re_, te_ = mean_re_te(trans, rot, batch["trans"], batch["ego_rot"])  # :246
    res["vis"] = {"vis/error_R": float(re_), "vis/error_t": float(te_) * 100}
    return res


# --------------------------------------------------------------------------------------
# ADD / ADD-S (lib/pysixd/pose_error.py:297-337) -- the accuracy-parity metric
# --------------------------------------------------------------------------------------
def add_metric(R_est, t_est, R_gt, t_gt, pts) -> float:
    pe = pts @ R_est.T + t_est.reshape(1, 3)
    pg = pts @ R_gt.T + t_gt.reshape(1, 3)
    return float(np.linalg.norm(pe - pg, axis=1).mean())


def adi_metric(R_est, t_est, R_gt, t_gt, pts) -> float:
    from scipy import spatial

    pe = pts @ R_est.T + t_est.reshape(1, 3)
    pg = pts @ R_gt.T + t_gt.reshape(1, 3)
    d, _ = spatial.cKDTree(pe).query(pg, k=1)
    return float(d.mean())


def leaf_state_dict(sd: Dict[str, Tensor], dtype=torch.float32, requires_grad: bool = True) -> Dict[str, Tensor]:
    """Clone a state dict into autograd leaves (parameters) + plain buffers."""
    out = {}
    for k, v in sd.items():
        if not v.dtype.is_floating_point:
            out[k] = v.clone()
            continue
        t = v.detach().clone().to(dtype)
        is_buf = k.endswith("running_mean") or k.endswith("running_var")
        if requires_grad and not is_buf:
            t.requires_grad_(True)
        out[k] = t
    return out
