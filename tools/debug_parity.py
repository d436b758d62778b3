"""Stage-by-stage comparison of the CUDA engine against the CPU oracle (debug aid, runs on the GPU box).
usage: python tools/debug_parity.py [fp32x3|bf16] [train|eval] [B]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

from gdr_net_b200 import GDRN as G
from gdr_net_b200 import synth
from gdr_net_b200.config import a6_config
from oracle import fixtures
from oracle import gdrn_oracle as O


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def nchw(pt):
    return pt.float().permute(0, 3, 1, 2)


def main():
    precision = sys.argv[1] if len(sys.argv) > 1 else "fp32x3"
    train = (sys.argv[2] if len(sys.argv) > 2 else "train") == "train"
    B = int(sys.argv[3]) if len(sys.argv) > 3 else 4
    sd = fixtures.calibrated_state_dict(0)
    cfg = a6_config(device="cuda")
    model, _ = G.build_model_optimizer(cfg, precision=precision)
    model.load_state_dict(sd)
    model.train(train)
    batch_cpu = synth.make_batch(B, seed=1)
    batch = {k: (v.cuda() if isinstance(v, torch.Tensor) else v) for k, v in batch_cpu.items()}
    eng = model.engine
    kw = synth.forward_kwargs(batch, train=True)
    aux = dict(roi_coord_2d=kw["roi_coord_2d"], roi_cams=kw["roi_cams"], roi_centers=kw["roi_centers"], roi_whs=kw["roi_whs"],
               roi_extents=kw["roi_extents"], resize_ratios=kw["resize_ratios"], gt_xyz=kw["gt_xyz"],
               gt_mask_trunc=kw["gt_mask_trunc"], gt_mask_visib=kw["gt_mask_visib"], gt_region=kw["gt_region"],
               gt_ego_rot=kw["gt_ego_rot"], gt_points=kw["gt_points"], sym_infos=None, gt_trans=kw["gt_trans"],
               gt_trans_ratio=kw["gt_trans_ratio"])
    aux = {k: (v.float().contiguous() if isinstance(v, torch.Tensor) and v.dtype != torch.long else v) for k, v in aux.items()}
    res = eng.forward(batch["roi_img"].float().contiguous(), aux, train_bn=train, do_loss=True)
    S = eng.saved
    torch.cuda.synchronize()

    # ---- oracle, stage by stage (same code as oracle.backbone_forward / head_forward, with taps)
    leaf = O.leaf_state_dict(sd)
    inter = {}
    x = batch_cpu["roi_img"]
    p = "backbone."
    u0 = F.conv2d(x, leaf[p + "conv1.weight"], None, stride=2, padding=3)
    inter["stem.u0"] = u0
    a0 = F.relu(O._bn(u0, leaf, p + "bn1", train, False))
    inter["stem.a0"] = a0
    cur = F.max_pool2d(a0, 3, 2, 1)
    inter["pool"] = cur
    bi_ = 0
    for li, nblk in enumerate(O.RESNET34_LAYERS, start=1):
        for bi in range(nblk):
            q = f"{p}layer{li}.{bi}."
            stride = 2 if (bi == 0 and li > 1) else 1
            idn = cur
            u1 = F.conv2d(cur, leaf[q + "conv1.weight"], None, stride=stride, padding=1)
            a1 = F.relu(O._bn(u1, leaf, q + "bn1", train, False))
            u2 = F.conv2d(a1, leaf[q + "conv2.weight"], None, stride=1, padding=1)
            o2 = O._bn(u2, leaf, q + "bn2", train, False)
            if (q + "downsample.0.weight") in leaf:
                idn = O._bn(F.conv2d(cur, leaf[q + "downsample.0.weight"], None, stride=stride), leaf, q + "downsample.1", train, False)
            cur = F.relu(o2 + idn)
            inter[f"blk{bi_}.u1"], inter[f"blk{bi_}.a1"], inter[f"blk{bi_}.u2"], inter[f"blk{bi_}.out"] = u1, a1, u2, cur
            bi_ += 1
    feat = cur
    hp = "rot_head_net.features."
    xh = F.conv_transpose2d(feat, leaf[hp + "0.weight"], None, stride=2, padding=1, output_padding=1)
    inter["deconv.u"] = xh
    xh = F.relu(O._bn(xh, leaf, hp + "1", train, False))
    inter["deconv.y"] = xh
    for i, (ci, bi, up) in enumerate([(3, 4, 0), (6, 7, 0), (10, 11, 1), (13, 14, 0), (17, 18, 1), (20, 21, 0)]):
        if up:
            xh = F.interpolate(xh, scale_factor=2, mode="bilinear", align_corners=True)
        uh = F.conv2d(xh, leaf[f"{hp}{ci}.weight"], None, padding=1)
        xh = F.relu(O._bn(uh, leaf, f"{hp}{bi}", train, False))
        inter[f"head{i}.u"], inter[f"head{i}.y"] = uh, xh
    head = F.conv2d(xh, leaf[hp + "23.weight"], leaf[hp + "23.bias"])
    inter["logits"] = head
    for t in inter.values():
        if t.requires_grad:
            t.retain_grad()

    print(f"== forward ({precision}, train_bn={train}, B={B}) ==")
    print("stem.u0", rel(nchw(S["stem"]["u0"]), inter["stem.u0"]))
    print("stem.a0", rel(nchw(S["stem"]["a0"]), inter["stem.a0"]))
    print("pool   ", rel(nchw(S["blocks"][0]["x_in"]), inter["pool"]))
    for i, Lb in enumerate(S["blocks"]):
        print(f"blk{i:02d} u1 {rel(nchw(Lb['u1']), inter[f'blk{i}.u1']):.2e} a1 {rel(nchw(Lb['a1']), inter[f'blk{i}.a1']):.2e} "
              f"u2 {rel(nchw(Lb['u2']), inter[f'blk{i}.u2']):.2e} out {rel(nchw(Lb['out']), inter[f'blk{i}.out']):.2e}")
    print("deconv.u", rel(nchw(S["deconv"]["u"]), inter["deconv.u"]), "deconv.y", rel(nchw(S["deconv"]["y"]), inter["deconv.y"]))
    for i, L in enumerate(S["head"]):
        print(f"head{i} u {rel(nchw(L['u']), inter[f'head{i}.u']):.2e} y {rel(nchw(L['y']), inter[f'head{i}.y']):.2e}")
    logits = S["logits"].view(B, 64, 64, 72)[..., :69].permute(0, 3, 1, 2)
    print("logits", rel(logits, inter["logits"]))
    coor_feat = torch.cat([head[:, 1:4], batch_cpu["roi_coord_2d"]], 1)
    region_sm = F.softmax(head[:, 5:], 1)
    xyz = (coor_feat[:, :3] - 0.5) * batch_cpu["roi_extent"].view(B, 3, 1, 1)
    pnp_in_ref = torch.cat([xyz, coor_feat[:, 3:], region_sm], 1)
    print("pnp_in", rel(nchw(S["pnp_in"])[:, :69], pnp_in_ref), "pad", float(nchw(S["pnp_in"])[:, 69:].abs().max()))
    xr = pnp_in_ref
    pq = "pnp_net."
    for i, (ci, gi) in enumerate([(0, 1), (3, 4), (6, 7)]):
        ur = F.conv2d(xr, leaf[f"{pq}features.{ci}.weight"], None, stride=2, padding=1)
        xr = F.relu(F.group_norm(ur, 32, leaf[f"{pq}features.{gi}.weight"], leaf[f"{pq}features.{gi}.bias"], 1e-5))
        print(f"pnp{i} u {rel(nchw(S['pnp'][i]['u']), ur):.2e} y {rel(nchw(S['pnp'][i]['y']), xr):.2e}")
    fl = xr.reshape(B, -1)
    h1 = F.leaky_relu(F.linear(fl, leaf[pq + "fc1.weight"], leaf[pq + "fc1.bias"]), 0.1)
    h2 = F.leaky_relu(F.linear(h1, leaf[pq + "fc2.weight"], leaf[pq + "fc2.bias"]), 0.1)
    print("h1", rel(S["h1"].float(), h1), "h2", rel(S["h2"].float(), h2))
    o = O.gdrn_forward(leaf, batch_cpu, train=train, do_loss=True)
    print("pred rot6d", rel(S["pred"][:, :6], o["rot6d"]), "t", rel(S["pred"][:, 6:9], o["pred_t_"]))
    print("rot", rel(res["rot"], o["rot"]), "trans", rel(res["trans"], o["trans"]))
    for i, n in enumerate(["loss_coor_x", "loss_coor_y", "loss_coor_z", "loss_mask", "loss_region", "loss_PM_R", "loss_centroid", "loss_z"]):
        print(n, float(res["losses"][i]), float(o["losses"][n]))

    # ---- backward
    sum(o["losses"].values()).backward()
    grads = eng.backward(torch.ones(8, device="cuda"))
    torch.cuda.synchronize()
    print("== backward: relative L2 error per parameter (reverse order) ==")
    for name, _p in reversed(eng.named_params):
        r = rel(grads[name], leaf[name].grad)
        flag = "  <<<<" if r > 5e-3 else ""
        print(f"{name:48s} {r:.2e} |g|={float(leaf[name].grad.norm()):.3e}{flag}")


if __name__ == "__main__":
    main()
