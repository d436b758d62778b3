"""BASELINE.json configs[3] / SURVEY 8(d) config 4: the stand-alone Patch-PnP (`ConvPnPNet.forward`, reference
conv_pnp_net.py:111-157) at batch 512 for both input widths (nIn = 67: xyz + 64 regions; 69: + 2-D coords) against the
CPU oracle's `pnp_forward` on the same seeded maps / weights, within north_star's 1e-3."""
import pytest
import torch

from gdr_net_b200 import synth

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]


def _rel(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def _maps(B, c_feat, seed):
    g = synth._gen(seed, "pnp_maps")
    coor = torch.rand(B, c_feat, 64, 64, generator=g)                       # normalised xyz (+ 2-D coords) in [0, 1)
    region = torch.softmax(2.0 * torch.randn(B, 64, 64, 64, generator=g), dim=1)  # region attention = a softmax over 64
    ext = 0.05 + 0.25 * torch.rand(B, 3, generator=g)
    return coor, region, ext


@pytest.mark.parametrize("B", [512, 3])
@pytest.mark.parametrize("c_feat", [3, 5])
def test_patch_pnp_matches_oracle(B, c_feat):
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from gdr_net_b200.GDRN import ConvPnPNet
    from oracle import gdrn_oracle as O

    net = ConvPnPNet(nIn=c_feat + 64).cuda()
    sd = synth.seeded_state_dict({"pnp_net." + k: v for k, v in net.state_dict().items()}, seed=11)
    net.load_state_dict({k[len("pnp_net."):]: v for k, v in sd.items()})
    coor, region, ext = _maps(B, c_feat, seed=5 + c_feat)
    with torch.no_grad():
        rot_ref, t_ref = O.pnp_forward(coor, region, ext, sd)
    for precision, tol in (("fp32x3", 1e-3), ("half", 5e-2)):
        net.precision = precision
        for it in range(2):  # second call replays the CUDA graph
            rot, t = net(coor.cuda(), region.cuda(), ext.cuda())
        torch.cuda.synchronize()
        assert rot.shape == (B, 6) and t.shape == (B, 3)
        r1, r2 = _rel(rot, rot_ref), _rel(t, t_ref)
        print(f"[{precision}] Patch-PnP B={B} nIn={c_feat + 64}: rot rel-L2 {r1:.2e}, t rel-L2 {r2:.2e}")
        assert r1 < tol and r2 < tol, (precision, r1, r2)
    # the input tensor is not modified (the reference de-normalises xyz IN PLACE, conv_pnp_net.py:120-122; callers never reuse it)
    with pytest.raises(ValueError):
        net(coor[:, :2].cuda(), region.cuda(), ext.cuda())
