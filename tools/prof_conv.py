"""Run one conv shape a few times (for ncu --set full captures). usage: prof_conv.py N H W Cin Cout k stride pad [wgrad]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gdr_net_b200 import ops
N, H, W, Cin, Cout, k, stride, pad = [int(a) for a in sys.argv[1:9]]
wgrad = len(sys.argv) > 9
x = ops.PT((N, H, W, Cin), 1); x.buf.normal_()
w = torch.randn(Cout, Cin, k, k, device="cuda") * 0.05
wp = ops.pack_conv_fwd(w, 1)
out = ops.PT((N, H // stride, W // stride, (Cout + 63) // 64 * 64), 1)
stats = torch.zeros(2, Cout, device="cuda")
dy = ops.PT(out.shape, 1); dy.buf.normal_()
ws = ops.Workspace()
for _ in range(6):
    if wgrad:
        ops.conv_wgrad(dy, x, ws, Cout, k, k, stride, pad)
    else:
        ops.conv_fwd(x, wp, Cout, k, k, stride, pad, out=out, stats=stats)
torch.cuda.synchronize()
