"""Micro-benchmark of the elementwise kernels at the layer sizes of the B=64 step (CUDA events, back-to-back launches)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gdr_net_b200 import ops
from gdr_net_b200.capi import C
from gdr_net_b200.ops import _stream

def timeit(fn, iters=50, warmup=5):
    for _ in range(warmup): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / iters * 1e3  # us

B = 64
for name, H, Cc in [("stem 128x128x64", 128, 64), ("layer1 64x64x64", 64, 64), ("layer2 32x32x128", 32, 128), ("layer3 16x16x256", 16, 256),
                    ("layer4 8x8x512", 8, 512), ("head 64x64x256", 64, 256)]:
    u = ops.PT((B, H, H, Cc), 1); u.buf.normal_()
    y = ops.like(u); g = ops.like(u); g.buf.normal_()
    gamma = torch.ones(Cc, device="cuda"); beta = torch.zeros(Cc, device="cuda")
    rm = torch.zeros(Cc, device="cuda"); rv = torch.ones(Cc, device="cuda")
    stats = torch.rand(2, Cc, device="cuda") * 1000; stats[1] += stats[0] ** 2 / (B * H * H) * 1.5
    mean = torch.zeros(Cc, device="cuda"); invstd = torch.ones(Cc, device="cuda"); sums = torch.zeros(2, Cc, device="cuda")
    dg = torch.zeros(Cc, device="cuda"); db = torch.zeros(Cc, device="cuda")
    rows = B * H * H
    nbytes = u.numel() * 2
    t_fwd = timeit(lambda: C.gdrn_bn_fwd(u.hi_ptr, None, None, None, y.hi_ptr, None, stats.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                                        rm.data_ptr(), rv.data_ptr(), mean.data_ptr(), invstd.data_ptr(), None, rows, Cc, 1e-5, 0.1, 1, 1, _stream()))
    scale = torch.ones(Cc, device="cuda"); shift = torch.zeros(Cc, device="cuda")
    t_act = timeit(lambda: ops.bn_act(u, scale, shift, True, out=y))
    t_bwd = timeit(lambda: ops.bn_bwd(g, None, y, u, mean, invstd, gamma, sums, dg, db, True))
    t_copy = timeit(lambda: y.buf.copy_(u.buf))
    print(f"{name:20s} {nbytes/1e6:7.1f} MB/tensor | bn_fwd {t_fwd:7.1f} us ({2*nbytes/t_fwd/1e6:5.2f} TB/s) | bn_act {t_act:7.1f} us | "
          f"bn_bwd(reduce+apply) {t_bwd:7.1f} us ({7*nbytes/t_bwd/1e6:5.2f} TB/s) | torch copy {t_copy:7.1f} us ({2*nbytes/t_copy/1e6:5.2f} TB/s)")

# gather kernels at their step sizes
x = ops.PT((B, 128, 128, 64), 1); x.buf.normal_()
t_mp = timeit(lambda: ops.maxpool_fwd(x, want_arg=True))
yp, arg = ops.maxpool_fwd(x, want_arg=True)
gp = ops.like(yp); gp.buf.normal_()
t_mpb = timeit(lambda: ops.maxpool_bwd(arg, gp))
print(f"maxpool 128x128x64: fwd {t_mp:7.1f} us  bwd {t_mpb:7.1f} us   (ideal ~34 / ~35 us at 5.5 TB/s)")
for H in (16, 32):
    xs = ops.PT((B, H, H, 256), 1); xs.buf.normal_()
    t_uf = timeit(lambda: ops.upsample2x_fwd(xs))
    gs = ops.PT((B, 2 * H, 2 * H, 256), 1); gs.buf.normal_()
    t_ub = timeit(lambda: ops.upsample2x_bwd(gs))
    mb = xs.numel() * 2 * 5 / 1e6
    print(f"upsample {H}->{2*H} x256: fwd {t_uf:7.1f} us  bwd {t_ub:7.1f} us   ({mb:.0f} MB each way, ideal {mb/5.5:.0f} us)")
xi = torch.randn(B, 3, 256, 256, device="cuda")
a_col = ops.PT((B * 128 * 128, 192), 1)
t_im = timeit(lambda: C.gdrn_stem_im2col(xi.data_ptr(), a_col.hi_ptr, a_col.lo_ptr, B, 256, 256, _stream()))
print(f"stem im2col: {t_im:7.1f} us  (writes {a_col.numel()*2/1e6:.0f} MB, ideal {a_col.numel()*2/5.5e6:.0f} us)")
