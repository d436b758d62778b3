"""Cost of the BatchNorm-statistics epilogue: conv forward with / without the per-channel sum / sum-of-squares (CUDA graph loop)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from gdr_net_b200 import ops
from tools.bench_unpack import graph_time

SHAPES = [(64, 64, 64, 64, 64), (64, 32, 32, 128, 128), (64, 16, 16, 256, 256), (64, 8, 8, 512, 512), (64, 64, 64, 256, 256),
          (64, 32, 32, 256, 256)]
for (N, H, W, Cin, Cout) in SHAPES:
    x = ops.PT.from_float(torch.randn(N, H, W, Cin, device="cuda"), 1)
    w = torch.randn(Cout, Cin, 3, 3, device="cuda") * 0.05
    wp = ops.pack_conv_fwd(w, 1)
    out = ops.PT((N, H, W, Cout), 1)
    stats = torch.zeros(2, Cout, device="cuda")
    t1 = graph_time(lambda: ops.conv_fwd(x, wp, Cout, 3, 3, 1, 1, out=out, stats=stats))
    t0 = graph_time(lambda: ops.conv_fwd(x, wp, Cout, 3, 3, 1, 1, out=out))
    fl = 2.0 * N * H * W * Cout * Cin * 9
    print(f"{H}x{W} {Cin}->{Cout}: with stats {t1:7.1f} us ({fl / t1 / 1e6:6.0f} TF/s)   without {t0:7.1f} us ({fl / t0 / 1e6:6.0f} TF/s)")
