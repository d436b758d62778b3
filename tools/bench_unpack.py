"""Cost of the split-K reduction (unpack_wgrad) right after its producer (conv_wgrad), per layer shape, inside a CUDA graph
(so the workspace is L2-warm exactly as in the train step).  usage: python tools/bench_unpack.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from gdr_net_b200 import ops

SHAPES = [(64, 64, 64, 64, 64, 3, 1, 1), (64, 32, 32, 128, 128, 3, 1, 1), (64, 16, 16, 256, 256, 3, 1, 1), (64, 8, 8, 512, 512, 3, 1, 1),
          (64, 64, 64, 256, 256, 3, 1, 1), (64, 32, 32, 256, 256, 3, 1, 1), (64, 64, 64, 64, 128, 3, 2, 1)]


def graph_time(fn, reps=20, inner=10):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(inner):
            fn()
    g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (reps * inner) * 1e3


def main():
    planes = 1
    ws = ops.Workspace()
    for (N, H, W, Cin, Cout, k, stride, pad) in SHAPES:
        x = ops.PT.from_float(torch.randn(N, H, W, Cin, device="cuda"), planes)
        dy = ops.PT.from_float(torch.randn(N, H // stride, W // stride, Cout, device="cuda"), planes)
        grad = torch.zeros(Cout, Cin, k, k, device="cuda")
        info = {}

        def wg():
            info["r"] = ops.conv_wgrad(dy, x, ws, Cout, k, k, stride, pad)

        def both():
            buf, ks, kss = ops.conv_wgrad(dy, x, ws, Cout, k, k, stride, pad)
            ops.unpack_wgrad(buf, grad, Cout, Cin, k, k, Cin, ks, kss, Cin * k * k, k * k, k, 1)

        t_w = graph_time(wg)
        t_b = graph_time(both)
        ks = info["r"][1]
        print(f"{H}x{W} {Cin}->{Cout} k{k} s{stride}: ksplit {ks:3d}  wgrad {t_w:7.1f} us  wgrad+unpack {t_b:7.1f} us  unpack {t_b - t_w:6.1f} us")


if __name__ == "__main__":
    main()
