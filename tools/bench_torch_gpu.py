"""Comparison arm: the reference algorithm as plain PyTorch ops (cuDNN / cuBLAS / ATen) on the SAME GPU -- i.e. what the
reference's own nn.Modules dispatch to on this box (the reference package itself is not installable on the GPU box).
Runs the oracle restatement (pinned bit-exact to the reference on CPU) on cuda with cudnn.benchmark=True
(reference configs/_base_/common_base.py:16) and TF32 convolutions allowed (PyTorch default), fwd+bwd, batch 64."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gdr_net_b200 import synth
from oracle import fixtures, gdrn_oracle as O

def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    amp = "--amp" in sys.argv
    torch.backends.cudnn.benchmark = True
    dev = torch.device("cuda")
    sd = synth.seeded_state_dict(fixtures.template_from_manifest(), 0)
    leaf = {k: (v.to(dev).requires_grad_(v.dtype.is_floating_point and "running" not in k)) for k, v in O.leaf_state_dict(sd, requires_grad=False).items()}
    batch = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in synth.make_batch(B, seed=100).items()}
    # the oracle's logging helper does numpy work on the host exactly like the reference (model_utils.py:40-52)
    def step():
        for v in leaf.values():
            if v.requires_grad:
                v.grad = None
        with torch.autocast("cuda", dtype=torch.float16, enabled=amp):
            o = O.gdrn_forward(leaf, batch, train=True, do_loss=True)
        total = sum(o["losses"].values())
        total.backward()
        return float(total)
    for _ in range(8):
        step()
    torch.cuda.synchronize()
    n = 15
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    out = dict(impl="torch/cuDNN restatement of the reference on the same GPU", batch=B, amp_fp16=amp, ms_per_step=round(ms, 3),
               crops_per_s=round(B / ms * 1e3, 1), tf32=torch.backends.cudnn.allow_tf32, cudnn=torch.backends.cudnn.version(),
               torch=torch.__version__)
    print(json.dumps(out))
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/torch_gpu_baseline%s.json" % ("_amp" if amp else ""), "w") as f:
        json.dump(out, f)

if __name__ == "__main__":
    main()
