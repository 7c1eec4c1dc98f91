"""ncu target / timing of the fused tcgen05 policy+value forward (fm_mlp_tc5.cu) at rollout batch sizes.
usage: python tools/prof_mlp.py [rows] [iters]"""
import sys
import torch
sys.path.insert(0, ".")
from factory_marl_b200.ppo import PolicyValueNet
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
pol = PolicyValueNet(178, 16, False, seed=0, capacity=rows)
x = torch.randn(rows, 178, device="cuda")
for use in (1, 0):
    pol.lib.fm_ppo_use_tcgen05(use)
    for _ in range(3):
        pol.forward(x, deterministic=True)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        pol.forward(x, deterministic=True)
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / iters
    flop = rows * 2 * 2 * (192 * 128 + 128 * 128 + 128 * 16)      # two nets, K padded to 192, outputs padded to 16
    print("rows %d %s forward (policy + value, incl. head): %.3f ms, %.1f TFLOP/s bf16" % (rows, "tcgen05" if use else "mma.sync", ms, flop / ms * 1e-9))
pol.lib.fm_ppo_use_tcgen05(1)
