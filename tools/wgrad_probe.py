"""The hidden layer's weight-gradient GEMM at BASELINE config 2 shapes, alone: dWfc^T[1024, 73728] = dpre^T . vlad over B = 256 rows.

    python tools/wgrad_probe.py [M] [N] [Kb]        (timing with CUDA events over a captured graph of 10 launches)"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("youtube-8m_b200", "tests", ""):
  sys.path.insert(0, os.path.join(ROOT, p))
import yt8m_native as nat

M = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
N = int(sys.argv[2]) if len(sys.argv) > 2 else 73728
Kb = int(sys.argv[3]) if len(sys.argv) > 3 else 256
dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(0)
a = torch.randn(Kb, M, device=dev, generator=g)
a_hi, a_lo = nat.split_bf16(a)
b = torch.randn(Kb, N, device=dev, generator=g).to(torch.bfloat16)
out = torch.empty(M, N, device=dev)
nat.wgrad(a_hi, a_lo, b, M, N, out=out)
torch.cuda.synchronize()
want = (a_hi.float() + a_lo.float()).t() @ b.float()
print("max err vs fp32 matmul: %.3e (scale %.1f)" % (float((out - want).abs().max()), float(want.abs().max())))
del want
for name, lo in (("hi+lo", a_lo), ("hi only", None)):
  for _ in range(3):
    nat.wgrad(a_hi, lo, b, M, N, out=out)
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(10):
    nat.wgrad(a_hi, lo, b, M, N, out=out)
  e1.record()
  torch.cuda.synchronize()
  us = e0.elapsed_time(e1) * 100
  print("wgrad %s M=%d N=%d Kb=%d: %.1f us  (%.0f GB/s of fp32 output, %.0f TFLOP/s)" % (
      name, M, N, Kb, us, M * N * 4 / us / 1e3, 2.0 * M * N * Kb * (2 if lo is not None else 1) / us / 1e6))
