"""Timing scan of the tiled NetVLAD kernel: how the launch time splits into start-up, per-tile and per-video cost."""
import math, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "youtube-8m_b200"))
import yt8m_native as nat
dev = "cuda:0"
T, D, K = 300, 1152, 64
cw = (torch.randn(K, D, device=dev) / math.sqrt(D)).to(torch.bfloat16)
cw2 = torch.randn(D * K, device=dev) / math.sqrt(D)
def t(B, nfv):
  x = torch.randn(B, T, D, device=dev).to(torch.bfloat16)
  nf = torch.full((B,), nfv, dtype=torch.int32, device=dev)
  fn = lambda: nat.netvlad_fwd_tiled(x, nf, cw, None, None, cw2, out_f16=True, two_kernels=TWO)
  for _ in range(3): fn()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(20): fn()
  e1.record(); torch.cuda.synchronize()
  us = e0.elapsed_time(e1) * 50
  parts = []
  if TWO:
    for flag in (1 << 22, 1 << 23):
      nat.debug_set_flags(flag)
      for _ in range(2): fn()
      e0.record()
      for _ in range(20): fn()
      e1.record(); torch.cuda.synchronize()
      parts.append(e0.elapsed_time(e1) * 50)
    nat.debug_set_flags(0)
  tiles = B * ((nfv + 63) // 64)
  print("B=%4d nf=%3d: %7.1f us  (%d videos, %d tiles)  assign %s aggregate %s" % (B, nfv, us, B, tiles, *(["%.1f" % p for p in parts] or ["-", "-"])), flush=True)
TWO = len(sys.argv) < 2 or sys.argv[1] != "one"
print("two kernels" if TWO else "one pass")
# the bench batch's length distribution
import random
def tb(B):
  g = torch.Generator().manual_seed(8)
  x = torch.randn(B, T, D, device=dev).to(torch.bfloat16)
  nf = torch.randint(30, T + 1, (B,), generator=g, dtype=torch.int32).to(dev)
  fn = lambda: nat.netvlad_fwd_tiled(x, nf, cw, None, None, cw2, out_f16=True, two_kernels=TWO)
  res = []
  K2 = 1 << 23
  for flag in (0, 1 << 22, K2, K2 | (1 << 24), K2 | (1 << 25), K2 | (1 << 25) | (1 << 26), K2 | (1 << 25) | (1 << 26) | (1 << 27), K2 | (1 << 28)):
    nat.debug_set_flags(flag)
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): fn()
    e1.record(); torch.cuda.synchronize()
    res.append(e0.elapsed_time(e1) * 50)
  nat.debug_set_flags(0)
  print("B=%4d nf~U{30..300}: %.1f us  assign %.1f aggregate %.1f  aggregate: no stores %.1f | no pass 2 %.1f | + no residual %.1f | + no norm exchange %.1f | hand-offs only %.1f" % (B, *res), flush=True)
tb(256); tb(512)
