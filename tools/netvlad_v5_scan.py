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
  fn = lambda: nat.netvlad_fwd_tiled(x, nf, cw, None, None, cw2, out_f16=True)
  for _ in range(3): fn()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(20): fn()
  e1.record(); torch.cuda.synchronize()
  us = e0.elapsed_time(e1) * 50
  tiles = B * ((nfv + 63) // 64)
  print("B=%4d nf=%3d: %7.1f us  (%d videos, %d tiles)" % (B, nfv, us, B, tiles), flush=True)
for B in (1, 33, 37, 66, 132, 264):
  t(B, 300)
for B in (33, 66, 132, 264, 528):
  t(B, 64)
for nfv in (64, 128, 192, 256, 300):
  t(132, nfv)
