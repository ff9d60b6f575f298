"""NetVLAD v4 (one-pass cluster kernel) bring-up: parity vs the oracle and vs v3 on one shape, then timing.

    python tools/netvlad_v4_check.py B T D [f16|bf16] [time]
"""
import math, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("youtube-8m_b200", "tests", ""):
  sys.path.insert(0, os.path.join(ROOT, p))
import yt8m_native as nat
import synth
from oracle import yt8m_oracle as O

B, T, D = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
fmt = sys.argv[4] if len(sys.argv) > 4 else "f16"
do_time = len(sys.argv) > 5
K = 64
dev = "cuda:0"
g = torch.Generator().manual_seed(B + T + D)
x, nf, _ = synth.model_input(B, T, D, seed=6, min_frames=min(30, T))
nf[0] = T
cw = synth.normal((D, K), g, 4.0)
scale = 1.0 + 0.1 * torch.randn(K, generator=g)
shift = 0.1 * torch.randn(K, generator=g)
cw2 = synth.normal((D, K), g, 1 / math.sqrt(D))
xb = x.to(dev).to(torch.bfloat16)
cwp = nat.pack_transpose(cw.to(dev))
args = (xb, nf.to(dev), cwp, scale.to(dev), shift.to(dev), cw2.to(dev))
f16 = fmt == "f16"

def run(force_v3):
  nat.debug_set_flags(4096 if force_v3 else 0)
  out = nat.netvlad_fwd(*args, out_f16=f16, want_stats=True)
  torch.cuda.synchronize()
  nat.debug_set_flags(0)
  return out[0].float().cpu(), out[3].cpu()

v4, st4 = run(False)
print("v4 ran: B=%d T=%d D=%d fmt=%s finite=%s" % (B, T, D, fmt, bool(torch.isfinite(v4).all())), flush=True)
if B * T * D <= 40 * 300 * 1152:
  want = O.netvlad_pool(x, nf, cw, scale, shift, cw2)
  l2 = float((v4 - want).norm() / want.norm())
  mx = float((v4 - want).abs().max() / want.abs().max())
  print("  vs oracle: l2 %.3e  max/scale %.3e  row-norm[0] %.6f" % (l2, mx, float(v4[0].norm())))
  bad = (v4 - want).abs().reshape(B, D, K)
  per_b = bad.amax(dim=(1, 2))
  per_d = bad.amax(dim=(0, 2))
  per_k = bad.amax(dim=(0, 1))
  print("  worst videos", per_b.topk(min(4, B)).indices.tolist(), "worst d", per_d.topk(4).indices.tolist(), "worst k", per_k.topk(4).indices.tolist())
v3, st3 = run(True)
print("  vs v3: l2 %.3e  max %.3e   stats: asum %.3e ssq %.3e" % (float((v4 - v3).norm() / v3.norm()), float((v4 - v3).abs().max()),
      float((st4[:, :K] - st3[:, :K]).abs().max()), float(((st4[:, K:] - st3[:, K:]).abs() / st3[:, K:].abs().clamp_min(1e-6)).max())))
if do_time:
  for name, flag in (("v4", 0), ("v4 without the L2 prefetch", 65536), ("v3", 4096)):
    nat.debug_set_flags(flag)
    for _ in range(3):
      nat.netvlad_fwd(*args, out_f16=f16)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
      nat.netvlad_fwd(*args, out_f16=f16)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 50
    print("  %s: %.1f us/launch  (%.0f GB/s of frames)" % (name, us, B * T * D * 2 / us / 1e3))
  nat.debug_set_flags(0)
