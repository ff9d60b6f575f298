"""NetVLAD v5 (four-CTA-cluster one-pass kernel) bring-up: parity vs the oracle and vs its predecessors on one shape, timing.

    python tools/netvlad_v5_check.py B T D [f16|bf16] [time] [K]
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
do_time = len(sys.argv) > 5 and sys.argv[5] == "time"
K = int(sys.argv[6]) if len(sys.argv) > 6 else 64
dev = "cuda:0"
g = torch.Generator().manual_seed(B + T + D)
x, nf, _ = synth.model_input(B, T, D, seed=6, min_frames=min(30, T))
nf[0] = T
if B > 3:
  nf[1], nf[2] = 0, 1                        # an empty and a one-frame video
cw = synth.normal((D, K), g, 4.0)
scale = 1.0 + 0.1 * torch.randn(K, generator=g)
shift = 0.1 * torch.randn(K, generator=g)
cw2 = synth.normal((D, K), g, 1 / math.sqrt(D))
xb = x.to(dev).to(torch.bfloat16)
cwp = nat.pack_transpose(cw.to(dev))
args = (xb, nf.to(dev), cwp, scale.to(dev), shift.to(dev), cw2.to(dev))
f16 = fmt == "f16"

def run(flags):
  nat.debug_set_flags(flags)
  out = nat.netvlad_fwd(*args, out_f16=f16, want_stats=True)
  torch.cuda.synchronize()
  nat.debug_set_flags(0)
  return out[0].float().cpu(), out[3].cpu()

idx8 = nat.netvlad_tiled_index(D, K, 8, dev)
idx4 = nat.netvlad_tiled_index(D, K, 4, dev)
c2t = cw2.to(dev).reshape(-1)[idx4].contiguous()
targs = (xb, nf.to(dev), cwp, scale.to(dev), shift.to(dev), c2t)

def run_tiled(two=True):
  out, st = nat.netvlad_fwd_tiled(*targs, out_f16=f16, want_stats=True, two_kernels=two)
  torch.cuda.synchronize()
  std = torch.empty_like(out)
  std[:, idx8] = out                          # tiled position p holds row-major element idx8[p]
  return std.float().cpu(), st.cpu()

v5, st5 = run_tiled()
one, st1 = run_tiled(False)
print("  two-kernel vs one-pass: max %.3e  stats %.3e" % (float((v5 - one).abs().max()), float((st5 - st1).abs().max())))
if B <= 8:
  u5, _ = run(1 << 20)                        # the same kernel with row-major epilogue accesses
  print("  tiled vs row-major epilogue: max %.3e" % float((v5 - u5).abs().max()))
print("v5 ran: B=%d T=%d D=%d K=%d fmt=%s finite=%s" % (B, T, D, K, fmt, bool(torch.isfinite(v5).all())), flush=True)
if B * T * D <= 40 * 300 * 1152:
  want = O.netvlad_pool(x, nf, cw, scale, shift, cw2)
  l2 = float((v5 - want).norm() / want.norm())
  mx = float((v5 - want).abs().max() / want.abs().max())
  print("  vs oracle: l2 %.3e  max/scale %.3e  row-norm[0] %.6f" % (l2, mx, float(v5[0].norm())))
  bad = (v5 - want).abs().reshape(B, D, K)
  print("  worst videos", bad.amax(dim=(1, 2)).topk(min(4, B)).indices.tolist(), "worst d", bad.amax(dim=(0, 2)).topk(4).indices.tolist(),
        "worst k", bad.amax(dim=(0, 1)).topk(4).indices.tolist())
old, sto = run(0)                             # the previous kernel for this shape (v4 / generic)
print("  vs previous kernel: l2 %.3e  max %.3e   stats: asum %.3e ssq %.3e" % (
    float((v5 - old).norm() / old.norm()), float((v5 - old).abs().max()), float((st5[:, :K] - sto[:, :K]).abs().max()),
    float(((st5[:, K:] - sto[:, K:]).abs() / sto[:, K:].abs().clamp_min(1e-6)).max())))
if do_time:
  real = int(nf.clamp(0, T).sum())
  for name, fn in (("two kernels (v6)", lambda: nat.netvlad_fwd_tiled(*targs, out_f16=f16)),
                   ("one pass (v5)", lambda: nat.netvlad_fwd_tiled(*targs, out_f16=f16, two_kernels=False)), ("previous", lambda: nat.netvlad_fwd(*args, out_f16=f16))):
    for _ in range(3):
      fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
      fn()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 50
    alg = real * D * 2 + B * D * K * 2
    print("  %s: %.1f us/launch  (%.0f GB/s algorithmic: %d real rows + descriptors)" % (name, us, alg / us / 1e3, real))
  nat.debug_set_flags(0)
