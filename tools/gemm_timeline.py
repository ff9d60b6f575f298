"""Decode the GEMM main-loop timeline of CTA 0 (yt8m_gemm.cuh debug stamps): when does each k-block land, when is its
ring slot free again, when do the epilogues run?   python tools/gemm_timeline.py [moe|fc]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("youtube-8m_b200", "tests", ""):
  sys.path.insert(0, os.path.join(ROOT, p))
import yt8m_native as nat

dev = "cuda:0"
what = sys.argv[1] if len(sys.argv) > 1 else "moe"
B, V = 256, 4716
if what == "moe":
  d_in, m = 1024, 2
  rows = nat.moe_packed_rows(V, m)
  wp = (torch.randn(rows, d_in, device=dev) * 0.02).to(torch.bfloat16)
  bp = torch.zeros(rows, device=dev)
  h = torch.randn(B, d_in, device=dev).to(torch.float16)
  wp16 = wp.to(torch.float16)
  fn = lambda: nat.moe_fwd(h, wp16, bp, V, m)
else:
  K, D = 64, 1152
  vl = torch.randn(B, K * D, device=dev).to(torch.float16)
  w = (torch.randn(1024, K * D, device=dev) * 0.01).to(torch.float16)
  fn = lambda: nat.linear(vl, w, act="relu6", out_f32=False, out_f16=True)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(3):
  fn()
flush.zero_()
buf = torch.zeros(128, dtype=torch.int64, device=dev)
nat.debug_set_timeline(buf)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
fn()
e1.record()
torch.cuda.synchronize()
nat.debug_set_timeline(None)
print("%s: %.1f us for the call" % (what, e0.elapsed_time(e1) * 1e3))
t = buf.cpu().tolist()
t0 = min(v for v in t if v > 0)
ev = []
for c in range(48):
  if t[c]:
    ev.append((t[c] - t0, "mma : k-block %d landed" % c))
  if t[64 + c]:
    ev.append((t[64 + c] - t0, "tma : slot for k-block %d free" % c))
for i in range(8):
  if t[48 + 2 * i]:
    ev.append((t[48 + 2 * i] - t0, "epi : tile %d accumulator complete" % i))
  if t[48 + 2 * i + 1]:
    ev.append((t[48 + 2 * i + 1] - t0, "epi : tile %d done" % i))
for ns, name in sorted(ev):
  print("%8.2f us  %s" % (ns / 1e3, name))
