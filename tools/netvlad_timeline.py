"""Decode the NetVLAD kernel's phase timeline (CTA 0, first 4 videos): where do the microseconds go?"""
import math, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "youtube-8m_b200"))
import yt8m_native as nat

dev = "cuda:0"
B, T, D = int(sys.argv[1]) if len(sys.argv) > 1 else 256, 300, 1152
K = int(sys.argv[2]) if len(sys.argv) > 2 else 64
FLAGS = int(sys.argv[3]) if len(sys.argv) > 3 else 0
LO = (int(sys.argv[4]) if len(sys.argv) > 4 else 1) != 0
x = torch.randn(B, T, D, device=dev).to(torch.bfloat16)
nf = torch.full((B,), T, dtype=torch.int32, device=dev)
cw = (torch.randn(K, D, device=dev) / math.sqrt(D)).to(torch.bfloat16)
cw2 = torch.randn(D, K, device=dev) / math.sqrt(D)
for _ in range(3):
  nat.netvlad_fwd(x, nf, cw, None, None, cw2, want_lo=LO)
nat.debug_set_flags(FLAGS)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
  nat.netvlad_fwd(x, nf, cw, None, None, cw2, want_lo=LO)
e1.record()
torch.cuda.synchronize()
print('flags=%d lo=%d  kernel %.1f us' % (FLAGS, LO, e0.elapsed_time(e1) * 200))
buf = torch.zeros(128, dtype=torch.int64, device=dev)
nat.debug_set_timeline(buf)
nat.netvlad_fwd(x, nf, cw, None, None, cw2, want_lo=LO)
torch.cuda.synchronize()
nat.debug_set_timeline(None)
nat.debug_set_flags(0)
t = buf.cpu().tolist()
names = {0: "prod: p0 start", 1: "prod: p0 loads issued", 2: "prod: p1 loads issued", 8: "mma: iter start", 9: "mma: first x tile landed",
         10: "mma: s_full committed", 11: "mma: a_ready seen", 12: "mma: last group committed", 16: "epi: iter start",
         17: "epi: s_full seen", 18: "epi: softmax done", 19: "epi: group0 ready", 20: "epi: group1 ready", 21: "epi: group2 ready",
         22: "epi: group3 ready", 23: "epi: group4 ready", 26: "epi: epilogue done", 3: "epi: group0 drained", 4: "epi: group1 drained", 5: "epi: group2 drained", 6: "epi: group3 drained",
         7: "epi: group4 drained", 13: "mma: group0 issued", 14: "mma: group1 issued", 15: "mma: group2 issued", 25: "mma: group3 issued",
         28: "mma: group4 issued", 27: "epi: rescale done"}
t0 = min(v for v in t if v > 0)
for it in range(4):
  ev = [(t[it * 32 + s] - t0, names[s]) for s in names if t[it * 32 + s] > 0]
  for ns, n in sorted(ev):
    print("video %d  %8.2f us  %s" % (it, ns / 1e3, n))
  print()
