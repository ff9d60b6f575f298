"""Decode the NetVLAD v5 phase timeline (cluster 0 / rank 0, first 3 videos): where do the microseconds go?"""
import math, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "youtube-8m_b200"))
import yt8m_native as nat

dev = "cuda:0"
B, T, D, K = int(sys.argv[1]) if len(sys.argv) > 1 else 256, 300, 1152, 64
x = torch.randn(B, T, D, device=dev).to(torch.bfloat16)
nf = torch.full((B,), T, dtype=torch.int32, device=dev)
cw = (torch.randn(K, D, device=dev) / math.sqrt(D)).to(torch.bfloat16)
cw2 = torch.randn(D, K, device=dev) / math.sqrt(D)
for _ in range(3):
  nat.netvlad_fwd_tiled(x, nf, cw, None, None, cw2, out_f16=True)
buf = torch.zeros(384, dtype=torch.int64, device=dev)
nat.debug_set_timeline(buf)
nat.netvlad_fwd_tiled(x, nf, cw, None, None, cw2, out_f16=True)
torch.cuda.synchronize()
nat.debug_set_timeline(None)
t = buf.cpu().tolist()
pos = [v for v in t if v > 0]
if not pos:
  raise SystemExit("no stamps: build with YT8M_NVCC_EXTRA=-DYT8M_V5_TIMELINE")
t0 = min(pos)
names = {}
for i in range(8):
  names[i] = "mma0: tile %d landed" % i
  names[8 + i] = "exch: tile %d S ready" % i
  names[16 + i] = "own : tile %d partials landed" % i
  names[24 + i] = "own : tile %d assignment sent" % i
  names[32 + i] = "mma1: tile %d assignment landed" % i
  names[40 + i] = "mma1: tile %d issued" % i
  names[56 + i] = "prod: tile %d slot free" % i
names.update({52: "epi : V in registers", 53: "epi : a_sum ready", 54: "epi : residual done", 55: "epi : ssq reduced", 64: "epi : peers' norms landed"})
names.update({48: "epi : video complete", 49: "epi : pass 1 done", 50: "epi : norms exchanged", 51: "epi : pass 2 done"})
for it in range(3):
  ev = [(t[it * 128 + s] - t0, names[s]) for s in names if t[it * 128 + s] > 0]
  for ns, n in sorted(ev):
    print("video %d  %8.2f us  %s" % (it, ns / 1e3, n))
  print()
