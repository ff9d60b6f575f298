"""Decode the persistent-LSTM step timeline (CTA 0, steps 8..15 of the first layer): where do the microseconds go?"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("youtube-8m_b200", "tests", ""):
  sys.path.insert(0, os.path.join(ROOT, p))
import yt8m_native as nat
import synth

dev = "cuda:0"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
T, D, H = 300, 1152, 1024
g = torch.Generator().manual_seed(3)
x = torch.randn(B, T, D, device=dev).to(torch.bfloat16)
nf = torch.full((B,), T, dtype=torch.int32, device=dev)
packed = [nat.lstm_pack(synth.xavier((D + H, 4 * H), g).to(dev), torch.zeros(4 * H, device=dev), D, H)]
for _ in range(2):
  nat.lstm_fwd(x, nf, [packed[0][0]], [packed[0][1]], H)
buf = torch.zeros(128, dtype=torch.int64, device=dev)
nat.debug_set_timeline(buf)
nat.lstm_fwd(x, nf, [packed[0][0]], [packed[0][1]], H)
torch.cuda.synchronize()
nat.debug_set_timeline(None)
t = buf.cpu().tolist()
names = ["tma: grid barrier passed", "mma: h tiles landed", "epi: accumulator complete", "epi: partials pushed",
         "epi: peers' partials landed", "epi: h_t stored", "epi: arrived on the grid barrier"]
t0 = min(v for v in t if v > 0)
for st in range(8):
  ev = sorted((t[st * 16 + s] - t0, names[s]) for s in range(7) if t[st * 16 + s] > 0)
  for ns, n in ev:
    print("step %2d  %8.2f us  %s" % (st + 8, ns / 1e3, n))
  print()
