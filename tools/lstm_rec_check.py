"""Persistent LSTM recurrence (yt8m_lstm_rec.cu) bring-up: parity vs the oracle and vs the per-step path, then timing.

    python tools/lstm_rec_check.py B T D H L [time]
"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("youtube-8m_b200", "tests", ""):
  sys.path.insert(0, os.path.join(ROOT, p))
import yt8m_native as nat
import synth
from oracle import yt8m_oracle as O

B, T, D, H, L = (int(v) for v in sys.argv[1:6])
do_time = len(sys.argv) > 6
dev = "cuda:0"
g = torch.Generator().manual_seed(B * 7 + T + H)
x = synth.bf16r(torch.randn(B, T, D, generator=g) * 0.5)
nf = torch.randint(1, T + 1, (B,), generator=g, dtype=torch.int32)
nf[0] = T
ws = []
for l in range(L):
  i = D if l == 0 else H
  ws.append((synth.xavier((i + H, 4 * H), g, 2.0), 0.1 * torch.randn(4 * H, generator=g)))
packed = [nat.lstm_pack(w.to(dev), bb.to(dev), D if l == 0 else H, H) for l, (w, bb) in enumerate(ws)]
xb = x.to(dev).to(torch.bfloat16)
nfd = nf.to(dev)


def run(per_step, seq=True):
  nat.debug_set_flags(8192 if per_step else 0)
  out = nat.lstm_fwd(xb, nfd, [p[0] for p in packed], [p[1] for p in packed], H, want_seq=seq, want_seq_bf16=seq)
  torch.cuda.synchronize()
  nat.debug_set_flags(0)
  return out


def rel(a, b):
  return float((a.float().cpu() - b).norm() / b.norm().clamp_min(1e-30))


st, seq, seq_bf = run(False)
print("rec ran: B=%d T=%d D=%d H=%d L=%d finite=%s" % (B, T, D, H, L, bool(torch.isfinite(st).all())), flush=True)
if B * T * (D + H) * H * L <= 64 * 300 * 2176 * 1024 * 2:
  outs, states = O.dynamic_rnn_lstm(x, nf, ws)
  want = O.lstm_model_state(states)
  print("  vs oracle: state %.3e  seq %.3e  seq_bf %.3e" % (rel(st, want), rel(seq, outs), rel(seq_bf, outs)))
  zero_ok = all(float(seq[b, int(nf[b]):].abs().max()) == 0.0 for b in range(B) if int(nf[b]) < T)
  print("  outputs past num_frames are zero:", zero_ok)
st0, seq0, _ = run(True)
print("  vs per-step path: state %.3e  seq %.3e" % (rel(st, st0.float().cpu()), rel(seq, seq0.float().cpu())))
if do_time:
  for name, flag in (("persistent", False), ("per-step", True)):
    for _ in range(2):
      run(flag, seq=False)
    nat.debug_set_flags(8192 if flag else 0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 5
    e0.record()
    for _ in range(n):
      nat.lstm_fwd(xb, nfd, [p[0] for p in packed], [p[1] for p in packed], H)
    e1.record()
    torch.cuda.synchronize()
    nat.debug_set_flags(0)
    ms = e0.elapsed_time(e1) / n
    print("  %s: %.3f ms/forward  (%.0f videos/s, %.2f us per step and layer)" % (name, ms, B / ms * 1e3, ms * 1e3 / (T * L)))
