"""Small invocations of the hand-rolled-protocol kernels (mbarrier / DSMEM / grid barrier) for compute-sanitizer:
the tiled NetVLAD kernels (one-pass v5, two-kernel v6), the one-pass v4, the tcgen05 GEMM (linear split-K, MoE head), the persistent
LSTM recurrence.  Every call is checked against the oracle as well, so a sanitizer-clean run is also a correct one."""
import math, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("youtube-8m_b200", "tests", ""):
  sys.path.insert(0, os.path.join(ROOT, p))
import yt8m_native as nat
import synth
from oracle import yt8m_oracle as O

dev = "cuda:0"
which = sys.argv[1] if len(sys.argv) > 1 else "all"
g = torch.Generator().manual_seed(3)

def rel(a, b):
  return float((a - b).norm() / b.norm().clamp_min(1e-9))

if which in ("all", "netvlad"):
  B, T, D, K = 9, 130, 256, 64
  x, nf, _ = synth.model_input(B, T, D, seed=6, min_frames=3)
  nf[1] = 0
  cw, cw2 = synth.normal((D, K), g, 4.0), torch.randn(D, K, generator=g) / math.sqrt(D)
  want = O.netvlad_pool(x, nf, cw, torch.ones(K), torch.zeros(K), cw2)
  xb, cwp = x.to(dev).to(torch.bfloat16), nat.pack_transpose(cw.to(dev))
  idx8, idx4 = nat.netvlad_tiled_index(D, K, 8, dev), nat.netvlad_tiled_index(D, K, 4, dev)
  c2t = cw2.to(dev).reshape(-1)[idx4].contiguous()
  for two in (False, True):
    out = nat.netvlad_fwd_tiled(xb, nf.to(dev), cwp, None, None, c2t, out_f16=True, two_kernels=two)
    got = torch.empty_like(out); got[:, idx8] = out
    torch.cuda.synchronize()
    print("netvlad tiled two_kernels=%s: rel %.2e" % (two, rel(got.float().cpu(), want)))
  out = nat.netvlad_fwd(xb, nf.to(dev), cwp, None, None, cw2.to(dev), out_f16=True)[0]
  torch.cuda.synchronize()
  print("netvlad v4: rel %.2e" % rel(out.float().cpu(), want))

if which in ("all", "gemm"):
  M, N, Kd = 70, 96, 4096
  a = synth.bf16r(torch.randn(M, Kd, generator=g))
  w = synth.bf16r(torch.randn(Kd, N, generator=g) / math.sqrt(Kd))
  got = nat.linear(a.to(dev).to(torch.bfloat16), nat.pack_transpose(w.to(dev)), n=N, k=Kd)["f32"]
  torch.cuda.synchronize()
  print("linear (split-K): rel %.2e" % rel(got.cpu(), a @ w))
  Bm, Dm, V, Mx = 40, 256, 300, 2
  xm = synth.bf16r(torch.randn(Bm, Dm, generator=g) / math.sqrt(Dm))
  gw, ew, eb = synth.xavier((Dm, V * (Mx + 1)), g, 6.0), synth.xavier((Dm, V * Mx), g, 6.0), 0.1 * torch.randn(V * Mx, generator=g)
  wp, bp = nat.moe_pack(gw.to(dev), ew.to(dev), eb.to(dev), V, Mx)
  got = nat.moe_fwd(xm.to(dev).to(torch.bfloat16), wp, bp, V, Mx)
  torch.cuda.synchronize()
  print("moe head: rel %.2e" % rel(got.cpu(), O.moe_model(xm, gw, ew, eb, V, Mx)))

if which in ("all", "lstm"):
  B, T, D, H, L = 5, 12, 64, 256, 2
  x, nf, _ = synth.model_input(B, T, D, seed=9, min_frames=2)
  ws = []
  for l in range(L):
    ind = D if l == 0 else H
    ws.append((synth.xavier((ind + H, 4 * H), g), 0.1 * torch.randn(4 * H, generator=g)))
  packs = [nat.lstm_pack(w.to(dev), b.to(dev), D if l == 0 else H, H) for l, (w, b) in enumerate(ws)]
  state, _, _ = nat.lstm_fwd(x.to(dev).to(torch.bfloat16), nf.to(dev), [p[0] for p in packs], [p[1] for p in packs], H)
  torch.cuda.synchronize()
  _, st = O.dynamic_rnn_lstm(x, nf, ws)
  print("lstm (persistent recurrence): rel %.2e" % rel(state.cpu(), O.lstm_model_state(st)))

if which in ("all", "train"):
  # the training step's new kernels: tensor-core assignment backward, vectorised norm backward, ticketed dcw2 / column sums,
  # in-place split-K accumulation of the reverse LSTM recurrence -- through the trainers (two steps: the self-cleaning
  # tickets and the zero-on-read accumulator are exercised on their second use)
  import yt8m_trainer
  Bf, T, Df, K, H, Vf, M = 6, 140, 256, 64, 128, 200, 2
  xf, nff, _ = synth.model_input(Bf, T, Df, seed=5, min_frames=1)
  nff[0] = 0
  yf = synth.labels(Bf, Vf, seed=5, per_video=3.4)
  sdf = {"cluster_weights": synth.normal((Df, K), g, 4.0), "cluster_biases": 0.1 * torch.randn(K, generator=g),
         "cluster_weights2": synth.normal((Df, K), g, 1 / math.sqrt(Df)), "hidden1_weights": synth.normal((K * Df, H), g, 12.0 / math.sqrt(K)),
         "hidden1_biases": 0.1 * torch.randn(H, generator=g), "gates/weights": synth.xavier((H, Vf * (M + 1)), g, 2.0),
         "experts/weights": synth.xavier((H, Vf * M), g, 2.0), "experts/biases": 0.1 * torch.randn(Vf * M, generator=g)}
  t = yt8m_trainer.NetVLADTrainer(Df, clusters=K, hidden=H, vocab=Vf, mixtures=M, device=torch.device(dev))
  t.import_state(sdf)
  for _ in range(2):
    t.step(xf.to(dev).to(torch.bfloat16), nff.to(dev), yf.to(dev))
  torch.cuda.synchronize()
  print("NetVLAD trainer: 2 steps, params finite %s" % bool(torch.isfinite(t.param).all()))
  Bl, Tl, Dl, Hl, Ll = 4, 200, 64, 256, 2
  xl, nfl, _ = synth.model_input(Bl, Tl, Dl, seed=6, min_frames=2)
  yl = synth.labels(Bl, Vf, seed=6, per_video=3.4)
  sdl = {"gates/weights": synth.xavier((Ll * 2 * Hl, Vf * (M + 1)), g, 2.0), "experts/weights": synth.xavier((Ll * 2 * Hl, Vf * M), g, 2.0),
         "experts/biases": 0.1 * torch.randn(Vf * M, generator=g)}
  for l in range(Ll):
    sdl[yt8m_trainer.LstmTrainer.SCOPE % l + "/weights"] = synth.xavier(((Dl if l == 0 else Hl) + Hl, 4 * Hl), g, 2.0)
    sdl[yt8m_trainer.LstmTrainer.SCOPE % l + "/biases"] = synth.bf16r(0.1 * torch.randn(4 * Hl, generator=g))
  tl = yt8m_trainer.LstmTrainer(Dl, hidden=Hl, layers=Ll, vocab=Vf, mixtures=M, device=torch.device(dev))
  tl.import_state(sdl)
  for _ in range(2):
    tl.step(xl.to(dev).to(torch.bfloat16), nfl.to(dev), yl.to(dev))
  torch.cuda.synchronize()
  print("LSTM trainer: 2 steps, params finite %s" % bool(torch.isfinite(tl.param).all()))
