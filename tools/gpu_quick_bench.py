"""Per-kernel timings at the BASELINE configs[1] shapes (CUDA events, L2 flushed between iterations)."""
import json
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "youtube-8m_b200"))
sys.path.insert(0, ROOT)
import yt8m_native as nat  # noqa: E402

dev = "cuda:0"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, iters=10, warm=3):
  for _ in range(warm):
    fn()
  ts = []
  for _ in range(iters):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
  ts.sort()
  return ts[len(ts) // 2]


res = {}
B, T, D, K, V = 256, 300, 1152, 64, 4716
x = torch.randn(B, T, D, device=dev).to(torch.bfloat16)
nf = torch.full((B,), T, dtype=torch.int32, device=dev)
cw = (torch.randn(K, D, device=dev) / math.sqrt(D)).to(torch.bfloat16)
cw2 = torch.randn(D, K, device=dev) / math.sqrt(D)
ms = timeit(lambda: nat.netvlad_fwd(x, nf, cw, None, None, cw2))
res["netvlad_k64_b256"] = {"ms": ms, "GBps_x": B * T * D * 2 / ms / 1e6, "videos_per_s": B / ms * 1e3}
ms = timeit(lambda: nat.netvlad_fwd(x, nf, cw, None, None, cw2, want_lo=True))
res["netvlad_k64_b256_hilo"] = {"ms": ms}
ms = timeit(lambda: nat.netvlad_fwd(x, nf, cw, None, None, cw2, out_f16=True))
res["netvlad_k64_b256_f16"] = {"ms": ms}

u8 = torch.randint(0, 256, (B, T, D), dtype=torch.uint8, device=dev)
ms = timeit(lambda: nat.l2norm_rows(u8, num_frames=nf))
res["dequant_l2norm_b256"] = {"ms": ms, "GBps": B * T * D * 3 / ms / 1e6}

vl = torch.randn(B, K * D, device=dev).to(torch.bfloat16)
wfc = (torch.randn(1024, K * D, device=dev) * 0.01).to(torch.bfloat16)
ms = timeit(lambda: nat.linear(vl, wfc, act="relu6", out_bf16=True))
res["fc_73728x1024_b256"] = {"ms": ms, "TFLOPs": 2 * B * 1024 * K * D / ms / 1e9, "GBps_w": 1024 * K * D * 2 / ms / 1e6}
ms = timeit(lambda: nat.linear(vl, wfc, a_lo=vl, act="relu6", out_bf16=True, out_lo=True))
res["fc_73728x1024_b256_hilo"] = {"ms": ms}
vl16 = vl.to(torch.float16)
wfc16 = wfc.to(torch.float16)
ms = timeit(lambda: nat.linear(vl16, wfc16, act="relu6", out_f32=False, out_f16=True))
res["fc_73728x1024_b256_f16"] = {"ms": ms}
nat.debug_set_flags(16384)          # experiment: 128-wide tiles (32 KB stages, 6-deep ring) instead of 256-wide
ms = timeit(lambda: nat.linear(vl16, wfc16, act="relu6", out_f32=False, out_f16=True))
res["fc_73728x1024_b256_f16_bn128"] = {"ms": ms}
nat.debug_set_flags(16384 | 256)
ms = timeit(lambda: nat.linear(vl16, wfc16, act="relu6", out_f32=False, out_f16=True))
res["fc_73728x1024_b256_f16_bn128_mt2"] = {"ms": ms}
nat.debug_set_flags(256)            # experiment: pair the two M tiles in one CTA (W read once per CTA, twice the splits)
ms = timeit(lambda: nat.linear(vl16, wfc16, act="relu6", out_f32=False, out_f16=True))
res["fc_73728x1024_b256_f16_mt2"] = {"ms": ms}
nat.debug_set_flags(0)

for (d_in, m) in [(1024, 2), (4096, 4)]:
  rows = nat.moe_packed_rows(V, m)
  wp = (torch.randn(rows, d_in, device=dev) * 0.02).to(torch.bfloat16)
  bp = torch.zeros(rows, device=dev)
  for bb in (256, 2048):
    h = torch.randn(bb, d_in, device=dev).to(torch.bfloat16)
    ms = timeit(lambda: nat.moe_fwd(h, wp, bp, V, m))
    res["moe_d%d_m%d_b%d" % (d_in, m, bb)] = {"ms": ms, "TFLOPs": 2 * bb * rows * d_in / ms / 1e9}
    if bb == 256:
      ms = timeit(lambda: nat.moe_fwd(h, wp, bp, V, m, x_lo=h))
      res["moe_d%d_m%d_b%d_hilo" % (d_in, m, bb)] = {"ms": ms}
      nat.debug_set_flags(32768)      # A/B: both M tiles in one CTA (W tile read once), one CTA per SM
      ms = timeit(lambda: nat.moe_fwd(h, wp, bp, V, m))
      res["moe_d%d_m%d_b%d_mt2" % (d_in, m, bb)] = {"ms": ms}
      nat.debug_set_flags(1024)       # A/B: one CTA per SM
      ms = timeit(lambda: nat.moe_fwd(h, wp, bp, V, m))
      res["moe_d%d_m%d_b%d_1cta" % (d_in, m, bb)] = {"ms": ms}
      nat.debug_set_flags(0)

# big square GEMM through the same main loop (tensor-pipe ceiling of this kernel)
a = torch.randn(8192, 8192, device=dev).to(torch.bfloat16)
w = torch.randn(8192, 8192, device=dev).to(torch.bfloat16)
ms = timeit(lambda: nat.linear(a, w, out_f32=False, out_bf16=True), iters=5)
res["gemm_8192"] = {"ms": ms, "TFLOPs": 2 * 8192 ** 3 / ms / 1e9}

# LSTM config 3 at B=64
Bl, H = 64, 1024
xl = (torch.randn(Bl, T, D, device=dev) * 0.03).to(torch.bfloat16)
nfl = torch.full((Bl,), T, dtype=torch.int32, device=dev)
w0 = (torch.randn(4 * H, D + H, device=dev) * 0.02).to(torch.bfloat16)
w1 = (torch.randn(4 * H, 2 * H, device=dev) * 0.02).to(torch.bfloat16)
b0 = torch.zeros(4 * H, device=dev)
ms = timeit(lambda: nat.lstm_fwd(xl, nfl, [w0, w1], [b0, b0], H), iters=3, warm=1)
res["lstm_l2_h1024_b64"] = {"ms": ms, "videos_per_s": Bl / ms * 1e3}

# LSTM config 3 training step at B=64 and B=128: persistent forward + BPTT + MoE-4 head on the 4096-d state + Adam
import yt8m_trainer  # noqa: E402
for Bt in ((64, 128) if "lstm" in sys.argv else ()):
  trn = yt8m_trainer.LstmTrainer(D, hidden=H, layers=2, vocab=V, mixtures=4)
  trn.param.normal_(0.0, 0.02)
  for l in range(2):
    trn.w_bf16[l].copy_(trn.p["w%d" % l])
  trn.head.w_bf16.copy_(trn.head.w)
  xt = (torch.randn(Bt, T, D, device=dev) * 0.03).to(torch.bfloat16)
  nft = torch.randint(30, T + 1, (Bt,), dtype=torch.int32, device=dev)
  yt = (torch.rand(Bt, V, device=dev) < 3.4 / V).float()
  ms = timeit(lambda: trn.step(xt, nft, yt), iters=3, warm=1)
  res["lstm_train_step_b%d" % Bt] = {"ms": ms, "videos_per_s": Bt / ms * 1e3}
  ms = timeit(lambda: trn.forward(xt, nft), iters=3, warm=1)
  res["lstm_train_fwd_b%d" % Bt] = {"ms": ms}
  del trn

for k, v in res.items():
  print(k, json.dumps(v))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "quick_bench.json"), "w"), indent=1)
