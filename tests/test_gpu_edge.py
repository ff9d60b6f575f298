"""GPU edge cases the reference's data can produce: single-video batches, one-frame / zero-frame videos, ragged
num_frames, tiny dimensions, all-padding rows -- CUDA vs oracle."""
import math

import pytest
import torch

import synth
from oracle import yt8m_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def nat():
  if not torch.cuda.is_available():
    pytest.skip("no CUDA device")
  import yt8m_native
  return yt8m_native


def bf(x):
  return x.to(torch.bfloat16).to(DEV)


def rel(got, want):
  return float((got.float().cpu() - want).abs().max() / want.abs().max().clamp_min(1e-30))


def test_netvlad_zero_one_and_full_frames(nat):
  g = torch.Generator().manual_seed(0)
  b, t, d, k = 5, 300, 1152, 64
  x, nf, _ = synth.model_input(b, t, d, seed=20)
  nf = torch.tensor([0, 1, 44, 129, 300], dtype=torch.int32)
  x = x * (torch.arange(t).unsqueeze(0) < nf.unsqueeze(1)).float().unsqueeze(2)     # reader pads with zeros
  cw, cw2 = synth.normal((d, k), g, 4.0), synth.normal((d, k), g, 1 / math.sqrt(d))
  want = O.netvlad_pool(x, nf, cw, torch.ones(k), torch.zeros(k), cw2)
  hi, lo, f32 = nat.netvlad_fwd(bf(x), nf.to(DEV), nat.pack_transpose(cw.to(DEV)), None, None, cw2.to(DEV), want_f32=True, want_lo=True)
  assert float(f32[0].abs().max()) == 0.0 and float(want[0].abs().max()) == 0.0           # no frames -> zero descriptor
  assert bool(torch.isfinite(f32).all())
  assert float((f32.cpu() - want).norm() / want.norm()) < 1e-3
  for i in range(1, b):
    assert abs(float(f32[i].norm()) - 1.0) < 1e-3


def test_single_video_and_single_row(nat):
  g = torch.Generator().manual_seed(1)
  x, nf, _ = synth.model_input(1, 300, 1152, seed=21)
  cw, cw2 = synth.normal((1152, 64), g, 4.0), synth.normal((1152, 64), g, 0.03)
  want = O.netvlad_pool(x, nf, cw, torch.ones(64), torch.zeros(64), cw2)
  cwp = nat.pack_transpose(cw.to(DEV))
  _, _, f32 = nat.netvlad_fwd(bf(x), nf.to(DEV), cwp, None, None, cw2.to(DEV), want_f32=True)
  assert float((f32.cpu() - want).norm() / want.norm()) < 1e-3
  # ... and a video's descriptor does not depend on what else is in the batch.  Not bit-exact: the cluster sums
  # are accumulated with shared-memory atomics in no fixed order, and a 1e-7 wobble can flip the rounding of the
  # lo half of the stash (2^-16 relative)
  x3, nf3, _ = synth.model_input(3, 300, 1152, seed=22)
  x3[1], nf3[1] = x[0], nf[0]
  _, _, f32b = nat.netvlad_fwd(bf(x3), nf3.to(DEV), cwp, None, None, cw2.to(DEV), want_f32=True)
  assert float((f32b[1] - f32[0]).abs().max()) <= 1e-4 * float(f32[0].abs().max())
  # MoE / linear with one row and tiny shapes
  d, v, m = 8, 3, 2
  gw, ew, eb = synth.xavier((d, v * 3), g, 2.0), synth.xavier((d, v * 2), g, 2.0), 0.1 * torch.randn(v * 2, generator=g)
  xr = synth.bf16r(torch.randn(1, d, generator=g))
  wp, bp = nat.moe_pack(gw.to(DEV), ew.to(DEV), eb.to(DEV), v, m)
  assert rel(nat.moe_fwd(bf(xr), wp, bp, v, m), O.moe_model(xr, gw, ew, eb, v, m)) < 1e-5
  w = synth.xavier((d, 1), g)
  got = nat.linear(bf(xr), nat.pack_transpose(w.to(DEV)), n=1, k=d)["f32"]
  assert rel(got, xr @ w) < 1e-5


def test_lstm_ragged_including_zero_length(nat):
  g = torch.Generator().manual_seed(2)
  b, t, d, h = 6, 9, 64, 32
  x = synth.bf16r(torch.randn(b, t, d, generator=g) * 0.5)
  nf = torch.tensor([0, 1, 9, 5, 2, 9], dtype=torch.int32)
  ws = [(synth.xavier((d + h, 4 * h), g, 2.0), 0.1 * torch.randn(4 * h, generator=g)), (synth.xavier((2 * h, 4 * h), g, 2.0), torch.zeros(4 * h))]
  outs, states = O.dynamic_rnn_lstm(x, nf, ws)
  packed = [nat.lstm_pack(w.to(DEV), bb.to(DEV), d if l == 0 else h, h) for l, (w, bb) in enumerate(ws)]
  state, seq, _ = nat.lstm_fwd(bf(x), nf.to(DEV), [p[0] for p in packed], [p[1] for p in packed], h, want_seq=True)
  assert rel(state, O.lstm_model_state(states)) < 1e-4
  assert float(state[0].abs().max()) == 0.0                      # zero-length video keeps the zero state
  assert rel(seq, outs) < 1e-4 and float(seq[0].abs().max()) == 0.0


def test_attention_ragged_and_single_frame(nat):
  g = torch.Generator().manual_seed(3)
  b, t, a, f = 4, 300, 8, 1152
  x, _, _ = synth.model_input(b, t, f, seed=22, min_frames=t)     # every frame present, then cut to nf below
  nf = torch.tensor([1, 2, 299, 300], dtype=torch.int32)
  x = x * (torch.arange(t).unsqueeze(0) < nf.unsqueeze(1)).float().unsqueeze(2)
  logits = torch.randn(b, t, a, generator=g)
  mask = O.sequence_mask(nf, t).unsqueeze(2)
  w = torch.softmax(logits, dim=1) * mask
  w = w / w.sum(dim=1, keepdim=True)
  want = torch.einsum("bta,btf->baf", w, x)
  got, _, _ = nat.attn_pool(logits.to(DEV), bf(x), nf.to(DEV), a, 0)
  assert rel(got, want) < 1e-5
  assert torch.allclose(got[0, 3].cpu(), x[0, 0], atol=1e-6)      # one valid frame: the pool IS that frame
  # non-zero-frame mask variant gives the same answer when padding rows are exactly zero
  got2, _, _ = nat.attn_pool(logits.to(DEV), bf(x), None, a, 0)
  assert rel(got2, want) < 1e-5


def test_l2norm_all_padding_and_uint8_extremes(nat):
  u8 = torch.zeros((2, 4, 64), dtype=torch.uint8)
  u8[1] = 255
  nf = torch.tensor([0, 4], dtype=torch.int32)
  out, f32 = nat.l2norm_rows(u8.to(DEV), num_frames=nf.to(DEV), want_f32=True)
  assert float(f32[0].abs().max()) == 0.0                        # padding rows stay zero, not Dequantize(0)
  assert torch.allclose(f32[1].cpu(), torch.full((4, 64), 1 / 8.0), atol=1e-6)


def test_topk_with_ties_and_k1(nat):
  x = torch.zeros(3, 40)
  x[0, 7] = x[0, 3] = 0.5
  x[1, 39] = 1.0
  idx, val = nat.topk_rows(x.to(DEV), 3)
  assert idx[0].tolist()[:2] == [3, 7] and val[0].tolist()[:2] == [0.5, 0.5]      # ties: lower class index first
  assert idx[1, 0].item() == 39 and idx[2].tolist() == [0, 1, 2]
  idx1, _ = nat.topk_rows(x.to(DEV), 1)
  assert idx1.flatten().tolist() == [3, 39, 0]


def test_ragged_ingest_equals_padded_ingest(nat):
  """yt8m_frames_unpack_u8 on a readers.PackedFrames == yt8m_l2norm_rows_fwd on the reader's padded uint8 batch, bit for bit
  (de-quantise + L2-normalise + zero padding); includes a zero-frame and a full-length video."""
  import readers
  import feature_transform
  g = torch.Generator().manual_seed(12)
  b, t, d = 7, 40, 1152
  u8 = torch.randint(0, 256, (b, t, d), generator=g, dtype=torch.uint8)
  nf = torch.tensor([40, 0, 1, 17, 39, 8, 40], dtype=torch.int32)
  u8 = u8 * (torch.arange(t).unsqueeze(0) < nf.unsqueeze(1)).unsqueeze(2).to(torch.uint8)      # the reader pads with zeros
  want = nat.l2norm_rows(u8.to(DEV), num_frames=nf.to(DEV))
  packed = readers.PackedFrames.from_padded(u8, nf)
  assert packed.data.shape[0] == int(nf.sum())
  got, _ = feature_transform.DefaultTransformer().transform(packed.pin_memory(), nf)
  assert got.shape == want.shape and torch.equal(got.view(torch.int16), want.view(torch.int16))
  assert float(got[1].float().abs().max()) == 0.0
  # oracle: Dequantize + l2_normalize of the real frames
  x = O.l2_normalize(O.dequantize(u8.float()))
  x = x * (torch.arange(t).unsqueeze(0) < nf.unsqueeze(1)).unsqueeze(2)
  assert float((got.float().cpu() - x).abs().max()) < 2 ** -8


@pytest.mark.parametrize("b,t,d", [(200, 96, 128), (1100, 96, 128), (90, 300, 256)])
def test_netvlad_one_pass_kernel_short_and_empty_videos(nat, b, t, d):
  """The one-pass cluster kernel (yt8m_netvlad_v4.cu: fp16 output, K = 64) streams only ceil(num_frames / 32) tiles per
  video and hands videos to the clusters longest first (B <= 1024) or round-robin (B > 1024).  Many one- and two-tile
  videos, empty videos and full-length ones in one batch, several videos per cluster: every descriptor vs the oracle."""
  g = torch.Generator().manual_seed(b)
  k = 64
  x = synth.bf16r(torch.randn(b, t, d, generator=g))
  x = x * torch.rsqrt((x * x).sum(dim=2, keepdim=True))
  nf = torch.randint(0, t + 1, (b,), generator=g, dtype=torch.int32)
  nf[:6] = torch.tensor([0, 1, t, 32, 33, 0], dtype=torch.int32)
  x = synth.bf16r(x * (torch.arange(t).unsqueeze(0) < nf.unsqueeze(1)).float().unsqueeze(2))
  cw, cw2 = synth.normal((d, k), g, 4.0), synth.normal((d, k), g, 1 / math.sqrt(d))
  want = O.netvlad_pool(x, nf, cw, torch.ones(k), torch.zeros(k), cw2)
  out = nat.netvlad_fwd(bf(x), nf.to(DEV), nat.pack_transpose(cw.to(DEV)), None, None, cw2.to(DEV), out_f16=True)[0]
  got = out.view(torch.float16).float().cpu() if out.dtype != torch.float16 else out.float().cpu()
  assert bool(torch.isfinite(got).all())
  assert float(got[0].abs().max()) == 0.0 and float(got[5].abs().max()) == 0.0          # no frames -> zero descriptor
  err = (got - want).norm(dim=1) / want.norm(dim=1).clamp_min(1e-6)
  assert float(err.max()) < 4e-3, (int(err.argmax()), float(err.max()), int(nf[int(err.argmax())]))


@pytest.mark.parametrize("two_kernels", [False, True])
@pytest.mark.parametrize("b,t,d,fmt,k", [(200, 96, 256, "f16", 64), (1100, 70, 256, "f16", 64), (90, 300, 1152, "f16", 64),
                                         (41, 300, 1024, "bf16", 64), (3, 64, 320, "f16", 64), (150, 129, 1280, "f16", 64),
                                         (60, 300, 1152, "f16", 128), (310, 96, 256, "bf16", 128), (5, 65, 1280, "f16", 128)])
def test_netvlad_tiled_kernel_short_and_empty_videos(nat, b, t, d, fmt, k, two_kernels):
  """yt8m_netvlad_fwd_tiled (yt8m_netvlad_v5.cu: a cluster of FOUR CTAs per video, 64-frame tiles, blocked descriptor / cw2
  layouts): only ceil(num_frames / 64) tiles per video are streamed, videos are scheduled longest first by a counting sort
  (several videos per cluster, more clusters than videos, uneven feature splits 5,5,4,4 / 4,4,4,4 / 2,1,1,1 / 5,5,5,5).  Empty,
  one-frame, tile-boundary and full-length videos in one batch: every descriptor, un-tiled, against the oracle; the saved
  statistics against their definitions.  K = 128 (BASELINE.json configs[3], the gated model): two X slots, cw2 from global memory,
  the accumulator corrected in TMEM between the passes.  two_kernels: the assignment + aggregation pair (yt8m_netvlad_v6.cu; D <= 1152, else the
  entry point falls back to the one-pass kernel)."""
  g = torch.Generator().manual_seed(b + d)
  if two_kernels and k != 64:
    pytest.skip("the two-kernel pair is built for K = 64")
  assert nat.netvlad_tiled_supported(t, d, k)
  x = synth.bf16r(torch.randn(b, t, d, generator=g))
  x = x * torch.rsqrt((x * x).sum(dim=2, keepdim=True))
  nf = torch.randint(0, t + 1, (b,), generator=g, dtype=torch.int32)
  nf[:3] = torch.tensor([0, 1, t], dtype=torch.int32)
  if b > 8:
    nf[3:8] = torch.tensor([min(64, t), min(65, t), 0, min(63, t), min(128, t)], dtype=torch.int32)
  x = synth.bf16r(x * (torch.arange(t).unsqueeze(0) < nf.unsqueeze(1)).float().unsqueeze(2))
  cw, cw2 = synth.normal((d, k), g, 4.0), torch.randn(d, k, generator=g) / math.sqrt(d)          # cw2 is NOT bf16-representable
  scale, shift = 1.0 + 0.1 * torch.randn(k, generator=g), 0.1 * torch.randn(k, generator=g)
  want = O.netvlad_pool(x, nf, cw, scale, shift, cw2)
  idx8, idx4 = nat.netvlad_tiled_index(d, k, 8, DEV), nat.netvlad_tiled_index(d, k, 4, DEV)
  c2t = cw2.to(DEV).reshape(-1)[idx4].contiguous()
  out, stats = nat.netvlad_fwd_tiled(bf(x), nf.to(DEV), nat.pack_transpose(cw.to(DEV)), scale.to(DEV), shift.to(DEV), c2t,
                                     out_f16=(fmt == "f16"), want_stats=True, two_kernels=two_kernels)
  got = torch.empty_like(out)
  got[:, idx8] = out                                      # tiled position p holds row-major element idx8[p]
  got = got.float().cpu()
  assert bool(torch.isfinite(got).all())
  assert float(got[0].abs().max()) == 0.0                 # no frames -> zero descriptor
  err = (got - want).norm(dim=1) / want.norm(dim=1).clamp_min(1e-6)
  tol = 4e-3 if fmt == "f16" else 8e-3                    # one fp16 / bf16 rounding of the output on top of the bf16 assignment
  assert float(err.max()) < tol, (int(err.argmax()), float(err.max()), int(nf[int(err.argmax())]))
  # stats = {a_sum[K], ||V[:, k]||^2 [K], sum_k ||V_k||^2 / max(||V_k||^2, eps)}: a_sum adds up to the number of frames
  st = stats.cpu()
  assert float((st[:, :k].sum(dim=1) - nf.clamp(0, t).float()).abs().max()) < 0.05 * max(t, 1) ** 0.5 + 0.6
  assert bool((st[:, k:2 * k] >= 0).all())
