"""GPU parity tests: every CUDA entry point of libyt8m_b200.so (called through the C ABI with raw device
pointers) against the CPU fp32 oracle on identical, bf16-representable seeded inputs.

Tolerance (north_star): 1e-3 relative on outputs.  Each test states how "relative" is measured.
"""
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


def rel_err(got, want):
  """max |got - want| / max |want|  (relative to the output scale)."""
  got = got.detach().float().cpu()
  want = want.detach().float().cpu()
  return float((got - want).abs().max() / want.abs().max().clamp_min(1e-30))


def bf(x):
  return x.to(torch.bfloat16).to(DEV)


def gen(seed):
  return torch.Generator().manual_seed(seed)


# ------------------------------------------------------------------------------------------------
# row transforms
# ------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("src", ["f32", "bf16", "u8"])
def test_l2norm_rows(nat, src):
  g = gen(1)
  b, t, d = 3, 40, 1152
  u8 = torch.randint(0, 256, (b, t, d), generator=g, dtype=torch.uint8)
  nf = torch.tensor([40, 17, 1], dtype=torch.int32)
  raw = O.dequantize(u8)
  mask = O.sequence_mask(nf, t).unsqueeze(2)
  if src == "u8":
    want = O.l2_normalize(raw * mask)
    got, got32 = nat.l2norm_rows(u8.to(DEV), num_frames=nf.to(DEV), want_f32=True)
  else:
    xin = synth.bf16r(raw * mask)
    want = O.l2_normalize(xin)
    xdev = xin.to(DEV) if src == "f32" else bf(xin)
    got, got32 = nat.l2norm_rows(xdev, want_f32=True)
  assert rel_err(got32, want) < 1e-5
  assert rel_err(got, want) < 2 ** -8          # bf16 rounding of the output only
  # padding rows stay exactly zero (l2_normalize of a zero row is 0)
  assert float(got32[1, 17:].abs().max()) == 0.0


def test_l2norm_zero_rows_and_no_normalize(nat):
  x = torch.zeros(4, 64)
  x[1] = 3.0
  got, got32 = nat.l2norm_rows(x.to(DEV), want_f32=True)
  assert float(got32[0].abs().max()) == 0.0
  assert abs(float(got32[1, 0]) - 3.0 / math.sqrt(64 * 9.0)) < 1e-6
  plain, plain32 = nat.l2norm_rows(x.to(DEV), normalize=False, want_f32=True)
  assert torch.equal(plain32.cpu(), x)


# ------------------------------------------------------------------------------------------------
# dense layer (tcgen05 GEMM + linear epilogue)
# ------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("m,n,k", [(128, 128, 64), (256, 1024, 1152), (37, 200, 136), (300, 24, 2304), (129, 4716, 1152),
                                   (256, 1024, 8192)])
@pytest.mark.parametrize("act", [None, "relu6", "sigmoid"])
def test_linear(nat, m, n, k, act):
  g = gen(m * 7 + n)
  a = synth.bf16r(torch.randn(m, k, generator=g))
  w = synth.xavier((k, n), g, gain=2.0)                  # TF layout [in, out]
  bias = torch.randn(n, generator=g) * 0.1
  scale = 1.0 + 0.1 * torch.randn(n, generator=g)
  fn = {None: None, "relu6": O.relu6, "sigmoid": torch.sigmoid}[act]
  want = O.fully_connected(a, w * 1.0, None) * scale + bias
  want = fn(want) if fn else want
  wp = nat.pack_transpose(w.to(DEV))
  assert torch.equal(wp[:, :k].float().cpu(), w.t().contiguous())       # packing is exact (bf16-representable)
  a_dev = torch.zeros(m, nat.pad8(k), dtype=torch.bfloat16, device=DEV)
  a_dev[:, :k] = bf(a)
  res = nat.linear(a_dev, wp, n=n, k=k, scale=scale.to(DEV), shift=bias.to(DEV), act=act, out_f32=True, out_bf16=True,
                   out_lo=True)
  assert rel_err(res["f32"], want) < 2e-5
  assert rel_err(res["hi"].float() + res["lo"].float(), want) < 3e-5   # hi+lo carries ~16 bits
  assert rel_err(res["hi"], want) < 2 ** -7


def test_linear_split_a(nat):
  """hi/lo split activation: fp32-level accuracy for a non-bf16-representable A."""
  g = gen(5)
  m, n, k = 200, 256, 1024
  a = torch.randn(m, k, generator=g)                     # NOT bf16 representable
  w = synth.xavier((k, n), g)
  want = a @ w
  wp = nat.pack_transpose(w.to(DEV))
  hi, lo = nat.split_bf16(a.to(DEV))
  got_split = nat.linear(hi, wp, a_lo=lo, n=n, k=k)["f32"]
  got_hi = nat.linear(hi, wp, n=n, k=k)["f32"]
  assert rel_err(got_split, want) < 5e-5
  assert rel_err(got_hi, want) > rel_err(got_split, want)  # the lo term matters


# ------------------------------------------------------------------------------------------------
# MoE head
# ------------------------------------------------------------------------------------------------

def _moe_weights(d, v, m, g, gain=1.0):
  gate_w = synth.xavier((d, v * (m + 1)), g, gain)
  expert_w = synth.xavier((d, v * m), g, gain)
  expert_b = 0.1 * torch.randn(v * m, generator=g)
  return gate_w, expert_w, expert_b


@pytest.mark.parametrize("b,d,v,m", [(128, 1024, 4716, 2), (200, 1152, 4716, 4), (5, 64, 30, 1), (64, 2048, 1000, 8),
                                     (33, 1352, 4716, 3)])
def test_moe(nat, b, d, v, m):
  g = gen(b + d + m)
  # gain chosen so the logits are O(1): exercises the softmax / sigmoid, not just their linear regime
  gain = math.sqrt(d) / 4
  gate_w, expert_w, expert_b = _moe_weights(d, v, m, g, gain)
  x = synth.bf16r(O.l2_normalize(torch.randn(b, d, generator=g)))
  want = O.moe_model(x, gate_w, expert_w, expert_b, v, m)
  wp, bp = nat.moe_pack(gate_w.to(DEV), expert_w.to(DEV), expert_b.to(DEV), v, m)
  xd = torch.zeros(b, nat.pad8(d), dtype=torch.bfloat16, device=DEV)
  xd[:, :d] = bf(x)
  got = nat.moe_fwd(xd, wp, bp, v, m, d=d)
  # probabilities: |dp| <= 1e-3 * p elementwise would need |dlogit| <= 1e-3; we check both views
  assert rel_err(got, want) < 1e-4
  gotc, wantc = got.cpu(), want
  assert float(((gotc - wantc).abs() / wantc.clamp_min(1e-6)).max()) < 1e-3


def test_moe_zero_weights_kat(nat):
  """Closed form: all-zero weights => p = M/(M+1) * 0.5."""
  b, d, v, m = 16, 64, 100, 2
  z = lambda *s: torch.zeros(*s, device=DEV)
  wp, bp = nat.moe_pack(z(d, v * (m + 1)), z(d, v * m), z(v * m), v, m)
  x = torch.randn(b, d, device=DEV).to(torch.bfloat16)
  got = nat.moe_fwd(x, wp, bp, v, m)
  assert torch.allclose(got.cpu(), torch.full((b, v), m / (m + 1) * 0.5), atol=1e-6)


def test_group_max(nat):
  x = torch.randn(6 * 8, 333)
  got = nat.group_max_rows(x.to(DEV), 8)
  assert torch.equal(got.cpu(), x.reshape(6, 8, 333).max(dim=1).values)


# ------------------------------------------------------------------------------------------------
# LSTM
# ------------------------------------------------------------------------------------------------

def _lstm_weights(in_dim, h, layers, g, gain=1.0):
  out = []
  for l in range(layers):
    i = in_dim if l == 0 else h
    out.append((synth.xavier((i + h, 4 * h), g, gain), 0.1 * torch.randn(4 * h, generator=g)))
  return out


@pytest.mark.parametrize("b,t,d,h,layers", [(4, 12, 64, 32, 1), (9, 20, 128, 64, 2), (130, 7, 1152, 128, 2),
                                                 # H % 256 == 0: the persistent recurrence (yt8m_lstm_rec.cu), 64 videos per launch
                                                 (3, 5, 64, 256, 1), (70, 9, 128, 256, 2), (64, 12, 1152, 1024, 2)])
def test_lstm(nat, b, t, d, h, layers):
  g = gen(b * 3 + t)
  x = synth.bf16r(torch.randn(b, t, d, generator=g) * 0.5)
  nf = torch.randint(1, t + 1, (b,), generator=g, dtype=torch.int32)
  nf[0] = t
  ws = _lstm_weights(d, h, layers, g, gain=2.0)
  outs, states = O.dynamic_rnn_lstm(x, nf, ws)
  want_state = O.lstm_model_state(states)
  packed = [nat.lstm_pack(w.to(DEV), bb.to(DEV), d if l == 0 else h, h) for l, (w, bb) in enumerate(ws)]
  state, seq, seq_bf = nat.lstm_fwd(bf(x), nf.to(DEV), [p[0] for p in packed], [p[1] for p in packed], h, want_seq=True,
                                    want_seq_bf16=True)
  assert rel_err(state, want_state) < 1e-4
  assert rel_err(seq, outs) < 1e-4
  assert rel_err(seq_bf, outs) < 2 ** -7
  # dynamic_rnn semantics: outputs past num_frames are exactly zero
  for bi in range(b):
    assert float(seq[bi, int(nf[bi]):].abs().max() if int(nf[bi]) < t else 0.0) == 0.0


def test_lstm_zero_weights_kat(nat):
  """Closed form: zero weights and bias => c_t = 0, h_t = 0."""
  b, t, d, h = 3, 5, 64, 32
  wp, bp = nat.lstm_pack(torch.zeros(d + h, 4 * h, device=DEV), torch.zeros(4 * h, device=DEV), d, h)
  x = torch.randn(b, t, d, device=DEV).to(torch.bfloat16)
  nf = torch.full((b,), t, dtype=torch.int32, device=DEV)
  state, _, _ = nat.lstm_fwd(x, nf, [wp], [bp], h)
  assert float(state.abs().max()) == 0.0


# ------------------------------------------------------------------------------------------------
# attention pooling
# ------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("mode", [0, 1])
def test_attn_pool_seqmask(nat, mode):
  g = gen(11 + mode)
  b, t, a, f = 5, 300, 8, 1152
  x, nf, _ = synth.model_input(b, t, f, seed=3)
  logits = torch.randn(b, t, a, generator=g) * 2
  mask = O.sequence_mask(nf, t).unsqueeze(2)
  if mode == 0:
    w = torch.softmax(logits, dim=1) * mask
    w = w / w.sum(dim=1, keepdim=True)
  else:
    w = torch.sigmoid(logits) * mask
    w = w / (w.sum(dim=1, keepdim=True) + 1e-8)
  want = torch.einsum("bta,btf->baf", w, x)
  got, hi, lo = nat.attn_pool(logits.to(DEV), bf(x), nf.to(DEV), a, mode)
  assert rel_err(got, want) < 1e-5
  assert rel_err(hi.float() + lo.float(), want) < 3e-5


def test_attn_pool_nonzero_mask_matches_attention_model(nat):
  """zt AttentionModel semantics (mask = frame has a non-zero entry) via the oracle's full restatement."""
  g = gen(21)
  b, t, a, d = 4, 300, 8, 1152
  x, nf, _ = synth.model_input(b, t, d, seed=4)
  w = synth.bf16r(torch.fmod(torch.randn(2 * d, a, generator=g), 2.0) * 0.1 * 20)
  bias = torch.full((a,), 0.1)
  want = O.attention_model_pool(x, nf, w, bias)                       # [B*A, D]
  # the mean-pooled half of the logits and the bias are constant over t: they cancel in softmax over T
  wp = nat.pack_transpose(w[:d].to(DEV))
  logits = nat.linear(bf(x).reshape(b * t, d), wp, n=a, k=d)["f32"]
  logits = logits.as_strided((b, t, a), (t * logits.stride(0), logits.stride(0), 1))
  got, _, _ = nat.attn_pool(logits, bf(x), None, a, 0)
  assert rel_err(got.reshape(b * a, d), want) < 1e-4


def test_attn_equal_logits_is_mean_pool(nat):
  """Closed form: equal logits => mean over the valid frames."""
  b, t, a, f = 2, 50, 4, 64
  x, nf, _ = synth.model_input(b, t, f, seed=5, min_frames=10)
  got, _, _ = nat.attn_pool(torch.zeros(b, t, a, device=DEV), bf(x), nf.to(DEV), a, 0)
  for bi in range(b):
    want = x[bi, :int(nf[bi])].mean(dim=0)
    assert torch.allclose(got[bi, 0].cpu(), want, atol=1e-6)


# ------------------------------------------------------------------------------------------------
# NetVLAD (fused)
# ------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("b,t,d,k", [(3, 300, 1152, 64), (2, 300, 1152, 128), (4, 100, 256, 32), (5, 128, 128, 64), (2, 257, 384, 64)])
def test_netvlad(nat, b, t, d, k):
  g = gen(b + t + k)
  x, nf, _ = synth.model_input(b, t, d, seed=6, min_frames=min(30, t))
  nf[0] = t
  cw = synth.normal((d, k), g, 4.0)                      # logits O(1): peaky assignments
  scale = 1.0 + 0.1 * torch.randn(k, generator=g)
  shift = 0.1 * torch.randn(k, generator=g)
  cw2 = synth.normal((d, k), g, 1 / math.sqrt(d))
  want = O.netvlad_pool(x, nf, cw, scale, shift, cw2)
  cwp = nat.pack_transpose(cw.to(DEV))
  hi, lo, f32 = nat.netvlad_fwd(bf(x), nf.to(DEV), cwp, scale.to(DEV), shift.to(DEV), cw2.to(DEV), want_f32=True, want_lo=True)
  # relative to the descriptor's scale (entries ~ 1/sqrt(D*K)); the assignment is bf16-rounded for
  # the second GEMM, so this is looser than the GEMM tests
  assert rel_err(f32, want) < 4e-3
  err_l2 = float((f32.cpu() - want).norm() / want.norm())
  assert err_l2 < 1e-3
  assert rel_err(hi.float() + lo.float(), f32) < 1e-4
  # every cluster column has unit norm / sqrt(K):  ||v||_2 == 1
  assert abs(float(f32[0].norm()) - 1.0) < 1e-3


# ------------------------------------------------------------------------------------------------
# loss / top-k / gating
# ------------------------------------------------------------------------------------------------

def test_xent(nat):
  g = gen(31)
  b, v = 37, 4716
  p = torch.rand(b, v, generator=g)
  y = synth.labels(b, v)
  want = O.cross_entropy_loss(p, y)
  loss, dp = nat.xent(p.to(DEV), y.to(DEV), want_grad=True)
  assert abs(float(loss) - float(want)) / float(want) < 1e-5
  pr = p.clone().requires_grad_(True)
  O.cross_entropy_loss(pr, y).backward()
  assert rel_err(dp, pr.grad) < 1e-5


def test_topk(nat):
  g = gen(32)
  x = torch.rand(50, 4716, generator=g)
  idx, val = nat.topk_rows(x.to(DEV), 20)
  wv, wi = torch.topk(x, 20, dim=1)
  assert torch.equal(val.cpu(), wv)
  assert torch.equal(idx.cpu().long(), wi)


def test_context_gate(nat):
  g = gen(33)
  x = torch.randn(64, 1024, generator=g)
  gg = torch.randn(64, 1024, generator=g)
  sc = torch.rand(1024, generator=g) + 0.5
  sh = torch.randn(1024, generator=g)
  out, hi, lo = nat.context_gate(x.to(DEV), gg.to(DEV), sc.to(DEV), sh.to(DEV))
  want = x * torch.sigmoid(gg * sc + sh)
  assert rel_err(out, want) < 1e-5


# ------------------------------------------------------------------------------------------------
# fp16 activation operands (YT8M_FMT_F16): fp16 A x bf16 W on the tensor cores, one MMA per tile
# ------------------------------------------------------------------------------------------------

def test_linear_f16_operand_exact_and_rounded(nat):
  g = gen(41)
  m, n, k = 200, 384, 1024
  a = torch.randn(m, k, generator=g)
  a16 = a.to(torch.float16)                                  # what the kernel sees
  w = synth.xavier((k, n), g)
  wp = nat.pack_transpose(w.to(DEV)).to(torch.float16)       # bf16 -> fp16: exact above 2^-17, < 2^-24 absolute below
  assert float((wp[:, :k].float().cpu() - w.t()).abs().max()) <= 2.0 ** -25
  w = wp[:, :k].float().cpu().t().contiguous()                # the weights the kernel sees
  res = nat.linear(a16.to(DEV), wp, n=n, k=k, out_f32=True, out_f16=True)
  with pytest.raises(nat.Yt8mError):                         # one 16-bit format per MMA: fp16 x bf16 is refused
    nat.linear(a16.to(DEV), nat.pack_transpose(w.to(DEV)), n=n, k=k)
  # fp16 x fp16 products are exact in the fp32 accumulator: against the fp16-rounded input only summation order differs
  assert rel_err(res["f32"], a16.float() @ w) < 2e-5
  # against the unrounded activation: 11 significant bits per element, averaged over k
  assert rel_err(res["f32"], a @ w) < 5e-4
  assert res["hi"].dtype == torch.float16
  assert rel_err(res["hi"], res["f32"]) < 2 ** -10
  # split-K path (few tiles, long K) writes fp16 from the finalize kernel
  m2, n2, k2 = 64, 128, 8192
  a2 = torch.randn(m2, k2, generator=g).to(torch.float16)
  w2 = synth.xavier((k2, n2), g)
  r2 = nat.linear(a2.to(DEV), nat.pack_transpose(w2.to(DEV)).to(torch.float16), n=n2, k=k2, act="relu6", out_f32=True, out_f16=True)
  want2 = O.relu6(a2.float() @ w2)
  assert rel_err(r2["f32"], want2) < 2e-5
  assert rel_err(r2["hi"], want2) < 2 ** -10


def test_moe_f16_operand(nat):
  g = gen(42)
  b, d, v, m = 130, 1024, 4716, 2
  x = (torch.rand(b, d, generator=g) * 6.0).to(torch.float16)          # a ReLU6 activation
  gw, ew, eb = _moe_weights(d, v, m, g, gain=0.3)
  wp, bp = nat.moe_pack(gw.to(DEV), ew.to(DEV), eb.to(DEV), v, m)
  got = nat.moe_fwd(x.to(DEV), wp.to(torch.float16), bp, v, m)
  assert rel_err(got, O.moe_model(x.float(), gw, ew, eb, v, m)) < 2e-5


@pytest.mark.parametrize("k", [64, 128])
def test_netvlad_f16_output(nat, k):
  g = gen(43 + k)
  b, t, d = 7, 300, 1152
  x, nf, _ = synth.model_input(b, t, d, seed=16)
  cw, cw2 = synth.normal((d, k), g, 4.0), synth.normal((d, k), g, 1 / math.sqrt(d))
  want = O.netvlad_pool(x, nf, cw, torch.ones(k), torch.zeros(k), cw2)
  cwp = nat.pack_transpose(cw.to(DEV))
  h16, lo, f32 = nat.netvlad_fwd(bf(x), nf.to(DEV), cwp, None, None, cw2.to(DEV), want_f32=True, out_f16=True)
  assert h16.dtype == torch.float16 and lo is None
  assert float((f32.cpu() - want).norm() / want.norm()) < 1.5e-3        # stash is fp16: one more 2^-11 rounding than hi+lo
  assert rel_err(h16, f32) < 2 ** -10
  # the fp16 descriptor as FC operand: 1e-3 of the output scale (north_star tolerance) with margin
  w = synth.xavier((d * k, 256), g)
  got = nat.linear(h16, nat.pack_transpose(w.to(DEV)).to(torch.float16), n=256, k=d * k)["f32"]
  assert rel_err(got, want @ w) < 1e-3
