"""GPU parity of the training step (wh/train.py:440-466 semantics) against torch-CPU autograd over the oracle
forward + the oracle's clip_by_norm / TF-Adam: gradients, loss, and the weights after several steps."""
import math

import pytest
import torch

import synth
from oracle import yt8m_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def tr():
  if not torch.cuda.is_available():
    pytest.skip("no CUDA device")
  import yt8m_trainer
  return yt8m_trainer


def _data(b, d, v, seed):
  g = torch.Generator().manual_seed(seed)
  x = torch.randn(b, d, generator=g)
  x = synth.bf16r(x * torch.rsqrt((x * x).sum(dim=1, keepdim=True)))
  return x, synth.labels(b, v, seed=seed, per_video=min(3.4, v / 4)), g


def _oracle_steps(kind, sd, x, y, vocab, mixtures, steps, l2=1e-8, reg_penalty=1.0, clip=1.0, base_lr=0.01):
  params = {k: t.clone().requires_grad_(True) for k, t in sd.items()}
  m = {k: torch.zeros_like(t) for k, t in sd.items()}
  v = {k: torch.zeros_like(t) for k, t in sd.items()}
  first = None
  for step in range(steps):
    if kind == "logistic":
      p = O.logistic_model(x, params["fully_connected/weights"], params["fully_connected/biases"])
    else:
      p = O.moe_model(x, params["gates/weights"], params["experts/weights"], params["experts/biases"], vocab, mixtures)
    label_loss = O.cross_entropy_loss(p, y)
    reg = sum(O.l2_regularizer(t, l2) for k, t in params.items() if k.endswith("weights"))
    grads_label = torch.autograd.grad(label_loss, list(params.values()), retain_graph=True)
    grads = torch.autograd.grad(label_loss + reg_penalty * reg, list(params.values()))
    if first is None:
      first = ({k: g.clone() for k, g in zip(params, grads_label)}, float(label_loss), float(reg))
    lr = O.exponential_decay(base_lr, step, x.shape[0], 4000000, 0.95)
    with torch.no_grad():
      for (k, t), g in zip(params.items(), grads):
        g = O.clip_by_norm(g, clip)
        new, m[k], v[k] = O.adam_step(t, g, m[k], v[k], step + 1, lr)
        t.copy_(new)
  return {k: t.detach() for k, t in params.items()}, first


@pytest.mark.parametrize("kind,b,d,v,mix", [("logistic", 128, 1152, 4716, 0), ("moe", 96, 256, 500, 2), ("moe", 64, 1024, 4716, 2),
                                            ("moe", 40, 136, 333, 4)])
def test_train_step_parity(tr, kind, b, d, v, mix):
  x, y, g = _data(b, d, v, seed=b + d)
  gain = math.sqrt(d) / 4
  if kind == "logistic":
    sd = {"fully_connected/weights": synth.xavier((d, v), g, gain), "fully_connected/biases": 0.1 * torch.randn(v, generator=g)}
  else:
    sd = {"gates/weights": synth.xavier((d, v * (mix + 1)), g, gain), "experts/weights": synth.xavier((d, v * mix), g, gain),
          "experts/biases": 0.1 * torch.randn(v * mix, generator=g)}
  t = tr.HeadTrainer(kind, d, v, mixtures=max(mix, 1))
  t.import_state({k: w.to(DEV) for k, w in sd.items()})
  t.keep_grads = True
  steps = 3
  for _ in range(steps):
    t.step(x.to(DEV).to(torch.bfloat16), y.to(DEV))
    if _ == 0:
      grad0 = t.grads_tf_layout(t.last_grad)
      loss0 = float(t.last["label_loss_local"])
      reg0 = t.reg_loss()
  torch.cuda.synchronize()
  want, (wgrads, wloss, wreg) = _oracle_steps(kind, sd, x, y, v, mix, steps)
  # loss and regulariser of the first step
  assert abs(loss0 - wloss) / wloss < 1e-4
  assert abs(reg0 - wreg) / max(wreg, 1e-30) < 1e-3
  # gradients of the label loss (before reg / clip): relative to each tensor's largest entry
  for k in wgrads:
    err = float((grad0[k] - wgrads[k]).abs().max() / wgrads[k].abs().max())
    assert err < 2e-3, (k, err)
  # weights after `steps` Adam steps: every step moves a weight by ~lr; compare the total displacement
  got = t.export_state()
  for k in want:
    dw_got, dw_want = got[k] - sd[k], want[k] - sd[k]
    bad = ((dw_got - dw_want).abs() > 0.05 * 0.01 * steps).float().mean()
    assert float(bad) < 2e-3, (k, float(bad))          # near-zero gradients make a few updates ill-conditioned
  # the bf16 operand copy follows the fp32 master
  assert torch.equal(t.w_bf16.float(), t.w.to(torch.bfloat16).float())


def test_moe_untouched_rows(tr):
  """Padding rows of the packed matrix and the (non-existent) gate biases never move."""
  b, d, v, mix = 16, 64, 30, 2
  x, y, g = _data(b, d, v, seed=3)
  sd = {"gates/weights": synth.xavier((d, v * 3), g, 4.0), "experts/weights": synth.xavier((d, v * 2), g, 4.0),
        "experts/biases": torch.zeros(v * 2)}
  t = tr.HeadTrainer("moe", d, v, mixtures=mix)
  t.import_state({k: w.to(DEV) for k, w in sd.items()})
  t.step(x.to(DEV).to(torch.bfloat16), y.to(DEV))
  gates, experts = tr._moe_row_index(v, mix)
  used = torch.zeros(t.rows, dtype=torch.bool)
  used[gates] = True
  used[experts] = True
  assert float(t.w.cpu()[~used].abs().max()) == 0.0
  assert float(t.b.cpu()[gates].abs().max()) == 0.0
  assert float(t.b.cpu()[experts].abs().max()) > 0.0


# ------------------------------------------------------------------------------------------------
# backward of the NetVLAD layer (kernel level): every gradient against torch autograd over the oracle forward
# ------------------------------------------------------------------------------------------------

def _rel_l2(got, want):
  got, want = got.detach().float().cpu(), want.detach().float().cpu()
  return float((got - want).norm() / want.norm().clamp_min(1e-30))


@pytest.mark.parametrize("k,with_scale", [(64, False), (64, True), (128, False)])
def test_netvlad_backward_kernels(k, with_scale):
  if not torch.cuda.is_available():
    pytest.skip("no CUDA device")
  import yt8m_native as nat
  g = torch.Generator().manual_seed(70 + k)
  b, t, d = 5, 300, 1152
  x, nf, _ = synth.model_input(b, t, d, seed=33)
  nf[0], nf[1] = t, 1
  cw = synth.normal((d, k), g, 4.0).requires_grad_(True)
  shift = (0.1 * torch.randn(k, generator=g)).requires_grad_(True)
  scale = (1.0 + 0.1 * torch.randn(k, generator=g)) if with_scale else torch.ones(k)
  cw2 = synth.normal((d, k), g, 1 / math.sqrt(d)).requires_grad_(True)
  dy = torch.randn(b, d * k, generator=g) * 1e-3
  y = O.netvlad_pool(x, nf, cw, scale, shift, cw2)
  (y * dy).sum().backward()

  xd = x.to(DEV).to(torch.bfloat16)
  cwp = nat.pack_transpose(cw.detach().to(DEV))
  sc = scale.to(DEV) if with_scale else None
  _, _, y32, stats = nat.netvlad_fwd(xd, nf.to(DEV), cwp, sc, shift.detach().to(DEV), cw2.detach().to(DEV), want_f32=True,
                                     want_lo=True, want_stats=True)
  assert _rel_l2(y32, y) < 1e-3
  dv, dasum, dcw2, dv_split = nat.netvlad_bwd_norm(dy.to(DEV), y32, stats, cw2.detach().to(DEV), want_split=True)
  assert float(((dv_split[0].float() + dv_split[1].float()) - dv).abs().max()) <= 2.0 ** -16 * float(dv.abs().max())
  z = nat.linear(xd.reshape(b * t, d), cwp, n=k, k=d, scale=sc, shift=shift.detach().to(DEV))["f32"]
  dz_hi, dz_lo, dshift = nat.netvlad_bwd_assign(xd, nf.to(DEV), z, dv, dasum, scale=sc)
  dcw_t = nat.wgrad(dz_hi, dz_lo, xd.reshape(b * t, d), k, d)                 # [K, D] = dL/dCw^T
  # the one-kernel tensor-core version (logits recomputed on chip, yt8m_netvlad_bwd_tc.cu): same gradients
  assert nat.netvlad_bwd_assign_fused_supported(t, d, k)
  fz_hi, fz_lo, fshift = nat.netvlad_bwd_assign_fused(xd, nf.to(DEV), cwp, sc, shift.detach().to(DEV), dv_split, dasum)
  fcw_t = nat.wgrad(fz_hi, fz_lo, xd.reshape(b * t, d), k, d)
  assert _rel_l2(fshift, shift.grad) < 1e-2
  assert _rel_l2(fcw_t.t(), cw.grad) < 1e-2
  assert _rel_l2(fz_hi.float() + fz_lo.float(), dz_hi.float() + dz_lo.float()) < 2e-3
  assert float((fz_hi.float().reshape(b, t, k))[1, 1:].abs().max()) == 0.0
  # gradients pass through bf16 assignments / bf16 hi-lo dz: 1e-2 of the gradient's norm
  assert _rel_l2(dcw2, cw2.grad) < 1e-2
  assert _rel_l2(dshift, shift.grad) < 1e-2
  assert _rel_l2(dcw_t.t(), cw.grad) < 1e-2
  # padded frames carry no gradient
  rows = dz_hi.float().reshape(b, t, k)
  assert float(rows[1, 1:].abs().max()) == 0.0


def test_act_bwd_and_wgrad_split():
  if not torch.cuda.is_available():
    pytest.skip("no CUDA device")
  import yt8m_native as nat
  g = torch.Generator().manual_seed(5)
  rows, cols = 37, 200
  pre = torch.randn(rows, cols, generator=g) * 4
  dy = torch.randn(rows, cols, generator=g)
  sc = 1.0 + 0.1 * torch.randn(cols, generator=g)
  for act, fn, dfn in (("relu6", O.relu6, lambda yy: ((yy > 0) & (yy < 6)).float()), ("sigmoid", torch.sigmoid, lambda yy: yy * (1 - yy)),
                       (None, lambda v: v, lambda yy: torch.ones_like(yy))):
    yy = fn(pre)
    hi, lo = nat.act_bwd(dy.to(DEV), yy.to(DEV), act=act, col_scale=sc.to(DEV))
    want = dy * dfn(yy) * sc
    assert float(((hi.float() + lo.float()).cpu() - want).abs().max()) < 1e-4 * float(want.abs().max())
  # wgrad with a long contraction and few output tiles takes the split path (fp32 atomics): same numbers
  kb, m, n = 4096, 64, 256
  a = synth.bf16r(torch.randn(kb, m, generator=g))
  bm = synth.bf16r(torch.randn(kb, n, generator=g))
  got = nat.wgrad(a.to(DEV).to(torch.bfloat16), None, bm.to(DEV).to(torch.bfloat16), m, n)
  assert float((got.cpu() - a.t() @ bm).abs().max()) < 1e-3 * float((a.t() @ bm).abs().max())


def test_dgrad_through_forward_weights_and_sliced_colsum():
  """nat.dgrad (dx = dz . W with W in the FORWARD's [out, in] bf16 layout: the MN-major GEMM on dz^T, no transposed weight
  copy) against fp32 matmul, for a batch that is not a multiple of 8 and for the split-K shape of the MoE head; and the
  column sums over a tall matrix (row slices + ticketed fold): exact against a float64 sum, identical run to run."""
  if not torch.cuda.is_available():
    pytest.skip("no CUDA device")
  import yt8m_native as nat
  g = torch.Generator().manual_seed(11)
  for b, rows, n in ((37, 256, 1000), (64, 2432, 128)):
    dz = torch.randn(b, rows, generator=g)
    w = synth.bf16r(torch.randn(rows, n, generator=g) / math.sqrt(rows))
    hi, lo = nat.split_bf16(dz.to(DEV))
    wb = torch.zeros((rows, nat.pad8(n)), dtype=torch.bfloat16, device=DEV)
    wb[:, :n] = w.to(DEV)
    got = nat.dgrad(hi, lo, wb, n)
    want = (hi.float() + lo.float()).cpu() @ w
    assert got.shape == (b, n)
    assert float((got.cpu() - want).abs().max()) < 2e-5 * float(want.abs().max()) + 1e-6
  rows, cols = 5000, 200
  m = torch.randn(rows, cols, generator=g)
  hi, lo = nat.split_bf16(m.to(DEV))
  want = (hi.double() + lo.double()).sum(dim=0).cpu()
  first = nat.colsum_bf16(hi, lo, cols).cpu()
  assert float((first.double() - want).abs().max()) < 1e-5 * float(want.abs().max()) + 1e-4
  for _ in range(3):                                       # the tickets clean up after themselves; fixed summation order
    assert torch.equal(nat.colsum_bf16(hi, lo, cols).cpu(), first)
  small = nat.colsum_bf16(hi[:100], lo[:100], cols).cpu()   # single-slice path
  assert float((small.double() - (hi[:100].double() + lo[:100].double()).sum(dim=0).cpu()).abs().max()) < 1e-4


# ------------------------------------------------------------------------------------------------
# the whole frame-level training step: NetVLAD + hidden FC + MoE head (BASELINE config 2, reduced sizes)
# ------------------------------------------------------------------------------------------------

def _netvlad_forward(params, x, nf, vocab, mixtures):
  k = params["cluster_weights"].shape[1]
  v = O.netvlad_pool(x, nf, params["cluster_weights"], torch.ones(k), params["cluster_biases"], params["cluster_weights2"])
  h = O.relu6(v @ params["hidden1_weights"] + params["hidden1_biases"])
  return O.moe_model(h, params["gates/weights"], params["experts/weights"], params["experts/biases"], vocab, mixtures)


def test_netvlad_train_step_parity(tr):
  g = torch.Generator().manual_seed(90)
  b, t, d, k, h, v, mix = 6, 300, 1152, 64, 256, 500, 2
  x, nf, _ = synth.model_input(b, t, d, seed=34)
  y = synth.labels(b, v, seed=34, per_video=3.4)
  sd = {"cluster_weights": synth.normal((d, k), g, 4.0), "cluster_biases": 0.1 * torch.randn(k, generator=g),
        "cluster_weights2": synth.normal((d, k), g, 1 / math.sqrt(d)),
        "hidden1_weights": synth.normal((k * d, h), g, 12.0 / math.sqrt(k)), "hidden1_biases": 0.1 * torch.randn(h, generator=g),
        "gates/weights": synth.xavier((h, v * (mix + 1)), g, 2.0), "experts/weights": synth.xavier((h, v * mix), g, 2.0),
        "experts/biases": 0.1 * torch.randn(v * mix, generator=g)}
  t_ = tr.NetVLADTrainer(d, clusters=k, hidden=h, vocab=v, mixtures=mix)
  t_.import_state(sd)
  t_.keep_grads = True
  xd, nfd, yd = x.to(DEV).to(torch.bfloat16), nf.to(DEV), y.to(DEV)
  p0 = t_.step(xd, nfd, yd)
  grad0 = t_.grads_tf_layout(t_.last_grad)
  loss0 = float(t_.last["label_loss_local"])
  # oracle: autograd over the fp32 forward
  params = {kk: w.clone().requires_grad_(True) for kk, w in sd.items()}
  pw = _netvlad_forward(params, x, nf, v, mix)
  lw = O.cross_entropy_loss(pw, y)
  gw = dict(zip(params, torch.autograd.grad(lw, list(params.values()))))
  assert float((p0.cpu() - pw.detach()).abs().max()) < 1e-3, float(pw.max())
  assert abs(loss0 - float(lw)) / float(lw) < 1e-3
  for kk in gw:
    assert float(gw[kk].norm()) > 0, kk
    err = _rel_l2(grad0[kk], gw[kk])
    assert err < 2e-2, (kk, err)
  # two more steps: every tensor moves, the bf16 operand copies follow the masters, nothing becomes non-finite
  for _ in range(2):
    t_.step(xd, nfd, yd)
  torch.cuda.synchronize()
  got = t_.export_state()
  for kk in sd:
    assert bool(torch.isfinite(got[kk]).all()), kk
    assert kk == "input_bn/beta" or float((got[kk] - sd[kk]).abs().max()) > 0, kk
  assert torch.equal(t_.cw_bf16.float(), t_.p["cw"].to(torch.bfloat16).float())
  assert torch.equal(t_.wfc_bf16.float(), t_.p["wfc"].to(torch.bfloat16).float())
  # Adam's first step moves every weight by ~lr * sign(g): check the displacement of the first step's direction
  moved = got["cluster_weights2"] - sd["cluster_weights2"]
  agree = (torch.sign(moved) == -torch.sign(gw["cluster_weights2"])).float().mean()
  assert float(agree) > 0.9


# ------------------------------------------------------------------------------------------------
# backward of the sequence poolers: LSTM (BPTT), attention pooling, context gating
# ------------------------------------------------------------------------------------------------

def _lstm_layers(d, h, layers, g, gain=2.0):
  ws = []
  for l in range(layers):
    in_dim = d if l == 0 else h
    ws.append((synth.xavier((in_dim + h, 4 * h), g, gain), synth.bf16r(0.1 * torch.randn(4 * h, generator=g))))
  return ws


@pytest.mark.parametrize("b,t,d,h,layers", [(5, 9, 64, 256, 2), (3, 17, 128, 256, 1), (66, 6, 64, 256, 2)])
def test_lstm_bwd_parity(b, t, d, h, layers):
  """yt8m_lstm_fwd_train + yt8m_lstm_bwd against autograd through the oracle's dynamic_rnn_lstm, with cotangents on
  BOTH the final state and the top-layer output sequence; includes rows frozen early and a one-frame video."""
  if not torch.cuda.is_available():
    pytest.skip("no CUDA device")
  import yt8m_native as nat
  import yt8m_trainer as tr
  g = torch.Generator().manual_seed(17 + b)
  x = synth.bf16r(torch.randn(b, t, d, generator=g) * 0.5)
  nf = torch.randint(1, t + 1, (b,), generator=g, dtype=torch.int32)
  nf[0], nf[1] = t, 1
  x = x * (torch.arange(t).unsqueeze(0) < nf.unsqueeze(1)).float().unsqueeze(2)
  ws = _lstm_layers(d, h, layers, g)
  r_state = torch.randn(b, layers * 2 * h, generator=g)
  r_seq = torch.randn(b, t, h, generator=g) * 0.3
  # oracle
  params = [(w.clone().requires_grad_(True), bb.clone().requires_grad_(True)) for w, bb in ws]
  outs, states = O.dynamic_rnn_lstm(x, nf, params)
  st = O.lstm_model_state(states)
  loss = (st * r_state).sum() + (outs * r_seq).sum()
  flat = [v for pr in params for v in pr]
  grads = torch.autograd.grad(loss, flat)
  # CUDA
  xd, nfd = x.to(DEV).to(torch.bfloat16), nf.to(DEV)
  packed = [nat.lstm_pack(w.to(DEV), bb.to(DEV), d if l == 0 else h, h) for l, (w, bb) in enumerate(ws)]
  wp, bp = [p[0] for p in packed], [p[1] for p in packed]
  state, seq, seq_hi, seq_lo = nat.lstm_fwd_train(xd, nfd, wp, bp, h, want_seq=True)
  assert float((state.cpu() - st.detach()).abs().max()) < 1e-4
  assert float((seq.cpu() - outs.detach()).abs().max()) < 1e-4
  assert float(((seq_hi[-1].float() + seq_lo[-1].float()).cpu() - outs.detach()).abs().max()) < 1e-4
  wt = [nat.pack_transpose(tr._lstm_tf_to_packed(w.to(DEV), h)) for w, _ in ws]
  dw, db = nat.lstm_bwd(xd, nfd, wp, bp, wt, h, seq_hi, seq_lo, dstate=r_state.to(DEV), dout_seq=r_seq.to(DEV))
  torch.cuda.synchronize()
  for l in range(layers):
    gw, gb = grads[2 * l], grads[2 * l + 1]
    got_w = tr._lstm_packed_to_tf(dw[l], h).cpu()
    got_b = db[l].view(h, 4).t().reshape(-1).cpu()
    assert float(gw.norm()) > 0
    # tolerance: the recurrent operand of the weight-gradient GEMM is the bf16 hi half of h (2^-9 relative per element)
    assert _rel_l2(got_w, gw) < 5e-3, (l, _rel_l2(got_w, gw))
    assert _rel_l2(got_b, gb) < 2e-3, (l, _rel_l2(got_b, gb))


@pytest.mark.parametrize("mode,use_nf", [(0, True), (1, True), (0, False)])
def test_attn_pool_bwd_parity(mode, use_nf):
  if not torch.cuda.is_available():
    pytest.skip("no CUDA device")
  import yt8m_native as nat
  g = torch.Generator().manual_seed(23 + mode)
  b, t, a, f = 4, 37, 8, 136
  nf = torch.tensor([37, 5, 20, 1], dtype=torch.int32)
  mask = (torch.arange(t).unsqueeze(0) < nf.unsqueeze(1)).float()
  feats = synth.bf16r(torch.randn(b, t, f, generator=g)) * mask.unsqueeze(2)
  logits = torch.randn(b, t, a, generator=g) * 2
  dout = torch.randn(b, a, f, generator=g)
  lg = logits.clone().requires_grad_(True)
  ft = feats.clone().requires_grad_(True)
  if mode == 0:
    m = mask if use_nf else (ft.detach().abs().sum(dim=2) > 0).float()
    w = torch.softmax(lg, dim=1) * m.unsqueeze(2)
    w = w / w.sum(dim=1, keepdim=True)
  else:
    w = torch.sigmoid(lg) * mask.unsqueeze(2)
    w = w / (w.sum(dim=1, keepdim=True) + 1e-8)
  out = torch.einsum("bta,btf->baf", w, ft)
  gl, gf = torch.autograd.grad((out * dout).sum(), [lg, ft])
  dl, df = nat.attn_pool_bwd(logits.to(DEV), feats.to(DEV).to(torch.bfloat16), nf.to(DEV) if use_nf else None, a, mode, dout.to(DEV),
                             want_dfeats=True)
  fwd = nat.attn_pool(logits.to(DEV), feats.to(DEV).to(torch.bfloat16), nf.to(DEV) if use_nf else None, a, mode)[0]
  assert float((fwd.cpu() - out.detach()).abs().max()) < 1e-4
  assert float((dl.cpu() - gl).abs().max()) < 1e-4 * max(1.0, float(gl.abs().max()))
  assert float((df.cpu() - gf).abs().max()) < 1e-4 * max(1.0, float(gf.abs().max()))
  # padded frames carry no gradient
  assert float(dl.cpu()[1, 5:].abs().max()) == 0.0


def test_context_gate_bwd_parity():
  if not torch.cuda.is_available():
    pytest.skip("no CUDA device")
  import yt8m_native as nat
  g = torch.Generator().manual_seed(29)
  rows, cols = 19, 100
  x = torch.randn(rows, cols, generator=g).requires_grad_(True)
  gg = torch.randn(rows, cols, generator=g).requires_grad_(True)
  sc = 1.0 + 0.1 * torch.randn(cols, generator=g)
  sh = 0.1 * torch.randn(cols, generator=g)
  dy = torch.randn(rows, cols, generator=g)
  y = x * torch.sigmoid(gg * sc + sh)
  # the direct path only: dL/dx through the gate input g is the caller's dgrad GEMM
  gx, gg_ = torch.autograd.grad((y * dy).sum(), [x, gg])
  dx, dg, dgh, dgl = nat.context_gate_bwd(dy.to(DEV), x.detach().to(DEV), gg.detach().to(DEV), sc.to(DEV), sh.to(DEV))
  assert float((dx.cpu() - gx).abs().max()) < 1e-5
  assert float((dg.cpu() - gg_).abs().max()) < 1e-5
  assert float(((dgh.float() + dgl.float()).cpu() - gg_).abs().max()) < 1e-4 * float(gg_.abs().max())


@pytest.mark.parametrize("memory", [False, True])
def test_lstm_train_step_parity(tr, memory):
  """LstmModel / LstmMemoryModel + MoE head: predictions, loss and every gradient of one step against autograd over the
  oracle (BASELINE config 3 at reduced sizes), then two more steps move every tensor."""
  g = torch.Generator().manual_seed(91)
  b, t, d, h, layers, v, mix = 6, 24, 128, 256, 2, 300, 2
  x, nf, _ = synth.model_input(b, t, d, seed=35, min_frames=3)
  y = synth.labels(b, v, seed=35, per_video=3.4)
  ws = _lstm_layers(d, h, layers, g)
  head_in = layers * h if memory else layers * 2 * h
  sd = {"gates/weights": synth.xavier((head_in, v * (mix + 1)), g, 2.0), "experts/weights": synth.xavier((head_in, v * mix), g, 2.0),
        "experts/biases": 0.1 * torch.randn(v * mix, generator=g)}
  for l, (w, bb) in enumerate(ws):
    sd[tr.LstmTrainer.SCOPE % l + "/weights"] = w
    sd[tr.LstmTrainer.SCOPE % l + "/biases"] = bb
  t_ = tr.LstmTrainer(d, hidden=h, layers=layers, vocab=v, mixtures=mix, memory=memory)
  t_.import_state(sd)
  t_.keep_grads = True
  xd, nfd, yd = x.to(DEV).to(torch.bfloat16), nf.to(DEV), y.to(DEV)
  p0 = t_.step(xd, nfd, yd)
  grad0 = t_.grads_tf_layout(t_.last_grad)
  loss0 = float(t_.last["label_loss_local"])
  params = {kk: w.clone().requires_grad_(True) for kk, w in sd.items()}
  lw_ = [(params[tr.LstmTrainer.SCOPE % l + "/weights"], params[tr.LstmTrainer.SCOPE % l + "/biases"]) for l in range(layers)]
  _, states = O.dynamic_rnn_lstm(x, nf, lw_)
  feat = O.lstm_memory_model_state(states) if memory else O.lstm_model_state(states)
  pw = O.moe_model(feat, params["gates/weights"], params["experts/weights"], params["experts/biases"], v, mix)
  lw = O.cross_entropy_loss(pw, y)
  gw = dict(zip(params, torch.autograd.grad(lw, list(params.values()))))
  assert float((p0.cpu() - pw.detach()).abs().max()) < 1e-3
  assert abs(loss0 - float(lw)) / float(lw) < 1e-3
  for kk in gw:
    assert float(gw[kk].norm()) > 0, kk
    err = _rel_l2(grad0[kk], gw[kk])
    assert err < 2e-2, (kk, err)
  for _ in range(2):
    t_.step(xd, nfd, yd)
  torch.cuda.synchronize()
  got = t_.export_state()
  for kk in sd:
    assert bool(torch.isfinite(got[kk]).all()), kk
    assert kk == "input_bn/beta" or float((got[kk] - sd[kk]).abs().max()) > 0, kk
  for l in range(layers):
    assert torch.equal(t_.w_bf16[l].float(), t_.p["w%d" % l].to(torch.bfloat16).float())


def test_gated_netvlad_train_step_parity(tr):
  """GatedNetVLADModel (K = 128 clusters, context gating, MoE-4: BASELINE config 4 at reduced sizes): predictions, loss
  and every gradient of one step against autograd over the oracle; then two more steps move every tensor."""
  g = torch.Generator().manual_seed(92)
  b, t, d, k, h, v, mix = 5, 100, 256, 128, 256, 300, 4
  x, nf, _ = synth.model_input(b, t, d, seed=36, min_frames=10)
  y = synth.labels(b, v, seed=36, per_video=3.4)
  sd = {"cluster_weights": synth.normal((d, k), g, 4.0), "cluster_biases": 0.1 * torch.randn(k, generator=g),
        "cluster_weights2": synth.normal((d, k), g, 1 / math.sqrt(d)),
        "hidden1_weights": synth.normal((k * d, h), g, 12.0 / math.sqrt(k)), "hidden1_biases": 0.1 * torch.randn(h, generator=g),
        "gating_weights": synth.normal((h, h), g, 1.0 / math.sqrt(h)), "gating_biases": 0.1 * torch.randn(h, generator=g),
        "gates/weights": synth.xavier((h, v * (mix + 1)), g, 2.0), "experts/weights": synth.xavier((h, v * mix), g, 2.0),
        "experts/biases": 0.1 * torch.randn(v * mix, generator=g)}
  t_ = tr.NetVLADTrainer(d, clusters=k, hidden=h, vocab=v, mixtures=mix, gating=True)
  t_.import_state(sd)
  t_.keep_grads = True
  xd, nfd, yd = x.to(DEV).to(torch.bfloat16), nf.to(DEV), y.to(DEV)
  p0 = t_.step(xd, nfd, yd)
  grad0 = t_.grads_tf_layout(t_.last_grad)
  loss0 = float(t_.last["label_loss_local"])
  params = {kk: w.clone().requires_grad_(True) for kk, w in sd.items()}
  vl = O.netvlad_pool(x, nf, params["cluster_weights"], torch.ones(k), params["cluster_biases"], params["cluster_weights2"])
  hid = O.relu6(vl @ params["hidden1_weights"] + params["hidden1_biases"])
  gated = O.context_gating(hid, params["gating_weights"], torch.ones(h), params["gating_biases"])
  pw = O.moe_model(gated, params["gates/weights"], params["experts/weights"], params["experts/biases"], v, mix)
  lw = O.cross_entropy_loss(pw, y)
  gw = dict(zip(params, torch.autograd.grad(lw, list(params.values()))))
  assert float((p0.cpu() - pw.detach()).abs().max()) < 1e-3
  assert abs(loss0 - float(lw.detach())) / float(lw.detach()) < 1e-3
  for kk in gw:
    assert float(gw[kk].norm()) > 0, kk
    err = _rel_l2(grad0[kk], gw[kk])
    assert err < 2e-2, (kk, err)
  for _ in range(2):
    t_.step(xd, nfd, yd)
  torch.cuda.synchronize()
  got = t_.export_state()
  for kk in sd:
    assert bool(torch.isfinite(got[kk]).all()), kk
    assert kk == "input_bn/beta" or float((got[kk] - sd[kk]).abs().max()) > 0, kk
  assert torch.equal(t_.wg_bf16.float(), t_.p["wg"].to(torch.bfloat16).float())


def test_group_max_rows_bwd():
  if not torch.cuda.is_available():
    pytest.skip("no CUDA device")
  import yt8m_native as nat
  g = torch.Generator().manual_seed(31)
  groups, heads, cols = 7, 8, 45
  x = torch.randn(groups * heads, cols, generator=g).requires_grad_(True)
  dout = torch.randn(groups, cols, generator=g)
  y = x.reshape(groups, heads, cols).max(dim=1).values
  gx, = torch.autograd.grad((y * dout).sum(), [x])
  got = nat.group_max_rows_bwd(x.detach().to(DEV), dout.to(DEV), heads)
  assert torch.equal(got.cpu(), gx)


def test_attention_train_step_parity(tr):
  """zt AttentionModel (A = 8 heads over the raw frames) + MoeExtendModel (MoE on B*A rows, max over heads): predictions,
  loss and every gradient of one step against autograd over the oracle (BASELINE config 5 pooling at reduced sizes)."""
  g = torch.Generator().manual_seed(93)
  b, t, d, a, v, mix = 5, 40, 128, 8, 300, 2
  x, nf, _ = synth.model_input(b, t, d, seed=37, min_frames=5)
  y = synth.labels(b, v, seed=37, per_video=3.4)
  sd = {"Attention/W": synth.bf16r(torch.randn(2 * d, a, generator=g) * 3.0), "Attention/b": torch.full((a,), 0.1),
        "gates/weights": synth.xavier((d, v * (mix + 1)), g, 6.0), "experts/weights": synth.xavier((d, v * mix), g, 6.0),
        "experts/biases": 0.1 * torch.randn(v * mix, generator=g)}
  t_ = tr.AttentionTrainer(d, heads=a, vocab=v, mixtures=mix)
  t_.import_state(sd)
  t_.keep_grads = True
  xd, nfd, yd = x.to(DEV).to(torch.bfloat16), nf.to(DEV), y.to(DEV)
  p0 = t_.step(xd, nfd, yd)
  grad0 = t_.grads_tf_layout(t_.last_grad)
  loss0 = float(t_.last["label_loss_local"])
  params = {kk: w.clone().requires_grad_(True) for kk, w in sd.items()}
  pooled = O.attention_model_pool(x, nf, params["Attention/W"], params["Attention/b"])
  pw = O.moe_extend_model(pooled, params["gates/weights"], params["experts/weights"], params["experts/biases"], v, mix, a)
  lw = O.cross_entropy_loss(pw, y)
  gw = dict(zip(params, torch.autograd.grad(lw, list(params.values()))))
  assert float((p0.cpu() - pw.detach()).abs().max()) < 1e-3
  assert abs(loss0 - float(lw.detach())) / float(lw.detach()) < 1e-3
  # softmax over T is shift invariant: the mean-pooled half of W and the bias get no label gradient
  assert float(gw["Attention/b"].abs().max()) < 1e-6 and float(grad0["Attention/b"].abs().max()) == 0.0
  assert float(gw["Attention/W"][d:].abs().max()) < 1e-6 and float(grad0["Attention/W"][d:].abs().max()) == 0.0
  for kk in gw:
    if kk == "Attention/b":
      continue
    assert float(gw[kk].norm()) > 0, kk
    err = _rel_l2(grad0[kk], gw[kk])
    assert err < 2e-2, (kk, err)
  for _ in range(2):
    t_.step(xd, nfd, yd)
  torch.cuda.synchronize()
  got = t_.export_state()
  for kk in sd:                      # every tensor moves: the shift-invariant ones through their L2 regulariser + Adam
    assert bool(torch.isfinite(got[kk]).all()), kk
    assert float((got[kk] - sd[kk]).abs().max()) > 0, kk


def test_chain_moe_train_step_parity(tr):
  """ChainMoeModel (support MoE -> concat -> main MoE, wh/all_video_models/chain_moe_model.py:9-49) without --multitask:
  predictions, loss and the gradients of BOTH heads against autograd over the oracle; input width D + S = 153 is not a
  multiple of 8 (padded operands)."""
  g = torch.Generator().manual_seed(94)
  b, d, v, s, mix = 12, 128, 300, 25, 2
  x, y, _ = _data(b, d, v, 38)
  sd = {"gates-support/weights": synth.xavier((d, s * (mix + 1)), g, 6.0), "experts-support/weights": synth.xavier((d, s * mix), g, 6.0),
        "experts-support/biases": 0.1 * torch.randn(s * mix, generator=g),
        "gates-main/weights": synth.xavier((d + s, v * (mix + 1)), g, 6.0), "experts-main/weights": synth.xavier((d + s, v * mix), g, 6.0),
        "experts-main/biases": 0.1 * torch.randn(v * mix, generator=g)}
  t_ = tr.ChainMoeTrainer(d, vocab=v, mixtures=mix, num_supports=s)
  t_.import_state({k: w.to(DEV) for k, w in sd.items()})
  t_.keep_grads = True
  p0 = t_.step(x.to(DEV), y.to(DEV))
  grad0 = t_.grads_tf_layout(t_.last_grad)
  loss0 = float(t_.last["label_loss_local"])
  params = {kk: w.clone().requires_grad_(True) for kk, w in sd.items()}
  sup = {"gate_w": params["gates-support/weights"], "expert_w": params["experts-support/weights"], "expert_b": params["experts-support/biases"]}
  main = {"gate_w": params["gates-main/weights"], "expert_w": params["experts-main/weights"], "expert_b": params["experts-main/biases"]}
  pw, _ = O.chain_moe_model(x, sup, main, v, s, mix)
  lw = O.cross_entropy_loss(pw, y)
  gw = dict(zip(params, torch.autograd.grad(lw, list(params.values()))))
  assert float((p0.cpu() - pw.detach()).abs().max()) < 1e-3
  assert abs(loss0 - float(lw.detach())) / float(lw.detach()) < 1e-3
  for kk in gw:
    assert float(gw[kk].norm()) > 0, kk
    err = _rel_l2(grad0[kk], gw[kk])
    assert err < 2e-2, (kk, err)
  for _ in range(2):
    t_.step(x.to(DEV), y.to(DEV))
  torch.cuda.synchronize()
  got = t_.export_state()
  for kk in sd:
    assert bool(torch.isfinite(got[kk]).all()), kk
    assert kk == "input_bn/beta" or float((got[kk] - sd[kk]).abs().max()) > 0, kk


@pytest.mark.parametrize("kind", ["max_pooling", "multi"])
def test_lstm_attention_train_step_parity(tr, kind):
  """LstmAttentionMaxPoolingModel / LstmMultiAttentionModel: one training step (LSTM forward that retains the sequences,
  attention pooling, MoE per head, max over heads, and the whole backward incl. BPTT driven by the gradient of the
  OUTPUT SEQUENCE) against autograd over the oracle's whole-model forward."""
  from oracle import model_oracle as MO
  g = torch.Generator().manual_seed(95)
  b, t, d, h, layers, a, v, mix = 4, 16, 64, 256, 2, 8, 200, 2
  x, nf, _ = synth.model_input(b, t, d, seed=39, min_frames=3)
  y = synth.labels(b, v, seed=39, per_video=3.4)
  sd = {}
  for l, (w, bb) in enumerate(_lstm_layers(d, h, layers, g)):
    sd[tr.LstmTrainer.SCOPE % l + "/weights"], sd[tr.LstmTrainer.SCOPE % l + "/biases"] = w, bb
  if kind == "max_pooling":
    att, gn, en, pool = "attention-", "gates-sub-moe", "experts-sub-moe", h
    sd[att + "/weights"] = synth.xavier((d + h, a), g, 8.0)
  else:
    att, gn, en, pool = "fully_connected", "gates", "experts", d
    sd[att + "/weights"] = synth.xavier((h, a), g, 8.0)
  sd[att + "/biases"] = 0.1 * torch.randn(a, generator=g)
  sd[gn + "/weights"] = synth.xavier((pool, v * (mix + 1)), g, 6.0)
  sd[en + "/weights"] = synth.xavier((pool, v * mix), g, 6.0)
  sd[en + "/biases"] = 0.1 * torch.randn(v * mix, generator=g)
  t_ = tr.LstmAttentionTrainer(d, hidden=h, layers=layers, heads=a, vocab=v, mixtures=mix, kind=kind)
  t_.import_state(sd)
  t_.keep_grads = True
  xd, nfd, yd = x.to(DEV).to(torch.bfloat16), nf.to(DEV), y.to(DEV)
  p0 = t_.step(xd, nfd, yd)
  grad0 = t_.grads_tf_layout(t_.last_grad)
  loss0 = float(t_.last["label_loss_local"])
  params = {kk: w.clone().requires_grad_(True) for kk, w in sd.items()}
  fwd = MO.lstm_attention_max_pooling if kind == "max_pooling" else MO.lstm_multi_attention
  pw = fwd(params, x, nf, v, mix, a, layers=layers)
  lw = O.cross_entropy_loss(pw, y)
  gw = dict(zip(params, torch.autograd.grad(lw, list(params.values()))))
  assert float((p0.cpu() - pw.detach()).abs().max()) < 2e-3      # the attention reads the bf16 copy of h (as the forward plugin)
  assert abs(loss0 - float(lw.detach())) / float(lw.detach()) < 2e-3
  for kk in gw:
    if kind == "max_pooling" and kk == att + "/biases":
      # softmax over T is shift invariant: the bias gradient is zero up to rounding in both implementations
      scale = float(gw[att + "/weights"].abs().max())
      assert float(gw[kk].abs().max()) < 1e-5 * scale and float(grad0[kk].abs().max()) < 1e-3 * scale, kk
      continue
    assert float(gw[kk].norm()) > 0, kk
    err = _rel_l2(grad0[kk], gw[kk])
    assert err < 3e-2, (kk, err)
  for _ in range(2):
    t_.step(xd, nfd, yd)
  torch.cuda.synchronize()
  got = t_.export_state()
  for kk in sd:
    assert bool(torch.isfinite(got[kk]).all()), kk
    if not (kind == "max_pooling" and kk == att + "/biases"):
      assert float((got[kk] - sd[kk]).abs().max()) > 0, kk


def test_l2norm_rows_bwd():
  if not torch.cuda.is_available():
    pytest.skip("no CUDA device")
  import yt8m_native as nat
  g = torch.Generator().manual_seed(41)
  x = torch.randn(9, 72, generator=g).requires_grad_(True)
  dy = torch.randn(9, 72, generator=g)
  y = O.l2_normalize(x, dim=1)
  gx, = torch.autograd.grad((y * dy).sum(), [x])
  got = nat.l2norm_rows_bwd(x.detach().to(DEV), dy.to(DEV))
  assert float((got.cpu() - gx).abs().max()) < 1e-5 * float(gx.abs().max())


def test_deep_combine_chain_train_step_parity(tr):
  """DeepCombineChainModel (wh/all_video_models/deep_combine_chain_model.py:24-49) without --multitask: 2 stacked MoE
  sub-predictions -> projection -> ReLU -> L2-normalise -> concat, main MoE; every gradient against autograd over the oracle."""
  g = torch.Generator().manual_seed(96)
  b, d, v, mix, layers, r = 10, 128, 300, 2, 2, 64
  x, y, _ = _data(b, d, v, 40)
  sd = {}
  for i in range(layers + 1):
    sc = "-prediction-%d" % i if i < layers else "--main"
    din = d + i * r
    sd["gates%s/weights" % sc] = synth.xavier((din, v * (mix + 1)), g, 6.0)
    sd["experts%s/weights" % sc] = synth.xavier((din, v * mix), g, 6.0)
    sd["experts%s/biases" % sc] = 0.1 * torch.randn(v * mix, generator=g)
  for i in range(layers):
    sd["relu-%d/weights" % i] = synth.xavier((v, r), g, 4.0)
    sd["relu-%d/biases" % i] = 0.1 * torch.randn(r, generator=g)
  t_ = tr.DeepCombineChainTrainer(d, vocab=v, mixtures=mix, layers=layers, relu_cells=r)
  t_.import_state({k: w.to(DEV) for k, w in sd.items()})
  t_.keep_grads = True
  p0 = t_.step(x.to(DEV), y.to(DEV))
  grad0 = t_.grads_tf_layout(t_.last_grad)
  loss0 = float(t_.last["label_loss_local"])
  params = {kk: w.clone().requires_grad_(True) for kk, w in sd.items()}
  lyr = [{"gate_w": params["gates-prediction-%d/weights" % i], "expert_w": params["experts-prediction-%d/weights" % i],
          "expert_b": params["experts-prediction-%d/biases" % i], "relu_w": params["relu-%d/weights" % i],
          "relu_b": params["relu-%d/biases" % i]} for i in range(layers)]
  main = {"gate_w": params["gates--main/weights"], "expert_w": params["experts--main/weights"], "expert_b": params["experts--main/biases"]}
  pw, _ = O.deep_combine_chain_model(x, lyr, main, v, mix)
  lw = O.cross_entropy_loss(pw, y)
  gw = dict(zip(params, torch.autograd.grad(lw, list(params.values()))))
  assert float((p0.cpu() - pw.detach()).abs().max()) < 1e-3
  assert abs(loss0 - float(lw.detach())) / float(lw.detach()) < 1e-3
  for kk in gw:
    assert float(gw[kk].norm()) > 0, kk
    err = _rel_l2(grad0[kk], gw[kk])
    assert err < 2e-2, (kk, err)
  for _ in range(2):
    t_.step(x.to(DEV), y.to(DEV))
  torch.cuda.synchronize()
  got = t_.export_state()
  for kk in sd:
    assert bool(torch.isfinite(got[kk]).all()), kk
    assert kk == "input_bn/beta" or float((got[kk] - sd[kk]).abs().max()) > 0, kk


def test_dbof_train_step_parity(tr):
  """DbofModel in its bias form (--dbof_add_batch_norm=False, max pooling) + MoE: pinned frame sample, every gradient of one
  step against autograd over the oracle (wh/all_frame_models/dbof_model.py:62-123)."""
  g = torch.Generator().manual_seed(97)
  b, t, d, c, h, n, v, mix = 6, 50, 128, 512, 256, 10, 300, 2
  x, nf, _ = synth.model_input(b, t, d, seed=41, min_frames=12)
  y = synth.labels(b, v, seed=41, per_video=3.4)
  fidx = (torch.rand(b, n, generator=g) * nf.float().unsqueeze(1)).to(torch.int64)
  sd = {"cluster_weights": synth.normal((d, c), g, 3.0), "cluster_biases": 0.1 * torch.randn(c, generator=g),
        "hidden1_weights": synth.normal((c, h), g, 1.5 / math.sqrt(c)), "hidden1_biases": 0.1 * torch.randn(h, generator=g),
        "gates/weights": synth.xavier((h, v * (mix + 1)), g, 4.0), "experts/weights": synth.xavier((h, v * mix), g, 4.0),
        "experts/biases": 0.1 * torch.randn(v * mix, generator=g)}
  t_ = tr.DbofTrainer(d, cluster_size=c, hidden=h, iterations=n, vocab=v, mixtures=mix)
  t_.import_state(sd)
  t_.keep_grads = True
  xd, nfd, yd = x.to(DEV).to(torch.bfloat16), nf.to(DEV), y.to(DEV)
  p0 = t_.step(xd, nfd, yd, frame_index=fidx)
  grad0 = t_.grads_tf_layout(t_.last_grad)
  loss0 = float(t_.last["label_loss_local"])
  params = {kk: w.clone().requires_grad_(True) for kk, w in sd.items()}
  pp = {"cluster_w": params["cluster_weights"], "cluster_b": params["cluster_biases"], "hidden_w": params["hidden1_weights"],
        "hidden_b": params["hidden1_biases"]}
  hid = O.dbof_pool(x, fidx, pp, add_batch_norm=False, pooling="max")
  pw = O.moe_model(hid, params["gates/weights"], params["experts/weights"], params["experts/biases"], v, mix)
  lw = O.cross_entropy_loss(pw, y)
  gw = dict(zip(params, torch.autograd.grad(lw, list(params.values()))))
  assert float((p0.cpu() - pw.detach()).abs().max()) < 1e-3
  assert abs(loss0 - float(lw.detach())) / float(lw.detach()) < 1e-3
  for kk in gw:
    assert float(gw[kk].norm()) > 0, kk
    err = _rel_l2(grad0[kk], gw[kk])
    assert err < 2e-2, (kk, err)
  for _ in range(2):
    t_.step(xd, nfd, yd, frame_index=fidx)
  torch.cuda.synchronize()
  got = t_.export_state()
  for kk in sd:
    assert bool(torch.isfinite(got[kk]).all()), kk
    assert kk == "input_bn/beta" or float((got[kk] - sd[kk]).abs().max()) > 0, kk


def test_batch_norm_training_kernels_match_autograd():
  """yt8m_bn_stats / yt8m_bn_fold / yt8m_col_affine_act / yt8m_bn_bwd against autograd over oracle.batch_norm(is_training=True)
  (slim.batch_norm at wh/all_frame_models/dbof_model.py:64-108): forward value, the moving-average update, and dgamma / dbeta /
  dx THROUGH the batch statistics, with and without the ReLU6 that follows; fp32 and bf16 inputs, a column count that is not a
  multiple of 32."""
  import yt8m_native as nat
  g = torch.Generator().manual_seed(5)
  for rows, cols, act, bf in ((300, 72, "relu6", False), (64, 200, None, True), (1000, 33, "relu6", False)):
    x = torch.randn(rows, cols, generator=g) * 2.0 + 0.5
    if bf:
      x = synth.bf16r(x)
    gamma, beta = 1.0 + 0.2 * torch.randn(cols, generator=g), 0.3 * torch.randn(cols, generator=g) + (2.0 if act else 0.0)
    mm, mv = torch.randn(cols, generator=g), torch.rand(cols, generator=g) + 0.5
    dy = torch.randn(rows, cols, generator=g)
    xr, gr, br = x.clone().requires_grad_(True), gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    yw, mm_w, mv_w = O.batch_norm(xr, gr, br, mm, mv, True)
    if act:
      yw = O.relu6(yw)
    dxw, dgw, dbw = torch.autograd.grad((yw * dy).sum(), [xr, gr, br])
    xd = x.to(DEV).to(torch.bfloat16) if bf else x.to(DEV)
    mmd, mvd = mm.to(DEV), mv.to(DEV)
    out, stats = nat.bn_train_fwd(xd, gamma.to(DEV), beta.to(DEV), mmd, mvd, act=act, want_bf16=True)
    assert float((out["f32"].cpu() - yw.detach()).abs().max()) < 2e-5
    assert float(((out["hi"].float() + out["lo"].float()).cpu() - yw.detach()).abs().max()) < 1e-4
    assert float((mmd.cpu() - mm_w).abs().max()) < 1e-6 and float((mvd.cpu() - mv_w).abs().max()) < 1e-5
    dg, db, dh, dl, dxf = nat.bn_train_bwd(dy.to(DEV), out["f32"], xd, stats, gamma.to(DEV), act=act)
    assert _rel_l2(dg.cpu(), dgw) < 1e-4 and _rel_l2(db.cpu(), dbw) < 1e-4
    assert _rel_l2(dxf.cpu(), dxw) < 1e-4 and _rel_l2((dh.float() + dl.float()).cpu(), dxw) < 1e-4


def test_dbof_batch_norm_train_step_parity(tr):
  """DbofModel with its DEFAULT flags (--dbof_add_batch_norm=True, max pooling) + MoE in training mode: batch statistics in the
  three slim.batch_norm layers, every gradient of one step against autograd over the oracle, and the moving averages after
  the step against the oracle's update (wh/all_frame_models/dbof_model.py:62-123, wh/train.py:449-456)."""
  g = torch.Generator().manual_seed(98)
  b, t, d, c, h, n, v, mix = 8, 50, 128, 512, 256, 10, 300, 2
  x, nf, _ = synth.model_input(b, t, d, seed=42, min_frames=12)
  y = synth.labels(b, v, seed=42, per_video=3.4)
  fidx = (torch.rand(b, n, generator=g) * nf.float().unsqueeze(1)).to(torch.int64)
  sd = {"cluster_weights": synth.normal((d, c), g, 1.0 / math.sqrt(d)), "hidden1_weights": synth.normal((c, h), g, 1.5 / math.sqrt(c)),
        "gates/weights": synth.xavier((h, v * (mix + 1)), g, 4.0), "experts/weights": synth.xavier((h, v * mix), g, 4.0),
        "experts/biases": 0.1 * torch.randn(v * mix, generator=g)}
  for scope, width in (("input_bn", d), ("cluster_bn", c), ("hidden1_bn", h)):
    sd[scope + "/gamma"] = 1.0 + 0.2 * torch.randn(width, generator=g)
    sd[scope + "/beta"] = 0.2 * torch.randn(width, generator=g) + (1.5 if scope != "input_bn" else 0.0)
    sd[scope + "/moving_mean"] = 0.1 * torch.randn(width, generator=g)
    sd[scope + "/moving_variance"] = 0.5 + torch.rand(width, generator=g)
  t_ = tr.DbofTrainer(d, cluster_size=c, hidden=h, iterations=n, vocab=v, mixtures=mix, batch_norm=True)
  t_.import_state(sd)
  t_.keep_grads = True
  xd, nfd, yd = x.to(DEV).to(torch.bfloat16), nf.to(DEV), y.to(DEV)
  p0 = t_.step(xd, nfd, yd, frame_index=fidx)
  grad0 = t_.grads_tf_layout(t_.last_grad)
  loss0 = float(t_.last["label_loss_local"])
  trainable = [kk for kk in sd if "moving" not in kk]
  params = {kk: sd[kk].clone().requires_grad_(True) for kk in trainable}

  def bn(scope):
    return {"gamma": params[scope + "/gamma"], "beta": params[scope + "/beta"], "mean": sd[scope + "/moving_mean"],
            "var": sd[scope + "/moving_variance"]}
  pp = {"cluster_w": params["cluster_weights"], "hidden_w": params["hidden1_weights"], "input_bn": bn("input_bn"),
        "cluster_bn": bn("cluster_bn"), "hidden1_bn": bn("hidden1_bn")}
  hid = O.dbof_pool(x, fidx, pp, is_training=True, add_batch_norm=True, pooling="max")
  pw = O.moe_model(hid, params["gates/weights"], params["experts/weights"], params["experts/biases"], v, mix)
  lw = O.cross_entropy_loss(pw, y)
  gw = dict(zip(params, torch.autograd.grad(lw, list(params.values()))))
  assert float((p0.cpu() - pw.detach()).abs().max()) < 1e-3
  assert abs(loss0 - float(lw.detach())) / float(lw.detach()) < 1e-3
  for kk in gw:
    if kk == "input_bn/beta":
      # a uniform shift of the input rows is removed again by cluster_bn's mean subtraction: the true gradient is ZERO (autograd
      # returns rounding noise), so the bar is absolute, on the scale of the sibling gamma gradient
      assert float(grad0[kk].norm()) < 1e-3 * float(gw["input_bn/gamma"].norm()), (float(grad0[kk].norm()), float(gw["input_bn/gamma"].norm()))
      continue
    assert float(gw[kk].norm()) > 0, kk
    err = _rel_l2(grad0[kk], gw[kk])
    assert err < 2e-2, (kk, err)
  # moving averages: decay 0.999 towards the batch moments of the input rows (the first layer's statistics are data only)
  rows = x[torch.arange(b).unsqueeze(1), fidx].reshape(b * n, d)
  _, mm_w, mv_w = O.batch_norm(rows, sd["input_bn/gamma"], sd["input_bn/beta"], sd["input_bn/moving_mean"], sd["input_bn/moving_variance"], True)
  got = t_.export_state()
  assert float((got["input_bn/moving_mean"] - mm_w).abs().max()) < 1e-6
  assert float((got["input_bn/moving_variance"] - mv_w).abs().max()) < 1e-6
  for _ in range(2):
    t_.step(xd, nfd, yd, frame_index=fidx)
  torch.cuda.synchronize()
  got = t_.export_state()
  for kk in sd:
    assert bool(torch.isfinite(got[kk]).all()), kk
    assert kk == "input_bn/beta" or float((got[kk] - sd[kk]).abs().max()) > 0, kk
