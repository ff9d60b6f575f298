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
  dv, dasum, dcw2 = nat.netvlad_bwd_norm(dy.to(DEV), y32, stats, cw2.detach().to(DEV))
  z = nat.linear(xd.reshape(b * t, d), cwp, n=k, k=d, scale=sc, shift=shift.detach().to(DEV))["f32"]
  dz_hi, dz_lo, dshift = nat.netvlad_bwd_assign(xd, nf.to(DEV), z, dv, dasum, scale=sc)
  dcw_t = nat.wgrad(dz_hi, dz_lo, xd.reshape(b * t, d), k, d)                 # [K, D] = dL/dCw^T
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
    assert float((got[kk] - sd[kk]).abs().max()) > 0, kk
  assert torch.equal(t_.cw_bf16.float(), t_.p["cw"].to(torch.bfloat16).float())
  assert torch.equal(t_.wfc_bf16.float(), t_.p["wfc"].to(torch.bfloat16).float())
  # Adam's first step moves every weight by ~lr * sign(g): check the displacement of the first step's direction
  moved = got["cluster_weights2"] - sd["cluster_weights2"]
  agree = (torch.sign(moved) == -torch.sign(gw["cluster_weights2"])).float().mean()
  assert float(agree) > 0.9
