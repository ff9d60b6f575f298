"""Closed-form known-answer tests of the CPU oracle (SURVEY.md §4 lists them).  The model arithmetic of
the reference runs inside TensorFlow 1.0, which is neither vendored nor installable here, and the
reference has no tests, so these KATs (plus TF-1.0's documented op semantics) are what pins it."""
import math

import pytest
import torch

from oracle import yt8m_oracle as O


def test_moe_zero_weights():
  b, d, v, m = 5, 16, 7, 2
  p = O.moe_model(torch.randn(b, d), torch.zeros(d, v * (m + 1)), torch.zeros(d, v * m), torch.zeros(v * m), v, m)
  assert torch.allclose(p, torch.full((b, v), m / (m + 1) * 0.5))


def test_moe_column_layout_is_class_major_mixture_minor():
  # one class, two mixtures: gate logits (g0, g1, g_dummy), expert logits (e0, e1)
  x = torch.ones(1, 1)
  gw = torch.tensor([[1.0, 2.0, 3.0, 0.0, 0.0, 0.0]])          # V=2: class0=(1,2,3) class1=(0,0,0)
  ew = torch.tensor([[0.5, -0.5, 0.0, 0.0]])
  p = O.moe_model(x, gw, ew, torch.zeros(4), 2, 2)
  g = torch.softmax(torch.tensor([1.0, 2.0, 3.0]), 0)
  want0 = g[0] * torch.sigmoid(torch.tensor(0.5)) + g[1] * torch.sigmoid(torch.tensor(-0.5))
  assert abs(float(p[0, 0]) - float(want0)) < 1e-6
  assert abs(float(p[0, 1]) - (2 / 3) * 0.5) < 1e-6


def test_lstm_zero_weights_and_gate_order():
  b, d, h = 3, 4, 2
  c, hh = O.basic_lstm_cell(torch.randn(b, d), torch.zeros(b, h), torch.zeros(b, h), torch.zeros(d + h, 4 * h),
                            torch.zeros(4 * h))
  assert float(c.abs().max()) == 0 and float(hh.abs().max()) == 0
  # gate order i, j, f, o and forget_bias added at the use site: bias on j only -> c = sigmoid(0)*tanh(bj)
  bias = torch.zeros(4 * h)
  bias[h:2 * h] = 1.5
  c, hh = O.basic_lstm_cell(torch.zeros(1, d), torch.ones(1, h), torch.zeros(1, h), torch.zeros(d + h, 4 * h), bias)
  want_c = 1.0 * torch.sigmoid(torch.tensor(1.0)) + 0.5 * math.tanh(1.5)
  assert torch.allclose(c, torch.full((1, h), float(want_c)), atol=1e-6)
  assert torch.allclose(hh, torch.tanh(c) * 0.5, atol=1e-6)


def test_dynamic_rnn_freezes_state_and_zeroes_output():
  torch.manual_seed(0)
  b, t, d, h = 2, 6, 3, 4
  x = torch.randn(b, t, d)
  w = [(torch.randn(d + h, 4 * h) * 0.3, torch.randn(4 * h) * 0.1), (torch.randn(2 * h, 4 * h) * 0.3, torch.zeros(4 * h))]
  nf = torch.tensor([6, 3])
  outs, states = O.dynamic_rnn_lstm(x, nf, w)
  outs3, states3 = O.dynamic_rnn_lstm(x[:, :3], torch.tensor([3, 3]), w)
  assert float(outs[1, 3:].abs().max()) == 0.0
  for (c, hh), (c3, h3) in zip(states, states3):
    assert torch.allclose(c[1], c3[1]) and torch.allclose(hh[1], h3[1])
  assert O.lstm_model_state(states).shape == (b, 4 * h)          # [c0, h0, c1, h1]
  assert torch.equal(O.lstm_model_state(states)[:, h:2 * h], states[0][1])
  assert O.lstm_memory_model_state(states).shape == (b, 2 * h)


def test_attention_equal_logits_is_mean_pool():
  torch.manual_seed(1)
  b, t, d, a, h = 2, 7, 5, 3, 4
  x, outs = torch.randn(b, t, d), torch.randn(b, t, h)
  nf = torch.tensor([7, 4])
  pooled = O.attention_softmax_pool(x, outs, nf, torch.zeros(d + h, a), torch.zeros(a))
  assert torch.allclose(pooled[1, 0], outs[1, :4].mean(dim=0), atol=1e-6)
  sig = O.attention_sigmoid_pool(x, outs, nf, torch.zeros(h, a), torch.zeros(a))
  assert torch.allclose(sig[1, 2], x[1, :4].mean(dim=0), atol=1e-5)


def test_attention_model_ignores_mean_half_and_bias():
  """softmax over T is shift invariant: the [mean] half of W and b cannot change the result."""
  torch.manual_seed(2)
  b, t, d, a = 2, 9, 6, 4
  x = torch.randn(b, t, d)
  x[1, 5:] = 0                                                    # padded frames -> masked by |x| > 0
  nf = torch.tensor([9, 5])
  w = torch.randn(2 * d, a)
  s1 = O.attention_model_pool(x, nf, w, torch.full((a,), 0.1))
  w2 = w.clone()
  w2[d:] = torch.randn(d, a)
  s2 = O.attention_model_pool(x, nf, w2, torch.randn(a))
  assert torch.allclose(s1, s2, atol=1e-5)


def test_l2_normalize_zero_row_and_dequantize():
  x = torch.zeros(2, 8)
  x[1] = 2.0
  y = O.l2_normalize(x)
  assert float(y[0].abs().max()) == 0.0
  assert torch.allclose(y[1], torch.full((8,), 1 / math.sqrt(8)))
  q = O.dequantize(torch.tensor([0, 255], dtype=torch.uint8))
  assert torch.allclose(q, torch.tensor([4 / 512 - 2, 4 + 4 / 512 - 2]))


def test_cross_entropy_epsilon():
  p = torch.tensor([[1.0, 0.0]])
  y = torch.tensor([[1.0, 0.0]])
  assert abs(float(O.cross_entropy_loss(p, y)) - (-2 * math.log(1 + 1e-5))) < 1e-7
  assert abs(float(O.cross_entropy_loss(1 - p, y)) - (-2 * math.log(1e-5))) < 1e-4


def test_adam_clip_decay():
  g = torch.tensor([3.0, 4.0])
  assert torch.allclose(O.clip_by_norm(g, 1.0), g / 5)
  assert torch.allclose(O.clip_by_norm(g, 10.0), g)
  p, m, v = O.adam_step(torch.zeros(2), g, torch.zeros(2), torch.zeros(2), 1, 0.01)
  # first step: m = 0.1 g, v = 0.001 g^2, lr_t = lr * sqrt(0.001) / 0.1 -> p = -lr * g / (|g| + eps*sqrt(1000))
  want = -0.01 * math.sqrt(0.001) / 0.1 * (0.1 * g) / (torch.sqrt(0.001 * g * g) + 1e-8)
  assert torch.allclose(p, want)
  assert O.exponential_decay(0.01, 3999, 1000, 4000000, 0.95) == 0.01
  assert abs(O.exponential_decay(0.01, 4000, 1000, 4000000, 0.95) - 0.0095) < 1e-12


def test_netvlad_properties():
  torch.manual_seed(3)
  b, t, d, k = 2, 10, 8, 4
  x = O.l2_normalize(torch.randn(b, t, d))
  nf = torch.tensor([10, 6])
  cw, cw2 = torch.randn(d, k), torch.randn(d, k) * 0.1
  v = O.netvlad_pool(x, nf, cw, torch.ones(k), torch.zeros(k), cw2)
  assert v.shape == (b, d * k)
  assert torch.allclose(v.norm(dim=1), torch.ones(b), atol=1e-5)
  # frames past num_frames do not matter
  x2 = x.clone()
  x2[1, 6:] = torch.randn(4, d)
  assert torch.allclose(v, O.netvlad_pool(x2, nf, cw, torch.ones(k), torch.zeros(k), cw2), atol=1e-6)
  # flatten order is D-major, K-minor and every cluster column has norm 1/sqrt(K)
  cols = v[0].reshape(d, k)
  assert torch.allclose(cols.norm(dim=0), torch.full((k,), 1 / math.sqrt(k)), atol=1e-5)


def test_chain_models_shapes():
  torch.manual_seed(4)
  b, d, v, m = 3, 6, 5, 2
  mk = lambda din, vv: {"gate_w": torch.randn(din, vv * (m + 1)), "expert_w": torch.randn(din, vv * m), "expert_b": torch.zeros(vv * m)}
  p, sp = O.chain_moe_model(torch.randn(b, d), mk(d, 4), mk(d + 4, v), v, 4, m)
  assert p.shape == (b, v) and sp.shape == (b, 4)
  layers = []
  din = d
  for _ in range(2):
    l = mk(din, v)
    l.update({"relu_w": torch.randn(v, 3), "relu_b": torch.zeros(3)})
    layers.append(l)
    din += 3
  p, sp = O.deep_combine_chain_model(torch.randn(b, d), layers, mk(din, v), v, m)
  assert p.shape == (b, v) and sp.shape == (b, 2 * v)
  assert float(p.min()) >= 0 and float(p.max()) <= 1


def test_lstm_oracle_against_torch_nn_lstm():
  """Independent cross-check of the BasicLSTMCell / dynamic_rnn restatement (TF-1.0 semantics): torch.nn.LSTM
  computes the same recurrence with gate order (i, f, g, o) and separate input / hidden matrices.  Mapping: TF kernel
  [x; h] x [i | j | f | o], forget_bias added to f  ->  torch weight_ih / weight_hh rows [i; f; g = j; o], bias_f += 1.
  Sequences are cut at num_frames: outputs past the end are zero and the state is the one at the last real frame."""
  import torch
  from oracle import yt8m_oracle as O
  g = torch.Generator().manual_seed(11)
  b, t, d, h = 3, 7, 5, 4
  x = torch.randn(b, t, d, generator=g, dtype=torch.float64)
  nf = torch.tensor([7, 3, 1])
  w = torch.randn(d + h, 4 * h, generator=g, dtype=torch.float64) * 0.5
  bias = torch.randn(4 * h, generator=g, dtype=torch.float64) * 0.1
  outs, states = O.dynamic_rnn_lstm(x, nf, [(w, bias)], forget_bias=1.0)
  lstm = torch.nn.LSTM(d, h, batch_first=True).double()
  wi, wj, wf, wo = w.chunk(4, dim=1)
  bi, bj, bf, bo = bias.chunk(4)
  with torch.no_grad():
    tw = torch.cat([wi, wf, wj, wo], dim=1).t()                # torch rows: i, f, g, o
    lstm.weight_ih_l0.copy_(tw[:, :d])
    lstm.weight_hh_l0.copy_(tw[:, d:])
    lstm.bias_ih_l0.copy_(torch.cat([bi, bf + 1.0, bj, bo]))
    lstm.bias_hh_l0.zero_()
  for i in range(b):
    n = int(nf[i])
    y, (hn, cn) = lstm(x[i:i + 1, :n])
    assert float((outs[i, :n] - y[0]).abs().max()) < 1e-12
    assert float(outs[i, n:].abs().max()) == 0.0 if n < t else True
    c_fin, h_fin = states[0]
    assert float((c_fin[i] - cn[0, 0]).abs().max()) < 1e-12 and float((h_fin[i] - hn[0, 0]).abs().max()) < 1e-12


def test_netvlad_and_attention_oracles_against_explicit_loops():
  """The vectorised oracle restatements against plain Python loops over the published definitions (small cases):
  NetVLAD (soft assignment over K, residual against cluster_w2, intra-norm per cluster, flatten D-major / K-minor, L2)
  and the softmax-over-T attention pooling with mask and renormalisation (lstm_attention_max_pooling_model.py:58-63)."""
  import numpy as np
  g = torch.Generator().manual_seed(13)
  b, t, d, k = 2, 6, 5, 3
  x = torch.randn(b, t, d, generator=g, dtype=torch.float64)
  nf = torch.tensor([6, 4])
  cw = torch.randn(d, k, generator=g, dtype=torch.float64)
  sc = 1 + 0.1 * torch.randn(k, generator=g, dtype=torch.float64)
  sh = 0.1 * torch.randn(k, generator=g, dtype=torch.float64)
  c2 = torch.randn(d, k, generator=g, dtype=torch.float64)
  got = O.netvlad_pool(x, nf, cw, sc, sh, c2).numpy()
  xs, cwn, scn, shn, c2n = x.numpy(), cw.numpy(), sc.numpy(), sh.numpy(), c2.numpy()
  for i in range(b):
    v = np.zeros((d, k))
    asum = np.zeros(k)
    for tt in range(int(nf[i])):
      logit = np.array([sum(xs[i, tt, dd] * cwn[dd, kk] for dd in range(d)) * scn[kk] + shn[kk] for kk in range(k)])
      e = np.exp(logit - logit.max())
      a = e / e.sum()
      asum += a
      for dd in range(d):
        for kk in range(k):
          v[dd, kk] += a[kk] * xs[i, tt, dd]
    v -= asum[None, :] * c2n
    for kk in range(k):
      v[:, kk] /= max(np.sqrt((v[:, kk] ** 2).sum()), 1e-6)
    flat = v.reshape(-1)                                   # index d * K + k
    flat = flat / np.sqrt((flat ** 2).sum())
    assert np.abs(got[i] - flat).max() < 1e-12
  # attention: weights = softmax over ALL T, times mask, renormalised; pooled[a] = sum_t w[t, a] * outputs[t]
  a_heads, hdim = 2, 4
  outs = torch.randn(b, t, hdim, generator=g, dtype=torch.float64)
  wa = torch.randn(d + hdim, a_heads, generator=g, dtype=torch.float64)
  ba = torch.randn(a_heads, generator=g, dtype=torch.float64)
  pooled = O.attention_softmax_pool(x, outs, nf, wa, ba).numpy()
  for i in range(b):
    logits = np.concatenate([xs[i], outs[i].numpy()], axis=1) @ wa.numpy() + ba.numpy()          # [T, A]
    for a in range(a_heads):
      e = np.exp(logits[:, a] - logits[:, a].max())
      w = e / e.sum()
      w = np.array([w[tt] if tt < int(nf[i]) else 0.0 for tt in range(t)])
      w = w / w.sum()
      want = sum(w[tt] * outs[i, tt].numpy() for tt in range(t))
      assert np.abs(pooled[i, a] - want).max() < 1e-12
