"""GPU parity of the model PLUGINS (the reference's create_model() surface) against whole-model CPU
forwards of the oracle, on identical seeded inputs and identical (bf16-representable) weights.

Bars (north_star): predictions within 1e-3 relative, GAP@20 within 1e-4.
"relative" here: |p_cuda - p_oracle| <= 1e-3 * max(p_oracle, 1e-3) elementwise.
"""
import numpy as np
import pytest
import torch

import synth
from oracle import gap_oracle, model_oracle

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
V = 4716


@pytest.fixture(scope="module")
def env():
  if not torch.cuda.is_available():
    pytest.skip("no CUDA device")
  import frame_level_models
  import video_level_models
  import yt8m_flags
  import yt8m_ops
  return frame_level_models, video_level_models, yt8m_flags.FLAGS, yt8m_ops


def build_and_run(ops, model, gains, seed=9, **kw):
  """First call creates the variables (reference initialisers); they are then rescaled so the logits
  are O(1) -- the default initialisers leave every probability within 1e-3 of M/(M+1)/2, a degenerate
  parity test -- and the forward is run again with the rescaled weights."""
  st = ops.get_store()
  st.reset(seed=seed)
  model.create_model(**kw)
  for name, v in st.vars.items():
    for pat, g in gains.items():
      if pat in name and (name.endswith("weights") or name.endswith("/W")):
        v.assign(synth.bf16r(v.value.cpu() * g))
  out = model.create_model(**kw)
  torch.cuda.synchronize()
  return out, {k: t.clone() for k, t in st.state_dict().items()}


def check(pred, want, labels=None):
  got = pred.detach().float().cpu()
  rel = ((got - want).abs() / want.clamp_min(1e-3)).max()
  assert float(rel) < 1e-3, "max relative error %.3e" % float(rel)
  if labels is not None:
    g1 = gap_oracle.gap(got.numpy(), labels.numpy(), 20)
    g2 = gap_oracle.gap(want.numpy(), labels.numpy(), 20)
    assert abs(g1 - g2) < 1e-4, (g1, g2)


def video_input(b, d, seed):
  g = torch.Generator().manual_seed(seed)
  x = torch.randn(b, d, generator=g)
  x = synth.bf16r(x * torch.rsqrt((x * x).sum(dim=1, keepdim=True)))
  return x


def test_logistic_model(env):
  flm, vlm, FLAGS, ops = env
  x = video_input(128, 1152, 1)
  y = synth.labels(128, V)
  out, sd = build_and_run(ops, vlm.LogisticModel(), {"fully_connected": 30.0}, model_input=x.to(DEV).to(torch.bfloat16),
                          vocab_size=V)
  check(out["predictions"], model_oracle.logistic(sd, x), y)


@pytest.mark.parametrize("m", [2, 4])
def test_moe_model(env, m):
  flm, vlm, FLAGS, ops = env
  x = video_input(96, 1024, 2)
  y = synth.labels(96, V)
  with FLAGS.override(moe_num_mixtures=m):
    out, sd = build_and_run(ops, vlm.MoeModel(), {"gates": 40.0, "experts": 40.0}, model_input=x.to(DEV), vocab_size=V)
  check(out["predictions"], model_oracle.moe(sd, x, V, m), y)


def test_chain_moe_model(env):
  flm, vlm, FLAGS, ops = env
  x = video_input(64, 1152, 3)
  with FLAGS.override(moe_num_mixtures=2, num_supports=25):
    out, sd = build_and_run(ops, vlm.ChainMoeModel(), {"gates": 30.0, "experts": 30.0}, model_input=x.to(DEV), vocab_size=V)
  want, want_sup = model_oracle.chain_moe(sd, x, V, 2, 25)
  check(out["predictions"], want, synth.labels(64, V))
  check(out["support_predictions"], want_sup)


def test_deep_combine_chain_model(env):
  flm, vlm, FLAGS, ops = env
  x = video_input(48, 1152, 4)
  with FLAGS.override(moe_num_mixtures=2, deep_chain_layers=3, deep_chain_relu_cells=200):
    out, sd = build_and_run(ops, vlm.DeepCombineChainModel(), {"gates": 30.0, "experts": 30.0, "relu-": 3.0},
                            model_input=x.to(DEV), vocab_size=V)
  want, want_sup = model_oracle.deep_combine_chain(sd, x, V, 2, 3)
  check(out["predictions"], want, synth.labels(48, V))
  check(out["support_predictions"], want_sup)


@pytest.mark.parametrize("fmt", ["f16", "bf16x2"])
@pytest.mark.parametrize("k,gating", [(64, False), (128, True)])
def test_netvlad_model(env, k, gating, fmt):
  flm, vlm, FLAGS, ops = env
  b = 6
  x, nf, _ = synth.model_input(b, seed=8)
  y = synth.labels(b, V)
  model = flm.GatedNetVLADModel() if gating else flm.NetVLADModel()
  with FLAGS.override(netvlad_cluster_size=k, netvlad_hidden_size=1024, moe_num_mixtures=2 if not gating else 4,
                      netvlad_operand_format=fmt):
    out, sd = build_and_run(ops, model, {"cluster_weights": 30.0, "gates": 8.0, "experts": 8.0},
                            model_input=x.to(DEV).to(torch.bfloat16), vocab_size=V, num_frames=nf.to(DEV))
  want = model_oracle.netvlad(sd, x, nf, V, 2 if not gating else 4, gating=gating)
  check(out["predictions"], want, y)


def test_lstm_model(env):
  flm, vlm, FLAGS, ops = env
  b = 6
  x, nf, _ = synth.model_input(b, seed=10)
  with FLAGS.override(lstm_cells="1024", lstm_layers=2, moe_num_mixtures=4):
    out, sd = build_and_run(ops, flm.LstmModel(), {"basic_lstm_cell": 1.0, "gates": 10.0, "experts": 10.0},
                            model_input=x.to(DEV).to(torch.bfloat16), vocab_size=V, num_frames=nf.to(DEV))
  check(out["predictions"], model_oracle.lstm_model(sd, x, nf, V, 4), synth.labels(b, V))


def test_lstm_memory_model(env):
  flm, vlm, FLAGS, ops = env
  b = 4
  x, nf, _ = synth.model_input(b, frames=60, seed=11)
  with FLAGS.override(lstm_cells="256", lstm_layers=2, moe_num_mixtures=2):
    out, sd = build_and_run(ops, flm.LstmMemoryModel(), {"basic_lstm_cell": 1.5, "gates": 10.0, "experts": 10.0},
                            model_input=x.to(DEV).to(torch.bfloat16), vocab_size=V, num_frames=nf.to(DEV))
  check(out["predictions"], model_oracle.lstm_memory_model(sd, x, nf, V, 2))


def test_lstm_attention_max_pooling_model(env):
  flm, vlm, FLAGS, ops = env
  b = 4
  x, nf, _ = synth.model_input(b, frames=80, seed=12)
  with FLAGS.override(lstm_cells="256", lstm_layers=2, moe_num_mixtures=2, lstm_attentions=8):
    out, sd = build_and_run(ops, flm.LstmAttentionMaxPoolingModel(),
                            {"basic_lstm_cell": 1.5, "attention-": 20.0, "gates": 10.0, "experts": 10.0},
                            model_input=x.to(DEV).to(torch.bfloat16), vocab_size=V, num_frames=nf.to(DEV))
  check(out["predictions"], model_oracle.lstm_attention_max_pooling(sd, x, nf, V, 2, 8))


def test_lstm_multi_attention_model(env):
  flm, vlm, FLAGS, ops = env
  b = 4
  x, nf, _ = synth.model_input(b, frames=80, seed=13)
  with FLAGS.override(lstm_cells="256", lstm_layers=2, moe_num_mixtures=2, attention_size=4):
    out, sd = build_and_run(ops, flm.LstmMultiAttentionModel(),
                            {"basic_lstm_cell": 1.5, "fully_connected": 10.0, "gates": 10.0, "experts": 10.0},
                            model_input=x.to(DEV).to(torch.bfloat16), vocab_size=V, num_frames=nf.to(DEV))
  check(out["predictions"], model_oracle.lstm_multi_attention(sd, x, nf, V, 2, 4))


def test_attention_model_with_moe_extend(env):
  flm, vlm, FLAGS, ops = env
  b = 8
  x, nf, _ = synth.model_input(b, seed=14)
  y = synth.labels(b, V)
  with FLAGS.override(moe_num_mixtures=4, moe_num_extend=8, video_level_classifier_model="MoeExtendModel"):
    out, sd = build_and_run(ops, flm.AttentionModel(), {"Attention/W": 30.0, "gates": 10.0, "experts": 10.0},
                            model_input=x.to(DEV).to(torch.bfloat16), vocab_size=V, num_frames=nf.to(DEV))
  check(out["predictions"], model_oracle.attention_model(sd, x, nf, V, 4, 8), y)


def test_dbof_model(env):
  flm, vlm, FLAGS, ops = env
  b = 6
  x, nf, _ = synth.model_input(b, seed=15)
  g = torch.Generator().manual_seed(16)
  fi = (torch.rand(b, 30, generator=g) * nf.unsqueeze(1)).to(torch.int64)
  with FLAGS.override(dbof_cluster_size=2048, dbof_hidden_size=1024, moe_num_mixtures=2, iterations=30):
    out, sd = build_and_run(ops, flm.DbofModel(), {"gates": 3.0, "experts": 3.0},
                            model_input=x.to(DEV).to(torch.bfloat16), vocab_size=V, num_frames=nf.to(DEV), frame_index=fi)
  check(out["predictions"], model_oracle.dbof(sd, x, fi, V, 2))


def test_plugin_surface(env):
  """find_class_by_name-style lookup and the error behaviour of the reference boundary."""
  flm, vlm, FLAGS, ops = env
  import models
  with pytest.raises(NotImplementedError):
    models.BaseModel().create_model(None)
  for name in ("LogisticModel", "MoeModel", "ChainMoeModel", "DeepCombineChainModel"):
    assert issubclass(getattr(vlm, name), models.BaseModel)
  for name in ("LstmModel", "LstmMemoryModel", "LstmAttentionMaxPoolingModel", "LstmMultiAttentionModel", "DbofModel",
               "NetVLADModel", "GatedNetVLADModel", "AttentionModel"):
    assert issubclass(getattr(flm, name), models.BaseModel)
  with pytest.raises(ValueError):
    with FLAGS.override(dbof_pooling_method="attention"):
      x, nf, _ = synth.model_input(2, frames=40, seed=1)
      flm.DbofModel().create_model(x.to(DEV).to(torch.bfloat16), vocab_size=V, num_frames=nf.to(DEV))


def test_captured_step_replay_matches_eager(env):
  """ops.CapturedStep: the CUDA-graph replay of create_model() gives the eager result (to the summation-order
  noise of the split-K fp32 reductions in the hidden FC), sees inputs refilled in place, and carries a live event
  pair around the NetVLAD kernel."""
  flm, vlm, FLAGS, ops = env
  b = 5
  x, nf, _ = synth.model_input(b, seed=21)
  x2, nf2, _ = synth.model_input(b, seed=22)
  xd, nfd = x.to(DEV).to(torch.bfloat16), nf.to(DEV)
  model = flm.NetVLADModel()
  with FLAGS.override(netvlad_cluster_size=64, netvlad_hidden_size=1024, moe_num_mixtures=2):
    ops.get_store().reset(seed=9)
    fn = lambda: model.create_model(model_input=xd, vocab_size=V, num_frames=nfd)["predictions"]
    eager1 = fn().clone()
    step = ops.CapturedStep(fn, time_tag="netvlad")
    got1 = step().clone()
    xd.copy_(x2.to(DEV).to(torch.bfloat16))
    nfd.copy_(nf2.to(DEV))
    got2 = step().clone()
    eager2 = fn().clone()
    torch.cuda.synchronize()
  assert torch.allclose(got1, eager1, rtol=2e-5, atol=1e-7)
  assert torch.allclose(got2, eager2, rtol=2e-5, atol=1e-7)
  assert float((got1 - got2).abs().max()) > 1e-5          # the refilled inputs were seen
  ms = step.kernel_ms()
  assert len(ms) == 1 and 0.0 < ms[0] < 50.0


def test_dbof_model_training_mode_batch_norm(env):
  """DbofModel.create_model(is_training=True) with the reference's default flags: slim.batch_norm with BATCH statistics in all
  three layers (wh/all_frame_models/dbof_model.py:64-108) against the oracle, and the moving averages it leaves behind."""
  flm, vlm, FLAGS, ops = env
  b = 8
  x, nf, _ = synth.model_input(b, seed=17)
  g = torch.Generator().manual_seed(18)
  fi = (torch.rand(b, 30, generator=g) * nf.unsqueeze(1)).to(torch.int64)
  kw = dict(model_input=x.to(DEV).to(torch.bfloat16), vocab_size=V, num_frames=nf.to(DEV), frame_index=fi)
  with FLAGS.override(dbof_cluster_size=2048, dbof_hidden_size=1024, moe_num_mixtures=2, iterations=30):
    st = ops.get_store()
    st.reset(seed=9)
    model = flm.DbofModel()
    model.create_model(**kw)                                   # creates the variables (inference form)
    for name, v in st.vars.items():
      if name.endswith("/weights"):
        v.assign(synth.bf16r(v.value.cpu() * 3.0))
    sd0 = {k: t.clone() for k, t in st.state_dict().items()}
    out = model.create_model(is_training=True, **kw)
    torch.cuda.synchronize()
    sd1 = st.state_dict()

  def bn(scope):
    return {"gamma": sd0[scope + "/gamma"], "beta": sd0[scope + "/beta"], "mean": sd0[scope + "/moving_mean"], "var": sd0[scope + "/moving_variance"]}
  p = {"cluster_w": sd0["cluster_weights"], "hidden_w": sd0["hidden1_weights"], "input_bn": bn("input_bn"), "cluster_bn": bn("cluster_bn"),
       "hidden1_bn": bn("hidden1_bn")}
  from oracle import yt8m_oracle as O
  h = O.dbof_pool(x, fi, p, is_training=True, add_batch_norm=True, pooling="max")
  want = O.moe_model(h, sd0["gates/weights"], sd0["experts/weights"], sd0["experts/biases"], V, 2)
  check(out["predictions"], want)
  rows = x[torch.arange(b).unsqueeze(1), fi].reshape(b * 30, -1)
  _, mm, mv = O.batch_norm(rows, sd0["input_bn/gamma"], sd0["input_bn/beta"], sd0["input_bn/moving_mean"], sd0["input_bn/moving_variance"], True)
  assert float((sd1["input_bn/moving_mean"] - mm).abs().max()) < 1e-6 and float((sd1["input_bn/moving_variance"] - mv).abs().max()) < 1e-6
  assert float((sd1["hidden1_bn/moving_mean"] - sd0["hidden1_bn/moving_mean"]).abs().max()) > 0


def test_cnn_deep_combine_chain_model(env):
  """CnnDeepCombineChainModel (wh/all_frame_models/cnn_deep_combine_chain_model.py): the temporal CNN as three
  tensor-core GEMMs over the shifted-concat operand + max over time + the chain of MoE stages, against the oracle (itself held
  to the reference source by tests/test_oracle_golden_models.py)."""
  flm, vlm, FLAGS, ops = env
  b = 12
  x, nf, _ = synth.model_input(b, frames=120, seed=31)
  nf[0], nf[1] = 120, 1
  x = x * (torch.arange(120).unsqueeze(0) < nf.unsqueeze(1)).float().unsqueeze(2)
  y = synth.labels(b, V)
  with FLAGS.override(moe_num_mixtures=2, deep_chain_layers=2, deep_chain_relu_cells=64):
    out, sd = build_and_run(ops, flm.CnnDeepCombineChainModel(), {"gates": 30.0, "experts": 30.0, "relu-": 3.0, "cnn": 1.0},
                            model_input=x.to(DEV).to(torch.bfloat16), vocab_size=V, num_frames=nf.to(DEV))
  want, want_sup = model_oracle.cnn_deep_combine_chain(sd, x, nf, V, 2, 2)
  check(out["predictions"], want, y)
  check(out["support_predictions"], want_sup)


def test_lstm_parallel_finaloutput_model(env):
  """LstmParallelFinaloutputModel (wh/all_frame_models/lstm_parallel_finaloutput_model.py) with the reference's
  modalities: rgb 1024 -> LSTM-1024 x 2, audio 128 -> LSTM-128 x 2, h states concatenated (2304-d) -> MoE."""
  flm, vlm, FLAGS, ops = env
  import yt8m_flags as flags
  for name, default in (("feature_names", "mean_rgb"), ("feature_sizes", "1024")):
    if name not in FLAGS:
      flags.DEFINE_string(name, default, "defined by the command-line front ends (train.py / eval.py / inference.py)")
  b = 5
  x, nf, _ = synth.model_input(b, frames=40, seed=32, min_frames=5)
  with FLAGS.override(lstm_cells="1024,128", lstm_layers=2, moe_num_mixtures=2, feature_names="rgb,audio", feature_sizes="1024,128",
                      video_level_classifier_model="MoeModel"):
    out, sd = build_and_run(ops, flm.LstmParallelFinaloutputModel(), {"basic_lstm_cell": 1.0, "gates": 10.0, "experts": 10.0},
                            model_input=x.to(DEV).to(torch.bfloat16), vocab_size=V, num_frames=nf.to(DEV))
  want = model_oracle.lstm_parallel_finaloutput(sd, x, nf, V, 2, [1024, 128], 2)
  check(out["predictions"], want)
