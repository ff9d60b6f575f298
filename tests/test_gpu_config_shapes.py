"""GPU parity of the model plugins AT THE BASELINE.json SHAPES (configs 2-5), against the whole-model CPU oracle.

tests/test_gpu_models.py holds the same plugins to the oracle at small batches; the kernels' schedulers, tile
pairings and split-K choices depend on the batch, so the benchmark shapes get their own cases here:

  config 2  NetVLAD K=64 + FC 73,728->1024 + MoE-2, B=256, T=300, D=1152
  config 3  LstmModel L=2 H=1024 + MoE-4 on the 4096-d state, B=64 per GPU
  config 4  GatedNetVLAD K=128 + FC 147,456->1024 + context gate + MoE-4, B=512
  config 5  AttentionModel A=8 + DeepCombineChainModel (3 layers), B in {64, 1024}; and its LSTM sibling
            LstmAttentionMaxPoolingModel at H=1024, T=300

Videos are independent on this path, so where the fp32 oracle would take minutes the GPU runs the FULL batch and the
oracle is evaluated on a fixed subset of its videos (stated per test).  Bars as in test_gpu_models.py:
|p_cuda - p_oracle| <= 1e-3 * max(p_oracle, 1e-3), GAP@20 within 1e-4.
"""
import pytest
import torch

import synth
from oracle import gap_oracle, model_oracle
from oracle import yt8m_oracle as O

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
  st = ops.get_store()
  st.reset(seed=seed)
  model.create_model(**kw)
  for name, v in st.vars.items():
    for pat, g in gains.items():
      if pat in name and (name.endswith("weights") or name.endswith("/W") or name.endswith("weights2")):
        v.assign(synth.bf16r(v.value.cpu() * g))
  out = model.create_model(**kw)
  torch.cuda.synchronize()
  sd = {k: t.clone() for k, t in st.state_dict().items()}
  st.reset(seed=seed)                       # release the packed copies of this (large) model
  torch.cuda.empty_cache()
  return out, sd


def check_rows(pred, want, rows, labels=None):
  got = pred.detach().float().cpu()[rows]
  rel = ((got - want).abs() / want.clamp_min(1e-3)).max()
  assert float(rel) < 1e-3, "max relative error %.3e" % float(rel)
  assert torch.isfinite(pred).all()
  if labels is not None:
    g1 = gap_oracle.gap(got.numpy(), labels[rows].numpy(), 20)
    g2 = gap_oracle.gap(want.numpy(), labels[rows].numpy(), 20)
    assert abs(g1 - g2) < 1e-4, (g1, g2)


def subset(b, n):
  """n video indices spread over the batch (first, last and evenly between), sorted."""
  if n >= b:
    return torch.arange(b)
  return torch.unique(torch.linspace(0, b - 1, n).round().to(torch.int64))


def test_config2_netvlad64_b256(env):
  """BASELINE configs[1]: the benchmark batch itself -- the one-pass kernel's longest-first schedule at B=256."""
  flm, vlm, FLAGS, ops = env
  b = 256
  x, nf, _ = synth.model_input(b, seed=8)
  y = synth.labels(b, V)
  with FLAGS.override(netvlad_cluster_size=64, netvlad_hidden_size=1024, moe_num_mixtures=2, netvlad_operand_format="f16",
                      video_level_classifier_model="MoeModel"):
    out, sd = build_and_run(ops, flm.NetVLADModel(), {"cluster_weights": 30.0, "gates": 8.0, "experts": 8.0},
                            model_input=x.to(DEV).to(torch.bfloat16), vocab_size=V, num_frames=nf.to(DEV))
  rows = subset(b, 96)
  want = model_oracle.netvlad(sd, x[rows], nf[rows], V, 2)
  check_rows(out["predictions"], want, rows, y)


def test_config4_gated_netvlad128_b512(env):
  """BASELINE configs[3]: K=128, context gating, MoE-4, global batch 512 on one GPU; oracle on 48 of the 512 videos."""
  flm, vlm, FLAGS, ops = env
  b = 512
  x, nf, _ = synth.model_input(b, seed=18)
  y = synth.labels(b, V)
  with FLAGS.override(netvlad_cluster_size=128, netvlad_hidden_size=1024, moe_num_mixtures=4, netvlad_operand_format="f16",
                      video_level_classifier_model="MoeModel"):
    out, sd = build_and_run(ops, flm.GatedNetVLADModel(), {"cluster_weights": 30.0, "gates": 8.0, "experts": 8.0},
                            model_input=x.to(DEV).to(torch.bfloat16), vocab_size=V, num_frames=nf.to(DEV))
  rows = subset(b, 48)
  want = model_oracle.netvlad(sd, x[rows], nf[rows], V, 4, gating=True)
  check_rows(out["predictions"], want, rows, y)


def test_config3_lstm1024_b64(env):
  """BASELINE configs[2]: 2 x LSTM-1024 over 300 frames + MoE-4 on [c0,h0,c1,h1]; 64 videos per GPU; oracle on 10."""
  flm, vlm, FLAGS, ops = env
  b = 64
  x, nf, _ = synth.model_input(b, seed=19)
  with FLAGS.override(lstm_cells="1024", lstm_layers=2, moe_num_mixtures=4, video_level_classifier_model="MoeModel"):
    out, sd = build_and_run(ops, flm.LstmModel(), {"basic_lstm_cell": 1.0, "gates": 10.0, "experts": 10.0},
                            model_input=x.to(DEV).to(torch.bfloat16), vocab_size=V, num_frames=nf.to(DEV))
  rows = subset(b, 10)
  want = model_oracle.lstm_model(sd, x[rows], nf[rows], V, 4)
  check_rows(out["predictions"], want, rows)


def test_config5_sibling_lstm_attention_h1024_t300(env):
  """LstmAttentionMaxPoolingModel at the reference's own size (H=1024, T=300, 8 heads, MoE-2); oracle on 6 of 16."""
  flm, vlm, FLAGS, ops = env
  b = 16
  x, nf, _ = synth.model_input(b, seed=20)
  with FLAGS.override(lstm_cells="1024", lstm_layers=2, moe_num_mixtures=2, lstm_attentions=8):
    out, sd = build_and_run(ops, flm.LstmAttentionMaxPoolingModel(),
                            {"basic_lstm_cell": 1.0, "attention-": 20.0, "gates": 10.0, "experts": 10.0},
                            model_input=x.to(DEV).to(torch.bfloat16), vocab_size=V, num_frames=nf.to(DEV))
  rows = subset(b, 6)
  want = model_oracle.lstm_attention_max_pooling(sd, x[rows], nf[rows], V, 2, 8)
  check_rows(out["predictions"], want, rows)


@pytest.mark.parametrize("b", [64, 1024])
def test_config5_attention_chain(env, b):
  """BASELINE configs[4]: 8-head attention pooling over 300 x 1152 + chained MoE (DeepCombineChainModel, 3 layers, MoE-4 per
  layer) applied per head, max over the heads; the ends of the batch sweep.  Oracle on 12 videos."""
  flm, vlm, FLAGS, ops = env
  x, nf, _ = synth.model_input(b, seed=21)
  y = synth.labels(b, V)
  with FLAGS.override(moe_num_mixtures=4, moe_num_extend=8, video_level_classifier_model="DeepCombineChainModel",
                      deep_chain_layers=3, deep_chain_relu_cells=256):
    out, sd = build_and_run(ops, flm.AttentionModel(), {"Attention/W": 30.0, "gates": 10.0, "experts": 10.0, "relu-": 3.0},
                            model_input=x.to(DEV).to(torch.bfloat16), vocab_size=V, num_frames=nf.to(DEV))
  assert tuple(out["predictions"].shape) == (b, V)
  rows = subset(b, 12)
  want = model_oracle.attention_chain(sd, x[rows], nf[rows], V, 4, 8, 3)
  check_rows(out["predictions"], want, rows, y)
