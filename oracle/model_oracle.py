"""Whole-model CPU forwards built from oracle/yt8m_oracle.py, driven by a variable dictionary that uses
the reference's variable names (the same names the product's variable store uses).  TEST
INFRASTRUCTURE ONLY -- see oracle/yt8m_oracle.py for who may import ``oracle/``.

Every function takes ``sd`` (name -> fp32 CPU tensor, TF layouts) and returns predictions [B, V].
"""
import torch

from oracle import yt8m_oracle as O

BN_EPS = 1e-3


def _moe(sd, x, v, m, gates="gates", experts="experts"):
  return O.moe_model(x, sd[gates + "/weights"], sd[experts + "/weights"], sd[experts + "/biases"], v, m)


def _bn_affine(sd, scope):
  s = sd[scope + "/gamma"] * torch.rsqrt(sd[scope + "/moving_variance"] + BN_EPS)
  return s, sd[scope + "/beta"] - sd[scope + "/moving_mean"] * s


def logistic(sd, x):
  """wh/all_video_models/logistic_model.py:23-26."""
  return O.logistic_model(x, sd["fully_connected/weights"], sd["fully_connected/biases"])


def moe(sd, x, vocab, mixtures):
  """wh/all_video_models/moe_model.py:38-65."""
  return _moe(sd, x, vocab, mixtures)


def chain_moe(sd, x, vocab, mixtures, num_supports):
  """wh/all_video_models/chain_moe_model.py:12-18."""
  sup = {"gate_w": sd["gates-support/weights"], "expert_w": sd["experts-support/weights"], "expert_b": sd["experts-support/biases"]}
  main = {"gate_w": sd["gates-main/weights"], "expert_w": sd["experts-main/weights"], "expert_b": sd["experts-main/biases"]}
  return O.chain_moe_model(x, sup, main, vocab, num_supports, mixtures)


def deep_combine_chain(sd, x, vocab, mixtures, layers):
  """wh/all_video_models/deep_combine_chain_model.py:24-49."""
  lyr = []
  for i in range(layers):
    lyr.append({"gate_w": sd["gates-prediction-%d/weights" % i], "expert_w": sd["experts-prediction-%d/weights" % i],
                "expert_b": sd["experts-prediction-%d/biases" % i], "relu_w": sd["relu-%d/weights" % i],
                "relu_b": sd["relu-%d/biases" % i]})
  main = {"gate_w": sd["gates--main/weights"], "expert_w": sd["experts--main/weights"], "expert_b": sd["experts--main/biases"]}
  return O.deep_combine_chain_model(x, lyr, main, vocab, mixtures)


def _lstm(sd, x, nf, layers):
  ws = [(sd["RNN/multi_rnn_cell/cell_%d/basic_lstm_cell/weights" % l], sd["RNN/multi_rnn_cell/cell_%d/basic_lstm_cell/biases" % l])
        for l in range(layers)]
  return O.dynamic_rnn_lstm(x, nf, ws)


def lstm_model(sd, x, nf, vocab, mixtures, layers=2):
  """wh/all_frame_models/lstm_model.py:30-55 + MoeModel."""
  _, states = _lstm(sd, x, nf, layers)
  return _moe(sd, O.lstm_model_state(states), vocab, mixtures)


def lstm_memory_model(sd, x, nf, vocab, mixtures, layers=2):
  """wh/all_frame_models/lstm_memory_model.py:47-73 + MoeModel."""
  _, states = _lstm(sd, x, nf, layers)
  return _moe(sd, O.lstm_memory_model_state(states), vocab, mixtures)


def lstm_attention_max_pooling(sd, x, nf, vocab, mixtures, heads, layers=2):
  """wh/all_frame_models/lstm_attention_max_pooling_model.py:29-68."""
  outs, _ = _lstm(sd, x, nf, layers)
  pooled = O.attention_softmax_pool(x, outs, nf, sd["attention-/weights"], sd["attention-/biases"])   # [B, A, H]
  b, a, h = pooled.shape
  p = _moe(sd, pooled.reshape(b * a, h), vocab, mixtures, "gates-sub-moe", "experts-sub-moe")
  return p.reshape(b, a, vocab).max(dim=1).values


def lstm_multi_attention(sd, x, nf, vocab, mixtures, heads, layers=2):
  """wh/all_frame_models/lstm_multi_attention_model.py:30-91 with MoeModel as the classifier."""
  outs, _ = _lstm(sd, x, nf, layers)
  pooled = O.attention_sigmoid_pool(x, outs, nf, sd["fully_connected/weights"], sd["fully_connected/biases"])
  b, a, d = pooled.shape
  p = _moe(sd, pooled.reshape(b * a, d), vocab, mixtures)
  return p.reshape(b, a, vocab).max(dim=1).values


def attention_model(sd, x, nf, vocab, mixtures, heads):
  """zt/frame_level_models.py:4355-4405 + MoeExtendModel (zt/video_level_models.py:2299-2330)."""
  state = O.attention_model_pool(x, nf, sd["Attention/W"], sd["Attention/b"])
  return O.moe_extend_model(state, sd["gates/weights"], sd["experts/weights"], sd["experts/biases"], vocab, mixtures, heads)


def attention_chain(sd, x, nf, vocab, mixtures, heads, layers):
  """zt/frame_level_models.py:4355-4405 pooling with --video_level_classifier_model=DeepCombineChainModel
  (wh/all_video_models/deep_combine_chain_model.py:24-49) applied to each of the B*A pooled rows, then the max over the A
  heads that MoeExtendModel takes (zt/video_level_models.py:2327-2328) -- BASELINE.json configs[4]."""
  state = O.attention_model_pool(x, nf, sd["Attention/W"], sd["Attention/b"])
  p, _ = deep_combine_chain(sd, state, vocab, mixtures, layers)
  return p.reshape(-1, heads, vocab).max(dim=1).values


def cnn_deep_combine_chain(sd, x, nf, vocab, mixtures, layers, filter_sizes=(1, 2, 3)):
  """wh/all_frame_models/cnn_deep_combine_chain_model.py:43-88 (variables by the reference's names)."""
  def moe_p(scope):
    return {"gate_w": sd["gates-%s/weights" % scope], "expert_w": sd["experts-%s/weights" % scope], "expert_b": sd["experts-%s/biases" % scope]}
  p = {"mean_relu_w": sd["mean-relu/weights"], "mean_relu_b": sd["mean-relu/biases"],
       "cnn": [[(fs, sd["cnn%dcnn-filter-len%d" % (l, fs)]) for fs in filter_sizes] for l in range(layers + 1)],
       "layers": [dict(moe_p("prediction-%d" % l), relu_w=sd["relu-%d/weights" % l], relu_b=sd["relu-%d/biases" % l]) for l in range(layers)],
       "main": moe_p("-main")}
  return O.cnn_deep_combine_chain_model(x, nf, p, vocab, mixtures, layers)


def lstm_parallel_finaloutput(sd, x, nf, vocab, mixtures, feature_sizes, layers=2):
  """wh/all_frame_models/lstm_parallel_finaloutput_model.py:32-73 + MoeModel."""
  stacks = [[(sd["RNN%d/multi_rnn_cell/cell_%d/basic_lstm_cell/weights" % (i, l)], sd["RNN%d/multi_rnn_cell/cell_%d/basic_lstm_cell/biases" % (i, l)])
             for l in range(layers)] for i in range(len(feature_sizes))]
  return _moe(sd, O.lstm_parallel_finaloutput_state(x, nf, feature_sizes, stacks), vocab, mixtures)


def dbof(sd, x, frame_index, vocab, mixtures, pooling="max"):
  """wh/all_frame_models/dbof_model.py:62-123 (inference-mode batch norm) + MoeModel."""
  def bn(scope):
    return {"gamma": sd[scope + "/gamma"], "beta": sd[scope + "/beta"], "mean": sd[scope + "/moving_mean"],
            "var": sd[scope + "/moving_variance"]}
  p = {"cluster_w": sd["cluster_weights"], "hidden_w": sd["hidden1_weights"], "input_bn": bn("input_bn"),
       "cluster_bn": bn("cluster_bn"), "hidden1_bn": bn("hidden1_bn")}
  h = O.dbof_pool(x, frame_index, p, is_training=False, add_batch_norm=True, pooling=pooling)
  return _moe(sd, h, vocab, mixtures)


def netvlad(sd, x, nf, vocab, mixtures, gating=False, relu=True):
  """NetVLAD (+ context gating) + hidden FC + MoeModel; not in /root/reference, see yt8m_oracle.netvlad_pool."""
  s, t = _bn_affine(sd, "cluster_bn")
  v = O.netvlad_pool(x, nf, sd["cluster_weights"], s, t, sd["cluster_weights2"])
  s_h, t_h = _bn_affine(sd, "hidden1_bn")
  h = (v @ sd["hidden1_weights"]) * s_h + t_h
  if relu:
    h = O.relu6(h)
  if gating:
    s_g, t_g = _bn_affine(sd, "gating_bn")
    h = O.context_gating(h, sd["gating_weights"], s_g, t_g)
  return _moe(sd, h, vocab, mixtures)
