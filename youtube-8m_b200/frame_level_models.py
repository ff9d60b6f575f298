"""Frame-level model plugins with the reference's surface (wh/frame_level_models.py:16-89):
``create_model(model_input [B, T, D], vocab_size, num_frames [B], **unused_params)`` ->
``{"predictions": [B, vocab_size]}``; the video-level head is resolved by name with
``getattr(video_level_models, FLAGS.video_level_classifier_model)`` exactly like the reference
(e.g. wh/all_frame_models/lstm_model.py:49-55).

Pooling runs in the hand-written sm_100a kernels (fused NetVLAD, LSTM step GEMM with fused gates,
attention pooling); ``model_input`` is a CUDA tensor (bf16 preferred; fp32 is converted on the device)
that has already been through the feature transformer (L2-normalised rows), as in wh/train.py:355-361.
"""
import math

import torch

import yt8m_flags as flags
import models
import video_level_models
import yt8m_ops as ops
import yt8m_native as nat

FLAGS = flags.FLAGS

flags.DEFINE_integer("iterations", 30, "Number of frames per batch for DBoF.")
flags.DEFINE_bool("dbof_add_batch_norm", True, "Adds batch normalization to the DBoF model.")
flags.DEFINE_bool("sample_random_frames", True,
                  "If true samples random frames (for frame level models). If false, a random"
                  "sequence of frames is sampled instead.")
flags.DEFINE_integer("dbof_cluster_size", 8192, "Number of units in the DBoF cluster layer.")
flags.DEFINE_integer("dbof_hidden_size", 1024, "Number of units in the DBoF hidden layer.")
flags.DEFINE_string("dbof_pooling_method", "max",
                    "The pooling method used in the DBoF cluster layer. Choices are 'average' and 'max'.")
flags.DEFINE_string("video_level_classifier_model", "MoeModel",
                    "Some Frame-Level models can be decomposed into a generalized pooling operation "
                    "followed by a classifier layer")
flags.DEFINE_bool("rnn_swap_memory", False, "If true, swap_memory = True. (accepted, ignored: 180 GB of HBM)")
flags.DEFINE_string("lstm_cells", "1024", "Number of LSTM cells.")
flags.DEFINE_integer("lstm_layers", 2, "Number of LSTM layers.")
flags.DEFINE_integer("attention_size", 1, "Number of attention layers.")
flags.DEFINE_integer("lstm_attentions", 8, "Attention size in lstm_attention_max_pooling_model.")
flags.DEFINE_bool("is_training", False, "used in batch normalization.")
# NetVLAD family (not in the reference; flag names follow the published WILLOW implementation)
flags.DEFINE_integer("netvlad_cluster_size", 64, "Number of NetVLAD clusters.")
flags.DEFINE_integer("netvlad_hidden_size", 1024, "Number of units in the NetVLAD hidden layer.")
flags.DEFINE_bool("netvlad_add_batch_norm", True, "Batch-normalise the assignment logits and the hidden layer.")
flags.DEFINE_bool("netvlad_relu", True, "ReLU6 after the NetVLAD hidden layer (as DBoF does).")
flags.DEFINE_string("netvlad_operand_format", "f16",
                    "How the descriptor and the hidden layer travel to the next GEMM: 'f16' = one IEEE fp16 tensor "
                    "(11 significant bits, one MMA per weight tile), 'bf16x2' = bf16 hi + lo pair (~16 bits, two MMAs).")

# flags of the reference's other 50 frame-level models (wh/frame_level_models.py:20-83, outside SURVEY.md §8): accepted so that
# command lines parse; the models themselves are not built (find_class_by_name raises StopIteration for them)
flags.DEFINE_string("cnn_filter_nums", '256,256,256', "reference flag of a model / loss outside SURVEY.md §8 (accepted, unused)")
flags.DEFINE_string("cnn_filter_sizes", '1,2,3', "reference flag of a model / loss outside SURVEY.md §8 (accepted, unused)")
flags.DEFINE_integer("cnn_num_filters", 512, "reference flag of a model / loss outside SURVEY.md §8 (accepted, unused)")
flags.DEFINE_integer("cnn_pooling_k", 4, "reference flag of a model / loss outside SURVEY.md §8 (accepted, unused)")
flags.DEFINE_integer("deep_cnn_base_size", 128, "reference flag of a model / loss outside SURVEY.md §8 (accepted, unused)")
flags.DEFINE_integer("distillchain_relu_cells", 256, "reference flag of a model / loss outside SURVEY.md §8 (accepted, unused)")
flags.DEFINE_integer("frame_seg_relu_cells", 256, "reference flag of a model / loss outside SURVEY.md §8 (accepted, unused)")
flags.DEFINE_integer("gru_cells", 1024, "reference flag of a model / loss outside SURVEY.md §8 (accepted, unused)")
flags.DEFINE_integer("gru_layers", 2, "reference flag of a model / loss outside SURVEY.md §8 (accepted, unused)")
flags.DEFINE_integer("lstm_look_back", 3, "reference flag of a model / loss outside SURVEY.md §8 (accepted, unused)")
flags.DEFINE_string("lstm_normalization", 'identical', "reference flag of a model / loss outside SURVEY.md §8 (accepted, unused)")
flags.DEFINE_integer("mm_label_embedding", 256, "reference flag of a model / loss outside SURVEY.md §8 (accepted, unused)")
flags.DEFINE_integer("multiscale_cnn_lstm_layers", 1, "reference flag of a model / loss outside SURVEY.md §8 (accepted, unused)")
flags.DEFINE_integer("num_attentions", 5, "reference flag of a model / loss outside SURVEY.md §8 (accepted, unused)")
flags.DEFINE_integer("positional_embedding_size", 32, "reference flag of a model / loss outside SURVEY.md §8 (accepted, unused)")
flags.DEFINE_string("video_level_classifier_support_model", 'MoeModel', "reference flag of a model / loss outside SURVEY.md §8 (accepted, unused)")
flags.DEFINE_string("wide_and_deep_models", 'FrameLevelLogisticModel,LstmMemoryModel', "reference flag of a model / loss outside SURVEY.md §8 (accepted, unused)")

BN_EPS = 1e-3      # slim.batch_norm default epsilon


# --------------------------------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------------------------------

def _classifier(name=None):
  return getattr(video_level_models, name or FLAGS.video_level_classifier_model)()


def _bn_affine(scope, channels, is_training):
  """slim.batch_norm(center=True, scale=True) in inference form: y = x * s + t with
  s = gamma * rsqrt(moving_var + eps), t = beta - moving_mean * s.  Variables keep slim's names.
  Training-mode batch statistics are outside the forward hot path built this round."""
  if is_training:
    raise NotImplementedError("batch-norm with batch statistics (is_training=True) is not built yet; "
                              "the forward path uses the moving statistics")
  st = ops.get_store()
  gamma = st.get(scope + "/gamma", (channels,), ops.ones_init, round_bf16=False)
  beta = st.get(scope + "/beta", (channels,), ops.zeros_init, round_bf16=False)
  mean = st.get(scope + "/moving_mean", (channels,), ops.zeros_init, trainable=False, round_bf16=False)
  var = st.get(scope + "/moving_variance", (channels,), ops.ones_init, trainable=False, round_bf16=False)
  ver = (gamma.version, beta.version, mean.version, var.version)

  def build():
    s = gamma.value * torch.rsqrt(var.value + BN_EPS)         # [C]-sized parameter folding
    return s, beta.value - mean.value * s

  return st.packed(gamma, "bn_affine", build, version=ver)


def _bn_train(scope, x, act, want_bf16=True):
  """slim.batch_norm(is_training=True) (+ activation) on a [rows, channels] activation: batch statistics, and the moving
  averages of the scope's variables updated in place like the reference's UPDATE_OPS (wh/train.py:449-456)."""
  st = ops.get_store()
  c = x.shape[1]
  gamma = st.get(scope + "/gamma", (c,), ops.ones_init, round_bf16=False)
  beta = st.get(scope + "/beta", (c,), ops.zeros_init, round_bf16=False)
  mean = st.get(scope + "/moving_mean", (c,), ops.zeros_init, trainable=False, round_bf16=False)
  var = st.get(scope + "/moving_variance", (c,), ops.ones_init, trainable=False, round_bf16=False)
  out, _ = nat.bn_train_fwd(x, gamma.value, beta.value, mean.value, var.value, act=act, want_bf16=want_bf16)
  mean.version += 1
  var.version += 1
  return out


def _lstm_stack(model_input, num_frames, want_seq=False, want_seq_bf16=False, hidden=None, scope="RNN"):
  """MultiRNNCell[BasicLSTMCell(lstm_cells, forget_bias=1.0)] x lstm_layers under
  dynamic_rnn(sequence_length=num_frames) (wh/all_frame_models/lstm_model.py:30-47).
  Returns (state [B, L*2*H] = [c0, h0, c1, h1, ...], outputs fp32, outputs bf16)."""
  hidden = int(FLAGS.lstm_cells) if hidden is None else hidden
  layers = FLAGS.lstm_layers
  x = ops.frames_operand(model_input)
  d = x.shape[2]
  st = ops.get_store()
  wps, bps = [], []
  for l in range(layers):
    in_dim = d if l == 0 else hidden
    cell = scope + "/multi_rnn_cell/cell_%d/basic_lstm_cell" % l
    w = st.get(cell + "/weights", (in_dim + hidden, 4 * hidden), ops.xavier_uniform)
    b = st.get(cell + "/biases", (4 * hidden,), ops.zeros_init, round_bf16=False)
    wp, bp = st.packed(w, "lstm", lambda w=w, b=b, in_dim=in_dim: nat.lstm_pack(w.value, b.value, in_dim, hidden),
                       version=(w.version, b.version))
    wps.append(wp)
    bps.append(bp)
  nf = num_frames.to(x.device, torch.int32)
  return nat.lstm_fwd(x, nf, wps, bps, hidden, forget_bias=1.0, want_seq=want_seq, want_seq_bf16=want_seq_bf16)


def _split_state(state, layers, hidden):
  """[B, L*2*H] -> per layer (c, h) views."""
  return [(state[:, (2 * l) * hidden:(2 * l + 1) * hidden], state[:, (2 * l + 1) * hidden:(2 * l + 2) * hidden])
          for l in range(layers)]


# --------------------------------------------------------------------------------------------------
# models
# --------------------------------------------------------------------------------------------------

class FrameLevelLogisticModel(models.BaseModel):
  """Mean-pool the valid frames, then a logistic layer (wh/all_frame_models/logistic_model.py:35-46)."""

  def create_model(self, model_input, vocab_size, num_frames, **unused_params):
    x = ops.frames_operand(model_input)
    b, t, d = x.shape
    nf = num_frames.to(x.device, torch.int32)
    # mean over valid frames == attention pooling with equal logits
    pooled, _, _ = nat.attn_pool(torch.zeros((b, t, 1), device=x.device), x, nf, 1, 0, want_bf16=False)
    out = ops.fully_connected(pooled.reshape(b, d), vocab_size, "fully_connected", activation_fn="sigmoid",
                              l2_penalty=1e-8, want_bf16=False)
    return {"predictions": out.f32}


class LstmModel(models.BaseModel):
  """wh/all_frame_models/lstm_model.py:13-57: the classifier input is the non-tuple state [c0,h0,c1,h1]."""

  def create_model(self, model_input, vocab_size, num_frames, **unused_params):
    state, _, _ = _lstm_stack(model_input, num_frames)
    return _classifier().create_model(model_input=state, original_input=model_input, vocab_size=vocab_size,
                                      **unused_params)


class LstmMemoryModel(models.BaseModel):
  """wh/all_frame_models/lstm_memory_model.py:13-73: classifier input = concat of the layers' c states."""

  def create_model(self, model_input, vocab_size, num_frames, dropout=False, keep_prob=None, noise_level=None,
                   **unused_params):
    if dropout or noise_level is not None:
      raise NotImplementedError("dropout / noise_level are training-time extras outside the hot path")
    hidden, layers = int(FLAGS.lstm_cells), FLAGS.lstm_layers
    state, _, _ = _lstm_stack(model_input, num_frames)
    final_state = torch.cat([c for c, _ in _split_state(state, layers, hidden)], dim=1)   # device copy
    return _classifier().create_model(model_input=final_state, original_input=model_input, vocab_size=vocab_size,
                                      num_frames=num_frames, **unused_params)


class LstmParallelFinaloutputModel(models.BaseModel):
  """wh/all_frame_models/lstm_parallel_finaloutput_model.py:13-73: one LSTM stack PER MODALITY (--feature_names / --feature_sizes,
  --lstm_cells="1024,128"): every slice of the frame features is L2-normalised per frame (:36), runs through its own
  MultiRNNCell under variable scope RNN<i> (:45-56), and the classifier sees the concatenation of the h states of every layer of
  every stack (:58-61)."""

  def create_model(self, model_input, vocab_size, num_frames, **unused_params):
    import utils
    lstm_sizes = [int(v) for v in str(FLAGS.lstm_cells).split(",")]
    _, feature_sizes = utils.GetListOfFeatureNamesAndSizes(FLAGS.feature_names, FLAGS.feature_sizes)
    assert len(lstm_sizes) == len(feature_sizes), \
        "length of lstm_sizes (={}) != length of feature_sizes (={})".format(len(lstm_sizes), len(feature_sizes))
    x = ops.frames_operand(model_input)
    b, t, d = x.shape
    if sum(feature_sizes) != d:
      raise ValueError("--feature_sizes add up to %d, the input has %d features" % (sum(feature_sizes), d))
    nf = num_frames.to(x.device, torch.int32)
    layers = FLAGS.lstm_layers
    hs, off = [], 0
    for i, (size, hidden) in enumerate(zip(feature_sizes, lstm_sizes)):
      sub = nat.l2norm_rows(x[:, :, off:off + size].contiguous())          # tf.split (device copy) + tf.nn.l2_normalize(dim=2)
      off += size
      state, _, _ = _lstm_stack(sub, nf, hidden=hidden, scope="RNN%d" % i)
      hs.extend(h for _, h in _split_state(state, layers, hidden))
    final_state = torch.cat(hs, dim=1)                                     # tf.concat (device copy)
    return _classifier().create_model(model_input=final_state, original_input=model_input, vocab_size=vocab_size,
                                      **unused_params)


class CnnDeepCombineChainModel(models.BaseModel):
  """wh/all_frame_models/cnn_deep_combine_chain_model.py:10-88: a temporal CNN (filters over 1, 2 and 3 consecutive frames, max
  over time) feeds a chain of MoE sub-predictions; every stage sees the mean-pooled input, a FRESH CNN descriptor and the
  ReLU-projected, L2-normalised predictions of all earlier stages.

  The convolution is a dense GEMM: the operand row of frame t is [x_t, x_{t-1}, x_{t-2}] (zero rows shifted in at the front,
  :24-29), built once and shared by every filter length (its first D*fs columns) and by all deep_chain_layers + 1 CNN stages --
  their filters of one length are stacked along the output axis, so the whole model runs THREE tensor-core GEMMs over the
  B*T frame rows, followed by the max over time (over all max_frames rows, as the reference's unmasked tf.reduce_max does)."""

  FILTER_SIZES = (1, 2, 3)

  def create_model(self, model_input, vocab_size, num_frames, num_mixtures=None, l2_penalty=1e-8, sub_scope="",
                   original_input=None, **unused_params):
    num_layers = FLAGS.deep_chain_layers
    relu_cells = FLAGS.deep_chain_relu_cells
    num_filters = (relu_cells, relu_cells, relu_cells * 2)
    x = ops.frames_operand(model_input)
    b, t, d = x.shape
    nf = num_frames.to(x.device, torch.int32)
    st = ops.get_store()
    # mean over the valid frames == attention pooling with equal logits (einsum("ijk,ij->ik") / num_frames, :55-57)
    pooled, _, _ = nat.attn_pool(torch.zeros((b, t, 1), device=x.device), x, nf, 1, 0, want_bf16=False)
    mean_input = ops.Act(f32=pooled.reshape(b, d))
    mean_relu = ops.fully_connected(mean_input, relu_cells, sub_scope + "mean-relu", activation_fn="relu", l2_penalty=l2_penalty,
                                    want_bf16=False)
    relu_layers = [ops.l2_normalize_rows(mean_relu)]
    # shifted-concat operand [B, T, 3D]: columns [i*D, (i+1)*D) hold the frames shifted down by i (device copies, no arithmetic)
    fs_max = max(self.FILTER_SIZES)
    xcat = torch.zeros((b, t, fs_max * d), dtype=x.dtype, device=x.device)
    for i in range(fs_max):
      xcat[:, i:, i * d:(i + 1) * d] = x[:, :t - i]
    rows = xcat.reshape(b * t, fs_max * d)
    # every CNN stage's filters of one length, stacked along the output axis: [D*fs, (L+1)*nf]
    pooled_cnn = []
    for fs, nfil in zip(self.FILTER_SIZES, num_filters):
      ws = [st.get(sub_scope + "cnn%dcnn-filter-len%d" % (l, fs), (d * fs, nfil), ops.truncated_normal(0.1), l2=l2_penalty)
            for l in range(num_layers + 1)]
      wp = st.packed(ws[0], "cnn_stack_len%d" % fs, lambda ws=ws: nat.pack_transpose(torch.cat([w.value for w in ws], dim=1)),
                     version=tuple(w.version for w in ws))
      out = nat.linear(rows, wp, n=(num_layers + 1) * nfil, k=d * fs)["f32"]            # [B*T, (L+1)*nf]
      pooled_cnn.append(nat.group_max_rows(out, t).reshape(b, num_layers + 1, nfil))     # max over time
    def cnn_descriptor(l):
      cat = torch.cat([pc[:, l] for pc in pooled_cnn], dim=1).contiguous()               # tf.concat over filter lengths
      return ops.l2_normalize_rows(cat)
    next_input = cnn_descriptor(0)
    support_predictions = []
    for layer in range(num_layers):
      sub_prediction = self.sub_model(next_input, vocab_size, sub_scope=sub_scope + "prediction-%d" % layer)
      support_predictions.append(sub_prediction)
      sub_relu = ops.fully_connected(sub_prediction, relu_cells, sub_scope + "relu-%d" % layer, activation_fn="relu",
                                     l2_penalty=l2_penalty, want_bf16=False)
      relu_layers.append(ops.l2_normalize_rows(sub_relu))
      next_input = ops.concat([mean_input, cnn_descriptor(layer + 1)] + relu_layers)
    main = self.sub_model(next_input, vocab_size, sub_scope=sub_scope + "-main")
    return {"predictions": main, "support_predictions": torch.cat(support_predictions, dim=1)}

  def sub_model(self, model_input, vocab_size, num_mixtures=None, l2_penalty=1e-8, sub_scope="", **unused_params):
    num_mixtures = num_mixtures or FLAGS.moe_num_mixtures
    return ops.moe_head(model_input, vocab_size, num_mixtures, "gates-" + sub_scope, "experts-" + sub_scope, l2_penalty)


class LstmAttentionMaxPoolingModel(models.BaseModel):
  """wh/all_frame_models/lstm_attention_max_pooling_model.py:10-98: attention logits from [x_t, h_t],
  softmax over T (masked, renormalised), weighted sum of the LSTM outputs, MoE per head, max over heads."""

  def create_model(self, model_input, vocab_size, num_frames, num_mixtures=None, l2_penalty=1e-8, sub_scope="",
                   original_input=None, **unused_params):
    hidden = int(FLAGS.lstm_cells)
    num_attentions = FLAGS.lstm_attentions
    x = ops.frames_operand(model_input)
    b, t, d = x.shape
    nf = num_frames.to(x.device, torch.int32)
    _, _, out_bf = _lstm_stack(x, nf, want_seq_bf16=True)
    # attention-<sub_scope>/weights is [D + H, A]: the logits GEMM is split over the two operands
    st = ops.get_store()
    scope = "attention-" + sub_scope
    w = st.get(scope + "/weights", (d + hidden, num_attentions), ops.xavier_uniform, l2=l2_penalty)
    bias = st.get(scope + "/biases", (num_attentions,), ops.zeros_init, round_bf16=False)
    wp = st.packed(w, "kmajor", lambda: nat.pack_transpose(w.value))            # [A, D+H]
    cat = torch.cat([x, out_bf], dim=2).reshape(b * t, d + hidden)               # device copy (tf.concat, :52)
    logits = nat.linear(cat, wp, n=num_attentions, k=d + hidden, shift=bias.value)["f32"]
    logits3 = logits.as_strided((b, t, num_attentions), (t * logits.stride(0), logits.stride(0), 1))
    pooled, hi, lo = nat.attn_pool(logits3, out_bf, nf, num_attentions, 0)
    act = ops.Act(f32=pooled.reshape(b * num_attentions, hidden), hi=hi.reshape(b * num_attentions, hidden),
                  lo=lo.reshape(b * num_attentions, hidden))
    moe = self.sub_moe(act, vocab_size, sub_scope="sub-moe")
    return {"predictions": nat.group_max_rows(moe, num_attentions)}

  def sub_moe(self, model_input, vocab_size, num_mixtures=None, l2_penalty=1e-8, sub_scope="", **unused_params):
    num_mixtures = num_mixtures or FLAGS.moe_num_mixtures
    return ops.moe_head(model_input, vocab_size, num_mixtures, "gates-" + sub_scope, "experts-" + sub_scope, l2_penalty)


class LstmMultiAttentionModel(models.BaseModel):
  """wh/all_frame_models/lstm_multi_attention_model.py:13-91: sigmoid attention over the LSTM outputs
  (masked, / (sum + 1e-8)), pools the RAW input, classifier per head, max over heads."""

  def create_model(self, model_input, vocab_size, num_frames, **unused_params):
    hidden = int(FLAGS.lstm_cells)
    attention_size = FLAGS.attention_size
    l2_penalty = unused_params.get("l2_penalty", 1e-8)
    x = ops.frames_operand(model_input)
    b, t, d = x.shape
    nf = num_frames.to(x.device, torch.int32)
    _, _, out_bf = _lstm_stack(x, nf, want_seq_bf16=True)
    st = ops.get_store()
    w = st.get("fully_connected/weights", (hidden, attention_size), ops.xavier_uniform, l2=l2_penalty)
    bias = st.get("fully_connected/biases", (attention_size,), ops.zeros_init, round_bf16=False)
    wp = st.packed(w, "kmajor", lambda: nat.pack_transpose(w.value))
    logits = nat.linear(out_bf.reshape(b * t, hidden), wp, n=attention_size, k=hidden, shift=bias.value)["f32"]
    logits3 = logits.as_strided((b, t, attention_size), (t * logits.stride(0), logits.stride(0), 1))
    pooled, hi, lo = nat.attn_pool(logits3, x, nf, attention_size, 1)
    act = ops.Act(f32=pooled.reshape(b * attention_size, d), hi=hi.reshape(b * attention_size, d),
                  lo=lo.reshape(b * attention_size, d))
    out = _classifier().create_model(model_input=act, original_input=model_input, vocab_size=vocab_size,
                                     **unused_params)["predictions"]
    return {"predictions": nat.group_max_rows(out, attention_size)}


class AttentionModel(models.BaseModel):
  """zt/frame_level_models.py:4355-4405 (LSTM-free multi-head attention pooling).

  logits = [x_t, mean_t(x)] . W + b, softmax over T, times the non-zero-frame mask, renormalised.  The
  mean-pooled half of W and the bias add the same constant to every frame of a video, and softmax
  over T is shift invariant, so only W[:D] reaches the kernels (the variables keep the reference
  shapes for checkpoint compatibility).  The head is FLAGS.video_level_classifier_model on B*A rows
  (MoeExtendModel in the reference's scripts, which takes the max over the A heads)."""

  def create_model(self, model_input, vocab_size, num_frames, l2_penalty=1e-8, **unused_params):
    num_extend = FLAGS.moe_num_extend
    x = ops.frames_operand(model_input)
    b, t, d = x.shape
    st = ops.get_store()
    w = st.get("Attention/W", (2 * d, num_extend), ops.truncated_normal(0.1), l2=l2_penalty)
    st.get("Attention/b", (num_extend,), ops.constant_init(0.1), l2=l2_penalty, round_bf16=False)
    wp = st.packed(w, "kmajor_top", lambda: nat.pack_transpose(w.value[:d].contiguous()))
    with nat.region("attention_pool"):
      if num_extend == 8 and d % 8 == 0 and d <= 4096:
        # one kernel: logits, masked softmax over the frames, weighted sum -- the frames leave HBM once
        pooled, hi, lo = nat.attn_pool_fused(x, wp, None, num_extend)
      else:
        logits = nat.linear(x.reshape(b * t, d), wp, n=num_extend, k=d)["f32"]
        logits3 = logits.as_strided((b, t, num_extend), (t * logits.stride(0), logits.stride(0), 1))
        pooled, hi, lo = nat.attn_pool(logits3, x, None, num_extend, 0)
    act = ops.Act(f32=pooled.reshape(b * num_extend, d), hi=hi.reshape(b * num_extend, d),
                  lo=lo.reshape(b * num_extend, d))
    out = _classifier().create_model(model_input=act, vocab_size=vocab_size, **unused_params)
    if out["predictions"].shape[0] == b * num_extend and num_extend > 1:
      # a row-wise classifier (MoeModel, ChainMoeModel, DeepCombineChainModel -- BASELINE configs[4] "attention pool +
      # chained MoE") scored every head: reduce with the max over heads MoeExtendModel takes (zt/video_level_models.py:2327-2328)
      out = dict(out)
      out["predictions"] = nat.group_max_rows(out["predictions"], num_extend)
      if "support_predictions" in out:
        out["support_predictions"] = nat.group_max_rows(out["support_predictions"].contiguous(), num_extend)
    return out


class DbofModel(models.BaseModel):
  """Deep Bag of Frames (wh/all_frame_models/dbof_model.py:13-124): sample `iterations` frames, BN,
  cluster projection D -> dbof_cluster_size, BN, ReLU6, max/avg pool over the samples, hidden FC, BN,
  ReLU6, classifier.  Batch-norm runs in inference form (moving statistics) this round.

  `frame_index` ([B, iterations] int64) may be passed to pin the sampled frames (the reference draws
  them with tf.random_uniform, wh/model_utils.py:56-74)."""

  def create_model(self, model_input, vocab_size, num_frames, iterations=None, add_batch_norm=None,
                   sample_random_frames=None, cluster_size=None, hidden_size=None, is_training=False,
                   frame_index=None, **unused_params):
    iterations = iterations or FLAGS.iterations
    add_batch_norm = add_batch_norm or FLAGS.dbof_add_batch_norm
    random_frames = sample_random_frames or FLAGS.sample_random_frames
    cluster_size = cluster_size or FLAGS.dbof_cluster_size
    hidden1_size = hidden_size or FLAGS.dbof_hidden_size
    method = FLAGS.dbof_pooling_method
    if method not in ("max", "average"):
      raise ValueError("Unrecognized pooling method: %s" % method)
    x = ops.frames_operand(model_input)
    b, t, d = x.shape
    nf = num_frames.to(x.device)
    if frame_index is None:
      frame_index = sample_frames(nf, iterations, random_frames)
    xs = x[torch.arange(b, device=x.device).unsqueeze(1), frame_index.to(x.device)]      # gather_nd (device copy)
    rows = xs.reshape(b * iterations, d)
    st = ops.get_store()
    cw = st.get("cluster_weights", (d, cluster_size), ops.random_normal(1 / math.sqrt(d)))
    hw = st.get("hidden1_weights", (cluster_size, hidden1_size), ops.random_normal(1 / math.sqrt(cluster_size)))
    if add_batch_norm and is_training:
      # batch statistics (dbof_model.py:64-108 with is_training=True): the statistics sit between the GEMMs, so the layers run
      # un-fused here; the moving averages of the three scopes are updated in place
      if method != "max":
        raise NotImplementedError("is_training=True is built for --dbof_pooling_method=max")
      cwp = st.packed(cw, "kmajor", lambda: nat.pack_transpose(cw.value))
      hwp = st.packed(hw, "kmajor", lambda: nat.pack_transpose(hw.value))
      r = _bn_train("input_bn", rows.contiguous(), None)
      z = nat.linear(r["hi"], cwp, a_lo=r["lo"], n=cluster_size, k=d)["f32"]
      act = _bn_train("cluster_bn", z, "relu6", want_bf16=False)["f32"]
      hi, lo = nat.split_bf16(nat.group_max_rows(act, iterations))
      z3 = nat.linear(hi, hwp, a_lo=lo, n=hidden1_size, k=cluster_size)["f32"]
      hidden = _bn_train("hidden1_bn", z3, "relu6")
      act = ops.Act(f32=hidden["f32"], hi=hidden["hi"], lo=hidden["lo"], cols=hidden1_size)
      return _classifier().create_model(model_input=act, original_input=model_input, vocab_size=vocab_size, **unused_params)
    rows_lo = None
    if add_batch_norm:
      s_in, t_in = _bn_affine("input_bn", d, is_training)
      rows, rows_lo = nat.col_affine(rows.contiguous(), s_in, t_in)        # input_bn applied exactly (hi + lo)
      scale, shift = _bn_affine("cluster_bn", cluster_size, is_training)
      s_h, t_h = _bn_affine("hidden1_bn", hidden1_size, is_training)
    else:
      scale = None
      shift = st.get("cluster_biases", (cluster_size,), ops.random_normal(1 / math.sqrt(d)), round_bf16=False).value
      s_h = None
      t_h = st.get("hidden1_biases", (hidden1_size,), ops.random_normal(0.01), round_bf16=False).value
    cwp = st.packed(cw, "kmajor", lambda: nat.pack_transpose(cw.value))
    act = nat.linear(rows, cwp, a_lo=rows_lo, n=cluster_size, k=d, scale=scale, shift=shift, act="relu6")["f32"]
    if method == "max":
      pooled = nat.group_max_rows(act, iterations)
    else:
      pooled, _, _ = nat.attn_pool(torch.zeros((b, iterations, 1), device=x.device),
                                   nat.l2norm_rows(act, normalize=False).reshape(b, iterations, -1), None, 1, 0,
                                   want_bf16=False)
      pooled = pooled.reshape(b, cluster_size)
    hwp = st.packed(hw, "kmajor", lambda: nat.pack_transpose(hw.value))
    hi, lo = nat.split_bf16(pooled)
    hidden = nat.linear(hi, hwp, a_lo=lo, n=hidden1_size, k=cluster_size, scale=s_h, shift=t_h, act="relu6",
                        out_bf16=True, out_lo=True)
    act = ops.Act(f32=hidden["f32"], hi=hidden["hi"], lo=hidden["lo"], cols=hidden1_size)
    return _classifier().create_model(model_input=act, original_input=model_input, vocab_size=vocab_size,
                                      **unused_params)


def sample_frames(num_frames, num_samples, random_frames=True, generator=None):
  """wh/model_utils.py:23-74 (SampleRandomSequence / SampleRandomFrames): index selection only."""
  b = num_frames.shape[0]
  nf = num_frames.to(torch.float32).unsqueeze(1)
  dev = num_frames.device
  if random_frames:
    return (torch.rand((b, num_samples), device=dev, generator=generator) * nf).to(torch.int64)
  max_start = torch.clamp(nf - num_samples, min=0)
  start = (torch.rand((b, 1), device=dev, generator=generator) * (max_start + 1)).to(torch.int64)
  idx = start + torch.arange(num_samples, device=dev).unsqueeze(0)
  return torch.minimum(idx, (nf - 1).to(torch.int64))


class NetVLADModel(models.BaseModel):
  """NetVLAD pooling + hidden FC + classifier.  NOT part of /root/reference: follows
  Miech, Laptev, Sivic 2017 in the idiom of DbofModel; definition pinned by oracle/yt8m_oracle.py
  (netvlad_pool).  The soft-assignment GEMM, masked softmax, residual aggregation GEMM, intra-norm and
  final L2 norm are ONE fused kernel (csrc/yt8m_netvlad.cu)."""

  gating = False

  def create_model(self, model_input, vocab_size, num_frames, cluster_size=None, hidden_size=None,
                   add_batch_norm=None, is_training=False, **unused_params):
    cluster_size = cluster_size or FLAGS.netvlad_cluster_size
    hidden1_size = hidden_size or FLAGS.netvlad_hidden_size
    add_batch_norm = FLAGS.netvlad_add_batch_norm if add_batch_norm is None else add_batch_norm
    x = ops.frames_operand(model_input)
    b, t, d = x.shape
    nf = num_frames.to(x.device, torch.int32)
    st = ops.get_store()
    cw = st.get("cluster_weights", (d, cluster_size), ops.random_normal(1 / math.sqrt(d)))
    cw2 = st.get("cluster_weights2", (d, cluster_size), ops.random_normal(1 / math.sqrt(d)), round_bf16=False)
    if add_batch_norm:
      scale, shift = _bn_affine("cluster_bn", cluster_size, is_training)
    else:
      scale = None
      shift = st.get("cluster_biases", (cluster_size,), ops.random_normal(1 / math.sqrt(d)), round_bf16=False).value
    cwp = st.packed(cw, "kmajor", lambda: nat.pack_transpose(cw.value))
    c2split = st.packed(cw2, "hi_lo", lambda: nat.split_bf16(cw2.value))     # residual centres as tensor-core operands
    if FLAGS.netvlad_operand_format not in ("f16", "bf16x2"):
      raise ValueError("--netvlad_operand_format must be 'f16' or 'bf16x2'")
    f16 = FLAGS.netvlad_operand_format == "f16"
    hw = st.get("hidden1_weights", (cluster_size * d, hidden1_size), ops.random_normal(1 / math.sqrt(cluster_size)))
    tiled = f16 and nat.netvlad_tiled_supported(t, d, cluster_size)
    if tiled:
      # the one-pass four-CTA-cluster kernel writes the descriptor in a blocked order (include/yt8m_b200.h); the hidden layer's
      # weight ROWS are gathered into the same order once, at packing time -- the product v . W is unchanged
      c2t = st.packed(cw2, "tiled", lambda: cw2.value.reshape(-1)[nat.netvlad_tiled_index(d, cluster_size, 4, cw2.value.device)].contiguous())
      vlad_hi, vlad_lo = nat.netvlad_fwd_tiled(x, nf, cwp, scale, shift, c2t, out_f16=True), None
      hwp = st.packed(hw, "kmajor_f16_tiled", lambda: nat.pack_transpose(
          hw.value[nat.netvlad_tiled_index(d, cluster_size, 8, hw.value.device)]).to(torch.float16))
    else:
      vlad_hi, vlad_lo, _ = nat.netvlad_fwd(x, nf, cwp, scale, shift, cw2.value, want_lo=not f16, cw2_split=c2split, out_f16=f16)
    if tiled:
      pass
    elif f16:   # the fp16 conversion of the packed weights (a dtype conversion: bf16 values >= 2^-17 are exact in fp16)
      hwp = st.packed(hw, "kmajor_f16", lambda: nat.pack_transpose(hw.value).to(torch.float16))
    else:
      hwp = st.packed(hw, "kmajor", lambda: nat.pack_transpose(hw.value))
    if add_batch_norm:
      s_h, t_h = _bn_affine("hidden1_bn", hidden1_size, is_training)
    else:
      s_h = None
      t_h = st.get("hidden1_biases", (hidden1_size,), ops.random_normal(0.01), round_bf16=False).value
    hidden = nat.linear(vlad_hi, hwp, a_lo=vlad_lo, n=hidden1_size, k=cluster_size * d, scale=s_h, shift=t_h,
                        act="relu6" if FLAGS.netvlad_relu else None, out_f32=self.gating or not f16, out_bf16=not f16,
                        out_lo=not f16, out_f16=f16)
    act = ops.Act(f32=hidden.get("f32"), hi=hidden["hi"], lo=hidden.get("lo"), cols=hidden1_size)
    if self.gating:
      gw = st.get("gating_weights", (hidden1_size, hidden1_size), ops.random_normal(1 / math.sqrt(hidden1_size)))
      if f16:
        gwp = st.packed(gw, "kmajor_f16", lambda: nat.pack_transpose(gw.value).to(torch.float16))
      else:
        gwp = st.packed(gw, "kmajor", lambda: nat.pack_transpose(gw.value))
      if add_batch_norm:
        s_g, t_g = _bn_affine("gating_bn", hidden1_size, is_training)
      else:
        s_g = None
        t_g = st.get("gating_biases", (hidden1_size,), ops.random_normal(1 / math.sqrt(hidden1_size)),
                     round_bf16=False).value
      g = nat.linear(act.hi, gwp, a_lo=act.lo, n=hidden1_size, k=hidden1_size)["f32"]
      y, yh, yl = nat.context_gate(act.f32.contiguous(), g.contiguous(), s_g, t_g)
      act = ops.Act(f32=y, hi=yh, lo=yl, cols=hidden1_size)
    return _classifier().create_model(model_input=act, original_input=model_input, vocab_size=vocab_size,
                                      **unused_params)


class GatedNetVLADModel(NetVLADModel):
  """NetVLAD + context gating y = x * sigmoid(BN(x . Wg)) before the classifier (same source)."""
  gating = True
