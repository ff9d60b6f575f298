"""CPU fp32 oracle for the yt8m hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product path
(``youtube-8m_b200/``) never does; it fails loudly when the CUDA library is missing.

Every function restates, in plain torch-CPU float32 (float64 on request), the arithmetic
of the reference lines cited in its docstring (paths relative to ``/root/reference``;
``wh/`` = ``youtube-8m-wangheda/``, ``zt/`` = ``youtube-8m-zhangteng/``).  The reference
executes these inside TensorFlow 1.0, which is NOT vendored under ``/root/reference``
(``.gitmodules:4-6`` declares the submodule, ``.SUBMODULES.json:8`` records it empty) and is not
installable here, and the reference holds no tests or golden vectors for this path
(SURVEY.md §4).  So for the model arithmetic:

    PINNED TO THE REFERENCE SOURCE, NOT TO A TENSORFLOW RUN -- ``oracle/make_model_golden.py`` exec's the reference's
    own create_model() / calculate_loss() / Dequantize code against ``oracle/tf_numpy_shim.py`` (a numpy stand-in for
    the TF / slim ops it calls, following TF-1.0's documented semantics) and
    ``tests/test_oracle_golden_models.py`` holds this module to those outputs (16 cases, 3e-6); TensorFlow's own
    kernels are not in the loop, so closed-form known-answer tests (tests/test_oracle_kat.py) and a torch.nn.LSTM
    cross-check back the shim's op semantics.  NetVLAD / context gating: PARITY UNPINNED (no upstream definition).

The metric path (GAP@20 / Hit@1 / PERR) IS pinned: ``oracle/gap_oracle.py`` is checked against
golden vectors produced by the reference's own ``average_precision_calculator.py`` /
``eval_util.py`` run in the build container (``oracle/make_golden.py``, fixtures in
``tests/golden/``).

NetVLAD and context gating do not exist in ``/root/reference`` at all; the
functions below follow the published definition (Miech, Laptev, Sivic 2017) in the idiom of
``wh/all_frame_models/dbof_model.py`` and are labelled as such.

Weights use the reference (TF) layout: matrices are ``[in, out]``.
"""
import math

import torch

DT = torch.float32


def bf16r(x):
  """Round to the nearest bf16-representable value, keep the working dtype.

  Both the oracle and the CUDA path see inputs/weights that are exactly representable in
  bf16 ("identical synthetic inputs"); all oracle arithmetic after that is fp32.
  """
  return x.to(torch.bfloat16).to(x.dtype)


# ----------------------------------------------------------------------------------------
# elementary TF-1.0 ops
# ----------------------------------------------------------------------------------------

def l2_normalize(x, dim=-1, eps=1e-12):
  """tf.nn.l2_normalize: x * rsqrt(max(sum(x^2), eps)).

  Used by wh/all_feature_transform/default_transformer.py:5-8 (every frame row) and
  wh/all_video_models/deep_combine_chain_model.py:46.
  """
  ss = (x * x).sum(dim=dim, keepdim=True)
  return x * torch.rsqrt(torch.clamp(ss, min=eps))


def dequantize(u8, max_quantized_value=2.0, min_quantized_value=-2.0):
  """wh/utils.py:23-38 with the reader's (max=2, min=-2) (wh/readers.py:191-192)."""
  rng = max_quantized_value - min_quantized_value
  scalar = rng / 255.0
  bias = (rng / 512.0) + min_quantized_value
  return u8.to(DT) * scalar + bias


def fully_connected(x, w, b=None, act=None):
  """slim.fully_connected: act(x.W + b), W: [in, out]; rank-3 input contracts the last dim."""
  y = x @ w
  if b is not None:
    y = y + b
  if act is not None:
    y = act(y)
  return y


def l2_regularizer(w, scale=1e-8):
  """slim.l2_regularizer(scale)(w) = scale * sum(w^2) / 2."""
  return scale * (w * w).sum() / 2.0


def relu6(x):
  return torch.clamp(x, 0.0, 6.0)


def batch_norm(x, gamma, beta, moving_mean, moving_var, is_training, eps=1e-3, decay=0.999):
  """slim.batch_norm(center=True, scale=True) as called at wh/all_frame_models/dbof_model.py:65-108.

  x: [rows, C].  Training mode normalises with the biased batch variance and returns updated
  moving statistics (decay 0.999); inference mode uses the moving statistics.
  Returns (y, new_moving_mean, new_moving_var).
  """
  if is_training:
    mean = x.mean(dim=0)
    var = x.var(dim=0, unbiased=False)
    y = (x - mean) * torch.rsqrt(var + eps) * gamma + beta
    new_mm = moving_mean * decay + mean.detach() * (1 - decay)
    new_mv = moving_var * decay + var.detach() * (1 - decay)
    return y, new_mm, new_mv
  y = (x - moving_mean) * torch.rsqrt(moving_var + eps) * gamma + beta
  return y, moving_mean, moving_var


# ----------------------------------------------------------------------------------------
# video-level heads
# ----------------------------------------------------------------------------------------

def logistic_model(x, w, b):
  """wh/all_video_models/logistic_model.py:23-26: sigmoid(x.W + b), W: [D, V]."""
  return torch.sigmoid(x @ w + b)


def moe_model(x, gate_w, expert_w, expert_b, vocab_size, num_mixtures):
  """wh/all_video_models/moe_model.py:38-65.

  gate_w: [D, V*(M+1)] (no bias, :40-46); expert_w: [D, V*M], expert_b: [V*M] (:47-52).
  Column layout is class-major / mixture-minor (row-major reshape at :54-59).
  """
  v, m = vocab_size, num_mixtures
  gate = (x @ gate_w).reshape(-1, m + 1)
  expert = (x @ expert_w + expert_b).reshape(-1, m)
  gating = torch.softmax(gate, dim=1)
  experts = torch.sigmoid(expert)
  p = (gating[:, :m] * experts).sum(dim=1)
  return p.reshape(-1, v)


def moe_extend_model(x, gate_w, expert_w, expert_b, vocab_size, num_mixtures, num_extend):
  """zt/video_level_models.py:2299-2330: MoE on B*A rows, then max over the A heads."""
  p = moe_model(x, gate_w, expert_w, expert_b, vocab_size, num_mixtures)
  return p.reshape(-1, num_extend, vocab_size).max(dim=1).values


def chain_moe_model(x, support, main, vocab_size, num_supports, num_mixtures):
  """wh/all_video_models/chain_moe_model.py:12-18.

  support / main: dicts with gate_w, expert_w, expert_b.  Returns (predictions, support_predictions).
  """
  sp = moe_model(x, support["gate_w"], support["expert_w"], support["expert_b"], num_supports, num_mixtures)
  main_in = torch.cat([x, sp], dim=1)
  p = moe_model(main_in, main["gate_w"], main["expert_w"], main["expert_b"], vocab_size, num_mixtures)
  return p, sp


def deep_combine_chain_model(x, layers, main, vocab_size, num_mixtures, relu_type="relu"):
  """wh/all_video_models/deep_combine_chain_model.py:24-49.

  layers: list of dicts {gate_w, expert_w, expert_b, relu_w [V, relu_cells], relu_b}.
  Returns (predictions, support_predictions = concat of the per-layer sub-predictions).
  """
  nxt = x
  supports = []
  for lyr in layers:
    sub = moe_model(nxt, lyr["gate_w"], lyr["expert_w"], lyr["expert_b"], vocab_size, num_mixtures)
    act = sub @ lyr["relu_w"] + lyr["relu_b"]
    act = torch.nn.functional.elu(act) if relu_type == "elu" else torch.relu(act)
    nxt = torch.cat([nxt, l2_normalize(act, dim=1)], dim=1)
    supports.append(sub)
  p = moe_model(nxt, main["gate_w"], main["expert_w"], main["expert_b"], vocab_size, num_mixtures)
  return p, torch.cat(supports, dim=1)


# ----------------------------------------------------------------------------------------
# LSTM (tf.contrib.rnn.BasicLSTMCell + MultiRNNCell + tf.nn.dynamic_rnn, TF 1.0 semantics)
# ----------------------------------------------------------------------------------------

def basic_lstm_cell(x, c, h, w, b, forget_bias=1.0):
  """BasicLSTMCell.__call__ (TF 1.0): one matrix [in+H, 4H], input order [x, h], gate split
  order i, j, f, o; c' = c*sigmoid(f + forget_bias) + sigmoid(i)*tanh(j); h' = tanh(c')*sigmoid(o).
  Call site: wh/all_frame_models/lstm_model.py:34-47.
  """
  g = torch.cat([x, h], dim=1) @ w + b
  i, j, f, o = g.chunk(4, dim=1)
  c2 = c * torch.sigmoid(f + forget_bias) + torch.sigmoid(i) * torch.tanh(j)
  h2 = torch.tanh(c2) * torch.sigmoid(o)
  return c2, h2


def dynamic_rnn_lstm(x, num_frames, layers, forget_bias=1.0):
  """tf.nn.dynamic_rnn(MultiRNNCell[BasicLSTMCell]*L, sequence_length=num_frames), zero initial state.

  For t >= num_frames[b] the output row is zero and the state row is carried through unchanged.
  x: [B, T, D]; layers: list of (w [in+H, 4H], b [4H]).
  Returns (outputs [B, T, H] of the top layer, [(c_l, h_l)] final states).
  """
  bsz, t_max, _ = x.shape
  hid = layers[0][0].shape[1] // 4
  cs = [torch.zeros(bsz, hid, dtype=x.dtype) for _ in layers]
  hs = [torch.zeros(bsz, hid, dtype=x.dtype) for _ in layers]
  outs = []
  for t in range(t_max):
    live = (t < num_frames).to(x.dtype).unsqueeze(1)
    inp = x[:, t, :]
    for l, (w, b) in enumerate(layers):
      c2, h2 = basic_lstm_cell(inp, cs[l], hs[l], w, b, forget_bias)
      cs[l] = live * c2 + (1 - live) * cs[l]
      hs[l] = live * h2 + (1 - live) * hs[l]
      inp = h2
    outs.append(live * inp)
  return torch.stack(outs, dim=1), list(zip(cs, hs))


def lstm_model_state(states):
  """wh/all_frame_models/lstm_model.py:34-52: state_is_tuple=False => [c0, h0, c1, h1] (4096-d)."""
  return torch.cat([t for ch in states for t in ch], dim=1)


def lstm_memory_model_state(states):
  """wh/all_frame_models/lstm_memory_model.py:61: concat of the c states (2048-d)."""
  return torch.cat([c for c, _ in states], dim=1)


# ----------------------------------------------------------------------------------------
# attention pooling
# ----------------------------------------------------------------------------------------

def sequence_mask(num_frames, max_frames, dtype=DT):
  """tf.sequence_mask (wh/all_frame_models/lstm_attention_max_pooling_model.py:34)."""
  return (torch.arange(max_frames).unsqueeze(0) < num_frames.unsqueeze(1)).to(dtype)


def attention_softmax_pool(x, outputs, num_frames, att_w, att_b):
  """wh/all_frame_models/lstm_attention_max_pooling_model.py:51-63.

  logits = [x, outputs].Wa + ba ([B,T,A]); softmax over T (dim=1, :59); times mask; renormalise
  over T (:60); pooled[b,a,:] = sum_t w[b,a,t] * outputs[b,t,:] (:63).  Returns [B, A, H].
  """
  logits = torch.cat([x, outputs], dim=2) @ att_w + att_b
  mask = sequence_mask(num_frames, x.shape[1], x.dtype)
  w = torch.softmax(logits, dim=1) * mask.unsqueeze(2)
  w = w / w.sum(dim=1, keepdim=True)
  return torch.einsum("bta,bth->bah", w, outputs)


def attention_sigmoid_pool(x, outputs, num_frames, att_w, att_b):
  """wh/all_frame_models/lstm_multi_attention_model.py:66-78.

  att = sigmoid(outputs.Wa + b) * mask; att /= (sum_t att + 1e-8); pooled = einsum(att, x)
  over the RAW input (:78).  Returns [B, A, D].
  """
  mask = sequence_mask(num_frames, x.shape[1], x.dtype).unsqueeze(2)
  att = torch.sigmoid(outputs @ att_w + att_b) * mask
  att = att / (att.sum(dim=1, keepdim=True) + 1e-8)
  return torch.einsum("bta,btd->bad", att, x)


def attention_model_pool(x, num_frames, w, b):
  """zt/frame_level_models.py:4372-4398 (LSTM-free AttentionModel).

  mask = frame has a non-zero entry (:4372-4375); mean = sum_t x / num_frames (:4380-4384);
  logits = [x_t, mean].W + b (W: [2D, A]); softmax over T, times mask, renormalise (:4393-4395);
  state[b,a,:] = sum_t atten[b,t,a] * x[b,t,:] (:4397).  Returns [B*A, D].
  """
  bsz, t_max, d = x.shape
  fmask = (x.abs().sum(dim=2) > 0).to(x.dtype)
  avg = x.sum(dim=1) / num_frames.to(x.dtype).unsqueeze(1)
  cat = torch.cat([x, avg.unsqueeze(1).expand(bsz, t_max, d)], dim=2)
  out = torch.softmax(cat @ w + b, dim=1) * fmask.unsqueeze(2)
  att = out / out.sum(dim=1, keepdim=True)
  state = torch.einsum("bta,btd->bad", att, x)
  return state.reshape(-1, d)


def shifted_concat_cnn(x, filters):
  """wh/all_frame_models/cnn_deep_combine_chain_model.py:13-41 (CnnDeepCombineChainModel.cnn): for every filter
  (length fs, matrix [D*fs, nf]) the input is concatenated with its copies shifted DOWN by 1 .. fs-1 frames (zero rows shifted
  in at the front, tf.pad + slice, :24-29) and contracted over the D*fs axis ("ijk,kl->ijl", :38); the outputs are concatenated
  along the channel axis.  x: [B, T, D]; filters: list of (fs, W)."""
  bsz, t_max, d = x.shape
  shifts = [x] + [torch.cat([torch.zeros(bsz, i, d, dtype=x.dtype), x[:, :t_max - i]], dim=1) for i in range(1, max(f for f, _ in filters))]
  return torch.cat([torch.cat(shifts[:fs], dim=2) @ w for fs, w in filters], dim=2)


def cnn_deep_combine_chain_model(x, num_frames, p, vocab_size, num_mixtures, num_layers):
  """wh/all_frame_models/cnn_deep_combine_chain_model.py:43-88.  p: mean_relu_w/b; cnn[l] = list of (fs, W) for l = 0..num_layers;
  layers[l] = MoE (gate_w, expert_w, expert_b) + relu_w / relu_b; main = MoE.  The max over time runs over ALL max_frames rows
  (:67, :83: tf.reduce_max without a mask -- padded rows contribute their shifted neighbours and zeros).
  Returns (predictions [B, V], support_predictions [B, V * num_layers])."""
  bsz, t_max, d = x.shape
  mask = sequence_mask(num_frames, t_max, x.dtype)
  mean_input = torch.einsum("ijk,ij->ik", x, mask) / num_frames.to(x.dtype).unsqueeze(1)
  relu_layers = [l2_normalize(torch.relu(mean_input @ p["mean_relu_w"] + p["mean_relu_b"]), dim=1)]

  def pooled_cnn(l):
    return l2_normalize(shifted_concat_cnn(x, p["cnn"][l]).max(dim=1).values, dim=1)

  next_input = pooled_cnn(0)
  supports = []
  for l in range(num_layers):
    lyr = p["layers"][l]
    sub = moe_model(next_input, lyr["gate_w"], lyr["expert_w"], lyr["expert_b"], vocab_size, num_mixtures)
    supports.append(sub)
    relu_layers.append(l2_normalize(torch.relu(sub @ lyr["relu_w"] + lyr["relu_b"]), dim=1))
    next_input = torch.cat([mean_input, pooled_cnn(l + 1)] + relu_layers, dim=1)
  main = moe_model(next_input, p["main"]["gate_w"], p["main"]["expert_w"], p["main"]["expert_b"], vocab_size, num_mixtures)
  return main, torch.cat(supports, dim=1)


def lstm_parallel_finaloutput_state(x, num_frames, feature_sizes, stacks):
  """wh/all_frame_models/lstm_parallel_finaloutput_model.py:32-64: the input is split by modality (:36), every slice is
  L2-normalised per frame (:36), run through its own MultiRNNCell stack (tuple state), and the h states of every layer of every
  stack are concatenated (:58-61).  stacks[i]: list of (W, b) per layer of modality i."""
  hs = []
  off = 0
  for size, layers in zip(feature_sizes, stacks):
    sub = l2_normalize(x[:, :, off:off + size], dim=2)
    off += size
    _, states = dynamic_rnn_lstm(sub, num_frames, layers)
    hs.extend(h for _, h in states)
  return torch.cat(hs, dim=1)


# ----------------------------------------------------------------------------------------
# DBoF (wh/all_frame_models/dbof_model.py) and NetVLAD (not in the reference)
# ----------------------------------------------------------------------------------------

def dbof_pool(x, frame_index, p, is_training=False, add_batch_norm=True, pooling="max"):
  """wh/all_frame_models/dbof_model.py:62-115 with the sampled frame indices given explicitly
  (the reference draws them with tf.random_uniform, wh/model_utils.py:56-74).

  p: dict with cluster_w [D, C], hidden_w [C, Hd] and either BN params
  (input_bn/cluster_bn/hidden1_bn = dicts gamma, beta, mean, var) or cluster_b / hidden_b.
  Returns the hidden activation [B, Hd] that feeds the video-level classifier (:117-123).
  """
  bsz = x.shape[0]
  xs = x[torch.arange(bsz).unsqueeze(1), frame_index]          # gather_nd
  n = xs.shape[1]
  r = xs.reshape(-1, xs.shape[2])
  if add_batch_norm:
    bn = p["input_bn"]
    r, _, _ = batch_norm(r, bn["gamma"], bn["beta"], bn["mean"], bn["var"], is_training)
  act = r @ p["cluster_w"]
  if add_batch_norm:
    bn = p["cluster_bn"]
    act, _, _ = batch_norm(act, bn["gamma"], bn["beta"], bn["mean"], bn["var"], is_training)
  else:
    act = act + p["cluster_b"]
  act = relu6(act).reshape(bsz, n, -1)
  act = act.max(dim=1).values if pooling == "max" else act.mean(dim=1)   # wh/model_utils.py:76-94
  act = act @ p["hidden_w"]
  if add_batch_norm:
    bn = p["hidden1_bn"]
    act, _, _ = batch_norm(act, bn["gamma"], bn["beta"], bn["mean"], bn["var"], is_training)
  else:
    act = act + p["hidden_b"]
  return relu6(act)


def netvlad_pool(x, num_frames, cluster_w, cluster_scale, cluster_shift, cluster_w2):
  """NetVLAD aggregation.  NOT derived from /root/reference (no such model there):
  follows Miech/Laptev/Sivic 2017 ("Learnable pooling with Context Gating") as recorded in
  SURVEY.md §8(c), in the idiom of wh/all_frame_models/dbof_model.py:73-92.  PARITY UNPINNED.

    logits = (X . cluster_w) * cluster_scale + cluster_shift     ([B*T, K]; the affine is the folded
             batch-norm (gamma*rsqrt(var+eps), beta - mean*that) or (1, cluster_biases))
    a      = softmax_K(logits), zeroed for padded frames t >= num_frames[b]
    a_sum  = sum_t a[t, k]
    vlad[d, k] = sum_t a[t, k] * x[t, d] - a_sum[k] * cluster_w2[d, k]
    intra-normalise each cluster column over d, flatten D-major/K-minor ([B, D*K]), L2-normalise.
  """
  bsz, t_max, d = x.shape
  k = cluster_w.shape[1]
  logits = (x.reshape(-1, d) @ cluster_w) * cluster_scale + cluster_shift
  a = torch.softmax(logits, dim=1).reshape(bsz, t_max, k)
  a = a * sequence_mask(num_frames, t_max, x.dtype).unsqueeze(2)
  a_sum = a.sum(dim=1, keepdim=True)                        # [B, 1, K]
  vlad = torch.einsum("btk,btd->bdk", a, x) - a_sum * cluster_w2.unsqueeze(0)
  vlad = l2_normalize(vlad, dim=1)
  return l2_normalize(vlad.reshape(bsz, d * k), dim=1)


def context_gating(x, gate_w, gate_scale, gate_shift):
  """Context gating y = x * sigmoid(affine(x . Wg)).  NOT derived from /root/reference (same
  source as netvlad_pool); the affine is the folded batch-norm or (1, bias).  PARITY UNPINNED."""
  return x * torch.sigmoid((x @ gate_w) * gate_scale + gate_shift)


# ----------------------------------------------------------------------------------------
# loss / optimiser glue
# ----------------------------------------------------------------------------------------

def cross_entropy_loss(predictions, labels, epsilon=10e-6):
  """wh/losses.py:114-130: mean_b sum_v -[y log(p+eps) + (1-y) log(1-p+eps)], eps = 10e-6 = 1e-5."""
  y = labels.to(predictions.dtype)
  ce = y * torch.log(predictions + epsilon) + (1 - y) * torch.log(1 - predictions + epsilon)
  return (-ce).sum(dim=1).mean()


def support_labels(labels, support_type, num_frequents=200, vertical_mapping=None):
  """wh/losses.py:222-258 (MultiTaskLoss.get_support): "label" = the labels themselves, "frequent" = the first
  num_frequents label columns, "vertical" = (labels . mapping) > 0.2 with a [V, num_verticals] 0/1 mapping, a comma list =
  the column concatenation of its parts."""
  y = labels.to(DT)
  if "," in support_type:
    return torch.cat([support_labels(labels, st, num_frequents, vertical_mapping) for st in support_type.split(",")], dim=1)
  if support_type == "label":
    return y
  if support_type == "frequent":
    return y[:, :num_frequents]
  if support_type == "vertical":
    return ((y @ vertical_mapping.to(DT)) > 0.2).to(DT)
  raise NotImplementedError(support_type)


def multitask_cross_entropy_loss(predictions, support_predictions, labels, sup_labels, support_loss_percent):
  """wh/losses.py:271-279: CE(predictions, labels) * (1 - p) + CE(support_predictions, support_labels) * p."""
  return (cross_entropy_loss(predictions, labels) * (1.0 - support_loss_percent) +
          cross_entropy_loss(support_predictions, sup_labels) * support_loss_percent)


def exponential_decay(base_lr, global_step, batch_size, decay_examples, decay):
  """wh/train.py:303-308, staircase=True: lr * decay ** floor(step*B / decay_examples)."""
  return base_lr * decay ** math.floor(global_step * batch_size / decay_examples)


def clip_by_norm(g, clip):
  """tf.clip_by_norm via wh/utils.py:164-174 (per tensor): g * clip / max(||g||_2, clip)."""
  n = torch.sqrt((g * g).sum())
  return g * clip / torch.maximum(n, torch.tensor(clip, dtype=g.dtype))


def adam_step(param, grad, m, v, step, lr, beta1=0.9, beta2=0.999, eps=1e-8):
  """tf.train.AdamOptimizer (TF 1.0): lr_t = lr*sqrt(1-b2^t)/(1-b1^t);
  theta -= lr_t * m / (sqrt(v) + eps)   (epsilon OUTSIDE the bias-corrected sqrt).
  ``step`` is the 1-based update count.  Returns (param, m, v)."""
  m = beta1 * m + (1 - beta1) * grad
  v = beta2 * v + (1 - beta2) * grad * grad
  lr_t = lr * math.sqrt(1 - beta2 ** step) / (1 - beta1 ** step)
  return param - lr_t * m / (torch.sqrt(v) + eps), m, v
