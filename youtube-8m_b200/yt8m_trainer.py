"""The training step of wh/train.py:262-479 for the video-level heads (LogisticModel, MoeModel), on the GPU
through the C ABI: forward, CrossEntropyLoss + its gradient, backward (fused MoE backward epilogue, MN-major
tcgen05 wgrad), L2 regulariser, per-tensor clip_by_norm, TF-1.0 Adam, exponential-decay learning rate.

Data parallel: one process per GPU; every rank computes gradients of ITS shard of the batch
with the loss gradient pre-scaled by B_local / B_global, then the single flat fp32 gradient buffer is summed over the
ranks once per step (NCCL over NVLink on the GPU box, gloo in the CPU tests; one all-reduce, or -- NetVLAD trainers --
three contiguous pieces started while the backward is still running, yt8m_dp.GradExchange) and every rank holds the
gradient of the mean loss over the global batch -- the single-process large-batch semantics of the reference's build_graph (its own multi-
worker mode is asynchronous parameter-server and is not reproduced).  Parameters and Adam state are
replicated and stay bit-identical across ranks because every rank applies the same update.

Master weights, gradients and Adam moments live in the packed [out, in] layout of the forward kernels;
``export_state()`` / ``import_state()`` convert to and from the reference's TF layouts and variable names.
"""
import math

import torch

import yt8m_native as nat
import yt8m_dp


def exponential_decay(base_lr, global_step, batch_size, decay_examples, decay):
  """tf.train.exponential_decay(base, global_step * batch_size, decay_examples, decay, staircase=True)
  (wh/train.py:303-308)."""
  return base_lr * decay ** math.floor(global_step * batch_size / float(decay_examples))


def adam_lr_t(lr, step, beta1=0.9, beta2=0.999):
  """TF-1.0 AdamOptimizer folds the bias corrections into the step size; `step` is the 1-based update count."""
  return lr * math.sqrt(1.0 - beta2 ** step) / (1.0 - beta1 ** step)


class _mix_loss(object):
  """The --multitask label loss main * (1 - p) + sum(support parts) * p, kept as 1-element device tensors until someone asks
  for the number (float()) -- no synchronisation inside the step."""

  def __init__(self, parts, pct):
    self.parts, self.pct = parts, pct

  def __float__(self):
    main = float(self.parts["main"])
    sup = sum(float(v) for k, v in self.parts.items() if k != "main")
    return main * (1.0 - self.pct) + sup * self.pct


def _moe_row_index(vocab, mixtures):
  """Packed-row index of every reference column: gates [V*(M+1)], experts [V*M] (class-major, mixture-minor)."""
  per = 2 * mixtures + 1
  cpt = 128 // per
  v = torch.arange(vocab)
  base = (v // cpt) * 128 + (v % cpt) * per
  gates = (base.unsqueeze(1) + torch.arange(mixtures + 1).unsqueeze(0)).reshape(-1)
  experts = (base.unsqueeze(1) + (mixtures + 1) + torch.arange(mixtures).unsqueeze(0)).reshape(-1)
  return gates, experts


class HeadTrainer(object):
  """Trains one video-level head.  kind = "logistic" | "moe"."""

  @staticmethod
  def flat_size(kind, in_dim, vocab, mixtures=2):
    rows = int(nat.moe_packed_rows(vocab, mixtures)) if kind == "moe" else vocab
    return rows * nat.pad8(in_dim) + rows

  def __init__(self, kind, in_dim, vocab, mixtures=2, l2_penalty=1e-8, device=None, group=None, storage=None):
    """storage: optional (param, grad, adam_m, adam_v) flat fp32 views of flat_size() elements inside a larger
    buffer (a frame-level trainer keeps ONE flat gradient buffer for the single all-reduce)."""
    assert kind in ("logistic", "moe")
    self.kind, self.d, self.v, self.m = kind, in_dim, vocab, mixtures
    self.l2 = l2_penalty
    self.dev = device or torch.device("cuda", torch.cuda.current_device())
    self.group = group
    self.world = yt8m_dp.world_size(group)
    self.dpad = nat.pad8(in_dim)
    if kind == "moe":
      self.per, self.rows = 2 * mixtures + 1, int(nat.moe_packed_rows(vocab, mixtures))
      if self.rows <= 0 or mixtures not in (1, 2, 4):
        raise ValueError("HeadTrainer: MoE training supports 1, 2 or 4 mixtures")
    else:
      self.per, self.rows = 0, vocab
    n_w, n_b = self.rows * self.dpad, self.rows
    # one flat buffer per role so that the gradient all-reduce is a single collective
    if storage is not None:
      self.param, self.grad, self.adam_m, self.adam_v = storage
      assert all(t.numel() == n_w + n_b and t.dtype == torch.float32 for t in storage)
    else:
      self.param = torch.zeros(n_w + n_b, dtype=torch.float32, device=self.dev)
      self.grad = torch.zeros_like(self.param)
      self.adam_m = torch.zeros_like(self.param)
      self.adam_v = torch.zeros_like(self.param)
    self.w, self.b = self.param[:n_w].view(self.rows, self.dpad), self.param[n_w:]
    self.gw, self.gb = self.grad[:n_w].view(self.rows, self.dpad), self.grad[n_w:]
    self.mw, self.mb = self.adam_m[:n_w].view(self.rows, self.dpad), self.adam_m[n_w:]
    self.vw, self.vb = self.adam_v[:n_w].view(self.rows, self.dpad), self.adam_v[n_w:]
    self.w_bf16 = torch.zeros((self.rows, self.dpad), dtype=torch.bfloat16, device=self.dev)
    self.global_step = 0
    self.last = {}
    self.keep_grads = False

  # ---- reference (TF) layout <-> packed -------------------------------------------------------------
  def import_state(self, sd):
    """sd: reference names/layouts.  logistic: fully_connected/{weights [D,V], biases [V]};
    moe: gates/weights [D,V(M+1)], experts/weights [D,VM], experts/biases [VM]."""
    self.param.zero_()
    if self.kind == "logistic":
      self.w[:, :self.d].copy_(sd["fully_connected/weights"].t())
      self.b.copy_(sd["fully_connected/biases"])
    else:
      g_idx, e_idx = _moe_row_index(self.v, self.m)
      g_idx, e_idx = g_idx.to(self.dev), e_idx.to(self.dev)
      self.w[g_idx, :self.d] = sd["gates/weights"].t().to(self.dev)
      self.w[e_idx, :self.d] = sd["experts/weights"].t().to(self.dev)
      self.b[e_idx] = sd["experts/biases"].to(self.dev)
    self.w_bf16.copy_(self.w)            # dtype conversion of the operand copy (weights are bf16-representable)

  def export_state(self):
    if self.kind == "logistic":
      return {"fully_connected/weights": self.w[:, :self.d].t().contiguous().cpu(), "fully_connected/biases": self.b.cpu().clone()}
    g_idx, e_idx = _moe_row_index(self.v, self.m)
    w = self.w.cpu()
    return {"gates/weights": w[g_idx, :self.d].t().contiguous(), "experts/weights": w[e_idx, :self.d].t().contiguous(),
            "experts/biases": self.b.cpu()[e_idx].clone()}

  # ---- forward / step -------------------------------------------------------------------------------
  def _operand(self, x):
    """x: [B, D] bf16 (exact) or fp32 (split hi/lo) on the GPU -> (hi, lo)."""
    if x.dtype == torch.bfloat16:
      if x.stride(0) % 8 or x.stride(1) != 1:
        buf = torch.zeros((x.shape[0], self.dpad), dtype=torch.bfloat16, device=x.device)
        buf[:, :self.d] = x
        x = buf[:, :self.d]
      return x, None
    return nat.split_bf16(x.float().contiguous())

  def forward(self, x):
    hi, lo = self._operand(x)
    if self.kind == "logistic":
      return nat.linear(hi, self.w_bf16, a_lo=lo, n=self.v, k=self.d, shift=self.b, act="sigmoid")["f32"], (hi, lo)
    return nat.moe_fwd(hi, self.w_bf16, self.b, self.v, self.m, x_lo=lo, d=self.d), (hi, lo)

  def backward(self, p, hi, lo, labels, global_batch, want_dx=False, loss_weight=1.0):
    """Loss + gradients of this rank's shard into self.gw / self.gb (NOT yet all-reduced).  Returns (loss, dx):
    dx [B_local, D] fp32 = dLoss/dx when want_dx (the head sits on top of a trainable frame-level model).
    loss_weight scales the gradient only (--multitask: 1 - support_loss_percent, wh/losses.py:279)."""
    b_local = p.shape[0]
    # d(mean over the GLOBAL batch)/dp: xent divides by the local batch, so rescale by B_local / B_global
    loss, dp = nat.xent(p, labels, want_grad=True, grad_scale=loss_weight * b_local / float(global_batch))
    return loss, self.backward_from_dp(dp, p, hi, lo, want_dx)

  def backward_from_dp(self, dp, p, hi, lo, want_dx=False):
    """Gradients of this head given dL/dp [rows, V] (the loss sits further up: e.g. after the max over attention heads).
    Returns dx [rows, D] fp32 when want_dx."""
    if self.kind == "logistic":
      dz_hi, dz_lo = nat.logistic_bwd_dz(dp, p)
    else:
      dz_hi, dz_lo = nat.moe_bwd_dlogits(hi, lo, self.w_bf16, self.b, dp, self.v, self.m, d=self.d)
    nat.wgrad(dz_hi, dz_lo, hi, self.rows, self.d, out=self.gw)          # dW^T[rows, D] = dZ^T . X  (lo of x is dropped)
    nat.colsum_bf16(dz_hi, dz_lo, self.rows, out=self.gb)
    dx = None
    if want_dx:
      # dX[B, D] = dZ[B, rows] . W[rows, D]: contraction over the rows of the forward's own bf16 operand copy
      dx = nat.dgrad(dz_hi, dz_lo, self.w_bf16, self.d)
    return dx

  def apply(self, lr_t, clip_gradient_norm=1.0, regularization_penalty=1.0):
    """L2 regulariser gradient, per-tensor clip, TF-Adam on this head's tensors (self.grad already all-reduced)."""
    sums_w = nat.grad_reg_sumsq(self.gw, self.w, self.l2 * regularization_penalty, self.per, self.m)
    sums_b = nat.grad_reg_sumsq(self.gb.view(self.rows, 1), self.b.view(self.rows, 1), 0.0, self.per, self.m)
    nat.clip_adam_step(self.w, self.gw, self.mw, self.vw, sums_w, clip_gradient_norm, lr_t, moe_per=self.per, moe_nmix=self.m,
                       param_bf16=self.w_bf16)
    nat.clip_adam_step(self.b.view(self.rows, 1), self.gb.view(self.rows, 1), self.mb.view(self.rows, 1), self.vb.view(self.rows, 1),
                       sums_b, clip_gradient_norm, lr_t, moe_per=self.per, moe_nmix=self.m, only_segment=1 if self.kind == "moe" else -1)
    return sums_w

  def step(self, x, labels, base_lr=0.01, lr_decay=0.95, lr_decay_examples=4000000, clip_gradient_norm=1.0,
           regularization_penalty=1.0, global_batch=None):
    """One optimiser step on this rank's shard (x [B_local, D], labels [B_local, V]).  Returns predictions;
    loss terms are left on the device in ``self.last`` (fetch with .item() only when logging)."""
    b_local = x.shape[0]
    global_batch = global_batch or b_local * self.world
    p, (hi, lo) = self.forward(x)
    loss, _ = self.backward(p, hi, lo, labels, global_batch)
    yt8m_dp.all_reduce_sum_(self.grad, self.group)                        # the ONE collective of the step
    if self.keep_grads:
      self.last_grad = self.grad.clone()                                  # tests: gradient parity before reg/clip
    lr = exponential_decay(base_lr, self.global_step, global_batch, lr_decay_examples, lr_decay)
    lr_t = adam_lr_t(lr, self.global_step + 1)
    sums_w = self.apply(lr_t, clip_gradient_norm, regularization_penalty)
    self.global_step += 1
    self.last = {"label_loss_local": loss, "sums_w": sums_w, "lr": lr}
    return p

  def grads_tf_layout(self, flat):
    """A flat packed gradient / parameter buffer -> reference names/layouts (tests, checkpoints)."""
    n_w = self.rows * self.dpad
    w, b = flat[:n_w].view(self.rows, self.dpad).cpu(), flat[n_w:].cpu()
    if self.kind == "logistic":
      return {"fully_connected/weights": w[:, :self.d].t().contiguous(), "fully_connected/biases": b.clone()}
    g_idx, e_idx = _moe_row_index(self.v, self.m)
    return {"gates/weights": w[g_idx, :self.d].t().contiguous(), "experts/weights": w[e_idx, :self.d].t().contiguous(),
            "experts/biases": b[e_idx].clone()}

  def reg_loss(self):
    """sum over weight tensors of l2 * ||W||^2 / 2 (slim.l2_regularizer) at the start of the last step."""
    s = self.last["sums_w"]
    return self.l2 * float(s[2] + s[3]) / 2.0


class NetVLADTrainer(object):
  """The training step for NetVLADModel + MoeModel (BASELINE config 2) on the GPU through the C ABI: fused NetVLAD
  forward, hidden FC, MoE head, CrossEntropyLoss, and the full backward -- MoE head (fused recompute epilogue +
  MN-major wgrad + dgrad), hidden FC (activation backward, wgrad, dgrad), NetVLAD layer (normalisation / residual /
  assignment-softmax backward kernels + the cluster-weight wgrad over all B*T frame rows) -- then per-tensor
  clip_by_norm and TF-Adam (wh/train.py:440-466).  NetVLAD itself is not part of the reference (oracle/
  yt8m_oracle.py:netvlad_pool); the bias variant (--netvlad_add_batch_norm=False) is the one trained here.

  One flat fp32 buffer holds every gradient; data parallelism sums it over the ranks once per step, in three contiguous
  pieces started as their gradients become final (yt8m_dp.GradExchange).
  Layouts: cluster_weights^T [K, D], cluster_biases [K], cluster_weights2 [D, K], hidden1_weights^T [H, K*D],
  hidden1_biases [H], then the packed MoE head.  import_state / export_state speak the model's TF names."""

  def __init__(self, feature_dim, clusters=64, hidden=1024, vocab=4716, mixtures=2, relu=True, l2_penalty=1e-8, device=None,
               group=None, gating=False):
    """gating=True: GatedNetVLADModel -- context gating y = h * sigmoid(h . Wg + bg) between the hidden layer and the
    classifier (BASELINE config 4; bias variant)."""
    self.d, self.k, self.h, self.v, self.m, self.relu = feature_dim, clusters, hidden, vocab, mixtures, relu
    self.gating = gating
    self.dev = device or torch.device("cuda", torch.cuda.current_device())
    self.group = group
    self.world = yt8m_dp.world_size(group)
    kd = clusters * feature_dim
    self.kd = kd
    sizes = [("cw", clusters * feature_dim), ("cb", clusters), ("c2", feature_dim * clusters), ("wfc", hidden * kd), ("bfc", hidden)]
    if gating:
      sizes += [("wg", hidden * hidden), ("bg", hidden)]
    sizes.append(("head", HeadTrainer.flat_size("moe", hidden, vocab, mixtures)))
    total = sum(n for _, n in sizes)
    self.param = torch.zeros(total, dtype=torch.float32, device=self.dev)
    self.grad = torch.zeros_like(self.param)
    self.adam_m = torch.zeros_like(self.param)
    self.adam_v = torch.zeros_like(self.param)
    self._off = {}
    off = 0
    for name, n in sizes:
      self._off[name] = (off, off + n)
      off += n
    shapes = {"cw": (clusters, feature_dim), "cb": (clusters, 1), "c2": (feature_dim, clusters), "wfc": (hidden, kd), "bfc": (hidden, 1)}
    if gating:
      shapes.update({"wg": (hidden, hidden), "bg": (hidden, 1)})            # Wg^T [out, in]
    self._names = list(shapes)
    self.p, self.g, self.am, self.av = {}, {}, {}, {}
    for name, shp in shapes.items():
      a, b = self._off[name]
      self.p[name], self.g[name] = self.param[a:b].view(shp), self.grad[a:b].view(shp)
      self.am[name], self.av[name] = self.adam_m[a:b].view(shp), self.adam_v[a:b].view(shp)
    a, b = self._off["head"]
    self.head = HeadTrainer("moe", hidden, vocab, mixtures, l2_penalty, self.dev, group,
                            storage=(self.param[a:b], self.grad[a:b], self.adam_m[a:b], self.adam_v[a:b]))
    self.cw_bf16 = torch.zeros((clusters, feature_dim), dtype=torch.bfloat16, device=self.dev)
    self.wfc_bf16 = torch.zeros((hidden, kd), dtype=torch.bfloat16, device=self.dev)
    self.wg_bf16 = torch.zeros((hidden, hidden), dtype=torch.bfloat16, device=self.dev) if gating else None
    self.global_step = 0
    self.keep_grads = False
    self.last = {}

  # ---- TF names / layouts ----------------------------------------------------------------------------
  def import_state(self, sd):
    dev = self.dev
    self.p["cw"].copy_(sd["cluster_weights"].t().to(dev))
    self.p["cb"].copy_(sd["cluster_biases"].view(-1, 1).to(dev))
    self.p["c2"].copy_(sd["cluster_weights2"].to(dev))
    self.p["wfc"].copy_(sd["hidden1_weights"].t().to(dev))
    self.p["bfc"].copy_(sd["hidden1_biases"].view(-1, 1).to(dev))
    self.head.import_state({k: sd[k] for k in ("gates/weights", "experts/weights", "experts/biases")})
    self.cw_bf16.copy_(self.p["cw"])
    self.wfc_bf16.copy_(self.p["wfc"])
    if self.gating:
      self.p["wg"].copy_(sd["gating_weights"].t().to(dev))
      self.p["bg"].copy_(sd["gating_biases"].view(-1, 1).to(dev))
      self.wg_bf16.copy_(self.p["wg"])

  def _tf_layout(self, views, head_flat):
    out = {"cluster_weights": views["cw"].t().contiguous().cpu(), "cluster_biases": views["cb"].reshape(-1).cpu().clone(),
           "cluster_weights2": views["c2"].cpu().clone(), "hidden1_weights": views["wfc"].t().contiguous().cpu(),
           "hidden1_biases": views["bfc"].reshape(-1).cpu().clone()}
    if self.gating:
      out["gating_weights"] = views["wg"].t().contiguous().cpu()
      out["gating_biases"] = views["bg"].reshape(-1).cpu().clone()
    out.update(self.head.grads_tf_layout(head_flat))
    return out

  def export_state(self):
    a, b = self._off["head"]
    return self._tf_layout(self.p, self.param[a:b])

  def grads_tf_layout(self, flat):
    views = {}
    for name in self._names:
      a, b = self._off[name]
      views[name] = flat[a:b].view(self.p[name].shape)
    a, b = self._off["head"]
    return self._tf_layout(views, flat[a:b])

  # ---- forward / step --------------------------------------------------------------------------------
  def forward(self, x, num_frames):
    """x bf16 [B, T, D] (L2-normalised frame rows), num_frames int32 [B] -> predictions + what the backward needs."""
    c2 = self.p["c2"]
    cb = self.p["cb"].view(-1)
    vh, vl, y32, stats = nat.netvlad_fwd(x, num_frames, self.cw_bf16, None, cb, c2, want_f32=True, want_lo=True, want_stats=True)
    hid = nat.linear(vh, self.wfc_bf16, a_lo=vl, n=self.h, k=self.kd, shift=self.p["bfc"].view(-1),
                     act="relu6" if self.relu else None, out_f32=True, out_bf16=True, out_lo=True)
    sv = {"vh": vh, "vl": vl, "y32": y32, "stats": stats, "hid": hid}
    top_hi, top_lo = hid["hi"], hid["lo"]
    if self.gating:
      g = nat.linear(hid["hi"], self.wg_bf16, a_lo=hid["lo"], n=self.h, k=self.h)["f32"]
      _, top_hi, top_lo = nat.context_gate(hid["f32"], g, None, self.p["bg"].view(-1))
      sv["g"] = g
    sv["top"] = (top_hi, top_lo)
    p = nat.moe_fwd(top_hi, self.head.w_bf16, self.head.b, self.v, self.m, x_lo=top_lo, d=self.h)
    return p, sv

  def step(self, x, num_frames, labels, base_lr=0.01, lr_decay=0.95, lr_decay_examples=4000000, clip_gradient_norm=1.0,
           regularization_penalty=1.0, global_batch=None):
    b, t, d = x.shape
    global_batch = global_batch or b * self.world
    p, sv = self.forward(x, num_frames)
    hid = sv["hid"]
    loss, dhid = self.head.backward(p, sv["top"][0], sv["top"][1], labels, global_batch, want_dx=True)
    # the gradient exchange: the flat buffer [cw cb c2 | wfc bfc (wg bg) | head] leaves in three contiguous pieces, each as
    # soon as it is final -- the classifier's 97 MB and the hidden layer's 302 MB travel while the rest of the backward runs
    xch = yt8m_dp.GradExchange(self.group, self.world, getattr(self, "wire_dtype", None))
    xch.start(self.grad[self._off["head"][0]:], "head")
    if self.gating:
      # context gating y = h * sigmoid(h . Wg + bg): direct path + the path through the gate logits
      dhid, _, dg_hi, dg_lo = nat.context_gate_bwd(dhid[:, :self.h].contiguous(), hid["f32"], sv["g"], None, self.p["bg"].view(-1))
      nat.wgrad(dg_hi, dg_lo, hid["hi"], self.h, self.h, out=self.g["wg"])             # dWg^T [out, in]
      nat.colsum_bf16(dg_hi, dg_lo, self.h, out=self.g["bg"].view(-1))
      nat.add_inplace(dhid, nat.dgrad(dg_hi, dg_lo, self.wg_bf16, self.h))              # through the forward's bf16 Wg^T [out, in]
    # hidden FC: h = act(vlad . Wfc + b)
    dpre_hi, dpre_lo = nat.act_bwd(dhid, hid["f32"], act="relu6" if self.relu else None)
    nat.wgrad(dpre_hi, dpre_lo, sv["vh"], self.h, self.kd, out=self.g["wfc"])          # dWfc^T [H, K*D]
    nat.colsum_bf16(dpre_hi, dpre_lo, self.h, out=self.g["bfc"].view(-1))
    xch.start(self.grad[self._off["wfc"][0]:self._off["head"][0]], "fc")
    dvlad = nat.dgrad(dpre_hi, dpre_lo, self.wfc_bf16, self.kd)                         # [B, K*D] through the forward's bf16 Wfc^T [H, K*D]
    # NetVLAD layer
    if nat.netvlad_bwd_assign_fused_supported(t, d, self.k):
      # one tcgen05 kernel: logits recomputed on chip, da = X . dV on the tensor cores, softmax backward in its epilogue
      dv, dasum, dc2, dv_split = nat.netvlad_bwd_norm(dvlad, sv["y32"], sv["stats"], self.p["c2"], want_split=True)
      dz_hi, dz_lo, dshift = nat.netvlad_bwd_assign_fused(x, num_frames, self.cw_bf16, None, self.p["cb"].view(-1), dv_split, dasum)
    else:
      dv, dasum, dc2 = nat.netvlad_bwd_norm(dvlad, sv["y32"], sv["stats"], self.p["c2"])
      z = nat.linear(x.reshape(b * t, d), self.cw_bf16, n=self.k, k=d, shift=self.p["cb"].view(-1))["f32"]
      dz_hi, dz_lo, dshift = nat.netvlad_bwd_assign(x, num_frames, z, dv, dasum)
    self.g["c2"].copy_(dc2)
    self.g["cb"].view(-1).copy_(dshift)
    nat.wgrad(dz_hi, dz_lo, x.reshape(b * t, d), self.k, d, out=self.g["cw"])           # dCw^T [K, D]
    xch.start(self.grad[:self._off["wfc"][0]], "pool")
    if self.keep_grads:
      xch.finish()
      self.last_grad = self.grad.clone()
    lr = exponential_decay(base_lr, self.global_step, global_batch, lr_decay_examples, lr_decay)
    lr_t = adam_lr_t(lr, self.global_step + 1)

    def update(tensors):
      for name, bf in tensors:
        sums = nat.grad_reg_sumsq(self.g[name], self.p[name], 0.0)
        nat.clip_adam_step(self.p[name], self.g[name], self.am[name], self.av[name], sums, clip_gradient_norm, lr_t, param_bf16=bf)

    # each piece is updated as it arrives: the classifier's Adam runs while the hidden layer's 302 MB are still on the wire
    xch.wait("head")
    self.head.apply(lr_t, clip_gradient_norm, regularization_penalty)
    xch.wait("fc")
    update([("wfc", self.wfc_bf16), ("bfc", None)] + ([("wg", self.wg_bf16), ("bg", None)] if self.gating else []))
    xch.wait("pool")
    update([("cw", self.cw_bf16), ("cb", None), ("c2", None)])
    self.global_step += 1
    self.head.global_step = self.global_step
    self.last = {"label_loss_local": loss, "lr": lr}
    return p


def _lstm_tf_to_packed(w_tf, hidden):
  """TF BasicLSTMCell kernel [in+H, 4H] (columns g*H + u, gates i, j, f, o) -> packed [4H, in+H] (rows 4u + g)."""
  k = w_tf.shape[0]
  return w_tf.t().reshape(4, hidden, k).permute(1, 0, 2).reshape(4 * hidden, k).contiguous()


def _lstm_packed_to_tf(w_packed, hidden):
  k = w_packed.shape[1]
  return w_packed.reshape(hidden, 4, k).permute(1, 0, 2).reshape(4 * hidden, k).t().contiguous()


class LstmTrainer(object):
  """The training step for LstmModel / LstmMemoryModel + MoeModel (BASELINE config 3) on the GPU through the C ABI:
  persistent-recurrence forward that retains every layer's output sequence (yt8m_lstm_fwd_train), MoE head,
  CrossEntropyLoss, MoE backward, back-propagation through time (yt8m_lstm_bwd), per-tensor clip_by_norm and TF-Adam
  (wh/train.py:440-466 over wh/all_frame_models/lstm_model.py:30-57 / lstm_memory_model.py:47-73).

  memory=False: classifier input = the non-tuple state [c0, h0, c1, h1, ...] (lstm_model.py:52);
  memory=True : classifier input = concat of the layers' c states (lstm_memory_model.py:61).
  One flat fp32 buffer holds every gradient; under data parallelism it is summed over the ranks once per step, the
  classifier's piece while the back-propagation through time is still running.
  Layouts: per layer the packed kernel [4H, in+H] (rows 4u+g) and bias [4H]; then the packed MoE head.
  import_state / export_state speak the reference's TF variable names and layouts."""

  SCOPE = "RNN/multi_rnn_cell/cell_%d/basic_lstm_cell"

  def __init__(self, feature_dim, hidden=1024, layers=2, vocab=4716, mixtures=2, memory=False, l2_penalty=1e-8, device=None,
               group=None):
    self.d, self.h, self.l, self.v, self.m, self.memory = feature_dim, hidden, layers, vocab, mixtures, memory
    self.dev = device or torch.device("cuda", torch.cuda.current_device())
    self.group = group
    self.world = yt8m_dp.world_size(group)
    self.head_in = layers * hidden if memory else layers * 2 * hidden
    sizes = []
    for l in range(layers):
      k = (feature_dim if l == 0 else hidden) + hidden
      sizes += [("w%d" % l, 4 * hidden * k), ("b%d" % l, 4 * hidden)]
    sizes.append(("head", HeadTrainer.flat_size("moe", self.head_in, vocab, mixtures)))
    total = sum(n for _, n in sizes)
    self.param = torch.zeros(total, dtype=torch.float32, device=self.dev)
    self.grad = torch.zeros_like(self.param)
    self.adam_m = torch.zeros_like(self.param)
    self.adam_v = torch.zeros_like(self.param)
    self._off, off = {}, 0
    for name, n in sizes:
      self._off[name] = (off, off + n)
      off += n
    self.p, self.g, self.am, self.av = {}, {}, {}, {}
    for l in range(layers):
      k = (feature_dim if l == 0 else hidden) + hidden
      for name, shp in (("w%d" % l, (4 * hidden, k)), ("b%d" % l, (4 * hidden, 1))):
        a, b = self._off[name]
        self.p[name], self.g[name] = self.param[a:b].view(shp), self.grad[a:b].view(shp)
        self.am[name], self.av[name] = self.adam_m[a:b].view(shp), self.adam_v[a:b].view(shp)
    a, b = self._off["head"]
    self.head = HeadTrainer("moe", self.head_in, vocab, mixtures, l2_penalty, self.dev, group,
                            storage=(self.param[a:b], self.grad[a:b], self.adam_m[a:b], self.adam_v[a:b]))
    self.w_bf16 = [torch.zeros(self.p["w%d" % l].shape, dtype=torch.bfloat16, device=self.dev) for l in range(layers)]
    if memory:
      # columns of the full state [c0, h0, c1, h1, ...] that feed the classifier
      self.state_cols = torch.cat([torch.arange(l * 2 * hidden, l * 2 * hidden + hidden) for l in range(layers)]).to(self.dev)
    self.global_step = 0
    self.keep_grads = False
    self.last = {}

  # ---- TF names / layouts ----------------------------------------------------------------------------
  def import_state(self, sd):
    for l in range(self.l):
      scope = self.SCOPE % l
      self.p["w%d" % l].copy_(_lstm_tf_to_packed(sd[scope + "/weights"].to(self.dev), self.h))
      b = sd[scope + "/biases"].to(self.dev)
      self.p["b%d" % l].copy_(b.reshape(4, self.h).t().reshape(-1, 1))
      self.w_bf16[l].copy_(self.p["w%d" % l])
    self.head.import_state({k: sd[k] for k in ("gates/weights", "experts/weights", "experts/biases")})

  def _tf_layout(self, flat):
    out = {}
    for l in range(self.l):
      scope = self.SCOPE % l
      a, b = self._off["w%d" % l]
      out[scope + "/weights"] = _lstm_packed_to_tf(flat[a:b].view(self.p["w%d" % l].shape), self.h).cpu()
      a, b = self._off["b%d" % l]
      out[scope + "/biases"] = flat[a:b].view(self.h, 4).t().reshape(-1).cpu().clone()
    a, b = self._off["head"]
    out.update(self.head.grads_tf_layout(flat[a:b]))
    return out

  def export_state(self):
    return self._tf_layout(self.param)

  def grads_tf_layout(self, flat):
    return self._tf_layout(flat)

  # ---- forward / step --------------------------------------------------------------------------------
  def forward(self, x, num_frames):
    """x bf16 [B, T, D] (L2-normalised frame rows), num_frames int32 [B] -> predictions + what the backward needs."""
    bs = [self.p["b%d" % l].view(-1) for l in range(self.l)]
    state, _, seq_hi, seq_lo = nat.lstm_fwd_train(x, num_frames, self.w_bf16, bs, self.h)
    feat = state.index_select(1, self.state_cols) if self.memory else state
    p, (hi, lo) = self.head.forward(feat)
    return p, {"seq_hi": seq_hi, "seq_lo": seq_lo, "hi": hi, "lo": lo, "bs": bs}

  def step(self, x, num_frames, labels, base_lr=0.01, lr_decay=0.95, lr_decay_examples=4000000, clip_gradient_norm=1.0,
           regularization_penalty=1.0, global_batch=None):
    b = x.shape[0]
    global_batch = global_batch or b * self.world
    p, sv = self.forward(x, num_frames)
    loss, dfeat = self.head.backward(p, sv["hi"], sv["lo"], labels, global_batch, want_dx=True)
    # the classifier's gradient (174 M floats with MoE-4 on the 4096-d state: 0.7 GB) is final here and travels under the whole
    # back-propagation through time; the recurrent layers' piece follows at the end (yt8m_dp.GradExchange)
    xch = yt8m_dp.GradExchange(self.group, self.world)
    xch.start(self.grad[self._off["head"][0]:], "head")
    if self.memory:
      dstate = torch.zeros((b, self.l * 2 * self.h), dtype=torch.float32, device=self.dev)
      dstate.index_copy_(1, self.state_cols, dfeat[:, :self.head_in].contiguous())
    else:
      dstate = dfeat[:, :self.head_in].contiguous()
    wt = [nat.pack_transpose(self.p["w%d" % l]) for l in range(self.l)]          # bf16 [in+H, 4H]: the dgrad operands
    nat.lstm_bwd(x, num_frames, self.w_bf16, sv["bs"], wt, self.h, sv["seq_hi"], sv["seq_lo"], dstate=dstate,
                 dw=[self.g["w%d" % l] for l in range(self.l)], db=[self.g["b%d" % l].view(-1) for l in range(self.l)])
    del wt
    xch.start(self.grad[:self._off["head"][0]], "rnn")
    if self.keep_grads:
      xch.finish()
      self.last_grad = self.grad.clone()
    lr = exponential_decay(base_lr, self.global_step, global_batch, lr_decay_examples, lr_decay)
    lr_t = adam_lr_t(lr, self.global_step + 1)
    xch.wait("head")
    self.head.apply(lr_t, clip_gradient_norm, regularization_penalty)
    xch.wait("rnn")
    for l in range(self.l):
      for name, bf in (("w%d" % l, self.w_bf16[l]), ("b%d" % l, None)):
        sums = nat.grad_reg_sumsq(self.g[name], self.p[name], 0.0)               # BasicLSTMCell carries no regulariser
        nat.clip_adam_step(self.p[name], self.g[name], self.am[name], self.av[name], sums, clip_gradient_norm, lr_t, param_bf16=bf)
    self.global_step += 1
    self.head.global_step = self.global_step
    self.last = {"label_loss_local": loss, "lr": lr}
    return p


class AttentionTrainer(object):
  """The training step for zt's LSTM-free AttentionModel + MoeExtendModel (multi-head attention pooling over the raw
  frames, MoE on the B*A pooled rows, max over the A heads; zt/frame_level_models.py:4355-4405 +
  zt/video_level_models.py:2272-2330) on the GPU through the C ABI: logits GEMM over all B*T frame rows, attention
  pooling, MoE head, max over heads, CrossEntropyLoss, and the backward of each (group-max routing, fused MoE backward,
  yt8m_attn_pool_bwd, the logits weight gradient as one MN-major GEMM over B*T rows), per-tensor clip and TF-Adam.

  softmax over T is shift invariant, so the mean-pooled half of Attention/W ([D:2D]) and Attention/b receive no label
  gradient -- only their L2 regulariser (l2 * tf.nn.l2_loss, :4389-4392) moves them, exactly as in the reference graph.
  Layouts: Attention/W^T [A, 2D], Attention/b [A], then the packed MoE head; one flat gradient buffer, one all-reduce."""

  def __init__(self, feature_dim, heads=8, vocab=4716, mixtures=2, l2_penalty=1e-8, device=None, group=None):
    self.d, self.a, self.v, self.m, self.l2 = feature_dim, heads, vocab, mixtures, l2_penalty
    self.dev = device or torch.device("cuda", torch.cuda.current_device())
    self.group = group
    self.world = yt8m_dp.world_size(group)
    sizes = [("aw", heads * 2 * feature_dim), ("ab", heads), ("head", HeadTrainer.flat_size("moe", feature_dim, vocab, mixtures))]
    total = sum(n for _, n in sizes)
    self.param = torch.zeros(total, dtype=torch.float32, device=self.dev)
    self.grad = torch.zeros_like(self.param)
    self.adam_m = torch.zeros_like(self.param)
    self.adam_v = torch.zeros_like(self.param)
    self._off, off = {}, 0
    for name, n in sizes:
      self._off[name] = (off, off + n)
      off += n
    self.p, self.g, self.am, self.av = {}, {}, {}, {}
    for name, shp in (("aw", (heads, 2 * feature_dim)), ("ab", (heads, 1))):
      a, b = self._off[name]
      self.p[name], self.g[name] = self.param[a:b].view(shp), self.grad[a:b].view(shp)
      self.am[name], self.av[name] = self.adam_m[a:b].view(shp), self.adam_v[a:b].view(shp)
    a, b = self._off["head"]
    self.head = HeadTrainer("moe", feature_dim, vocab, mixtures, l2_penalty, self.dev, group,
                            storage=(self.param[a:b], self.grad[a:b], self.adam_m[a:b], self.adam_v[a:b]))
    self.aw_bf16 = torch.zeros((heads, 2 * feature_dim), dtype=torch.bfloat16, device=self.dev)
    self.global_step = 0
    self.keep_grads = False
    self.last = {}

  # ---- TF names / layouts ----------------------------------------------------------------------------
  def import_state(self, sd):
    self.p["aw"].copy_(sd["Attention/W"].t().to(self.dev))
    self.p["ab"].copy_(sd["Attention/b"].view(-1, 1).to(self.dev))
    self.aw_bf16.copy_(self.p["aw"])
    self.head.import_state({k: sd[k] for k in ("gates/weights", "experts/weights", "experts/biases")})

  def _tf_layout(self, flat):
    a, b = self._off["aw"]
    out = {"Attention/W": flat[a:b].view(self.a, 2 * self.d).t().contiguous().cpu()}
    a, b = self._off["ab"]
    out["Attention/b"] = flat[a:b].cpu().clone()
    a, b = self._off["head"]
    out.update(self.head.grads_tf_layout(flat[a:b]))
    return out

  def export_state(self):
    return self._tf_layout(self.param)

  def grads_tf_layout(self, flat):
    return self._tf_layout(flat)

  # ---- forward / step --------------------------------------------------------------------------------
  def forward(self, x, num_frames=None):
    """x bf16 [B, T, D] (padded frames are all-zero rows: the mask of :4372-4375).  Returns predictions [B, V]."""
    b, t, d = x.shape
    logits = nat.linear(x.reshape(b * t, d), self.aw_bf16, n=self.a, k=d)["f32"]                # W[:D] only (shift invariance)
    logits3 = logits.as_strided((b, t, self.a), (t * logits.stride(0), logits.stride(0), 1))
    _, hi, lo = nat.attn_pool(logits3, x, None, self.a, 0)
    hi, lo = hi.reshape(b * self.a, d), lo.reshape(b * self.a, d)
    p_heads = nat.moe_fwd(hi, self.head.w_bf16, self.head.b, self.v, self.m, x_lo=lo, d=d)      # [B*A, V]
    return nat.group_max_rows(p_heads, self.a), {"logits3": logits3, "hi": hi, "lo": lo, "p_heads": p_heads}

  def step(self, x, num_frames, labels, base_lr=0.01, lr_decay=0.95, lr_decay_examples=4000000, clip_gradient_norm=1.0,
           regularization_penalty=1.0, global_batch=None):
    b, t, d = x.shape
    global_batch = global_batch or b * self.world
    p, sv = self.forward(x)
    loss, dp = nat.xent(p, labels, want_grad=True, grad_scale=b / float(global_batch))
    dp_heads = nat.group_max_rows_bwd(sv["p_heads"], dp, self.a)                                # gradient to the arg-max head
    dpooled = self.head.backward_from_dp(dp_heads, sv["p_heads"], sv["hi"], sv["lo"], want_dx=True)
    dlogits, _ = nat.attn_pool_bwd(sv["logits3"], x, None, self.a, 0, dpooled[:, :d].contiguous().view(b, self.a, d))
    dl_hi, dl_lo = nat.split_bf16(dlogits.view(b * t, self.a))
    self.g["aw"].zero_()                                                                        # the mean-pooled half: no label gradient
    nat.wgrad(dl_hi, dl_lo, x.reshape(b * t, d), self.a, d, out=self.g["aw"])                   # dW[:D]^T [A, D] into the [A, 2D] rows
    self.g["ab"].zero_()
    yt8m_dp.all_reduce_sum_(self.grad, self.group)                                              # the ONE collective of the step
    if self.keep_grads:
      self.last_grad = self.grad.clone()
    lr = exponential_decay(base_lr, self.global_step, global_batch, lr_decay_examples, lr_decay)
    lr_t = adam_lr_t(lr, self.global_step + 1)
    for name, bf in (("aw", self.aw_bf16), ("ab", None)):
      sums = nat.grad_reg_sumsq(self.g[name], self.p[name], self.l2 * regularization_penalty)
      nat.clip_adam_step(self.p[name], self.g[name], self.am[name], self.av[name], sums, clip_gradient_norm, lr_t, param_bf16=bf)
    self.head.apply(lr_t, clip_gradient_norm, regularization_penalty)
    self.global_step += 1
    self.head.global_step = self.global_step
    self.last = {"label_loss_local": loss, "lr": lr}
    return p


class ChainMoeTrainer(object):
  """The training step for ChainMoeModel (wh/all_video_models/chain_moe_model.py:9-49) without --multitask: a support
  MoE over `num_supports` categories, its predictions concatenated to the input, the main MoE on the concatenation;
  the label loss sits on the main predictions and reaches the support head through the concatenated columns
  (wh/train.py:412-413).  Two packed MoE heads in ONE flat gradient buffer: one all-reduce per step."""

  def __init__(self, in_dim, vocab=4716, mixtures=2, num_supports=25, l2_penalty=1e-8, device=None, group=None):
    self.d, self.v, self.m, self.s = in_dim, vocab, mixtures, num_supports
    self.dev = device or torch.device("cuda", torch.cuda.current_device())
    self.group = group
    self.world = yt8m_dp.world_size(group)
    n_sup = HeadTrainer.flat_size("moe", in_dim, num_supports, mixtures)
    n_main = HeadTrainer.flat_size("moe", in_dim + num_supports, vocab, mixtures)
    self.param = torch.zeros(n_sup + n_main, dtype=torch.float32, device=self.dev)
    self.grad = torch.zeros_like(self.param)
    self.adam_m = torch.zeros_like(self.param)
    self.adam_v = torch.zeros_like(self.param)
    sl = lambda t, a, b: t[a:b]
    self.support = HeadTrainer("moe", in_dim, num_supports, mixtures, l2_penalty, self.dev, group,
                               storage=tuple(sl(t, 0, n_sup) for t in (self.param, self.grad, self.adam_m, self.adam_v)))
    self.main = HeadTrainer("moe", in_dim + num_supports, vocab, mixtures, l2_penalty, self.dev, group,
                            storage=tuple(sl(t, n_sup, n_sup + n_main) for t in (self.param, self.grad, self.adam_m, self.adam_v)))
    self._n_sup = n_sup
    self.global_step = 0
    self.keep_grads = False
    self.last = {}

  @staticmethod
  def _sub(sd, scope):
    return {"gates/weights": sd["gates%s/weights" % scope], "experts/weights": sd["experts%s/weights" % scope],
            "experts/biases": sd["experts%s/biases" % scope]}

  def import_state(self, sd):
    self.support.import_state(self._sub(sd, "-support"))
    self.main.import_state(self._sub(sd, "-main"))

  def _tf_layout(self, flat):
    out = {}
    for scope, head, part in (("-support", self.support, flat[:self._n_sup]), ("-main", self.main, flat[self._n_sup:])):
      for k, v in head.grads_tf_layout(part).items():
        name, leaf = k.split("/")
        out["%s%s/%s" % (name, scope, leaf)] = v
    return out

  def export_state(self):
    return self._tf_layout(self.param)

  def grads_tf_layout(self, flat):
    return self._tf_layout(flat)

  def step(self, x, labels, base_lr=0.01, lr_decay=0.95, lr_decay_examples=4000000, clip_gradient_norm=1.0,
           regularization_penalty=1.0, global_batch=None, support_labels=None, support_loss_percent=0.1):
    """support_labels [B, num_supports] (--multitask, wh/train.py:394-413 + wh/losses.py:271-279): the label loss becomes
    CE(main) * (1 - support_loss_percent) + CE(support predictions, support labels) * support_loss_percent."""
    b = x.shape[0]
    global_batch = global_batch or b * self.world
    pct = support_loss_percent if support_labels is not None else 0.0
    sp, (s_hi, s_lo) = self.support.forward(x)                                        # [B, S] support predictions
    main_in = torch.cat([x.float()[:, :self.d], sp[:, :self.s]], dim=1)               # tf.concat (device copy)
    p, (m_hi, m_lo) = self.main.forward(main_in)
    loss, d_in = self.main.backward(p, m_hi, m_lo, labels, global_batch, want_dx=True, loss_weight=1.0 - pct)
    d_sp = d_in[:, self.d:self.d + self.s].contiguous()                               # the columns that came from the support head
    loss_parts = {"main": loss}
    if support_labels is not None:
      sup_loss, d_sup = nat.xent(sp, support_labels, want_grad=True, grad_scale=pct * b / float(global_batch))
      nat.add_inplace(d_sp, d_sup)
      loss_parts["support"] = sup_loss
    self.support.backward_from_dp(d_sp, sp, s_hi, s_lo)
    yt8m_dp.all_reduce_sum_(self.grad, self.group)                                    # the ONE collective of the step
    if self.keep_grads:
      self.last_grad = self.grad.clone()
    lr = exponential_decay(base_lr, self.global_step, global_batch, lr_decay_examples, lr_decay)
    lr_t = adam_lr_t(lr, self.global_step + 1)
    self.support.apply(lr_t, clip_gradient_norm, regularization_penalty)
    self.main.apply(lr_t, clip_gradient_norm, regularization_penalty)
    self.global_step += 1
    self.last = {"label_loss_local": _mix_loss(loss_parts, pct), "lr": lr, "support_predictions": sp}
    return p


class LstmAttentionTrainer(object):
  """The training step for the two LSTM + multi-head attention models of the reference, on the GPU through the C ABI:

  kind="max_pooling": LstmAttentionMaxPoolingModel (wh/all_frame_models/lstm_attention_max_pooling_model.py:29-68) --
      logits = [x_t, h_t] . Wa + ba, softmax over T (masked, renormalised), pooled = sum_t w . h_t, MoE ("sub-moe") on
      the B*A rows, max over heads;
  kind="multi": LstmMultiAttentionModel (wh/all_frame_models/lstm_multi_attention_model.py:30-91) --
      att = sigmoid(h_t . Wa + ba) masked, / (sum + 1e-8), pooled = sum_t att . x_t (the RAW input), MoeModel, max over heads.

  Backward: group-max routing, fused MoE backward, yt8m_attn_pool_bwd (d logits, and d h_t for "max_pooling"), the
  attention weight gradient as one MN-major GEMM over B*T rows, d h_t through the logits GEMM, then back-propagation
  through time with the gradient of the top layer's OUTPUT SEQUENCE (yt8m_lstm_bwd, dout_seq), clip + TF-Adam.
  As in the forward plugins the attention reads the bf16 (hi) copy of the LSTM outputs."""

  SCOPE = LstmTrainer.SCOPE

  def __init__(self, feature_dim, hidden=1024, layers=2, heads=8, vocab=4716, mixtures=2, kind="max_pooling", l2_penalty=1e-8,
               device=None, group=None):
    assert kind in ("max_pooling", "multi")
    self.d, self.h, self.l, self.a, self.v, self.m, self.kind, self.l2 = feature_dim, hidden, layers, heads, vocab, mixtures, kind, l2_penalty
    self.dev = device or torch.device("cuda", torch.cuda.current_device())
    self.group = group
    self.world = yt8m_dp.world_size(group)
    self.att_in = feature_dim + hidden if kind == "max_pooling" else hidden      # rows of the attention matrix
    self.pool_dim = hidden if kind == "max_pooling" else feature_dim             # what the heads pool
    sizes = []
    for l in range(layers):
      k = (feature_dim if l == 0 else hidden) + hidden
      sizes += [("w%d" % l, 4 * hidden * k), ("b%d" % l, 4 * hidden)]
    sizes += [("wa", heads * self.att_in), ("ba", heads), ("head", HeadTrainer.flat_size("moe", self.pool_dim, vocab, mixtures))]
    total = sum(n for _, n in sizes)
    self.param = torch.zeros(total, dtype=torch.float32, device=self.dev)
    self.grad = torch.zeros_like(self.param)
    self.adam_m = torch.zeros_like(self.param)
    self.adam_v = torch.zeros_like(self.param)
    self._off, off = {}, 0
    for name, n in sizes:
      self._off[name] = (off, off + n)
      off += n
    shapes = {"wa": (heads, self.att_in), "ba": (heads, 1)}
    for l in range(layers):
      k = (feature_dim if l == 0 else hidden) + hidden
      shapes["w%d" % l], shapes["b%d" % l] = (4 * hidden, k), (4 * hidden, 1)
    self.p, self.g, self.am, self.av = {}, {}, {}, {}
    for name, shp in shapes.items():
      a, b = self._off[name]
      self.p[name], self.g[name] = self.param[a:b].view(shp), self.grad[a:b].view(shp)
      self.am[name], self.av[name] = self.adam_m[a:b].view(shp), self.adam_v[a:b].view(shp)
    a, b = self._off["head"]
    self.head = HeadTrainer("moe", self.pool_dim, vocab, mixtures, l2_penalty, self.dev, group,
                            storage=(self.param[a:b], self.grad[a:b], self.adam_m[a:b], self.adam_v[a:b]))
    self.w_bf16 = [torch.zeros(self.p["w%d" % l].shape, dtype=torch.bfloat16, device=self.dev) for l in range(layers)]
    self.wa_bf16 = torch.zeros((heads, self.att_in), dtype=torch.bfloat16, device=self.dev)
    self.att_name = "attention-" if kind == "max_pooling" else "fully_connected"
    self.moe_names = ("gates-sub-moe", "experts-sub-moe") if kind == "max_pooling" else ("gates", "experts")
    self.global_step = 0
    self.keep_grads = False
    self.last = {}

  # ---- TF names / layouts ----------------------------------------------------------------------------
  def import_state(self, sd):
    for l in range(self.l):
      scope = self.SCOPE % l
      self.p["w%d" % l].copy_(_lstm_tf_to_packed(sd[scope + "/weights"].to(self.dev), self.h))
      self.p["b%d" % l].copy_(sd[scope + "/biases"].to(self.dev).reshape(4, self.h).t().reshape(-1, 1))
      self.w_bf16[l].copy_(self.p["w%d" % l])
    self.p["wa"].copy_(sd[self.att_name + "/weights"].t().to(self.dev))
    self.p["ba"].copy_(sd[self.att_name + "/biases"].view(-1, 1).to(self.dev))
    self.wa_bf16.copy_(self.p["wa"])
    gn, en = self.moe_names
    self.head.import_state({"gates/weights": sd[gn + "/weights"], "experts/weights": sd[en + "/weights"], "experts/biases": sd[en + "/biases"]})

  def _tf_layout(self, flat):
    out = {}
    for l in range(self.l):
      scope = self.SCOPE % l
      a, b = self._off["w%d" % l]
      out[scope + "/weights"] = _lstm_packed_to_tf(flat[a:b].view(self.p["w%d" % l].shape), self.h).cpu()
      a, b = self._off["b%d" % l]
      out[scope + "/biases"] = flat[a:b].view(self.h, 4).t().reshape(-1).cpu().clone()
    a, b = self._off["wa"]
    out[self.att_name + "/weights"] = flat[a:b].view(self.a, self.att_in).t().contiguous().cpu()
    a, b = self._off["ba"]
    out[self.att_name + "/biases"] = flat[a:b].cpu().clone()
    a, b = self._off["head"]
    gn, en = self.moe_names
    for k, v in self.head.grads_tf_layout(flat[a:b]).items():
      name, leaf = k.split("/")
      out["%s/%s" % (gn if name == "gates" else en, leaf)] = v
    return out

  def export_state(self):
    return self._tf_layout(self.param)

  def grads_tf_layout(self, flat):
    return self._tf_layout(flat)

  # ---- step ------------------------------------------------------------------------------------------
  def step(self, x, num_frames, labels, base_lr=0.01, lr_decay=0.95, lr_decay_examples=4000000, clip_gradient_norm=1.0,
           regularization_penalty=1.0, global_batch=None):
    b, t, d = x.shape
    h, a = self.h, self.a
    global_batch = global_batch or b * self.world
    bs = [self.p["b%d" % l].view(-1) for l in range(self.l)]
    _, _, seq_hi, seq_lo = nat.lstm_fwd_train(x, num_frames, self.w_bf16, bs, h)
    out_bf = seq_hi[-1]                                                            # bf16 outputs of the top layer [B, T, H]
    if self.kind == "max_pooling":
      att_op = torch.cat([x, out_bf], dim=2).reshape(b * t, d + h)                 # tf.concat (device copy)
      feats, mode = out_bf, 0
    else:
      att_op = out_bf.reshape(b * t, h)
      feats, mode = x, 1
    logits = nat.linear(att_op, self.wa_bf16, n=a, k=self.att_in, shift=self.p["ba"].view(-1))["f32"]
    logits3 = logits.as_strided((b, t, a), (t * logits.stride(0), logits.stride(0), 1))
    _, hi, lo = nat.attn_pool(logits3, feats, num_frames, a, mode)
    hi, lo = hi.reshape(b * a, self.pool_dim), lo.reshape(b * a, self.pool_dim)
    p_heads = nat.moe_fwd(hi, self.head.w_bf16, self.head.b, self.v, self.m, x_lo=lo, d=self.pool_dim)
    p = nat.group_max_rows(p_heads, a)
    # ---- backward
    loss, dp = nat.xent(p, labels, want_grad=True, grad_scale=b / float(global_batch))
    dp_heads = nat.group_max_rows_bwd(p_heads, dp, a)
    dpooled = self.head.backward_from_dp(dp_heads, p_heads, hi, lo, want_dx=True)
    dlogits, dfeats = nat.attn_pool_bwd(logits3, feats, num_frames, a, mode,
                                        dpooled[:, :self.pool_dim].contiguous().view(b, a, self.pool_dim),
                                        want_dfeats=(self.kind == "max_pooling"))
    dl_hi, dl_lo = nat.split_bf16(dlogits.view(b * t, a))
    nat.wgrad(dl_hi, dl_lo, att_op, a, self.att_in, out=self.g["wa"])              # dWa^T [A, att_in]
    nat.colsum_bf16(dl_hi, dl_lo, a, out=self.g["ba"].view(-1))
    wa_t = nat.pack_transpose(self.p["wa"])                                        # bf16 [att_in, 8]: rows of Wa, K = A contiguous
    wa_h = wa_t[d:] if self.kind == "max_pooling" else wa_t                        # the rows that multiply h_t
    dseq = nat.linear(dl_hi, wa_h, a_lo=dl_lo, n=h, k=a)["f32"]                    # d h_t through the logits [B*T, H]
    if dfeats is not None:
      nat.add_inplace(dseq, dfeats.view(b * t, h))                                 # + d h_t through the pooling
    wt = [nat.pack_transpose(self.p["w%d" % l]) for l in range(self.l)]
    nat.lstm_bwd(x, num_frames, self.w_bf16, bs, wt, h, seq_hi, seq_lo, dstate=None, dout_seq=dseq.view(b, t, h),
                 dw=[self.g["w%d" % l] for l in range(self.l)], db=[self.g["b%d" % l].view(-1) for l in range(self.l)])
    del wt
    yt8m_dp.all_reduce_sum_(self.grad, self.group)                                 # the ONE collective of the step
    if self.keep_grads:
      self.last_grad = self.grad.clone()
    lr = exponential_decay(base_lr, self.global_step, global_batch, lr_decay_examples, lr_decay)
    lr_t = adam_lr_t(lr, self.global_step + 1)
    for l in range(self.l):
      for name, bf in (("w%d" % l, self.w_bf16[l]), ("b%d" % l, None)):
        sums = nat.grad_reg_sumsq(self.g[name], self.p[name], 0.0)
        nat.clip_adam_step(self.p[name], self.g[name], self.am[name], self.av[name], sums, clip_gradient_norm, lr_t, param_bf16=bf)
    sums = nat.grad_reg_sumsq(self.g["wa"], self.p["wa"], self.l2 * regularization_penalty)   # slim l2_regularizer on the weights
    nat.clip_adam_step(self.p["wa"], self.g["wa"], self.am["wa"], self.av["wa"], sums, clip_gradient_norm, lr_t, param_bf16=self.wa_bf16)
    sums = nat.grad_reg_sumsq(self.g["ba"], self.p["ba"], 0.0)
    nat.clip_adam_step(self.p["ba"], self.g["ba"], self.am["ba"], self.av["ba"], sums, clip_gradient_norm, lr_t)
    self.head.apply(lr_t, clip_gradient_norm, regularization_penalty)
    self.global_step += 1
    self.head.global_step = self.global_step
    self.last = {"label_loss_local": loss, "lr": lr}
    return p


class DeepCombineChainTrainer(object):
  """The training step for DeepCombineChainModel (wh/all_video_models/deep_combine_chain_model.py:9-85) without
  --multitask: `layers` stacked MoE sub-predictions, each projected (V -> relu_cells), ReLU, L2-normalised and
  concatenated to the next MoE's input; the main MoE on the last concatenation carries the label loss.  Backward walks
  the chain in reverse: the gradient of a concatenation splits into the part that continues down the chain and the part
  that goes through L2-normalise (yt8m_l2norm_rows_bwd), ReLU, the projection (wgrad + dgrad) and the layer's MoE."""

  def __init__(self, in_dim, vocab=4716, mixtures=2, layers=3, relu_cells=256, l2_penalty=1e-8, device=None, group=None):
    self.d, self.v, self.m, self.nl, self.r, self.l2 = in_dim, vocab, mixtures, layers, relu_cells, l2_penalty
    self.dev = device or torch.device("cuda", torch.cuda.current_device())
    self.group = group
    self.world = yt8m_dp.world_size(group)
    self.vpad = nat.pad8(vocab)
    dims = [in_dim + i * relu_cells for i in range(layers + 1)]               # input width of layer i (the last: the main head)
    sizes = []
    for i in range(layers):
      sizes += [("moe%d" % i, HeadTrainer.flat_size("moe", dims[i], vocab, mixtures)), ("wr%d" % i, relu_cells * self.vpad),
                ("br%d" % i, relu_cells)]
    sizes.append(("main", HeadTrainer.flat_size("moe", dims[layers], vocab, mixtures)))
    total = sum(n for _, n in sizes)
    self.param = torch.zeros(total, dtype=torch.float32, device=self.dev)
    self.grad = torch.zeros_like(self.param)
    self.adam_m = torch.zeros_like(self.param)
    self.adam_v = torch.zeros_like(self.param)
    self._off, off = {}, 0
    for name, n in sizes:
      self._off[name] = (off, off + n)
      off += n
    bufs = (self.param, self.grad, self.adam_m, self.adam_v)
    self.heads, self.p, self.g, self.am, self.av, self.wr_bf16 = [], {}, {}, {}, {}, []
    for i in range(layers + 1):
      a, b = self._off["moe%d" % i if i < layers else "main"]
      self.heads.append(HeadTrainer("moe", dims[i], vocab, mixtures, l2_penalty, self.dev, group, storage=tuple(t[a:b] for t in bufs)))
    for i in range(layers):
      for name, shp in (("wr%d" % i, (relu_cells, self.vpad)), ("br%d" % i, (relu_cells, 1))):   # Wr^T [relu_cells, V] (K-major)
        a, b = self._off[name]
        self.p[name], self.g[name] = self.param[a:b].view(shp), self.grad[a:b].view(shp)
        self.am[name], self.av[name] = self.adam_m[a:b].view(shp), self.adam_v[a:b].view(shp)
      self.wr_bf16.append(torch.zeros((relu_cells, self.vpad), dtype=torch.bfloat16, device=self.dev))
    self.dims = dims
    self.global_step = 0
    self.keep_grads = False
    self.last = {}

  # ---- TF names / layouts ----------------------------------------------------------------------------
  @staticmethod
  def _scope(i, layers):
    return "-prediction-%d" % i if i < layers else "--main"

  def import_state(self, sd):
    for i, head in enumerate(self.heads):
      sc = self._scope(i, self.nl)
      head.import_state({"gates/weights": sd["gates%s/weights" % sc], "experts/weights": sd["experts%s/weights" % sc],
                         "experts/biases": sd["experts%s/biases" % sc]})
    for i in range(self.nl):
      self.p["wr%d" % i].zero_()
      self.p["wr%d" % i][:, :self.v].copy_(sd["relu-%d/weights" % i].t().to(self.dev))
      self.p["br%d" % i].copy_(sd["relu-%d/biases" % i].view(-1, 1).to(self.dev))
      self.wr_bf16[i].copy_(self.p["wr%d" % i])

  def _tf_layout(self, flat):
    out = {}
    for i, head in enumerate(self.heads):
      sc = self._scope(i, self.nl)
      a, b = self._off["moe%d" % i if i < self.nl else "main"]
      for k, v in head.grads_tf_layout(flat[a:b]).items():
        name, leaf = k.split("/")
        out["%s%s/%s" % (name, sc, leaf)] = v
    for i in range(self.nl):
      a, b = self._off["wr%d" % i]
      out["relu-%d/weights" % i] = flat[a:b].view(self.r, self.vpad)[:, :self.v].t().contiguous().cpu()
      a, b = self._off["br%d" % i]
      out["relu-%d/biases" % i] = flat[a:b].cpu().clone()
    return out

  def export_state(self):
    return self._tf_layout(self.param)

  def grads_tf_layout(self, flat):
    return self._tf_layout(flat)

  # ---- step ------------------------------------------------------------------------------------------
  def step(self, x, labels, base_lr=0.01, lr_decay=0.95, lr_decay_examples=4000000, clip_gradient_norm=1.0,
           regularization_penalty=1.0, global_batch=None, support_labels=None, support_loss_percent=0.1):
    """support_labels [B, layers * V] (--multitask; the reference's chain scripts pass --support_type="label,label,..."): adds
    CE(concat of the sub-predictions, support labels) * support_loss_percent and weighs the main loss by the rest
    (wh/losses.py:271-279)."""
    b = x.shape[0]
    global_batch = global_batch or b * self.world
    pct = support_loss_percent if support_labels is not None else 0.0
    if support_labels is not None and support_labels.shape[1] != self.nl * self.v:
      raise ValueError("support labels have %d columns, the model emits %d support predictions" % (support_labels.shape[1], self.nl * self.v))
    cur = x.float()[:, :self.d].contiguous()
    saved = []
    for i in range(self.nl):
      sub, (s_hi, s_lo) = self.heads[i].forward(cur)                                   # [B, V] sub-prediction
      p_hi, p_lo = nat.split_bf16(sub[:, :self.v])
      z = nat.linear(p_hi, self.wr_bf16[i], a_lo=p_lo, n=self.r, k=self.vpad, shift=self.p["br%d" % i].view(-1), act="relu")["f32"]
      _, n32 = nat.l2norm_rows(z.contiguous(), want_f32=True)
      saved.append((sub, s_hi, s_lo, p_hi, z))
      cur = torch.cat([cur, n32], dim=1)                                               # tf.concat (device copy)
    p, (m_hi, m_lo) = self.heads[self.nl].forward(cur)
    loss, dcur = self.heads[self.nl].backward(p, m_hi, m_lo, labels, global_batch, want_dx=True, loss_weight=1.0 - pct)
    loss_parts = {"main": loss}
    for i in range(self.nl - 1, -1, -1):
      sub, s_hi, s_lo, p_hi, z = saved[i]
      dn = dcur[:, self.dims[i]:self.dims[i] + self.r].contiguous()                    # the columns that came from this layer
      dz = nat.l2norm_rows_bwd(z, dn)
      dzr_hi, dzr_lo = nat.act_bwd(dz, z, act="relu")
      nat.wgrad(dzr_hi, dzr_lo, p_hi, self.r, self.v, out=self.g["wr%d" % i])          # dWr^T [relu_cells, V]
      nat.colsum_bf16(dzr_hi, dzr_lo, self.r, out=self.g["br%d" % i].view(-1))
      wr_t = nat.pack_transpose(self.p["wr%d" % i])                                    # bf16 [V(pad), relu_cells]: dgrad operand
      dsub = nat.linear(dzr_hi, wr_t, a_lo=dzr_lo, n=self.v, k=self.r)["f32"].contiguous()   # dL/d sub-prediction [B, V]
      if support_labels is not None:                                                   # this layer's share of the support loss
        sl_i = support_labels[:, i * self.v:(i + 1) * self.v].contiguous()
        sup_loss, d_sup = nat.xent(sub, sl_i, want_grad=True, grad_scale=pct * b / float(global_batch))
        nat.add_inplace(dsub, d_sup)
        loss_parts["support%d" % i] = sup_loss
      dx_i = self.heads[i].backward_from_dp(dsub, sub, s_hi, s_lo, want_dx=(i > 0))
      if i > 0:
        dnext = dcur[:, :self.dims[i]].contiguous()                                    # what continues down the chain ...
        nat.add_inplace(dnext, dx_i[:, :self.dims[i]].contiguous())                    # ... + the path through this layer's MoE
        dcur = dnext
    yt8m_dp.all_reduce_sum_(self.grad, self.group)                                     # the ONE collective of the step
    if self.keep_grads:
      self.last_grad = self.grad.clone()
    lr = exponential_decay(base_lr, self.global_step, global_batch, lr_decay_examples, lr_decay)
    lr_t = adam_lr_t(lr, self.global_step + 1)
    for i in range(self.nl):
      sums = nat.grad_reg_sumsq(self.g["wr%d" % i], self.p["wr%d" % i], self.l2 * regularization_penalty)
      nat.clip_adam_step(self.p["wr%d" % i], self.g["wr%d" % i], self.am["wr%d" % i], self.av["wr%d" % i], sums, clip_gradient_norm,
                         lr_t, param_bf16=self.wr_bf16[i])
      sums = nat.grad_reg_sumsq(self.g["br%d" % i], self.p["br%d" % i], 0.0)
      nat.clip_adam_step(self.p["br%d" % i], self.g["br%d" % i], self.am["br%d" % i], self.av["br%d" % i], sums, clip_gradient_norm, lr_t)
    for head in self.heads:
      head.apply(lr_t, clip_gradient_norm, regularization_penalty)
    self.global_step += 1
    self.last = {"label_loss_local": _mix_loss(loss_parts, pct), "lr": lr}
    return p


class DbofTrainer(object):
  """The training step for DbofModel (wh/all_frame_models/dbof_model.py:62-123, --dbof_pooling_method=max) + MoeModel: sample
  `iterations` frames per video, cluster projection D -> cluster_size, ReLU6, max over the sampled frames, hidden FC, ReLU6, MoE
  head; the backward routes the pooled gradient to the arg-max frame of every (video, cluster) (yt8m_group_max_rows_bwd) and
  computes the cluster-weight gradient as one MN-major GEMM over all B*iterations sampled rows.

  batch_norm=True (the reference default, --dbof_add_batch_norm=True): slim.batch_norm(is_training=True) on the sampled input
  rows, after the cluster projection and after the hidden FC (dbof_model.py:64-108) -- batch statistics (biased variance),
  backward THROUGH the statistics (yt8m_bn_bwd), and the moving averages (decay 0.999) that wh/train.py:449-456 runs as
  update ops.  Data parallel: every rank normalises with the statistics of ITS shard (no collective in the forward pass); the
  batch moments ride at the tail of the flat gradient buffer, so the ONE all-reduce also averages them and every replica
  applies the same moving-average update.
  batch_norm=False: the bias form (cluster_biases / hidden1_biases)."""

  BN = (("input_bn", "d"), ("cluster_bn", "c"), ("hidden1_bn", "h"))

  def __init__(self, feature_dim, cluster_size=8192, hidden=1024, iterations=30, vocab=4716, mixtures=2, l2_penalty=1e-8,
               device=None, group=None, batch_norm=False):
    self.d, self.c, self.h, self.n, self.v, self.m = feature_dim, cluster_size, hidden, iterations, vocab, mixtures
    self.bn = batch_norm
    self.dev = device or torch.device("cuda", torch.cuda.current_device())
    self.group = group
    self.world = yt8m_dp.world_size(group)
    dims = {"d": feature_dim, "c": cluster_size, "h": hidden}
    if batch_norm:
      sizes = [("cw", cluster_size * feature_dim), ("wh", hidden * cluster_size)]
      for scope, k in self.BN:
        sizes += [(scope + "/gamma", dims[k]), (scope + "/beta", dims[k])]
    else:
      sizes = [("cw", cluster_size * feature_dim), ("cb", cluster_size), ("wh", hidden * cluster_size), ("bh", hidden)]
    sizes.append(("head", HeadTrainer.flat_size("moe", hidden, vocab, mixtures)))
    self._param_names = [n for n, _ in sizes if n != "head"]
    if batch_norm:                                   # batch moments: all-reduced with the gradients, never optimised
      for scope, k in self.BN:
        sizes += [(scope + "/mean", dims[k]), (scope + "/var", dims[k])]
    total = sum(n for _, n in sizes)
    self.param = torch.zeros(total, dtype=torch.float32, device=self.dev)
    self.grad = torch.zeros_like(self.param)
    self.adam_m = torch.zeros_like(self.param)
    self.adam_v = torch.zeros_like(self.param)
    self._off, off = {}, 0
    for name, n in sizes:
      self._off[name] = (off, off + n)
      off += n
    shapes = {"cw": (cluster_size, feature_dim), "wh": (hidden, cluster_size), "cb": (cluster_size, 1), "bh": (hidden, 1)}
    self.p, self.g, self.am, self.av = {}, {}, {}, {}
    for name, n in sizes:
      if name == "head":
        continue
      a, b = self._off[name]
      shp = shapes.get(name, (n, 1))
      self.p[name], self.g[name] = self.param[a:b].view(shp), self.grad[a:b].view(shp)
      self.am[name], self.av[name] = self.adam_m[a:b].view(shp), self.adam_v[a:b].view(shp)
    a, b = self._off["head"]
    self.head = HeadTrainer("moe", hidden, vocab, mixtures, l2_penalty, self.dev, group,
                            storage=(self.param[a:b], self.grad[a:b], self.adam_m[a:b], self.adam_v[a:b]))
    self.cw_bf16 = torch.zeros((cluster_size, feature_dim), dtype=torch.bfloat16, device=self.dev)
    self.wh_bf16 = torch.zeros((hidden, cluster_size), dtype=torch.bfloat16, device=self.dev)
    # moving statistics (non-trainable variables of the reference graph)
    self.moving = {}
    if batch_norm:
      for scope, k in self.BN:
        self.moving[scope + "/moving_mean"] = torch.zeros(dims[k], device=self.dev)
        self.moving[scope + "/moving_variance"] = torch.ones(dims[k], device=self.dev)
    self.sample_random_frames = True
    self.global_step = 0
    self.keep_grads = False
    self.last = {}

  def import_state(self, sd):
    dev = self.dev
    self.p["cw"].copy_(sd["cluster_weights"].t().to(dev))
    self.p["wh"].copy_(sd["hidden1_weights"].t().to(dev))
    if self.bn:
      for scope, _ in self.BN:
        self.p[scope + "/gamma"].copy_(sd[scope + "/gamma"].view(-1, 1).to(dev))
        self.p[scope + "/beta"].copy_(sd[scope + "/beta"].view(-1, 1).to(dev))
        self.moving[scope + "/moving_mean"].copy_(sd[scope + "/moving_mean"].to(dev))
        self.moving[scope + "/moving_variance"].copy_(sd[scope + "/moving_variance"].to(dev))
    else:
      self.p["cb"].copy_(sd["cluster_biases"].view(-1, 1).to(dev))
      self.p["bh"].copy_(sd["hidden1_biases"].view(-1, 1).to(dev))
    self.cw_bf16.copy_(self.p["cw"])
    self.wh_bf16.copy_(self.p["wh"])
    self.head.import_state({k: sd[k] for k in ("gates/weights", "experts/weights", "experts/biases")})

  def _tf_layout(self, flat, with_moving=False):
    v = {}
    for name in self._param_names:
      a, b = self._off[name]
      v[name] = flat[a:b].view(self.p[name].shape)
    out = {"cluster_weights": v["cw"].t().contiguous().cpu(), "hidden1_weights": v["wh"].t().contiguous().cpu()}
    if self.bn:
      for scope, _ in self.BN:
        out[scope + "/gamma"] = v[scope + "/gamma"].reshape(-1).cpu().clone()
        out[scope + "/beta"] = v[scope + "/beta"].reshape(-1).cpu().clone()
        if with_moving:
          out[scope + "/moving_mean"] = self.moving[scope + "/moving_mean"].cpu().clone()
          out[scope + "/moving_variance"] = self.moving[scope + "/moving_variance"].cpu().clone()
    else:
      out["cluster_biases"] = v["cb"].reshape(-1).cpu().clone()
      out["hidden1_biases"] = v["bh"].reshape(-1).cpu().clone()
    a, b = self._off["head"]
    out.update(self.head.grads_tf_layout(flat[a:b]))
    return out

  def export_state(self):
    return self._tf_layout(self.param, with_moving=True)

  def grads_tf_layout(self, flat):
    return self._tf_layout(flat)

  def _bn_forward(self, scope, x, act, want_bf16):
    out, stats = nat.bn_train_fwd(x, self.p[scope + "/gamma"].view(-1), self.p[scope + "/beta"].view(-1), act=act, want_bf16=want_bf16)
    # this rank's batch moments, pre-divided by the world size: the all-reduce turns them into the cross-rank average
    self.g[scope + "/mean"].view(-1).copy_(stats[0]).div_(self.world) if self.world > 1 else self.g[scope + "/mean"].view(-1).copy_(stats[0])
    self.g[scope + "/var"].view(-1).copy_(stats[1]).div_(self.world) if self.world > 1 else self.g[scope + "/var"].view(-1).copy_(stats[1])
    return out, stats

  def step(self, x, num_frames, labels, base_lr=0.01, lr_decay=0.95, lr_decay_examples=4000000, clip_gradient_norm=1.0,
           regularization_penalty=1.0, global_batch=None, frame_index=None):
    """frame_index ([B, iterations] int64) pins the sampled frames (tests); by default they are drawn like
    wh/model_utils.py:56-74."""
    import frame_level_models
    b, t, d = x.shape
    n = self.n
    global_batch = global_batch or b * self.world
    if frame_index is None:
      frame_index = frame_level_models.sample_frames(num_frames, n, self.sample_random_frames)
    rows = x[torch.arange(b, device=x.device).unsqueeze(1), frame_index.to(x.device)].reshape(b * n, d).contiguous()   # gather_nd
    if self.bn:
      r, st1 = self._bn_forward("input_bn", rows, None, True)                                                          # hi + lo operand
      z2 = nat.linear(r["hi"], self.cw_bf16, a_lo=r["lo"], n=self.c, k=d)["f32"]                                       # [B*n, C]
      a2, st2 = self._bn_forward("cluster_bn", z2, "relu6", False)
      act = a2["f32"]
      rows_op = r["hi"]
    else:
      act = nat.linear(rows, self.cw_bf16, n=self.c, k=d, shift=self.p["cb"].view(-1), act="relu6")["f32"]             # [B*n, C]
      rows_op = rows
    pooled = nat.group_max_rows(act, n)                                                                                # [B, C]
    q_hi, q_lo = nat.split_bf16(pooled)
    if self.bn:
      z3 = nat.linear(q_hi, self.wh_bf16, a_lo=q_lo, n=self.h, k=self.c)["f32"]
      hid, st3 = self._bn_forward("hidden1_bn", z3, "relu6", True)
    else:
      hid = nat.linear(q_hi, self.wh_bf16, a_lo=q_lo, n=self.h, k=self.c, shift=self.p["bh"].view(-1), act="relu6", out_f32=True,
                       out_bf16=True, out_lo=True)
    p = nat.moe_fwd(hid["hi"], self.head.w_bf16, self.head.b, self.v, self.m, x_lo=hid["lo"], d=self.h)
    # ---- backward
    loss, dhid = self.head.backward(p, hid["hi"], hid["lo"], labels, global_batch, want_dx=True)
    dhid = dhid[:, :self.h].contiguous()
    if self.bn:
      dg, db, dpre_hi, dpre_lo, _ = nat.bn_train_bwd(dhid, hid["f32"], z3, st3, self.p["hidden1_bn/gamma"].view(-1), act="relu6")
      self.g["hidden1_bn/gamma"].view(-1).copy_(dg)
      self.g["hidden1_bn/beta"].view(-1).copy_(db)
    else:
      dpre_hi, dpre_lo = nat.act_bwd(dhid, hid["f32"], act="relu6")
      nat.colsum_bf16(dpre_hi, dpre_lo, self.h, out=self.g["bh"].view(-1))
    nat.wgrad(dpre_hi, dpre_lo, q_hi, self.h, self.c, out=self.g["wh"])                  # d hidden1_weights^T [H, C]
    wh_t = nat.pack_transpose(self.p["wh"])                                              # bf16 [C, H]: the dgrad operand
    dpooled = nat.linear(dpre_hi, wh_t, a_lo=dpre_lo, n=self.c, k=self.h)["f32"]
    del wh_t
    dact = nat.group_max_rows_bwd(act, dpooled.contiguous(), n)                          # to the arg-max frame of every (video, cluster)
    if self.bn:
      dg, db, dz_hi, dz_lo, _ = nat.bn_train_bwd(dact, act, z2, st2, self.p["cluster_bn/gamma"].view(-1), act="relu6")
      self.g["cluster_bn/gamma"].view(-1).copy_(dg)
      self.g["cluster_bn/beta"].view(-1).copy_(db)
    else:
      dz_hi, dz_lo = nat.act_bwd(dact, act, act="relu6")
      nat.colsum_bf16(dz_hi, dz_lo, self.c, out=self.g["cb"].view(-1))
    nat.wgrad(dz_hi, dz_lo, rows_op, self.c, d, out=self.g["cw"])                        # d cluster_weights^T [C, D]
    if self.bn:
      # input_bn has trainable gamma / beta: the cluster projection's input gradient, then the statistics' backward (no dx:
      # the frames themselves are data)
      cw_t = nat.pack_transpose(self.p["cw"])                                            # bf16 [D, C]
      drows = nat.linear(dz_hi, cw_t, a_lo=dz_lo, n=d, k=self.c)["f32"]
      del cw_t
      dg, db, _, _, _ = nat.bn_train_bwd(drows, None, rows, st1, self.p["input_bn/gamma"].view(-1), act=None, want_dx=False)
      self.g["input_bn/gamma"].view(-1).copy_(dg)
      self.g["input_bn/beta"].view(-1).copy_(db)
    yt8m_dp.all_reduce_sum_(self.grad, self.group)                                       # the ONE collective of the step
    if self.keep_grads:
      self.last_grad = self.grad.clone()
    lr = exponential_decay(base_lr, self.global_step, global_batch, lr_decay_examples, lr_decay)
    lr_t = adam_lr_t(lr, self.global_step + 1)
    bf = {"cw": self.cw_bf16, "wh": self.wh_bf16}
    for name in self._param_names:
      sums = nat.grad_reg_sumsq(self.g[name], self.p[name], 0.0)
      nat.clip_adam_step(self.p[name], self.g[name], self.am[name], self.av[name], sums, clip_gradient_norm, lr_t, param_bf16=bf.get(name))
    self.head.apply(lr_t, clip_gradient_norm, regularization_penalty)
    if self.bn:                                                                          # update ops (wh/train.py:449-456)
      for scope, _ in self.BN:
        nat.bn_moving_update(self.moving[scope + "/moving_mean"], self.moving[scope + "/moving_variance"],
                             self.g[scope + "/mean"].view(-1), self.g[scope + "/var"].view(-1))
    self.global_step += 1
    self.head.global_step = self.global_step
    self.last = {"label_loss_local": loss, "lr": lr}
    return p
