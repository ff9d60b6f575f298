"""The training step of wh/train.py:262-479 for the video-level heads (LogisticModel, MoeModel), on the GPU
through the C ABI: forward, CrossEntropyLoss + its gradient, backward (fused MoE backward epilogue, MN-major
tcgen05 wgrad), L2 regulariser, per-tensor clip_by_norm, TF-1.0 Adam, exponential-decay learning rate.

Data parallel: one process per GPU; every rank computes gradients of ITS shard of the batch
with the loss gradient pre-scaled by 1/world, then ONE all-reduce(sum) over a single flat fp32 gradient buffer
(NCCL over NVLink on the GPU box, gloo in the CPU tests) gives every rank the gradient of the mean loss over
the global batch -- the single-process large-batch semantics of the reference's build_graph (its own multi-
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

  def __init__(self, kind, in_dim, vocab, mixtures=2, l2_penalty=1e-8, device=None, group=None):
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

  def step(self, x, labels, base_lr=0.01, lr_decay=0.95, lr_decay_examples=4000000, clip_gradient_norm=1.0,
           regularization_penalty=1.0, global_batch=None):
    """One optimiser step on this rank's shard (x [B_local, D], labels [B_local, V]).  Returns predictions;
    loss terms are left on the device in ``self.last`` (fetch with .item() only when logging)."""
    b_local = x.shape[0]
    global_batch = global_batch or b_local * self.world
    p, (hi, lo) = self.forward(x)
    # d(mean over the GLOBAL batch)/dp: xent divides by the local batch, so rescale by B_local / B_global
    loss, dp = nat.xent(p, labels, want_grad=True, grad_scale=b_local / float(global_batch))
    if self.kind == "logistic":
      dz_hi, dz_lo = nat.logistic_bwd_dz(dp, p)
    else:
      dz_hi, dz_lo = nat.moe_bwd_dlogits(hi, lo, self.w_bf16, self.b, dp, self.v, self.m, d=self.d)
    nat.wgrad(dz_hi, dz_lo, hi, self.rows, self.d, out=self.gw)          # dW^T[rows, D] = dZ^T . X  (lo of x is dropped)
    nat.colsum_bf16(dz_hi, dz_lo, self.rows, out=self.gb)
    yt8m_dp.all_reduce_sum_(self.grad, self.group)                        # the ONE collective of the step
    if self.keep_grads:
      self.last_grad = self.grad.clone()                                  # tests: gradient parity before reg/clip
    lr = exponential_decay(base_lr, self.global_step, global_batch, lr_decay_examples, lr_decay)
    lr_t = adam_lr_t(lr, self.global_step + 1)
    sums_w = nat.grad_reg_sumsq(self.gw, self.w, self.l2 * regularization_penalty, self.per, self.m)
    sums_b = nat.grad_reg_sumsq(self.gb.view(self.rows, 1), self.b.view(self.rows, 1), 0.0, self.per, self.m)
    nat.clip_adam_step(self.w, self.gw, self.mw, self.vw, sums_w, clip_gradient_norm, lr_t, moe_per=self.per, moe_nmix=self.m,
                       param_bf16=self.w_bf16)
    nat.clip_adam_step(self.b.view(self.rows, 1), self.gb.view(self.rows, 1), self.mb.view(self.rows, 1), self.vb.view(self.rows, 1),
                       sums_b, clip_gradient_norm, lr_t, moe_per=self.per, moe_nmix=self.m, only_segment=1 if self.kind == "moe" else -1)
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
