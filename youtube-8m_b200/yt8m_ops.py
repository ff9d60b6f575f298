"""Host-side glue between the model plugins and the CUDA library: activations handles, the variable
store (the stand-in for TF-1.0 variable scopes / collections the reference relies on), weight
packing caches and thin op functions.  No arithmetic happens here: every op forwards to
``yt8m_native`` (the C ABI); PyTorch is used for device memory and streams only.
"""
import math

import torch

import yt8m_native as nat


def device():
  return torch.device("cuda", torch.cuda.current_device())


# ------------------------------------------------------------------------------------------------
# activations
# ------------------------------------------------------------------------------------------------

class Act(object):
  """A 2-D activation [rows, cols] living on the GPU as a bf16 hi (+ lo) pair and/or fp32.

  The tensor cores take bf16 operands; ``hi + lo`` carries ~16 mantissa bits of an fp32 activation
  through them (two MMAs against the same weight tile).  Frame features are bf16-exact: lo is None.
  hi/lo row stride is a multiple of 8 elements (TMA), pad columns are zero.
  """
  __slots__ = ("f32", "hi", "lo", "cols")

  def __init__(self, f32=None, hi=None, lo=None, cols=None):
    self.f32, self.hi, self.lo = f32, hi, lo
    self.cols = cols if cols is not None else (f32.shape[1] if f32 is not None else hi.shape[1])

  @property
  def rows(self):
    return (self.f32 if self.f32 is not None else self.hi).shape[0]

  def operand(self):
    """(hi, lo) bf16 operand views for a GEMM; splits the fp32 value on first use."""
    if self.hi is None:
      self.hi, self.lo = nat.split_bf16(self.f32)
    return self.hi, self.lo

  def float(self):
    if self.f32 is None:
      v = self.hi[:, :self.cols].float()
      if self.lo is not None:
        v = v + self.lo[:, :self.cols].float()
      self.f32 = v
    return self.f32[:, :self.cols]


def as_act(x):
  """Accepts an Act, or a 2-D torch tensor (fp32 -> split on demand; bf16 -> exact operand)."""
  if isinstance(x, Act):
    return x
  if not torch.is_tensor(x):
    raise TypeError("expected a tensor or Act, got %r" % type(x))
  if x.dim() != 2:
    raise ValueError("expected a [rows, cols] tensor, got shape %s" % (tuple(x.shape),))
  x = x.to(device())
  if x.dtype == torch.bfloat16:
    if x.stride(1) != 1 or x.stride(0) % 8 != 0:
      cols = x.shape[1]
      buf = torch.zeros((x.shape[0], nat.pad8(cols)), dtype=torch.bfloat16, device=x.device)
      buf[:, :cols] = x
      x = buf[:, :cols]
    return Act(hi=x, cols=x.shape[1])
  return Act(f32=x.float().contiguous())


def concat(acts):
  """tf.concat(axis=1) of activations -> one operand (hi/lo buffers written side by side)."""
  acts = [as_act(a) for a in acts]
  rows = acts[0].rows
  total = sum(a.cols for a in acts)
  dev = device()
  hi = torch.zeros((rows, nat.pad8(total)), dtype=torch.bfloat16, device=dev)
  lo = torch.zeros((rows, nat.pad8(total)), dtype=torch.bfloat16, device=dev)
  f32 = torch.empty((rows, total), dtype=torch.float32, device=dev)
  off = 0
  for a in acts:
    f = a.float()
    f32[:, off:off + a.cols] = f                                  # device-side copy (no arithmetic)
    nat.split_bf16(f if f.stride(1) == 1 else f.contiguous(), out_hi=hi[:, off:off + a.cols], out_lo=lo[:, off:off + a.cols])
    off += a.cols
  return Act(f32=f32, hi=hi[:, :total], lo=lo[:, :total], cols=total)


# ------------------------------------------------------------------------------------------------
# variables (TF-1.0 variable-scope / collection stand-in)
# ------------------------------------------------------------------------------------------------

class Variable(object):
  __slots__ = ("name", "value", "trainable", "l2", "version", "grad")

  def __init__(self, name, value, trainable=True, l2=None):
    self.name, self.value, self.trainable, self.l2 = name, value, trainable, l2
    self.version = 0
    self.grad = None

  def assign(self, value):
    self.value.copy_(value.to(self.value.device, self.value.dtype))
    self.version += 1


class VariableStore(object):
  """name -> Variable, created on first use like ``tf.get_variable`` under the default graph.

  Also the registry behind ``tf.losses.get_regularization_losses()`` (wh/train.py:440-442): every
  variable created with an ``l2`` scale contributes l2 * sum(w^2) / 2.
  """

  def __init__(self, seed=9):
    self.vars = {}
    self.seed = seed
    self._packed = {}
    self._gen = None
    self.scope = []

  def reset(self, seed=None):
    self.vars.clear()
    self._packed.clear()
    self._gen = None
    if seed is not None:
      self.seed = seed

  def _generator(self):
    if self._gen is None:
      self._gen = torch.Generator().manual_seed(self.seed)
    return self._gen

  def full_name(self, name):
    return "/".join([s for s in self.scope if s] + [name])

  def get(self, name, shape, init, trainable=True, l2=None, round_bf16=True):
    key = self.full_name(name)
    v = self.vars.get(key)
    if v is not None:
      if tuple(v.value.shape) != tuple(shape):
        raise ValueError("variable %s exists with shape %s, requested %s" % (key, tuple(v.value.shape), tuple(shape)))
      return v
    val = init(tuple(shape), self._generator())
    if round_bf16:
      # weights are kept bf16-representable so that the tensor-core operand copy is exact
      val = val.to(torch.bfloat16).to(torch.float32)
    v = Variable(key, val.to(device()), trainable, l2)
    self.vars[key] = v
    return v

  def packed(self, var, kind, builder, version=None):
    """Cached device copy of ``var`` in a kernel layout; rebuilt when the variable changes
    (``version`` overrides var.version for caches that depend on several variables)."""
    key = (var.name, kind)
    ver = var.version if version is None else version
    hit = self._packed.get(key)
    if hit is not None and hit[0] == ver:
      return hit[1]
    val = builder()
    self._packed[key] = (ver, val)
    return val

  def trainable_variables(self):
    return [v for v in self.vars.values() if v.trainable]

  def state_dict(self):
    return {k: v.value.detach().cpu() for k, v in self.vars.items()}

  def load_state_dict(self, sd, strict=True):
    if strict:
      # both directions: a model variable that the checkpoint lacks would silently keep its random initialisation
      missing = sorted(k for k in self.vars if k not in sd)
      if missing:
        raise KeyError("model variables missing from the checkpoint: %s" % ", ".join(missing))
    for k, t in sd.items():
      if k in self.vars:
        self.vars[k].assign(t)
      elif strict:
        raise KeyError("checkpoint variable %s not in the model" % k)
      else:
        self.vars[k] = Variable(k, t.to(device()).float())


_DEFAULT_STORE = VariableStore()


def get_store():
  return _DEFAULT_STORE


class variable_scope(object):
  def __init__(self, name):
    self.name = name

  def __enter__(self):
    _DEFAULT_STORE.scope.append(self.name)

  def __exit__(self, *exc):
    _DEFAULT_STORE.scope.pop()
    return False


# initialisers (signatures: (shape, generator) -> cpu fp32 tensor)
def xavier_uniform(shape, gen):
  """slim.fully_connected default weights_initializer (xavier_initializer, uniform)."""
  lim = math.sqrt(6.0 / (shape[0] + shape[1]))
  return (torch.rand(shape, generator=gen) * 2 - 1) * lim


def zeros_init(shape, gen):
  return torch.zeros(shape)


def ones_init(shape, gen):
  return torch.ones(shape)


def constant_init(c):
  return lambda shape, gen: torch.full(shape, float(c))


def random_normal(stddev):
  return lambda shape, gen: torch.randn(shape, generator=gen) * stddev


def truncated_normal(stddev):
  def init(shape, gen):
    t = torch.randn(shape, generator=gen)
    bad = t.abs() > 2
    while bad.any():
      t[bad] = torch.randn(int(bad.sum()), generator=gen)
      bad = t.abs() > 2
    return t * stddev
  return init


# ------------------------------------------------------------------------------------------------
# ops
# ------------------------------------------------------------------------------------------------

def fully_connected(x, num_outputs, scope, activation_fn="relu", use_bias=True, l2_penalty=None,
                    weights_initializer=xavier_uniform, biases_initializer=zeros_init, scale=None, want_bf16=True):
  """slim.fully_connected(x, num_outputs, activation_fn=..., weights_regularizer=l2(l2_penalty), scope=scope).

  Variables: <scope>/weights [in, out] and <scope>/biases [out] (TF layout).  Note slim's DEFAULT
  activation is ReLU; every reference call site on the hot path overrides it.
  """
  x = as_act(x)
  st = get_store()
  w = st.get(scope + "/weights", (x.cols, num_outputs), weights_initializer, l2=l2_penalty)
  b = st.get(scope + "/biases", (num_outputs,), biases_initializer, round_bf16=False) if use_bias else None
  hi, lo = x.operand()
  if hi.dtype == torch.float16:      # fp16 activation (e.g. a NetVLAD hidden layer feeding a chain model): fp16 weight copy
    wp = st.packed(w, "kmajor_f16", lambda: nat.pack_transpose(w.value).to(torch.float16))
  else:
    wp = st.packed(w, "kmajor", lambda: nat.pack_transpose(w.value))
  res = nat.linear(hi, wp, a_lo=lo, n=num_outputs, k=x.cols, scale=scale, shift=b.value if b is not None else None,
                   act=activation_fn, out_f32=True, out_bf16=want_bf16, out_lo=want_bf16)
  return Act(f32=res["f32"], hi=res.get("hi"), lo=res.get("lo"), cols=num_outputs)


def moe_head(x, vocab_size, num_mixtures, gates_scope, experts_scope, l2_penalty=1e-8):
  """The MoE head of wh/all_video_models/moe_model.py:38-65 (also the sub_moe / sub_model copies):
  gates FC (no bias) + experts FC, softmax(M+1) x sigmoid(M), summed over mixtures -- one fused kernel.
  Variables keep the reference names/layouts: <gates>/weights [D, V(M+1)], <experts>/weights [D, VM],
  <experts>/biases [VM]."""
  x = as_act(x)
  st = get_store()
  d, v, m = x.cols, vocab_size, num_mixtures
  gw = st.get(gates_scope + "/weights", (d, v * (m + 1)), xavier_uniform, l2=l2_penalty)
  ew = st.get(experts_scope + "/weights", (d, v * m), xavier_uniform, l2=l2_penalty)
  eb = st.get(experts_scope + "/biases", (v * m,), zeros_init, round_bf16=False)

  def build():
    return nat.moe_pack(gw.value, ew.value, eb.value, v, m)

  hi, lo = x.operand()
  if hi.dtype == torch.float16:      # fp16 activation: the weight operand must be fp16 too (one format per MMA)
    def build16():
      wp, bp = build()
      return wp.to(torch.float16), bp
    wp, bp = st.packed(gw, "moe_f16", build16, version=(gw.version, ew.version, eb.version))
  else:
    wp, bp = st.packed(gw, "moe", build, version=(gw.version, ew.version, eb.version))
  return nat.moe_fwd(hi, wp, bp, v, m, x_lo=lo, d=d)


def l2_normalize_rows(x):
  """tf.nn.l2_normalize(x, dim=1) on a 2-D activation."""
  x = as_act(x)
  f = x.float()
  f = f if f.is_contiguous() and f.shape[1] % 8 == 0 else _pad_cols(f)
  out_bf, out_f = nat.l2norm_rows(f, want_f32=True)
  return Act(f32=out_f[:, :x.cols], cols=x.cols)


def _pad_cols(f):
  rows, cols = f.shape
  buf = torch.zeros((rows, nat.pad8(cols)), dtype=torch.float32, device=f.device)
  buf[:, :cols] = f
  return buf


def frames_operand(model_input):
  """[B, T, D] frame features -> contiguous bf16 on the GPU (the layout the TMA tensor maps expect).
  fp32 input is converted by the row kernel (no normalisation: the transformer already ran)."""
  x = model_input.to(device())
  if x.dtype == torch.bfloat16:
    return x.contiguous()
  return nat.l2norm_rows(x.float().contiguous(), normalize=False)


class CapturedStep(object):
  """CUDA-graph replay of one fixed-shape plugin call (`fn()` -> tensor or dict of tensors): the launch-bound
  host side of a step (Python, ctypes, tensor-map encoding, output allocation) is paid once at capture.
  `fn` must read its inputs from fixed device buffers (refill them in place between replays) and launch only on
  the current stream.  `time_tag`: library calls whose name contains it get a CUDA-event pair captured around
  them (see yt8m_native.kernel_timer_begin); `kernel_ms()` returns those durations for the LAST replay."""

  def __init__(self, fn, warmup=2, time_tag=None):
    dev_stream = torch.cuda.Stream()
    dev_stream.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(dev_stream):
      for _ in range(max(warmup, 1)):     # lazy weight packing, function attributes, workspaces: all before capture
        fn()
    torch.cuda.current_stream().wait_stream(dev_stream)
    torch.cuda.synchronize()
    self.graph = torch.cuda.CUDAGraph()
    self._pairs = []
    if time_tag:
      nat.kernel_timer_begin(time_tag)
    try:
      with torch.cuda.graph(self.graph, stream=dev_stream):
        self.output = fn()
    finally:
      if time_tag:
        self._pairs = nat.kernel_timer_end(read=False)

  def __call__(self):
    self.graph.replay()
    return self.output

  def kernel_ms(self):
    return nat.kernel_timer_read(self._pairs)

