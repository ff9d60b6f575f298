"""ctypes binding of libyt8m_b200.so (the C ABI declared in include/yt8m_b200.h).

This is the ONLY route from the Python plugins to arithmetic: there is no CPU or PyTorch fallback.
If the library is missing or fails to load, importing this module raises.  Every wrapper takes
torch CUDA tensors (used purely as device-memory handles), passes raw pointers + the current CUDA
stream, and raises ``Yt8mError`` with ``yt8m_last_error()`` on a non-zero status.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libyt8m_b200.so")

c_void_p, c_int, c_ll, c_float, c_size_t = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_float, ctypes.c_size_t


class Yt8mError(RuntimeError):
  pass


def _load():
  if not os.path.exists(LIB_PATH):
    raise ImportError(
        "libyt8m_b200.so not found at %s -- build it with `python youtube-8m_b200/build_native.py` "
        "(there is no fallback path)" % LIB_PATH)
  return ctypes.CDLL(LIB_PATH)


_lib = _load()

_SIGS = {
    "yt8m_version": (c_int, []),
    "yt8m_last_error": (ctypes.c_char_p, []),
    "yt8m_launch_count": (c_ll, []),
    "yt8m_l2norm_rows_fwd": (c_int, [c_void_p, c_int, c_ll, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "yt8m_frames_unpack_u8": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "yt8m_linear_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "yt8m_linear_fwd": (c_int, [c_void_p, c_void_p, c_ll, c_void_p, c_ll, c_int, c_int, c_int, c_void_p, c_void_p, c_int,
                                c_int, c_int, c_void_p, c_void_p, c_void_p, c_ll, c_void_p, c_size_t, c_void_p]),
    "yt8m_pack_transpose_bf16": (c_int, [c_void_p, c_int, c_int, c_void_p, c_ll, c_void_p]),
    "yt8m_moe_packed_rows": (c_ll, [c_int, c_int]),
    "yt8m_moe_pack_weights": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_ll, c_void_p, c_void_p]),
    "yt8m_moe_fwd": (c_int, [c_void_p, c_void_p, c_ll, c_void_p, c_ll, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p,
                             c_ll, c_void_p]),
    "yt8m_group_max_rows": (c_int, [c_void_p, c_ll, c_int, c_int, c_void_p, c_void_p]),
    "yt8m_lstm_pack_weights": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "yt8m_lstm_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int]),
    "yt8m_lstm_fwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_float, c_void_p,
                              c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "yt8m_lstm_fwd_train": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_float, c_void_p,
                                    c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "yt8m_lstm_bwd_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int]),
    "yt8m_lstm_bwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_float,
                              c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "yt8m_attn_pool_bwd": (c_int, [c_void_p, c_ll, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_ll,
                                   c_void_p, c_void_p]),
    "yt8m_context_gate_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_ll, c_int, c_void_p, c_void_p, c_void_p,
                                      c_void_p, c_ll, c_void_p]),
    "yt8m_add_inplace": (c_int, [c_void_p, c_void_p, c_ll, c_void_p]),
    "yt8m_l2norm_rows_bwd": (c_int, [c_void_p, c_void_p, c_ll, c_int, c_void_p, c_void_p]),
    "yt8m_group_max_rows_bwd": (c_int, [c_void_p, c_void_p, c_ll, c_int, c_int, c_void_p, c_void_p]),
    "yt8m_attn_pool_fwd": (c_int, [c_void_p, c_ll, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                   c_void_p, c_void_p]),
    "yt8m_attn_pool_fused": (c_int, [c_void_p, c_void_p, c_ll, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "yt8m_netvlad_fwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                 c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_ll, c_int, c_void_p, c_void_p]),
    "yt8m_netvlad_tiled_supported": (c_int, [c_int, c_int, c_int]),
    "yt8m_netvlad_tiled_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "yt8m_netvlad_fwd_tiled": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                       c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "yt8m_netvlad_bwd_norm": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                      c_void_p, c_void_p, c_void_p]),
    "yt8m_netvlad_bwd_assign_fused_supported": (c_int, [c_int, c_int, c_int]),
    "yt8m_netvlad_bwd_assign_fused": (c_int, [c_void_p, c_void_p, c_void_p, c_ll, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                              c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "yt8m_netvlad_bwd_assign": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                        c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "yt8m_act_bwd": (c_int, [c_void_p, c_void_p, c_ll, c_int, c_int, c_void_p, c_void_p, c_void_p, c_ll, c_void_p]),
    "yt8m_debug_set_timeline": (c_int, [c_void_p]),
    "yt8m_debug_set_flags": (c_int, [c_int]),
    "yt8m_context_gate_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_ll, c_int, c_void_p, c_void_p, c_void_p,
                                      c_void_p]),
    "yt8m_col_affine": (c_int, [c_void_p, c_int, c_ll, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "yt8m_split_bf16": (c_int, [c_void_p, c_ll, c_int, c_ll, c_void_p, c_void_p, c_ll, c_void_p]),
    "yt8m_xent_fwd_bwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_float, c_void_p]),
    "yt8m_logistic_bwd_dz": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_ll, c_void_p]),
    "yt8m_wgrad": (c_int, [c_void_p, c_void_p, c_ll, c_void_p, c_ll, c_int, c_int, c_int, c_void_p, c_ll, c_void_p]),
    "yt8m_colsum_bf16": (c_int, [c_void_p, c_void_p, c_ll, c_int, c_int, c_void_p, c_void_p]),
    "yt8m_moe_bwd_dlogits": (c_int, [c_void_p, c_void_p, c_ll, c_void_p, c_ll, c_void_p, c_void_p, c_ll, c_int, c_int, c_int, c_int,
                                     c_void_p, c_void_p, c_ll, c_void_p]),
    "yt8m_grad_reg_sumsq": (c_int, [c_void_p, c_void_p, c_ll, c_int, c_float, c_int, c_int, c_void_p, c_void_p]),
    "yt8m_clip_adam_step": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_ll, c_int, c_void_p, c_float, c_float, c_float,
                                    c_float, c_float, c_int, c_int, c_int, c_void_p, c_void_p]),
    "yt8m_bn_stats": (c_int, [c_void_p, c_int, c_ll, c_int, c_ll, c_void_p, c_void_p, c_void_p]),
    "yt8m_bn_fold": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_void_p]),
    "yt8m_col_affine_act": (c_int, [c_void_p, c_int, c_ll, c_int, c_ll, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_ll, c_void_p]),
    "yt8m_bn_bwd": (c_int, [c_void_p, c_ll, c_void_p, c_ll, c_void_p, c_int, c_ll, c_void_p, c_void_p, c_float, c_void_p, c_int, c_ll, c_int,
                            c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_ll, c_void_p]),
    "yt8m_topk_rows": (c_int, [c_void_p, c_ll, c_int, c_int, c_void_p, c_void_p, c_void_p]),
}
for _name, (_res, _args) in _SIGS.items():
  _fn = getattr(_lib, _name)          # AttributeError here == header/library mismatch: fail loudly
  _fn.restype = _res
  _fn.argtypes = _args

EXPORTS = tuple(_SIGS)

ACT = {"none": 0, None: 0, "relu": 1, "relu6": 2, "sigmoid": 3, "tanh": 4}
SRC_F32, SRC_BF16, SRC_U8 = 0, 1, 2


def version():
  return _lib.yt8m_version()


def launch_count():
  """CUDA kernels launched by the library so far (bench.py's gpu_launches)."""
  return _lib.yt8m_launch_count()


_KT = None   # kernel timer state: {"tag": str, "ev": [(start_event, end_event), ...]}


def kernel_timer_begin(tag):
  """bench.py: record a CUDA-event pair (on the launching stream) around every library call whose
  name contains `tag`, so a kernel's duration is measured live inside the timed region."""
  global _KT
  _KT = {"tag": tag, "ev": []}


def kernel_timer_end(read=True):
  """Stops timing.  read=True: returns the per-call durations in ms (synchronises); read=False: returns the
  event pairs (captured steps: read them with kernel_timer_read() after the replays)."""
  global _KT
  kt, _KT = _KT, None
  if not kt:
    return []
  if not read:
    return kt["ev"]
  return kernel_timer_read(kt["ev"])


def kernel_timer_read(pairs):
  torch.cuda.synchronize()
  return [a.elapsed_time(b) for a, b in pairs]


class region(object):
  """``with nat.region("attention_pool"):`` -- a named group of library calls.  A no-op unless bench.py's kernel timer is
  armed with exactly this tag, in which case ONE CUDA-event pair brackets the whole group (e.g. the logits GEMM + the
  pooling kernel of the attention pooler)."""

  def __init__(self, tag):
    self.tag, self.e1 = tag, None

  def __enter__(self):
    global _KT
    if _KT is not None and _KT["tag"] == self.tag:
      ext = torch.cuda.is_current_stream_capturing()
      e0 = torch.cuda.Event(enable_timing=True, external=ext)
      self.e1 = torch.cuda.Event(enable_timing=True, external=ext)
      e0.record()
      _KT["ev"].append((e0, self.e1))
    return self

  def __exit__(self, *exc):
    if self.e1 is not None:
      self.e1.record()
    return False


def _call(name, *args):
  fn = getattr(_lib, name)
  if _KT is not None and _KT["tag"] in name:
    # inside a stream capture the pair becomes event-record NODES of the graph (external events): every replay
    # re-records them, so elapsed_time() after a replay is that replay's kernel duration
    ext = torch.cuda.is_current_stream_capturing()
    e0, e1 = torch.cuda.Event(enable_timing=True, external=ext), torch.cuda.Event(enable_timing=True, external=ext)
    e0.record()
    rc = fn(*args)
    e1.record()
    _KT["ev"].append((e0, e1))
  else:
    rc = fn(*args)
  _check(rc, name)


def _check(rc, what):
  if rc != 0:
    raise Yt8mError("%s failed (%d): %s" % (what, rc, _lib.yt8m_last_error().decode()))


def _p(t):
  if t is None:
    return None
  if not t.is_cuda:
    raise Yt8mError("yt8m_b200 needs CUDA tensors (got a %s tensor); there is no CPU path" % t.device)
  return t.data_ptr()


def _stream():
  return torch.cuda.current_stream().cuda_stream


def _bf16(shape, device):
  return torch.empty(shape, dtype=torch.bfloat16, device=device)


def _f32(shape, device):
  return torch.empty(shape, dtype=torch.float32, device=device)


def pad8(n):
  return (n + 7) // 8 * 8


# ------------------------------------------------------------------------------------------------
# wrappers
# ------------------------------------------------------------------------------------------------

def l2norm_rows(x, normalize=True, num_frames=None, want_f32=False):
  """x: [..., dim] fp32 / bf16 / uint8 (contiguous).  Returns bf16 (and fp32 if asked)."""
  assert x.is_contiguous()
  dim = x.shape[-1]
  rows = x.numel() // dim
  src = {torch.float32: SRC_F32, torch.bfloat16: SRC_BF16, torch.uint8: SRC_U8}[x.dtype]
  out = _bf16(x.shape, x.device)
  of = _f32(x.shape, x.device) if want_f32 else None
  fpv = x.shape[-2] if (num_frames is not None and x.dim() >= 3) else 0
  _check(_lib.yt8m_l2norm_rows_fwd(_p(x), src, rows, dim, int(bool(normalize)), _p(num_frames), fpv, _p(out), _p(of),
                                   _stream()), "yt8m_l2norm_rows_fwd")
  return (out, of) if want_f32 else out


def frames_unpack_u8(packed, row_offsets, num_frames, max_frames, normalize=True, out=None):
  """packed uint8 [sum nf, D] (device), row_offsets int64 [B], num_frames int32 [B] -> padded bf16 [B, max_frames, D]."""
  b, d = num_frames.shape[0], packed.shape[1]
  if packed.shape[0] == 0:                       # a batch of empty videos: the kernel still needs a valid pointer
    packed = torch.zeros((1, d), dtype=torch.uint8, device=packed.device)
  if out is None:
    out = _bf16((b, max_frames, d), packed.device)
  _call("yt8m_frames_unpack_u8", _p(packed), _p(row_offsets), _p(num_frames), b, max_frames, d, 1 if normalize else 0, _p(out), None,
        _stream())
  return out


def split_bf16(x, out_hi=None, out_lo=None, want_lo=True):
  """fp32 [rows, cols] -> bf16 hi, lo (row stride padded to 8 elements, pad columns zero).
  out_hi/out_lo may be column-slices of wider (pre-zeroed) buffers, e.g. to build a concatenation."""
  assert x.dim() == 2 and x.stride(1) == 1
  rows, cols = x.shape
  if out_hi is None:
    out_hi = torch.zeros((rows, pad8(cols)), dtype=torch.bfloat16, device=x.device)[:, :cols]
    out_lo = torch.zeros((rows, pad8(cols)), dtype=torch.bfloat16, device=x.device)[:, :cols] if want_lo else None
  _check(_lib.yt8m_split_bf16(_p(x), rows, cols, x.stride(0), _p(out_hi), _p(out_lo), out_hi.stride(0), _stream()),
         "yt8m_split_bf16")
  return out_hi, out_lo


def pack_transpose(w_kn):
  """fp32 [K, N] (TF layout) -> bf16 [N, pad8(K)] K-contiguous."""
  k, n = w_kn.shape
  out = _bf16((n, pad8(k)), w_kn.device)
  _check(_lib.yt8m_pack_transpose_bf16(_p(w_kn.contiguous()), k, n, _p(out), out.stride(0), _stream()), "yt8m_pack_transpose_bf16")
  return out


_ws_cache = {}


def _workspace(nbytes, device):
  if torch.cuda.is_current_stream_capturing():
    # a captured step owns its workspace (graph-private pool); never share it with eager calls
    return torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
  key = (device.index, torch.cuda.current_stream().cuda_stream)
  buf = _ws_cache.get(key)
  if buf is None or buf.numel() < nbytes:
    buf = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
    _ws_cache[key] = buf
  return buf


FMT_BF16, FMT_F16 = 0, 1
SUMS_FLOATS = 4104                 # include/yt8m_b200.h: YT8M_SUMS_FLOATS


def _fmt(t):
  """Operand format of a 16-bit activation tensor: the torch dtype says which (yt8m_b200.h: YT8M_FMT_*)."""
  if t.dtype == torch.float16:
    return FMT_F16
  if t.dtype == torch.bfloat16:
    return FMT_BF16
  raise Yt8mError("activation operands must be torch.bfloat16 or torch.float16, got %s" % t.dtype)


def linear(a_hi, w_packed, a_lo=None, n=None, k=None, scale=None, shift=None, act=None, out_f32=True, out_bf16=False,
           out_lo=False, out_f16=False):
  """act((A . W^T) * scale + shift).  a_hi/a_lo: bf16 [M, lda] (or a_hi fp16 alone); w_packed: bf16 [N, ldw].
  Returns dict with the requested outputs ('f32', 'hi', 'lo'); out_f16 makes 'hi' an fp16 tensor."""
  m = a_hi.shape[0]
  n = n or w_packed.shape[0]
  k = k or min(a_hi.shape[1], w_packed.shape[1])
  dev = a_hi.device
  ld_out = pad8(n)
  of = _f32((m, ld_out), dev) if out_f32 else None
  if out_f16:
    oh, ol = torch.empty((m, ld_out), dtype=torch.float16, device=dev), None
  else:
    oh = _bf16((m, ld_out), dev) if out_bf16 else None
    ol = _bf16((m, ld_out), dev) if (out_bf16 and out_lo) else None
  if ld_out != n:
    for t in (of, oh, ol):
      if t is not None:
        t.zero_()
  if w_packed.dtype != a_hi.dtype:
    raise Yt8mError("linear: activation (%s) and weight (%s) operands must share one 16-bit format" % (a_hi.dtype, w_packed.dtype))
  ws_bytes = _lib.yt8m_linear_workspace_bytes(m, n, k)
  ws = _workspace(ws_bytes, dev)
  _call("yt8m_linear_fwd", _p(a_hi), _p(a_lo), a_hi.stride(0), _p(w_packed), w_packed.stride(0), m, n, k, _p(scale),
        _p(shift), ACT[act], _fmt(a_hi), FMT_F16 if out_f16 else FMT_BF16, _p(of), _p(oh), _p(ol), ld_out, _p(ws), ws.numel(),
        _stream())
  res = {}
  if of is not None:
    res["f32"] = of[:, :n]
  if oh is not None:
    res["hi"] = oh[:, :n]
  if ol is not None:
    res["lo"] = ol[:, :n]
  return res


def moe_packed_rows(vocab, num_mixtures):
  return _lib.yt8m_moe_packed_rows(vocab, num_mixtures)


def moe_pack(gate_w, expert_w, expert_b, vocab, num_mixtures):
  """TF-layout fp32 gate [D, V(M+1)], expert [D, VM], bias [VM] -> (w_packed bf16 [rows, pad8(D)], bias_packed)."""
  d = gate_w.shape[0]
  rows = moe_packed_rows(vocab, num_mixtures)
  if rows <= 0:
    raise Yt8mError("moe_pack: unsupported vocab=%d mixtures=%d" % (vocab, num_mixtures))
  wp = _bf16((rows, pad8(d)), gate_w.device)
  bp = _f32((rows,), gate_w.device)
  _check(_lib.yt8m_moe_pack_weights(_p(gate_w.contiguous()), _p(expert_w.contiguous()), _p(expert_b.contiguous()), d, vocab,
                                    num_mixtures, _p(wp), wp.stride(0), _p(bp), _stream()), "yt8m_moe_pack_weights")
  return wp, bp


def moe_fwd(x_hi, w_packed, bias_packed, vocab, num_mixtures, x_lo=None, d=None):
  b = x_hi.shape[0]
  d = d or min(x_hi.shape[1], w_packed.shape[1])
  if w_packed.dtype != x_hi.dtype:
    raise Yt8mError("moe_fwd: activation (%s) and weight (%s) operands must share one 16-bit format" % (x_hi.dtype, w_packed.dtype))
  out = _f32((b, vocab), x_hi.device)
  _call("yt8m_moe_fwd", _p(x_hi), _p(x_lo), x_hi.stride(0), _p(w_packed), w_packed.stride(0), _p(bias_packed), b, d, vocab,
        num_mixtures, _fmt(x_hi), _p(out), vocab, _stream())
  return out


def group_max_rows(x, heads):
  rows, cols = x.shape
  groups = rows // heads
  out = _f32((groups, cols), x.device)
  _check(_lib.yt8m_group_max_rows(_p(x.contiguous()), groups, heads, cols, _p(out), _stream()), "yt8m_group_max_rows")
  return out


def lstm_pack(w_tf, b_tf, in_dim, hidden):
  wp = _bf16((4 * hidden, in_dim + hidden), w_tf.device)
  bp = _f32((4 * hidden,), w_tf.device)
  _check(_lib.yt8m_lstm_pack_weights(_p(w_tf.contiguous()), _p(b_tf.contiguous()), in_dim, hidden, _p(wp), _p(bp), _stream()),
         "yt8m_lstm_pack_weights")
  return wp, bp


def lstm_fwd(x, num_frames, w_packed, b_packed, hidden, forget_bias=1.0, want_seq=False, want_seq_bf16=False):
  """x: bf16 [B, T, D]; w_packed/b_packed: per-layer lists.  Returns (state [B, L*2*H], out_seq, out_seq_bf16)."""
  b, t, d = x.shape
  layers = len(w_packed)
  dev = x.device
  state = _f32((b, layers * 2 * hidden), dev)
  seq = _f32((b, t, hidden), dev) if want_seq else None
  seq_bf = _bf16((b, t, hidden), dev) if want_seq_bf16 else None
  ws_bytes = _lib.yt8m_lstm_workspace_bytes(b, t, d, hidden, layers)
  ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
  wp = (c_void_p * layers)(*[w.data_ptr() for w in w_packed])
  bp = (c_void_p * layers)(*[v.data_ptr() for v in b_packed])
  _call("yt8m_lstm_fwd", _p(x), _p(num_frames), b, t, d, hidden, layers, ctypes.cast(wp, c_void_p), ctypes.cast(bp, c_void_p),
        float(forget_bias), _p(state), _p(seq), _p(seq_bf), _p(ws), ws_bytes, _stream())
  return state, seq, seq_bf


def _ptr_array(tensors):
  arr = (c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])
  return arr, ctypes.cast(arr, c_void_p)


def lstm_fwd_train(x, num_frames, w_packed, b_packed, hidden, forget_bias=1.0, want_seq=False):
  """Training forward: as lstm_fwd, additionally returning every layer's output sequence as bf16 (hi, lo) lists
  (what lstm_bwd needs).  Returns (state [B, L*2*H], out_seq or None, seq_hi[L], seq_lo[L])."""
  b, t, d = x.shape
  layers = len(w_packed)
  dev = x.device
  state = _f32((b, layers * 2 * hidden), dev)
  seq = _f32((b, t, hidden), dev) if want_seq else None
  seq_hi = [_bf16((b, t, hidden), dev) for _ in range(layers)]
  seq_lo = [_bf16((b, t, hidden), dev) for _ in range(layers)]
  ws_bytes = _lib.yt8m_lstm_workspace_bytes(b, t, d, hidden, layers)
  ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
  k1, wp = _ptr_array(w_packed)
  k2, bp = _ptr_array(b_packed)
  k3, sh = _ptr_array(seq_hi)
  k4, sl = _ptr_array(seq_lo)
  _call("yt8m_lstm_fwd_train", _p(x), _p(num_frames), b, t, d, hidden, layers, wp, bp, float(forget_bias), _p(state), _p(seq), sh, sl,
        _p(ws), ws_bytes, _stream())
  return state, seq, seq_hi, seq_lo


def lstm_bwd(x, num_frames, w_packed, b_packed, wt_packed, hidden, seq_hi, seq_lo, dstate=None, dout_seq=None, dw=None, db=None,
             forget_bias=1.0):
  """Back-propagation through time.  wt_packed[l] = bf16 transpose [in_l + H, 4H] of w_packed[l]; dstate fp32
  [B, L*2*H] and/or dout_seq fp32 [B, T, H].  Returns (dw[L] fp32 [4H, in_l + H], db[L] fp32 [4H]) in the packed layout."""
  b, t, d = x.shape
  layers = len(w_packed)
  dev = x.device
  if dw is None:
    dw = [_f32((4 * hidden, (d if l == 0 else hidden) + hidden), dev) for l in range(layers)]
  if db is None:
    db = [_f32((4 * hidden,), dev) for _ in range(layers)]
  ws_bytes = _lib.yt8m_lstm_bwd_workspace_bytes(b, t, d, hidden, layers)
  ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
  keep = [_ptr_array(v) for v in (w_packed, b_packed, wt_packed, seq_hi, seq_lo, dw, db)]
  wp, bp, wtp, sh, sl, dwp, dbp = [k[1] for k in keep]
  _call("yt8m_lstm_bwd", _p(x), _p(num_frames), b, t, d, hidden, layers, wp, bp, wtp, float(forget_bias), sh, sl, _p(dstate),
        _p(dout_seq), dwp, dbp, _p(ws), ws_bytes, _stream())
  return dw, db


def attn_pool_bwd(logits, feats, num_frames, heads, mode, dout, want_dfeats=False):
  """Backward of attn_pool: dout fp32 [B, A, F] -> (dlogits fp32 [B, T, A], dfeats fp32 [B, T, F] or None)."""
  b, t, f = feats.shape
  dl = _f32((b, t, heads), feats.device)
  df = _f32((b, t, f), feats.device) if want_dfeats else None
  _call("yt8m_attn_pool_bwd", _p(logits), logits.stride(1), _p(feats), _p(num_frames), b, t, heads, f, mode, _p(dout.contiguous()),
        _p(dl), heads, _p(df), _stream())
  return dl, df


def context_gate_bwd(dy, x, g, scale=None, shift=None, want_bf16=True):
  """Backward of context_gate: returns (dx_direct fp32, dg fp32, dg_hi, dg_lo)."""
  rows, cols = x.shape
  dx, dg = _f32((rows, cols), x.device), _f32((rows, cols), x.device)
  ld = pad8(cols)
  gh = torch.zeros((rows, ld), dtype=torch.bfloat16, device=x.device) if want_bf16 else None
  gl = torch.zeros((rows, ld), dtype=torch.bfloat16, device=x.device) if want_bf16 else None
  _call("yt8m_context_gate_bwd", _p(dy.contiguous()), _p(x.contiguous()), _p(g.contiguous()), _p(scale), _p(shift), rows, cols, _p(dx),
        _p(dg), _p(gh), _p(gl), ld, _stream())
  return dx, dg, (gh[:, :cols] if want_bf16 else None), (gl[:, :cols] if want_bf16 else None)


def group_max_rows_bwd(x, dout, heads):
  """Backward of group_max_rows: x [G*heads, C] (the forward's input), dout [G, C] -> dx [G*heads, C]."""
  rows, cols = x.shape
  dx = _f32((rows, cols), x.device)
  _call("yt8m_group_max_rows_bwd", _p(x.contiguous()), _p(dout.contiguous()), rows // heads, heads, cols, _p(dx), _stream())
  return dx


def l2norm_rows_bwd(x, dy):
  """Backward of l2norm_rows on fp32 [rows, dim]: x = the un-normalised input, dy = dL/dy -> dL/dx."""
  rows, dim = x.shape
  dx = _f32((rows, dim), x.device)
  _call("yt8m_l2norm_rows_bwd", _p(x.contiguous()), _p(dy.contiguous()), rows, dim, _p(dx), _stream())
  return dx


def add_inplace(y, x):
  """y += x (fp32, same shape, contiguous)."""
  assert y.is_contiguous() and x.is_contiguous() and y.numel() == x.numel()
  _call("yt8m_add_inplace", _p(y), _p(x), y.numel(), _stream())
  return y


def attn_pool(logits, feats, num_frames, heads, mode, want_bf16=True):
  """logits fp32 [B, T, >=A]; feats bf16 [B, T, F] -> fp32 [B, A, F] (+ bf16 hi/lo)."""
  b, t, f = feats.shape
  out = _f32((b, heads, f), feats.device)
  oh = _bf16((b, heads, f), feats.device) if want_bf16 else None
  ol = _bf16((b, heads, f), feats.device) if want_bf16 else None
  _check(_lib.yt8m_attn_pool_fwd(_p(logits), logits.stride(1), _p(feats), _p(num_frames), b, t, heads, f, mode, _p(out), _p(oh),
                                 _p(ol), _stream()), "yt8m_attn_pool_fwd")
  return out, oh, ol


def attn_pool_fused(x, w_packed, num_frames, heads, want_bf16=True):
  """x bf16 [B, T, D]; w_packed bf16 [A, >= D] (K-major attention weights) -> fp32 [B, A, D] (+ bf16 hi / lo): logits, masked
  softmax over the frames and the weighted sum in one kernel (num_frames None = the non-zero-frame mask)."""
  b, t, d = x.shape
  out = _f32((b, heads, d), x.device)
  oh = _bf16((b, heads, d), x.device) if want_bf16 else None
  ol = _bf16((b, heads, d), x.device) if want_bf16 else None
  _call("yt8m_attn_pool_fused", _p(x), _p(w_packed), w_packed.stride(0), _p(num_frames), b, t, d, heads, _p(out), _p(oh), _p(ol), _stream())
  return out, oh, ol


def netvlad_fwd(x, num_frames, cw_packed, scale, shift, cw2, want_f32=False, want_lo=False, cw2_split=None, out_f16=False,
                want_stats=False):
  """x bf16 [B, T, D]; cw_packed bf16 [K, D]; cw2 fp32 [D, K] (cw2_split = its (hi, lo) bf16 copies, made on
  demand) -> bf16 hi [B, D*K] (+lo, +fp32); out_f16: one fp16 tensor instead of hi (+lo).  want_stats appends
  the fp32 [B, 2K+1] tensor the backward needs."""
  b, t, d = x.shape
  k = cw_packed.shape[0]
  if cw2_split is None and k == 64:
    cw2_split = split_bf16(cw2.contiguous())
  c2h, c2l = cw2_split if cw2_split is not None else (None, None)
  if out_f16:
    oh, ol = torch.empty((b, d * k), dtype=torch.float16, device=x.device), None
  else:
    oh = _bf16((b, d * k), x.device)
    # the fp32 output is the rescaled STASH (un-normalised descriptor parked in out_hi [+ out_lo]): ask for the lo
    # half whenever fp32 is wanted, or it would only be bf16-accurate
    ol = _bf16((b, d * k), x.device) if (want_lo or want_f32) else None
  of = _f32((b, d * k), x.device) if want_f32 else None
  stats = _f32((b, 2 * k + 1), x.device) if want_stats else None
  _call("yt8m_netvlad_fwd", _p(x), _p(num_frames), b, t, d, k, _p(cw_packed), _p(scale), _p(shift), _p(cw2), _p(c2h), _p(c2l),
        _p(of), _p(oh), _p(ol), d * k, FMT_F16 if out_f16 else FMT_BF16, _p(stats), _stream())
  if want_stats:
    return oh, (ol if want_lo else None), of, stats
  return oh, (ol if want_lo else None), of


def netvlad_tiled_supported(t, d, k):
  return bool(_lib.yt8m_netvlad_tiled_supported(t, d, k))


_tiled_index_cache = {}


def netvlad_tiled_index(d, k, width, device):
  """int64 [D*K]: idx[tiled position] = row-major position d*K + k for the blocked layouts of yt8m_netvlad_fwd_tiled
  (width 8: the 16-bit descriptor; width 4: fp32 cw2).  `tensor.reshape(-1)[idx]` (or `rows[idx]` for the next layer's
  weight) produces the tiled order -- an index gather, no arithmetic."""
  key = (d, k, width, str(device))
  if key not in _tiled_index_cache:
    dd = torch.arange(d).unsqueeze(1)
    kk = torch.arange(k).unsqueeze(0)
    pos = ((dd // 32) * (k // width) + kk // width) * (32 * width) + (dd % 32) * width + kk % width
    idx = torch.empty(d * k, dtype=torch.int64)
    idx[pos.reshape(-1)] = (dd * k + kk).reshape(-1)
    _tiled_index_cache[key] = idx.to(device)
  return _tiled_index_cache[key]


def netvlad_fwd_tiled(x, num_frames, cw_packed, scale, shift, cw2_tiled, out_f16=True, want_stats=False, two_kernels=False):
  """NetVLAD with the blocked layouts (see include/yt8m_b200.h): returns the TILED descriptor [B, D*K] (fp16 or bf16)
  (+ stats).  two_kernels: pass the assignment scratch, which selects the assignment + aggregation kernel pair (measured
  slower than the one-pass kernel as soon as the batch's frames outgrow L2: it reads them twice)."""
  b, t, d = x.shape
  k = cw_packed.shape[0]
  out = torch.empty((b, d * k), dtype=torch.float16 if out_f16 else torch.bfloat16, device=x.device)
  stats = _f32((b, 2 * k + 1), x.device) if want_stats else None
  ws_bytes = _lib.yt8m_netvlad_tiled_workspace_bytes(b, t, d, k) if two_kernels else 0
  ws = _workspace(ws_bytes, x.device) if ws_bytes else None
  _call("yt8m_netvlad_fwd_tiled", _p(x), _p(num_frames), b, t, d, k, _p(cw_packed), _p(scale), _p(shift), _p(cw2_tiled), _p(out),
        FMT_F16 if out_f16 else FMT_BF16, _p(stats), _p(ws), ws.numel() if ws is not None else 0, _stream())
  return (out, stats) if want_stats else out


def netvlad_bwd_norm(dy, y, stats, cw2, want_dcw2=True, want_split=False):
  """dy, y fp32 [B, D*K]; stats [B, 2K+1]; cw2 fp32 [D, K] -> (dv [B, D*K], dasum [B, K], dcw2 [D, K] or None)
  (+ (dv_hi, dv_lo) bf16 [B, D*K] when want_split: the operands of netvlad_bwd_assign_fused)."""
  d, k = cw2.shape
  b = dy.shape[0]
  dv = _f32((b, d * k), dy.device)
  dasum = _f32((b, k), dy.device)
  dcw2 = _f32((d, k), dy.device) if want_dcw2 else None
  hi = _bf16((b, d * k), dy.device) if want_split else None
  lo = _bf16((b, d * k), dy.device) if want_split else None
  _check(_lib.yt8m_netvlad_bwd_norm(_p(dy.contiguous()), _p(y.contiguous()), _p(stats), _p(cw2.contiguous()), b, d, k, _p(dv),
                                    _p(dasum), _p(dcw2), _p(hi), _p(lo), _stream()), "yt8m_netvlad_bwd_norm")
  return (dv, dasum, dcw2, (hi, lo)) if want_split else (dv, dasum, dcw2)


def netvlad_bwd_assign_fused_supported(t, d, k):
  return bool(_lib.yt8m_netvlad_bwd_assign_fused_supported(t, d, k))


def netvlad_bwd_assign_fused(x, num_frames, cw_packed, scale, shift, dv_split, dasum, want_dshift=True):
  """The assignment backward in one tcgen05 kernel (logits recomputed on chip): x bf16 [B, T, D]; cw_packed bf16 [K, >=D];
  scale / shift [K] or None; dv_split = (hi, lo) bf16 [B, D*K]; dasum [B, K] -> (dzs_hi, dzs_lo bf16 [B*T, K], dshift)."""
  b, t, d = x.shape
  k = dasum.shape[1]
  hi, lo = _bf16((b * t, k), x.device), _bf16((b * t, k), x.device)
  dshift = _f32((k,), x.device) if want_dshift else None
  _check(_lib.yt8m_netvlad_bwd_assign_fused(_p(x), _p(num_frames), _p(cw_packed), cw_packed.stride(0), _p(scale), _p(shift),
                                            _p(dv_split[0]), _p(dv_split[1]), _p(dasum), b, t, d, k, _p(hi), _p(lo), _p(dshift),
                                            _stream()), "yt8m_netvlad_bwd_assign_fused")
  return hi, lo, dshift


def netvlad_bwd_assign(x, num_frames, z, dv, dasum, scale=None, want_dshift=True):
  """x bf16 [B, T, D]; z fp32 [B*T, K]; dv [B, D*K]; dasum [B, K] -> (dzs_hi, dzs_lo bf16 [B*T, K], dshift [K] or None)."""
  b, t, d = x.shape
  k = dasum.shape[1]
  ws = _f32((b * t, k), x.device)
  hi, lo = _bf16((b * t, k), x.device), _bf16((b * t, k), x.device)
  dshift = _f32((k,), x.device) if want_dshift else None
  _check(_lib.yt8m_netvlad_bwd_assign(_p(x), _p(num_frames), _p(z.contiguous()), _p(dv), _p(dasum), _p(scale), b, t, d, k, _p(ws),
                                      _p(hi), _p(lo), _p(dshift), _stream()), "yt8m_netvlad_bwd_assign")
  return hi, lo, dshift


def act_bwd(dy, y, act=None, col_scale=None):
  """d_pre = dy * act'(y) * col_scale as bf16 (hi, lo) [rows, pad8(cols)] views."""
  rows, cols = dy.shape
  ld = pad8(cols)
  hi, lo = _bf16((rows, ld), dy.device), _bf16((rows, ld), dy.device)
  if ld != cols:
    hi.zero_()
    lo.zero_()
  _check(_lib.yt8m_act_bwd(_p(dy.contiguous()), _p(y.contiguous()), rows, cols, ACT[act], _p(col_scale), _p(hi), _p(lo), ld,
                           _stream()), "yt8m_act_bwd")
  return hi[:, :cols], lo[:, :cols]


BN_EPS, BN_DECAY = 1e-3, 0.999       # slim.batch_norm defaults


def _src(t):
  return {torch.float32: SRC_F32, torch.bfloat16: SRC_BF16}[t.dtype]


def bn_train_fwd(x, gamma, beta, moving_mean=None, moving_var=None, act=None, want_f32=True, want_bf16=False):
  """slim.batch_norm(is_training=True) on x [rows, cols] (fp32 or bf16, row stride >= cols) followed by `act`: batch statistics,
  folded affine + activation, and the in-place moving-average update.  Returns (out dict f32 / hi / lo, (mean, var))."""
  rows, cols = x.shape
  dev = x.device
  mean, var = _f32((cols,), dev), _f32((cols,), dev)
  _call("yt8m_bn_stats", _p(x), _src(x), rows, cols, x.stride(0), _p(mean), _p(var), _stream())
  scale, shift = _f32((cols,), dev), _f32((cols,), dev)
  _call("yt8m_bn_fold", _p(gamma), _p(beta), _p(mean), _p(var), BN_EPS, cols, _p(scale), _p(shift), _p(moving_mean), _p(moving_var), BN_DECAY,
        _stream())
  ld = pad8(cols)
  of = _f32((rows, ld), dev) if want_f32 else None
  oh = torch.zeros((rows, ld), dtype=torch.bfloat16, device=dev) if want_bf16 else None
  ol = torch.zeros((rows, ld), dtype=torch.bfloat16, device=dev) if want_bf16 else None
  _call("yt8m_col_affine_act", _p(x), _src(x), rows, cols, x.stride(0), _p(scale), _p(shift), ACT[act], _p(of), _p(oh), _p(ol), ld, _stream())
  out = {}
  if of is not None:
    out["f32"] = of[:, :cols]
  if oh is not None:
    out["hi"], out["lo"] = oh[:, :cols], ol[:, :cols]
  return out, (mean, var)


def bn_moving_update(moving_mean, moving_var, mean, var):
  """moving <- decay * moving + (1 - decay) * batch, in place (the UPDATE_OPS of slim.batch_norm)."""
  cols = mean.shape[0]
  scratch = _f32((2, cols), mean.device)
  _call("yt8m_bn_fold", None, None, _p(mean), _p(var), BN_EPS, cols, _p(scratch[0]), _p(scratch[1]), _p(moving_mean), _p(moving_var),
        BN_DECAY, _stream())


def bn_train_bwd(dy, y, x, stats, gamma, act=None, want_dx=True):
  """Backward of bn_train_fwd: dy, y fp32 [rows, cols] (y = the forward's fp32 output; ignored when act is None), x = the forward's
  input.  Returns (dgamma, dbeta, dx_hi, dx_lo, dx_f32) -- dx as bf16 hi/lo operands AND fp32 (None when not wanted)."""
  rows, cols = x.shape
  dev = x.device
  mean, var = stats
  dgamma, dbeta = _f32((cols,), dev), _f32((cols,), dev)
  ld = pad8(cols)
  dxf = _f32((rows, ld), dev) if want_dx else None
  dh = torch.zeros((rows, ld), dtype=torch.bfloat16, device=dev) if want_dx else None
  dl = torch.zeros((rows, ld), dtype=torch.bfloat16, device=dev) if want_dx else None
  yy = y if ACT[act] != 0 else None
  _call("yt8m_bn_bwd", _p(dy), dy.stride(0), _p(yy), yy.stride(0) if yy is not None else 0, _p(x), _src(x), x.stride(0), _p(mean), _p(var),
        BN_EPS, _p(gamma), ACT[act], rows, cols, _p(dgamma), _p(dbeta), _p(dxf), _p(dh), _p(dl), ld, _stream())
  if not want_dx:
    return dgamma, dbeta, None, None, None
  return dgamma, dbeta, dh[:, :cols], dl[:, :cols], dxf[:, :cols]


def debug_set_timeline(buf):
  """buf: int64 CUDA tensor with >= 128 elements, or None to switch the stamps off."""
  _check(_lib.yt8m_debug_set_timeline(_p(buf)), "yt8m_debug_set_timeline")


def debug_set_flags(flags):
  _check(_lib.yt8m_debug_set_flags(int(flags)), "yt8m_debug_set_flags")


def context_gate(x, g, scale=None, shift=None, want_bf16=True):
  rows, cols = x.shape
  out = _f32((rows, cols), x.device)
  oh = _bf16((rows, cols), x.device) if want_bf16 else None
  ol = _bf16((rows, cols), x.device) if want_bf16 else None
  _check(_lib.yt8m_context_gate_fwd(_p(x.contiguous()), _p(g.contiguous()), _p(scale), _p(shift), rows, cols, _p(out), _p(oh),
                                    _p(ol), _stream()), "yt8m_context_gate_fwd")
  return out, oh, ol


def col_affine(x, scale, shift, want_f32=False):
  """x [rows, cols] fp32/bf16 contiguous (cols % 8 == 0) -> (hi, lo[, f32]) of x * scale + shift."""
  assert x.is_contiguous() and x.shape[1] % 8 == 0
  rows, cols = x.shape
  src = SRC_F32 if x.dtype == torch.float32 else SRC_BF16
  hi, lo = _bf16((rows, cols), x.device), _bf16((rows, cols), x.device)
  of = _f32((rows, cols), x.device) if want_f32 else None
  _check(_lib.yt8m_col_affine(_p(x), src, rows, cols, _p(scale), _p(shift), _p(of), _p(hi), _p(lo), _stream()), "yt8m_col_affine")
  return (hi, lo, of) if want_f32 else (hi, lo)


def xent(pred, labels, want_grad=False, grad_scale=1.0):
  b, v = pred.shape
  loss = _f32((1,), pred.device)
  dp = _f32((b, v), pred.device) if want_grad else None
  _check(_lib.yt8m_xent_fwd_bwd(_p(pred.contiguous()), _p(labels.contiguous()), b, v, _p(loss), _p(dp), float(grad_scale),
                                _stream()), "yt8m_xent_fwd_bwd")
  return loss, dp


def topk_rows(x, k):
  rows, cols = x.shape
  idx = torch.empty((rows, k), dtype=torch.int32, device=x.device)
  val = _f32((rows, k), x.device)
  _check(_lib.yt8m_topk_rows(_p(x.contiguous()), rows, cols, k, _p(idx), _p(val), _stream()), "yt8m_topk_rows")
  return idx, val


# ------------------------------------------------------------------------------------------------
# training-step wrappers
# ------------------------------------------------------------------------------------------------

def logistic_bwd_dz(dp, p):
  b, v = p.shape
  ld = pad8(v)
  hi = torch.zeros((b, ld), dtype=torch.bfloat16, device=p.device)
  lo = torch.zeros((b, ld), dtype=torch.bfloat16, device=p.device)
  _check(_lib.yt8m_logistic_bwd_dz(_p(dp.contiguous()), _p(p.contiguous()), b, v, _p(hi), _p(lo), ld, _stream()),
         "yt8m_logistic_bwd_dz")
  return hi, lo


def wgrad(a_hi, a_lo, b, m, n, out=None):
  """out[m, n] = A^T . B with A [Kb, >=m] (hi/lo) and B [Kb, >=n] stored batch-major."""
  kb = a_hi.shape[0]
  if out is None:
    out = _f32((m, n), a_hi.device)
  _check(_lib.yt8m_wgrad(_p(a_hi), _p(a_lo), a_hi.stride(0), _p(b), b.stride(0), m, n, kb, _p(out), out.stride(0), _stream()),
         "yt8m_wgrad")
  return out


def dgrad(dz_hi, dz_lo, w_rows, n):
  """dx[B, n] fp32 = dz[B, rows] . W[rows, :n] with W in the layout the FORWARD uses (bf16 [rows, ld >= n], the output
  features of the layer as rows).  The contraction runs over the rows of both stored operands once dz is transposed (a few
  MB), so this is the MN-major GEMM of yt8m_wgrad -- no transposed copy of the weight matrix (the first version rebuilt a
  bf16 W^T from the fp32 master every step: 160 us for the 302 MB hidden layer of BASELINE config 2)."""
  b, rows = dz_hi.shape[0], w_rows.shape[0]
  bp = pad8(b)

  def tr(z):
    if z is None:
      return None
    t = torch.zeros((rows, bp), dtype=torch.bfloat16, device=z.device) if bp != b else torch.empty((rows, bp), dtype=torch.bfloat16,
                                                                                                    device=z.device)
    t[:, :b] = z[:, :rows].t()
    return t

  return wgrad(tr(dz_hi), tr(dz_lo), w_rows, b, n)


def colsum_bf16(hi, lo, cols, out=None):
  if out is None:
    out = _f32((cols,), hi.device)
  _check(_lib.yt8m_colsum_bf16(_p(hi), _p(lo), hi.stride(0), hi.shape[0], cols, _p(out), _stream()), "yt8m_colsum_bf16")
  return out


def moe_bwd_dlogits(x_hi, x_lo, w_packed, bias_packed, dp, vocab, num_mixtures, d=None):
  b = x_hi.shape[0]
  d = d or min(x_hi.shape[1], w_packed.shape[1])
  rows = w_packed.shape[0]
  hi = _bf16((b, rows), x_hi.device)
  lo = _bf16((b, rows), x_hi.device)
  _check(_lib.yt8m_moe_bwd_dlogits(_p(x_hi), _p(x_lo), x_hi.stride(0), _p(w_packed), w_packed.stride(0), _p(bias_packed), _p(dp),
                                   dp.stride(0), b, d, vocab, num_mixtures, _p(hi), _p(lo), rows, _stream()), "yt8m_moe_bwd_dlogits")
  return hi, lo


def grad_reg_sumsq(grad, param, l2, moe_per=0, moe_nmix=0):
  rows, row_len = (param.shape[0], param.shape[1]) if param.dim() == 2 else (1, param.numel())
  sums = _f32((SUMS_FLOATS,), param.device)     # [0..3] results, the rest: scratch of the deterministic reduction
  _check(_lib.yt8m_grad_reg_sumsq(_p(grad), _p(param), rows, row_len, float(l2), moe_per, moe_nmix, _p(sums), _stream()),
         "yt8m_grad_reg_sumsq")
  return sums


def clip_adam_step(param, grad, m, v, sums, clip, lr_t, beta1=0.9, beta2=0.999, eps=1e-8, moe_per=0, moe_nmix=0, only_segment=-1,
                   param_bf16=None):
  rows, row_len = (param.shape[0], param.shape[1]) if param.dim() == 2 else (1, param.numel())
  _check(_lib.yt8m_clip_adam_step(_p(param), _p(grad), _p(m), _p(v), rows, row_len, _p(sums), float(clip), float(lr_t), float(beta1),
                                  float(beta2), float(eps), moe_per, moe_nmix, only_segment, _p(param_bf16), _stream()),
         "yt8m_clip_adam_step")
