"""A numpy stand-in for the few TensorFlow-1.0 / slim ops that the reference's hot-path model files call, so that the
REFERENCE'S OWN create_model() / calculate_loss() code can be executed in the build container (Python 3, no TensorFlow)
to generate golden vectors (oracle/make_model_golden.py -> tests/golden/model_golden.json).

Test infrastructure only (like everything under oracle/).  What is the reference's and what is ours:
  * the graph-building code that runs is the reference's, read from /root/reference and exec'd unchanged except for the
    py2 `print x` statements (and `xrange`), which do not touch arithmetic;
  * the ops below follow TF-1.0's documented semantics: slim.fully_connected = act(x . W + b) with ReLU as the
    DEFAULT activation and "biases_initializer=None => no bias", tf.nn.softmax over `dim`, tf.nn.l2_normalize with
    epsilon 1e-12, tf.sequence_mask, tf.einsum, and BasicLSTMCell / MultiRNNCell / dynamic_rnn(sequence_length) -- gate
    order i, j, f, o, forget_bias added to f, zero output and frozen state past the sequence end.  The LSTM here is a second,
    numpy restatement of that published algorithm, written independently of oracle/yt8m_oracle.py.
Variables are not initialised randomly: the harness puts every weight into `STORE` under the TF variable name."""
import collections
import contextlib
import sys
import types

import numpy as np

STORE = {}          # TF variable name -> numpy array, filled by the harness
SCOPES = []         # active tf.variable_scope names


class T(np.ndarray):
  """ndarray with the two Tensor methods the reference calls."""

  def get_shape(self):
    shape = self.shape

    class _S(object):
      def as_list(self_inner):
        return list(shape)
    return _S()

  def set_shape(self, shape):
    assert list(self.shape) == list(shape), (self.shape, shape)


def t(x, dtype=np.float32):
  return np.asarray(x, dtype=dtype).view(T)


def _var(name):
  full = "/".join(SCOPES + [name])
  if full not in STORE:
    raise KeyError("the reference asked for variable %r; the harness did not provide it" % full)
  return np.asarray(STORE[full], dtype=np.float32)


# ---- tf.nn ------------------------------------------------------------------------------------------------------
def _softmax(x, dim=-1, name=None):
  x = np.asarray(x, dtype=np.float32)
  e = np.exp(x - x.max(axis=dim, keepdims=True))
  return t(e / e.sum(axis=dim, keepdims=True))


def _sigmoid(x, name=None):
  return t(1.0 / (1.0 + np.exp(-np.asarray(x, dtype=np.float32))))


def _l2_normalize(x, dim, epsilon=1e-12, name=None):
  x = np.asarray(x, dtype=np.float32)
  ss = (x * x).sum(axis=dim, keepdims=True)
  return t(x / np.sqrt(np.maximum(ss, epsilon)))


def _dynamic_rnn(cell, inputs, sequence_length=None, swap_memory=False, dtype=None, **unused):
  """tf.nn.dynamic_rnn(MultiRNNCell[BasicLSTMCell]): returns (outputs of the top layer, per-layer (c, h) states)."""
  x = np.asarray(inputs, dtype=np.float32)
  bsz, t_max, _ = x.shape
  n = np.full(bsz, t_max) if sequence_length is None else np.asarray(sequence_length).astype(np.int64)
  cells = cell.cells
  cs = [np.zeros((bsz, c.num_units), np.float32) for c in cells]
  hs = [np.zeros((bsz, c.num_units), np.float32) for c in cells]
  outs = np.zeros((bsz, t_max, cells[-1].num_units), np.float32)
  for step in range(t_max):
    live = (step < n)[:, None]
    inp = x[:, step]
    for l, c in enumerate(cells):
      w = _var("multi_rnn_cell/cell_%d/basic_lstm_cell/weights" % l)
      b = _var("multi_rnn_cell/cell_%d/basic_lstm_cell/biases" % l)
      g = np.concatenate([inp, hs[l]], axis=1) @ w + b
      i, j, f, o = np.split(g, 4, axis=1)
      sig = lambda v: 1.0 / (1.0 + np.exp(-v))
      c2 = cs[l] * sig(f + c.forget_bias) + sig(i) * np.tanh(j)
      h2 = np.tanh(c2) * sig(o)
      cs[l] = np.where(live, c2, cs[l])
      hs[l] = np.where(live, h2, hs[l])
      inp = h2
    outs[:, step] = np.where(live, inp, 0.0)
  # state layout (TF 1.0): a tuple cell returns LSTMStateTuple(c, h); a non-tuple BasicLSTMCell returns concat([c, h], 1)
  # and a non-tuple MultiRNNCell concatenates its layers' states along the columns
  per_layer = [LSTMStateTuple(t(c), t(h)) if cl.state_is_tuple else t(np.concatenate([c, h], axis=1))
               for cl, c, h in zip(cells, cs, hs)]
  if cell.state_is_tuple:
    return t(outs), tuple(per_layer)
  return t(outs), t(np.concatenate([np.asarray(s_) for s_ in per_layer], axis=1))


LSTMStateTuple = collections.namedtuple("LSTMStateTuple", ("c", "h"))


class _BasicLSTMCell(object):
  def __init__(self, num_units, forget_bias=1.0, state_is_tuple=True, **unused):
    self.num_units, self.forget_bias, self.state_is_tuple = int(num_units), float(forget_bias), bool(state_is_tuple)


class _MultiRNNCell(object):
  def __init__(self, cells, state_is_tuple=True):
    self.cells, self.state_is_tuple = list(cells), bool(state_is_tuple)
    if not state_is_tuple and any(c.state_is_tuple for c in self.cells):
      raise ValueError("MultiRNNCell(state_is_tuple=False) over tuple cells is an error in TF 1.0 as well")


# ---- slim -------------------------------------------------------------------------------------------------------
def _relu(x, name=None):
  return t(np.maximum(np.asarray(x, dtype=np.float32), 0.0))


_DEFAULT = object()


def _fully_connected(inputs, num_outputs, activation_fn=_DEFAULT, biases_initializer=_DEFAULT, weights_regularizer=None,
                     weights_initializer=None, scope=None, **unused):
  scope = scope or "fully_connected"
  x = np.asarray(inputs, dtype=np.float32)
  w = _var(scope + "/weights")
  assert w.shape == (x.shape[-1], num_outputs), (scope, w.shape, x.shape, num_outputs)
  y = x @ w
  if biases_initializer is not None:
    y = y + _var(scope + "/biases")
  act = _relu if activation_fn is _DEFAULT else activation_fn        # slim's default activation is ReLU
  return t(act(y)) if act is not None else t(y)


@contextlib.contextmanager
def _scope(name, *args, **kw):
  SCOPES.append(name)
  try:
    yield
  finally:
    SCOPES.pop()


@contextlib.contextmanager
def _name_scope(name, *args, **kw):
  yield


def _reduce(fn):
  def op(x, axis=None, keep_dims=False, name=None, reduction_indices=None):
    if axis is None:
      axis = reduction_indices
    return t(fn(np.asarray(x, dtype=np.float32), axis=axis if axis is None or isinstance(axis, int) else tuple(axis), keepdims=keep_dims))
  return op


def _sequence_mask(lengths, maxlen=None, dtype=np.float32, name=None):
  lengths = np.asarray(lengths).astype(np.int64)
  return t((np.arange(maxlen)[None, :] < lengths[:, None]).astype(np.float32))


def install(flag_values):
  """Puts stub modules `tensorflow`, `tensorflow.contrib.slim`, ... into sys.modules; returns the tf stub."""
  tf = types.ModuleType("tensorflow")
  tf.float32, tf.int32, tf.int64 = np.float32, np.int32, np.int64
  nn = types.ModuleType("tensorflow.nn")
  nn.softmax, nn.sigmoid, nn.relu, nn.l2_normalize, nn.dynamic_rnn = _softmax, _sigmoid, _relu, _l2_normalize, _dynamic_rnn
  nn.elu = lambda x, name=None: t(np.where(np.asarray(x) > 0, x, np.exp(np.minimum(np.asarray(x, dtype=np.float32), 0)) - 1))
  nn.embedding_lookup = lambda params, ids, **kw: t(np.asarray(params)[np.asarray(ids).astype(np.int64)])
  tf.nn = nn
  tf.reshape = lambda x, shape, name=None: t(np.reshape(np.asarray(x, dtype=np.float32), shape))
  tf.reduce_sum, tf.reduce_mean, tf.reduce_max = _reduce(np.sum), _reduce(np.mean), _reduce(np.max)
  tf.concat = lambda values, axis, name=None: t(np.concatenate([np.asarray(v, dtype=np.float32) for v in values], axis=axis))
  tf.cast = lambda x, dtype, name=None: t(np.asarray(x).astype(np.float32)) if np.dtype(dtype) == np.float32 else np.asarray(x).astype(dtype)
  tf.log = lambda x, name=None: t(np.log(np.asarray(x, dtype=np.float32)))
  tf.negative = lambda x, name=None: t(-np.asarray(x, dtype=np.float32))
  tf.square = lambda x, name=None: t(np.square(np.asarray(x, dtype=np.float32)))
  tf.maximum = lambda a, b, name=None: (max(a, b) if isinstance(a, int) and isinstance(b, int) else t(np.maximum(a, b)))
  tf.expand_dims = lambda x, axis=None, name=None, dim=None: np.expand_dims(np.asarray(x), axis if axis is not None else dim).view(T)   # TF 1.0 also takes dim=
  tf.einsum = lambda eq, *ops: t(np.einsum(eq, *[np.asarray(o, dtype=np.float32) for o in ops]))
  tf.sequence_mask = _sequence_mask
  tf.name_scope, tf.variable_scope = _name_scope, _scope
  tf.constant_initializer = lambda value, **kw: np.asarray(value, dtype=np.float32)
  # tf.get_variable: a constant initialiser IS the value (lookup tables); a random initialiser means "trainable variable": the
  # harness supplies its value under scope/name
  def _get_variable(name, shape=None, dtype=None, trainable=True, initializer=None, **kw):
    if isinstance(initializer, tuple) and initializer and initializer[0] == "initial value":
      v = _var(name)
      assert tuple(v.shape) == tuple(shape), (name, v.shape, shape)
      return t(v)
    return t(np.reshape(initializer, shape))
  tf.get_variable = _get_variable
  tf.truncated_normal_initializer = lambda mean=0.0, stddev=1.0, **kw: ("initial value", None)
  # ops used by CnnDeepCombineChainModel / LstmParallelFinaloutputModel (wh/all_frame_models/cnn_deep_combine_chain_model.py:24-41,
  # :105-116; lstm_parallel_finaloutput_model.py:36)
  tf.pad = lambda x, paddings, name=None: t(np.pad(np.asarray(x, dtype=np.float32), [tuple(int(v) for v in p_) for p_ in paddings]))
  tf.split = lambda x, sizes, axis=0, name=None: [t(a) for a in np.split(np.asarray(x, dtype=np.float32), np.cumsum(sizes)[:-1], axis=axis)]
  nn.embedding_lookup = lambda params, ids, name=None: t(np.asarray(params)[np.asarray(ids).astype(np.int64)])
  # ops used by zt's AttentionModel (zt/frame_level_models.py:4372-4398)
  tf.abs = lambda x, name=None: t(np.abs(np.asarray(x, dtype=np.float32)))
  tf.shape = lambda x, name=None: tuple(np.asarray(x).shape)
  tf.ones = lambda shape, dtype=np.float32, name=None: t(np.ones(shape, np.float32))
  tf.zeros = lambda shape, dtype=np.float32, name=None: t(np.zeros(shape, np.float32))
  tf.greater = lambda a, b, name=None: np.asarray(a) > np.asarray(b)
  tf.where = lambda c, a, b, name=None: t(np.where(np.asarray(c), a, b))
  tf.tile = lambda x, multiples, name=None: t(np.tile(np.asarray(x, dtype=np.float32), multiples))
  tf.truncated_normal = lambda shape, stddev=1.0, **kw: ("initial value", tuple(shape))
  tf.constant = lambda value, shape=None, dtype=None, name=None: ("initial value", shape) if shape is not None else t(value)
  def _variable(initial_value, name=None, **kw):
    if name is None:                      # unnamed variables (wh/all_frame_models/dbof_model.py:73-113): creation order
      return t(STORE["__unnamed__"].pop(0))
    return t(_var(name))                  # the harness supplies the value under scope/name
  tf.Variable = _variable
  # ops used by DbofModel + model_utils (wh/model_utils.py:56-94, wh/all_frame_models/dbof_model.py:62-123)
  tf.random_normal = lambda shape, mean=0.0, stddev=1.0, **kw: ("initial value", tuple(shape))
  tf.random_uniform = lambda shape, **kw: t(np.reshape(STORE["__random_uniform__"], shape))    # pinned by the harness
  tf.range = lambda n, name=None: np.arange(n)
  tf.multiply = lambda a, b, name=None: t(np.asarray(a, dtype=np.float32) * np.asarray(b, dtype=np.float32))
  tf.minimum = lambda a, b, name=None: (min(a, b) if isinstance(a, int) and isinstance(b, int) else np.minimum(a, b))
  tf.stack = lambda values, axis=0, name=None: np.stack([np.asarray(v) for v in values], axis=axis)
  tf.gather_nd = lambda params, indices, name=None: t(np.asarray(params)[tuple(np.moveaxis(np.asarray(indices).astype(np.int64), -1, 0))])
  tf.matmul = lambda a, b, name=None: t(np.asarray(a, dtype=np.float32) @ np.asarray(b, dtype=np.float32))
  tf.summary = types.SimpleNamespace(histogram=lambda *a, **k: None, scalar=lambda *a, **k: None)
  nn.relu6 = lambda x, name=None: t(np.clip(np.asarray(x, dtype=np.float32), 0.0, 6.0))
  tf.add_to_collection = lambda name, value: None
  tf.GraphKeys = types.SimpleNamespace(REGULARIZATION_LOSSES="regularization_losses")
  nn.l2_loss = lambda x, name=None: float((np.asarray(x, dtype=np.float32) ** 2).sum() / 2)
  nn.xw_plus_b = lambda x, w, b, name=None: t(np.asarray(x, dtype=np.float32) @ np.asarray(w, dtype=np.float32) + np.asarray(b, dtype=np.float32))
  # ops used by the frame reader (wh/readers.py:21-56 resize_axis, :159-186 get_video_matrix)
  tf.uint8 = np.uint8
  tf.decode_raw = lambda strings, dtype, name=None: np.stack([np.frombuffer(bytes(b_), dtype=dtype) for b_ in strings])
  tf.convert_to_tensor = lambda x, **kw: t(x)
  tf.unstack = lambda x, **kw: [int(v) for v in x]
  tf.zeros_like = lambda x, **kw: np.zeros_like(np.asarray(x))
  # tf.slice: size -1 = "everything from begin to the end of that dimension"
  tf.slice = lambda x, begin, size, name=None: t(np.asarray(x)[tuple(slice(int(b_), None if int(n_) < 0 else int(b_) + int(n_))
                                                                        for b_, n_ in zip(begin, size))])
  tf.fill = lambda dims, value, name=None: t(np.full([int(d_) for d_ in dims], value, dtype=np.float32))
  contrib = types.ModuleType("tensorflow.contrib")
  rnn = types.ModuleType("tensorflow.contrib.rnn")
  rnn.BasicLSTMCell, rnn.MultiRNNCell = _BasicLSTMCell, _MultiRNNCell
  slim = types.ModuleType("tensorflow.contrib.slim")
  slim.fully_connected = _fully_connected
  slim.l2_regularizer = lambda scale: ("l2", scale)

  def _batch_norm(inputs, center=True, scale=True, is_training=True, scope=None, decay=0.999, epsilon=0.001, **kw):
    """slim.batch_norm in inference form (moving statistics), epsilon = 0.001."""
    if is_training:
      raise NotImplementedError("the harness runs slim.batch_norm with is_training=False only")
    x = np.asarray(inputs, dtype=np.float32)
    y = (x - _var(scope + "/moving_mean")) / np.sqrt(_var(scope + "/moving_variance") + epsilon)
    if scale:
      y = y * _var(scope + "/gamma")
    if center:
      y = y + _var(scope + "/beta")
    return t(y)
  slim.batch_norm = _batch_norm
  contrib.rnn, contrib.slim = rnn, slim
  contrib.layers = types.SimpleNamespace(l2_regularizer=lambda scale: ("l2", scale))     # cnn_deep_combine_chain_model.py:36
  tf.contrib = contrib
  flags = types.ModuleType("tensorflow.flags")
  flags.FLAGS = flag_values
  for kind in ("string", "integer", "float", "bool", "boolean"):
    setattr(flags, "DEFINE_" + kind, lambda name, default, help=None: (hasattr(flag_values, name) or setattr(flag_values, name, default)))
  tf.flags = flags
  import logging as _logging
  tf.logging = _logging                                  # `from tensorflow import logging` (wh/utils.py:20)
  for name, mod in (("tensorflow", tf), ("tensorflow.nn", nn), ("tensorflow.contrib", contrib), ("tensorflow.contrib.rnn", rnn),
                    ("tensorflow.contrib.slim", slim), ("tensorflow.flags", flags)):
    sys.modules[name] = mod
  return tf
