"""Video-level model plugins with the reference's surface (wh/video_level_models.py:16-52): flags at
import time, classes looked up by name, ``create_model(model_input, vocab_size, **unused_params)``
returning ``{"predictions": [batch, vocab_size] probabilities, ...}``.

The arithmetic runs in hand-written sm_100a kernels behind the C ABI (libyt8m_b200.so); there is no
fallback.  ``model_input`` is a [rows, features] CUDA tensor (fp32 or bf16) or a ``yt8m_ops.Act``.
"""
import yt8m_flags as flags
import models
import yt8m_ops as ops
import yt8m_native as nat

FLAGS = flags.FLAGS

flags.DEFINE_integer(
    "moe_num_mixtures", 2,
    "The number of mixtures (excluding the dummy 'expert') used for MoeModel.")
flags.DEFINE_integer("deep_chain_layers", 3, "The number of layers used for DeepChainModel")
flags.DEFINE_integer("deep_chain_relu_cells", 200, "The number of relu cells used for DeepChainModel")
flags.DEFINE_string("deep_chain_relu_type", "relu",
                    "The type of relu cells used for DeepChainModel (options are elu and relu)")
flags.DEFINE_bool("deep_chain_use_length", False, "The number of relu cells used for DeepChainModel")
flags.DEFINE_integer("num_supports", 25, "Number of total support categories.")
flags.DEFINE_integer("moe_num_extend", 8, "The number of attention outputs, used for MoeExtendModel.")

# flags of reference video-level models outside SURVEY.md §8 (wh/video_level_models.py:19-46): accepted so that command lines parse
flags.DEFINE_integer("divergence_model_count", 8, "reference flag of a model / loss outside SURVEY.md §8 (accepted, unused)")
flags.DEFINE_integer("hidden_chain_layers", 4, "reference flag of a model / loss outside SURVEY.md §8 (accepted, unused)")
flags.DEFINE_integer("hidden_chain_relu_cells", 256, "reference flag of a model / loss outside SURVEY.md §8 (accepted, unused)")


class LogisticModel(models.BaseModel):
  """Logistic model with L2 regularization (wh/all_video_models/logistic_model.py:9-26)."""

  def create_model(self, model_input, vocab_size, l2_penalty=1e-8, original_input=None, **unused_params):
    out = ops.fully_connected(model_input, vocab_size, "fully_connected", activation_fn="sigmoid",
                              l2_penalty=l2_penalty, want_bf16=False)
    return {"predictions": out.f32}


class MoeModel(models.BaseModel):
  """A softmax over a mixture of logistic models (wh/all_video_models/moe_model.py:9-65)."""

  def create_model(self, model_input, vocab_size, num_mixtures=None, l2_penalty=1e-8, sub_scope="",
                   original_input=None, **unused_params):
    num_mixtures = num_mixtures or FLAGS.moe_num_mixtures
    p = ops.moe_head(model_input, vocab_size, num_mixtures, "gates" + sub_scope, "experts" + sub_scope, l2_penalty)
    return {"predictions": p}


class MoeExtendModel(models.BaseModel):
  """MoE on B*A attention rows, max over the A heads (zt/video_level_models.py:2272-2330)."""

  def create_model(self, model_input, vocab_size, num_mixtures=None, l2_penalty=1e-8, **unused_params):
    num_mixtures = num_mixtures or FLAGS.moe_num_mixtures
    num_extends = FLAGS.moe_num_extend
    p = ops.moe_head(model_input, vocab_size, num_mixtures, "gates", "experts", l2_penalty)
    return {"predictions": nat.group_max_rows(p, num_extends)}


class ChainMoeModel(models.BaseModel):
  """Support MoE -> concat -> main MoE (wh/all_video_models/chain_moe_model.py:9-49)."""

  def create_model(self, model_input, vocab_size, num_mixtures=None, l2_penalty=1e-8, sub_scope="",
                   original_input=None, **unused_params):
    num_supports = FLAGS.num_supports
    x = ops.as_act(model_input)
    support = self.sub_model(x, num_supports, sub_scope=sub_scope + "-support")
    main_input = ops.concat([x, support])
    main = self.sub_model(main_input, vocab_size, sub_scope=sub_scope + "-main")
    return {"predictions": main, "support_predictions": support}

  def sub_model(self, model_input, vocab_size, num_mixtures=None, l2_penalty=1e-8, sub_scope="", **unused_params):
    num_mixtures = num_mixtures or FLAGS.moe_num_mixtures
    return ops.moe_head(model_input, vocab_size, num_mixtures, "gates" + sub_scope, "experts" + sub_scope, l2_penalty)


class DeepCombineChainModel(models.BaseModel):
  """Stacked MoE sub-predictions, each projected (4716 -> relu_cells), ReLU, L2-normalised and
  concatenated to the next MoE's input (wh/all_video_models/deep_combine_chain_model.py:9-85)."""

  def create_model(self, model_input, vocab_size, num_mixtures=None, l2_penalty=1e-8, sub_scope="",
                   original_input=None, dropout=False, keep_prob=None, noise_level=None, num_frames=None,
                   **unused_params):
    num_layers = FLAGS.deep_chain_layers
    relu_cells = FLAGS.deep_chain_relu_cells
    if FLAGS.deep_chain_relu_type == "elu":
      raise NotImplementedError("deep_chain_relu_type=elu is not on the B200 hot path (default is relu)")
    if dropout or noise_level is not None:
      raise NotImplementedError("dropout / noise_level are training-time extras outside the hot path")
    next_input = ops.as_act(model_input)
    support_predictions = []
    for layer in range(num_layers):
      sub_prediction = self.sub_model(next_input, vocab_size, sub_scope=sub_scope + "prediction-%d" % layer)
      sub_relu = ops.fully_connected(sub_prediction, relu_cells, sub_scope + "relu-%d" % layer, activation_fn="relu",
                                     l2_penalty=l2_penalty, want_bf16=False)
      relu_norm = ops.l2_normalize_rows(sub_relu)
      next_input = ops.concat([next_input, relu_norm])
      support_predictions.append(sub_prediction)
    main = self.sub_model(next_input, vocab_size, sub_scope=sub_scope + "-main")
    import torch  # device-side concatenation of the per-layer outputs (a copy, no arithmetic)
    return {"predictions": main, "support_predictions": torch.cat(support_predictions, dim=1)}

  def sub_model(self, model_input, vocab_size, num_mixtures=None, l2_penalty=1e-8, sub_scope="", **unused_params):
    num_mixtures = num_mixtures or FLAGS.moe_num_mixtures
    return ops.moe_head(model_input, vocab_size, num_mixtures, "gates-" + sub_scope, "experts-" + sub_scope, l2_penalty)
