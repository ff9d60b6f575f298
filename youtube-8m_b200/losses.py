"""Loss plugins (wh/losses.py).  CrossEntropyLoss is the default --label_loss (wh/train.py:88)."""
import yt8m_flags as flags
import yt8m_native as nat

flags.DEFINE_integer("num_classes", 4716, "number of classes")
flags.DEFINE_bool("label_smoothing", False, "whether do label smoothing")

# flags of the reference's other losses (wh/losses.py:22-44): accepted so that its command lines parse; only CrossEntropyLoss is built
flags.DEFINE_float("label_smoothing_epsilon", 0.1, "reference flag of a model / loss outside SURVEY.md §8 (accepted, unused)")
flags.DEFINE_float("batch_agreement", 0.1, "reference flag of a model / loss outside SURVEY.md §8 (accepted, unused)")
flags.DEFINE_float("false_positive_punishment", 1.0, "reference flag of a model / loss outside SURVEY.md §8 (accepted, unused)")
flags.DEFINE_float("false_negative_punishment", 1.0, "reference flag of a model / loss outside SURVEY.md §8 (accepted, unused)")
flags.DEFINE_integer("num_frequents", 200, "reference flag of a model / loss outside SURVEY.md §8 (accepted, unused)")
flags.DEFINE_integer("num_verticals", 25, "reference flag of a model / loss outside SURVEY.md §8 (accepted, unused)")
flags.DEFINE_float("support_loss_percent", 0.1, "reference flag of a model / loss outside SURVEY.md §8 (accepted, unused)")
flags.DEFINE_string("support_type", 'vertical', "reference flag of a model / loss outside SURVEY.md §8 (accepted, unused)")
flags.DEFINE_string("vertical_file", 'resources/vertical.tsv', "reference flag of a model / loss outside SURVEY.md §8 (accepted, unused)")


class BaseLoss(object):
  """Inherit from this class when implementing new losses (wh/losses.py:56-74)."""

  def calculate_loss(self, unused_predictions, unused_labels, **unused_params):
    raise NotImplementedError()


class CrossEntropyLoss(BaseLoss):
  """wh/losses.py:110-130: mean_b sum_v -[y log(p + 1e-5) + (1 - y) log(1 - p + 1e-5)]
  (the reference writes epsilon = 10e-6).  Returns a 1-element CUDA tensor."""

  def calculate_loss(self, predictions, labels, weights=None, **unused_params):
    if weights is not None:
      raise NotImplementedError("per-video loss weights (boosting pipeline) are outside the hot path")
    if flags.FLAGS.label_smoothing:
      raise NotImplementedError("label_smoothing is outside the hot path (default False)")
    loss, _ = nat.xent(predictions, labels.to(predictions.device).float())
    return loss

  def calculate_loss_and_grad(self, predictions, labels, grad_scale=1.0):
    """Also returns dLoss/dpredictions (the backward entry of the train step)."""
    return nat.xent(predictions, labels.to(predictions.device).float(), want_grad=True, grad_scale=grad_scale)
