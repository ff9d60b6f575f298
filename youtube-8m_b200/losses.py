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
flags.DEFINE_integer("num_frequents", 200, "Number of total frequent categories.")
flags.DEFINE_integer("num_verticals", 25, "Number of total vertical categories.")
flags.DEFINE_float("support_loss_percent", 0.1, "the part that support loss (in multi-task scenario) take in the whole loss function.")
flags.DEFINE_string("support_type", 'vertical', "type of support label, vertical or frequent or vertical,frequent.")
flags.DEFINE_string("vertical_file", 'resources/vertical.tsv', "Location of label-vertical mapping file.")


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


class MultiTaskLoss(BaseLoss):
  """wh/losses.py:214-258: the support labels of the --multitask losses, derived from the video labels on the HOST (they are
  part of the input batch: a slice, a concatenation, or a 0/1 lookup through the label -> vertical table)."""

  _vertical = None

  def get_support(self, labels, support_type=None):
    """labels: bool / float [B, V] (torch CPU tensor or numpy) -> float32 numpy [B, S]."""
    import numpy as np
    if support_type is None:
      support_type = flags.FLAGS.support_type
    y = np.asarray(labels.cpu() if hasattr(labels, "cpu") else labels).astype(np.float32)
    if "," in support_type:
      return np.concatenate([self.get_support(y, st) for st in support_type.split(",")], axis=1)
    if support_type == "vertical":
      if MultiTaskLoss._vertical is None:
        vm = np.zeros((flags.FLAGS.num_classes, flags.FLAGS.num_verticals), dtype=np.float32)
        with open(flags.FLAGS.vertical_file) as f:                   # "<class> <vertical>" per line (wh/losses.py:237-242)
          for line in f:
            group = [int(t) for t in line.strip().split()]
            if len(group) == 2:
              vm[group[0], group[1]] = 1
        MultiTaskLoss._vertical = vm
      counts = np.zeros((y.shape[0], MultiTaskLoss._vertical.shape[1]), dtype=np.float32)
      rows, cols = np.nonzero(y)                                      # table lookup of every positive label
      np.add.at(counts, rows, MultiTaskLoss._vertical[cols])
      return (counts > 0.2).astype(np.float32)
    if support_type == "frequent":
      return y[:, :flags.FLAGS.num_frequents]
    if support_type == "label":
      return y
    raise NotImplementedError()

  def calculate_loss(self, unused_predictions, unused_support_predictions, unused_labels, **unused_params):
    raise NotImplementedError()


class MultiTaskCrossEntropyLoss(MultiTaskLoss):
  """wh/losses.py:271-279: CE(predictions, labels) * (1 - p) + CE(support_predictions, support_labels) * p with
  p = --support_loss_percent (the loss of the reference's chain-model scripts, --multitask=True)."""

  def calculate_loss(self, predictions, support_predictions, labels, **unused_params):
    import torch
    sup = torch.from_numpy(self.get_support(labels)).to(predictions.device)
    ce = CrossEntropyLoss()
    pct = flags.FLAGS.support_loss_percent
    main = ce.calculate_loss(predictions, labels, **unused_params)
    support = ce.calculate_loss(support_predictions.contiguous(), sup, **unused_params)
    return float(main) * (1.0 - pct) + float(support) * pct
