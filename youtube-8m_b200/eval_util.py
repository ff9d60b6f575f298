"""Evaluation metrics with the reference's API (wh/eval_util.py:28-254): Hit@1, PERR, GAP@k and the
streaming EvaluationMetrics.  These run on the host over fetched numpy arrays, as in the reference;
this implementation is vectorised (argpartition / argsort over whole batches instead of per-row Python
loops and heap pushes) and is checked against golden vectors produced by the reference's own code
(tests/golden/gap_golden.json).  Tie-breaking follows the reference: a fixed shuffle seeded with 0, then
a stable descending sort (wh/average_precision_calculator.py:216-222,247-253).
"""
import random

import numpy


def calculate_hit_at_one(predictions, actuals):
  """Mean over the batch of actuals[row, argmax(predictions[row])] (wh/eval_util.py:28-42)."""
  top = numpy.argmax(predictions, 1)
  return float(numpy.average(actuals[numpy.arange(actuals.shape[0]), top]))


def calculate_precision_at_equal_recall_rate(predictions, actuals):
  """PERR (wh/eval_util.py:74-99): per video, the precision among its num_labels best classes."""
  num_videos = actuals.shape[0]
  counts = actuals.sum(axis=1).astype(numpy.int64)
  total = 0.0
  for n in numpy.unique(counts):
    rows = numpy.nonzero(counts == n)[0]
    if n <= 0:
      # a label-free video contributes 0: the reference's argpartition(..., -0)[-0:] selects every class and none is a hit
      # (wh/eval_util.py:87-96; golden case "empty_label_rows")
      continue
    # same selection rule as the reference (numpy.argpartition per row), batched over equal n
    idx = numpy.argpartition(predictions[rows], -n, axis=1)[:, -n:]
    r = rows[:, None]
    hits = (actuals[r, idx] * (predictions[r, idx] > 0)).sum(axis=1)
    total += float((hits / n).sum())
  return total / num_videos


def calculate_recall_at_n(predictions, actuals, n):
  """wh/eval_util.py:45-71."""
  num_videos = actuals.shape[0]
  idx = numpy.argpartition(predictions, -n, axis=1)[:, -n:]
  r = numpy.arange(num_videos)[:, None]
  hits = (actuals[r, idx] * (predictions[r, idx] > 0)).sum(axis=1)
  return float(numpy.mean(hits / actuals.sum(axis=1)))


def top_k_by_class(predictions, labels, k=20):
  """wh/eval_util.py:123-157 in array form: returns (values [N*k], labels [N*k], classes [N*k],
  per-class positive counts [num_classes])."""
  if k <= 0:
    raise ValueError("k must be a positive integer.")
  k = min(k, predictions.shape[1])
  idx = numpy.argpartition(predictions, -k, axis=1)[:, -k:]
  r = numpy.arange(predictions.shape[0])[:, None]
  return predictions[r, idx].ravel(), labels[r, idx].ravel(), idx.ravel(), labels.sum(axis=0)


def _ap(predictions, actuals, total_num_positives, n=None):
  """Non-interpolated AP with the reference's tie-breaking (average_precision_calculator.py:179-253)."""
  m = len(predictions)
  if m == 0:
    return 0.0
  random.seed(0)
  perm = numpy.asarray(random.sample(range(m), m))
  p, a = predictions[perm], actuals[perm]
  order = numpy.argsort(-p, kind="stable")
  numpos = total_num_positives if total_num_positives is not None else int((a > 0).sum())
  if numpos == 0:
    return 0.0
  if n is not None:
    numpos = min(numpos, n)
    order = order[:n]
  hit = (a[order] > 0).astype(numpy.float64)
  prec = numpy.cumsum(hit) / numpy.arange(1, len(order) + 1)
  return float((prec * hit).sum() / numpos)


def calculate_gap(predictions, actuals, top_k=20):
  """Global average precision over the pooled per-video top_k (wh/eval_util.py:102-120)."""
  p, l, _, pos = top_k_by_class(predictions, actuals, top_k)
  return _ap(p, l, float(pos.sum()))


def step_metrics_from_topk(top_val, top_idx, actuals, top_k=20):
  """Hit@1, PERR and GAP@top_k of one step from the per-video top-K' entries (K' >= top_k, descending) that the GPU
  extracted (yt8m_topk_rows) -- the per-step metrics of the training log line (wh/train.py:578-591) without fetching
  the B x 4716 predictions: 2 * K' numbers per video cross PCIe and the host work is O(B * K') instead of O(B * V).

  Equal to calculate_hit_at_one / calculate_precision_at_equal_recall_rate / calculate_gap on the full predictions up to
  ties between equal scores.  Returns (hit_at_one, perr, gap); perr is None when some video has more than K' positive
  labels (the caller then falls back to the full predictions for that step)."""
  top_val, top_idx = numpy.asarray(top_val), numpy.asarray(top_idx).astype(numpy.int64)
  n_videos, kp = top_idx.shape
  if top_k > kp:
    raise ValueError("top_k (%d) exceeds the %d entries extracted per video" % (top_k, kp))
  r = numpy.arange(n_videos)[:, None]
  hit_lab = actuals[r, top_idx].astype(numpy.float64)            # label of every extracted class
  hit_at_one = float(numpy.average(hit_lab[:, 0]))
  counts = actuals.sum(axis=1).astype(numpy.int64)
  perr = None
  if counts.max(initial=0) <= kp:
    within = numpy.arange(kp)[None, :] < counts[:, None]         # the num_labels best classes of every video
    hits = (hit_lab * (top_val > 0) * within).sum(axis=1)
    with numpy.errstate(divide="ignore", invalid="ignore"):
      perr = float(numpy.where(counts > 0, hits / numpy.maximum(counts, 1), 0.0).sum() / n_videos)   # label-free video: 0
  k = min(top_k, actuals.shape[1])
  gap = _ap(top_val[:, :k].ravel(), hit_lab[:, :k].ravel(), float(actuals.sum()))
  return hit_at_one, perr, gap


class EvaluationMetrics(object):
  """A class to store the evaluation metrics (wh/eval_util.py:167-254)."""

  def __init__(self, num_class, top_k):
    if not isinstance(num_class, int) or num_class <= 1:
      raise ValueError("num_class must be a positive integer.")
    self.num_class = num_class
    self.top_k = top_k
    self.clear()

  def accumulate(self, predictions, labels, loss):
    batch_size = labels.shape[0]
    mean_hit_at_one = calculate_hit_at_one(predictions, labels)
    mean_perr = calculate_precision_at_equal_recall_rate(predictions, labels)
    mean_loss = float(numpy.mean(loss))
    p, l, c, pos = top_k_by_class(predictions, labels, self.top_k)
    self._p.append(p)
    self._l.append(l)
    self._c.append(c)
    self._pos += pos
    self.num_examples += batch_size
    self.sum_hit_at_one += mean_hit_at_one * batch_size
    self.sum_perr += mean_perr * batch_size
    self.sum_loss += mean_loss * batch_size
    return {"hit_at_one": mean_hit_at_one, "perr": mean_perr, "loss": mean_loss}

  def get(self):
    if self.num_examples <= 0:
      raise ValueError("total_sample must be positive.")
    p, l, c = numpy.concatenate(self._p), numpy.concatenate(self._l), numpy.concatenate(self._c)
    gap = _ap(p, l, float(self._pos.sum()))
    # per-class AP (mean_average_precision_calculator.py:88-112): heaps are per class, so the
    # within-class order is the accumulation order
    aps = []
    order = numpy.argsort(c, kind="stable")
    bounds = numpy.searchsorted(c[order], numpy.arange(self.num_class + 1))
    for v in range(self.num_class):
      sel = order[bounds[v]:bounds[v + 1]]
      aps.append(_ap(p[sel], l[sel], float(self._pos[v])) if len(sel) else 0.0)
    return {"avg_hit_at_one": self.sum_hit_at_one / self.num_examples, "avg_perr": self.sum_perr / self.num_examples,
            "avg_loss": self.sum_loss / self.num_examples, "aps": aps, "gap": gap}

  def clear(self):
    self.sum_hit_at_one = 0.0
    self.sum_perr = 0.0
    self.sum_loss = 0.0
    self.num_examples = 0
    self._p, self._l, self._c = [], [], []
    self._pos = numpy.zeros(self.num_class, dtype=numpy.float64)
