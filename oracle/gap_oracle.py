"""CPU restatement of the reference's metric path (GAP@k, Hit@1, PERR, per-class AP).  TEST
INFRASTRUCTURE ONLY (see oracle/yt8m_oracle.py for who may import ``oracle/``).

Pinned: tests/test_gap_oracle.py checks every function here against golden vectors produced by
the reference's OWN code (wh/eval_util.py, wh/average_precision_calculator.py,
wh/mean_average_precision_calculator.py) run in the build container by oracle/make_golden.py.

Deliberately written as straight loops over sorted lists -- a second, independent formulation of
the arithmetic the reference does with heaps -- so that it can check both the reference fixtures
and the product's vectorised implementation (youtube-8m_b200/eval_util.py).
"""
import heapq
import random

import numpy as np


def hit_at_one(predictions, actuals):
  """wh/eval_util.py:28-42: mean over rows of actuals[row, argmax(predictions[row])]."""
  top = np.argmax(predictions, axis=1)
  return float(np.mean(actuals[np.arange(actuals.shape[0]), top]))


def perr(predictions, actuals):
  """wh/eval_util.py:74-99: per row, precision among the num_labels highest predictions
  (only those > 0 count), averaged over rows."""
  total = 0.0
  for row in range(actuals.shape[0]):
    n = int(np.sum(actuals[row]))
    idx = np.argpartition(predictions[row], -n)[-n:]
    got = sum(float(actuals[row][j]) for j in idx if predictions[row][j] > 0)
    total += got / idx.size
  return total / actuals.shape[0]


def top_k_pairs(predictions, actuals, k=20):
  """wh/eval_util.py:123-165: the (prediction, label) pairs of the k best classes of every row,
  flattened CLASS-major / video-minor exactly like top_k_by_class + flatten (the order matters only for
  how the reference breaks ties), and the total number of positives in ``actuals``."""
  k = min(k, predictions.shape[1])
  per_class_p = [[] for _ in range(predictions.shape[1])]
  per_class_l = [[] for _ in range(predictions.shape[1])]
  for row in range(predictions.shape[0]):
    idx = np.argpartition(predictions[row], -k)[-k:]
    for j in idx:
      per_class_p[j].append(predictions[row][j])
      per_class_l[j].append(actuals[row][j])
  ps = [x for c in per_class_p for x in c]
  ls = [x for c in per_class_l for x in c]
  return ps, ls, float(np.sum(actuals))


def heap_order(ps, ls, heap=None):
  """wh/average_precision_calculator.py:127-133: the calculator keeps (prediction, actual) in a heapq
  list; peek_ap_at_n reads them back in heap-array order."""
  heap = [] if heap is None else heap
  for p, l in zip(ps, ls):
    heapq.heappush(heap, (p, l))
  return heap


def ap_at_n(predictions, actuals, n=None, total_num_positives=None):
  """wh/average_precision_calculator.py:179-253: non-interpolated AP; ties are broken by the
  reference's fixed shuffle (random.seed(0); random.sample), then a stable descending sort."""
  predictions = np.asarray(predictions)
  actuals = np.asarray(actuals)
  random.seed(0)
  perm = random.sample(range(len(predictions)), len(predictions))
  predictions, actuals = predictions[perm], actuals[perm]
  order = sorted(range(len(predictions)), key=lambda i: predictions[i], reverse=True)
  numpos = int(np.sum(actuals > 0)) if total_num_positives is None else total_num_positives
  if numpos == 0:
    return 0.0
  if n is not None:
    numpos = min(numpos, n)
  r = len(order) if n is None else min(len(order), n)
  ap, hits = 0.0, 0.0
  for i in range(r):
    if actuals[order[i]] > 0:
      hits += 1
      ap += hits / (i + 1) / numpos
  return ap


def gap(predictions, actuals, top_k=20):
  """wh/eval_util.py:102-120 (calculate_gap): AP over the pooled per-video top-k pairs with
  numpos = all positives in ``actuals``."""
  p, l, numpos = top_k_pairs(predictions, actuals, top_k)
  heap = heap_order(p, l)
  hp, hl = zip(*heap)
  return ap_at_n(np.array(hp), np.array(hl), n=None, total_num_positives=numpos)


class StreamingGap:
  """wh/eval_util.py:167-254 (EvaluationMetrics) restricted to hit@1 / perr / loss / gap: accumulate
  mini-batches into one heap, report epoch averages."""

  def __init__(self, top_k=20):
    self.top_k = top_k
    self.clear()

  def clear(self):
    self._heap, self._pos = [], 0.0
    self.sum_hit, self.sum_perr, self.sum_loss, self.n = 0.0, 0.0, 0.0, 0

  def accumulate(self, predictions, labels, loss):
    b = labels.shape[0]
    p, l, pos = top_k_pairs(predictions, labels, self.top_k)
    heap_order(p, l, self._heap)
    self._pos += pos
    self.sum_hit += hit_at_one(predictions, labels) * b
    self.sum_perr += perr(predictions, labels) * b
    self.sum_loss += float(np.mean(loss)) * b
    self.n += b

  def get(self):
    if self.n <= 0:
      raise ValueError("total_sample must be positive.")
    hp, hl = zip(*self._heap)
    return {"avg_hit_at_one": self.sum_hit / self.n, "avg_perr": self.sum_perr / self.n,
            "avg_loss": self.sum_loss / self.n,
            "gap": ap_at_n(np.array(hp), np.array(hl), n=None, total_num_positives=self._pos)}
