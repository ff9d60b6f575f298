"""Utility functions shared by train.py / eval.py / inference.py (wh/utils.py)."""
import glob
import logging
import os
import re

import numpy
import torch


def Dequantize(feat_vector, max_quantized_value=2, min_quantized_value=-2):  # noqa: N802 (reference name)
  """Host-side statement of wh/utils.py:23-38 (the GPU path fuses it into yt8m_l2norm_rows_fwd)."""
  assert max_quantized_value > min_quantized_value
  quantized_range = max_quantized_value - min_quantized_value
  scalar = quantized_range / 255.0
  bias = (quantized_range / 512.0) + min_quantized_value
  return feat_vector * scalar + bias


def GetListOfFeatureNamesAndSizes(feature_names, feature_sizes):  # noqa: N802
  """wh/utils.py:140-162."""
  list_of_feature_names = [n.strip() for n in feature_names.split(",")]
  list_of_feature_sizes = [int(s) for s in feature_sizes.split(",")]
  if len(list_of_feature_names) != len(list_of_feature_sizes):
    logging.error("length of the feature names (=" + str(len(list_of_feature_names)) + ") != length of feature "
                  "sizes (=" + str(len(list_of_feature_sizes)) + ")")
  return list_of_feature_names, list_of_feature_sizes


def FormatEpochInfo(epoch_info_dict):  # noqa: N802
  """The epoch line of wh/utils.py:100-138 (AddEpochSummary) without the TensorBoard writer."""
  mean_ap = numpy.mean(epoch_info_dict["aps"])
  return ("epoch/eval number {0} | Avg_Hit@1: {1:.3f} | Avg_PERR: {2:.3f} "
          "| MAP: {3:.3f} | GAP: {4:.3f} | Avg_Loss: {5:3f}").format(
              epoch_info_dict["epoch_id"], epoch_info_dict["avg_hit_at_one"], epoch_info_dict["avg_perr"], mean_ap,
              epoch_info_dict["gap"], epoch_info_dict["avg_loss"])


def find_class_by_name(name, modules):
  """Searches the provided modules for the named class and returns it (wh/train.py:212-215): the first module
  attribute with that name wins; a missing name raises StopIteration like the reference's next()."""
  modules = [getattr(module, name, None) for module in modules]
  return next(a for a in modules if a)


# ---- checkpoints: train_dir/model.ckpt-<global_step> --------------------------------------------------

def checkpoint_path(train_dir, step):
  return os.path.join(train_dir, "model.ckpt-%d" % step)


def latest_checkpoint(train_dir):
  """Like tf.train.latest_checkpoint: the model.ckpt-<step> with the largest step, or None."""
  best, best_step = None, -1
  for p in glob.glob(os.path.join(train_dir, "model.ckpt-*")):
    m = re.search(r"model\.ckpt-(\d+)$", p)
    if m and int(m.group(1)) > best_step:
      best, best_step = p, int(m.group(1))
  return best


def save_checkpoint(train_dir, step, variables, optimizer_state=None, flags_dict=None, max_to_keep=3):
  """variables: {reference variable name: CPU tensor in the reference (TF) layout}."""
  os.makedirs(train_dir, exist_ok=True)
  path = checkpoint_path(train_dir, step)
  torch.save({"global_step": step, "variables": variables, "optimizer": optimizer_state, "flags": flags_dict}, path)
  kept = sorted(glob.glob(os.path.join(train_dir, "model.ckpt-*")), key=lambda p: int(p.rsplit("-", 1)[1]))
  for old in kept[:-max_to_keep]:
    os.remove(old)
  return path


def load_checkpoint(path):
  return torch.load(path, map_location="cpu", weights_only=False)
