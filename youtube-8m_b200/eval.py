"""Binary for evaluating models on the YouTube-8M dataset -- the command line of wh/eval.py (flags :33-84,
loop :208-322): restore a checkpoint, run the model plugin forward over the eval split, accumulate
Hit@1 / PERR / MAP / GAP with EvaluationMetrics, print the reference's epoch line."""
import logging
import sys
import time

import torch

import eval_util
import feature_transform
import frame_level_models
import losses
import readers
import utils
import video_level_models
import yt8m_flags as flags
import yt8m_ops as ops

FLAGS = flags.FLAGS

if __name__ == "__main__":
  flags.DEFINE_string("train_dir", "/tmp/yt8m_model/", "The directory to load the model files from.")
  flags.DEFINE_string("model_checkpoint_path", "", "The file path to load the model from.")
  flags.DEFINE_string("distill_data_pattern", None, "File glob defining the distillation data pattern (accepted, unused)")
  flags.DEFINE_string("eval_data_pattern", "", "File glob defining the evaluation dataset in tensorflow.SequenceExample format.")
  flags.DEFINE_string("feature_names", "mean_rgb", "Name of the feature to use for training.")
  flags.DEFINE_string("feature_sizes", "1024", "Length of the feature vectors.")
  flags.DEFINE_bool("frame_features", False, "If set, then --eval_data_pattern must be frame-level features.")
  flags.DEFINE_string("model", "LogisticModel", "Which architecture to use for the model.")
  flags.DEFINE_integer("batch_size", 1024, "How many examples to process per batch.")
  flags.DEFINE_string("label_loss", "CrossEntropyLoss", "Loss computed on validation data")
  flags.DEFINE_integer("num_readers", 8, "How many threads to use for reading input files. (accepted, unused)")
  flags.DEFINE_bool("run_once", False, "Whether to run eval only once.")
  flags.DEFINE_integer("top_k", 20, "How many predictions to output per video.")
  flags.DEFINE_bool("multitask", False, "Whether to consider support_predictions")
  flags.DEFINE_bool("dropout", False, "Whether to consider dropout")
  flags.DEFINE_float("keep_prob", 1.0, "probability to keep output (used in dropout, keep it unchanged in validationg and test)")
  flags.DEFINE_float("noise_level", 0.0, "standard deviation of noise (added to hidden nodes)")


def get_reader():
  feature_names, feature_sizes = utils.GetListOfFeatureNamesAndSizes(FLAGS.feature_names, FLAGS.feature_sizes)
  if FLAGS.frame_features:
    return readers.YT8MFrameFeatureReader(feature_names=feature_names, feature_sizes=feature_sizes)
  return readers.YT8MAggregatedFeatureReader(feature_names=feature_names, feature_sizes=feature_sizes)


def restore(checkpoint, model, example_input, num_frames, vocab_size):
  """Variables are created by a first create_model call (the reference rebuilds the graph, wh/eval.py:168-182),
  then overwritten from the checkpoint (Saver.restore, :230-240)."""
  st = ops.get_store()
  st.reset(seed=9)
  kw = {"num_frames": num_frames} if num_frames is not None else {}
  model.create_model(example_input, vocab_size=vocab_size, is_training=False, **kw)
  ck = utils.load_checkpoint(checkpoint)
  st.load_state_dict(ck["variables"], strict=True)
  return ck["global_step"]


def evaluation_loop(reader, model, checkpoint, last_global_step_val):
  """wh/eval.py:208-322.  Returns the global step evaluated."""
  transformer = utils.find_class_by_name(FLAGS.feature_transformer, [feature_transform])()
  loss_fn = utils.find_class_by_name(FLAGS.label_loss, [losses])()
  evl_metrics = eval_util.EvaluationMetrics(reader.num_classes, FLAGS.top_k)
  global_step_val, examples_processed = None, 0
  packed = {"packed": True} if FLAGS.frame_features else {}         # readers.PackedFrames: no padding over PCIe
  for video_ids, feats, labels, num_frames in reader.prepare_reader(FLAGS.eval_data_pattern, FLAGS.batch_size, 1, **packed):
    t0 = time.time()
    nf = num_frames.cuda() if FLAGS.frame_features else None
    x, _ = transformer.transform(feats.cuda(non_blocking=True), nf)
    if global_step_val is None:
      global_step_val = restore(checkpoint, model, x, nf, reader.num_classes)
      if global_step_val == last_global_step_val:
        logging.info("skip this checkpoint global_step_val=%s (same as the previous one).", global_step_val)
        return global_step_val
    kw = {"num_frames": nf} if nf is not None else {}
    result = model.create_model(x, vocab_size=reader.num_classes, is_training=False, **kw)
    p = result["predictions"]
    y = labels.cuda(non_blocking=True).float()
    if FLAGS.multitask:                            # wh/eval.py:188-190: the loss also sees the support predictions
      loss = float(loss_fn.calculate_loss(p, result["support_predictions"], labels))
    else:
      loss = float(loss_fn.calculate_loss(p, y))
    it = evl_metrics.accumulate(p.cpu().numpy(), labels.numpy().astype("float32"), loss)
    examples_processed += labels.shape[0]
    logging.info("examples_processed: %d | global_step %s | Batch Hit@1: %.3f | Batch PERR: %.3f | Batch Loss: %.3f | "
                 "Examples_per_sec: %.3f", examples_processed, global_step_val, it["hit_at_one"], it["perr"], it["loss"],
                 labels.shape[0] / max(time.time() - t0, 1e-9))
  logging.info("Done with batched inference. Now calculating global performance metrics.")
  epoch_info = evl_metrics.get()
  epoch_info["epoch_id"] = global_step_val
  logging.info(utils.FormatEpochInfo(epoch_info))
  return global_step_val, epoch_info


def evaluate():
  if not FLAGS.eval_data_pattern:
    raise IOError("'eval_data_pattern' was not specified. Nothing to evaluate.")
  reader = get_reader()
  model = utils.find_class_by_name(FLAGS.model, [frame_level_models, video_level_models])()
  last = -1
  while True:
    ckpt = FLAGS.model_checkpoint_path or utils.latest_checkpoint(FLAGS.train_dir)
    if not ckpt:
      logging.info("No checkpoint file found.")
    else:
      res = evaluation_loop(reader, model, ckpt, last)
      last = res[0] if isinstance(res, tuple) else res
    if FLAGS.run_once:
      break
    time.sleep(60)


def main(unused_argv=None):
  logging.basicConfig(level=logging.INFO, format="%(levelname)s:%(message)s")
  FLAGS.parse()
  if not torch.cuda.is_available():
    raise SystemExit("eval.py: no CUDA device; the yt8m_b200 path has no CPU fallback")
  evaluate()


if __name__ == "__main__":
  main(sys.argv)
