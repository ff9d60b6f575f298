"""Binary that writes a model's FULL prediction vectors as TFRecords for the ensemble stage -- the command line and
wire format of wh/inference-pre-ensemble.py (flags :41-88, writer :291-308): files
``<output_dir>/predictions-%04d.tfrecord`` of --file_size Examples, each with ``video_id`` (bytes), ``labels`` (int64
indices of the positive classes) and ``predictions`` (float list, one confidence per class).  The model forward runs
on the GPU through the C ABI; TFRecord framing (masked CRC-32C) and the Example encoding are the readers.py codecs."""
import logging
import os
import sys
import time

import numpy as np
import torch

import feature_transform
import frame_level_models
import readers
import utils
import video_level_models
import yt8m_flags as flags
from eval import restore

FLAGS = flags.FLAGS

if __name__ == "__main__":
  flags.DEFINE_string("train_dir", "/tmp/yt8m_model/", "The directory to load the model files from.")
  flags.DEFINE_string("model_checkpoint_path", None, "The file path to load the model from.")
  flags.DEFINE_string("output_dir", "", "The file to save the predictions to.")
  flags.DEFINE_string("input_data_pattern", "", "File glob defining the evaluation dataset in tensorflow.SequenceExample format.")
  flags.DEFINE_string("distill_data_pattern", None, "File glob defining the distillation data pattern (accepted, unused)")
  flags.DEFINE_bool("frame_features", False, "If set, then --input_data_pattern must be frame-level features.")
  flags.DEFINE_integer("batch_size", 8192, "How many examples to process per batch.")
  flags.DEFINE_string("feature_names", "mean_rgb", "Name of the feature to use for training.")
  flags.DEFINE_string("feature_sizes", "1024", "Length of the feature vectors.")
  flags.DEFINE_integer("file_size", 4096, "Number of examples per output file.")
  flags.DEFINE_string("model", "YouShouldSpecifyAModel", "Which architecture to use for the model.")
  flags.DEFINE_integer("num_readers", 1, "How many threads to use for reading input files. (accepted, unused)")
  flags.DEFINE_integer("top_k", 20, "How many predictions to output per video. (accepted, unused: the full vector is written)")
  flags.DEFINE_bool("dropout", False, "Whether to consider dropout")
  flags.DEFINE_float("keep_prob", 1.0, "probability to keep output (used in dropout, keep it unchanged in validationg and test)")
  flags.DEFINE_float("noise_level", 0.0, "standard deviation of noise (added to hidden nodes)")


def get_output_feature(video_id, labels, features, feature_names):
  """wh/inference-pre-ensemble.py:301-308: one tf.train.Example (serialised)."""
  vid = video_id if isinstance(video_id, (bytes, bytearray)) else str(video_id).encode("utf-8")
  feats = {"video_id": ("bytes", [bytes(vid)]), "labels": ("int64", [int(v) for v in labels])}
  for name, values in zip(feature_names, features):
    feats[name] = ("float", np.asarray(values, dtype=np.float32))
  return readers.encode_example(feats)


def write_to_record(output_dir, id_batch, label_batch, predictions, filenum, num_examples_processed):
  """wh/inference-pre-ensemble.py:291-299."""
  path = os.path.join(output_dir, "predictions-%04d.tfrecord" % filenum)
  recs = []
  for i in range(num_examples_processed):
    label = np.nonzero(label_batch[i, :])[0]
    recs.append(get_output_feature(id_batch[i], label, [predictions[i, :]], ["predictions"]))
  readers.write_tfrecord(path, recs)
  return path


def inference(reader, model, checkpoint, data_pattern, output_dir, batch_size, file_size):
  """wh/inference-pre-ensemble.py:205-288.  Batches need not divide file_size (the reference asserts they do)."""
  if os.path.exists(output_dir):
    raise IOError("Output path exists! path='" + output_dir + "'")
  os.makedirs(output_dir)
  transformer = utils.find_class_by_name(FLAGS.feature_transformer, [feature_transform])()
  restored, start = False, time.time()
  ids, labs, preds, held, filenum, total = [], [], [], 0, 0, 0

  def flush(n):
    nonlocal ids, labs, preds, held, filenum
    vid = [v for chunk in ids for v in chunk]
    lab, prd = np.concatenate(labs, axis=0), np.concatenate(preds, axis=0)
    write_to_record(output_dir, vid[:n], lab[:n], prd[:n], filenum, n)
    filenum += 1
    ids, labs, preds, held = ([vid[n:]], [lab[n:]], [prd[n:]], len(vid) - n) if len(vid) > n else ([], [], [], 0)

  packed = {"packed": True} if FLAGS.frame_features else {}         # readers.PackedFrames: no padding over PCIe
  for video_ids, feats, labels, num_frames in reader.prepare_reader(data_pattern, batch_size, 1, **packed):
    nf = num_frames.cuda() if FLAGS.frame_features else None
    x, _ = transformer.transform(feats.cuda(non_blocking=True), nf)
    if not restored:
      restore(checkpoint, model, x, nf, reader.num_classes)
      restored = True
    kw = {"num_frames": nf} if nf is not None else {}
    p = model.create_model(x, vocab_size=reader.num_classes, is_training=False, **kw)["predictions"]
    ids.append(list(video_ids))
    labs.append(labels.numpy())
    preds.append(p.cpu().numpy())
    held += len(video_ids)
    total += len(video_ids)
    logging.info("num examples processed: " + str(total) + " elapsed seconds: " + "{0:.2f}".format(time.time() - start))
    while held >= file_size:
      flush(file_size)
  if held > 0:
    flush(held)
  logging.info("Done with inference. The output file was written to " + output_dir)
  return filenum


def main(unused_argv=None):
  logging.basicConfig(level=logging.INFO, format="%(levelname)s:%(message)s")
  FLAGS.parse()
  if not torch.cuda.is_available():
    raise SystemExit("inference-pre-ensemble.py: no CUDA device; the yt8m_b200 path has no CPU fallback")
  feature_names, feature_sizes = utils.GetListOfFeatureNamesAndSizes(FLAGS.feature_names, FLAGS.feature_sizes)
  reader = (readers.YT8MFrameFeatureReader if FLAGS.frame_features else readers.YT8MAggregatedFeatureReader)(
      feature_names=feature_names, feature_sizes=feature_sizes)
  if not FLAGS.output_dir:
    raise ValueError("'output_dir' was not specified. Unable to continue with inference.")
  if not FLAGS.input_data_pattern:
    raise ValueError("'input_data_pattern' was not specified. Unable to continue with inference.")
  ckpt = FLAGS.model_checkpoint_path or utils.latest_checkpoint(FLAGS.train_dir)
  if not ckpt:
    raise IOError("Unable to find a checkpoint in " + FLAGS.train_dir)
  model = utils.find_class_by_name(FLAGS.model, [frame_level_models, video_level_models])()
  inference(reader, model, ckpt, FLAGS.input_data_pattern, FLAGS.output_dir, FLAGS.batch_size, FLAGS.file_size)


if __name__ == "__main__":
  main(sys.argv)
