"""Binary for generating predictions over a set of videos -- the command line of wh/inference.py (flags
:35-73, output format :76-87,163): ``VideoId,LabelConfidencePairs`` with the top_k classes per video.
The top-k extraction runs on the GPU (yt8m_topk_rows); only k (class, confidence) pairs per video cross PCIe."""
import logging
import sys
import time

import torch

import feature_transform
import frame_level_models
import readers
import utils
import video_level_models
import yt8m_flags as flags
import yt8m_native as nat
from eval import restore

FLAGS = flags.FLAGS

if __name__ == "__main__":
  flags.DEFINE_string("train_dir", "/tmp/yt8m_model/", "The directory to load the model files from.")
  flags.DEFINE_string("model_checkpoint_path", "", "The file path to load the model from.")
  flags.DEFINE_string("output_file", "", "The file to save the predictions to.")
  flags.DEFINE_string("input_data_pattern", "", "File glob defining the evaluation dataset in tensorflow.SequenceExample format.")
  flags.DEFINE_bool("frame_features", False, "If set, then --input_data_pattern must be frame-level features.")
  flags.DEFINE_integer("batch_size", 8192, "How many examples to process per batch.")
  flags.DEFINE_string("feature_names", "mean_rgb", "Name of the feature to use for training.")
  flags.DEFINE_string("feature_sizes", "1024", "Length of the feature vectors.")
  flags.DEFINE_string("model", "LogisticModel", "Which architecture the checkpoint holds.")
  flags.DEFINE_integer("num_readers", 1, "How many threads to use for reading input files. (accepted, unused)")
  flags.DEFINE_integer("top_k", 20, "How many predictions to output per video.")
  flags.DEFINE_bool("dropout", False, "Whether to consider dropout")
  flags.DEFINE_float("keep_prob", 1.0, "probability to keep output (used in dropout, keep it unchanged in validationg and test)")


def format_lines(video_ids, top_idx, top_val):
  """wh/inference.py:76-87: 'id,cls conf cls conf ...' sorted by descending confidence."""
  for vid, idx, val in zip(video_ids, top_idx, top_val):
    vid = vid.decode("utf-8") if isinstance(vid, (bytes, bytearray)) else str(vid)
    yield vid + "," + " ".join("%i %f" % (int(i), float(v)) for i, v in zip(idx, val)) + "\n"


def inference(reader, model, checkpoint, data_pattern, out_file_location, batch_size, top_k):
  transformer = utils.find_class_by_name(FLAGS.feature_transformer, [feature_transform])()
  restored, n, start = False, 0, time.time()
  with open(out_file_location, "w+") as out_file:
    out_file.write("VideoId,LabelConfidencePairs\n")
    packed = {"packed": True} if FLAGS.frame_features else {}       # readers.PackedFrames: no padding over PCIe
    for video_ids, feats, _, num_frames in reader.prepare_reader(data_pattern, batch_size, 1, **packed):
      nf = num_frames.cuda() if FLAGS.frame_features else None
      x, _ = transformer.transform(feats.cuda(non_blocking=True), nf)
      if not restored:
        restore(checkpoint, model, x, nf, reader.num_classes)
        restored = True
      kw = {"num_frames": nf} if nf is not None else {}
      p = model.create_model(x, vocab_size=reader.num_classes, is_training=False, **kw)["predictions"]
      idx, val = nat.topk_rows(p, min(top_k, p.shape[1]))
      n += len(video_ids)
      logging.info("num examples processed: " + str(n) + " elapsed seconds: " + "{0:.2f}".format(time.time() - start))
      for line in format_lines(video_ids, idx.cpu().numpy(), val.cpu().numpy()):
        out_file.write(line)
      out_file.flush()
  logging.info("Done with inference. The output file was written to " + out_file_location)


def main(unused_argv=None):
  logging.basicConfig(level=logging.INFO, format="%(levelname)s:%(message)s")
  FLAGS.parse()
  if not torch.cuda.is_available():
    raise SystemExit("inference.py: no CUDA device; the yt8m_b200 path has no CPU fallback")
  feature_names, feature_sizes = utils.GetListOfFeatureNamesAndSizes(FLAGS.feature_names, FLAGS.feature_sizes)
  reader = (readers.YT8MFrameFeatureReader if FLAGS.frame_features else readers.YT8MAggregatedFeatureReader)(
      feature_names=feature_names, feature_sizes=feature_sizes)
  if not FLAGS.output_file:
    raise ValueError("'output_file' was not specified. Unable to continue with inference.")
  if not FLAGS.input_data_pattern:
    raise ValueError("'input_data_pattern' was not specified. Unable to continue with inference.")
  ckpt = FLAGS.model_checkpoint_path or utils.latest_checkpoint(FLAGS.train_dir)
  if not ckpt:
    raise IOError("Unable to find a checkpoint in " + FLAGS.train_dir)
  model = utils.find_class_by_name(FLAGS.model, [frame_level_models, video_level_models])()
  inference(reader, model, ckpt, FLAGS.input_data_pattern, FLAGS.output_file, FLAGS.batch_size, FLAGS.top_k)


if __name__ == "__main__":
  main(sys.argv)
