"""Binary for training models on the YouTube-8M dataset -- the command line of wh/train.py (flags at
:38-137, loop semantics at :504-622) on the B200 path.

    python train.py --train_data_pattern='train*.tfrecord' --model=MoeModel --feature_names="mean_rgb,mean_audio" \
        --feature_sizes="1024,128" --train_dir=/tmp/yt8m_model --start_new_model

Multi-GPU: launch with torchrun (one process per GPU); every rank reads a disjoint slice of each batch and the
step does ONE NCCL all-reduce of the flat gradient buffer (synchronous data parallel; the reference's
asynchronous parameter-server mode is not reproduced -- SURVEY.md §5).
"""
import logging
import os
import shutil
import sys
import time

import torch

import eval_util
import feature_transform  # noqa: F401 (registers flags)
import frame_level_models
import losses  # noqa: F401
import readers
import utils
import video_level_models
import yt8m_dp
import yt8m_native as nat
import yt8m_flags as flags

FLAGS = flags.FLAGS

if __name__ == "__main__":
  flags.DEFINE_string("train_dir", "/tmp/yt8m_model/", "The directory to save the model files in.")
  flags.DEFINE_string("train_data_pattern", "", "File glob for the training dataset (TFRecords of Example / SequenceExample).")
  flags.DEFINE_string("feature_names", "mean_rgb", "Name of the feature to use for training.")
  flags.DEFINE_string("feature_sizes", "1024", "Length of the feature vectors.")
  flags.DEFINE_bool("frame_features", False, "If set, then --train_data_pattern must be frame-level features.")
  flags.DEFINE_string("model", "LogisticModel", "Which architecture to use for the model.")
  flags.DEFINE_bool("multitask", False, "Whether to consider support_predictions")
  flags.DEFINE_bool("start_new_model", False, "If set, this will not resume from a checkpoint and will instead create a new model instance.")
  flags.DEFINE_integer("batch_size", 1024, "How many examples to process per batch for training.")
  flags.DEFINE_string("label_loss", "CrossEntropyLoss", "Which loss function to use for training the model.")
  flags.DEFINE_float("regularization_penalty", 1, "How much weight to give to the regularization loss (the label loss has a weight of 1).")
  flags.DEFINE_float("base_learning_rate", 0.01, "Which learning rate to start with.")
  flags.DEFINE_float("learning_rate_decay", 0.95, "Learning rate decay factor to be applied every learning_rate_decay_examples.")
  flags.DEFINE_float("learning_rate_decay_examples", 4000000, "Multiply current learning rate by learning_rate_decay every learning_rate_decay_examples.")
  flags.DEFINE_integer("num_epochs", 5, "How many passes to make over the dataset before halting training.")
  flags.DEFINE_integer("max_steps", None, "The maximum number of iterations of the training loop.")
  flags.DEFINE_float("keep_checkpoint_every_n_hours", 1.0, "How many hours before saving a new checkpoint")
  flags.DEFINE_integer("keep_checkpoint_interval", 15, "How many minutes to wait before saving a new checkpoint")
  flags.DEFINE_integer("num_readers", 8, "How many batches the background input thread may run ahead of the training step "
                       "(the reference's queue-runner thread count, wh/train.py:199-209); 0 = read synchronously.")
  flags.DEFINE_integer("shuffle_seed", 0, "Seed of the input shuffle (file order per epoch + a 5*batch_size record buffer, "
                       "as string_input_producer(shuffle=True) + shuffle_batch_join do in the reference).")
  flags.DEFINE_string("dp_gradient_dtype", "float32", "Wire format of the data-parallel gradient exchange of the NetVLAD trainers: "
                      "float32 (the reference's arithmetic: towers sum fp32 gradients, wh/train.py:461-466) or bfloat16 (half the "
                      "bytes on NVLink, every rank's gradient rounded to 8 significant bits before the sum).")
  flags.DEFINE_bool("shuffle_input", True, "Shuffle the training input like the reference does; False replays the files in sorted order.")
  flags.DEFINE_string("optimizer", "AdamOptimizer", "What optimizer class to use.")
  flags.DEFINE_float("clip_gradient_norm", 1.0, "Norm to clip gradients to.")
  flags.DEFINE_bool("log_device_placement", False, "Whether to write the device on which every op will run into the logs on startup.")
  flags.DEFINE_integer("recall_at_n", 100, "N in recall@N.")
  flags.DEFINE_bool("dropout", False, "Whether to consider dropout")
  flags.DEFINE_float("keep_prob", 1.0, "probability to keep output (used in dropout, keep it unchanged in validationg and test)")
  flags.DEFINE_float("noise_level", 0.0, "standard deviation of noise (added to hidden nodes)")
  # the boosting / distillation pipeline of the reference (wh/train.py:104-137): accepted so that its command lines parse;
  # enabling any of them is refused (outside SURVEY.md §8)
  flags.DEFINE_bool("reweight", False, "Whether to reweight samples (boosting pipeline; not built)")
  flags.DEFINE_string("sample_vocab_file", "", "Where the vocabulary of the sample weights is (boosting pipeline; not built)")
  flags.DEFINE_string("sample_freq_file", "", "Where the sample weights are (boosting pipeline; not built)")
  flags.DEFINE_bool("distillation_features", False, "If set, *DistillationFeatureReader will be used (not built)")
  flags.DEFINE_bool("distillation_as_input", False, "If set, distillation_predictions will be given to the model (not built)")
  flags.DEFINE_bool("distillation_as_boosting", False, "If set, boosting via distillation predictions (not built)")
  flags.DEFINE_integer("distillation_type", 0, "Type of distillation, options are 0, 1 and 2 (not built)")
  flags.DEFINE_float("distillation_percent", 0.0, "If larger than 0, final_loss = distillation_loss * percent + normal_loss * (1.0 - percent) (not built)")


def get_reader():
  """wh/train.py:730-748."""
  feature_names, feature_sizes = utils.GetListOfFeatureNamesAndSizes(FLAGS.feature_names, FLAGS.feature_sizes)
  if FLAGS.frame_features:
    return readers.YT8MFrameFeatureReader(feature_names=feature_names, feature_sizes=feature_sizes)
  return readers.YT8MAggregatedFeatureReader(feature_names=feature_names, feature_sizes=feature_sizes)


class Trainer(object):
  """A Trainer to train a Tensorflow graph -- here: a HeadTrainer driving the CUDA training step
  (wh/train.py:482-728 keeps the same responsibilities: resume, loop, log line, checkpoints)."""

  def __init__(self, model_name, reader, train_dir, is_master=True):
    self.model_name, self.reader, self.train_dir, self.is_master = model_name, reader, train_dir, is_master

  def remove_training_directory(self, train_dir):
    """wh/train.py:641-652."""
    try:
      logging.info("Removing existing train directory.")
      shutil.rmtree(train_dir)
    except OSError:
      logging.error("Failed to delete directory " + train_dir + " when starting a new model. Please delete it manually and try again.")

  def build_model(self, in_dim):
    import yt8m_trainer
    model_cls = utils.find_class_by_name(self.model_name, [frame_level_models, video_level_models])
    if FLAGS.optimizer != "AdamOptimizer":
      raise NotImplementedError("only --optimizer=AdamOptimizer (the reference default) is built")
    self.multitask = None
    if FLAGS.multitask:
      # wh/train.py:394-413: the loss takes (predictions, support_predictions, labels); built for the two chain models, whose
      # training scripts are the ones that pass --multitask=True --label_loss=MultiTaskCrossEntropyLoss
      if FLAGS.label_loss != "MultiTaskCrossEntropyLoss" or model_cls not in (video_level_models.ChainMoeModel,
                                                                              video_level_models.DeepCombineChainModel):
        raise NotImplementedError("--multitask is built for --label_loss=MultiTaskCrossEntropyLoss with ChainMoeModel / "
                                  "DeepCombineChainModel")
      self.multitask = losses.MultiTaskCrossEntropyLoss()
    elif FLAGS.label_loss != "CrossEntropyLoss":
      raise NotImplementedError("only --label_loss=CrossEntropyLoss (the reference default; MultiTaskCrossEntropyLoss with "
                                "--multitask) is built")
    if model_cls in (frame_level_models.NetVLADModel, frame_level_models.GatedNetVLADModel):
      if FLAGS.netvlad_add_batch_norm or FLAGS.video_level_classifier_model != "MoeModel":
        raise NotImplementedError("train.py --model=%s: the CUDA training step is built for "
                                  "--netvlad_add_batch_norm=False with --video_level_classifier_model=MoeModel" % self.model_name)
      if FLAGS.dp_gradient_dtype not in ("float32", "bfloat16"):
        raise ValueError("--dp_gradient_dtype must be float32 or bfloat16")
      t = yt8m_trainer.NetVLADTrainer(in_dim, clusters=FLAGS.netvlad_cluster_size, hidden=FLAGS.netvlad_hidden_size,
                                      vocab=self.reader.num_classes, mixtures=FLAGS.moe_num_mixtures, relu=FLAGS.netvlad_relu,
                                      gating=model_cls is frame_level_models.GatedNetVLADModel)
      t.wire_dtype = torch.bfloat16 if FLAGS.dp_gradient_dtype == "bfloat16" else None
      return t
    if model_cls in (frame_level_models.LstmModel, frame_level_models.LstmMemoryModel):
      if FLAGS.video_level_classifier_model != "MoeModel":
        raise NotImplementedError("train.py --model=%s: the CUDA training step is built for "
                                  "--video_level_classifier_model=MoeModel" % self.model_name)
      return yt8m_trainer.LstmTrainer(in_dim, hidden=int(FLAGS.lstm_cells), layers=FLAGS.lstm_layers, vocab=self.reader.num_classes,
                                      mixtures=FLAGS.moe_num_mixtures, memory=model_cls is frame_level_models.LstmMemoryModel)
    if model_cls in (frame_level_models.LstmAttentionMaxPoolingModel, frame_level_models.LstmMultiAttentionModel):
      multi = model_cls is frame_level_models.LstmMultiAttentionModel
      if multi and FLAGS.video_level_classifier_model != "MoeModel":
        raise NotImplementedError("train.py --model=LstmMultiAttentionModel: built for --video_level_classifier_model=MoeModel")
      return yt8m_trainer.LstmAttentionTrainer(in_dim, hidden=int(FLAGS.lstm_cells), layers=FLAGS.lstm_layers,
                                               heads=FLAGS.attention_size if multi else FLAGS.lstm_attentions,
                                               vocab=self.reader.num_classes, mixtures=FLAGS.moe_num_mixtures,
                                               kind="multi" if multi else "max_pooling")
    if model_cls is frame_level_models.DbofModel:
      if FLAGS.dbof_pooling_method != "max" or FLAGS.video_level_classifier_model != "MoeModel":
        raise NotImplementedError("train.py --model=DbofModel: the CUDA training step is built for "
                                  "--dbof_pooling_method=max and --video_level_classifier_model=MoeModel")
      t = yt8m_trainer.DbofTrainer(in_dim, cluster_size=FLAGS.dbof_cluster_size, hidden=FLAGS.dbof_hidden_size,
                                   iterations=FLAGS.iterations, vocab=self.reader.num_classes, mixtures=FLAGS.moe_num_mixtures,
                                   batch_norm=FLAGS.dbof_add_batch_norm)
      t.sample_random_frames = FLAGS.sample_random_frames
      return t
    if model_cls is frame_level_models.AttentionModel:
      if FLAGS.video_level_classifier_model != "MoeExtendModel":
        raise NotImplementedError("train.py --model=AttentionModel: the CUDA training step is built for "
                                  "--video_level_classifier_model=MoeExtendModel (the reference scripts' pairing)")
      return yt8m_trainer.AttentionTrainer(in_dim, heads=FLAGS.moe_num_extend, vocab=self.reader.num_classes,
                                           mixtures=FLAGS.moe_num_mixtures)
    if model_cls is video_level_models.ChainMoeModel:
      return yt8m_trainer.ChainMoeTrainer(in_dim, vocab=self.reader.num_classes, mixtures=FLAGS.moe_num_mixtures,
                                          num_supports=FLAGS.num_supports)
    if model_cls is video_level_models.DeepCombineChainModel:
      if FLAGS.deep_chain_relu_type != "relu":
        raise NotImplementedError("train.py --model=DeepCombineChainModel: built with --deep_chain_relu_type=relu")
      return yt8m_trainer.DeepCombineChainTrainer(in_dim, vocab=self.reader.num_classes, mixtures=FLAGS.moe_num_mixtures,
                                                  layers=FLAGS.deep_chain_layers, relu_cells=FLAGS.deep_chain_relu_cells)
    if model_cls is video_level_models.LogisticModel:
      kind = "logistic"
    elif model_cls is video_level_models.MoeModel:
      kind = "moe"
    else:
      raise NotImplementedError(
          "train.py: the CUDA training step is built for LogisticModel, MoeModel, NetVLADModel, GatedNetVLADModel, "
          "LstmModel, LstmMemoryModel, LstmAttentionMaxPoolingModel, LstmMultiAttentionModel, AttentionModel (+ MoeExtendModel) "
          "DbofModel (bias form), ChainMoeModel and DeepCombineChainModel this round; "
          "%s runs forward-only (eval.py / inference.py)" % self.model_name)
    return yt8m_trainer.HeadTrainer(kind, in_dim, self.reader.num_classes, mixtures=FLAGS.moe_num_mixtures)

  def run(self, start_new_model=False):
    rank, world = yt8m_dp.rank(), yt8m_dp.world_size()
    if self.is_master and start_new_model:
      self.remove_training_directory(self.train_dir)
    in_dim = sum(self.reader.feature_sizes)
    trainer = self.build_model(in_dim)
    transformer = utils.find_class_by_name(FLAGS.feature_transformer, [feature_transform])()
    latest = None if start_new_model else utils.latest_checkpoint(self.train_dir)
    if latest:
      logging.info("Restoring from checkpoint %s", latest)
      ck = utils.load_checkpoint(latest)
      trainer.import_state({k: v.cuda() for k, v in ck["variables"].items()})
      if ck.get("optimizer"):
        trainer.adam_m.copy_(ck["optimizer"]["m"])
        trainer.adam_v.copy_(ck["optimizer"]["v"])
      trainer.global_step = ck["global_step"]
    else:
      import yt8m_ops as ops
      logging.info("No checkpoint file found. Building a new model.")
      ops.get_store().reset(seed=9)
      model = utils.find_class_by_name(self.model_name, [frame_level_models, video_level_models])()
      if FLAGS.frame_features:                                              # creates the variables
        model.create_model(torch.zeros((2, self.reader.max_frames, in_dim), dtype=torch.bfloat16, device="cuda"),
                           vocab_size=self.reader.num_classes, num_frames=torch.ones(2, dtype=torch.int32, device="cuda"))
      else:
        model.create_model(torch.zeros((2, in_dim), device="cuda"), vocab_size=self.reader.num_classes)
      trainer.import_state({k: v.value for k, v in ops.get_store().vars.items()})
    logging.info("Entering training loop.")
    steps, last_save = 0, time.time()
    # frame-level batches travel as readers.PackedFrames: only the real frames cross PCIe (the padding is made on the GPU)
    packed = {"packed": True} if FLAGS.frame_features else {}
    # Input pipeline (wh/train.py:199-209): the file order is re-shuffled every epoch and records pass through a shuffle
    # buffer of 5 * batch_size; every rank derives the SAME record order from --shuffle_seed and parses only its own rows of
    # each global batch; a background thread keeps --num_readers batches ahead of the step.  Like the reference's
    # shuffle_batch_join (allow_smaller_final_batch=False) a final batch with fewer videos than ranks is dropped -- every
    # rank sees the same global row count, so they all skip it together and nobody waits alone in the all-reduce.
    batches = self.reader.prepare_reader(FLAGS.train_data_pattern, FLAGS.batch_size, FLAGS.num_epochs, shuffle=FLAGS.shuffle_input,
                                         seed=FLAGS.shuffle_seed, shard=(rank, world), **packed)
    if FLAGS.num_readers > 0:
      batches = readers.prefetch(batches, depth=min(FLAGS.num_readers, 4))
    for video_ids, feats, labels, num_frames, n_global in batches:
      if n_global < world:
        logging.info("dropping a final batch of %d videos (fewer than the %d ranks)", n_global, world)
        continue
      steps += 1
      t0 = time.time()
      x, nf = transformer.transform(feats.cuda(non_blocking=True), num_frames)
      y = labels.cuda(non_blocking=True).float()
      frame_args = (nf.to("cuda", torch.int32),) if FLAGS.frame_features else ()
      extra = {}
      if self.multitask is not None:
        extra = {"support_labels": torch.from_numpy(self.multitask.get_support(labels)).cuda(non_blocking=True),
                 "support_loss_percent": FLAGS.support_loss_percent}
      p = trainer.step(x, *frame_args, y, FLAGS.base_learning_rate, FLAGS.learning_rate_decay, FLAGS.learning_rate_decay_examples,
                       FLAGS.clip_gradient_norm, FLAGS.regularization_penalty, global_batch=n_global, **extra)
      if self.is_master:
        # the log line's metrics (wh/train.py:578-591) from the per-video top-32 extracted on the GPU: 64 numbers per video
        # cross PCIe instead of 4716, and the host loop is O(B * 32)
        lv = labels.numpy().astype("float32")
        kk = min(32, p.shape[1])
        ti, tv = nat.topk_rows(p, kk)
        hit1, perr, gap = eval_util.step_metrics_from_topk(tv.cpu().numpy(), ti.cpu().numpy(), lv, top_k=min(20, kk))
        if perr is None:                                                    # a video with more than 32 labels: full predictions
          perr = eval_util.calculate_precision_at_equal_recall_rate(p.cpu().numpy(), lv)
        loss_val = float(trainer.last["label_loss_local"])
        seconds = time.time() - t0
        logging.info("training step " + str(trainer.global_step) + "| Hit@1: " + ("%.2f" % hit1) +
                     " PERR: " + ("%.2f" % perr) + " GAP: " +
                     ("%.2f" % gap) + " Recall@%d: " % FLAGS.recall_at_n + "N/A" + " Loss: " + str(loss_val) +
                     " Examples/sec: %.1f" % (n_global / max(seconds, 1e-9)))
        if time.time() - last_save > FLAGS.keep_checkpoint_interval * 60:
          self.save(trainer)
          last_save = time.time()
      if FLAGS.max_steps is not None and steps > FLAGS.max_steps:
        logging.info("Done training -- max_steps limit reached.")
        break
    else:
      logging.info("Done training -- epoch limit reached.")
    if self.is_master:
      self.save(trainer)
    logging.info("Exited training loop.")
    return trainer

  def save(self, trainer):
    path = utils.save_checkpoint(self.train_dir, trainer.global_step, trainer.export_state(),
                                 {"m": trainer.adam_m.cpu(), "v": trainer.adam_v.cpu()},
                                 {"model": self.model_name, "moe_num_mixtures": FLAGS.moe_num_mixtures,
                                  "feature_names": FLAGS.feature_names, "feature_sizes": FLAGS.feature_sizes})
    logging.info("Saved checkpoint %s", path)


def main(unused_argv=None):
  logging.basicConfig(level=logging.INFO, format="%(levelname)s:%(message)s")
  rest = FLAGS.parse()
  if rest:
    logging.warning("ignoring positional arguments %s", rest)
  off = [n for n in ("reweight", "distillation_features", "distillation_as_input", "distillation_as_boosting") if getattr(FLAGS, n)]
  if off or FLAGS.distillation_percent > 0 or FLAGS.distillation_type != 0:
    raise NotImplementedError("train.py: the boosting / distillation pipeline (--reweight, --distillation_*) is outside the hot path")
  if FLAGS.dropout or FLAGS.keep_prob != 1.0 or FLAGS.noise_level != 0.0:
    # accepted for command-line compatibility, but never silently ignored: a run that asks for them would train a different,
    # unregularised model
    raise NotImplementedError("train.py: --dropout / --keep_prob / --noise_level (training-time regularisers of models outside "
                              "SURVEY.md §8) are not built; leave them at their defaults")
  if not torch.cuda.is_available():
    raise SystemExit("train.py: no CUDA device; the yt8m_b200 path has no CPU fallback")
  rank, world, local_rank = yt8m_dp.init_from_env()
  torch.cuda.set_device(local_rank)
  logging.info("rank %d / %d: torch %s", rank, world, torch.__version__)
  Trainer(FLAGS.model, get_reader(), FLAGS.train_dir, is_master=(rank == 0)).run(start_new_model=FLAGS.start_new_model)


if __name__ == "__main__":
  main(sys.argv)
