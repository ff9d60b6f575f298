"""CPU tests of the host-side data path: TFRecord framing (masked CRC-32C), Example / SequenceExample wire
decoding, reader batching / padding / truncation semantics (wh/readers.py), checkpoint helpers and
find_class_by_name's error behaviour (wh/train.py:212-215)."""
import os

import numpy as np
import pytest
import torch

import readers
import utils


def test_crc32c_known_answers():
  # RFC 3720 test vectors for CRC-32C (Castagnoli)
  assert readers.crc32c(b"") == 0
  assert readers.crc32c(bytes(32)) == 0x8A9136AA
  assert readers.crc32c(bytes([0xFF] * 32)) == 0x62A8AB43
  assert readers.crc32c(bytes(range(32))) == 0x46DD794E
  assert readers.crc32c(b"123456789") == 0xE3069283


def _video_records(n, classes=20):
  return [readers.encode_example({"video_id": ("bytes", [b"vid%d" % i]), "labels": ("int64", [i % classes, (i + 5) % classes, classes + 3]),
                                  "mean_rgb": ("float", np.arange(8, dtype=np.float32) + i),
                                  "mean_audio": ("float", np.full(4, i, dtype=np.float32))}) for i in range(n)]


def test_aggregated_reader_roundtrip(tmp_path):
  p = str(tmp_path / "a.tfrecord")
  readers.write_tfrecord(p, _video_records(5))
  r = readers.YT8MAggregatedFeatureReader(num_classes=20, feature_names=["mean_rgb", "mean_audio"], feature_sizes=[8, 4])
  batches = list(r.prepare_reader(str(tmp_path / "*.tfrecord"), batch_size=3, verify_crc=True))
  assert [len(b[0]) for b in batches] == [3, 2]
  ids, f, l, nf = batches[0]
  assert ids == [b"vid0", b"vid1", b"vid2"] and f.shape == (3, 12) and f.dtype == torch.float32
  assert f[2, :8].tolist() == list(np.arange(8) + 2.0) and f[2, 8:].tolist() == [2.0] * 4
  assert l.shape == (3, 20) and l.sum(dim=1).tolist() == [2, 2, 2]        # labels >= num_classes are dropped
  assert nf.tolist() == [1, 1, 1]
  two_epochs = sum(len(b[0]) for b in r.prepare_reader(p, batch_size=4, num_epochs=2))
  assert two_epochs == 10


def test_frame_reader_pads_and_truncates(tmp_path):
  recs = [readers.encode_sequence_example({"video_id": ("bytes", [b"v%d" % i]), "labels": ("int64", [1, 2])},
                                          {"rgb": [("bytes", [bytes([j] * 6)]) for j in range(3 + 2 * i)],
                                           "audio": [("bytes", [bytes([9] * 2)]) for j in range(3 + 2 * i)]}) for i in range(2)]
  p = str(tmp_path / "f.tfrecord")
  readers.write_tfrecord(p, recs)
  r = readers.YT8MFrameFeatureReader(num_classes=20, feature_names=["rgb", "audio"], feature_sizes=[6, 2], max_frames=4)
  (ids, f, l, nf), = list(r.prepare_reader(p, batch_size=8))
  assert f.dtype == torch.uint8 and f.shape == (2, 4, 8)
  assert nf.tolist() == [3, 4]                                              # 5 frames truncated to max_frames = 4
  assert f[0, :, 0].tolist() == [0, 1, 2, 0] and f[0, :, 7].tolist() == [9, 9, 9, 0]   # zero padding past num_frames
  assert f[1, :, 0].tolist() == [0, 1, 2, 3]
  assert l[0].nonzero().flatten().tolist() == [1, 2]


def test_corrupt_and_missing_files(tmp_path):
  p = str(tmp_path / "a.tfrecord")
  readers.write_tfrecord(p, _video_records(2))
  data = bytearray(open(p, "rb").read())
  data[20] ^= 0xFF
  open(p, "wb").write(bytes(data))
  with pytest.raises(IOError):
    list(readers.tfrecord_iterator(p, verify=True))
  open(p, "wb").write(bytes(data[:30]))
  with pytest.raises(IOError):
    list(readers.tfrecord_iterator(p))
  r = readers.YT8MAggregatedFeatureReader(num_classes=20, feature_names=["mean_rgb"], feature_sizes=[8])
  with pytest.raises(IOError):                                               # same failure as wh/train.py:193-195
    list(r.prepare_reader(str(tmp_path / "nothing*.tfrecord")))
  with pytest.raises(AssertionError):
    readers.YT8MFrameFeatureReader(feature_names=["rgb", "audio"], feature_sizes=[1024])
  with pytest.raises(NotImplementedError):
    readers.BaseReader().prepare_reader(None)


def test_utils(tmp_path):
  import video_level_models, frame_level_models
  assert utils.find_class_by_name("MoeModel", [frame_level_models, video_level_models]) is video_level_models.MoeModel
  with pytest.raises(StopIteration):
    utils.find_class_by_name("NoSuchModel", [frame_level_models, video_level_models])
  assert utils.GetListOfFeatureNamesAndSizes("rgb, audio", "1024,128") == (["rgb", "audio"], [1024, 128])
  q = utils.Dequantize(np.array([0.0, 255.0]))
  assert np.allclose(q, [4 / 512 - 2, 4 + 4 / 512 - 2])
  d = str(tmp_path / "ckpt")
  for step in (10, 20, 30, 40):
    utils.save_checkpoint(d, step, {"w": torch.ones(2) * step})
  assert sorted(os.listdir(d)) == ["model.ckpt-20", "model.ckpt-30", "model.ckpt-40"]       # max_to_keep = 3
  assert utils.latest_checkpoint(d).endswith("model.ckpt-40")
  assert utils.load_checkpoint(utils.latest_checkpoint(d))["variables"]["w"].tolist() == [40.0, 40.0]
  info = utils.FormatEpochInfo({"epoch_id": 7, "avg_hit_at_one": 0.5, "avg_perr": 0.25, "aps": [0.5, 1.0], "gap": 0.75, "avg_loss": 3.0})
  assert info.startswith("epoch/eval number 7 | Avg_Hit@1: 0.500 | Avg_PERR: 0.250 | MAP: 0.750 | GAP: 0.750")


def test_packed_frames_equal_the_padded_batch(tmp_path):
  """packed=True yields readers.PackedFrames: the same batch without the zero padding (only real frames cross PCIe)."""
  rs = np.random.RandomState(4)
  recs, lens = [], [3, 9, 1, 6, 12]                                   # 12 > max_frames: truncated like the padded reader
  for i, n in enumerate(lens):
    fr = rs.randint(0, 256, (n, 16)).astype(np.uint8)
    recs.append(readers.encode_sequence_example(
        {"video_id": ("bytes", [b"p%d" % i]), "labels": ("int64", [i])},
        {"rgb": [("bytes", [fr[j, :8].tobytes()]) for j in range(n)], "audio": [("bytes", [fr[j, 8:].tobytes()]) for j in range(n)]}))
  p = str(tmp_path / "f.tfrecord")
  readers.write_tfrecord(p, recs)
  r = readers.YT8MFrameFeatureReader(num_classes=10, feature_names=["rgb", "audio"], feature_sizes=[8, 8], max_frames=10)
  (ids_a, padded, lab_a, nf_a), = list(r.prepare_reader(p, batch_size=8))
  (ids_b, packed, lab_b, nf_b), = list(r.prepare_reader(p, batch_size=8, packed=True))
  assert isinstance(packed, readers.PackedFrames) and ids_a == ids_b and torch.equal(lab_a, lab_b) and torch.equal(nf_a, nf_b)
  assert nf_b.tolist() == [3, 9, 1, 6, 10]
  assert packed.shape == tuple(padded.shape) and packed.data.shape == (29, 16) and packed.offsets.tolist() == [0, 3, 12, 13, 19]
  assert torch.equal(packed.to_padded(), padded)
  assert packed.nbytes() == 29 * 16 + 5 * 4 + 5 * 8
  # slicing by video = data-parallel sharding
  sh = packed[1:4]
  assert sh.shape == (3, 10, 16) and sh.offsets.tolist() == [0, 9, 10] and torch.equal(sh.to_padded(), padded[1:4])
  assert packed[2:2].shape[0] == 0
  # synthetic batches: dropping the padding and restoring it is the identity
  again = readers.PackedFrames.from_padded(padded, nf_a)
  assert torch.equal(again.data, packed.data) and torch.equal(again.offsets, packed.offsets)


def test_pre_ensemble_prediction_records(tmp_path):
  """The wire format of wh/inference-pre-ensemble.py:291-308: predictions-%04d.tfrecord files whose Examples hold video_id,
  the indices of the positive labels and the full float prediction vector -- written with our codecs, read back."""
  import importlib.util
  spec = importlib.util.spec_from_file_location(
      "inference_pre_ensemble", os.path.join(os.path.dirname(readers.__file__), "inference-pre-ensemble.py"))
  mod = importlib.util.module_from_spec(spec)
  try:
    spec.loader.exec_module(mod)
  except ImportError as e:                      # the module imports the CUDA binding (built library required)
    pytest.skip(str(e))
  rs = np.random.RandomState(9)
  n, v = 5, 12
  ids = [b"vid%d" % i for i in range(n)]
  labels = rs.rand(n, v) < 0.3
  preds = rs.rand(n, v).astype(np.float32)
  path = mod.write_to_record(str(tmp_path), ids, labels, preds, 3, n)
  assert path.endswith("predictions-0003.tfrecord")
  recs = list(readers.tfrecord_iterator(path, True))
  assert len(recs) == n
  for i, rec in enumerate(recs):
    ex = readers.parse_example(rec)
    assert ex["video_id"][1][0] == ids[i]
    assert list(ex["labels"][1]) == list(np.nonzero(labels[i])[0])
    assert np.array_equal(np.asarray(ex["predictions"][1], dtype=np.float32), preds[i])


def test_shuffled_sharded_input_pipeline(tmp_path):
  """ADVICE r1 (train.py input): the record order is shuffled per epoch through a 5 * batch buffer, is a pure function of
  the seed (every data-parallel rank derives the same global batches), differs between epochs and from the file order; a
  rank parses only its own rows and the ranks' shards tile every batch; the short tail is dropped on request."""
  for f in range(3):
    readers.write_tfrecord(str(tmp_path / ("s%d.tfrecord" % f)), _video_records(40)[f * 13:(f + 1) * 13])
  pat = str(tmp_path / "s*.tfrecord")
  r = readers.YT8MAggregatedFeatureReader(num_classes=20, feature_names=["mean_rgb", "mean_audio"], feature_sizes=[8, 4])
  plain = [i for b in r.prepare_reader(pat, batch_size=8) for i in b[0]]
  a = [b[0] for b in r.prepare_reader(pat, batch_size=8, num_epochs=2, shuffle=True, seed=3)]
  b = [b[0] for b in r.prepare_reader(pat, batch_size=8, num_epochs=2, shuffle=True, seed=3)]
  c = [b[0] for b in r.prepare_reader(pat, batch_size=8, num_epochs=2, shuffle=True, seed=4)]
  flat = [i for ids in a for i in ids]
  assert a == b and a != c                                           # a function of the seed only
  assert sorted(flat) == sorted(plain * 2)                           # every record exactly once per epoch
  assert flat[:39] != plain and flat[:39] != flat[39:]               # shuffled, and differently in the second epoch
  assert [len(x) for x in a] == [8] * 9 + [6]
  assert [len(x[0]) for x in r.prepare_reader(pat, batch_size=8, num_epochs=2, shuffle=True, seed=3, drop_remainder=True)] == [8] * 9
  world = 4
  per_rank = [list(r.prepare_reader(pat, batch_size=8, num_epochs=2, shuffle=True, seed=3, shard=(k, world))) for k in range(world)]
  for step, ids in enumerate(a):
    got = [i for k in range(world) for i in per_rank[k][step][0]]
    assert got == ids and all(per_rank[k][step][4] == len(ids) for k in range(world))
  # a tail batch with fewer videos than ranks: some shards are EMPTY but well-formed, and every rank sees n_global
  tail = [list(r.prepare_reader(pat, batch_size=37, shard=(k, 8)))[-1] for k in range(8)]
  assert [len(t[0]) for t in tail] == [1, 1, 0, 0, 0, 0, 0, 0] and all(t[4] == 2 for t in tail)
  assert tail[5][1].shape == (0, 12) and tail[5][2].shape == (0, 20)
  assert readers.shard_range(9, (7, 8)) == (8, 9) and readers.shard_range(3, (5, 8)) == (3, 3)


def test_prefetch_thread_and_missing_labels(tmp_path):
  recs = [readers.encode_example({"video_id": ("bytes", [b"u%d" % i]), "mean_rgb": ("float", np.ones(8, dtype=np.float32)),
                                  "mean_audio": ("float", np.zeros(4, dtype=np.float32))}) for i in range(5)]   # no `labels` feature
  p = str(tmp_path / "u.tfrecord")
  readers.write_tfrecord(p, recs)
  r = readers.YT8MAggregatedFeatureReader(num_classes=20, feature_names=["mean_rgb", "mean_audio"], feature_sizes=[8, 4])
  out = list(readers.prefetch(r.prepare_reader(p, batch_size=2), depth=2))
  assert [len(b[0]) for b in out] == [2, 2, 1] and all(int(b[2].sum()) == 0 for b in out)   # tf.VarLenFeature: missing = empty
  seq = [readers.encode_sequence_example({"video_id": ("bytes", [b"w"])}, {"rgb": [("bytes", [bytes([1] * 6)])] * 3})]
  p2 = str(tmp_path / "w.tfrecord")
  readers.write_tfrecord(p2, seq)
  fr = readers.YT8MFrameFeatureReader(num_classes=20, feature_names=["rgb"], feature_sizes=[6], max_frames=4)
  (ids, f, l, nf), = list(fr.prepare_reader(p2, batch_size=4))
  assert nf.tolist() == [3] and int(l.sum()) == 0

  def boom():
    yield 1
    raise ValueError("reader failed")
  with pytest.raises(ValueError):
    list(readers.prefetch(boom()))
