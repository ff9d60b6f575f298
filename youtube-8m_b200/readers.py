"""YouTube-8M readers with the reference's classes (wh/readers.py:58-459): TFRecord files of
tf.train.Example (video-level: ``mean_rgb`` / ``mean_audio`` float features) or tf.train.SequenceExample
(frame-level: per-frame uint8-quantised ``rgb`` / ``audio`` byte strings), ``video_id`` + sparse ``labels``.

The reference parses these inside TensorFlow queue runners; here the framing (length + masked CRC-32C) and
the protobuf wire format are decoded on the host with no TensorFlow dependency, and batches are handed to
the GPU still QUANTISED (uint8 [B, max_frames, D], zero padded -- wh/readers.py:21-56,186): de-quantisation
(wh/utils.py:23-38) and the L2 normalisation are fused into one kernel behind DefaultTransformer.
"""
import glob
import random
import struct

import numpy as np
import torch

# ------------------------------------------------------------------------------------------------
# TFRecord framing
# ------------------------------------------------------------------------------------------------
_CRC_TABLE = None


def _crc32c_table():
  global _CRC_TABLE
  if _CRC_TABLE is None:
    t = []
    for i in range(256):
      c = i
      for _ in range(8):
        c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
      t.append(c)
    _CRC_TABLE = t
  return _CRC_TABLE


def crc32c(data):
  t = _crc32c_table()
  c = 0xFFFFFFFF
  for b in data:
    c = t[(c ^ b) & 0xFF] ^ (c >> 8)
  return c ^ 0xFFFFFFFF


def masked_crc32c(data):
  c = crc32c(data)
  return ((((c >> 15) | (c << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


def tfrecord_iterator(path, verify=False):
  """Yields the serialized records of one TFRecord file."""
  with open(path, "rb") as f:
    while True:
      head = f.read(12)
      if not head:
        return
      if len(head) < 12:
        raise IOError("truncated TFRecord header in %s" % path)
      (length,), (len_crc,) = struct.unpack("<Q", head[:8]), struct.unpack("<I", head[8:])
      if verify and masked_crc32c(head[:8]) != len_crc:
        raise IOError("corrupt TFRecord length CRC in %s" % path)
      data = f.read(length)
      tail = f.read(4)
      if len(data) < length or len(tail) < 4:
        raise IOError("truncated TFRecord in %s" % path)
      if verify and masked_crc32c(data) != struct.unpack("<I", tail)[0]:
        raise IOError("corrupt TFRecord data CRC in %s" % path)
      yield data


def write_tfrecord(path, records):
  with open(path, "wb") as f:
    for data in records:
      head = struct.pack("<Q", len(data))
      f.write(head + struct.pack("<I", masked_crc32c(head)) + data + struct.pack("<I", masked_crc32c(data)))


# ------------------------------------------------------------------------------------------------
# protobuf wire format of tf.train.Example / SequenceExample (field numbers from example.proto / feature.proto)
# ------------------------------------------------------------------------------------------------

def _varint(buf, pos):
  res, shift = 0, 0
  while True:
    b = buf[pos]
    pos += 1
    res |= (b & 0x7F) << shift
    if not b & 0x80:
      return res, pos
    shift += 7


def _fields(buf):
  """Yields (field_number, wire_type, value) of one message; length-delimited values are memoryviews."""
  pos, n = 0, len(buf)
  while pos < n:
    key, pos = _varint(buf, pos)
    fnum, wt = key >> 3, key & 7
    if wt == 0:
      val, pos = _varint(buf, pos)
    elif wt == 2:
      ln, pos = _varint(buf, pos)
      val = buf[pos:pos + ln]
      pos += ln
    elif wt == 5:
      val = buf[pos:pos + 4]
      pos += 4
    elif wt == 1:
      val = buf[pos:pos + 8]
      pos += 8
    else:
      raise ValueError("unsupported protobuf wire type %d" % wt)
    yield fnum, wt, val


def _parse_feature(buf):
  """Feature { bytes_list = 1; float_list = 2; int64_list = 3 } -> ("bytes"|"float"|"int64", values)."""
  for fnum, _, val in _fields(buf):
    if fnum == 1:
      return "bytes", [bytes(v) for f, _, v in _fields(val) if f == 1]
    if fnum == 2:
      out = []
      for f, wt, v in _fields(val):
        if f == 1:
          out.append(np.frombuffer(v, dtype="<f4"))          # packed (wt 2) or single fixed32 (wt 5)
      return "float", np.concatenate(out) if out else np.zeros(0, np.float32)
    if fnum == 3:
      out = []
      for f, wt, v in _fields(val):
        if f != 1:
          continue
        if wt == 0:
          out.append(v)
        else:
          p = 0
          while p < len(v):
            x, p = _varint(v, p)
            out.append(x)
      return "int64", np.asarray(out, dtype=np.int64)
  return "none", []


def _parse_features(buf):
  """Features { map<string, Feature> feature = 1 }"""
  out = {}
  for fnum, _, entry in _fields(buf):
    if fnum != 1:
      continue
    key, feat = None, None
    for f, _, v in _fields(entry):
      if f == 1:
        key = bytes(v).decode("utf-8")
      elif f == 2:
        feat = _parse_feature(v)
    out[key] = feat
  return out


def parse_example(data):
  """tf.train.Example { Features features = 1 } -> {name: (kind, values)}"""
  buf = memoryview(data)
  for fnum, _, val in _fields(buf):
    if fnum == 1:
      return _parse_features(val)
  return {}


def parse_sequence_example(data):
  """SequenceExample { Features context = 1; FeatureLists feature_lists = 2 } -> (context, {name: [Feature...]})"""
  buf = memoryview(data)
  context, lists = {}, {}
  for fnum, _, val in _fields(buf):
    if fnum == 1:
      context = _parse_features(val)
    elif fnum == 2:
      for f, _, entry in _fields(val):                       # map<string, FeatureList>
        if f != 1:
          continue
        key, feats = None, []
        for g, _, v in _fields(entry):
          if g == 1:
            key = bytes(v).decode("utf-8")
          elif g == 2:
            feats = [_parse_feature(x) for h, _, x in _fields(v) if h == 1]
        lists[key] = feats
  return context, lists


# ---- encoders (synthetic data / tests) ---------------------------------------------------------------

def _enc_varint(x):
  out = bytearray()
  while True:
    b = x & 0x7F
    x >>= 7
    out.append(b | (0x80 if x else 0))
    if not x:
      return bytes(out)


def _ld(fnum, payload):
  return _enc_varint((fnum << 3) | 2) + _enc_varint(len(payload)) + payload


def _enc_feature(kind, values):
  if kind == "bytes":
    return _ld(1, b"".join(_ld(1, v) for v in values))
  if kind == "float":
    return _ld(2, _ld(1, np.asarray(values, dtype="<f4").tobytes()))
  return _ld(3, _ld(1, b"".join(_enc_varint(int(v)) for v in values)))


def _enc_features(feats):
  return b"".join(_ld(1, _ld(1, k.encode()) + _ld(2, _enc_feature(kind, vals))) for k, (kind, vals) in feats.items())


def encode_example(feats):
  return _ld(1, _enc_features(feats))


def encode_sequence_example(context, feature_lists):
  fl = b"".join(_ld(1, _ld(1, k.encode()) + _ld(2, b"".join(_ld(1, _enc_feature(kind, vals)) for kind, vals in lst)))
                for k, lst in feature_lists.items())
  return _ld(1, _enc_features(context)) + _ld(2, fl)


# ------------------------------------------------------------------------------------------------
# readers
# ------------------------------------------------------------------------------------------------

class BaseReader(object):
  """Inherit from this class when implementing new readers (wh/readers.py:58-63)."""

  def prepare_reader(self, unused_filename_queue):
    raise NotImplementedError()


def _files(pattern):
  files = sorted(glob.glob(pattern)) if isinstance(pattern, str) else list(pattern)
  if not files:
    # same failure mode as wh/train.py:193-195
    raise IOError("Unable to find training files. data_pattern='" + str(pattern) + "'.")
  return files


def _label_vector(feature, num_classes):
  """tf.VarLenFeature + sparse_to_indicator (wh/readers.py:101-107, 225-228): a missing or empty `labels` feature is an
  empty label set (unlabeled test-set records); ids at or beyond num_classes are dropped."""
  lab = np.zeros(num_classes, dtype=bool)
  if feature is not None and feature[1] is not None and len(feature[1]):
    idx = np.asarray(feature[1], dtype=np.int64)
    lab[idx[(idx >= 0) & (idx < num_classes)]] = True
  return lab


def record_batches(data_pattern, batch_size, num_epochs=1, verify_crc=False, shuffle=False, seed=0, shuffle_buffer=None,
                   drop_remainder=False):
  """Batches of RAW records.  shuffle=True reproduces the reference's input randomisation (wh/train.py:199-209 /
  wh/readers.py via tf.train.string_input_producer(shuffle=True) + shuffle_batch_join(capacity=5*batch_size)): the file
  list is re-shuffled every epoch and records pass through a shuffle buffer of `shuffle_buffer` (default 5 * batch_size)
  records.  The order depends only on (seed, file list, record counts), never on record contents, so every data-parallel
  rank that runs this generator with the same seed sees the same batches and can parse just its own rows.
  drop_remainder: the reference's shuffle_batch_join (allow_smaller_final_batch=False) never emits a short final batch."""
  rng = random.Random(seed)
  cap = (5 * batch_size if shuffle_buffer is None else shuffle_buffer) if shuffle else 0
  pool, batch = [], []
  for _ in range(num_epochs):
    files = _files(data_pattern)
    if shuffle:
      rng.shuffle(files)
    for path in files:
      for rec in tfrecord_iterator(path, verify_crc):
        if cap:
          if len(pool) < cap:
            pool.append(rec)
            continue
          j = rng.randrange(cap)
          rec, pool[j] = pool[j], rec
        batch.append(rec)
        if len(batch) == batch_size:
          yield batch
          batch = []
  rng.shuffle(pool)
  for rec in pool:
    batch.append(rec)
    if len(batch) == batch_size:
      yield batch
      batch = []
  if batch and not drop_remainder:
    yield batch


def shard_range(n_rows, shard):
  """Rows [lo, hi) of a batch owned by rank r of w (balanced: sizes differ by at most one)."""
  if shard is None:
    return 0, n_rows
  r, w = shard
  base, extra = divmod(n_rows, w)
  lo = r * base + min(r, extra)
  return lo, lo + base + (1 if r < extra else 0)


def prefetch(generator, depth=2):
  """Runs `generator` in a background thread, `depth` batches ahead of the consumer -- the stand-in for the reference's
  queue-runner threads (wh/train.py:199-209): TFRecord parsing overlaps the GPU step."""
  import queue
  import threading
  q = queue.Queue(maxsize=max(depth, 1))
  end = object()

  def work():
    try:
      for item in generator:
        q.put(item)
      q.put(end)
    except BaseException as e:                         # surfaces in the consumer
      q.put(e)

  threading.Thread(target=work, daemon=True).start()
  while True:
    item = q.get()
    if item is end:
      return
    if isinstance(item, BaseException):
      raise item
    yield item


class YT8MAggregatedFeatureReader(BaseReader):
  """Video-level Examples (wh/readers.py:66-125)."""

  def __init__(self, num_classes=4716, feature_sizes=(1024,), feature_names=("mean_inc3",)):
    assert len(feature_names) == len(feature_sizes), \
        "length of feature_names (={}) != length of feature_sizes (={})".format(len(feature_names), len(feature_sizes))
    self.num_classes, self.feature_sizes, self.feature_names = num_classes, list(feature_sizes), list(feature_names)

  def prepare_reader(self, data_pattern, batch_size=1024, num_epochs=1, verify_crc=False, shuffle=False, seed=0, shard=None,
                     drop_remainder=False, shuffle_buffer=None):
    """Generator of (video_ids [B], features fp32 [B, sum(sizes)], labels bool [B, C], num_frames ones [B]).
    shard=(rank, world): only this rank's rows of every batch are parsed and returned, and a 5th element carries the
    GLOBAL row count of the batch."""
    for recs in record_batches(data_pattern, batch_size, num_epochs, verify_crc, shuffle, seed, shuffle_buffer, drop_remainder):
      lo, hi = shard_range(len(recs), shard)
      ids, feats, labels = [], [], []
      for rec in recs[lo:hi]:
        ex = parse_example(rec)
        ids.append(ex["video_id"][1][0])
        feats.append(np.concatenate([ex[n][1][:s] for n, s in zip(self.feature_names, self.feature_sizes)]))
        labels.append(_label_vector(ex.get("labels"), self.num_classes))
      d = sum(self.feature_sizes)
      out = (ids, torch.from_numpy(np.stack(feats)) if feats else torch.zeros((0, d)),
             torch.from_numpy(np.stack(labels)) if labels else torch.zeros((0, self.num_classes), dtype=torch.bool),
             torch.ones(len(ids), dtype=torch.int32))
      yield out if shard is None else out + (len(recs),)


class PackedFrames(object):
  """A frame-level batch WITHOUT the reader's zero padding: `data` uint8 [sum(num_frames), D] holds only the real
  frames, video b = rows [offsets[b], offsets[b] + num_frames[b]).  It stands where the padded
  [B, max_frames, D] tensor of wh/readers.py:186 stood (same .shape, slicing by video, .cuda()), but only the real
  frames cross PCIe; DefaultTransformer.transform() expands it on the GPU (yt8m_frames_unpack_u8: de-quantise +
  L2-normalise + zero padding in one pass)."""

  dtype = torch.uint8

  def __init__(self, data, num_frames, max_frames, offsets=None):
    self.data, self.num_frames, self.max_frames = data, num_frames.to(torch.int32), int(max_frames)
    if offsets is None:
      offsets = torch.cumsum(self.num_frames.to(torch.int64), 0) - self.num_frames.to(torch.int64)
    self.offsets = offsets

  @property
  def shape(self):
    return (int(self.num_frames.shape[0]), self.max_frames, int(self.data.shape[1]))

  @property
  def is_cuda(self):
    return self.data.is_cuda

  def __len__(self):
    return int(self.num_frames.shape[0])

  def dim(self):
    return 3

  def __getitem__(self, idx):
    """Slice by video (data-parallel sharding of a batch: rank r keeps videos [lo, hi))."""
    if not isinstance(idx, slice):
      raise TypeError("PackedFrames can only be sliced by a range of videos")
    lo, hi, step = idx.indices(len(self))
    assert step == 1
    if hi <= lo:
      return PackedFrames(self.data[:0], self.num_frames[:0], self.max_frames)
    r0 = int(self.offsets[lo])
    r1 = int(self.offsets[hi - 1]) + int(self.num_frames[hi - 1])
    return PackedFrames(self.data[r0:r1], self.num_frames[lo:hi], self.max_frames, self.offsets[lo:hi] - r0)

  def pin_memory(self):
    return PackedFrames(self.data.pin_memory(), self.num_frames.pin_memory(), self.max_frames, self.offsets.pin_memory())

  def cuda(self, non_blocking=False):
    if self.is_cuda:
      return self
    return PackedFrames(self.data.cuda(non_blocking=non_blocking), self.num_frames.cuda(non_blocking=non_blocking), self.max_frames,
                        self.offsets.cuda(non_blocking=non_blocking))

  def nbytes(self):
    return self.data.numel() + self.num_frames.numel() * 4 + self.offsets.numel() * 8

  def to_padded(self):
    """The reference reader's padded uint8 [B, max_frames, D] (host side; tests)."""
    b, t, d = self.shape
    out = torch.zeros((b, t, d), dtype=torch.uint8)
    data, off, nf = self.data.cpu(), self.offsets.cpu(), self.num_frames.cpu()
    for i in range(b):
      n = int(nf[i])
      out[i, :n] = data[int(off[i]):int(off[i]) + n]
    return out

  @staticmethod
  def from_padded(u8, num_frames):
    """Drop the padding of a [B, max_frames, D] uint8 batch (synthetic data)."""
    b, t, d = u8.shape
    nf = num_frames.to(torch.int64).clamp(0, t)
    keep = (torch.arange(t).unsqueeze(0) < nf.unsqueeze(1))
    return PackedFrames(u8[keep].contiguous(), nf.to(torch.int32), t)


class YT8MFrameFeatureReader(BaseReader):
  """Frame-level SequenceExamples (wh/readers.py:128-259).  Batches keep the features uint8-quantised;
  frames beyond max_frames are dropped and shorter videos are zero padded (resize_axis, :21-56)."""

  def __init__(self, num_classes=4716, feature_sizes=(1024,), feature_names=("inc3",), max_frames=300):
    assert len(feature_names) == len(feature_sizes), \
        "length of feature_names (={}) != length of feature_sizes (={})".format(len(feature_names), len(feature_sizes))
    self.num_classes, self.feature_sizes, self.feature_names = num_classes, list(feature_sizes), list(feature_names)
    self.max_frames = max_frames

  def get_video_matrix(self, frames, feature_size):
    """decode_raw uint8 + pad / truncate to max_frames (wh/readers.py:159-186; de-quantisation happens on the GPU)."""
    mat = np.frombuffer(b"".join(frames), dtype=np.uint8).reshape(-1, feature_size)
    n = min(mat.shape[0], self.max_frames)
    out = np.zeros((self.max_frames, feature_size), dtype=np.uint8)
    out[:n] = mat[:n]
    return out, n

  def prepare_reader(self, data_pattern, batch_size=1024, num_epochs=1, verify_crc=False, packed=False, shuffle=False, seed=0,
                     shard=None, drop_remainder=False, shuffle_buffer=None):
    """Generator of (video_ids, features uint8 [B, max_frames, D], labels bool [B, C], num_frames int32 [B]).
    packed=True: the features come as a PackedFrames (real frames only; same .shape) instead of the padded tensor.
    shuffle / seed / shard / drop_remainder: see record_batches and YT8MAggregatedFeatureReader.prepare_reader."""
    d = sum(self.feature_sizes)
    for recs in record_batches(data_pattern, batch_size, num_epochs, verify_crc, shuffle, seed, shuffle_buffer, drop_remainder):
      lo, hi = shard_range(len(recs), shard)
      ids, mats, labels, nfs = [], [], [], []
      for rec in recs[lo:hi]:
        ctx, lists = parse_sequence_example(rec)
        parts, nf = [], -1
        for name, size in zip(self.feature_names, self.feature_sizes):
          m, n = self.get_video_matrix([f[1][0] for f in lists[name]], size)
          if nf != -1 and n != nf:
            raise ValueError("feature %s has %d frames, expected %d" % (name, n, nf))
          nf = n
          parts.append(m)
        ids.append(ctx["video_id"][1][0])
        mats.append(np.concatenate(parts, axis=1))
        labels.append(_label_vector(ctx.get("labels"), self.num_classes))
        nfs.append(nf)
      nft = torch.tensor(nfs, dtype=torch.int32)
      if packed:
        data = np.concatenate([m[:n] for m, n in zip(mats, nfs)], axis=0) if mats else np.zeros((0, d), dtype=np.uint8)
        feats = PackedFrames(torch.from_numpy(data), nft, self.max_frames)
      else:
        feats = torch.from_numpy(np.stack(mats)) if mats else torch.zeros((0, self.max_frames, d), dtype=torch.uint8)
      lab = torch.from_numpy(np.stack(labels)) if labels else torch.zeros((0, self.num_classes), dtype=torch.bool)
      out = (ids, feats, lab, nft)
      yield out if shard is None else out + (len(recs),)
