"""Feature transformers (wh/feature_transform.py + wh/all_feature_transform/default_transformer.py).

DefaultTransformer L2-normalises every frame row on the GPU (fused with the uint8 de-quantisation
of the reader when the raw features are still quantised)."""
import torch

import readers
import yt8m_flags as flags
import yt8m_native as nat

flags.DEFINE_string("feature_transformer", "DefaultTransformer",
                    "how to preprocess feature, defaults to identical, which means no transform")
# flags of the reference's EngineerTransformer / ResolutionTransformer (outside SURVEY.md §8): accepted so that command lines parse
flags.DEFINE_string("engineer_types", "identical,avg,std,diff", "how to preprocess feature (EngineerTransformer; not built)")
flags.DEFINE_integer("time_resolution", 8, "how many frames are merged into one (ResolutionTransformer; not built)")


class DefaultTransformer(object):
  """model_input = tf.nn.l2_normalize(model_input_raw, last_dim) (default_transformer.py:5-8).

  Accepts fp32 / bf16 features, or the raw uint8 features of the TFRecords (wh/readers.py:178-186), in
  which case Dequantize (wh/utils.py:23-38) and the zero padding past num_frames are fused in -- either as the
  reader's padded [B, max_frames, D] tensor or as a readers.PackedFrames (real frames only).
  Returns bf16 (the tensor-core operand dtype) and num_frames unchanged."""

  def transform(self, model_input_raw, num_frames, **unused_params):
    x = model_input_raw
    if isinstance(x, readers.PackedFrames):
      # ragged batch: only the real frames were uploaded; de-quantise + L2-normalise + zero padding in one pass
      x = x.cuda(non_blocking=True)
      return nat.frames_unpack_u8(x.data, x.offsets, x.num_frames, x.max_frames, normalize=True), num_frames
    if not x.is_cuda:
      x = x.cuda(non_blocking=True)
    if x.dtype not in (torch.float32, torch.bfloat16, torch.uint8):
      x = x.float()
    nf = num_frames.to(x.device, torch.int32) if (num_frames is not None and x.dim() == 3) else None
    out = nat.l2norm_rows(x.contiguous(), normalize=True, num_frames=nf if x.dtype == torch.uint8 else None)
    return out, num_frames


class IdenticalTransformer(object):
  """wh/all_feature_transform/identical_transformer.py: no transform."""

  def transform(self, model_input_raw, num_frames, **unused_params):
    return model_input_raw, num_frames
