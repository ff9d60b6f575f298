"""CPU-side checks: the C-ABI library loads and exports every symbol include/yt8m_b200.h declares, the
flag shim parses like the reference's flags, the plugin registries expose the reference class names, and
the product path refuses to run without a GPU (no fallback)."""
import os
import re
import subprocess

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
  import build_native
  so = build_native.build()
  header = open(os.path.join(ROOT, "include", "yt8m_b200.h")).read()
  declared = set(re.findall(r"\b(yt8m_[a-z0-9_]+)\s*\(", header))
  out = subprocess.run(["nm", "-D", "--defined-only", so], capture_output=True, text=True).stdout
  exported = set(re.findall(r" T (yt8m_[a-z0-9_]+)", out))
  assert declared and not (declared - exported), sorted(declared - exported)
  import yt8m_native
  assert set(yt8m_native.EXPORTS) == declared, sorted(set(yt8m_native.EXPORTS) ^ declared)
  assert yt8m_native.version() >= 100
  assert yt8m_native.moe_packed_rows(4716, 2) == 128 * ((4716 + 24) // 25)
  assert yt8m_native.moe_packed_rows(4716, 64) == -1


def test_ctypes_signatures_match_the_header_prototypes():
  """Every binding in yt8m_native._SIGS has the argument COUNT and, per argument, the KIND (pointer / 32-bit int / 64-bit
  int / size_t / float) of the prototype in include/yt8m_b200.h -- a drift between the header and the ctypes table (an
  argument added on one side only) corrupts the call frame silently; this catches it without a GPU."""
  import ctypes
  import yt8m_native
  header = open(os.path.join(ROOT, "include", "yt8m_b200.h")).read()
  header = re.sub(r"/\*.*?\*/", " ", header, flags=re.S)
  header = re.sub(r"//[^\n]*", " ", header)
  protos = dict((m.group(2), (m.group(1), m.group(3))) for m in
                re.finditer(r"\b([A-Za-z_][A-Za-z0-9_ ]*?[\s\*]+)(yt8m_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", header))

  def kind_of_c(decl):
    decl = decl.strip()
    if "*" in decl or decl.startswith("yt8m_stream_t"):
      return "ptr"
    base = re.sub(r"\b(const|unsigned)\b", "", decl).split()
    t = " ".join(base[:-1]) if len(base) > 1 else base[0]
    return {"int": "i32", "long long": "i64", "size_t": "size", "float": "f32"}.get(t, t)

  def kind_of_ctypes(t):
    return {ctypes.c_void_p: "ptr", ctypes.c_char_p: "ptr", ctypes.c_int: "i32", ctypes.c_longlong: "i64",
            ctypes.c_size_t: "size", ctypes.c_float: "f32"}.get(t, str(t))

  assert set(protos) == set(yt8m_native._SIGS)
  for name, (res, args) in yt8m_native._SIGS.items():
    ret_decl, arg_decl = protos[name]
    c_args = [a for a in (x.strip() for x in arg_decl.split(",")) if a and a != "void"]
    assert len(c_args) == len(args), (name, len(c_args), len(args))
    assert [kind_of_c(a) for a in c_args] == [kind_of_ctypes(t) for t in args], (name, c_args)
    want_ret = "ptr" if "*" in ret_decl else kind_of_c(ret_decl + " x")
    assert kind_of_ctypes(res) == want_ret, (name, ret_decl)


def test_sass_contains_tcgen05_and_tma():
  import build_native
  sass = subprocess.run(["cuobjdump", "-sass", build_native.build()], capture_output=True, text=True).stdout
  for needle in ("UTCHMMA", "UTMALDG", "LDTM"):
    assert needle in sass, needle
  assert "HMMA.16816" not in sass              # no legacy mma.sync path


def test_no_cpu_fallback():
  import yt8m_native
  if torch.cuda.is_available():
    pytest.skip("CPU-only check")
  with pytest.raises(yt8m_native.Yt8mError):
    yt8m_native.l2norm_rows(torch.zeros(2, 8))
  import video_level_models
  with pytest.raises(Exception):
    video_level_models.LogisticModel().create_model(torch.zeros(2, 8), vocab_size=4)


def test_flags_shim():
  import yt8m_flags as flags
  fv = flags.FlagValues()
  fv._define("model", "LogisticModel", "", "string")
  fv._define("batch_size", 1024, "", "integer")
  fv._define("base_learning_rate", 0.01, "", "float")
  fv._define("frame_features", False, "", "bool")
  fv._define("start_new_model", False, "", "bool")
  rest = fv.parse(["--model=LstmModel", "--batch_size", "128", "--frame_features", "--base_learning_rate=0.001",
                   "--nostart_new_model", "positional"])
  assert rest == ["positional"]
  assert (fv.model, fv.batch_size, fv.frame_features, fv.base_learning_rate, fv.start_new_model) == ("LstmModel", 128, True, 0.001, False)
  with pytest.raises(ValueError):
    fv.parse(["--no_such_flag=1"])
  assert fv.parse(["--no_such_flag=1"], known_only=True) == ["--no_such_flag=1"]
  fv.parse(["--frame_features=false"])
  assert fv.frame_features is False
  with fv.override(batch_size=7):
    assert fv.batch_size == 7
  assert fv.batch_size == 128
  with pytest.raises(ValueError):
    fv._define("model", 3, "", "integer")


def test_plugin_registries_and_flag_defaults():
  import frame_level_models, video_level_models, models
  from yt8m_flags import FLAGS
  FLAGS.reset()
  # defaults copied from wh/frame_level_models.py:20-83 and wh/video_level_models.py:19-46
  assert FLAGS.moe_num_mixtures == 2 and FLAGS.lstm_cells == "1024" and FLAGS.lstm_layers == 2
  assert FLAGS.lstm_attentions == 8 and FLAGS.video_level_classifier_model == "MoeModel"
  assert FLAGS.dbof_cluster_size == 8192 and FLAGS.iterations == 30 and FLAGS.deep_chain_layers == 3
  for mod, names in ((video_level_models, ["LogisticModel", "MoeModel", "ChainMoeModel", "DeepCombineChainModel", "MoeExtendModel"]),
                     (frame_level_models, ["LstmModel", "LstmMemoryModel", "LstmAttentionMaxPoolingModel", "LstmMultiAttentionModel",
                                           "DbofModel", "AttentionModel", "NetVLADModel", "GatedNetVLADModel", "FrameLevelLogisticModel"])):
    for n in names:
      assert issubclass(getattr(mod, n), models.BaseModel), n
  with pytest.raises(NotImplementedError):
    models.BaseModel().create_model(None)


def test_sample_frames_host_logic():
  import frame_level_models as flm
  g = torch.Generator().manual_seed(0)
  nf = torch.tensor([300, 5, 1])
  idx = flm.sample_frames(nf, 30, True, generator=g)
  assert idx.shape == (3, 30) and bool((idx < nf.unsqueeze(1)).all()) and bool((idx >= 0).all())
  seq = flm.sample_frames(nf, 30, False, generator=g)
  assert bool((seq[0, 1:] - seq[0, :-1] == 1).all())            # a contiguous run when the video is long enough
  assert bool((seq[1] <= 4).all()) and int(seq[2].max()) == 0     # clamped to the last valid frame


def test_trainer_layout_helpers_roundtrip():
  """Host-side layout maps of the trainers (no GPU): TF BasicLSTMCell kernel [in+H, 4H] (columns g*H + u) <-> packed
  [4H, in+H] (rows 4u + g, the layout of yt8m_lstm_pack_weights); MoE reference columns <-> packed class-major rows."""
  import torch
  import yt8m_trainer as tr
  g = torch.Generator().manual_seed(1)
  h, k = 8, 5
  w_tf = torch.randn(k, 4 * h, generator=g)
  wp = tr._lstm_tf_to_packed(w_tf, h)
  assert wp.shape == (4 * h, k)
  for u in (0, 3, 7):
    for gate in range(4):
      assert torch.equal(wp[4 * u + gate], w_tf[:, gate * h + u])
  assert torch.equal(tr._lstm_packed_to_tf(wp, h), w_tf)
  # MoE: vocab 7, 2 mixtures -> 5 rows per class, 25 classes per 128-row tile
  gates, experts = tr._moe_row_index(7, 2)
  assert gates.tolist()[:6] == [0, 1, 2, 5, 6, 7] and experts.tolist()[:4] == [3, 4, 8, 9]
  assert len(set(gates.tolist()) | set(experts.tolist())) == 7 * 5
  gates, experts = tr._moe_row_index(30, 2)              # class 25 starts the second tile
  assert gates.tolist()[25 * 3] == 128 and experts.tolist()[25 * 2] == 131
  # learning-rate schedule and Adam bias correction (wh/train.py:303-308; TF-1.0 AdamOptimizer)
  assert tr.exponential_decay(0.01, 0, 1024, 4000000, 0.95) == 0.01
  assert abs(tr.exponential_decay(0.01, 3907, 1024, 4000000, 0.95) - 0.0095) < 1e-12      # 3907 * 1024 > 4e6: one decay
  assert abs(tr.adam_lr_t(0.01, 1) - 0.01 * (1 - 0.999) ** 0.5 / (1 - 0.9)) < 1e-15


def test_cli_and_model_flags_match_the_reference():
  """SURVEY.md §8b "CLI flags that must survive": every flags.DEFINE_* of the reference's train.py / eval.py / inference.py /
  inference-pre-ensemble.py and of its model / loss / transform flag modules (read from the reference SOURCE by
  oracle/make_flag_golden.py -> tests/golden/flags_golden.json) exists here with the same kind and default.  Our files are
  read the same way (ast), so the `if __name__ == "__main__"` blocks of the command lines are covered too."""
  import ast
  import json
  import os
  root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
  golden = json.load(open(os.path.join(root, "tests", "golden", "flags_golden.json")))

  def defines(path):
    found = {}
    for node in ast.walk(ast.parse(open(path).read())):
      if isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute) and node.func.attr.startswith("DEFINE_"):
        try:
          found[ast.literal_eval(node.args[0])] = [node.func.attr[7:].replace("boolean", "bool"), ast.literal_eval(node.args[1])]
        except Exception:
          pass
    return found

  pkg = os.path.join(root, "youtube-8m_b200")
  modules = {}
  for f in ("frame_level_models.py", "video_level_models.py", "losses.py", "feature_transform.py"):
    modules.update(defines(os.path.join(pkg, f)))                        # module-level flags share one registry
  assert len(golden) == 8
  for fname, ref in golden.items():
    mine = defines(os.path.join(pkg, fname))
    scope = dict(modules, **mine) if fname in ("frame_level_models.py", "video_level_models.py", "losses.py", "feature_transform.py") else mine
    for name, (kind, default) in ref.items():
      assert name in scope, "%s: reference flag --%s is not defined" % (fname, name)
      k2, d2 = scope[name]
      assert k2 == kind.replace("boolean", "bool"), (fname, name, kind, k2)
      assert d2 == default and type(d2) == type(default), "%s: --%s default %r != reference %r" % (fname, name, d2, default)


def test_netvlad_tiled_index_is_the_documented_permutation():
  """include/yt8m_b200.h (yt8m_netvlad_fwd_tiled): element (d, k) sits at ((d/32)*(K/w) + k/w)*(32*w) + (d%32)*w + k%w with
  w = 8 (16-bit descriptor) or 4 (fp32 cw2); the helper returns the gather index tiled position -> row-major position."""
  import torch
  import yt8m_native as nat
  for d, k, w in ((320, 64, 8), (1152, 64, 4), (256, 64, 8)):
    idx = nat.netvlad_tiled_index(d, k, w, "cpu")
    assert sorted(idx.tolist()) == list(range(d * k))
    dd = torch.randint(0, d, (50,))
    kk = torch.randint(0, k, (50,))
    pos = ((dd // 32) * (k // w) + kk // w) * (32 * w) + (dd % 32) * w + kk % w
    assert torch.equal(idx[pos], dd * k + kk)
