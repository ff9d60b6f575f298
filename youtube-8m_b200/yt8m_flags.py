"""A small stand-in for ``tensorflow.flags`` (the reference's config system):
``flags.DEFINE_*`` at import time anywhere, one global ``FLAGS`` resolved lazily from ``sys.argv``.

Kept deliberately compatible with how the reference uses it (e.g. wh/train.py:38-137,
wh/frame_level_models.py:20-83): same DEFINE_string / DEFINE_integer / DEFINE_float / DEFINE_bool
signatures, ``--name=value`` / ``--name value`` / ``--[no]name`` parsing, attribute access on FLAGS.
"""
import sys


class _Flag(object):
  __slots__ = ("name", "default", "help", "kind", "value", "present")

  def __init__(self, name, default, help_, kind):
    self.name, self.default, self.help, self.kind = name, default, help_, kind
    self.value, self.present = default, False


def _to_bool(s):
  if isinstance(s, bool):
    return s
  v = str(s).strip().lower()
  if v in ("1", "true", "t", "yes", "y"):
    return True
  if v in ("0", "false", "f", "no", "n"):
    return False
  raise ValueError("not a boolean: %r" % (s,))


_CONVERT = {"string": lambda s: None if s is None else str(s), "integer": lambda s: None if s is None else int(s),
            "float": lambda s: None if s is None else float(s), "bool": _to_bool}


class FlagValues(object):
  def __init__(self):
    object.__setattr__(self, "_flags", {})
    object.__setattr__(self, "_parsed", False)
    object.__setattr__(self, "_argv", None)

  # -- definition ---------------------------------------------------------------------------
  def _define(self, name, default, help_, kind):
    if name in self._flags:
      # the reference defines some flags in several scripts; same name + kind is tolerated
      if self._flags[name].kind != kind:
        raise ValueError("flag --%s redefined with a different type" % name)
      return
    self._flags[name] = _Flag(name, default, help_, kind)
    object.__setattr__(self, "_parsed", False)

  # -- parsing --------------------------------------------------------------------------------
  def parse(self, argv=None, known_only=False):
    """Parses ``argv`` (default sys.argv[1:]); returns the unparsed positional remainder."""
    args = list(sys.argv[1:] if argv is None else argv)
    rest = []
    i = 0
    while i < len(args):
      a = args[i]
      i += 1
      if not a.startswith("-") or a in ("-", "--"):
        rest.append(a)
        continue
      body = a.lstrip("-")
      name, eq, val = body.partition("=")
      f = self._flags.get(name)
      if f is None and not eq and name.startswith("no") and name[2:] in self._flags and self._flags[name[2:]].kind == "bool":
        f = self._flags[name[2:]]
        f.value, f.present = False, True
        continue
      if f is None:
        if known_only:
          rest.append(a)
          continue
        raise ValueError("Unknown command line flag '%s'" % name)
      if f.kind == "bool" and not eq:
        f.value, f.present = True, True
        continue
      if not eq:
        if i >= len(args):
          raise ValueError("flag --%s needs a value" % name)
        val = args[i]
        i += 1
      f.value, f.present = _CONVERT[f.kind](val), True
    object.__setattr__(self, "_parsed", True)
    return rest

  # -- access ---------------------------------------------------------------------------------
  def __getattr__(self, name):
    flags = object.__getattribute__(self, "_flags")
    if name not in flags:
      raise AttributeError(name)
    if not object.__getattribute__(self, "_parsed"):
      self.parse(known_only=True)
    return flags[name].value

  def __setattr__(self, name, value):
    if name not in self._flags:
      raise AttributeError("no flag named %s" % name)
    self._flags[name].value = value

  def __contains__(self, name):
    return name in self._flags

  def reset(self):
    """Back to defaults (used by tests)."""
    for f in self._flags.values():
      f.value, f.present = f.default, False
    object.__setattr__(self, "_parsed", True)

  def override(self, **kw):
    """Context manager: temporarily set flags (tests / bench)."""
    fv = self

    class _Ctx(object):
      def __enter__(self_inner):
        if not fv._parsed:
          fv.parse(known_only=True)
        self_inner.old = {k: fv._flags[k].value for k in kw}
        for k, v in kw.items():
          fv._flags[k].value = v
        return fv

      def __exit__(self_inner, *exc):
        for k, v in self_inner.old.items():
          fv._flags[k].value = v
        return False

    return _Ctx()

  def flag_values_dict(self):
    if not self._parsed:
      self.parse(known_only=True)
    return {k: f.value for k, f in self._flags.items()}


FLAGS = FlagValues()


def DEFINE_string(name, default, help):  # noqa: N802 (reference spelling)
  FLAGS._define(name, default, help, "string")


def DEFINE_integer(name, default, help):  # noqa: N802
  FLAGS._define(name, default, help, "integer")


def DEFINE_float(name, default, help):  # noqa: N802
  FLAGS._define(name, default, help, "float")


def DEFINE_bool(name, default, help):  # noqa: N802
  FLAGS._define(name, default, help, "bool")


DEFINE_boolean = DEFINE_bool
