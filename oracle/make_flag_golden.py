"""Generate tests/golden/flags_golden.json from the REFERENCE's own sources: every flags.DEFINE_* of the command lines
and model-flag modules on the hot path, as {file: {flag: [kind, default]}} (SURVEY.md §8b "CLI flags that must survive").

Run in the build container only (needs /root/reference):

    python oracle/make_flag_golden.py

The reference is Python 2 / TensorFlow 1 and is NOT imported: the DEFINE calls are read with the `ast` module after a
lib2to3-free normalisation of the two py2-only constructs that appear in these files (`print x` statements), so the
defaults are the literal values of the reference source.
"""
import ast
import json
import os
import re

REF = "/root/reference/youtube-8m-wangheda"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "flags_golden.json")
FILES = ["train.py", "eval.py", "inference.py", "inference-pre-ensemble.py", "frame_level_models.py", "video_level_models.py",
         "losses.py", "feature_transform.py"]


def py3ify(src):
  """`print a, b` -> `print(a, b)` (the only py2 syntax in these files); everything else parses under Python 3."""
  out = []
  for line in src.splitlines():
    m = re.match(r"^(\s*)print (.*)$", line)
    if m and not m.group(2).lstrip().startswith("("):
      line = "%sprint(%s)" % (m.group(1), m.group(2))
    out.append(line)
  return "\n".join(out)


def defines(path):
  tree = ast.parse(py3ify(open(path, errors="ignore").read()))
  found = {}
  for node in ast.walk(tree):
    if isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute) and node.func.attr.startswith("DEFINE_"):
      kind = node.func.attr[len("DEFINE_"):]
      try:
        name = ast.literal_eval(node.args[0])
        default = ast.literal_eval(node.args[1])
      except Exception:
        continue
      found[name] = [kind, default]
  return found


def main():
  golden = {f: defines(os.path.join(REF, f)) for f in FILES if os.path.exists(os.path.join(REF, f))}
  json.dump(golden, open(OUT, "w"), indent=1, sort_keys=True)
  print({f: len(v) for f, v in golden.items()})


if __name__ == "__main__":
  main()
