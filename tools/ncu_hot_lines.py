"""Hottest SASS instructions (warp-stall samples) of one kernel in an .ncu-rep, with a little context.

    python tools/ncu_hot_lines.py gpurun_out/prof.ncu-rep <launch-id> [top]
"""
import csv, io, subprocess, sys
rep, lid = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--launch-skip", lid, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr_i = next(i for i, r in enumerate(rows) if "# Samples" in r)
hdr = rows[hdr_i]
body = []
for r in rows[hdr_i + 1:]:
  if len(r) != len(hdr) or r == hdr:
    break
  body.append(r)
ci = {h: i for i, h in enumerate(hdr)}
print(rows[0][1][:120])
S = ci["# Samples"]
tot = sum(int(r[S] or 0) for r in body)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
order = sorted(range(len(body)), key=lambda i: -int(body[i][S] or 0))[:top]
print("total samples", tot)
for i in order:
  r = body[i]
  st = sorted(((int(r[ci[s]] or 0), s) for s in stalls), reverse=True)[:2]
  print("%5.1f%%  #%-5d %-70s %s" % (100.0 * int(r[S] or 0) / tot, i, r[ci["Source"]][:70], " ".join("%s=%d" % (s[6:], n) for n, s in st if n)))
