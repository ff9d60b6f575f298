"""Summarise gpurun_out/{launches.csv, prof.ncu-rep, bench.json} into profiles/<tag>_*.{md,csv,json} (committed evidence)."""
import csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
out = os.path.join(ROOT, "profiles")
os.makedirs(out, exist_ok=True)
g = os.path.join(ROOT, "gpurun_out")

# ---- launch list: per-kernel time share of one bench step
rows = [r for r in csv.reader(open(os.path.join(g, "launches.csv"))) if r and r[0].isdigit()]
agg = {}
for r in rows:
  name, val = r[4], float(r[-1])
  unit = r[-2]
  us = val / 1000.0 if unit in ("ns", "nsecond") else val
  short = name.split("(")[0].replace("void ", "")[:90]
  a = agg.setdefault(short, [0, 0.0])
  a[0] += 1
  a[1] += us
total = sum(v[1] for v in agg.values())
with open(os.path.join(out, tag + "_launches.md"), "w") as f:
  f.write("# ncu launch list (%s): `ncu --metrics gpu__time_duration.sum --clock-control none` over `bench.py --steps 2 --warmup 3`\n\n" % tag)
  f.write("Cold-cache, serialised per-launch times: compare SHARES, not absolutes.\n\n| kernel | launches | total us | share |\n|---|---:|---:|---:|\n")
  for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    f.write("| `%s` | %d | %.1f | %.1f%% |\n" % (k, n, us, 100 * us / total))

# ---- full-set capture: key metrics per kernel
raw = subprocess.run(["ncu", "-i", os.path.join(g, "prof.ncu-rep"), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw)))
hdr, units = rr[0], rr[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "smsp__cycles_active.avg"]
idx = {h: i for i, h in enumerate(hdr)}
summary = {}
with open(os.path.join(out, tag + "_ncu_full.md"), "w") as f:
  f.write("# ncu --set full summary (%s)\n\n`ncu --set full --clock-control none --import-source on -k regex:netvlad|gemm_tcgen05|l2norm_rows` over one bench step.\n\n" % tag)
  for r in rr[2:]:
    name = r[idx["Kernel Name"]].replace("void ", "")
    short = name.split("(")[0][:100]
    f.write("## `%s`\n\n| metric | value | unit |\n|---|---:|---|\n" % short)
    d = {}
    for w in want:
      if w in idx:
        f.write("| %s | %s | %s |\n" % (w, r[idx[w]], units[idx[w]]))
        d[w] = (r[idx[w]], units[idx[w]])
    f.write("\n")
    summary[short] = d

def to_bytes(v, u):
  x = float(v.replace(",", ""))
  return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)

traffic = {}
for k, d in summary.items():
  if "dram__bytes_read.sum" in d:
    traffic[k] = to_bytes(*d["dram__bytes_read.sum"]) + to_bytes(*d["dram__bytes_write.sum"])
nv = [v for k, v in traffic.items() if "netvlad" in k]
json.dump({"netvlad_fused_kernel_dram_bytes_per_launch": nv[0] if nv else None, "per_kernel_dram_bytes": traffic},
          open(os.path.join(out, "traffic.json"), "w"), indent=1)
if os.path.exists(os.path.join(g, "bench.json")):
  open(os.path.join(out, tag + "_bench.json"), "w").write(open(os.path.join(g, "bench.json")).read())
print(open(os.path.join(out, tag + "_launches.md")).read())
print(json.dumps(traffic, indent=1))
