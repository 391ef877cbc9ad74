#!/usr/bin/env python
"""Turn gpurun_out/{launches.csv, *.ncu-rep, bench json} into the tracked summaries under profiles/.

    python tools/summarize_profiles.py <tag> <launches.csv> <report.ncu-rep> <bench.log> [note]

Writes profiles/<tag>_launches.csv (our kernels only: name, grid, block, ns), profiles/<tag>_ncu_raw.csv
(selected ncu --set full metrics per captured kernel) and profiles/<tag>_summary.md."""
import collections
import csv
import io
import json
import os
import subprocess
import sys

tag, launches, rep, benchlog = sys.argv[1:5]
note = sys.argv[5] if len(sys.argv) > 5 else ""
command = sys.argv[6] if len(sys.argv) > 6 else "python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = os.path.join(ROOT, "profiles")
os.makedirs(out, exist_ok=True)
OURS = ("stft_kernel", "gram_tc_kernel", "gram_tma_kernel", "gram_eig_kernel", "gram_eig1_kernel", "gram_reduce", "gram_simt", "eig_power", "eig_jacobi", "eig_sort", "svd_", "lognorm",
        "minmax", "quantfilt", "patch_kernel", "unpatch", "csd_", "rescale", "moments", "norm_apply", "img_", "blur_", "morph_",
        "meansub_", "u8_lut")

rows = list(csv.reader(open(launches)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
ix = {h: i for i, h in enumerate(rows[hi])}
mine, other_ns = [], 0.0
for r in rows[hi + 1:]:
    if len(r) <= ix["Metric Value"]:
        continue
    name = r[ix["Kernel Name"]]
    ns = float(r[ix["Metric Value"]].replace(",", ""))
    if any(k in name for k in OURS):
        mine.append((name.split("(")[0][-60:], r[ix["Grid Size"]], r[ix["Block Size"]], ns))
    else:
        other_ns += ns
with open(os.path.join(out, f"{tag}_launches.csv"), "w") as f:
    w = csv.writer(f)
    w.writerow(["kernel", "grid", "block", "gpu__time_duration_ns"])
    w.writerows(mine)
agg = collections.OrderedDict()
for n, g, b, ns in mine:
    a = agg.setdefault(n, [0, 0.0])
    a[0] += 1
    a[1] += ns
tot = sum(v[1] for v in agg.values()) or 1.0

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
           "launch__grid_size", "launch__block_size", "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
           "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rr[0], rr[1], rr[2:]
hx = {h: i for i, h in enumerate(hdr)}
with open(os.path.join(out, f"{tag}_ncu_raw.csv"), "w") as f:
    w = csv.writer(f)
    w.writerow(["Kernel Name"] + [f"{m} [{units[hx[m]]}]" for m in METRICS if m in hx])
    for r in data:
        w.writerow([r[hx["Kernel Name"]][:70]] + [r[hx[m]] for m in METRICS if m in hx])

bench = None
if benchlog != "-":
    bench = json.loads([l for l in open(benchlog).read().strip().split("\n") if l.startswith("{")][-1])
with open(os.path.join(out, f"{tag}_summary.md"), "w") as f:
    f.write(f"# {tag}: ncu launch list + full capture summary\n\n{note}\n\n")
    f.write(f"Command: `{command}` on one B200 (gpurun); launch list from\n"
            "`ncu --metrics gpu__time_duration.sum --clock-control none`, per-kernel metrics from `ncu --set full "
            "--clock-control none --import-source on`.\nncu launch times are cold-cache and serialised: compare SHARES.\n\n")
    f.write("## Launch list (our kernels; setup kernels of torch's synthetic-input generation excluded: "
            f"{other_ns / 1e6:.2f} ms)\n\n| kernel | launches | total us | share |\n|---|---|---|---|\n")
    for n, (c, ns) in agg.items():
        f.write(f"| `{n}` | {c} | {ns / 1e3:.1f} | {100 * ns / tot:.1f} % |\n")
    if bench is not None:
      f.write("\n## bench.py (same build, not under ncu)\n\n")
      f.write(f"* {bench['ms_per_step']:.4f} ms per 40-channel shot = {bench['value'] / 1e9:.1f} G samples/s; e2e "
            f"{(bench['e2e']['value'] or 0) / 1e9:.2f} G samples/s\n")
      ks = bench.get("kernels", {})
      ktot = sum(v["ms_per_launch"] for v in ks.values()) or 1.0
      f.write("\n| kernel group (CUDA events in bench.py) | ms / launch | share |\n|---|---|---|\n")
      for k, v in ks.items():
          f.write(f"| {k} | {v['ms_per_launch']:.4f} | {100 * v['ms_per_launch'] / ktot:.1f} % |\n")
      f.write(f"\nroofline: `{json.dumps(bench.get('roofline'))}`\n\n")
    f.write("## ncu --set full (per captured launch)\n\n| kernel | us | DRAM rd MB | DRAM wr MB | DRAM % | SM % | issue % | warps % | tensor % | regs |\n|---|---|---|---|---|---|---|---|---|---|\n")
    for r in data:
        g = lambda m: r[hx[m]] if m in hx else ""
        f.write(f"| `{r[hx['Kernel Name']][:48]}` | {g('gpu__time_duration.sum')} | {g('dram__bytes_read.sum')} | {g('dram__bytes_write.sum')} | "
                f"{g('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')[:5]} | {g('sm__throughput.avg.pct_of_peak_sustained_elapsed')[:5]} | "
                f"{g('smsp__issue_active.avg.pct_of_peak_sustained_active')[:5]} | {g('sm__warps_active.avg.pct_of_peak_sustained_active')[:5]} | "
                f"{g('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active')[:5]} | {g('launch__registers_per_thread')} |\n")
print("wrote", out)
