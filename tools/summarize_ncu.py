"""Summarise ncu outputs of one gpurun call (gpurun_out/<tag>/) into profiles/<name>.md.

  python tools/summarize_ncu.py gpurun_out/r01b profiles/r01_ncu_tf32_engine.md

launches.csv  : `ncu --metrics gpu__time_duration.sum` launch list -> share of device time per kernel
prof_*.raw.csv: `ncu --set full` raw page of single launches -> the counters the roofline cites
"""
import csv
import glob
import os
import re
import sys
from collections import OrderedDict

KEYS = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max"]


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    name = re.sub(r"<unnamed>::", "", name)
    return name[:110]


def launches(path, out):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 14 and r[0].isdigit()]
    agg = OrderedDict()
    for r in rows:
        if r[12] != "gpu__time_duration.sum":
            continue
        v = float(r[14].replace(",", ""))
        v = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[13], 1e-6) * v
        d = agg.setdefault(short(r[4]), [0, 0.0])
        d[0] += 1
        d[1] += v
    tot = sum(d[1] for d in agg.values())
    out.append("## Launch list (%s): %d launches, %.1f ms of device time (cold-cache, serialised: shares only)\n" % (os.path.basename(path), len(rows), tot))
    out.append("| kernel | launches | ms | share |\n|---|---|---|---|")
    for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
        out.append("| `%s` | %d | %.3f | %.1f%% |" % (k, n, ms, 100 * ms / tot))
    out.append("")


def full(path, out):
    rows = list(csv.reader(open(path, errors="replace")))
    if len(rows) < 3:
        return
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = dict(zip(hdr, zip(units, vals)))
        out.append("## ncu --set full: %s -- `%s`\n" % (os.path.basename(path), short(d.get("Kernel Name", ("", "?"))[1])))
        out.append("| metric | value | unit |\n|---|---|---|")
        for k in KEYS:
            if k in d:
                out.append("| %s | %s | %s |" % (k, d[k][1], d[k][0]))
        out.append("")


def main():
    src, dst = sys.argv[1], sys.argv[2]
    out = ["# ncu summary of %s\n" % src]
    for extra in sys.argv[3:]:
        out.append(extra + "\n")
    for p in sorted(glob.glob(os.path.join(src, "launches*.csv"))):
        launches(p, out)
    for p in sorted(glob.glob(os.path.join(src, "prof_*.raw.csv"))):
        full(p, out)
    open(dst, "w").write("\n".join(out) + "\n")
    print("wrote", dst)


if __name__ == "__main__":
    main()
