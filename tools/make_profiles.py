"""Turn the ncu outputs of tools/profile_round.sh into the committed summaries under profiles/:

    python tools/make_profiles.py r01          (reads gpurun_out/r01_*.csv / *.json)

  profiles/<tag>_launches_step.csv / .md   launch list of one bench step + per-kernel totals
  profiles/<tag>_traffic.json              DRAM bytes per launch and GB/s per kernel family (bench.py reads the
                                           dense-layer figure as roofline.traffic)
  profiles/<tag>_top_kernels_ncu.md        one row per distinct (kernel, shape) of the --set full capture
  profiles/<tag>_bench.json, <tag>_stft_bench.json   the bench lines of the same run"""
import collections
import csv
import json
import os
import re
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
HBM_PEAK = 6548.8


def short(name):
    name = re.sub(r"\(.*", "", name).replace("void ", "")
    return name.replace("wmk::<unnamed>::", "").replace("wmk::", "").replace("unnamed>::", "")


def family(name):
    for key, fam in (("leff_block", "leff_fused"), ("attn_block", "attention"), ("gemm_tcgen05", "gemm"), ("attention", "attention"), ("dwconv", "dwconv"), ("layernorm", "layernorm"),
                     ("im2col", "layout"), ("s2d_pad", "layout"), ("copy_cols", "layout"), ("stft", "frontend"), ("iir", "attack"), ("awgn", "attack")):
        if key in name:
            return fam
    return "other"


def launches(tag):
    src = os.path.join(ROOT, "gpurun_out", tag + "_launches.csv")
    with open(src) as f:
        lines = [l for l in f if not l.startswith("==")]
    recs = collections.OrderedDict()
    for row in csv.DictReader(lines):
        d = recs.setdefault(row["ID"], {"name": short(row["Kernel Name"]), "by": 0.0, "us": 0.0})
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        if "time" in row["Metric Name"]:
            d["us"] = v / 1e3 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1e3)
        else:
            d["by"] += v * UNIT[u]
    shutil.copy(src, os.path.join(ROOT, "profiles", tag + "_launches_step.csv"))
    agg, fams = collections.OrderedDict(), collections.OrderedDict()
    for d in recs.values():
        a = agg.setdefault(d["name"], [0, 0.0, 0.0])
        a[0] += 1; a[1] += d["us"]; a[2] += d["by"]
        f = fams.setdefault(family(d["name"]), {"launches": 0, "dram_bytes": 0.0, "us": 0.0})
        f["launches"] += 1; f["dram_bytes"] += d["by"]; f["us"] += d["us"]
    tot = sum(a[1] for a in agg.values())
    with open(os.path.join(ROOT, "profiles", tag + "_launches_step.md"), "w") as f:
        f.write("| kernel | launches | total us | share | avg us | DRAM GB/s |\n|---|---:|---:|---:|---:|---:|\n")
        for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("| `%s` | %d | %.1f | %.1f%% | %.1f | %.0f |\n" % (n, a[0], a[1], 100 * a[1] / tot, a[1] / a[0], a[2] / a[1] / 1e3 if a[1] else 0))
        f.write("\ntotal %.1f us over %d launches of ONE timed bench step (ncu NVTX range wmk_timed_step; cold-cache, serialised "
                "by ncu: compare shares, not absolutes)\n" % (tot, len(recs)))
    for f in fams.values():
        f["dram_bytes_per_launch"] = f["dram_bytes"] / f["launches"]
        f["dram_gbs"] = f["dram_bytes"] / f["us"] / 1e3 if f["us"] else 0.0
    json.dump({"source": "ncu --nvtx-include wmk_timed_step/ --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum "
                         "--clock-control none over the %d launches of one bench step (64 x 3 s); profiles/%s_launches_step.csv"
                         % (len(recs), tag), "families": fams},
              open(os.path.join(ROOT, "profiles", tag + "_traffic.json"), "w"), indent=1)
    print("launches:", len(recs), "total us %.0f" % tot)


def top(tag):
    src = os.path.join(ROOT, "gpurun_out", tag + "_top_raw.csv")
    rows = list(csv.reader(open(src)))
    hdr, units = rows[0], rows[1]
    col = {k: hdr.index(k) for k in hdr}

    def val(r, k, scale_unit=False):
        if k not in col or r[col[k]] == "":
            return float("nan")
        v = float(r[col[k]].replace(",", ""))
        if scale_unit:
            u = units[col[k]]
            v *= UNIT.get(u, 1)
            if u in ("ns", "nsecond"):
                v /= 1e3
            elif u in ("ms", "msecond"):
                v *= 1e3
        return v
    seen, out = set(), []
    for r in rows[2:]:
        name = short(r[col["Kernel Name"]])
        us = val(r, "gpu__time_duration.sum", True)
        rd, wr = val(r, "dram__bytes_read.sum", True), val(r, "dram__bytes_write.sum", True)
        key = (name, round(rd / 1e7), round(wr / 1e7))
        if key in seen or us < 20:
            continue
        seen.add(key)
        gbs = (rd + wr) / us / 1e3
        out.append("| `%s` | %s x %s | %d | %.1f | %.0f | %.0f | %.0f (%.0f%%) | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f |" % (
            name, r[col["launch__grid_size"]], r[col["launch__block_size"]], val(r, "launch__registers_per_thread"), us,
            rd / 1e6, wr / 1e6, gbs, 100 * gbs / HBM_PEAK, val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
            val(r, "sm__warps_active.avg.pct_of_peak_sustained_active"), val(r, "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
            val(r, "lts__throughput.avg.pct_of_peak_sustained_elapsed"), val(r, "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
            val(r, "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active")))
    with open(os.path.join(ROOT, "profiles", tag + "_top_kernels_ncu.md"), "w") as f:
        f.write("# ncu --set full --clock-control none, first launches of one timed bench step (64 x 3 s), B200\n\n"
                "One row per distinct (kernel, traffic) pair; launches shorter than 20 us omitted.  DRAM GB/s = (dram__bytes_read.sum + "
                "dram__bytes_write.sum) / gpu__time_duration; %% of the measured %.1f GB/s copy peak in brackets.  Times are cold-cache "
                "and serialised by the profiler.\n\n" % HBM_PEAK)
        f.write("| kernel | grid x block | regs | us | DRAM read MB | DRAM write MB | DRAM GB/s | issue active % | warps active % | "
                "L1/TEX % | L2 % | FMA pipe % | tensor (hmma) pipe % |\n|---|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|\n")
        f.write("\n".join(out) + "\n")
    print("top kernels:", len(out))


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    launches(tag)
    top(tag)
    for n in ("_bench.json", "_stft_bench.json"):
        src = os.path.join(ROOT, "gpurun_out", tag + n)
        if os.path.exists(src):
            shutil.copy(src, os.path.join(ROOT, "profiles", tag + n))
