"""Summarise an ncu report (--set full --import-source on): headline metrics, stall reasons, and the
executed-instruction / stall-sample split between the kernel's barrier-separated phases.

    python tools/ncu_phases.py gpurun_out/prof_x.ncu-rep [kernel-substring]"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "launch__grid_size", "launch__block_size",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
        "sm__inst_executed_pipe_tensor_op_hmma.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed"]


def run(args):
    return subprocess.run(["ncu", "-i"] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    sub = sys.argv[2] if len(sys.argv) > 2 else ""
    rows = list(csv.reader(io.StringIO(run([rep, "--page", "raw", "--csv"]))))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        if sub not in name:
            continue
        print("##", name[:100])
        for k in KEYS:
            if k in hdr:
                print("  %-70s %s %s" % (k, r[hdr.index(k)], units[hdr.index(k)]))
    rows = list(csv.reader(io.StringIO(run([rep, "--page", "source", "--csv", "--print-source", "sass"]))))
    i = 0
    while i < len(rows):
        if rows[i] and rows[i][0] == "Kernel Name":
            kname = rows[i][1]
            hdr = rows[i + 1]
            i += 2
            data = []
            while i < len(rows) and rows[i] and rows[i][0] != "Kernel Name":
                data.append(rows[i])
                i += 1
            if sub not in kname:
                continue
            isrc, ie, iss = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
            stalls = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
            tot = sum(int(r[ie] or 0) for r in data)
            ts = sum(int(r[iss] or 0) for r in data)
            print("## phases of", kname[:80], "| warp instructions", tot, "| samples", ts)
            print("  stalls:", ", ".join("%s=%d" % (k[6:], sum(int(r[hdr.index(k)] or 0) for r in data)) for k in stalls
                                       if sum(int(r[hdr.index(k)] or 0) for r in data) > ts * 0.01))
            acc = samp = 0
            ops = {}
            n = 0
            for r in data:
                acc += int(r[ie] or 0)
                samp += int(r[iss] or 0)
                op = r[isrc].split()[0] if r[isrc].split() else ""
                if op.startswith("@"):
                    op = r[isrc].split()[1]
                op = op.split(".")[0]
                ops[op] = ops.get(op, 0) + int(r[ie] or 0)
                if "BAR.SYNC" in r[isrc] or "EXIT" in r[isrc]:
                    if acc:
                        top = sorted(ops.items(), key=lambda kv: -kv[1])[:6]
                        print("  phase %d: %5.1f%% of instructions, %5.1f%% of samples | %s" % (
                            n, 100.0 * acc / max(1, tot), 100.0 * samp / max(1, ts),
                            " ".join("%s:%.0f%%" % (k, 100.0 * v / acc) for k, v in top)))
                    n += 1
                    acc = samp = 0
                    ops = {}
        else:
            i += 1


if __name__ == "__main__":
    main()
