"""Summarise an .ncu-rep offline: headline metrics + warp-stall samples per SASS instruction (top N) and per mbarrier
wait.  Usage: python tools/ncu_stalls.py gpurun_out/x.ncu-rep [topN]"""
import csv
import io
import re
import subprocess
import sys


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return dict(zip(rows[0], rows[2] if len(rows) > 2 else rows[1]))


KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct", "launch__registers_per_thread",
        "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]


def main():
    rep = sys.argv[1]
    top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    r = raw(rep)
    for k in KEYS:
        if k in r:
            print(f"{k:75s} {r[k]}")
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    data = rows[2:]
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(x[idx["# Samples"]]) for x in data)
    print("total samples", tot)
    agg = {s: sum(int(x[idx[s]]) for x in data) for s in stalls}
    for s, v in sorted(agg.items(), key=lambda x: -x[1])[:10]:
        print(f"  {s:28s}{v:8d} {100 * v / tot:5.1f}%")
    print("-- top instructions by samples")
    order = sorted(range(len(data)), key=lambda i: -int(data[i][idx["# Samples"]]))[:top_n]
    for i in order:
        x = data[i]
        st = sorted(((s, int(x[idx[s]])) for s in stalls if int(x[idx[s]]) > 0), key=lambda t: -t[1])[:2]
        prev = data[i - 1][idx["Source"]].strip()[:50] if i else ""
        print(f"  {x[idx['# Samples']]:>6} exec {x[idx['Instructions Executed']]:>9}  {x[idx['Source']].strip()[:58]:58s} {st}  <- {prev}")
    print("-- mbarrier waits")
    for i, x in enumerate(data):
        if "TRYWAIT" in x[idx["Source"]]:
            m = re.search(r"\+0x([0-9a-f]+)\]", x[idx["Source"]])
            nxt = data[i + 1]
            print(f"  off {m.group(1) if m else '?':>6} exec {x[idx['Instructions Executed']]:>9} samples "
                  f"{int(x[idx['# Samples']]) + int(nxt[idx['# Samples']]):>6}")


if __name__ == "__main__":
    main()
