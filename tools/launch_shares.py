"""Per-kernel totals and shares from an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file X.csv …`).
Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.
    python tools/launch_shares.py profiles/r2c_launches_bench_steps2.csv [title]"""
import csv
import sys


def main():
    path = sys.argv[1]
    title = sys.argv[2] if len(sys.argv) > 2 else path
    rows = []
    with open(path, newline="") as f:
        lines = f.readlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
    for r in csv.DictReader(lines[start:]):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            rows.append((r["Kernel Name"], float(r["Metric Value"].replace(",", ""))))
    agg = {}
    for name, ns in rows:
        short = name.replace("void ", "").replace("stad::", "")
        short = short.split("(")[0] if short.startswith(("<unnamed>::", "gemm_kernel")) else short[:58]
        d = agg.setdefault(short, [0, 0.0])
        d[0] += 1
        d[1] += ns
    total = sum(v[1] for v in agg.values()) or 1.0
    print(f"# {title}")
    print("# per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes; unit ns")
    print(f"{'kernel':58s} {'launches':>8s} {'total':>12s} {'avg':>10s} {'share':>7s}")
    for name, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{name:58s} {n:8d} {ns:12.1f} {ns / n:10.1f} {100 * ns / total:6.1f}%")


if __name__ == "__main__":
    main()
