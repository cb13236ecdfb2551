"""Summarise an `ncu --set full` capture of the GEMM launches of one block into profiles/ncu_gemm_dram.json
(DRAM bytes per launch next to the algorithmic bytes), the source of bench.py's roofline.traffic.
    ncu -i gpurun_out/prof_gemm.ncu-rep --page raw --csv | python tools/ncu_dram_summary.py [--batch 64]"""
import argparse
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EPI = {"24, 1": "patch_embed", "8, 1": "patch_embed", "1, 0": "qkv", "20, 0": "proj_or_fc2", "3, 0": "fc1", "4, 0": "fc2_last"}  # <256, EPI, patch(, pair)>


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--model", default="vit_base_patch16_224")
    ap.add_argument("--dim", type=int, default=768)
    a = ap.parse_args()
    rows = list(csv.reader(sys.stdin))
    h = rows[0]
    col = {n: h.index(n) for n in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum",
                                   "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")}
    unit = {n: rows[1][i] for n, i in col.items()}
    scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
    M, D = a.batch * 1568, a.dim
    alg = {"patch_embed": a.batch * 3 * 16 * 224 * 224 * 2 + D * 1536 * 2 + M * D * 2,
           "qkv": 2 * (M * D + 3 * D * D + M * 3 * D), "proj": 2 * (M * D + D * D + 2 * M * D),
           "fc1": 2 * (M * D + 4 * D * D + M * 4 * D), "fc2": 2 * (M * 4 * D + 4 * D * D + 2 * M * D)}
    per, detail, n_resid = {}, [], 0
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        if "gemm_kernel" not in name:
            continue
        key = next((v for k, v in EPI.items() if f"<256, {k}>" in name or f"<256, {k}, " in name), None)
        if key == "proj_or_fc2":
            key = "proj" if n_resid % 2 == 0 else "fc2"
            n_resid += 1
        if key == "fc2_last":
            key = "fc2"
        rd = float(r[col["dram__bytes_read.sum"]]) * scale[unit["dram__bytes_read.sum"]]
        wr = float(r[col["dram__bytes_write.sum"]]) * scale[unit["dram__bytes_write.sum"]]
        detail.append({"kernel": key, "template": name[name.find("gemm_kernel"):][:28], "dram_read": rd, "dram_write": wr,
                       "us": float(r[col["gpu__time_duration.sum"]]),
                       "tensor_pipe_active_pct": float(r[col["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]])})
        per.setdefault(key, rd + wr)
    out = {"model": a.model, "batch": a.batch, "dram_bytes_per_launch": per, "algorithmic_bytes_per_launch": alg,
           "launches": detail,
           "source": "ncu --set full --clock-control none (one launch of each GEMM of a block, pair tiles, cold cache, serialised); "
                     "writes below the algorithmic figure stay in the 126 MB L2 for the next kernel"}
    path = os.path.join(ROOT, "profiles", "ncu_gemm_dram.json")
    json.dump(out, open(path, "w"), indent=1)
    print(json.dumps(per))


if __name__ == "__main__":
    main()
