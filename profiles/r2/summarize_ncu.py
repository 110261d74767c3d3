"""ncu raw-page CSV -> a small JSON/markdown summary (one row per profiled launch):
  python profiles/r2/summarize_ncu.py raw.csv out.json"""
import csv
import json
import re
import sys

KEEP = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active": "tensor_hmma_pct",
    "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active": "tensor_inst_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_active_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
    "launch__registers_per_thread": "regs",
    "launch__grid_size": "grid",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "l1tex__data_bank_conflicts_pipe_lsu.sum": "smem_bank_conflicts",
}
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
tens = [h for h in hdr if "tensor" in h and "pct" in h]
out = []
for r in rows[2:]:
    name = re.sub(r"\(.*", "", r[ix["Kernel Name"]]).replace("void ", "").replace("pg::tc::", "").replace("pg::", "")
    rec = {"kernel": name[:90]}
    for k, short in KEEP.items():
        if k in ix and r[ix[k]] != "":
            try:
                rec[short] = float(r[ix[k]].replace(",", ""))
                rec[short + "_unit"] = units[ix[k]]
            except ValueError:
                pass
    for h in tens:
        if h not in KEEP and r[ix[h]] not in ("", "n/a"):
            try:
                rec[h] = float(r[ix[h]].replace(",", ""))
            except ValueError:
                pass
    out.append(rec)
json.dump(out, open(sys.argv[2], "w"), indent=1)
for rec in out:
    print(rec["kernel"][:70], {k: v for k, v in rec.items() if k in ("duration", "dram_read", "dram_write", "dram_pct", "tensor_active_pct", "regs", "issue_active_pct")})
