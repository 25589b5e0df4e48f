"""Condense `ncu --page raw --csv` and `--page source --csv` exports into the JSON summaries kept under profiles/.
usage: python tools/ncu_summary.py RAW.csv SOURCE.csv OUT.json ["note"]"""
import collections
import csv
import json
import re
import sys

raw, src, out = sys.argv[1], sys.argv[2], sys.argv[3]
note = sys.argv[4] if len(sys.argv) > 4 else ""
rows = list(csv.reader(open(raw)))
hdr, vals = rows[0], rows[2]
d = dict(zip(hdr, vals))
keep = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__waves_per_multiprocessor", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
        "smsp__inst_executed.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_eligible.avg.per_cycle_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
        "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
        "sm__cycles_active.avg", "dram__throughput.avg.pct_of_peak_sustained_elapsed"]
metrics = {}
for k in keep:
    if k in d and d[k] != "":
        try:
            metrics[k] = float(d[k].replace(",", ""))
        except ValueError:
            pass
stalls = {k.replace("smsp__pcsamp_warps_issue_stalled_", ""): int(float(v)) for k, v in d.items()
          if k.startswith("smsp__pcsamp_warps_issue_stalled_") and not k.endswith("_not_issued") and v not in ("", "0")}
srows = list(csv.reader(open(src)))
shdr = srows[1]
ix = {h: i for i, h in enumerate(shdr)}
ops = collections.Counter()
for r in srows[2:]:
    if len(r) != len(shdr):
        break
    t = r[ix["Source"]].split()
    if not t:
        continue
    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    ops[op] += int(r[ix["Instructions Executed"]])
json.dump({"kernel": d.get("Kernel Name", rows[2][hdr.index("Kernel Name")] if "Kernel Name" in hdr else ""), "note": note,
           "metrics": metrics, "warp_instructions_by_opcode": dict(ops.most_common(30)),
           "stall_samples": dict(sorted(stalls.items(), key=lambda kv: -kv[1]))}, open(out, "w"), indent=1)
print(out, metrics.get("gpu__time_duration.sum"), dict(ops.most_common(6)))
