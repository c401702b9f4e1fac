"""profiles/rN_ncu_full_selected.csv + rN_traffic.json from the raw page of one `ncu --set full` capture
(ncu -i X.ncu-rep --page raw --csv > X.csv).   Usage: python tools/select_ncu_metrics.py X.csv out_selected.csv out_traffic.json [tag]"""
import csv, json, statistics, sys

WANT = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic"]
rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}


def find(name):
    for h, i in col.items():
        if h == name or h.endswith("." + name) or h.endswith(name):
            return i
    return None


idx = [(m, find(m)) for m in WANT]
idx = [(m, i) for m, i in idx if i is not None]
out = [["ID", "Kernel Name", "Block Size", "Grid Size"] + [m for m, _ in idx], ["", "", "", ""] + [units[i] for _, i in idx]]
per = {}
for r in rows[2:]:
    if len(r) != len(hdr):
        continue
    name = r[col["Kernel Name"]].split("(")[0]
    out.append([r[col["ID"]], name, r[col["Block Size"]], r[col["Grid Size"]]] + [r[i] for _, i in idx])

    def val(m):
        i = find(m)
        v = float(r[i].replace(",", "")) if r[i] else 0.0
        u = units[i]
        return v * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}.get(u, 1.0)
    per.setdefault(name, []).append(val("dram__bytes_read.sum") + val("dram__bytes_write.sum"))
csv.writer(open(sys.argv[2], "w", newline="")).writerows(out)
names = {"unnamed>::evaluate_kernel": "evaluate_kernel(candidate)", "unnamed>::accumulate_kernel": "accumulate_kernel", "unnamed>::chol_kernel": "step_kernel"}
traffic = {names.get(k, k): int(statistics.median(v)) for k, v in per.items()}
json.dump({"source": f"{sys.argv[2]} (ncu --set full --clock-control none, median over the captured launches of a running solve, tools/gpu_evidence_r2.sh {sys.argv[4] if len(sys.argv) > 4 else ''})",
           "dram_bytes_per_launch": traffic}, open(sys.argv[3], "w"), indent=1)
print(len(out) - 2, "launches;", traffic)
