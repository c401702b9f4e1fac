"""The two scripts that reduce ncu exports to the tables under profiles/ run on the committed evidence and agree with what
profiles/r2_summary.md states (the dominant kernel of the launch list; the DRAM traffic bench.py reads for `roofline.traffic`)."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def test_launch_list_summary_names_the_step_kernel_first():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sum_ncu_launches.py"), os.path.join(ROOT, "profiles", "r2_launches.csv")],
                         capture_output=True, text=True, check=True).stdout.splitlines()
    assert "chol_kernel" in out[0] and "evaluate_kernel" in out[1]
    share = float(out[0].split("share=")[1])
    assert 0.3 < share < 0.5
    assert not any("nvjet" in ln or "cublas" in ln.lower() or "cutlass" in ln.lower() for ln in out)  # only the repo's own kernels


def test_traffic_file_matches_the_selected_capture():
    import csv

    tr = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))["dram_bytes_per_launch"]
    rows = list(csv.reader(open(os.path.join(ROOT, "profiles", "r2_ncu_full_selected.csv"))))
    hdr = rows[0]
    i_rd, i_wr, i_name = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("Kernel Name")
    units = rows[1]
    scale = {"Kbyte": 1e3, "Mbyte": 1e6, "byte": 1.0}
    chol = sorted(float(r[i_rd]) * scale[units[i_rd]] + float(r[i_wr]) * scale[units[i_wr]] for r in rows[2:] if "chol_kernel" in r[i_name])
    assert chol and chol[0] <= tr["step_kernel"] <= chol[-1]
    assert 3.0e5 < tr["step_kernel"] < 5.0e5  # ~1.35 x the 305 KB of algorithmic bytes per factoring launch
