#!/usr/bin/env python3
"""profiles/traffic.json from `ncu --set full` captures: per-launch DRAM bytes and L1TEX data-pipe
wavefronts of the kernels bench.py reports a roofline for.

    python tools/traffic_from_ncu.py <workload key> <commit> <summary name> rep [rep ...]

workload key: "<length kind>-<contigs>-<samples>-<bins>" (bench.py's wl_key); commit: the commit
whose kernels the capture ran (the entry is valid while those kernel sources are unchanged).
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KERNELS = ("scan_kernel", "dist2_kernel", "prob_kernel", "tc_dist_kernel", "bound_kernel",
           "cand_score_kernel", "row_norms_kernel", "tc_prep_kernel")


def main():
    key, commit, source = sys.argv[1:4]
    entry = {"source": source, "commit": commit}
    for rep in sys.argv[4:]:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True,
                             text=True, check=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        head = rows[0]
        col = {h: i for i, h in enumerate(head)}
        for row in rows[2:]:
            name = row[col["Kernel Name"]]
            short = next((k for k in KERNELS if k in name), None)
            if short is None or short in entry:
                continue

            def val(metric):
                return float(row[col[metric]].replace(",", ""))
            unit_r = rows[1][col["dram__bytes_read.sum"]]
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit_r]
            unit_w = rows[1][col["dram__bytes_write.sum"]]
            scale_w = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit_w]
            entry[short] = int(val("dram__bytes_read.sum") * scale + val("dram__bytes_write.sum") * scale_w)
            if short == "scan_kernel":
                entry["scan_kernel_smem_wavefronts"] = int(val("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"))
                # shared + global: the per-SM average of the triage section x the SMs of the launch
                entry["scan_kernel_lsu_wavefronts"] = int(
                    val("SM_A.TriageCompute.l1tex__data_pipe_lsu_wavefronts.avg") * val("launch__grid_size"))
    path = os.path.join(ROOT, "profiles", "traffic.json")
    data = json.load(open(path))
    data[key] = entry
    json.dump(data, open(path, "w"), indent=1)
    print(json.dumps(entry, indent=1))


if __name__ == "__main__":
    main()
