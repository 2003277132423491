#!/bin/bash
# compute-sanitizer over the smoke pass and the small parity tests (race /
# memory checking of the kernels).  Run on the GPU box:
#     gpurun --timeout 1500 -- 'bash tools/sanitize.sh'
# Summaries land in gpurun_out/sanitize_*.txt; copy them to profiles/ (profiles/README.md).
# racecheck looks at shared memory: scan_kernel's private-column byte counters (non-atomic
# read-modify-writes by design, csrc/mcg_scan.cu), pair_kernel / dist2_kernel staging,
# walk_kernel's per-thread table columns.  memcheck covers global and shared accesses.
set -u
mkdir -p gpurun_out
SAN=/usr/local/cuda/bin/compute-sanitizer
TESTS="tests/test_gpu_parity.py::test_scan_edge_lengths tests/test_gpu_parity.py::test_scan_golden_ascii \
tests/test_gpu_parity.py::test_rule_c_golden tests/test_gpu_parity.py::test_rules_a_b_golden \
tests/test_gpu_parity.py::test_fused_call_matches_staged tests/test_gpu_parity.py::test_bfs_depth_overwrite_quirk \
tests/test_resident.py::test_bfs_resident_triangle_quirk_and_empty tests/test_resident.py::test_labels_set_patches_single_entries_and_batches \
tests/test_prune.py::test_candidate_list_overflow_falls_back_to_all_pairs"
for tool in memcheck racecheck; do
    log=gpurun_out/sanitize_${tool}.txt
    echo "# compute-sanitizer --tool $tool ($(date -u +%FT%TZ))" > $log
    echo "## smoke" >> $log
    timeout 600 $SAN --tool $tool --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" >> $log 2>&1
    echo "exit: $?" >> $log
    echo "## pytest $TESTS" >> $log
    timeout 1200 $SAN --tool $tool --print-limit 20 python -m pytest -x -q -m gpu $TESTS >> $log 2>&1
    echo "exit: $?" >> $log
    grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|smoke ok|exit:" $log > gpurun_out/sanitize_${tool}_summary.txt
done
cat gpurun_out/sanitize_*_summary.txt
