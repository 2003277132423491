#!/bin/bash
# One GPU visit of a round: parity tests, the bench line, the ncu launch list of the same command
# and `--set full` captures of the kernels that carry the step.  Run on the GPU box:
#     gpurun --timeout 2400 -- 'bash tools/gpu_round.sh r02h'
# Everything lands in gpurun_out/<tag>_*; copy what should be judged to profiles/.
set -u
TAG=${1:-run}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > $OUT/${TAG}_gpu.txt 2>&1
nproc >> $OUT/${TAG}_gpu.txt

if [ "${SKIP_TESTS:-0}" != 1 ]; then
    timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.txt 2>&1
    echo "pytest exit: $?" >> $OUT/${TAG}_pytest.txt
    tail -3 $OUT/${TAG}_pytest.txt
fi

# the bench line as the driver runs it (all configs), with its wall clock
t0=$(date +%s)
timeout 1500 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.log
echo "bench exit: $? wall: $(( $(date +%s) - t0 )) s" | tee -a $OUT/${TAG}_bench.log
t0=$(date +%s)
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_reference_arm.json 2>> $OUT/${TAG}_bench.log
echo "reference arm exit: $? wall: $(( $(date +%s) - t0 )) s" | tee -a $OUT/${TAG}_bench.log

if [ "${SKIP_NCU:-0}" != 1 ]; then
    # launch list (cold-cache, serialised: shares of the step, not absolute times)
    MCG_BENCH_CONFIGS="" MCG_BENCH_NO_PYTHON_REF=1 timeout 900 ncu --metrics gpu__time_duration.sum \
        --clock-control none -c 700 --csv --log-file $OUT/${TAG}_launches.csv \
        python bench.py --steps 2 --warmup 3 --e2e-steps 1 --cpu-budget 1 > $OUT/${TAG}_ncu_bench.log 2>&1
    echo "ncu launch list exit: $?"
    MCG_BENCH_CONFIGS="" MCG_BENCH_NO_PYTHON_REF=1 MCG_BENCH_CONTIGS=125000 MCG_BENCH_SAMPLES=50 \
        MCG_BENCH_BINS=500 MCG_BENCH_KIND=megahit timeout 900 ncu \
        --metrics gpu__time_duration.sum --clock-control none -c 700 --csv \
        --log-file $OUT/${TAG}_launches_cfg3_shard.csv \
        python bench.py --steps 2 --warmup 3 --e2e-steps 1 --cpu-budget 1 > $OUT/${TAG}_ncu_bench3.log 2>&1
    echo "ncu cfg3 launch list exit: $?"
    # full captures: one launch of every kernel of the resident step
    MCG_BENCH_CONFIGS="" MCG_BENCH_NO_PYTHON_REF=1 timeout 1200 ncu --set full --clock-control none \
        --import-source on -k regex:'scan_kernel|row_norms|cov_terms|tc_|bound_|cand_|dist2|prob_kernel' \
        --launch-skip 45 -c 28 -o $OUT/${TAG}_full -f \
        python bench.py --steps 2 --warmup 3 --e2e-steps 1 --cpu-budget 1 > $OUT/${TAG}_ncu_full.log 2>&1
    echo "ncu full exit: $?"
    python tools/ncu_summary.py $OUT/${TAG}_full.ncu-rep > $OUT/${TAG}_ncu_full_summary.txt 2>&1
fi
ls -la $OUT | tail -20
