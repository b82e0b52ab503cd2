#!/bin/bash
# Round-2 GPU call 7 (one B200): the pair-allocation deadlock fix (one cta_group::2 CTA per SM) under the reproducer, the
# whole GPU suite after the clean-up, bench with / without the pair GEMM, stress.
set -u
OUT=gpurun_out
mkdir -p $OUT
LOG=$OUT/r2_call7.log
: > $LOG
step() {
    local name=$1 limit=$2; shift 2
    echo "== $name" | tee -a $LOG
    local t0=$(date +%s)
    timeout "$limit" "$@" > "$OUT/$name.out" 2> "$OUT/$name.err"
    local rc=$?
    echo "   rc=$rc $(( $(date +%s) - t0 ))s $(grep -h 'no progress\|fault records\|flight recorder' "$OUT/$name.err" | cut -c1-600 | tr '\n' ' ') ($(tail -c 300 "$OUT/$name.out" | tr '\n' ' '))" | tee -a $LOG
}
BENCH="python bench.py --steps 20 --warmup 5 --workload object_relation --skip-cpu --hang-seconds 30 --min-timed-ms 0"
for i in 1 2 3 4 5 6; do step c7_fix_$i 120 env OPENVIIC_FLIGHT=1 $BENCH; done
step c7_tests_gpu 1500 python -m pytest tests -q -m gpu -s
step c7_bench 300 python bench.py --steps 20 --warmup 5
step c7_bench_no2cta 200 env OPENVIIC_GEMM_2CTA=0 python bench.py --steps 20 --warmup 5 --skip-cpu
step c7_bench_ort 200 python bench.py --steps 20 --warmup 5 --workload object_relation --skip-cpu
step c7_bench_m2 200 python bench.py --steps 20 --warmup 5 --workload meshed_memory --skip-cpu
step c7_stress 200 python tools/stress.py --iters 100 --seconds 40 --host-every 2
cat $LOG
