#!/bin/bash
# Round-2 GPU call 8 (one B200): pair GEMM retired + encoder chains + PDL mask-read fix: suite, reproducer, benches.
set -u
OUT=gpurun_out
mkdir -p $OUT
LOG=$OUT/r2_call8.log
: > $LOG
step() {
    local name=$1 limit=$2; shift 2
    echo "== $name" | tee -a $LOG
    local t0=$(date +%s)
    timeout "$limit" "$@" > "$OUT/$name.out" 2> "$OUT/$name.err"
    local rc=$?
    echo "   rc=$rc $(( $(date +%s) - t0 ))s $(grep -h 'no progress\|fault records\|flight recorder' "$OUT/$name.err" | cut -c1-500 | tr '\n' ' ') ($(tail -c 250 "$OUT/$name.out" | tr '\n' ' '))" | tee -a $LOG
}
step c8_tests_gpu 1500 python -m pytest tests -q -m gpu -s
REPRO="python bench.py --steps 20 --warmup 5 --workload object_relation --skip-cpu --hang-seconds 30 --min-timed-ms 0"
for i in 1 2 3 4; do step c8_repro_$i 120 env OPENVIIC_FLIGHT=1 $REPRO; done
step c8_bench 300 python bench.py --steps 20 --warmup 5
step c8_bench_noenc 200 env OPENVIIC_ENC_CHAINS=0 python bench.py --steps 20 --warmup 5 --skip-cpu
step c8_bench_ort 200 python bench.py --steps 20 --warmup 5 --workload object_relation --skip-cpu
step c8_bench_m2 200 python bench.py --steps 20 --warmup 5 --workload meshed_memory --skip-cpu
step c8_bench_ort2 200 python bench.py --steps 20 --warmup 5 --workload object_relation --skip-cpu
step c8_stress 200 python tools/stress.py --iters 100 --seconds 40 --host-every 2
step c8_stress_ort 200 python tools/stress.py --iters 60 --seconds 40 --host-every 2 --workload object_relation
cat $LOG
