#!/bin/bash
# Round-2 GPU call 2 (one B200): suite incl. the new full-size parity tests, bench with shared weights, stress.
set -u
OUT=gpurun_out
mkdir -p $OUT
LOG=$OUT/r2_call2.log
: > $LOG
step() {
    local name=$1 limit=$2; shift 2
    echo "== $name" | tee -a $LOG
    local t0=$(date +%s)
    timeout "$limit" "$@" > "$OUT/$name.out" 2> "$OUT/$name.err"
    echo "   rc=$? $(( $(date +%s) - t0 ))s ($(tail -c 400 "$OUT/$name.out" | tr '\n' ' '))" | tee -a $LOG
}
step c2_tests_gpu 900 python -m pytest tests -q -m gpu -s
step c2_bench 300 python bench.py --steps 20 --warmup 5
step c2_bench_tc0 200 env OPENVIIC_CROSS_TC=0 python bench.py --skip-cpu --steps 20
step c2_bench_enc_tc 200 env OPENVIIC_ENC_TC=1 python bench.py --skip-cpu --steps 20
step c2_stress 300 python tools/stress.py --iters 120 --seconds 90
cat $LOG
