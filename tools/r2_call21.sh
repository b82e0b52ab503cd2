#!/bin/bash
# Round-2 GPU call 21 (one B200): single-instruction exp2 in the vocabulary / softmax loops.
set -u
OUT=gpurun_out
mkdir -p $OUT
LOG=$OUT/r2_call21.log
: > $LOG
step() {
    local name=$1 limit=$2; shift 2
    echo "== $name" | tee -a $LOG
    local t0=$(date +%s)
    timeout "$limit" "$@" > "$OUT/$name.out" 2> "$OUT/$name.err"
    local rc=$?
    echo "   rc=$rc $(( $(date +%s) - t0 ))s ($(tail -c 300 "$OUT/$name.out" | tr '\n' ' '))" | tee -a $LOG
}
step c21_tests 900 python -m pytest tests/test_gpu_fused_decode.py tests/test_gpu_engine.py tests/test_gpu_fullsize.py tests/test_gpu_ops.py tests/test_gpu_modules.py -q -x
step c21_bench 300 python bench.py --steps 20 --warmup 5 --skip-cpu
step c21_trace 200 python tools/trace_chain.py
cat $LOG
