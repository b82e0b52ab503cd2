#!/bin/bash
# Round-2 GPU call 13 (one B200): pipelined chain epilogues, training step (tests + throughput).
set -u
OUT=gpurun_out
mkdir -p $OUT
LOG=$OUT/r2_call14.log
: > $LOG
step() {
    local name=$1 limit=$2; shift 2
    echo "== $name" | tee -a $LOG
    local t0=$(date +%s)
    timeout "$limit" "$@" > "$OUT/$name.out" 2> "$OUT/$name.err"
    local rc=$?
    echo "   rc=$rc $(( $(date +%s) - t0 ))s $(grep -h 'no progress\|fault records\|flight recorder' "$OUT/$name.err" | cut -c1-500 | tr '\n' ' ') ($(tail -c 300 "$OUT/$name.out" | tr '\n' ' '))" | tee -a $LOG
}
step c14_fused 900 python -m pytest tests/test_gpu_fused_decode.py tests/test_gpu_engine.py tests/test_gpu_fullsize.py -q -x -s
step c14_bench 300 python bench.py --steps 20 --warmup 5 --skip-cpu
step c14_bench_m2 300 python bench.py --steps 20 --warmup 5 --skip-cpu --workload meshed_memory
step c14_trace 200 python tools/trace_chain.py
cat $LOG
