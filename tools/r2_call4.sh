#!/bin/bash
# Round-2 GPU call 4 (one B200): job-table chains (plain + meshed) after the epilogue-race fix, module-path Camo, where a
# chain's time goes (issuer / producer wait accounting), cross-attention launch times, bench, stress with a watchdog.
set -u
OUT=gpurun_out
mkdir -p $OUT
LOG=$OUT/r2_call4.log
: > $LOG
step() {
    local name=$1 limit=$2; shift 2
    echo "== $name" | tee -a $LOG
    local t0=$(date +%s)
    timeout "$limit" "$@" > "$OUT/$name.out" 2> "$OUT/$name.err"
    echo "   rc=$? $(( $(date +%s) - t0 ))s ($(tail -c 700 "$OUT/$name.out" | tr '\n' ' '))" | tee -a $LOG
}
step c4_tests_gpu 1500 python -m pytest tests -q -m gpu -s
step c4_trace_chain 200 python tools/trace_chain.py
step c4_probe_cross 200 python tests/gpu_scripts/probe_cross_stream.py
step c4_bench 300 python bench.py --steps 20 --warmup 5
step c4_bench_m2 300 python bench.py --steps 20 --warmup 5 --workload meshed_memory --skip-cpu
step c4_bench_ort 300 python bench.py --steps 20 --warmup 5 --workload object_relation --skip-cpu
step c4_stress 200 python tools/stress.py --iters 100 --seconds 40
step c4_stress_tc1 200 env OPENVIIC_CROSS_TC=1 python tools/stress.py --iters 100 --seconds 40
step c4_stress_m2 200 python tools/stress.py --iters 60 --seconds 40 --workload meshed_memory
cat $LOG
