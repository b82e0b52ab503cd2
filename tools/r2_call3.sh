#!/bin/bash
# Round-2 GPU call 3 (one B200): streamed cross-attention (numerics, launch time, grid/stage sweep), the reworked parity
# tests (near-tie at first divergence, bf16-operand yardstick, full-size configs), bench + timing ablations.
set -u
OUT=gpurun_out
mkdir -p $OUT
LOG=$OUT/r2_call3.log
: > $LOG
step() {
    local name=$1 limit=$2; shift 2
    echo "== $name" | tee -a $LOG
    local t0=$(date +%s)
    timeout "$limit" "$@" > "$OUT/$name.out" 2> "$OUT/$name.err"
    echo "   rc=$? $(( $(date +%s) - t0 ))s ($(tail -c 600 "$OUT/$name.out" | tr '\n' ' '))" | tee -a $LOG
}
step c3_probe_cross 200 python tests/gpu_scripts/probe_cross_stream.py
step c3_probe_cross_i2 200 env OPENVIIC_XATTN_IMAGES=2 python tests/gpu_scripts/probe_cross_stream.py
step c3_probe_cross_i8 200 env OPENVIIC_XATTN_IMAGES=8 python tests/gpu_scripts/probe_cross_stream.py
step c3_tests_gpu 1500 python -m pytest tests -q -m gpu -s
step c3_bench 300 python bench.py --steps 20 --warmup 5
step c3_bench_i2 200 env OPENVIIC_XATTN_IMAGES=2 python bench.py --skip-cpu --steps 20
step c3_bench_i8 200 env OPENVIIC_XATTN_IMAGES=8 python bench.py --skip-cpu --steps 20
step c3_bench_tc1 200 env OPENVIIC_CROSS_TC=1 python bench.py --skip-cpu --steps 20
for ab in 1 2 4 8; do
step c3_ablate_$ab 200 env OPENVIIC_DBG_ABLATE=$ab python bench.py --skip-cpu --steps 20
done
step c3_stress 300 python tools/stress.py --iters 100 --seconds 60
cat $LOG
