#!/bin/bash
# Round-2 GPU call 1 (one B200): baseline suite, the stress / determinism loop that hunts the round-1 device fault
# (GPU core dump on exception), the never-run attention variants, bench with each switch.
set -u
OUT=gpurun_out
mkdir -p $OUT
LOG=$OUT/r2_call1.log
: > $LOG
step() {   # step <name> <timeout-seconds> <command...>
    local name=$1 limit=$2; shift 2
    echo "== $name" | tee -a $LOG
    local t0=$(date +%s)
    timeout "$limit" "$@" > "$OUT/$name.out" 2> "$OUT/$name.err"
    echo "   rc=$? $(( $(date +%s) - t0 ))s ($(tail -c 400 "$OUT/$name.out" | tr '\n' ' '))" | tee -a $LOG
}
step tests_gpu 600 python -m pytest tests -x -q -m gpu
step stress_default 400 python tools/stress.py --iters 150 --seconds 120
step bench_default 300 python bench.py --steps 20 --warmup 5
step probe_attention 300 python tests/gpu_scripts/probe_attention_variants.py
step bench_cross_tc 200 env OPENVIIC_CROSS_TC=1 python bench.py --skip-cpu --steps 20
step bench_self_split 200 env OPENVIIC_SELF_SPLIT=1 python bench.py --skip-cpu --steps 20
step bench_enc_tc 200 env OPENVIIC_ENC_TC=1 python bench.py --skip-cpu --steps 20
step bench_all3 200 env OPENVIIC_CROSS_TC=1 OPENVIIC_SELF_SPLIT=1 OPENVIIC_ENC_TC=1 python bench.py --skip-cpu --steps 20
step bench_epi1 200 env OPENVIIC_CHAIN_EPI=1 python bench.py --skip-cpu --steps 20
step bench_epi2 200 env OPENVIIC_CHAIN_EPI=2 python bench.py --skip-cpu --steps 20
step stress_variants 300 env OPENVIIC_CROSS_TC=1 OPENVIIC_SELF_SPLIT=1 OPENVIIC_ENC_TC=1 python tools/stress.py --iters 100 --seconds 60
step stress_nopair 300 env OPENVIIC_CHAIN_PAIR=0 python tools/stress.py --iters 100 --seconds 60
ls -la $OUT/core_* 2>/dev/null | tee -a $LOG
cat $LOG
