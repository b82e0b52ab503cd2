#!/bin/bash
# First GPU call of round 2 (one B200): everything written after round 1's GPU budget ran out, in one go.
#   gpurun --timeout 1500 -- 'bash tools/r2_first_call.sh'
# Every step has its own timeout and log under gpurun_out/; a failing step does not stop the later ones.
set -u
OUT=gpurun_out
mkdir -p $OUT
step() {   # step <name> <timeout-seconds> <command...>
    local name=$1 limit=$2; shift 2
    echo "== $name" | tee -a $OUT/r2_first_call.log
    timeout "$limit" "$@" > "$OUT/$name.out" 2> "$OUT/$name.err"
    echo "   rc=$? ($(tail -c 300 "$OUT/$name.out" | tr '\n' ' '))" | tee -a $OUT/r2_first_call.log
}
# 1. the new GPU tests first (predictor), then the whole suite
step tests_predict 400 python -m pytest tests/test_gpu_predict.py -x -q
step tests_gpu 900 python -m pytest tests -x -q -m gpu
# 2. the default bench line (new: saturated roofline figure, breadcrumbs on stderr)
step bench_default 400 python bench.py
step trace_chain 300 python tools/trace_chain.py
# 3. the opt-in attention kernels: numerics + launch times, then the suite and the bench with both switches on
step probe_attention 400 python tests/gpu_scripts/probe_attention_variants.py
step tests_gpu_variants 900 env OPENVIIC_CROSS_TC=1 OPENVIIC_SELF_SPLIT=1 OPENVIIC_ENC_TC=1 python -m pytest tests -x -q -m gpu
step bench_cross_tc 300 env OPENVIIC_CROSS_TC=1 python bench.py --skip-cpu
step bench_self_split 300 env OPENVIIC_SELF_SPLIT=1 python bench.py --skip-cpu
step bench_enc_tc 300 env OPENVIIC_ENC_TC=1 python bench.py --skip-cpu
step bench_both 300 env OPENVIIC_CROSS_TC=1 OPENVIIC_SELF_SPLIT=1 OPENVIIC_ENC_TC=1 python bench.py --skip-cpu
# 3b. the chain kernels with software-pipelined TMEM loads in the epilogues (separate instantiation, read at engine creation)
step tests_fused_epi 600 env OPENVIIC_CHAIN_EPI=1 python -m pytest tests/test_gpu_fused_decode.py tests/test_gpu_engine.py -x -q
step bench_chain_epi 300 env OPENVIIC_CHAIN_EPI=1 python bench.py --skip-cpu
step tests_fused_epi2 600 env OPENVIIC_CHAIN_EPI=2 python -m pytest tests/test_gpu_fused_decode.py tests/test_gpu_engine.py -x -q
step bench_chain_epi2 300 env OPENVIIC_CHAIN_EPI=2 python bench.py --skip-cpu
# 4. launch list of one batch with the variants on (share of the step per kernel), only after the runs above passed
step ncu_launches_variants 400 env OPENVIIC_CROSS_TC=1 OPENVIIC_SELF_SPLIT=1 OPENVIIC_ENC_TC=1 ncu --metrics gpu__time_duration.sum \
    --clock-control none -c 800 --csv --log-file $OUT/r02_launches_variants.csv python tools/one_batch.py
# 5. full captures of the two decode attention kernels (default and variant) for the roofline of the HBM-bound part
step ncu_attention_default 500 ncu --set full --clock-control none --import-source on -k regex:decode_.*attention --launch-skip 100 -c 12 \
    -o $OUT/r02_attention_default python tools/one_batch.py
step ncu_attention_variants 500 env OPENVIIC_CROSS_TC=1 OPENVIIC_SELF_SPLIT=1 OPENVIIC_ENC_TC=1 ncu --set full --clock-control none \
    --import-source on -k regex:decode_.*attention --launch-skip 100 -c 12 -o $OUT/r02_attention_variants python tools/one_batch.py
cat $OUT/r2_first_call.log
