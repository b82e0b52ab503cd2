#!/bin/bash
# Round-2 GPU call 9 (one B200): suite, the bench lines of the three single-GPU workloads, launch list and ncu captures.
set -u
OUT=gpurun_out
mkdir -p $OUT
LOG=$OUT/r2_call9.log
: > $LOG
step() {
    local name=$1 limit=$2; shift 2
    echo "== $name" | tee -a $LOG
    local t0=$(date +%s)
    timeout "$limit" "$@" > "$OUT/$name.out" 2> "$OUT/$name.err"
    local rc=$?
    echo "   rc=$rc $(( $(date +%s) - t0 ))s $(grep -h 'no progress\|fault records\|flight recorder' "$OUT/$name.err" | cut -c1-500 | tr '\n' ' ') ($(tail -c 250 "$OUT/$name.out" | tr '\n' ' '))" | tee -a $LOG
}
step c9_tests_gpu 1500 python -m pytest tests -q -m gpu -s
step c9_bench 300 python bench.py --steps 20 --warmup 5
step c9_bench_ort 300 python bench.py --steps 20 --warmup 5 --workload object_relation
step c9_bench_m2 300 python bench.py --steps 20 --warmup 5 --workload meshed_memory
step c9_trace_chain 200 python tools/trace_chain.py
step c9_one_batch 200 python tools/one_batch.py
# launch list of two eager batches (the second is warm): per-launch durations, cold-cache and serialised
step c9_launches 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $OUT/r02_launches.csv python tools/one_batch.py
# ncu --set full: the encoder and the first two decode steps of the warm batch, then decode step 10
step c9_ncu_head 900 ncu --set full --clock-control none --launch-skip 310 --launch-count 40 -o $OUT/r02_ncu_head -f python tools/one_batch.py
step c9_ncu_step10 900 ncu --set full --clock-control none --import-source on --launch-skip 470 --launch-count 15 -o $OUT/r02_ncu_step10 -f python tools/one_batch.py
for r in r02_ncu_head r02_ncu_step10; do
    ncu -i $OUT/$r.ncu-rep --page raw --csv > $OUT/$r.raw.csv 2> /dev/null
done
ls -la $OUT/*.ncu-rep $OUT/*.raw.csv | tee -a $LOG
cat $LOG
