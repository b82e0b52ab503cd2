#!/bin/bash
# Round-2 GPU call 15 (one B200): the final library -- full suite, stress, bench lines, training throughput, ncu evidence.
set -u
OUT=gpurun_out
mkdir -p $OUT
LOG=$OUT/r2_call15.log
: > $LOG
step() {
    local name=$1 limit=$2; shift 2
    echo "== $name" | tee -a $LOG
    local t0=$(date +%s)
    timeout "$limit" "$@" > "$OUT/$name.out" 2> "$OUT/$name.err"
    local rc=$?
    echo "   rc=$rc $(( $(date +%s) - t0 ))s $(grep -h 'no progress\|fault records\|flight recorder' "$OUT/$name.err" | cut -c1-500 | tr '\n' ' ') ($(tail -c 300 "$OUT/$name.out" | tr '\n' ' '))" | tee -a $LOG
}
step c15_tests_gpu 1800 python -m pytest tests -q -m gpu -s
step c15_stress 200 python tools/stress.py --iters 100 --seconds 40 --host-every 2
step c15_stress_m2 200 python tools/stress.py --iters 60 --seconds 40 --host-every 2 --workload meshed_memory
step c15_bench 400 python bench.py --steps 20 --warmup 5
step c15_bench_train 300 python tools/bench_train.py
step c15_trace_chain 200 python tools/trace_chain.py
step c15_launches 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $OUT/r02_launches.csv python tools/one_batch.py
step c15_ncu_head 900 ncu --set full --clock-control none --launch-skip 310 --launch-count 26 -o /tmp/r02_ncu_head -f python tools/one_batch.py
ncu -i /tmp/r02_ncu_head.ncu-rep --page raw --csv > $OUT/r02_ncu_head.raw.csv 2> /dev/null
step c15_ncu_chain 900 ncu --set full --clock-control none --import-source on -k regex:decode_chain --launch-skip 218 --launch-count 7 -o /tmp/r02_ncu_chain -f python tools/one_batch.py
ncu -i /tmp/r02_ncu_chain.ncu-rep --page raw --csv > $OUT/r02_ncu_chain.raw.csv 2> /dev/null
ncu -i /tmp/r02_ncu_chain.ncu-rep --page source --csv 2> /dev/null | gzip -9 > $OUT/r02_ncu_chain.source.csv.gz
du -sh $OUT | tee -a $LOG
cat $LOG
