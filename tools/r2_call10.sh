#!/bin/bash
# Round-2 GPU call 10 (one B200): suite on the final library, bench lines of the three single-GPU workloads, launch list,
# ncu captures exported to CSV on the box (the .ncu-rep files are too large to bring back).
set -u
OUT=gpurun_out
mkdir -p $OUT
LOG=$OUT/r2_call10.log
: > $LOG
step() {
    local name=$1 limit=$2; shift 2
    echo "== $name" | tee -a $LOG
    local t0=$(date +%s)
    timeout "$limit" "$@" > "$OUT/$name.out" 2> "$OUT/$name.err"
    local rc=$?
    echo "   rc=$rc $(( $(date +%s) - t0 ))s $(grep -h 'no progress\|fault records\|flight recorder' "$OUT/$name.err" | cut -c1-500 | tr '\n' ' ') ($(tail -c 250 "$OUT/$name.out" | tr '\n' ' '))" | tee -a $LOG
}
step c10_tests_gpu 1500 python -m pytest tests -q -m gpu -s
step c10_bench 300 python bench.py --steps 20 --warmup 5
step c10_bench_ort 300 python bench.py --steps 20 --warmup 5 --workload object_relation
step c10_bench_m2 300 python bench.py --steps 20 --warmup 5 --workload meshed_memory
step c10_trace_chain 200 python tools/trace_chain.py
step c10_launches 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $OUT/r02_launches.csv python tools/one_batch.py
# ncu --set full: the encoder and the first decode step of the warm batch (no source import: size)
step c10_ncu_head 900 ncu --set full --clock-control none --launch-skip 310 --launch-count 26 -o /tmp/r02_ncu_head -f python tools/one_batch.py
ncu -i /tmp/r02_ncu_head.ncu-rep --page raw --csv > $OUT/r02_ncu_head.raw.csv 2> /dev/null
# the seven chain launches of decode step 10, with source
step c10_ncu_chain 900 ncu --set full --clock-control none --import-source on -k regex:decode_chain --launch-skip 218 --launch-count 7 -o /tmp/r02_ncu_chain -f python tools/one_batch.py
ncu -i /tmp/r02_ncu_chain.ncu-rep --page raw --csv > $OUT/r02_ncu_chain.raw.csv 2> /dev/null
ncu -i /tmp/r02_ncu_chain.ncu-rep --page source --csv 2> /dev/null | gzip -9 > $OUT/r02_ncu_chain.source.csv.gz
ls -la /tmp/*.ncu-rep $OUT/*.csv $OUT/*.gz | tee -a $LOG
du -sh $OUT | tee -a $LOG
cat $LOG
