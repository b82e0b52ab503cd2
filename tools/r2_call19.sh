#!/bin/bash
# Round-2 GPU call 19 (one B200): final full suite + launch list of the training step.
set -u
OUT=gpurun_out
mkdir -p $OUT
LOG=$OUT/r2_call19.log
: > $LOG
step() {
    local name=$1 limit=$2; shift 2
    echo "== $name" | tee -a $LOG
    local t0=$(date +%s)
    timeout "$limit" "$@" > "$OUT/$name.out" 2> "$OUT/$name.err"
    local rc=$?
    echo "   rc=$rc $(( $(date +%s) - t0 ))s ($(tail -c 300 "$OUT/$name.out" | tr '\n' ' '))" | tee -a $LOG
}
step c19_tests_gpu 1800 python -m pytest tests -q -m gpu
step c19_smoke 300 python __graft_entry__.py smoke
step c19_train_launches 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $OUT/r02_train_launches.csv python tools/one_train_step.py
cat $LOG
