#!/bin/bash
# Clean-loop evidence for the fixed device fault: N short bench runs per workload on one B200, rc and captions/s each.
set -u
OUT=gpurun_out
mkdir -p $OUT
LOG=$OUT/r02_clean_loops.log
: > $LOG
run() {
    local wl=$1 i=$2
    local t0=$(date +%s)
    timeout 120 python bench.py --steps 20 --warmup 5 --skip-cpu --no-others --workload $wl --hang-seconds 40 --min-timed-ms 200 \
        > $OUT/loop.out 2> $OUT/loop.err
    local rc=$?
    local val=$(python -c "
import json
try:
    d = json.loads(open('$OUT/loop.out').read().strip().splitlines()[-1]); print(round(d['value']), round(d['e2e']['value']))
except Exception as e:
    print('no-json')")
    echo "$wl run $i rc=$rc $(( $(date +%s) - t0 ))s captions/s (device, e2e): $val $(grep -h 'no progress\|fault records' $OUT/loop.err | cut -c1-200)" | tee -a $LOG
}
for i in $(seq 1 ${1:-10}); do run object_relation $i; done
for i in $(seq 1 ${2:-6}); do run standard_grid $i; done
for i in $(seq 1 ${3:-4}); do run meshed_memory $i; done
echo "clean runs: $(grep -c 'rc=0' $LOG) of $(wc -l < $LOG)" | tee -a $LOG
