#!/bin/bash
# Round-2 GPU call 5 (one B200): new module-path tests (Camo, DLCT, adaptive attention); hang hunt with the flight
# recorder on (bench e2e loops and host-heavy stress, each under a watchdog).
set -u
OUT=gpurun_out
mkdir -p $OUT
LOG=$OUT/r2_call5.log
: > $LOG
step() {
    local name=$1 limit=$2; shift 2
    echo "== $name" | tee -a $LOG
    local t0=$(date +%s)
    timeout "$limit" "$@" > "$OUT/$name.out" 2> "$OUT/$name.err"
    echo "   rc=$? $(( $(date +%s) - t0 ))s ($(tail -c 500 "$OUT/$name.out" | tr '\n' ' '))" | tee -a $LOG
}
step c5_tests_new 600 python -m pytest tests/test_gpu_modules.py tests/test_gpu_fused_decode.py -q -s
export OPENVIIC_FLIGHT=1
for i in 1 2 3 4; do
step c5_hunt_ort_$i 150 python bench.py --steps 20 --warmup 5 --workload object_relation --skip-cpu --hang-seconds 45
done
for i in 1 2 3; do
step c5_hunt_stress_$i 150 python tools/stress.py --iters 80 --seconds 30 --host-every 2 --hang-seconds 30
done
for i in 1 2; do
step c5_hunt_std_$i 150 python bench.py --steps 20 --warmup 5 --skip-cpu --hang-seconds 45
done
grep -l "no progress\|\"hang\": true" $OUT/c5_hunt_*.err $OUT/c5_hunt_*.out 2>/dev/null | tee -a $LOG
cat $LOG
