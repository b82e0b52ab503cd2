#!/bin/bash
# Multi-GPU check, launched exactly as the driver does: gpurun --gpus N -- 'bash tools/r2_multi.sh N [repeats]'
set -u
N=${1:-4}
REP=${2:-2}
OUT=gpurun_out
mkdir -p $OUT
LOG=$OUT/r2_multi_n$N.log
: > $LOG
nvidia-smi -L | tee -a $LOG
for i in $(seq 1 $REP); do
    echo "== N=$N run $i" | tee -a $LOG
    t0=$(date +%s)
    timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + i)) \
        bench.py --gpus $N --steps 20 --warmup 5 > $OUT/multi_n${N}_$i.out 2> $OUT/multi_n${N}_$i.err
    rc=$?
    echo "   rc=$rc $(( $(date +%s) - t0 ))s" | tee -a $LOG
    python - $OUT/multi_n${N}_$i.out <<'PY' | tee -a $LOG
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(f"   value {d['value']:.0f} e2e {d['e2e']['value']:.0f} n_gpus {d['n_gpus']} ms/step {d['ms_per_step']:.3f}")
except Exception as e:
    print("   no JSON line:", e)
PY
    grep -n "fault records\|Error\|error\|Traceback" $OUT/multi_n${N}_$i.err | head -10 | tee -a $LOG
done
