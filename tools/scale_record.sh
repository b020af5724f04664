#!/bin/bash
# Record lines of one workload at N GPUs: tools/scale_record.sh <ngpus> <workload> <tag>
N=${1:-8}; WL=${2:-cfg3}; TAG=${3:-rec}
if [ "$N" = "1" ]; then L="python"; else L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512"; fi
$L bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-eager-arm --workload $WL \
   > gpurun_out/${TAG}_${WL}_${N}gpu.json 2> gpurun_out/${TAG}_${WL}_${N}gpu.err
python - <<P
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_${WL}_${N}gpu.json").read().strip().splitlines()[-1])
    print("${WL} N=${N}: value %.0f  ms %.3f  e2e %.0f (%.3f ms)  no_overlap %s  opt %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], (d.get("no_overlap") or {}).get("ms_per_step"), (d.get("with_optimizer_step") or {}).get("ms_per_step")))
except Exception as e:
    print("${WL} N=${N}: failed", e); print(open("gpurun_out/${TAG}_${WL}_${N}gpu.err").read()[-1500:])
P
