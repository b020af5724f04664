#!/bin/bash
# 8-GPU A/B runs of the data-parallel step: SMs reserved for NCCL / NCCL channel cap / e2e input format.
# usage: tools/scale_experiment.sh <ngpus> <tag>
N=${1:-8}; TAG=${2:-exp}
run() {  # name, env...
  name=$1; shift
  env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-eager-arm --no-optimizer-arm ${WL:+--workload $WL} \
      > gpurun_out/${TAG}_${name}.json 2> gpurun_out/${TAG}_${name}.err
  python - <<P
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_${name}.json").read().strip().splitlines()[-1])
    print("${name}: value %.0f  ms %.3f  e2e %.0f (%.3f ms)  no_overlap %s  gemm_ms %.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], (d.get("no_overlap") or {}).get("ms_per_step"), d["roofline"]["ms_per_step"]), d["e2e"].get("host_numa"))
except Exception as e:
    print("${name}: failed", e); print(open("gpurun_out/${TAG}_${name}.err").read()[-1500:])
P
}
run default SER_SM_RESERVE=0
run res8_ch8 SER_SM_RESERVE=8 NCCL_MAX_NCHANNELS=8
run res24 SER_SM_RESERVE=24
run padded SER_SM_RESERVE=0 BENCH_E2E_PADDED=1
