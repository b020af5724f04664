#!/bin/bash
# One-GPU evidence pass for profiles/: tools/evidence_run.sh <tag>
T=${1:-r02e}; O=gpurun_out
python tools/attn_check.py all --time > $O/${T}_attention_tcgen05_vs_mma_sync.txt 2>&1
python tools/gemm_bench.py --cold --iters 16 > $O/${T}_gemm_bench_cold.txt 2>&1
python tools/gemm_bench.py > $O/${T}_gemm_bench_hot.txt 2>&1
python tools/gpu_parity_report.py all > $O/${T}_parity_report.txt 2>&1
python bench.py --steps 30 --warmup 5 --profile-detail $O/${T}_event_profile_cfg2.txt > $O/${T}_bench_cfg2.json 2> $O/${T}_bench_cfg2.err
python bench.py --workload cfg3 --steps 20 --warmup 5 --no-cpu-baseline --no-eager-arm > $O/${T}_bench_cfg3.json 2> $O/${T}_bench_cfg3.err
python bench.py --workload cfg4 --steps 10 --warmup 3 --no-cpu-baseline --profile-detail $O/${T}_event_profile_cfg4.txt > $O/${T}_bench_cfg4.json 2> $O/${T}_bench_cfg4.err
python bench.py --workload cfg5 --steps 3 --warmup 1 --profile-detail $O/${T}_event_profile_cfg5.txt > $O/${T}_bench_cfg5.json 2> $O/${T}_bench_cfg5.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/${T}_bench_cfg2_reference_arm.json 2> $O/${T}_ref.err
python tools/graph_timeline.py --workload cfg2 --out $O/${T}_graph_timeline_cfg2_cupti.txt > /dev/null 2>&1
python tools/graph_timeline.py --workload cfg4 --out $O/${T}_graph_timeline_cfg4_cupti.txt > /dev/null 2>&1
SER_PDL=0 python tools/pooling_cupti.py > $O/${T}_pooling_cupti.txt 2>&1
for f in cfg2 cfg3 cfg4 cfg5; do python - <<P
import json
try:
    d=json.loads(open("$O/${T}_bench_$f.json").read().strip().splitlines()[-1])
    print("$f", round(d["value"]), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"],3), d["roofline"].get("frac_net_of_event_floor"))
except Exception as e: print("$f failed", e)
P
done
tail -3 $O/${T}_parity_report.txt
