#!/bin/bash
# Trimmed one-GPU evidence pass of the final build: tools/evidence_final.sh <tag>
T=${1:-r02g}; O=gpurun_out
timeout 300 python bench.py --steps 30 --warmup 5 --profile-detail $O/${T}_event_profile_cfg2.txt > $O/${T}_bench_cfg2.json 2> $O/${T}_bench_cfg2.err
timeout 120 python bench.py --workload cfg3 --steps 20 --warmup 5 --no-cpu-baseline --no-eager-arm > $O/${T}_bench_cfg3.json 2> $O/${T}_bench_cfg3.err
timeout 150 python bench.py --workload cfg4 --steps 10 --warmup 3 --no-cpu-baseline --no-eager-arm --profile-detail $O/${T}_event_profile_cfg4.txt > $O/${T}_bench_cfg4.json 2> $O/${T}_bench_cfg4.err
timeout 150 python bench.py --workload cfg5 --steps 3 --warmup 1 --no-cpu-baseline --no-eager-arm --profile-detail $O/${T}_event_profile_cfg5.txt > $O/${T}_bench_cfg5.json 2> $O/${T}_bench_cfg5.err
timeout 60 python tools/graph_timeline.py --workload cfg2 --out $O/${T}_graph_timeline_cfg2_cupti.txt > /dev/null 2>&1
timeout 60 python tools/graph_timeline.py --workload cfg4 --out $O/${T}_graph_timeline_cfg4_cupti.txt > /dev/null 2>&1
timeout 150 python tools/gpu_parity_report.py all > $O/${T}_parity_report.txt 2>&1
B="python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-eager-arm --no-optimizer-arm"
timeout 120 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 520 -c 250 --csv --log-file $O/${T}_ncu_launches_bench_cfg2.csv $B > $O/ncu_bench.log 2>&1
python tools/ncu_launches_summary.py $O/${T}_ncu_launches_bench_cfg2.csv --json $O/${T}_traffic.json > $O/${T}_ncu_launches_summary.txt 2>&1
for f in cfg2 cfg3 cfg4 cfg5; do python - <<P
import json
try:
    d=json.loads(open("$O/${T}_bench_$f.json").read().strip().splitlines()[-1])
    print("$f", round(d["value"]), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"],3), d["roofline"].get("frac_net_of_event_floor"))
except Exception as e: print("$f failed", e)
P
done
tail -3 $O/${T}_parity_report.txt; head -8 $O/${T}_ncu_launches_summary.txt
