#!/bin/bash
# ncu evidence of one build (one GPU): launch list of two eager steps + --set full of the top kernels; text only comes back
T=${1:-r02e}; O=gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-eager-arm --no-optimizer-arm"
$B > $O/plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 520 -c 250 --csv --log-file $O/${T}_ncu_launches_bench_cfg2.csv $B > $O/ncu_bench.log 2>&1
python tools/ncu_launches_summary.py $O/${T}_ncu_launches_bench_cfg2.csv --json $O/${T}_traffic.json > $O/${T}_ncu_launches_summary.txt 2>&1
A="python tools/attn_profile.py 128 1500 256 2 bwd 0.1 bits"
$A > $O/plain_attn.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:attn5 -s 9 -c 3 -o $O/attn5 $A > $O/ncu_attn.log 2>&1
ncu -i $O/attn5.ncu-rep --page details > $O/${T}_ncu_full_attn5_keepbits.txt 2>&1
G="python tools/gemm_bench.py --no-ref --iters 1 --cold --cases 1,0,2,11"
$G > $O/plain_gemm.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gemm_tc -c 16 -o $O/gemm $G > $O/ncu_gemm.log 2>&1
ncu -i $O/gemm.ncu-rep --page details > $O/${T}_ncu_full_gemm_cold.txt 2>&1
C="python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-eager-arm --no-optimizer-arm"
ncu --set full --clock-control none --import-source on -k regex:clf_stack -s 6 -c 2 -o $O/clf $C > $O/ncu_clf.log 2>&1
ncu -i $O/clf.ncu-rep --page details > $O/${T}_ncu_full_clf_stack.txt 2>&1
rm -f $O/*.ncu-rep
ls -la $O | tail -12; head -12 $O/${T}_ncu_launches_summary.txt
