#!/bin/bash
# Everything profiles/ is built from, in one GPU-box call (1 GPU).  Usage: bash tools/collect_evidence.sh <tag>
# Outputs land in gpurun_out/ev_<tag>/ ; copy the summaries into profiles/ afterwards (tools/ncu_summary.py for .ncu-rep).
tag=${1:-r2}
out=gpurun_out/ev_$tag
mkdir -p $out
python bench.py > $out/bench_n1.json 2> $out/bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_reference_cpu.json 2> $out/bench_reference_cpu.err
python bench.py --impl reference --ref-device cuda --steps 3 > $out/stock_torch_cuda.json 2> $out/stock_torch_cuda.err
for w in resnet_dgrn_fwd vit_dgrn_train infer512 infer1024 infer512_dgrn infer1024_dgrn; do
  python bench.py --workload $w --steps 10 > $out/bench_$w.json 2> $out/bench_$w.err
done
for k in gemm epi attn misc dgrn; do python tools/bench_kernels.py $k > $out/bench_kernels_$k.txt 2>&1; done
python tools/bench_kernels.py gemm --backend 5 --bexact --only leff > $out/bench_kernels_gemm_1xtf32.txt 2>&1
python tools/profile_step.py --gemm --top 60 > $out/profile_step.txt 2>&1
# ncu: launch list of one step with DRAM bytes (shares + traffic), then one --set full capture per kernel class
python tools/one_step.py > $out/one_step_plain.log 2>&1 &&
ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    --csv --log-file $out/step_launches.csv python tools/one_step.py > $out/one_step_ncu.log 2>&1
python tools/summarize_launches.py $out/step_launches.csv $out/step_traffic.json > $out/step_launch_summary.csv 2>&1
gzip -f $out/step_launches.csv
cap() {   # cap <name> <kernel regex> <skip> <count> <cmd...>
  name=$1; rx=$2; sk=$3; cn=$4; shift 4
  "$@" > $out/plain_$name.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$rx -s $sk -c $cn -o $out/$name "$@" > $out/ncu_$name.log 2>&1
  # gpurun brings back at most 64 MiB: keep the text summary and the per-instruction page, drop the report itself
  python tools/ncu_summary.py $out/$name.ncu-rep > $out/ncu_summary_$name.txt 2>&1
  ncu -i $out/$name.ncu-rep --page source --csv 2>/dev/null | gzip > $out/ncu_source_$name.csv.gz
  rm -f $out/$name.ncu-rep
}
cap gemm_compute_3x gemm_tc 2 1 python tools/one_gemm.py 4096 3584 896 0 1 0
cap gemm_compute_1x gemm_tc 2 1 python tools/one_gemm.py 4096 3584 896 0 1 5 1
cap gemm_hbm_plain gemm_tc 2 1 python tools/one_gemm.py 262144 448 112 0 1 0 0 plain
cap gemm_hbm_gelu gemm_tc 2 1 python tools/one_gemm.py 262144 448 112 0 1 0 0 gelu
cap attn win_attn 2 2 python tools/one_attn.py
cap kernels_dgrn "dcn_|gemm_tc" 4 4 python tools/bench_kernels.py dgrn
cap kernels_misc "dwconv|layernorm" 30 4 python tools/bench_kernels.py misc
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $out/gpu.csv
