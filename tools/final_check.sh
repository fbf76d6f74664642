#!/bin/bash
# Round-end check in one GPU-box call: the full GPU test suite, smoke, the default bench line, the reference arm and the
# ncu launch list of one step (time + DRAM bytes).  Outputs in gpurun_out/final_<tag>/.
tag=${1:-r2}
out=gpurun_out/final_$tag
mkdir -p $out
python -m pytest tests/ -x -q -m gpu > $out/pytest_gpu.txt 2>&1
python __graft_entry__.py > $out/smoke.txt 2>&1
python bench.py > $out/bench_n1.json 2> $out/bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_reference_cpu.json 2> $out/bench_reference_cpu.err
python tools/profile_step.py --gemm --top 40 > $out/profile_step.txt 2>&1
python tools/one_step.py > $out/one_step_plain.log 2>&1 &&
ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    --csv --log-file $out/step_launches.csv python tools/one_step.py > $out/one_step_ncu.log 2>&1
python tools/summarize_launches.py $out/step_launches.csv $out/step_traffic.json > $out/step_launch_summary.csv 2>&1
gzip -f $out/step_launches.csv
