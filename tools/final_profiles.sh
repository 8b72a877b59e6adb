#!/bin/bash
# The set of measurements committed under profiles/ for one state of the tree (run under gpurun, 1 GPU):
#   bash tools/final_profiles.sh <tag>
tag=$1
python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/${tag}_gpu_tests.log
python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference_arm.json 2>> gpurun_out/${tag}_bench.err
python bench.py --beams 360 --steps 20 --warmup 5 --no-cpu > gpurun_out/${tag}_bench_chain360.json 2>> gpurun_out/${tag}_bench.err
cmd="python bench.py --no-e2e --no-cpu --no-sustained --steps 2 --warmup 1"
$cmd > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches.csv $cmd > gpurun_out/ncu0.log 2>&1
$cmd > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:icp_align_kernel -s 1 -c 1 -f -o gpurun_out/${tag}_align $cmd > gpurun_out/ncu1.log 2>&1
$cmd > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:icp_align_kernel -s 4 -c 2 -f -o gpurun_out/${tag}_exh $cmd > gpurun_out/ncu2.log 2>&1
cmd2="python tools/slam_pipeline.py --sgd-steps 2"
$cmd2 > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"sgd_chain|proximity_closest" -c 3 -f -o gpurun_out/${tag}_sgd_prox $cmd2 > gpurun_out/ncu3.log 2>&1
python tools/sgd_bench.py --profile > gpurun_out/${tag}_sgd_bench.json 2> gpurun_out/${tag}_sgd_kernels.txt
python tools/slam_pipeline.py > gpurun_out/${tag}_slam_pipeline.json 2>> gpurun_out/${tag}_bench.err
tail -2 gpurun_out/${tag}_gpu_tests.log; tail -c 400 gpurun_out/${tag}_bench.json
