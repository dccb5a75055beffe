# usage: bash tools/gpu_final.sh <tag>  -- the round's evidence: tests, bench (both arms), launch list, ncu of both kernels
TAG=${1:-r01}
set -x
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
python __graft_entry__.py smoke 2>&1 | tail -1
python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -2 gpurun_out/${TAG}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err
SHORT="python bench.py --steps 2 --warmup 1 --prewarm 1536 --no-cpu"
$SHORT > gpurun_out/${TAG}_short.json 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 1560 -c 40 --csv --log-file gpurun_out/${TAG}_launches.csv $SHORT > gpurun_out/${TAG}_ncu1.log 2>&1
$SHORT > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sf_step_kernel -s 1537 -c 1 -o gpurun_out/${TAG}_step $SHORT > gpurun_out/${TAG}_ncu2.log 2>&1
$SHORT > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sf_observe_kernel -s 1 -c 1 -o gpurun_out/${TAG}_obs $SHORT > gpurun_out/${TAG}_ncu3.log 2>&1
tail -2 gpurun_out/${TAG}_ncu3.log
