# usage: bash tools/gpu_profile.sh <tag>   (run under gpurun; writes gpurun_out/<tag>_*)
TAG=${1:-r01}
set -x
python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -2 gpurun_out/${TAG}_bench.err; cat gpurun_out/${TAG}_bench.json
SHORT="python bench.py --steps 2 --warmup 1 --prewarm 256 --envs 32768 --no-cpu"
$SHORT > gpurun_out/${TAG}_short.json 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 280 -c 40 --csv --log-file gpurun_out/${TAG}_launches.csv $SHORT > gpurun_out/${TAG}_ncu1.log 2>&1
$SHORT > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sf_step_kernel -s 262 -c 2 -o gpurun_out/${TAG}_step $SHORT > gpurun_out/${TAG}_ncu2.log 2>&1
tail -3 gpurun_out/${TAG}_ncu2.log
