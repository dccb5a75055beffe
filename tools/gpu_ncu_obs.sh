# usage: bash tools/gpu_ncu_obs.sh <tag>  -- ncu --set full of one observation kernel launch (16,384 arenas)
TAG=${1:-r01}
set -x
SHORT="python bench.py --steps 2 --warmup 1 --prewarm 512 --envs 16384 --no-cpu"
$SHORT > gpurun_out/${TAG}_obs_short.json 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sf_observe_kernel -s 1 -c 1 -o gpurun_out/${TAG}_obs $SHORT > gpurun_out/${TAG}_ncu_obs.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_obs.log
