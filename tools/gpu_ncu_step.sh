# usage: bash tools/gpu_ncu_step.sh <tag> [envs] [prewarm]  -- ncu --set full of one steady-state step launch
TAG=${1:-r01}
ENVS=${2:-65536}
PRE=${3:-1024}
set -x
SHORT="python bench.py --steps 2 --warmup 1 --prewarm $PRE --envs $ENVS --no-cpu --no-obs"
$SHORT > gpurun_out/${TAG}_short.json 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sf_step_kernel -s $((PRE+1)) -c 1 -o gpurun_out/${TAG}_step $SHORT > gpurun_out/${TAG}_ncu2.log 2>&1
tail -3 gpurun_out/${TAG}_ncu2.log
cat gpurun_out/${TAG}_short.json | tail -1 | cut -c1-400
