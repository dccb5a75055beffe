# usage: bash tools/gpu_final.sh <tag>  -- the round's evidence on one B200: tests, both bench arms, the launch
# list of the TIMED region (NVTX range of bench.py), ncu --set full of one timed launch of each kernel
TAG=${1:-r02}
set -x
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -2 | tee gpurun_out/${TAG}_tests.log
python __graft_entry__.py smoke 2>&1 | tail -1
python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -2 gpurun_out/${TAG}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err
# the same command as the bench (same set-up, same populations), profiled: never a bench value
SHORT="python bench.py --steps 20 --warmup 3 --no-cpu"
ncu --nvtx --nvtx-include "sf_timed_device/" --nvtx-include "sf_timed_observe/" --nvtx-include "sf_timed_observe_nhwc/" --metrics gpu__time_duration.sum --clock-control none \
    --csv --log-file gpurun_out/${TAG}_launches.csv $SHORT > gpurun_out/${TAG}_ncu1.log 2>&1
ncu --nvtx --nvtx-include "sf_timed_device/" --set full --clock-control none --import-source on -k regex:sf_step_kernel -s 10 -c 1 \
    -o gpurun_out/${TAG}_step $SHORT > gpurun_out/${TAG}_ncu2.log 2>&1
ncu --nvtx --nvtx-include "sf_timed_observe/" --set full --clock-control none --import-source on -k regex:sf_observe_kernel -s 2 -c 1 \
    -o gpurun_out/${TAG}_obs $SHORT > gpurun_out/${TAG}_ncu3.log 2>&1
ncu --nvtx --nvtx-include "sf_timed_observe_nhwc/" --set full --clock-control none --import-source on -k regex:sf_observe_kernel -s 2 -c 1 \
    -o gpurun_out/${TAG}_obs_nhwc $SHORT > gpurun_out/${TAG}_ncu4.log 2>&1
tail -n 2 gpurun_out/${TAG}_ncu3.log; tail -n 2 gpurun_out/${TAG}_ncu4.log
# BASELINE configs[4] on this GPU's shard, and where its policy forward spends its time
python bench.py --workload royale16 --steps 3 --warmup 3 > gpurun_out/${TAG}_royale16.json 2> gpurun_out/${TAG}_royale16.err
python tools/gpu_policy_forward.py > gpurun_out/${TAG}_policy_forward_now.txt 2>&1
