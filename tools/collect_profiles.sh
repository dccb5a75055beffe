#!/bin/bash
# usage: bash tools/collect_profiles.sh <tag>  -- turn gpurun_out/<tag>_* (tools/gpu_final.sh) into the tracked files under profiles/
set -e
TAG=${1:-r02}
cd "$(dirname "$0")/.."
cp gpurun_out/${TAG}_bench.json profiles/${TAG}_bench.json
cp gpurun_out/${TAG}_bench_reference.json profiles/${TAG}_bench_reference.json
cp gpurun_out/${TAG}_launches.csv profiles/${TAG}_launches.csv
cp gpurun_out/${TAG}_tests.log profiles/${TAG}_gpu_tests.log 2>/dev/null || true
ncu -i gpurun_out/${TAG}_step.ncu-rep --page source --print-source cuda,sass --csv 2>/dev/null > /tmp/${TAG}_src_step.csv
ncu -i gpurun_out/${TAG}_obs.ncu-rep --page source --print-source cuda,sass --csv 2>/dev/null > /tmp/${TAG}_src_obs.csv
ncu -i gpurun_out/${TAG}_obs_nhwc.ncu-rep --page source --print-source cuda,sass --csv 2>/dev/null > /tmp/${TAG}_src_obs_nhwc.csv
cp gpurun_out/${TAG}_royale16.json profiles/${TAG}_royale16.json 2>/dev/null || true
{
  echo "# ncu --set full of ONE sf_step_kernel<0> launch inside the timed region of the bench command itself"
  echo "# (tools/gpu_final.sh: python bench.py --steps 20 --warmup 3 --no-cpu, NVTX range sf_timed_device): 131,072 arenas,"
  echo "# the populations the bench times"
  python tools/ncu_summary.py gpurun_out/${TAG}_step.ncu-rep
  echo; echo "# stall reasons (share of the warp samples)"
  python tools/ncu_stalls.py gpurun_out/${TAG}_step.ncu-rep
  echo; echo "# per function (samples, warp instructions, average active lanes)"
  python tools/ncu_funcs.py /tmp/${TAG}_src_step.csv strikeforce_b200/csrc/sf_core.cuh 2>/dev/null
  echo; echo "# hottest source lines"
  python tools/ncu_lines.py /tmp/${TAG}_src_step.csv 30
} > profiles/${TAG}_step_kernel_ncu.txt
{
  echo "# ncu --set full of ONE sf_observe_kernel launch inside the timed region of the bench command (NVTX range sf_timed_observe), 131,072 observations"
  python tools/ncu_summary.py gpurun_out/${TAG}_obs.ncu-rep
  echo; echo "# stall reasons (share of the warp samples)"
  python tools/ncu_stalls.py gpurun_out/${TAG}_obs.ncu-rep
  echo; python tools/ncu_lines.py /tmp/${TAG}_src_obs.csv 15
} > profiles/${TAG}_observe_kernel_ncu.txt
{
  echo "# ncu --set full of ONE sf_observe_kernel<true> launch (channel-innermost layout, SF_OBS_NHWC) inside the timed region of the bench command (NVTX range sf_timed_observe_nhwc), 131,072 observations"
  python tools/ncu_summary.py gpurun_out/${TAG}_obs_nhwc.ncu-rep
  echo; echo "# stall reasons (share of the warp samples)"
  python tools/ncu_stalls.py gpurun_out/${TAG}_obs_nhwc.ncu-rep
  echo; python tools/ncu_lines.py /tmp/${TAG}_src_obs_nhwc.csv 15
} > profiles/${TAG}_observe_nhwc_kernel_ncu.txt
ncu -i gpurun_out/${TAG}_obs_nhwc.ncu-rep --page raw --csv 2>/dev/null > profiles/${TAG}_observe_nhwc_kernel_raw.csv
ncu -i gpurun_out/${TAG}_step.ncu-rep --page raw --csv 2>/dev/null > profiles/${TAG}_step_kernel_raw.csv
ncu -i gpurun_out/${TAG}_obs.ncu-rep --page raw --csv 2>/dev/null > profiles/${TAG}_observe_kernel_raw.csv
python tools/ncu_mem_lines.py /tmp/${TAG}_src_step.csv > profiles/${TAG}_step_kernel_mem_lines.txt 2>/dev/null || true
# SASS evidence: instruction mix per kernel of the shipped library (no tensor-core ops: there is no contraction on this path)
{
  echo "# tools/sass_summary.py strikeforce_b200/libstrikeforce_b200.so (cuobjdump -sass): instructions per kernel and what they are made of"
  python tools/sass_summary.py strikeforce_b200/libstrikeforce_b200.so
} > profiles/${TAG}_sass_summary.txt
python - <<PY
import csv, json
def dram(path):
    rows = list(csv.reader(open(path)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    def b(k):
        i = hdr.index(k)
        return float(vals[i]) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[i]]
    return b("dram__bytes_read.sum") + b("dram__bytes_write.sum")
bench = json.loads(open("profiles/${TAG}_bench.json").read().strip().splitlines()[-1])
algo = bench["roofline"]["algo_bytes_per_launch"]
step, obs = dram("profiles/${TAG}_step_kernel_raw.csv"), dram("profiles/${TAG}_observe_kernel_raw.csv")
obs_cl = dram("profiles/${TAG}_observe_nhwc_kernel_raw.csv")
json.dump({"note": "dram__bytes_read.sum + dram__bytes_write.sum of ONE launch from an ncu --set full capture taken inside the timed region of the "
                   "bench command itself (tools/gpu_final.sh), next to the algorithmic bytes of a launch of that same region; bench.py copies "
                   "dram_bytes_per_launch into roofline.traffic when it runs the same arena count",
           "envs_per_gpu": bench["config"]["envs_per_gpu"],
           "sf_step_kernel": {"dram_bytes_per_launch": step, "algo_bytes_per_launch": algo, "traffic_over_algorithmic": step / algo,
                              "source": "profiles/${TAG}_step_kernel_raw.csv"},
           "sf_observe_kernel": {"dram_bytes_per_launch": obs, "algo_bytes_per_launch": bench["config"]["envs_per_gpu"] * 123008,
                                 "traffic_over_algorithmic": obs / (bench["config"]["envs_per_gpu"] * 123008),
                                 "source": "profiles/${TAG}_observe_kernel_raw.csv"},
           "sf_observe_kernel_nhwc": {"dram_bytes_per_launch": obs_cl, "algo_bytes_per_launch": bench["config"]["envs_per_gpu"] * 123008,
                                      "traffic_over_algorithmic": obs_cl / (bench["config"]["envs_per_gpu"] * 123008),
                                      "source": "profiles/${TAG}_observe_nhwc_kernel_raw.csv"}},
          open("profiles/traffic.json", "w"), indent=1)
PY
echo collected
