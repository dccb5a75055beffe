#!/bin/bash
# usage: bash tools/collect_profiles.sh <tag>  -- turn gpurun_out/<tag>_* (tools/gpu_final.sh) into the tracked files under profiles/
set -e
TAG=${1:-r01}
cd "$(dirname "$0")/.."
cp gpurun_out/${TAG}_bench.json profiles/${TAG}_bench.json
cp gpurun_out/${TAG}_bench_reference.json profiles/${TAG}_bench_reference.json
cp gpurun_out/${TAG}_launches.csv profiles/${TAG}_launches.csv
ncu -i gpurun_out/${TAG}_step.ncu-rep --page source --print-source cuda,sass --csv 2>/dev/null > /tmp/${TAG}_src_step.csv
ncu -i gpurun_out/${TAG}_obs.ncu-rep --page source --print-source cuda,sass --csv 2>/dev/null > /tmp/${TAG}_src_obs.csv
{
  echo "# ncu --set full, sf_step_kernel<0>, 131072 arenas, steady-state populations (bench.py --steps 2 --warmup 1 --prewarm 1536 --no-cpu)"
  python tools/ncu_summary.py gpurun_out/${TAG}_step.ncu-rep
  echo; echo "# per function (samples, warp instructions, average active lanes)"
  python tools/ncu_funcs.py /tmp/${TAG}_src_step.csv strikeforce_b200/csrc/sf_core.cuh 2>/dev/null
  echo; echo "# hottest source lines"
  python tools/ncu_lines.py /tmp/${TAG}_src_step.csv 30
} > profiles/${TAG}_step_kernel_ncu.txt
{
  echo "# ncu --set full, sf_observe_kernel, 131072 observations"
  python tools/ncu_summary.py gpurun_out/${TAG}_obs.ncu-rep
  echo; python tools/ncu_lines.py /tmp/${TAG}_src_obs.csv 15
} > profiles/${TAG}_observe_kernel_ncu.txt
ncu -i gpurun_out/${TAG}_step.ncu-rep --page raw --csv 2>/dev/null > profiles/${TAG}_step_kernel_raw.csv
ncu -i gpurun_out/${TAG}_obs.ncu-rep --page raw --csv 2>/dev/null > profiles/${TAG}_observe_kernel_raw.csv
python - <<PY
import csv, json
def dram(path):
    rows = list(csv.reader(open(path)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    def b(k):
        i = hdr.index(k)
        return float(vals[i]) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[i]]
    return b("dram__bytes_read.sum") + b("dram__bytes_write.sum")
json.dump({"note": "dram__bytes_read.sum + dram__bytes_write.sum per launch from one ncu --set full capture (tools/gpu_final.sh); bench.py copies them into roofline.traffic when it runs the same arena count",
           "envs_per_gpu": 131072,
           "sf_step_kernel": {"dram_bytes_per_launch": dram("profiles/${TAG}_step_kernel_raw.csv"), "source": "profiles/${TAG}_step_kernel_raw.csv"},
           "sf_observe_kernel": {"dram_bytes_per_launch": dram("profiles/${TAG}_observe_kernel_raw.csv"), "source": "profiles/${TAG}_observe_kernel_raw.csv"}},
          open("profiles/traffic.json", "w"), indent=1)
PY
echo collected
