# usage: bash tools/gpu_iter.sh <tag>  -- parity tests + steady-state bench (no CPU arm) + ncu of one step
TAG=${1:-it}
set -x
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -2 gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
print("value %.3e e2e %.3e ms/step %.3f frac %.4f algoB %.0f pop %s draws %.1f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["algo_bytes_per_env_step"], d["config"]["mean_population"], d["rng_draws_per_env_step"]))
PY
bash tools/gpu_ncu_step.sh ${TAG} 65536 1024
