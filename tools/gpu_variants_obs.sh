# usage: bash tools/gpu_variants_obs.sh -- time the observation kernel of every build_variants/*.so (PREWARM default 2048 = bench.py's)
for f in strikeforce_b200/libstrikeforce_b200.so build_variants/*.so; do
  SF_LIB_PATH=$PWD/$f python bench.py --steps 6 --warmup 3 --no-cpu --prewarm ${PREWARM:-2048} 2>/dev/null | tail -1 > gpurun_out/x.json; python -c "
import json; d=json.load(open('gpurun_out/x.json')); w=d['with_observation']; print('$f', 'step %.3f ms; observe %.3f ms frac %.3f' % (d['ms_per_step'], w['observe_kernel_ms'], w['roofline']['frac']))"
done
