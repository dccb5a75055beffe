# usage: bash tools/gpu_variants.sh  -- time every build_variants/*.so on the steady-state bench
# (BENCH_ARGS="--no-obs" skips the step+observation row; PREWARM = untimed age-spreading steps)
for f in strikeforce_b200/libstrikeforce_b200.so build_variants/*.so; do
  SF_LIB_PATH=$PWD/$f python bench.py --steps 10 --warmup 3 --no-cpu --prewarm ${PREWARM:-1024} ${BENCH_ARGS:-} 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
o=d.get('with_observation')
c=(o or {}).get('channels_last')
print('$f', 'value %.3e ms/step %.3f e2e %.3e' % (d['value'], d['ms_per_step'], d['e2e']['value']), ('obs %.3f ms = %.0f GB/s' % (o['observe_kernel_ms'], o['roofline']['achieved'])) if o else '', ('nhwc %.3f ms = %.0f GB/s' % (c['observe_kernel_ms'], c['roofline']['achieved'])) if c else '')"
done
