# usage: bash tools/gpu_obs_run.sh -- the observation kernel against the number of CONSECUTIVE observations a CTA
# writes before it moves on (SF_OBS_RUN; 1 = grid-stride over single observations)
for r in 0 1 4 16; do
  SF_OBS_RUN=$r python bench.py --steps 10 --warmup 3 --no-cpu --prewarm ${PREWARM:-1024} 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
o=d['with_observation']; c=o['channels_last']
print('SF_OBS_RUN=$r', 'obs %.3f ms = %.0f GB/s (%.1f%%)' % (o['observe_kernel_ms'], o['roofline']['achieved'], 100*o['roofline']['frac']), 'nhwc %.3f ms = %.0f GB/s (%.1f%%)' % (c['observe_kernel_ms'], c['roofline']['achieved'], 100*c['roofline']['frac']), 'step+obs %.3f / %.3f ms' % (o['ms_per_step'], c['ms_per_step']))"
done
