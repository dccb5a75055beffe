# usage: bash tools/gpu_envs_sweep.sh -- step-kernel throughput against the number of arenas per GPU
for n in ${SWEEP:-65536 131072 262144 524288}; do
  python bench.py --steps 10 --warmup 3 --no-cpu --envs $n 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('envs $n', 'value %.3e ms/step %.3f e2e %.3e frac %.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac']))"
done
