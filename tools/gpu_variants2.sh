run() { # label, lib, env...
  label=$1; lib=$2; shift 2
  env "$@" SF_LIB_PATH=$PWD/$lib python bench.py --steps 10 --warmup 3 --no-cpu --prewarm ${PREWARM:-2048} 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$label', 'value %.3e ms/step %.3f e2e %.3e' % (d['value'], d['ms_per_step'], d['e2e']['value']))"
}
run "cta1024 gran32" strikeforce_b200/libstrikeforce_b200.so SF_L2_FETCH=32
run "cta1024 gran64" strikeforce_b200/libstrikeforce_b200.so SF_L2_FETCH=64
run "cta1024 gran128" strikeforce_b200/libstrikeforce_b200.so SF_L2_FETCH=128
run "cta896 gran32" build_variants/lib_cta896.so SF_L2_FETCH=32
run "cta896 gran64" build_variants/lib_cta896.so SF_L2_FETCH=64
