# usage: bash tools/gpu_subwave.sh  -- batches smaller than one wave: which number of arenas per warp (SF_LANES_PER_WARP) is best?
for envs in ${SUBWAVE_ENVS:-4096 16384 32768 65536}; do
  for lpw in ${SUBWAVE_LPW:-1 2 4 8 16 32}; do
    SF_LANES_PER_WARP=$lpw python tools/quick_bench.py --mode Squad --envs $envs --prewarm ${PREWARM:-1536} --steps 30 2>/dev/null | grep "timed" | sed "s/^/Squad envs $envs arenas-per-warp $lpw: /" | cut -c1-120
  done
done
