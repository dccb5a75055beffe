# usage: bash tools/gpu_subwave.sh  -- batches smaller than one wave: arenas per warp chosen from the batch
# size (sf_launch_step) against the full-batch mapping (32 arenas per warp)
for mode in Solo Squad; do
for envs in 4096 32768 65536; do
  for lpw in 32 auto; do
    if [ $lpw = auto ]; then unset SF_LANES_PER_WARP; else export SF_LANES_PER_WARP=$lpw; fi
    python tools/quick_bench.py --mode $mode --envs $envs --prewarm 768 --steps 40 2>/dev/null | grep "timed" | sed "s/^/$mode envs $envs lanes-per-warp $lpw: /" | cut -c1-150
  done
done
done
