set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
python __graft_entry__.py smoke 2>&1 | tail -5
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
timeout 300 python tools/quick_bench.py --envs 4096 --mode Solo --prewarm 200 --steps 20 2>&1 | tail -8
timeout 600 python tools/quick_bench.py --envs 131072 --mode Squad --prewarm 256 --steps 20 2>&1 | tail -8
