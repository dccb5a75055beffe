#!/usr/bin/env python
"""Quick device timing of the step kernel (development aid; bench.py is the contract)."""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from strikeforce_b200 import config as sfcfg  # noqa: E402
from strikeforce_b200.sim import BatchedArena  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=131072)
ap.add_argument("--mode", default="Squad")
ap.add_argument("--prewarm", type=int, default=512)
ap.add_argument("--steps", type=int, default=50)
ap.add_argument("--max-steps", type=int, default=2048)
ap.add_argument("--table", default="28")
ap.add_argument("--obs", type=int, default=0)
args = ap.parse_args()
table = sfcfg.ACTIONS28 if args.table == "28" else sfcfg.ACTIONS9
t0 = time.time()
sim = BatchedArena(args.envs, mode=args.mode, level=1, auto_reset=True, max_steps=args.max_steps)
torch.cuda.synchronize()
print("create+reset %.2fs, device bytes %.2f GB" % (time.time() - t0, sim.device_bytes / 1e9))
t = 0
t0 = time.time()
for _ in range(args.prewarm):
    sim.step(sim.synth_actions(t, table))
    t += 1
torch.cuda.synchronize()
print("prewarm %d steps: %.2fs -> %.3e steps/s" % (args.prewarm, time.time() - t0, args.prewarm * args.envs / max(time.time() - t0, 1e-9)))
print("population mean:", sim.population().float().mean(0).tolist())
st0 = sim.stats()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
acts = []
for i in range(args.steps):
    acts.append(sim.synth_actions(t + i, table, out=torch.empty_like(sim._synth)))
torch.cuda.synchronize()
obs = None
ev[0].record()
for i in range(args.steps):
    sim.step(acts[i])
    if args.obs:
        obs = sim.observe(1, out=obs)
    ev[i + 1].record()
torch.cuda.synchronize()
ms = ev[0].elapsed_time(ev[-1])
st1 = sim.stats()
algo = st1["algo_bytes"] - st0["algo_bytes"]
print("timed %d steps: %.3f ms/step -> %.3e env-steps/s; algo bytes/step %.1f KB/env; achieved %.1f GB/s; draws/step %.1f"
      % (args.steps, ms / args.steps, args.steps * args.envs / (ms / 1e3), algo / args.steps / args.envs / 1e3,
         (algo + (args.obs * args.steps * args.envs * 123008)) / (ms / 1e3) / 1e9,
         (st1["rng_draws"] - st0["rng_draws"]) / args.steps / args.envs))
print("stats:", {k: st1[k] - st0[k] for k in st1})
