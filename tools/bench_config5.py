#!/usr/bin/env python
"""BASELINE.json configs[4] on one GPU's shard: Battle Royale, 16 players in 4 teams per arena,
32,768 arenas (262,144 over 8 GPUs), observations of all 16 players built on the device and fed to
one batched AgentModel forward per tick (random-init weights of the reference's architecture,
bots/bot-0.5/Modules.hpp), commands back into the step -- nothing leaves the device.
Not the headline (bench.py measures that); prints one JSON line with three rows:
step only, step + observations, step + observations + policy forward.
The map is the reference's 3x30x100 (its dimensions are compile-time constants of the reference)."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from strikeforce_b200 import bots, policy  # noqa: E402
from strikeforce_b200 import config as sfcfg  # noqa: E402
from strikeforce_b200.sim import BatchedArena  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=32768)
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--prewarm", type=int, default=512)
ap.add_argument("--policy-chunk", type=int, default=32768, help="observations per forward call")
args = ap.parse_args()
teams = [1, 2, 3, 4] * 4
P = len(teams)
E = args.envs
sim = BatchedArena(E, mode="Royale", teams=teams, auto_reset=True, max_steps=2048)
t = 0
for _ in range(args.prewarm):
    sim.step(sim.synth_actions(t, sfcfg.ACTIONS28))
    t += 1
torch.cuda.synchronize()
pop = sim.population().float().mean(0).tolist()


def timed(fn, n):
    for _ in range(args.warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


acts = [sim.synth_actions(t + i, sfcfg.ACTIONS28, out=torch.empty((E, P), dtype=torch.uint8, device=sim.device))
        for i in range(args.steps + args.warmup)]
it = iter(range(10 ** 9))
ms_step = timed(lambda: sim.step(acts[next(it) % len(acts)]), args.steps)
obs = torch.empty((E, P, sfcfg.OBS_CH, sfcfg.OBS_WIN, sfcfg.OBS_WIN), dtype=torch.float32, device=sim.device)
mask = (1 << P) - 1


def step_obs():
    sim.observe(mask, out=obs)
    sim.step(acts[next(it) % len(acts)])


ms_obs = timed(step_obs, max(2, args.steps // 2))
model = policy.AgentModel().to(sim.device)
agent = policy.PolicyAgent(model, E * P, device=sim.device, seed=1, chunk=args.policy_chunk)
custom = bots.Custom(agent).prepare(sim)
actions = torch.empty((E, P), dtype=torch.uint8, device=sim.device)


def tick():
    sim.observe(mask, out=obs)
    with torch.no_grad():
        idx = agent.predict(obs.view(E * P, sfcfg.OBS_CH, sfcfg.OBS_WIN, sfcfg.OBS_WIN))
    actions.copy_(custom._table[idx].view(E, P))
    sim.step(actions)


ms_tick = timed(tick, max(2, args.steps // 4))
print(json.dumps({
    "workload": "royale16 (BASELINE.json configs[4] on the reference's 3x30x100 map)", "envs_per_gpu": E, "players": P,
    "mean_population": dict(zip(["humans", "zombies", "bullets", "chests", "built", "portals"], [round(x, 2) for x in pop])),
    "step_only": {"ms_per_step": ms_step, "env_steps_per_s": E / (ms_step / 1e3)},
    "step_plus_16_observations": {"ms_per_step": ms_obs, "env_steps_per_s": E / (ms_obs / 1e3),
                                  "observation_GBps": E * P * sfcfg.OBS_LEN * 4 / ((ms_obs - ms_step) / 1e3) / 1e9},
    "step_obs_policy_forward": {"ms_per_tick": ms_tick, "env_steps_per_s": E / (ms_tick / 1e3),
                                "agent_decisions_per_s": E * P / (ms_tick / 1e3)},
    "device_memory_GB": torch.cuda.max_memory_allocated() / 1e9}))
sim.close()
