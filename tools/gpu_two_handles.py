#!/usr/bin/env python
"""Why does a batch larger than one wave step super-linearly slower (profiles/r01_envs_sweep.txt)?  The same
262,144 arenas as ONE handle (rows of 262,144 elements: 1 MB between two slots of an array) and as TWO
handles of 131,072 (rows half as long, the same total memory), stepped back to back."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from strikeforce_b200 import config as sfcfg  # noqa: E402
from strikeforce_b200.sim import BatchedArena  # noqa: E402


def run(sizes, prewarm=768, steps=30):
    sims = []
    base = 0
    for n in sizes:
        sims.append(BatchedArena(n, mode="Squad", level=1, level_max=10, auto_reset=True, max_steps=2048, env_id_base=base))
        base += n
    t = 0
    for _ in range(prewarm):
        for s in sims:
            s.step(s.synth_actions(t, sfcfg.ACTIONS28))
        t += 1
    acts = [[s.synth_actions(t + i, sfcfg.ACTIONS28, out=torch.empty_like(s._synth)) for i in range(steps)] for s in sims]
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        for s, a in zip(sims, acts):
            s.step(a[i])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print("handles %s: %.3f ms per step of %d arenas -> %.3e env-steps/s" % (sizes, ms, sum(sizes), sum(sizes) / (ms / 1e3)), flush=True)
    for s in sims:
        s.close()


for sizes in ([131072], [262144], [131072, 131072], [524288], [131072] * 4):
    run(sizes)
