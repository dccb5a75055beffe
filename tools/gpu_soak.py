#!/usr/bin/env python
"""Long randomised parity soak on the GPU (not part of the test-suite): several thousand steps per
arena with auto-reset, every mode, alphabets that include '_' (quit), '3' and bytes outside
valid_commands, small capacities to provoke SF_OVERFLOW.  State hash vs the C oracle every step."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import common  # noqa: E402
import sfo  # noqa: E402
from strikeforce_b200 import config as sfcfg  # noqa: E402
from strikeforce_b200 import data as sfdata  # noqa: E402
from strikeforce_b200.sim import BatchedArena  # noqa: E402

arena = sfdata.load_default()
FULL = sfcfg.ACTIONS28 + b"_3" + b"12p~ \x00\xff"
ROYALE_CAPS = dict(cap_portals=128, cap_built=1000, cap_bullets=128)
CASES = [  # mode, level range, envs, steps, table, agents, player, caps, max_steps (Battle Royale: 16 players, 4 teams)
    (sfcfg.MODE_SOLO, (1, 3), 48, 3000, FULL, False, "account1", None, 1500),
    (sfcfg.MODE_TIMER, (1, 2), 32, 2500, sfcfg.ACTIONS28, False, "synthetic", None, 0),
    (sfcfg.MODE_SQUAD, (1, 10), 48, 3000, FULL, False, "account1", None, 2048),
    (sfcfg.MODE_SQUAD, (2, 4), 32, 2000, sfcfg.ACTIONS9, True, "account1", None, 1000),
    (sfcfg.MODE_SOLO, (1, 1), 48, 2500, sfcfg.ACTIONS28, False, "account1",
     dict(cap_humans=16, cap_zombies=24, cap_bullets=12, cap_built=24, cap_portals=12), 0),
    (sfcfg.MODE_ROYALE, (1, 1), 40, 2500, FULL, False, "account1", ROYALE_CAPS, 1200),
    (sfcfg.MODE_ROYALE, (1, 1), 40, 4000, sfcfg.ACTIONS28, False, "new_player", ROYALE_CAPS, 0),
]
TEAMS = [1, 2, 3, 4] * 4
t00 = time.time()
for mode, (l0, l1), n, steps, table, agents, player, caps, max_steps in CASES:
    base = 777
    teams = TEAMS if mode == sfcfg.MODE_ROYALE else None
    sim = BatchedArena(n, mode=mode, level=l0, level_max=l1, squad_agents=agents, auto_reset=True, max_steps=max_steps,
                       env_id_base=base, player=player, caps=caps, teams=teams)
    span = l1 - l0 + 1
    oracles, levels = [], []
    for e in range(n):
        lvl = l0 + (base + e) % span
        cfg = sfcfg.make_config(arena, mode=mode, level_min=lvl, squad_agents=agents, max_steps=max_steps, player=player,
                                caps=caps, teams=teams)
        o = sfo.Arena(cfg)
        o.reset(lvl, common.synth_tb(base + e), common.synth_serial(base + e, 0))
        oracles.append(o)
        levels.append(lvl)
    episode = [0] * n
    ends = {}
    for t in range(steps):
        act = common.synth_actions(range(base, base + n), sim.n_agents, t, table)
        sim.step(torch.from_numpy(act).to(sim.device))
        out = sim.step_out().cpu().numpy()
        h = sim.state_hash().cpu().numpy().view(np.uint64)
        for e, o in enumerate(oracles):
            st = o.step(bytes(act[e]))
            assert out[e, 0] == st, ("status", mode, e, t, out[e, 0], st)
            if st != 0:
                ends[st] = ends.get(st, 0) + 1
                episode[e] += 1
                o.reset(levels[e], common.synth_tb(base + e), common.synth_serial(base + e, episode[e]))
            if h[e] != np.uint64(o.state_hash()):
                print("STATE DIFFERS", mode, e, t)
                print("\n".join(sfo.diff_records(o.dump(), sim.export_env(e), 20)))
                sys.exit(1)
    pop = sim.population().float().mean(0).tolist()
    print("soak ok: mode %d agents %d envs %d steps %d episodes %s ends %s mean pop %s (%.0f s)" % (
        mode, agents, n, steps, sum(episode), ends, [round(x, 1) for x in pop], time.time() - t00), flush=True)
    sim.close()
print("ALL SOAKS OK")
