"""Randomised parity soak on the GPU: many steps per arena with auto-reset in every mode, alphabets
that include '_' (quit), '3' and bytes outside valid_commands, small capacities that provoke
SF_OVERFLOW, Battle Royale with 16 players.  Status and canonical-state hash against the C oracle
after EVERY step.  `python tools/gpu_soak.py` runs the same cases several times longer."""
import numpy as np
import pytest

import common
import sfo
from strikeforce_b200 import config as sfcfg

pytestmark = pytest.mark.gpu

FULL = sfcfg.ACTIONS28 + b"_3" + b"12p~ \x00\xff"
ROYALE_CAPS = dict(cap_portals=128, cap_built=1000, cap_bullets=128)
TEAMS = [1, 2, 3, 4] * 4
# name: mode, level range, arenas, steps (test / tool), alphabet, squad agents, sheet, capacities, max_steps
CASES = {
    "solo-junk-bytes": (sfcfg.MODE_SOLO, (1, 3), 48, (700, 3000), FULL, False, "account1", None, 500),
    "timer": (sfcfg.MODE_TIMER, (1, 2), 32, (600, 2500), sfcfg.ACTIONS28, False, "synthetic", None, 0),
    "squad-junk-bytes": (sfcfg.MODE_SQUAD, (1, 10), 48, (900, 3000), FULL, False, "account1", None, 700),
    "squad-agents": (sfcfg.MODE_SQUAD, (2, 4), 32, (500, 2000), sfcfg.ACTIONS9, True, "account1", None, 400),
    "solo-small-capacities": (sfcfg.MODE_SOLO, (1, 1), 48, (900, 2500), sfcfg.ACTIONS28, False, "account1",
                              dict(cap_humans=16, cap_zombies=24, cap_bullets=12, cap_built=24, cap_portals=12), 0),
    "royale-junk-bytes": (sfcfg.MODE_ROYALE, (1, 1), 40, (700, 2500), FULL, False, "account1", ROYALE_CAPS, 500),
    "royale-new-player": (sfcfg.MODE_ROYALE, (1, 1), 40, (800, 4000), sfcfg.ACTIONS28, False, "new_player", ROYALE_CAPS, 0),
}


def run_case(torch, arena, name, long=False, base=777):
    from strikeforce_b200.sim import BatchedArena
    mode, (l0, l1), n, steps, table, agents, player, caps, max_steps = CASES[name]
    steps = steps[1 if long else 0]
    teams = TEAMS if mode == sfcfg.MODE_ROYALE else None
    sim = BatchedArena(n, mode=mode, level=l0, level_max=l1, squad_agents=agents, auto_reset=True, max_steps=max_steps,
                       env_id_base=base, player=player, caps=caps, teams=teams)
    try:
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
        episode, ends = [0] * n, {}
        for t in range(steps):
            act = common.synth_actions(range(base, base + n), sim.n_agents, t, table)
            sim.step(torch.from_numpy(act).to(sim.device))
            out = sim.step_out().cpu().numpy()
            h = sim.state_hash().cpu().numpy().view(np.uint64)
            for e, o in enumerate(oracles):
                st = o.step(bytes(act[e]))
                assert out[e, 0] == st, "status differs: %s arena %d step %d: %d vs %d" % (name, e, t, out[e, 0], st)
                if st != 0:
                    ends[st] = ends.get(st, 0) + 1
                    episode[e] += 1
                    o.reset(levels[e], common.synth_tb(base + e), common.synth_serial(base + e, episode[e]))
                if h[e] != np.uint64(o.state_hash()):
                    raise AssertionError("state differs: %s arena %d step %d\n%s" % (
                        name, e, t, "\n".join(sfo.diff_records(o.dump(), sim.export_env(e), 20))))
        pop = [round(x, 1) for x in sim.population().float().mean(0).tolist()]
        return dict(case=name, arenas=n, steps=steps, episodes=sum(episode), ends=ends, mean_population=pop)
    finally:
        sim.close()


@pytest.mark.parametrize("name", sorted(CASES))
def test_soak(name, arena_data):
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device; there is no CPU fallback to test"
    r = run_case(torch, arena_data, name)
    assert r["episodes"] > 0 or CASES[name][8] == 0
