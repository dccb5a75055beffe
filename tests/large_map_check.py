"""Parity on a LARGER arena than the reference's (BASELINE.json configs[4], "large custom map"): run as a
script with SF_GEOMETRY=ROWSxCOLS in the environment (tests/test_large_map.py does), because the arena's
dimensions are compile-time constants of every library involved -- as they are in the reference
(gameplay.hpp:37), which therefore cannot play this map: PARITY HERE IS PINNED AGAINST THE C ORACLE ONLY
(oracle/sf_oracle.c built with the same -DSF_ROWS / -DSF_COLS; at 30x100 that oracle is pinned against the
unmodified reference).

    SF_GEOMETRY=40x128 python tests/large_map_check.py host    # the device tick compiled for the host
    SF_GEOMETRY=40x128 python tests/large_map_check.py gpu     # the CUDA library through the C ABI

Squad with the whole alphabet (blocks, portals, bullets crossing the old border through its doors) and
Battle Royale with 16 players placed anywhere on the large map; status and canonical-state hash after every
step, observations of every player at intervals."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, os.path.join(ROOT, "oracle"), HERE, os.path.join(HERE, "hostcheck")):
    sys.path.insert(0, p)
import common  # noqa: E402
import sfo  # noqa: E402
from strikeforce_b200 import config as sfcfg  # noqa: E402
from strikeforce_b200 import data as sfdata  # noqa: E402

TEAMS = [1, 2, 3, 4] * 4
CASES = [  # name, mode, arenas, steps, agents per arena, teams
    ("squad", sfcfg.MODE_SQUAD, 10, 700, 1, None),
    ("royale", sfcfg.MODE_ROYALE, 8, 500, 16, TEAMS),
]


def oracles_for(arena, mode, n, base, teams, max_steps):
    out = []
    for e in range(n):
        lvl = 1 if mode == sfcfg.MODE_ROYALE else 1 + (base + e) % 10
        o = sfo.Arena(sfcfg.make_config(arena, mode=mode, level_min=lvl, max_steps=max_steps, teams=teams))
        o.reset(lvl, common.synth_tb(base + e), common.synth_serial(base + e, 0))
        out.append((o, lvl))
    return out


def beyond_reference(o):
    """humans, zombies, bullets and dynamic cells of an oracle arena that lie outside the reference's 30 x 100"""
    n = 0
    for (kind, index), f in sfo.parse_record(o.dump()).items():
        if kind == 3 and f[0]:
            r, c = f[5], f[6]
        elif kind == 4:
            r, c = f[2], f[3]
        elif kind == 5:
            r, c = f[1], f[2]
        elif kind == 7:
            r, c = index // sfcfg.COLS % sfcfg.ROWS, index % sfcfg.COLS
        else:
            continue
        n += r >= sfdata.REF_ROWS or c >= sfdata.REF_COLS
    return n


def run_host(arena):
    import hostcheck
    for name, mode, n, steps, agents, teams in CASES:
        base, max_steps = 3000, 512
        cfg = sfcfg.make_config(arena, n_envs=n, mode=mode, level_min=1, level_max=1 if teams else 10, auto_reset=True,
                                max_steps=max_steps, env_id_base=base, teams=teams)
        hs = hostcheck.HostSim(cfg)
        ora = oracles_for(arena, mode, n, base, teams, max_steps)
        episode = [0] * n
        for t in range(steps):
            act = common.synth_actions(range(base, base + n), agents, t, sfcfg.ACTIONS28)
            hs.step(act.tobytes())
            for e, (o, lvl) in enumerate(ora):
                st = o.step(bytes(act[e]))
                assert hs.step_out(e)["status"] == st, "%s: status, step %d arena %d" % (name, t, e)
                if st != sfcfg.RUNNING:
                    episode[e] += 1
                    o.reset(lvl, common.synth_tb(base + e), common.synth_serial(base + e, episode[e]))
                assert np.uint64(hs.state_hash(e)) == np.uint64(o.state_hash()), "%s: state, step %d arena %d" % (name, t, e)
        out = sum(beyond_reference(o) for o, _ in ora)
        assert out > 0, "%s: nothing ever left the reference's 30 x 100" % name
        print("host %s: %d arenas x %d steps on %dx%d, %d episodes: bit-exact (%d entities / dynamic cells outside the reference's "
              "30x100 at the end)" % (name, n, steps, sfcfg.ROWS, sfcfg.COLS, sum(episode), out))


def run_gpu(arena):
    import torch
    from strikeforce_b200.sim import BatchedArena
    assert torch.cuda.is_available()
    for name, mode, n, steps, agents, teams in CASES:
        n, base, max_steps = n * 8, 3000, 512
        sim = BatchedArena(n, mode=mode, level=1, level_max=1 if teams else 10, auto_reset=True, max_steps=max_steps,
                           env_id_base=base, teams=teams)
        ora = oracles_for(arena, mode, n, base, teams, max_steps)
        episode = [0] * n
        mask = (1 << agents) - 1
        try:
            for t in range(steps):
                act = common.synth_actions(range(base, base + n), agents, t, sfcfg.ACTIONS28)
                if t % 50 == 0:
                    a = sim.observe(mask)
                    b = sim.observe(mask, channels_last=True)
                    assert torch.equal(a.view(torch.int32), b.contiguous().view(torch.int32)), "%s: layouts differ" % name
                    a_h = a.cpu().numpy()
                    for e, (o, _) in enumerate(ora):
                        for slot in range(agents):
                            try:
                                ref = o.observe(slot)
                            except RuntimeError:
                                continue
                            assert (a_h[e, slot].reshape(-1).view(np.uint32) == ref.view(np.uint32)).all(), \
                                "%s: observation, step %d arena %d slot %d" % (name, t, e, slot)
                sim.step(torch.from_numpy(act).to(sim.device))
                out = sim.step_out().cpu().numpy()
                h = sim.state_hash().cpu().numpy().view(np.uint64)
                for e, (o, lvl) in enumerate(ora):
                    st = o.step(bytes(act[e]))
                    assert out[e, 0] == st, "%s: status, step %d arena %d" % (name, t, e)
                    if st != sfcfg.RUNNING:
                        episode[e] += 1
                        o.reset(lvl, common.synth_tb(base + e), common.synth_serial(base + e, episode[e]))
                    assert h[e] == np.uint64(o.state_hash()), "%s: state, step %d arena %d" % (name, t, e)
            out = sum(beyond_reference(o) for o, _ in ora)
            assert out > 0, "%s: nothing ever left the reference's 30 x 100" % name
            print("gpu %s: %d arenas x %d steps on %dx%d, %d episodes: bit-exact (%d entities / dynamic cells outside the "
                  "reference's 30x100 at the end)" % (name, n, steps, sfcfg.ROWS, sfcfg.COLS, sum(episode), out))
        finally:
            sim.close()


if __name__ == "__main__":
    assert sfcfg.GEOMETRY_TAG, "run with SF_GEOMETRY=ROWSxCOLS"
    arena = sfdata.load_default()
    (run_gpu if sys.argv[1:] == ["gpu"] else run_host)(arena)
    print("LARGE MAP OK")
