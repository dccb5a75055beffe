#!/usr/bin/env python
"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libsfref.so).

Run in the build container, where /root/reference exists (oracle/ref_harness/build_ref.sh
builds the library from the reference's own headers).  Each fixture holds, for one seeded
match: the action stream, the per-step status and canonical-state hash, a few full canonical
records and a few observations as produced by the reference's own gameplay::bot().
One process can host one arena only (the reference keeps its state in globals), so the
matches are generated one after the other.
"""
import os
import zlib
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import sfref  # noqa: E402
from strikeforce_b200 import config as sfcfg  # noqa: E402

CAPS = [sfcfg.DEFAULT_CAPS[k] for k in ("cap_humans", "cap_zombies", "cap_bullets", "cap_chests", "cap_built",
                                        "cap_portals")]

# name, mode, level, tb, serial, steps, alphabet, squad_agents, player template
CASES = [
    ("solo_l1_9act", sfcfg.MODE_SOLO, 1, 1700000000, 123456789, 1500, sfcfg.ACTIONS9, False, "account1"),
    ("solo_l1_28act", sfcfg.MODE_SOLO, 1, 1700000007, 987654321, 1500, sfcfg.ACTIONS28, False, "account1"),
    ("timer_l2_28act", sfcfg.MODE_TIMER, 2, 1700000011, 55555555, 1200, sfcfg.ACTIONS28, False, "account1"),
    ("squad_l1_28act", sfcfg.MODE_SQUAD, 1, 1700000042, 424242424, 1200, sfcfg.ACTIONS28, False, "account1"),
    ("squad_l3_agents", sfcfg.MODE_SQUAD, 3, 1700000099, 99999999, 1000, sfcfg.ACTIONS9, True, "account1"),
    ("solo_l4_longseed", sfcfg.MODE_SOLO, 4, 123456789012345, 987654321098, 600, sfcfg.ACTIONS28, False, "account1"),
]


def main():
    for name, mode, level, tb, serial, steps, table, agents, player in CASES:
        rng = np.random.default_rng(zlib.crc32(name.encode()))
        sfref.reset(mode, level, tb, serial, squad_agents=agents, caps=CAPS)
        n_agents = 10 if (agents and mode == sfcfg.MODE_SQUAD) else 1
        actions = np.frombuffer(bytes(table), dtype=np.uint8)[rng.integers(len(table), size=(steps, n_agents))]
        status = np.zeros(steps, dtype=np.int32)
        hashes = np.zeros(steps + 1, dtype=np.uint64)
        hashes[0] = sfref.state_hash()
        records, rec_steps, obs, obs_steps = [], [], [], []
        for t in range(steps):
            if t % 100 == 0:
                obs.append(sfref.observe(0))
                obs_steps.append(t)
            status[t] = sfref.step(bytes(actions[t]))
            hashes[t + 1] = sfref.state_hash()
            if t % 250 == 249 or status[t] != 0:
                records.append(sfref.dump())
                rec_steps.append(t)
            if status[t] != 0:
                status, hashes, actions = status[:t + 1], hashes[:t + 2], actions[:t + 1]
                break
        rec_len = np.array([len(r) for r in records], dtype=np.int64)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), mode=mode, level=level, tb=tb, serial=serial,
                            squad_agents=int(agents), player=player, actions=actions, status=status, hashes=hashes,
                            records=np.concatenate(records), rec_len=rec_len, rec_steps=np.array(rec_steps),
                            obs=np.stack(obs), obs_steps=np.array(obs_steps), counters=np.array(list(sfref.counters().values())),
                            population=np.array(list(sfref.population().values())))
        print(name, "steps", len(status), "final status", int(status[-1]), sfref.population())
    # known answers of the small pure functions
    kat_seeds = [(1700000000, 123456789), (0, 0), (1771155561, 1073741823), (99999999999999, 31337)]
    draws = []
    for tb, serial in kat_seeds:
        sfref.srand(tb, serial)
        draws.append([sfref.rand() for _ in range(64)])
    cd = [(x, y, sfref.compute_damage(x, y)) for x in (0, 1, 2, 99, 100, 150, 225, 275, 1000, 1135, 4096, 65536, 1000000)
          for y in (1, 2, 3, 50, 100, 255)]
    np.savez_compressed(os.path.join(HERE, "kat.npz"), seeds=np.array(kat_seeds, dtype=np.int64),
                        draws=np.array(draws, dtype=np.int32), compute_damage=np.array(cd, dtype=np.int64))
    print("kat ok")


if __name__ == "__main__":
    main()
