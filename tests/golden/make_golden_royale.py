#!/usr/bin/env python
"""Generate tests/golden/royale_*.npz: Battle Royale matches played by the UNMODIFIED reference
(oracle/_ref/libsfref.so) through its own replay reader -- the one way the reference runs its online
modes without a match server (gameplay.hpp:1762-1806, 1847-1859, 966-986).

The replay file must list, step by step, the command of `ind` and of every other player alive when
human_action runs; which players are alive is taken from the C model (oracle/sf_oracle.c) while the
file is written.  The fixture itself -- status, state hash after every step, observations -- is what
the reference then produces from that file; had the model been wrong about a death, the reference
would read a shifted command stream and the two would part within a step.
Run in the build container (needs oracle/_ref)."""
import os
import sys
import tempfile
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import sfo  # noqa: E402
import sfref  # noqa: E402
from strikeforce_b200 import config as sfcfg  # noqa: E402
from strikeforce_b200 import data as sfdata  # noqa: E402
from strikeforce_b200 import replay  # noqa: E402

CAPS = dict(sfcfg.DEFAULT_CAPS, cap_portals=128, cap_built=1000, cap_bullets=128)
CAP_KEYS = ("cap_humans", "cap_zombies", "cap_bullets", "cap_chests", "cap_built", "cap_portals")

# name, teams, tb, serial, steps, player sheet (the two short matches run until they end: one is won
# -- every rival dead --, one is lost)
CASES = [
    ("royale_16p_4teams", [1, 2, 3, 4] * 4, 1700000123, 424242, 1000, "account1"),
    ("royale_4p_2teams", [1, 2, 2, 1], 1700000777, 31337, 2500, "account1"),
    ("royale_2p_newplayer_a", [1, 2], 1700000003, 1015, 8000, "new_player"),
    ("royale_3p_newplayer_b", [1, 2, 2], 1700000002, 2002, 8000, "new_player"),
    # every player with a sheet of its own, as announced over the wire (get_info, gameplay.hpp:131-149)
    ("royale_6p_mixed_sheets", [1, 2, 3, 1, 2, 3], 1700000003, 3003, 3000,
     ["account1", "new_player", "synthetic", "new_player", "account1", "synthetic"]),
    # the copy of a match that belongs to a player other than 0 ("players ind team" of the header,
    # gameplay.hpp:1797-1799): credits, the corpse that keeps its cell and the end of the match follow `ind`
    ("royale_5p_ind2", [1, 2, 1, 3, 2], 1700000055, 55055, 3000,
     ["new_player", "account1", "account1", "synthetic", "new_player"], 2),
    ("royale_4p_ind3_short", [1, 2, 2, 1], 1700000009, 9009, 8000, "new_player", 3),
]


def main(which):
    """One match per process: the reference keeps its humans in globals, and a slot that held a
    player of team 3 in the match before would hand that team to the NPC spawned into it."""
    arena = sfdata.load_default()
    for case in [CASES[which]]:
        name, teams, tb, serial, steps, player = case[:6]
        ind = case[6] if len(case) > 6 else 0
        names = player if isinstance(player, list) else [player] * len(teams)
        sheet = np.stack([arena.player_sheet(n) for n in names])
        P = len(teams)
        rng = np.random.default_rng(zlib.crc32(name.encode()))
        table = np.frombuffer(bytes(sfcfg.ACTIONS28), dtype=np.uint8)
        actions = table[rng.integers(len(table), size=(steps, P))]
        cfg = sfcfg.make_config(arena, mode=sfcfg.MODE_ROYALE, teams=teams, auto_reset=False, caps=CAPS, sheets=names, ind=ind)
        o = sfo.Arena(cfg)
        o.reset(1, tb, serial)
        file_cmds = bytearray()
        n = 0
        for t in range(steps):
            if o.step_a() != 0:
                break
            rec = sfo.parse_record(o.dump())
            file_cmds.append(actions[t, ind])
            file_cmds.extend(actions[t, i] for i in range(P) if i != ind and rec[(3, i)][0])
            n = t + 1
            if o.step_b(bytes(actions[t])) != 0:
                break
        path = os.path.join(tempfile.mkdtemp(), name + ".sf_sample")
        replay.write_royale(path, tb, serial, sheet, teams, bytes(file_cmds), ind=ind)
        sfref.reset_replay(sfcfg.MODE_ROYALE, 1, path, caps=[CAPS[k] for k in CAP_KEYS])
        status = np.zeros(n, dtype=np.int32)
        hashes = np.zeros(n + 1, dtype=np.uint64)
        hashes[0] = sfref.state_hash()
        obs, obs_steps, obs_last, records, rec_steps = [], [], [], [], []
        for t in range(n):
            if t % 100 == 0:
                try:
                    obs.append(sfref.observe(0))
                except RuntimeError:  # player 0 is a remote player here (ind != 0) and has died
                    obs.append(np.full(sfref.OBS_LEN, np.nan, dtype=np.float32))
                try:
                    obs_last.append(sfref.observe(P - 1))
                except RuntimeError:  # that player is dead: its agent is gone
                    obs_last.append(np.full(sfref.OBS_LEN, np.nan, dtype=np.float32))
                obs_steps.append(t)
            status[t] = sfref.step(b"+" * P)  # in replay mode every command comes from the file
            hashes[t + 1] = sfref.state_hash()
            if t % 250 == 249 or status[t] != 0:
                records.append(sfref.dump())
                rec_steps.append(t)
            if status[t] != 0:
                status, hashes = status[:t + 1], hashes[:t + 2]
                break
        actions = actions[:len(status)]
        assert o.status() == status[-1] and np.uint64(o.state_hash()) == hashes[-1] or status[-1] in (5, 6)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), mode=sfcfg.MODE_ROYALE, level=1, tb=tb, serial=serial,
                            squad_agents=0, player=names[ind], sheets=np.array(names), teams=np.array(teams), ind=ind,
                            caps=np.array([CAPS[k] for k in CAP_KEYS]),
                            actions=actions, status=status, hashes=hashes, records=np.concatenate(records),
                            rec_len=np.array([len(r) for r in records], dtype=np.int64), rec_steps=np.array(rec_steps),
                            obs=np.stack(obs), obs_last=np.stack(obs_last), obs_steps=np.array(obs_steps),
                            counters=np.array(list(sfref.counters().values())),
                            population=np.array(list(sfref.population().values())))
        print(name, "steps", len(status), "final status", int(status[-1]), sfref.population(), sfref.counters())


if __name__ == "__main__":
    if len(sys.argv) > 1:
        main(int(sys.argv[1]))
    else:
        import subprocess
        for i in range(len(CASES)):
            subprocess.check_call([sys.executable, os.path.abspath(__file__), str(i)])
