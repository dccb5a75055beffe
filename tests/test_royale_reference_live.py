"""Battle Royale against the unmodified reference, live: a fresh seeded match is written as a match
file (strikeforce_b200.replay.write_royale), played by the reference through its own replay reader
("Battle Royal" in replay mode, gameplay.hpp:1762-1806, 1847-1859, 966-986) and followed by the C
oracle step for step.  The reference keeps its humans in process globals, so the match runs in a
process of its own (tests/golden/make_golden_royale.py explains why)."""
import os
import subprocess
import sys
import textwrap

import pytest

import sfref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = textwrap.dedent("""
    import os, sys, tempfile
    import numpy as np
    root = sys.argv[1]
    for p in (root, os.path.join(root, "oracle"), os.path.join(root, "tests")):
        sys.path.insert(0, p)
    import sfo, sfref, common
    from strikeforce_b200 import config as sfcfg, data as sfdata, replay

    arena = sfdata.load_default()
    teams, names = [2, 1, 3, 1, 2], ["new_player", "account1", "synthetic", "new_player", "account1"]
    caps = dict(sfcfg.DEFAULT_CAPS, cap_portals=128, cap_built=1000, cap_bullets=128)
    cfg = sfcfg.make_config(arena, mode=sfcfg.MODE_ROYALE, teams=teams, sheets=names, auto_reset=False, caps=caps)
    tb, serial, steps, P = 1700004242, 987654, 700, len(teams)
    o = sfo.Arena(cfg)
    o.reset(1, tb, serial)
    dumps, stats, cmds = [o.dump()], [], bytearray()
    for t in range(steps):
        act = common.synth_actions([11], P, t, sfcfg.ACTIONS28)[0]
        if o.step_a() != 0:
            break
        rec = sfo.parse_record(o.dump())
        cmds.append(act[0])
        cmds.extend(act[i] for i in range(1, P) if rec[(3, i)][0])
        stats.append(o.step_b(bytes(act)))
        dumps.append(o.dump())
        if stats[-1] != 0:
            break
    path = os.path.join(tempfile.mkdtemp(), "live.sf_sample")
    replay.write_royale(path, tb, serial, np.stack([arena.player_sheet(n) for n in names]), teams, bytes(cmds))
    sfref.reset_replay(sfcfg.MODE_ROYALE, 1, path, caps=[caps[k] for k in common.CAP_KEYS])
    d = sfref.dump()
    assert len(d) == len(dumps[0]) and (d == dumps[0]).all(), sfo.diff_records(dumps[0], d)
    for t, st in enumerate(stats):
        assert sfref.step(b"+" * P) == st, ("status", t)
        if st in (0, 1, 2, 3):
            d = sfref.dump()
            assert len(d) == len(dumps[t + 1]) and (d == dumps[t + 1]).all(), (t, sfo.diff_records(dumps[t + 1], d))
    for slot in (0, P - 1):  # what two of the players' own clients would observe
        try:
            a, b = sfref.observe(slot), o.observe(slot)
        except RuntimeError:
            continue
        assert (a.view(np.uint32) == b.view(np.uint32)).all(), ("observation", slot)
    print("LIVE-OK", len(stats), stats[-1])
""")


def test_reference_plays_a_royale_match_file_like_the_oracle():
    if not sfref.available():
        pytest.skip("oracle/_ref/libsfref.so not built (needs /root/reference)")
    r = subprocess.run([sys.executable, "-c", SCRIPT, ROOT], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "LIVE-OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
