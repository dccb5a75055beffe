"""The device tick (strikeforce_b200/csrc/sf_core.cuh) compiled for the host and compared with
the C oracle -- a CPU-side check of the CUDA algorithm (layout, log-domain RNG, lock-step
phases, observation features) for containers without a GPU.  The GPU tests repeat this
through the real kernels and the C ABI."""
import glob
import os

import numpy as np
import pytest

import common
import hostcheck
import sfo
from strikeforce_b200 import config as sfcfg

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = sorted(p for p in glob.glob(os.path.join(GOLDEN, "*.npz")) if not p.endswith("kat.npz"))


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[:-4] for p in CASES])
def test_device_tick_matches_reference_golden(path, arena_data):
    g = np.load(path)
    cfg = sfcfg.make_config(arena_data, n_envs=2, level_min=int(g["level"]), auto_reset=False, **common.golden_kwargs(g))
    hs = hostcheck.HostSim(cfg)
    hs.reset(1, int(g["tb"]), int(g["serial"]))  # arena 1 plays the golden match, arena 0 another one
    assert np.uint64(hs.state_hash(1)) == g["hashes"][0]
    obs = dict(zip(g["obs_steps"].tolist(), g["obs"]))
    obs_last = dict(zip(g["obs_steps"].tolist(), g["obs_last"])) if "obs_last" in g else {}
    idle = bytes(b"+" * hs.n_agents)
    for t, act in enumerate(g["actions"]):
        if t in obs and not np.isnan(obs[t][0]):
            assert (hs.observe(1, 0).view(np.uint32) == obs[t].view(np.uint32)).all(), "observation, step %d" % t
        if t in obs_last and not np.isnan(obs_last[t][0]):
            got = hs.observe(1, hs.n_agents - 1)
            assert (got.view(np.uint32) == obs_last[t].view(np.uint32)).all(), "observation of the last player, step %d" % t
        hs.step(idle + bytes(act))
        assert hs.status(1) == g["status"][t], "status, step %d" % t
        assert np.uint64(hs.state_hash(1)) == g["hashes"][t + 1], "state hash, step %d" % t


def test_auto_reset_follows_the_seed_chain(arena_data):
    """Episodes shorter and longer than the pending-stream warm-up (128 steps)."""
    for max_steps, steps in ((40, 130), (150, 320)):
        n, base = 3, 5
        cfg = sfcfg.make_config(arena_data, n_envs=n, mode=sfcfg.MODE_SOLO, level_min=1, auto_reset=True,
                                max_steps=max_steps, env_id_base=base)
        hs = hostcheck.HostSim(cfg)
        oracles = common.make_oracles(arena_data, n, sfcfg.MODE_SOLO, 1, max_steps=max_steps, env_id_base=base)
        episode = [0] * n
        for t in range(steps):
            act = common.synth_actions(range(base, base + n), 1, t, sfcfg.ACTIONS28)
            hs.step(act.tobytes())
            for e, o in enumerate(oracles):
                st = o.step(bytes(act[e]))
                assert hs.step_out(e)["status"] == st
                if st != sfcfg.RUNNING:
                    episode[e] += 1
                    o.reset(1, common.synth_tb(base + e), common.synth_serial(base + e, episode[e]))
                d0, d1 = o.dump(), hs.dump(e)
                assert len(d0) == len(d1) and (d0 == d1).all(), sfo.diff_records(d0, d1)
        assert hs.stats()["episodes"] == sum(episode) and hs.stats()["steps"] == n * steps


def test_split_step_and_p2_observation(arena_data):
    """sf_step_a / sf_step_b with the P2 observation point in between (gameplay.hpp:933)."""
    cfg = sfcfg.make_config(arena_data, n_envs=1, mode=sfcfg.MODE_SQUAD, level_min=2, squad_agents=True, auto_reset=False)
    hs = hostcheck.HostSim(cfg)
    o = sfo.Arena(cfg)
    o.reset(2, 1700000555, 777)
    hs.reset(0, 1700000555, 777)
    rng = np.random.default_rng(3)
    for t in range(300):
        act = bytes(sfcfg.ACTIONS9[i] for i in rng.integers(9, size=10))
        o.step_a()
        hs.step(None, half=1)
        if o.status() != sfcfg.RUNNING:
            break
        if t % 25 == 0:
            for slot in (1, 6):
                try:
                    ref = o.observe(slot)
                except RuntimeError:
                    continue
                assert (hs.observe(0, slot).view(np.uint32) == ref.view(np.uint32)).all()
        o.step_b(act)
        hs.step(act, half=2)
        assert hs.step_out(0) == o.step_out()
        if o.status() != sfcfg.RUNNING:
            break
        d0, d1 = o.dump(), hs.dump(0)
        assert len(d0) == len(d1) and (d0 == d1).all(), sfo.diff_records(d0, d1)


def test_overflow_guard_and_truncation(arena_data):
    caps = dict(cap_humans=12, cap_zombies=8, cap_bullets=6, cap_built=8, cap_portals=8)
    n = 6
    cfg = sfcfg.make_config(arena_data, n_envs=n, mode=sfcfg.MODE_SOLO, level_min=1, auto_reset=False, caps=caps,
                            max_steps=250)
    hs = hostcheck.HostSim(cfg)
    oracles = common.make_oracles(arena_data, n, sfcfg.MODE_SOLO, 1, max_steps=250, caps=caps)
    seen = set()
    for t in range(260):
        act = common.synth_actions(range(n), 1, t, sfcfg.ACTIONS28)
        hs.step(act.tobytes())
        for e, o in enumerate(oracles):
            st = o.step(bytes(act[e]))
            assert hs.status(e) == st
            seen.add(st)
            if st in (sfcfg.RUNNING, sfcfg.TRUNCATED, sfcfg.DEAD):
                assert np.uint64(hs.state_hash(e)) == np.uint64(o.state_hash())
    assert sfcfg.OVERFLOW in seen or sfcfg.TRUNCATED in seen


def test_host_helpers_match_oracle():
    L = hostcheck.lib()
    for x in (0, 1, 2, 99, 100, 150, 225, 275, 1000, 1135, 4096, 65536, 1000000):
        for y in (1, 2, 3, 50, 100, 255):
            assert L.hc_compute_damage(x, y) == sfo.compute_damage(x, y)
    for n in list(range(0, 3000, 7)) + [15000, 999999, 1000000, 1048575]:
        assert np.float32(L.hc_obs_transform_milli(n)).view(np.uint32) == \
            np.float32(sfo.obs_transform(np.float32(n / 1000.0))).view(np.uint32)


def test_portal_heavy_stream_matches_oracle(arena_data):
    """A command stream that keeps building portals and walking into them: humans that wait on an
    entrance whose exit is taken (HS_ON_ENT, sf_obey) and the exit watch flag of sf_portal_damage
    (SfDev::misc bit 26) both come and go; every seventh step the whole state is compared."""
    n, steps, table = 4, 1500, b"]]]]wasdwasdqe+xz[u"
    cfg = sfcfg.make_config(arena_data, n_envs=n, mode=sfcfg.MODE_SQUAD, level_min=10, auto_reset=False)
    hs = hostcheck.HostSim(cfg)
    oracles = common.make_oracles(arena_data, n, sfcfg.MODE_SQUAD, 10)
    for e in range(n):
        hs.reset(e, common.synth_tb(e), common.synth_serial(e, 0))
    alive, watch_seen = [True] * n, set()
    for t in range(steps):
        act = common.synth_actions(range(n), 1, t, table)
        hs.step(act.tobytes())
        for e, o in enumerate(oracles):
            if not alive[e]:
                continue
            st = o.step(bytes(act[e]))
            assert hs.step_out(e)["status"] == st, (t, e)
            if st != sfcfg.RUNNING:
                alive[e] = False
                continue
            watch_seen.add((hs.misc(e) >> 26) & 1)
            if t % 7 == 0:
                d0, d1 = o.dump(), hs.dump(e)
                assert len(d0) == len(d1) and (d0 == d1).all(), (t, e, sfo.diff_records(d0, d1))
    assert watch_seen == {0, 1}


def test_royale_auto_reset_follows_the_seed_chain(arena_data):
    """Battle Royale with auto-reset: every new episode seeds the next stream of the chain and then
    draws the players' ways and cells from it (gameplay.hpp:1847-1859)."""
    teams, n, base, max_steps, steps = [1, 2, 2, 1, 3], 3, 7, 60, 200
    cfg = sfcfg.make_config(arena_data, n_envs=n, mode=sfcfg.MODE_ROYALE, teams=teams, auto_reset=True,
                            max_steps=max_steps, env_id_base=base)
    hs = hostcheck.HostSim(cfg)
    ocfg = sfcfg.make_config(arena_data, mode=sfcfg.MODE_ROYALE, teams=teams, max_steps=max_steps)
    oracles = []
    for e in range(n):
        o = sfo.Arena(ocfg)
        o.reset(1, common.synth_tb(base + e), common.synth_serial(base + e, 0))
        oracles.append(o)
    episode = [0] * n
    for t in range(steps):
        act = common.synth_actions(range(base, base + n), len(teams), t, sfcfg.ACTIONS28)
        hs.step(act.tobytes())
        for e, o in enumerate(oracles):
            st = o.step(bytes(act[e]))
            assert hs.step_out(e)["status"] == st
            if st != sfcfg.RUNNING:
                episode[e] += 1
                o.reset(1, common.synth_tb(base + e), common.synth_serial(base + e, episode[e]))
            d0, d1 = o.dump(), hs.dump(e)
            assert len(d0) == len(d1) and (d0 == d1).all(), sfo.diff_records(d0, d1)
    assert hs.stats()["episodes"] == sum(episode) == 9


SPILL_STRESS = ("-DSF_BT_SLOTS=7", "-DSF_BT_FULL=5")  # a bullet-flag table that overflows all the time


@pytest.mark.parametrize("mode,extra", [(sfcfg.MODE_SQUAD, ()), (sfcfg.MODE_SQUAD, SPILL_STRESS), (sfcfg.MODE_TIMER, SPILL_STRESS)],
                         ids=["squad", "squad-tiny-flag-table", "timer-tiny-flag-table"])
def test_soak_against_the_oracle(arena_data, mode, extra):
    """Long episodes with auto-reset, levels 1-10, the whole 28-symbol alphabet: the populations the
    benchmark times (tens of humans, zombies and bullets, portals, built cells) and the paths only they
    reach -- bullets over exits, radiation taking a cell's flag over, absorbed bullets, a full flag
    table spilling into the overlay (the tiny-table builds do that constantly).  State hash and step
    status against the C oracle after every step."""
    n, steps, base = 12, 1300, 2000
    cfg = sfcfg.make_config(arena_data, n_envs=n, mode=mode, level_min=1, level_max=10, auto_reset=True, max_steps=1024,
                            env_id_base=base)
    hs = hostcheck.HostSim(cfg, extra=extra)
    oracles = []
    for e in range(n):
        lvl = 1 + (base + e) % 10
        o = sfo.Arena(sfcfg.make_config(arena_data, mode=mode, level_min=lvl, max_steps=1024))
        o.reset(lvl, common.synth_tb(base + e), common.synth_serial(base + e, 0))
        oracles.append(o)
    episode = [0] * n
    for t in range(steps):
        act = common.synth_actions(range(base, base + n), 1, t, sfcfg.ACTIONS28)
        hs.step(act.tobytes())
        for e, o in enumerate(oracles):
            st = o.step(bytes(act[e]))
            assert hs.step_out(e)["status"] == st, "status: step %d arena %d" % (t, e)
            if st != sfcfg.RUNNING:
                episode[e] += 1
                o.reset(1 + (base + e) % 10, common.synth_tb(base + e), common.synth_serial(base + e, episode[e]))
            assert np.uint64(hs.state_hash(e)) == np.uint64(o.state_hash()), "state: step %d arena %d" % (t, e)
    assert sum(episode) >= n  # every arena went through at least one truncation and reset
