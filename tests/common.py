"""Shared helpers of the parity tests: drive the C oracle (oracle/sf_oracle.c) and a device
or host-check simulator with the same seeds and action streams and compare canonical state."""
import numpy as np

import sfo
from strikeforce_b200 import config as sfcfg

SYNTH_TB0 = 1700000000


def synth_tb(e):
    return SYNTH_TB0 + e


def synth_serial(e, k=0):
    return (123456789 + 7919 * e + 104729 * k) & ((1 << 30) - 1)


def splitmix_draw(e, agent, t):
    """include/sf_synth.h sf_synth_draw_at"""
    m = (1 << 64) - 1
    s = (42 + e * 1000003 + agent + t * 0x9E3779B97F4A7C15) & m
    z = (s + 0x9E3779B97F4A7C15) & m
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & m
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & m
    return z ^ (z >> 31)


def synth_actions(env_ids, n_agents, t, table):
    out = np.empty((len(env_ids), n_agents), dtype=np.uint8)
    for i, e in enumerate(env_ids):
        for a in range(n_agents):
            out[i, a] = table[splitmix_draw(e, a, t) % len(table)]
    return out


def make_oracles(arena_data, n_envs, mode, level, squad_agents=False, player="account1", max_steps=0, caps=None,
                 env_id_base=0):
    cfg = sfcfg.make_config(arena_data, mode=mode, level_min=level, squad_agents=squad_agents, max_steps=max_steps,
                            player=player, caps=caps)
    arenas = []
    for e in range(n_envs):
        a = sfo.Arena(cfg)
        ge = env_id_base + e
        a.reset(level, synth_tb(ge), synth_serial(ge, 0))
        arenas.append(a)
    return arenas


CAP_KEYS = ("cap_humans", "cap_zombies", "cap_bullets", "cap_chests", "cap_built", "cap_portals")


def golden_kwargs(g):
    """make_config / BatchedArena keyword arguments of a tests/golden fixture (the Battle Royale
    ones also carry the teams of their players and the capacities they were played with)."""
    kw = dict(mode=int(g["mode"]), squad_agents=bool(g["squad_agents"]), player=str(g["player"]))
    if "teams" in g:
        kw["teams"] = [int(t) for t in g["teams"]]
        kw["caps"] = dict(zip(CAP_KEYS, (int(c) for c in g["caps"])))
        if "sheets" in g:
            kw["sheets"] = [str(n) for n in g["sheets"]]
        if "ind" in g:
            kw["ind"] = int(g["ind"])
    return kw


MATCH_TABLE = b"+qeawsd"  # commands that never attack (tests/test_match_server.py)


def record_crcs(rec):
    """(crc of the whole canonical record with the header field `ind` cleared, crc of the record
    without header and cells): the second one is what every client's copy of an online match shares
    with every other copy even after kills (credits go to `ind` only, a dead `ind` keeps its cell)."""
    import zlib
    rec = np.array(rec, dtype=np.int32)
    rec[10] = 0
    kept, i = [], 0
    while i + 3 <= len(rec):
        n = 3 + int(rec[i + 2])
        if int(rec[i]) not in (1, 7):
            kept.append(rec[i:i + n])
        i += n
    return zlib.crc32(rec.tobytes()), zlib.crc32(np.concatenate(kept).tobytes())


# ------------------------------------------------------------------ whole-trajectory checks

MASK64 = (1 << 64) - 1


def mix64(x):
    """include/sf_canon.h sf_mix64 on a Python int"""
    x &= MASK64
    x ^= x >> 30
    x = (x * 0xbf58476d1ce4e5b9) & MASK64
    x ^= x >> 27
    x = (x * 0x94d049bb133111eb) & MASK64
    return x ^ (x >> 31)


def torch_mix64(torch, x):
    """sf_mix64 on an int64 tensor holding uint64 bit patterns (multiplication wraps; the shifts are made logical)"""
    def lshr(v, s):
        return (v >> s) & ((1 << (64 - s)) - 1)

    def i64(c):
        return c - (1 << 64) if c >= (1 << 63) else c
    x = x ^ lshr(x, 30)
    x = x * i64(0xbf58476d1ce4e5b9)
    x = x ^ lshr(x, 27)
    x = x * i64(0x94d049bb133111eb)
    return x ^ lshr(x, 31)


def _trace_worker(job):
    """One process of oracle_traces: plays the synthetic workload (include/sf_synth.h, auto-reset)
    for its arenas inside the C oracle and returns, per arena, (episodes, checksum over the state hash
    after every step, last hash, observation of slot 0 at the end as uint32 words or None)."""
    kw, envs, steps, table, with_obs = job
    import sfo as _sfo
    from strikeforce_b200 import data as _data
    arena = _data.load_default()
    span = kw["level_max"] - kw["level_min"] + 1
    out = []
    arenas = {}
    for e in envs:
        lvl = kw["level_min"] + e % span
        if lvl not in arenas:
            c = sfcfg.make_config(arena, mode=kw["mode"], level_min=lvl, squad_agents=kw.get("squad_agents", False),
                                  max_steps=kw.get("max_steps", 0), player=kw.get("player", "account1"),
                                  teams=kw.get("teams"), caps=kw.get("caps"))
            arenas[lvl] = _sfo.Arena(c)
        n, chk, last, obs = arenas[lvl].run_trace(e, lvl, steps, table, with_obs)
        out.append((e, n, chk, last, None if obs is None else obs.view(np.uint32).copy()))
    return out


def oracle_traces(kw, envs, steps, table, with_obs=False, procs=None):
    """{env: (episodes, trajectory checksum, last hash, obs words)} from the C oracle, all host cores."""
    import multiprocessing as mp
    import os
    envs = list(envs)
    procs = procs or min(len(envs), os.cpu_count() or 1)
    jobs = [(kw, envs[i::procs], steps, bytes(table), with_obs) for i in range(procs)]
    ctx = mp.get_context("spawn")  # the parent may hold a CUDA context: never fork it
    with ctx.Pool(procs) as pool:
        res = pool.map(_trace_worker, jobs)
    return {r[0]: r[1:] for part in res for r in part}
