"""The N>1 path on CPU: two gloo ranks each simulate their shard (host-check build of the device
tick), reduce the episode statistics, and must reproduce the single-process batch exactly."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N_TOTAL, STEPS, MAX_STEPS = 7, 90, 40


def _simulate(env_id_base, n_local):
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests", "hostcheck"), os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import common
    import hostcheck
    from strikeforce_b200 import config as sfcfg
    from strikeforce_b200 import data as sfdata
    cfg = sfcfg.make_config(sfdata.load_default(), n_envs=n_local, mode=sfcfg.MODE_SQUAD, level_min=1, level_max=10,
                            auto_reset=True, max_steps=MAX_STEPS, env_id_base=env_id_base)
    hs = hostcheck.HostSim(cfg)
    for t in range(STEPS):
        act = common.synth_actions(range(env_id_base, env_id_base + n_local), hs.n_agents, t, sfcfg.ACTIONS28)
        hs.step(act.tobytes())
    hashes = np.array([hs.state_hash(e) for e in range(n_local)], dtype=np.uint64)
    stats = np.array(list(hs.stats().values()), dtype=np.int64)
    return hashes, stats


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from strikeforce_b200 import dist as sfdist
    base, n_local = sfdist.shard(N_TOTAL, rank, world)
    hashes, stats = _simulate(base, n_local)
    st = sfdist.reduce_stats(torch.from_numpy(stats.copy()))
    gathered = [None] * world
    dist.all_gather_object(gathered, (base, hashes.tolist()))
    worst = sfdist.max_over_ranks(float(rank + 1), "cpu")
    if rank == 0:
        out.put((st.tolist(), gathered, worst))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_reproduce_the_single_batch():
    sys.path.insert(0, ROOT)
    from strikeforce_b200 import dist as sfdist
    assert [sfdist.shard(7, r, 2) for r in range(2)] == [(0, 4), (4, 3)]
    assert sum(sfdist.shard(1048576, r, 8)[1] for r in range(8)) == 1048576
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    stats2, gathered, worst = out.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    hashes1, stats1 = _simulate(0, N_TOTAL)
    merged = np.zeros(N_TOTAL, dtype=np.uint64)
    for base, h in gathered:
        merged[base:base + len(h)] = np.array(h, dtype=np.uint64)
    assert (merged == hashes1).all(), "sharded arenas differ from the single batch"
    assert stats2 == stats1.tolist(), "reduced statistics differ"
    assert stats1[0] == N_TOTAL * STEPS and stats1[1] >= N_TOTAL * 2
    assert worst == 2.0
