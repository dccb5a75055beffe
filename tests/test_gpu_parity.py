"""GPU parity tests: the CUDA path, called through the C ABI, against the C oracle on the same
seeded inputs (bit-exact: state hash of the canonical record every step, step outputs,
observations as fp32 words)."""
import numpy as np
import pytest

import common
from strikeforce_b200 import config as sfcfg

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device; there is no CPU fallback to test")
    return torch


def _run_parity(torch, arena_data, mode, level, n_envs, steps, table, squad_agents=False, player="account1",
                max_steps=0, caps=None, obs_every=0, obs_mask=1):
    from strikeforce_b200.sim import BatchedArena
    sim = BatchedArena(n_envs, mode=mode, level=level, squad_agents=squad_agents, auto_reset=False,
                       max_steps=max_steps, player=player, caps=caps)
    oracles = common.make_oracles(arena_data, n_envs, mode, level, squad_agents, player, max_steps, caps)
    try:
        h_dev = sim.state_hash().cpu().numpy().view(np.uint64)
        h_ora = np.array([o.state_hash() for o in oracles], dtype=np.uint64)
        assert (h_dev == h_ora).all(), "state after reset differs"
        done = np.zeros(n_envs, dtype=bool)
        for t in range(steps):
            act = sim.synth_actions(t, table)
            act_h = act.cpu().numpy()
            assert (act_h == common.synth_actions(range(n_envs), sim.n_agents, t, table)).all()
            sim.step(act)
            out = sim.step_out().cpu().numpy()
            for e, o in enumerate(oracles):
                o.step(bytes(act_h[e]))
                so = o.step_out()
                assert out[e, 0] == so["status"], "status differs: env %d step %d: %d vs %d" % (e, t, out[e, 0], so["status"])
                if so["status"] in (sfcfg.RUNNING, sfcfg.WIN, sfcfg.DEAD, sfcfg.TIMEOUT, sfcfg.TRUNCATED) and not done[e]:
                    ref = [so[k] for k in ("status", "d_kills", "d_teams_kills", "d_loot", "d_hp", "d_damage",
                                           "d_effect", "episode_steps")]
                    assert out[e].tolist() == ref, "step_out differs: env %d step %d" % (e, t)
                done[e] |= so["status"] != sfcfg.RUNNING
            h_dev = sim.state_hash().cpu().numpy().view(np.uint64)
            for e, o in enumerate(oracles):
                if o.status() in (sfcfg.OVERFLOW, sfcfg.UB_GUARD):
                    continue  # aborted mid-step: only the status is defined
                if h_dev[e] != np.uint64(o.state_hash()):
                    diff = sfo_diff(o, sim, e)
                    pytest.fail("state differs: env %d step %d\n%s" % (e, t, "\n".join(diff)))
            if obs_every and t % obs_every == 0:
                nsel = bin(obs_mask).count("1")
                obs = sim.observe(agent_mask=obs_mask).cpu().numpy().reshape(n_envs, nsel, -1)
                slots = [i for i in range(32) if (obs_mask >> i) & 1]
                for e, o in enumerate(oracles):
                    for j, slot in enumerate(slots):
                        try:
                            ref = o.observe(slot)
                        except RuntimeError:
                            continue  # no active agent on that slot: the reference builds nothing
                        assert (obs[e, j].view(np.uint32) == ref.view(np.uint32)).all(), \
                            "observation differs: env %d slot %d step %d" % (e, slot, t)
        return sim.stats()
    finally:
        sim.close()


def sfo_diff(oracle, sim, env):
    import sfo
    return sfo.diff_records(oracle.dump(), sim.export_env(env))


def test_rng_known_answers(torch_cuda, arena_data):
    """random.hpp:54-76 on the device against the reference's known answers and the oracle."""
    import sfo
    from strikeforce_b200.sim import BatchedArena
    sim = BatchedArena(32)
    try:
        tb = [1700000000, 0, 1771155561] + [1700000000 + i for i in range(200)]
        serial = [123456789, 0, 1073741823] + [common.synth_serial(i) for i in range(200)]
        out = sim.rng_stream(tb, serial, 64)
        assert out[:16, 0].tolist() == [285, 813, 468, 757, 905, 105, 521, 980, 354, 530, 55, 911, 702, 226, 255, 98]
        assert out[:8, 1].tolist() == [669, 110, 539, 452, 774, 356, 42, 334]
        assert out[:8, 2].tolist() == [881, 744, 641, 806, 562, 855, 824, 187]
        for i in range(len(tb)):
            r = sfo.Rng(tb[i], serial[i])
            assert out[:, i].tolist() == [r.rand() for _ in range(64)]
    finally:
        sim.close()


def test_solo_parity(torch_cuda, arena_data):
    st = _run_parity(torch_cuda, arena_data, sfcfg.MODE_SOLO, 1, 96, 400, sfcfg.ACTIONS9, obs_every=50)
    assert st["steps"] == 96 * 400


def test_timer_parity_full_alphabet(torch_cuda, arena_data):
    _run_parity(torch_cuda, arena_data, sfcfg.MODE_TIMER, 2, 64, 600, sfcfg.ACTIONS28, player="synthetic", obs_every=100)


def test_squad_parity(torch_cuda, arena_data):
    _run_parity(torch_cuda, arena_data, sfcfg.MODE_SQUAD, 1, 64, 600, sfcfg.ACTIONS28, obs_every=100)


def test_squad_agents_parity(torch_cuda, arena_data):
    _run_parity(torch_cuda, arena_data, sfcfg.MODE_SQUAD, 3, 48, 500, sfcfg.ACTIONS9, squad_agents=True, obs_every=100,
                obs_mask=0b1000001001)


def test_long_episode_parity(torch_cuda, arena_data):
    """Mid-episode populations (dozens of zombies, NPC-built blocks and portals)."""
    _run_parity(torch_cuda, arena_data, sfcfg.MODE_SOLO, 1, 8, 2500, sfcfg.ACTIONS28, obs_every=500)


def test_overflow_and_truncation_status(torch_cuda, arena_data):
    caps = dict(cap_humans=12, cap_zombies=8, cap_bullets=6, cap_built=8, cap_portals=8)
    _run_parity(torch_cuda, arena_data, sfcfg.MODE_SOLO, 1, 32, 400, sfcfg.ACTIONS28, caps=caps, max_steps=300)


@pytest.mark.parametrize("pinned", [False, True], ids=["pageable", "pinned"])
def test_step_host_matches_device_step(torch_cuda, arena_data, pinned):
    """sf_step_host with a pageable result buffer (copied after the step) and a page-locked one
    (written by the kernel itself)."""
    import torch
    from strikeforce_b200.sim import BatchedArena
    a = BatchedArena(64, mode="Solo", auto_reset=True, max_steps=50)
    b = BatchedArena(64, mode="Solo", auto_reset=True, max_steps=50)
    try:
        if pinned:
            keep = torch.zeros((64, 8), dtype=torch.int32).pin_memory()
            out_h = keep.numpy().view(sfcfg.STEP_OUT_DTYPE).reshape(64)
        else:
            out_h = np.zeros(64, dtype=sfcfg.STEP_OUT_DTYPE)
        for t in range(120):  # crosses two auto-resets
            act = a.synth_actions(t)
            a.step(act)
            act_h = act.cpu().pin_memory() if pinned else act.cpu()  # page-locked buffers are used in place
            b.step_host(act_h.numpy(), out_h)
            out_d = a.step_out().cpu().numpy()
            assert (out_d == out_h.view(np.int32).reshape(64, 8)).all()
        assert (a.state_hash() == b.state_hash()).all()
        st = a.stats()
        assert st["episodes"] == 64 * 2 and st["truncated"] + st["deaths"] + st["wins"] == st["episodes"]
    finally:
        a.close(), b.close()


def test_golden_matches_through_the_c_abi(torch_cuda, arena_data):
    """The reference's own trajectories (tests/golden, made by the unmodified reference) replayed on
    the GPU: custom seeds through sf_reset, state hash after every step, observations bit for bit."""
    import glob
    import os
    from strikeforce_b200.sim import BatchedArena
    golden = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    for path in sorted(glob.glob(os.path.join(golden, "*.npz"))):
        if path.endswith("kat.npz"):
            continue
        g = np.load(path)
        sim = BatchedArena(33, level=int(g["level"]), auto_reset=False, **common.golden_kwargs(g))
        try:
            e = 32  # the last arena of a partly filled warp
            sim.reset([e], [int(g["tb"])], [int(g["serial"])])
            assert np.uint64(sim.state_hash()[e].item() & 0xFFFFFFFFFFFFFFFF) == g["hashes"][0]
            obs = dict(zip(g["obs_steps"].tolist(), g["obs"]))
            obs_last = dict(zip(g["obs_steps"].tolist(), g["obs_last"])) if "obs_last" in g else {}
            act = torch_cuda.full((33, sim.n_agents), ord("+"), dtype=torch_cuda.uint8, device=sim.device)
            hashes = []
            for t, a in enumerate(g["actions"]):
                if t in obs and not np.isnan(obs[t][0]):
                    o = sim.observe(1)[e].reshape(-1).cpu().numpy()
                    assert (o.view(np.uint32) == obs[t].view(np.uint32)).all(), "%s observation step %d" % (path, t)
                if t in obs_last and not np.isnan(obs_last[t][0]):  # Battle Royale: the last player's view
                    o = sim.observe(1 << (sim.n_agents - 1))[e].reshape(-1).cpu().numpy()
                    assert (o.view(np.uint32) == obs_last[t].view(np.uint32)).all(), "%s last player, step %d" % (path, t)
                act[e] = torch_cuda.from_numpy(a.copy()).to(sim.device)
                sim.step(act)
                hashes.append(sim.state_hash()[e:e + 1])
            got = torch_cuda.cat(hashes).cpu().numpy().view(np.uint64)
            bad = np.nonzero(got != g["hashes"][1:])[0]
            assert len(bad) == 0, "%s: state differs from the reference at step %d" % (path, bad[0])
            assert sim.step_out()[e, 0].item() == g["status"][-1]
        finally:
            sim.close()


def test_auto_reset_chain_parity(torch_cuda, arena_data):
    """Auto-reset installs the pending stream (short and long episodes) and follows sf_synth.h."""
    from strikeforce_b200.sim import BatchedArena
    for max_steps, steps in ((40, 130), (150, 320)):
        n, base = 40, 1000
        sim = BatchedArena(n, mode="Solo", level=1, auto_reset=True, max_steps=max_steps, env_id_base=base)
        oracles = common.make_oracles(arena_data, n, sfcfg.MODE_SOLO, 1, max_steps=max_steps, env_id_base=base)
        episode = [0] * n
        try:
            for t in range(steps):
                act = sim.synth_actions(t, sfcfg.ACTIONS28)
                act_h = act.cpu().numpy()
                sim.step(act)
                out = sim.step_out().cpu().numpy()
                h = sim.state_hash().cpu().numpy().view(np.uint64)
                for e, o in enumerate(oracles):
                    st = o.step(bytes(act_h[e]))
                    assert out[e, 0] == st
                    if st != sfcfg.RUNNING:
                        episode[e] += 1
                        o.reset(1, common.synth_tb(base + e), common.synth_serial(base + e, episode[e]))
                    assert h[e] == np.uint64(o.state_hash()), "env %d step %d" % (e, t)
            assert sim.stats()["episodes"] == sum(episode)
        finally:
            sim.close()


def test_replay_of_sf_sample_logs(torch_cuda, arena_data, tmp_path):
    """.sf_sample logs (strikeforce_b200/replay.py) replayed on the GPU follow the oracle."""
    import sfo
    from strikeforce_b200 import replay
    from strikeforce_b200.sim import BatchedArena
    rng = np.random.default_rng(21)
    sheet = arena_data.player_sheet("account1")
    logs = []
    for i in range(5):
        cmds = bytes(sfcfg.ACTIONS28[j] for j in rng.integers(28, size=200))
        path = tmp_path / ("m%d.sf_sample" % i)
        replay.write(str(path), 1700000000 + 17 * i, 4242 + i, sheet, cmds, name="p%d" % i)
        logs.append(replay.read(str(path)))
    sim = BatchedArena(8, mode="Solo", level=2, auto_reset=False, player=sheet)
    try:
        status = replay.drive(sim, logs)
        h = sim.state_hash().cpu().numpy().view(np.uint64)
        cfg = sfcfg.make_config(arena_data, mode=sfcfg.MODE_SOLO, level_min=2, player=sheet)
        for e, l in enumerate(logs):
            o = sfo.Arena(cfg)
            o.reset(2, l.tb, l.serial)
            for t, c in enumerate(l.commands):
                st = o.step(bytes([c]))
                assert status[t, e] == st
                if st:
                    break
            assert h[e] == np.uint64(o.state_hash())
    finally:
        sim.close()


def _sampled_parity(torch, arena_data, n_envs, mode, level_min, level_max, steps, table, sample, player="account1",
                    squad_agents=False, obs_at=(), teams=None, obs_mask=1):
    """BASELINE.json-size batches: every arena steps on the GPU, a sample of them is followed by the
    oracle state for state (the whole-batch property: a checksum of all state hashes is finite work
    for the GPU only; the oracle cannot follow 10^5 arenas)."""
    from strikeforce_b200.sim import BatchedArena
    sim = BatchedArena(n_envs, mode=mode, level=level_min, level_max=level_max, squad_agents=squad_agents,
                       auto_reset=False, player=player, teams=teams)
    span = level_max - level_min + 1
    oracles = {}
    for e in sample:
        lvl = level_min + e % span
        cfg = sfcfg.make_config(arena_data, mode=mode, level_min=lvl, squad_agents=squad_agents, player=player, teams=teams)
        o = sfo.Arena(cfg)
        o.reset(lvl, common.synth_tb(e), common.synth_serial(e, 0))
        oracles[e] = o
    idx = torch.tensor(sample, device=sim.device)
    try:
        for t in range(steps):
            if t in obs_at:
                slots = [b for b in range(32) if (obs_mask >> b) & 1]
                obs = sim.observe(obs_mask)
                picked = obs[idx].reshape(len(sample), len(slots), -1).cpu().numpy()
                del obs
                for i, e in enumerate(sample):
                    for j, slot in enumerate(slots):
                        try:
                            ref = oracles[e].observe(slot)
                        except RuntimeError:  # that human died: its agent is gone (deleteAgent)
                            continue
                        assert (picked[i, j].view(np.uint32) == ref.view(np.uint32)).all(), \
                            "observation differs: env %d slot %d step %d" % (e, slot, t)
            act = sim.synth_actions(t, table)
            sim.step(act)
            h = sim.state_hash()[idx].cpu().numpy().view(np.uint64)
            st = sim.step_out()[idx, 0].cpu().numpy()
            ref_act = common.synth_actions(sample, sim.n_agents, t, table)
            for i, e in enumerate(sample):
                s = oracles[e].step(bytes(ref_act[i]))
                assert st[i] == s, "status differs: env %d step %d" % (e, t)
                if s in (sfcfg.RUNNING, sfcfg.DEAD, sfcfg.WIN, sfcfg.TIMEOUT):
                    assert h[i] == np.uint64(oracles[e].state_hash()), "state differs: env %d step %d" % (e, t)
        stats = sim.stats()
        assert stats["steps"] + stats["overflows"] + stats["ub_guards"] >= n_envs * steps - stats["episodes"] * steps
        return stats
    finally:
        sim.close()


import sfo  # noqa: E402  (oracle binding; test infrastructure)


def _trajectory_parity(torch, arena_data, n_envs, mode, level_min, level_max, steps, table, sample, player="account1",
                       squad_agents=False, max_steps=0, with_obs=True):
    """BASELINE.json-size batches at the survey's cadence: every arena of the batch plays
    `steps` env-steps of the synthetic workload on the GPU with auto-reset, and for every arena the hash of
    its canonical state after EVERY step is folded into a running checksum on the device
    (chk = mix64(chk ^ hash)).  The C oracle plays the sampled arenas on all host cores
    (sfo_run_trace) and must arrive at the same checksum, the same final state and, bit for bit, the
    same final observation: one mismatch anywhere in a trajectory changes its checksum."""
    from strikeforce_b200.sim import BatchedArena
    kw = dict(mode=mode, level_min=level_min, level_max=level_max, squad_agents=squad_agents, max_steps=max_steps,
              player=player)
    sim = BatchedArena(n_envs, mode=mode, level=level_min, level_max=level_max, squad_agents=squad_agents,
                       auto_reset=True, max_steps=max_steps, player=player)
    try:
        chk = torch.zeros(n_envs, dtype=torch.int64, device=sim.device)
        for t in range(steps):
            sim.step(sim.synth_actions(t, table))
            chk = common.torch_mix64(torch, chk ^ sim.state_hash())
        idx = torch.tensor(sample, device=sim.device)
        got_chk = chk[idx].cpu().numpy().view(np.uint64)
        got_last = sim.state_hash()[idx].cpu().numpy().view(np.uint64)
        obs = sim.observe(1)[idx].reshape(len(sample), -1).cpu().numpy().view(np.uint32) if with_obs else None
        stats = sim.stats()
        ref = common.oracle_traces(kw, sample, steps, table, with_obs)
        bad = [e for i, e in enumerate(sample) if np.uint64(ref[e][1]) != got_chk[i] or np.uint64(ref[e][2]) != got_last[i]]
        assert not bad, "%d of %d trajectories differ from the oracle, first: arena %d" % (len(bad), len(sample), bad[0])
        if with_obs:
            for i, e in enumerate(sample):
                if ref[e][3][0] != np.float32(-1.0).view(np.uint32):  # the player is alive: its agent sees
                    assert (obs[i] == ref[e][3]).all(), "final observation differs: arena %d" % e
        # the statistics the handle reduced on the device count every step of every arena
        assert stats["steps"] + stats["overflows"] + stats["ub_guards"] == n_envs * steps
        if len(sample) == n_envs:
            assert stats["episodes"] == sum(r[0] for r in ref.values())
        return stats
    finally:
        sim.close()


def test_config2_solo_4096_every_arena(torch_cuda, arena_data):
    """BASELINE.json configs[1]: Solo, 4,096 arenas on one GPU, random 9-symbol actions, 1,024 steps;
    EVERY arena's whole trajectory is compared with the CPU oracle (4,096 x 1,024)."""
    _trajectory_parity(torch_cuda, arena_data, 4096, sfcfg.MODE_SOLO, 1, 1, 1024, sfcfg.ACTIONS9, list(range(4096)))


def test_config2_solo_4096_first_steps_lockstep(torch_cuda, arena_data):
    """The same batch followed state for state (status, hash, observation) through its first steps."""
    _sampled_parity(torch_cuda, arena_data, 4096, sfcfg.MODE_SOLO, 1, 1, 24, sfcfg.ACTIONS9, list(range(0, 4096, 16)),
                    obs_at=(23,))


def test_config3_timer_65536_sampled(torch_cuda, arena_data):
    """configs[2]: Timer mode, 65,536 arenas, consumables / throwables sheet, levels 1-10, 28-symbol
    alphabet, observation tensors emitted on the device; the whole trajectories (768 steps) of 256 sampled
    arenas against the oracle, plus a lock-step walk with observations over the first steps."""
    rng = np.random.default_rng(3)
    sample = sorted(set(rng.integers(0, 65536, size=254).tolist()) | {0, 65535})
    _trajectory_parity(torch_cuda, arena_data, 65536, sfcfg.MODE_TIMER, 1, 10, 768, sfcfg.ACTIONS28, sample, player="synthetic")
    _sampled_parity(torch_cuda, arena_data, 65536, sfcfg.MODE_TIMER, 1, 10, 40, sfcfg.ACTIONS28, sample[:64],
                    player="synthetic", obs_at=(0, 39))


def test_config4_squad_shard_131072_sampled(torch_cuda, arena_data):
    """configs[3], the headline configuration: one GPU's shard (131,072 arenas) of the Squad 5v5 batch,
    blocks and portals in the alphabet, levels 1-10, episodes truncated at 2,048 steps as in bench.py;
    the whole 1,024-step trajectories of 512 sampled arenas and their final observations against the
    oracle -- the populations the bench times (tens of humans, zombies, built cells) are reached well
    inside that -- plus a lock-step walk over the first steps."""
    rng = np.random.default_rng(4)
    sample = sorted(set(rng.integers(0, 131072, size=510).tolist()) | {0, 131071})
    _trajectory_parity(torch_cuda, arena_data, 131072, sfcfg.MODE_SQUAD, 1, 10, 1024, sfcfg.ACTIONS28, sample, max_steps=2048)
    _sampled_parity(torch_cuda, arena_data, 131072, sfcfg.MODE_SQUAD, 1, 10, 30, sfcfg.ACTIONS28, sample[:48])


def test_config5_royale_32768_sampled(torch_cuda, arena_data):
    """configs[4] on the reference's map: Battle Royale placement (gameplay.hpp:1847-1859), 16 players
    in 4 teams, one GPU's shard (32,768 arenas) of the 262,144-arena batch; 96 sampled arenas followed
    by the oracle, observations of three players of every sampled arena compared bit for bit."""
    rng = np.random.default_rng(5)
    sample = sorted(set(rng.integers(0, 32768, size=94).tolist()) | {0, 32767})
    _sampled_parity(torch_cuda, arena_data, 32768, sfcfg.MODE_ROYALE, 1, 1, 80, sfcfg.ACTIONS28, sample,
                    teams=[1, 2, 3, 4] * 4, obs_at=(0, 79), obs_mask=(1 << 0) | (1 << 6) | (1 << 15))


def test_royale_policy_loop(torch_cuda, arena_data):
    """configs[4]: observations of all 16 players -> one batched AgentModel forward per tick -> the 16
    commands of every arena, without leaving the device."""
    from strikeforce_b200 import bots, policy
    from strikeforce_b200.sim import BatchedArena
    torch = torch_cuda
    torch.manual_seed(0)
    sim = BatchedArena(48, mode="Royale", teams=[1, 2, 3, 4] * 4, auto_reset=True, max_steps=20)
    try:
        agent = policy.PolicyAgent(policy.AgentModel(), 48 * 16, device=sim.device, seed=3)
        stats = bots.play(sim, bots.Custom(agent), 8)
        assert stats["steps"] + stats["overflows"] + stats["ub_guards"] == 48 * 8
    finally:
        sim.close()


def test_policy_loop_stays_on_the_device(torch_cuda, arena_data):
    """observe -> batched AgentModel forward -> sampled commands -> step, as gameplay::play() does
    with a bot (gameplay.hpp:956, 933), for the player and for agent-driven squad humans (P2)."""
    from strikeforce_b200 import bots, policy
    from strikeforce_b200.sim import BatchedArena
    torch = torch_cuda
    torch.manual_seed(0)
    for agents in (False, True):
        sim = BatchedArena(96, mode="Squad", level=2, squad_agents=agents, auto_reset=True, max_steps=30)
        try:
            rows = 96 * (9 if agents else 1)  # the largest batch a predict() call sees
            model = policy.AgentModel()

            class TwoAgents(bots.Custom):  # the player and the squad humans keep separate recurrent states
                def __init__(self):
                    super().__init__()
                    self.p1 = policy.PolicyAgent(model, 96, device=sim.device, seed=1)
                    self.p2 = policy.PolicyAgent(model, rows, device=sim.device, seed=2)

                def bot(self, s, agent_mask=1, phase=sfcfg.OBS_P1):
                    self.agent = self.p1 if phase == sfcfg.OBS_P1 else self.p2
                    return super().bot(s, agent_mask, phase)

                def agents(self):
                    return [self.p1, self.p2]

            c = TwoAgents()
            c.agent = c.p1
            stats = bots.play(sim, c, 12)
            assert stats["steps"] + stats["overflows"] + stats["ub_guards"] == 96 * 12
        finally:
            sim.close()


@pytest.mark.parametrize("seat_player", ["scripted", "device-agent"])
def test_match_host_steps_a_gpu_arena(torch_cuda, arena_data, seat_player):
    """a match hosted with the reference server's wire protocol
    (strikeforce_b200.match_server) whose arenas live on the GPU -- one per SEAT (sf_config.royale_ind):
    the copy of the match each seat's own client holds, kill credits and corpses included; two scripted
    socket clients and one host-played seat, every seat's arena compared with the oracle of that seat fed
    the relayed commands.  The host's own seat is played from a script or by a DEVICE AGENT: the seat's arena
    is observed from that player's position (sf_observe on the arena whose ind is the seat), the batched
    AgentModel picks the command, and the host sends it to the other clients like any player's byte."""
    import socket
    import threading
    import test_match_server as tms
    from strikeforce_b200 import match_server as ms
    from strikeforce_b200.sim import BatchedArena
    torch = torch_cuda
    teams, tb, serial, T = [1, 2, 1], 1700000321, 777, 40
    sheet = "p\n" + "\n".join(str(int(v)) for v in arena_data.player_sheet("account1"))
    acts = np.stack([common.synth_actions([9], 3, t, sfcfg.ACTIONS28)[0] for t in range(T)])
    listener = socket.socket()
    listener.bind(("127.0.0.1", 0))
    listener.listen(4)
    port = listener.getsockname()[1]
    host = ms.MatchHost(teams, "pw", tb, serial, local_seats={1: sheet})
    lobby = threading.Thread(target=host.accept, args=(listener,), daemon=True)
    lobby.start()
    clients = []
    for seat in (0, 2):
        c = tms.ScriptedClient(port, b"pw", sheet, [(int(acts[t, seat]), set()) for t in range(T)])
        c.start()
        clients.append(c)
        tms.wait_for_seat(host, seat)
    lobby.join(20)
    host.handshake()
    sims = [BatchedArena(1, mode="Royale", teams=teams, auto_reset=False, ind=s) for s in range(3)]
    sim = sims[0]
    oracles = []
    for s in range(3):
        oa = sfo.Arena(sfcfg.make_config(arena_data, mode=sfcfg.MODE_ROYALE, teams=teams, auto_reset=False, ind=s))
        oa.reset(1, tb, serial)
        oracles.append(oa)
    o = oracles[0]
    try:
        for sm in sims:
            sm.reset([0], [tb], [serial])
        tick = [0]

        def step(row):
            for sm, oa in zip(sims, oracles):
                sm.step(torch.tensor(list(row), dtype=torch.uint8, device=sm.device).view(1, 3))
                oa.step(row)
                assert np.uint64(sm.state_hash().cpu().numpy().view(np.uint64)[0]) == np.uint64(oa.state_hash())

        agent = None
        if seat_player == "device-agent":
            from strikeforce_b200 import policy as sfpolicy
            torch.manual_seed(5)
            agent = sfpolicy.PolicyAgent(sfpolicy.AgentModel(), 1, device=sims[1].device, seed=9, t_initial=2)
        played = []

        def policy(seat):
            tick[0] += 1
            if agent is None:
                return int(acts[tick[0] - 1, 1])
            obs = sims[seat].observe(1 << seat, channels_last=agent.channels_last)  # what this seat's own client would see
            cmd = int(sfcfg.ACTIONS9[int(agent.predict(obs.flatten(0, 1))[0])])
            played.append(cmd)
            return cmd

        winner, ticks = ms.host_match(host, step, policy, max_ticks=T)
        assert ticks == T and winner == 0
        if agent is not None:
            assert len(played) == T and played[:2] == [ord("+")] * 2 and all(c in sfcfg.ACTIONS9 for c in played)
            for c, pos in zip(clients, (0, 1)):  # the agent's commands reached seats 0 and 2 as the bytes of player 1
                c.join(10)
                assert [r[pos] for r in c.received] == played
        assert np.uint64(sim.state_hash()[0].item() & 0xFFFFFFFFFFFFFFFF) == np.uint64(o.state_hash())
        for c in clients:
            c.join(10)
            assert c.error is None and len(c.received) == T
    finally:
        for sm in sims:
            sm.close()
        host.close(), listener.close()


def test_observation_layouts_and_the_path_beyond_the_table(torch_cuda, arena_data, monkeypatch):
    """sf_observe's two memory layouts hold the same values bit for bit ([32][31][31] as the reference builds it,
    Custom.hpp:139, and SF_OBS_NHWC with the channel innermost), and so does a handle whose feature table has
    room for 2 dynamic cells only (SF_OBS_TABLE_ROWS): every other dynamic cell of a window then takes the
    describe-during-copy-out path that a full table reaches only in very crowded windows.  The plain layout
    of the first arenas is checked against the oracle as everywhere else."""
    from strikeforce_b200.sim import BatchedArena
    torch = torch_cuda
    n, mask = 96, 0b1011
    kw = dict(mode="Squad", level=3, squad_agents=True, auto_reset=False)
    sim = BatchedArena(n, **kw)
    monkeypatch.setenv("SF_OBS_TABLE_ROWS", "2")
    small = BatchedArena(n, **kw)
    monkeypatch.delenv("SF_OBS_TABLE_ROWS")
    oracles = common.make_oracles(arena_data, 6, sfcfg.MODE_SQUAD, 3, squad_agents=True)
    try:
        for t in range(120):
            act = sim.synth_actions(t, sfcfg.ACTIONS28)
            if t % 10 == 0:
                a = sim.observe(mask)
                views = [sim.observe(mask, channels_last=True), small.observe(mask), small.observe(mask, channels_last=True)]
                assert views[0].shape == a.shape and not views[0].is_contiguous()
                assert views[0].flatten(0, 1).is_contiguous(memory_format=torch.channels_last)
                for v in views:
                    assert torch.equal(a.view(torch.int32), v.contiguous().view(torch.int32)), "layouts differ at step %d" % t
                a_h = a.cpu().numpy()
                for e, o in enumerate(oracles):
                    for j, slot in enumerate((0, 1, 3)):
                        try:
                            ref = o.observe(slot)
                        except RuntimeError:
                            continue
                        assert (a_h[e, j].reshape(-1).view(np.uint32) == ref.view(np.uint32)).all()
            if t % 20 == 10:  # the P2 observation point, between the two halves of the step, in both layouts
                sim.step_a()
                small.step_a()
                p2 = sim.observe(mask & ~1, sfcfg.OBS_P2)
                for v in (sim.observe(mask & ~1, sfcfg.OBS_P2, channels_last=True), small.observe(mask & ~1, sfcfg.OBS_P2),
                          small.observe(mask & ~1, sfcfg.OBS_P2, channels_last=True)):
                    assert torch.equal(p2.view(torch.int32), v.contiguous().view(torch.int32)), "P2 layouts differ at step %d" % t
                sim.step_b(act)
                small.step_b(act)
            else:
                sim.step(act)
                small.step(act)
            act_h = act.cpu().numpy()
            for e, o in enumerate(oracles):
                o.step(bytes(act_h[e]))
    finally:
        sim.close()
        small.close()


def test_a_handle_runs_on_its_own_device(torch_cuda, arena_data):
    """A handle is bound to the device that was current in sf_create; every entry point switches to it and
    puts the caller's device back (SfDeviceGuard).  Needs two GPUs; the single-GPU case checks that the
    calls leave the current device alone."""
    from strikeforce_b200.sim import BatchedArena
    torch = torch_cuda
    n_dev = torch.cuda.device_count()
    torch.cuda.set_device(0)
    sim = BatchedArena(64, mode="Solo", level=1, auto_reset=False)
    oracles = common.make_oracles(arena_data, 64, sfcfg.MODE_SOLO, 1)
    try:
        other = 1 if n_dev > 1 else 0
        torch.cuda.set_device(other)  # the caller moves on to another GPU; the handle stays where it is
        for t in range(20):
            act = sim.synth_actions(t, sfcfg.ACTIONS28)
            sim.step(act)
            assert torch.cuda.current_device() == other
            act_h = act.cpu().numpy()
            for e, o in enumerate(oracles):
                o.step(bytes(act_h[e]))
        h = sim.state_hash().cpu().numpy().view(np.uint64)
        assert all(h[e] == np.uint64(o.state_hash()) for e, o in enumerate(oracles))
        if n_dev < 2:
            pytest.skip("one GPU: the cross-device half of this test did not run")
    finally:
        torch.cuda.set_device(0)
        sim.close()
