"""The C++ host side of the boundary (strikeforce_b200/host/bot-b200/{Agent,Custom}.hpp over the C ABI):
the reference's bot plugin surface, compiled against libtorch.

* `b200_play` (product): gameplay::play() for a batch with the template agent -- must leave the arenas in
  exactly the state the Python mirror (strikeforce_b200.bots.play) leaves them in.
* `oracle/_ref/host_policy_check` (test harness, built where the reference tree is): the same binding
  driven by the reference's OWN AgentModel (bots/bot-0.5/Modules.hpp, unmodified, one per arena) on the
  device tensors sf_observe fills; its probabilities (5e-6 absolute, fp32 with TF32 off: batched rows
  against single-sample tensors), its greedy commands and the arenas' states must equal those of
  strikeforce_b200.policy / bots on the same arenas.  That closes the loop reference network -> device
  observations -> commands -> tick on the GPU."""
import json
import os
import subprocess

import numpy as np
import pytest

from strikeforce_b200 import config as sfcfg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PLAY = os.path.join(ROOT, "strikeforce_b200", "host", "b200_play")
CHECK = os.path.join(ROOT, "oracle", "_ref", "host_policy_check")


def test_host_sources_and_blob_layout(arena_data, tmp_path):
    """CPU part: the headers are there, and the configuration blob is the struct + the two map arrays."""
    for f in ("bot-b200/Agent.hpp", "bot-b200/Custom.hpp", "b200_play.cpp", "build.sh"):
        assert os.path.exists(os.path.join(ROOT, "strikeforce_b200", "host", f))
    cfg = sfcfg.make_config(arena_data, n_envs=5, mode=sfcfg.MODE_SQUAD, level_min=2, squad_agents=True)
    path = tmp_path / "cfg.bin"
    sfcfg.dump_config(cfg, str(path))
    raw = path.read_bytes()
    import ctypes as C
    assert len(raw) == C.sizeof(sfcfg.SfConfig) + sfcfg.CELLS * 3
    back = sfcfg.SfConfig.from_buffer_copy(raw[:C.sizeof(sfcfg.SfConfig)])
    assert (back.n_envs, back.mode, back.level_min, back.squad_agents, back.abi_version) == (5, sfcfg.MODE_SQUAD, 2, 1, cfg.abi_version)
    cells = np.frombuffer(raw, dtype=np.uint8, count=sfcfg.CELLS, offset=C.sizeof(sfcfg.SfConfig))
    assert set(bytes(cells)) <= set(b"#.^vO") and (cells == np.asarray(arena_data.map_cells).reshape(-1)).all()


@pytest.mark.gpu
def test_cpp_host_plays_like_the_python_mirror(arena_data, tmp_path):
    import torch
    assert torch.cuda.is_available() and os.path.exists(PLAY), "needs a CUDA device and the built host (strikeforce_b200/host/build.sh)"
    from strikeforce_b200 import bots
    from strikeforce_b200.sim import BatchedArena
    n, ticks = 96, 40
    cfg = sfcfg.make_config(arena_data, n_envs=n, mode=sfcfg.MODE_SQUAD, level_min=1, level_max=4, auto_reset=True, max_steps=25)
    blob = tmp_path / "cfg.bin"
    sfcfg.dump_config(cfg, str(blob))
    out = subprocess.run([PLAY, str(blob), str(ticks), "idle"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    got = json.loads(out.stdout.strip().splitlines()[-1])
    sim = BatchedArena(n, mode="Squad", level=1, level_max=4, auto_reset=True, max_steps=25)
    try:
        stats = bots.play(sim, bots.Custom(bots.Agent()), ticks)
        x = 0
        for h in sim.state_hash().cpu().numpy().view(np.uint64).tolist():
            x ^= h
        assert got["steps"] == stats["steps"] == n * ticks and got["episodes"] == stats["episodes"] > 0
        assert got["rng_draws"] == stats["rng_draws"] and got["kills"] == stats["kills"]
        assert got["hash_xor"] == x, "the arenas hosted from C++ differ from those hosted from Python"
    finally:
        sim.close()


@pytest.mark.gpu
@pytest.mark.parametrize("layout", ["nchw", "nhwc"])
def test_reference_network_drives_the_binding(arena_data, tmp_path, layout):
    """... `nhwc`: the binding asks sf_observe for channel-innermost observations (SF_OBS_NHWC) and hands the
    reference's network the same [B,32,31,31] tensor with channels-last strides: same commands, same arenas."""
    import torch
    if not os.path.exists(CHECK):
        pytest.skip("oracle/_ref/host_policy_check is built where the reference tree is (oracle/ref_harness/build_host_check.sh)")
    assert torch.cuda.is_available()
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    import test_policy_model as tpm
    from strikeforce_b200 import policy
    from strikeforce_b200.sim import BatchedArena
    n, ticks = 6, 10
    cfg = sfcfg.make_config(arena_data, n_envs=n, mode=sfcfg.MODE_SOLO, level_min=1, level_max=3, auto_reset=False, env_id_base=40)
    blob, res = tmp_path / "cfg.bin", tmp_path / "res.bin"
    sfcfg.dump_config(cfg, str(blob))
    out = subprocess.run([CHECK, str(blob), str(ticks), str(res)] + (["nhwc"] if layout == "nhwc" else []),
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr
    raw = res.read_bytes()
    per = n * 9 * 4 + n + n * 8
    assert len(raw) == per * ticks
    sim = BatchedArena(n, mode="Solo", level=1, level_max=3, auto_reset=False, env_id_base=40)
    try:
        model = tpm.formula_model().to(sim.device)
        state = model.initial_state(n, sim.device)
        table = torch.tensor(list(sfcfg.ACTIONS9), dtype=torch.uint8, device=sim.device)
        worst = 0.0
        for t in range(ticks):
            off = per * t
            p_ref = np.frombuffer(raw, dtype=np.float32, count=n * 9, offset=off).reshape(n, 9)
            a_ref = np.frombuffer(raw, dtype=np.uint8, count=n, offset=off + n * 36)
            h_ref = np.frombuffer(raw, dtype=np.uint64, count=n, offset=off + n * 36 + n)
            obs = sim.observe(1).view(n, sfcfg.OBS_CH, sfcfg.OBS_WIN, sfcfg.OBS_WIN)
            with torch.no_grad():
                p, _, state = model(obs, state)
            worst = max(worst, float(np.abs(p.cpu().numpy() - p_ref).max()))
            act = p.argmax(1)
            state = policy.AgentModel.with_action(state, act)
            sym = table[act]
            assert bytes(sym.cpu().numpy()) == bytes(a_ref), "tick %d: commands differ (largest probability gap so far %g)" % (t, worst)
            sim.step(sym.view(n, 1))
            assert (sim.state_hash().cpu().numpy().view(np.uint64) == h_ref).all(), "tick %d: arenas differ" % t
        assert worst < 5e-6, "probabilities differ from the reference network's by %g" % worst
    finally:
        sim.close()
