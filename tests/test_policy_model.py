"""the batched policy network (strikeforce_b200/policy.py) against outputs of the
reference's OWN AgentModel (bots/bot-0.5/Modules.hpp, run with libtorch by
oracle/ref_harness/policy_oracle.cpp; fixture tests/golden/policy_golden.bin).  Parameters and
observations come from a closed integer-hash formula on both sides, so the fixture holds only the
network's outputs.  Tolerance 2e-6 absolute on probabilities / values (fp32 reductions in a
different order: batched rows instead of single-sample tensors)."""
import os
import struct

import numpy as np
import pytest
import torch

from strikeforce_b200.policy import AgentModel, PolicyAgent

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "policy_golden.bin")
TOL = 2e-6


def hash_unit(i, k):
    u = (i.astype(np.uint64) * 2654435761 + k * 40503 + 12345) & 0xFFFFFFFF
    u ^= u >> 15
    u = (u * 2246822519) & 0xFFFFFFFF
    u ^= u >> 13
    return ((u >> 8).astype(np.float32) / np.float32(16777216.0)).astype(np.float32)


def read_golden():
    out, pos = {}, 0
    data = open(GOLDEN, "rb").read()
    while pos < len(data):
        n = struct.unpack_from("<i", data, pos)[0]
        name = data[pos + 4:pos + 4 + n].decode()
        pos += 4 + n
        nd = struct.unpack_from("<i", data, pos)[0]
        shape = struct.unpack_from("<%dq" % nd, data, pos + 4)
        pos += 4 + 8 * nd
        cnt = int(np.prod(shape)) if nd else 1
        out[name] = np.frombuffer(data, dtype=np.float32, count=cnt, offset=pos).reshape(shape).copy()
        pos += 4 * cnt
    return out


def formula_model():
    m = AgentModel()
    with torch.no_grad():
        for k, (_, p) in enumerate(m.named_parameters()):  # same registration order as the reference
            i = np.arange(p.numel(), dtype=np.uint64)
            w = ((hash_unit(i, k) - np.float32(0.5)) * np.float32(0.16)).astype(np.float32)
            p.copy_(torch.from_numpy(w).view(p.shape))
    return m.eval()


def formula_obs(s):
    i = np.arange(32 * 31 * 31, dtype=np.uint64)
    x = np.where(hash_unit(i, 1000 + s) < np.float32(0.08), hash_unit(i, 2000 + s) * np.float32(1.5), np.float32(0))
    return torch.from_numpy(x.astype(np.float32)).view(1, 32, 31, 31)


def test_parameter_names_match_the_reference():
    names = [n for n, _ in AgentModel().named_parameters()]
    assert names[:4] == ["backbone.cnn.conv0.weight", "backbone.cnn.conv1.weight", "backbone.cnn.conv2.weight",
                         "backbone.cnn.conv3.weight"]
    assert "backbone.combined_processor.0.weight" in names and "policy.1.bias" in names and len(names) == 30


def test_batched_forward_matches_the_reference_network():
    g = read_golden()
    steps = len([k for k in g if k.startswith("p:")])
    m = formula_model()
    # the golden sequence is row 1 of a batch of three; rows 0 and 2 see other inputs
    B = 3
    st = m.initial_state(B)
    with torch.no_grad():
        for s in range(steps):
            x = torch.cat([formula_obs(s + 7), formula_obs(s), formula_obs(s + 13)])
            p, v, st = m(x, st)
            assert np.abs(p[1].numpy() - g["p:%d" % s]).max() < TOL, "policy, step %d" % s
            assert np.abs(v[1].numpy() - g["v:%d" % s]).max() < TOL, "value, step %d" % s
            assert abs(float(p[1].sum()) - 1.0) < 1e-5
            act = torch.tensor([(s + 2) % 9, (s * 4 + 1) % 9, (s + 5) % 9])  # row 1: the action the oracle fed back
            st = AgentModel.with_action(st, act)


def test_gru_step_is_torch_gru_and_channels_last_is_the_same_network():
    """The recurrent layers run as two matrix products and the gate formula instead of nn.GRU's sequence
    kernel (policy._gru_step): same parameters, same result as the module.  The channel-innermost copy of
    the network (observations written with SF_OBS_NHWC) gives the same probabilities and values."""
    from strikeforce_b200 import policy
    torch.manual_seed(3)
    m = formula_model()
    with torch.no_grad():
        for g in (m.backbone.gru0, m.backbone.gru1):
            x, h = torch.rand(7, 160) - 0.5, torch.rand(7, 160) - 0.5
            _, hn = g(x.view(1, 7, 160), h.view(1, 7, 160))
            assert (hn.view(7, 160) - policy._gru_step(g, x, h)).abs().max() < 1e-6
        x = torch.cat([formula_obs(s) for s in range(4)])
        st = m.initial_state(4)
        p, v, _ = m(x, st)
        buf = x.permute(0, 2, 3, 1).contiguous()  # the memory sf_observe writes with SF_OBS_NHWC
        m2 = formula_model().to(memory_format=torch.channels_last)
        p2, v2, _ = m2(buf.permute(0, 3, 1, 2), st)
        assert (p - p2).abs().max() < TOL and (v - v2).abs().max() < TOL


def test_policy_agent_is_seeded_and_batched():
    m = formula_model()
    a1, a2 = PolicyAgent(m, 4, device="cpu", seed=3, t_initial=0), PolicyAgent(m, 4, device="cpu", seed=3, t_initial=0)
    x = torch.cat([formula_obs(s) for s in range(4)])
    for _ in range(3):
        r1, r2 = a1.predict(x), a2.predict(x)
        assert r1.shape == (4,) and torch.equal(r1, r2) and int(r1.max()) < 9


def test_chunked_predict_is_the_same_policy():
    """PolicyAgent(chunk=...) bounds the activations of very large batches (configs[4]: 524,288
    observations per tick); rows are independent, so chunking changes neither the sampled actions
    (same generator order) nor the recurrent state."""
    import torch
    from strikeforce_b200 import policy
    torch.manual_seed(0)
    m = policy.AgentModel()
    a = policy.PolicyAgent(m, 10, device="cpu", seed=5, t_initial=1)
    b = policy.PolicyAgent(m, 10, device="cpu", seed=5, chunk=4, t_initial=1)
    x = torch.rand(10, 32, 31, 31)
    for _ in range(3):
        assert a.predict(x).tolist() == b.predict(x).tolist()
        for s, t in zip(a.state, b.state):
            assert torch.allclose(s, t, atol=1e-6)


def test_policy_agent_follows_the_reference_agent_around_the_network():
    """Agent::predict (bots/bot-0.5/Agent.hpp:178-215): action 0 for the first T_initial calls without
    touching the network's memory; without SLOWMOTION action 0 is drawn half of the time; a new
    game (reset_rows) starts both over for the rows it names."""
    torch.manual_seed(1)
    m = formula_model()
    x = torch.cat([formula_obs(s) for s in range(6)])
    a = PolicyAgent(m, 6, device="cpu", seed=7)
    s0 = [t.clone() for t in a.state]
    for _ in range(10):  # T_initial = 10 (Agent.hpp:274)
        assert a.predict(x).tolist() == [0] * 6
        assert all(torch.equal(u, v) for u, v in zip(a.state, s0)), "the warm-up must not run the recurrent network"
    acts = torch.stack([a.predict(x) for _ in range(12)])
    assert int(acts.max()) > 0 and not all(torch.equal(u, v) for u, v in zip(a.state, s0))
    # a new game for rows 1 and 4 only
    before = [t.clone() for t in a.state]
    mask = torch.tensor([False, True, False, False, True, False])
    a.reset_rows(mask)
    for t, b0, z in zip(a.state, before, s0):
        assert torch.equal(t[mask], z[mask]) and torch.equal(t[~mask], b0[~mask])
    nxt = a.predict(x)
    assert nxt[mask].tolist() == [0, 0] and a.calls.tolist() == [10, 1, 10, 10, 1, 10]
    # the build without SLOWMOTION: P(action 0) = 0.5 whatever the network says
    b = PolicyAgent(m, 6, device="cpu", seed=11, t_initial=0, slowmotion=False)
    draws = torch.stack([b.predict(x) for _ in range(400)])
    frac0 = float((draws == 0).float().mean())
    assert 0.44 < frac0 < 0.56, frac0


@pytest.mark.gpu
def test_cuda_forward_matches_the_reference_network():
    """The forward the policy loop actually runs -- batched, on the B200, cuDNN / cuBLAS kernels --
    against the reference's own AgentModel (policy_golden.bin).  fp32 with TF32 off; tolerance 5e-6
    absolute on probabilities and values (the reduction orders of the GPU kernels differ from the
    reference's single-sample CPU kernels).  Row 1 of a batch of 64 replays the golden sequence, the
    other rows see other inputs: rows must not leak into each other."""
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    g = read_golden()
    steps = len([k for k in g if k.startswith("p:")])
    dev = torch.device("cuda", 0)
    m = formula_model().to(dev)
    B = 64
    st = m.initial_state(B, dev)
    worst = 0.0
    with torch.no_grad():
        for s in range(steps):
            x = torch.cat([formula_obs(s + 3 * r) if r != 1 else formula_obs(s) for r in range(B)]).to(dev)
            p, v, st = m(x, st)
            worst = max(worst, float(np.abs(p[1].cpu().numpy() - g["p:%d" % s]).max()),
                        float(np.abs(v[1].cpu().numpy() - g["v:%d" % s]).max()))
            act = torch.tensor([(s + r) % 9 if r != 1 else (s * 4 + 1) % 9 for r in range(B)], device=dev)
            st = AgentModel.with_action(st, act)
    assert worst < 5e-6, "CUDA forward differs from the reference network by %g" % worst
