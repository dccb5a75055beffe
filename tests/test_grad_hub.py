"""strikeforce_b200.grad_hub against libtorch's own AdamW run around the
reference's AgentModel the way AgentServer::aggregate_and_update does it
(bots/bot-0.5/AgentServer.cpp:465-524; oracle/ref_harness/hub_oracle.cpp ->
tests/golden/hub_golden.bin): two gloo ranks play the two clients.  Tolerance 2e-7 absolute on the
update vector (entries are ~1e-3: the Python AdamW forms the first moment with lerp_, libtorch with
mul_/add_, one rounding apart)."""
import os
import socket
import struct
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden", "hub_golden.bin")
HIDDEN, ROUNDS, TOL = 8, 3, 2e-7


def hash_unit(i, k):
    u = (i.astype(np.uint64) * 2654435761 + k * 40503 + 12345) & 0xFFFFFFFF
    u ^= u >> 15
    u = (u * 2246822519) & 0xFFFFFFFF
    u ^= u >> 13
    return ((u >> 8).astype(np.float32) / np.float32(16777216.0)).astype(np.float32)


def formula(shape, k, scale):
    n = int(np.prod(shape))
    w = ((hash_unit(np.arange(n, dtype=np.uint64), k) - np.float32(0.5)) * np.float32(scale)).astype(np.float32)
    return torch.from_numpy(w).view(shape)


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


def _client(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from strikeforce_b200.grad_hub import GradientHub
    from strikeforce_b200.policy import AgentModel
    model = AgentModel(hidden=HIDDEN)
    params = list(model.parameters())
    with torch.no_grad():
        for k, p in enumerate(params):
            p.copy_(formula(p.shape, k, 0.16))
    hub = GradientHub(params, lr=1e-3)
    updates = []
    for r in range(ROUNDS):
        grads = [formula(p.shape, 5000 + 1000 * r + 100 * rank + i, 0.02) for i, p in enumerate(params)]
        updates.append([u.numpy().copy() for u in hub.step(grads)])
    idle = hub.step(contributing=False)  # a round in which no client has gradients changes nothing
    out.put((rank, updates, idle is None, hub.version, [p.detach().numpy().copy() for p in params]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_reproduce_the_reference_hub():
    g = read_golden()
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_client, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((out.get(timeout=240) for _ in range(2)), key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    worst = 0.0
    for rank, updates, idle_none, version, params in res:
        assert idle_none and version == ROUNDS
        for r in range(ROUNDS):
            for i, u in enumerate(updates[r]):
                ref = g["u:%d:%d" % (r, i)]
                assert u.shape == ref.shape
                worst = max(worst, float(np.abs(u - ref).max()))
    assert worst < TOL, "update vectors differ from libtorch's by %g" % worst
    # the replicas of the server's model stay bit-identical on every rank
    for a, b in zip(res[0][4], res[1][4]):
        assert (a.view(np.uint32) == b.view(np.uint32)).all()


def test_one_contributor_counts_once():
    """The mean is over the CONTRIBUTING clients (:470-474, 489), single process."""
    sys.path.insert(0, ROOT)
    from strikeforce_b200.grad_hub import GradientHub
    w = torch.nn.Parameter(torch.tensor([1.0, -2.0, 3.0]))
    hub = GradientHub([w], lr=1e-3)
    u = hub.step([torch.tensor([0.5, 0.5, -0.5])])
    ref = torch.nn.Parameter(torch.tensor([1.0, -2.0, 3.0]))
    opt = torch.optim.AdamW([ref], lr=1e-3, foreach=False)
    ref.grad = torch.tensor([0.5, 0.5, -0.5])
    opt.step()
    assert torch.equal(w.detach(), ref.detach()) and torch.allclose(u[0], ref.detach() - torch.tensor([1.0, -2.0, 3.0]))
    v = torch.nn.Parameter(torch.tensor([1.0, -2.0, 3.0]))
    GradientHub.apply_update([v], u)
    assert torch.allclose(v.detach(), w.detach())
