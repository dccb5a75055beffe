"""Batched twin of the reference's policy network.

``AgentModel`` of ``bots/bot-0.5/Modules.hpp`` (lines 29-180) -- a 4-layer stride-2 CNN
(31 -> 15 -> 7 -> 3 -> 1, no bias), two single-layer GRUs, a 5-cell "point of view" gather, a
linear mixer and two residual heads -- evaluates ONE observation per call and keeps the GRU state
and the last action inside the module (``reset_memory`` / ``update_actions``).  Here the same
arithmetic runs on a batch ``[B, 32, 31, 31]`` that never leaves the device, with the recurrent
state as explicit ``[B, 160]`` tensors; every "divide by the mean absolute value" of the
reference (a reduction over the whole tensor of its single sample) is a per-sample reduction.
The layers are library calls (cuDNN / cuBLAS through torch), exactly as the reference's are
libtorch calls; parameter names and shapes match the reference's ``named_parameters()`` so that a
``model.pt`` state can be exchanged.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


def _l1norm(x, size):
    """x * size / (sum |x| + 1e-8) per sample (Modules.hpp:45, 109, 113, ...)."""
    flat = x.reshape(x.shape[0], -1)
    return x * (size / (flat.abs().sum(dim=1) + 1e-8)).view(-1, *([1] * (x.dim() - 1)))


def _gru_step(gru, x, h):
    """One time step of a one-layer ``nn.GRU`` (Modules.hpp:88, 92 call it with a sequence of one) as two
    matrix products over the batch and the gate arithmetic of torch's GRU (gates in the order r, z, n;
    n = tanh(W_in x + b_in + r * (W_hn h + b_hn)); h' = (1 - z) * n + z * h).  The module keeps its
    parameters -- same names, same checkpoints -- but is not called: cuDNN's persistent-RNN kernel, which
    nn.GRU dispatches to, takes 10 ms per call for a batch of 32,768 rows with sequence length 1 (61% of the
    whole forward, profiles/r02_policy_forward.txt); this is 0.3 ms."""
    gi = F.linear(x, gru.weight_ih_l0, gru.bias_ih_l0)
    gh = F.linear(h, gru.weight_hh_l0, gru.bias_hh_l0)
    i_r, i_z, i_n = gi.chunk(3, dim=1)
    h_r, h_z, h_n = gh.chunk(3, dim=1)
    r = torch.sigmoid(i_r + h_r)
    z = torch.sigmoid(i_z + h_z)
    n = torch.tanh(i_n + r * h_n)
    return n + z * (h - n)


class ResB(nn.Module):
    """Modules.hpp:29-52"""

    def __init__(self, hidden, layers):
        super().__init__()
        for i in range(layers):
            setattr(self, "lin%d" % i, nn.Linear(hidden, hidden))
        self.n = layers

    def forward(self, x):
        x = _l1norm(x, x[0].numel())
        for i in range(self.n):
            y = F.relu(getattr(self, "lin%d" % i)(x)) + x
            x = _l1norm(y, y[0].numel())
        return x


class GameCNN(nn.Module):
    """Modules.hpp:54-74"""

    def __init__(self, channels, d_out, layers):
        super().__init__()
        for i in range(layers):
            setattr(self, "conv%d" % i, nn.Conv2d(channels if i == 0 else d_out, d_out, 3, stride=2, padding=0, bias=False))
        self.n = layers

    def forward(self, x):
        for i in range(self.n):
            x = getattr(self, "conv%d" % i)(x)
        return x


class Backbone(nn.Module):
    """Modules.hpp:76-136; the recurrent state is passed in and returned instead of being kept."""

    def __init__(self, channels=32, grid=31, hidden=160, actions=9):
        super().__init__()
        self.channels, self.grid, self.hidden, self.actions = channels, grid, hidden, actions
        self.cnn = GameCNN(channels, hidden, 4)
        self.gru0 = nn.GRU(hidden, hidden, num_layers=1)
        self.combined_processor = nn.Sequential(nn.Linear(2 * hidden + actions, hidden))
        self.gru1 = nn.GRU(hidden, hidden, num_layers=1)

    def initial_state(self, batch, device=None):
        """reset_memory(), Modules.hpp:94-99: zero GRU states and the action one-hot on index 0."""
        a = torch.zeros(batch, self.actions, device=device)
        a[:, 0] = 1
        return torch.zeros(batch, self.hidden, device=device), torch.zeros(batch, self.hidden, device=device), a

    def forward(self, x, h0, h1, action_input):
        B, H = x.shape[0], self.hidden
        feat = _l1norm(self.cnn(x), H).reshape(B, H)        # [B, H, 1, 1] -> [B, H]
        h0n = _gru_step(self.gru0, feat, h0)                # a sequence of one: the output IS the new state
        out_seq = _l1norm(h0n, H)
        c = self.grid // 2                                   # the five cells around the agent, :115-121
        cells = [(-1, 0), (0, -1), (0, 0), (0, 1), (1, 0)]
        pov = torch.cat([x[:, :, c + dr, c + dc] for dr, dc in cells] + [action_input], dim=1)
        combined = torch.cat([out_seq + feat, _l1norm(pov, H)], dim=1)
        gated = _l1norm(self.combined_processor(combined), H)
        h1n = _gru_step(self.gru1, gated, h1)
        out = _l1norm(h1n, H) + gated
        return out, h0n, h1n


class AgentModel(nn.Module):
    """Modules.hpp:138-180: ``forward`` returns (p [B, 9], v [B], new state)."""

    def __init__(self, channels=32, grid=31, hidden=160, actions=9, head_layers=3):
        super().__init__()
        self.backbone = Backbone(channels, grid, hidden, actions)
        self.value = nn.Sequential(ResB(hidden, head_layers), nn.Linear(hidden, 1))
        self.policy = nn.Sequential(ResB(hidden, head_layers), nn.Linear(hidden, actions))

    def initial_state(self, batch, device=None):
        return self.backbone.initial_state(batch, device)

    def forward(self, x, state):
        h0, h1, a = state
        gated, h0, h1 = self.backbone(x, h0, h1, a)
        p = torch.softmax(self.policy(gated), dim=-1) + 1e-8
        v = torch.sigmoid(self.value(gated)).view(-1)
        return p, v, (h0, h1, a)

    @staticmethod
    def with_action(state, actions):
        """update_actions(one_hot), Modules.hpp:101-103: the chosen action is fed back next call."""
        h0, h1, a = state
        return h0, h1, F.one_hot(actions, a.shape[1]).to(a.dtype)


class PolicyAgent:
    """``Agent`` of bots/bot-0.5/Agent.hpp:178-224 for a batch: ``predict`` runs the network on the
    observations of ``sf_observe`` and samples an action per row with a seeded device generator
    (the reference samples with std::random_device, which cannot be reproduced).

    What the reference's ``predict`` does around the network is kept, per row:

    * the first ``t_initial`` (= ``T_initial`` = 10, Agent.hpp:274) calls of an episode return action 0
      without running the model (:179-180; ``cnt`` counts the calls of an ``Agent``, and a game creates a
      fresh one);
    * ``slowmotion=True`` is the shipped build (macros.hpp:18 defines SLOWMOTION): the action is drawn
      from the network's distribution as it is.  ``slowmotion=False`` is the build without that macro
      (:204-208): action 0 gets probability 0.5 and the others are scaled by ``0.5 / (1 - p0 + 1e-5)``;
    * ``reset_rows(mask)`` is the new ``Agent`` of the next game (``reset_memory``, Modules.hpp:94-99):
      zero GRU states, the action one-hot back on index 0, the call counter back to 0 -- call it with
      the arenas whose ``step_out`` status is terminal (``bots.play`` does);
    * ``channels_last=True`` (the default) keeps the convolution weights channel-innermost, to go with
      observations written in that layout (``BatchedArena.observe(channels_last=True)``; ``bots.Custom`` asks for
      the layout its agent wants): same values, no transposes inside the convolution library
      (profiles/r02_policy_forward.txt).  Observations in the reference's [32][31][31] order are accepted either way."""

    def __init__(self, model: AgentModel, batch, device="cuda", seed=0, training=False, chunk=0, t_initial=10,
                 slowmotion=True, channels_last=True):
        self.model = model.to(device).eval()
        self.channels_last = channels_last  # bots.Custom asks sf_observe for the matching layout
        if channels_last:  # for observations written with SF_OBS_NHWC: the convolutions then run without a transpose
            self.model = self.model.to(memory_format=torch.channels_last)
        self.state = model.initial_state(batch, device)
        self.gen = torch.Generator(device=device)
        self.gen.manual_seed(seed)
        self.training = training
        self.chunk = chunk  # >0: rows per forward call (bounds the activations of very large batches)
        self.t_initial = t_initial
        self.slowmotion = slowmotion
        self.calls = torch.zeros(batch, dtype=torch.int32, device=device)  # Agent::cnt per row

    def _sample(self, p):
        if not self.slowmotion:  # Agent.hpp:204-208
            p = p.clone()
            p[:, 1:] *= (0.5 / (1 - p[:, 0] + 1e-5)).unsqueeze(1)
            p[:, 0] = 0.5
        return torch.multinomial(p, 1, generator=self.gen).view(-1)

    @torch.no_grad()
    def predict(self, obs):
        B = obs.shape[0]
        # rows still inside their first t_initial calls answer 0 and leave the network's memory alone
        live = self.calls >= self.t_initial
        self.calls += (~live).to(self.calls.dtype)
        act = torch.zeros(B, dtype=torch.int64, device=obs.device)
        step = self.chunk if self.chunk and B > self.chunk else B
        for lo in range(0, B, step):
            hi = min(B, lo + step)
            sub = tuple(s[lo:hi] for s in self.state)
            p, _, st = self.model(obs[lo:hi], sub)
            a = torch.where(live[lo:hi], self._sample(p), act[lo:hi])
            keep = live[lo:hi].unsqueeze(1)
            for dst, new, old in zip(self.state, AgentModel.with_action(st, a), sub):
                dst[lo:hi] = torch.where(keep, new, old)
            act[lo:hi] = a
        return act

    def reset_rows(self, mask):
        """A new game for the rows of ``mask`` (bool [batch]): what constructing a new ``Agent`` does."""
        h0, h1, a = self.state
        m = mask.to(h0.device).unsqueeze(1)
        h0.masked_fill_(m, 0), h1.masked_fill_(m, 0), a.masked_fill_(m, 0)
        a[:, 0] = torch.where(m.view(-1), torch.ones_like(a[:, 0]), a[:, 0])
        self.calls.masked_fill_(m.view(-1), 0)

    def update(self, actions, imitate):
        return None

    def in_training(self):
        return self.training

    def is_manual(self):
        return False
