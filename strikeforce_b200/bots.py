"""Host-side mirror of the reference's bot plugin surface, batched.

The reference selects a bot folder at compile time (``selected_agent.hpp:25``,
``selected_custom.hpp:25``); the folder defines ``class Agent`` and three ``gameplay::`` member
functions (``prepare``, ``bot``, ``view``; gameplay.hpp:477-485, README.md:288-300).  Here the
same names take whole batches:

* ``Agent.predict(obs)``: ``obs`` is the fp32 device tensor ``[B, 32, 31, 31]`` built by
  ``sf_observe`` (one row per arena and driven human) instead of one ``std::vector<float>`` of
  30,752 values (bots/bot-0.5/Agent.hpp:178); it returns int64 indices ``[B]`` into the action
  string, as ``int predict(const std::vector<float>&)`` does for one agent;
* ``Agent.update(actions, imitate)`` mirrors ``void update(int action, bool imitate)``
  (gameplay.hpp:975, 998); ``in_training`` / ``is_manual`` as in bots/bot-0/Agent.hpp:27-37;
* ``Custom.prepare / bot / view`` mirror bots/bot-0.5/Custom.hpp:137-168: ``bot`` = observe ->
  predict -> table lookup, all on the device.

``play`` is the loop of ``gameplay::play()`` (gameplay.hpp:1443-1472) for a batch: P1 observation
and the player's command at the loop top (``get_my_action``, :956); with agent-driven squad
humans the step is split so that they observe at P2 (``get_command``, :933); in Battle Royale all
players observe at the loop top, each from its own position (BASELINE.json configs[4]).
"""
from __future__ import annotations

import torch

from . import config as sfcfg


class Agent:
    """Template agent (bots/bot-0/Agent.hpp:27-37): always action 0 ('+')."""

    def __init__(self, training=True):
        self.training = training

    def predict(self, obs: torch.Tensor) -> torch.Tensor:
        return torch.zeros(obs.shape[0], dtype=torch.int64, device=obs.device)

    def update(self, actions: torch.Tensor, imitate) -> None:
        return None

    def in_training(self) -> bool:
        return self.training

    def is_manual(self) -> bool:
        return False


class RandomAgent(Agent):
    """Uniform random policy with a seeded device generator (the reference samples with
    std::random_device, bots/bot-0.5/Agent.hpp:212, which is not reproducible)."""

    def __init__(self, n_actions=9, seed=0, device="cuda", training=False):
        super().__init__(training)
        self.n_actions = n_actions
        self.gen = torch.Generator(device=device)
        self.gen.manual_seed(seed)

    def predict(self, obs):
        return torch.randint(0, self.n_actions, (obs.shape[0],), generator=self.gen, device=obs.device)


class Custom:
    """bots/bot-0.5/Custom.hpp: ``prepare`` fixes the action table and creates the agent,
    ``bot`` turns observations into command symbols, ``view`` is a no-op hook."""

    action = sfcfg.ACTIONS9  # gameplay::action = "+xzqeawsd", Custom.hpp:162

    def __init__(self, agent: Agent | None = None, channels_last=None):
        self.agent = agent
        # the layout sf_observe writes (SF_OBS_NHWC; shape and values are the same): None = what the agent asks for
        self._channels_last = channels_last
        self._table = None

    @property
    def channels_last(self):
        if self._channels_last is not None:
            return self._channels_last
        return bool(getattr(self.agent, "channels_last", False))

    def prepare(self, sim):
        if self.agent is None:
            self.agent = Agent()
        self._table = torch.tensor(list(self.action), dtype=torch.uint8, device=sim.device)
        return self

    def bot(self, sim, agent_mask=1, phase=sfcfg.OBS_P1):
        """uint8 device tensor [n_envs, n_selected] of command symbols for the selected humans."""
        obs = sim.observe(agent_mask, phase, channels_last=self.channels_last)
        n_envs, nsel = obs.shape[0], obs.shape[1]
        idx = self.agent.predict(obs.flatten(0, 1))
        self.agent.update(idx, False)
        return self._table[idx].view(n_envs, nsel)

    def view(self, sim):
        return None

    def agents(self):
        """every agent object this Custom drives (one by default)"""
        return [self.agent]

    def new_games(self, sim, done):
        """``done``: bool device tensor [n_envs], arenas whose episode ended in the last step.  The
        reference plays one game per process and builds a new ``Agent`` for the next one
        (``prepare``, Custom.hpp:161-165); with auto-reset the batch equivalent is to reset the rows
        of those arenas (every driven human of an arena is a run of consecutive rows)."""
        for a in self.agents():
            reset = getattr(a, "reset_rows", None)
            rows = getattr(a, "calls", None)
            if reset is not None and rows is not None:
                reset(done.repeat_interleave(rows.shape[0] // sim.n_envs))


def _finish_step(sim, custom):
    custom.view(sim)
    custom.new_games(sim, sim.step_out()[:, 0] != sfcfg.RUNNING)


def play(sim, custom: Custom, steps: int):
    """The loop of gameplay::play() (gameplay.hpp:1443-1472) for a batch of arenas; returns the
    device-reduced statistics.  Observations, predictions and commands stay on the device."""
    custom.prepare(sim)
    actions = torch.full((sim.n_envs, sim.n_agents), ord("+"), dtype=torch.uint8, device=sim.device)
    squad_mask = ((1 << sim.n_agents) - 1) & ~1
    if sim.cfg.mode == sfcfg.MODE_ROYALE:
        # Battle Royale: every player is the `ind` of its own client and observes at the loop top
        # (get_my_action, :956); the commands of the others arrive over the wire (:977-986)
        for _ in range(steps):
            actions[:] = custom.bot(sim, (1 << sim.n_agents) - 1, sfcfg.OBS_P1)
            sim.step(actions)
            _finish_step(sim, custom)
        return sim.stats()
    for _ in range(steps):
        actions[:, 0:1] = custom.bot(sim, 1, sfcfg.OBS_P1)  # get_my_action, :956
        if squad_mask:
            sim.step_a()
            actions[:, 1:] = custom.bot(sim, squad_mask, sfcfg.OBS_P2)  # get_command -> bot(), :933
            custom.view(sim)
            sim.step_b(actions)
        else:
            sim.step(actions)
        _finish_step(sim, custom)
    return sim.stats()
