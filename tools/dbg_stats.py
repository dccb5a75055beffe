import sys, torch, numpy as np
sys.path.insert(0, '/root/repo')
from strikeforce_b200 import config as sfcfg
from strikeforce_b200.sim import BatchedArena
for E in (4096, 65536, 131072):
    sim = BatchedArena(E, mode="Squad", level=1, level_max=10, auto_reset=True, max_steps=2048)
    print(E, "after create", sim.stats())
    for t in range(3):
        sim.step(sim.synth_actions(t, sfcfg.ACTIONS28))
    print(E, "after 3 steps", sim.stats())
    ids = np.arange(E, dtype=np.int32)
    sim.reset(ids[ids % 16 == 0])
    print(E, "after partial reset", sim.stats())
    for t in range(3):
        sim.step(sim.synth_actions(t, sfcfg.ACTIONS28))
    print(E, "after 3 more", sim.stats())
    sim.close()
