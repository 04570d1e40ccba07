import sys, numpy as np
sys.path.insert(0,'/root/repo')
import gpu_nbody_simulation_b200 as bh
from gpu_nbody_simulation_b200 import initial_conditions as ic
n=int(sys.argv[1]); mode=sys.argv[2]
pos,vel,mass=ic.uniform_disk(n,seed=12345)
with bh.Simulation(n) as sim:
    sim.set_bodies(pos,vel,mass)
    if mode=='forces':
        sim.build_tree(); sim.compute_forces(); f=sim.forces(); print('forces ok', np.isfinite(f).all(), float(np.abs(f).max()))
    else:
        sim.step(1); p=sim.positions(); print('step ok', np.isfinite(p).all())
