import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import biolith_b200 as bb
data, true = bb.simulate_occupancy("occu", random_seed=0)
for chains in (1, 5, 64):
    t0 = time.perf_counter()
    res = bb.fit(bb.models.occu, **data, num_chains=chains, num_samples=1000, num_warmup=1000)
    dt = time.perf_counter() - t0
    info = res.mcmc.info
    print(f"config1 occu S=100 J=52 chains={chains}: fit wall {dt:.2f}s, global steps {info['global_steps']}, "
          f"{1e6*info['wall_s']/info['global_steps']:.1f} us/step, psi {res.samples['psi'].mean():.3f} (z {true['z'].mean():.3f})")
