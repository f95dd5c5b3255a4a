"""K2d timing with sites in original order vs sorted by detection count (instruction-cache locality of the
per-n1 instantiations)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import biolith_b200 as bb
from biolith_b200 import _lib
import bench

lib = _lib.load()
model, X, W, y, chains, shard = bench.make_data("occu_rn_200k_x10_k50", 0)
n1 = np.nansum(y[0, :, 0, :], axis=1)
order = np.argsort(n1, kind="stable")
theta = np.random.default_rng(1000).uniform(-2, 2, size=(256, 10)).astype(np.float32)
tm = (np.concatenate([bench.make_data.true_theta]) + 0.01 * np.random.default_rng(1).standard_normal((256, 10))).astype(np.float32)
for name, idx in (("original order", np.arange(X.shape[0])), ("sorted by n1", order)):
    with bb.OccupancyLikelihood(model, X[idx], W[idx], y[:, idx], max_abundance=50, max_chains=256) as lk:
        for tn, th in (("uniform", theta), ("mode", tm)):
            ms = bench._time_eval(lib, lk, th, 0, steps=5)
            print(name, tn, round(ms, 3), "ms")
print("n1 histogram", np.bincount(n1.astype(int)))
