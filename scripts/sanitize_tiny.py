import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np
import sister_b200
vol = np.random.default_rng(1).integers(0, 253, (16, 24, 40), dtype=np.uint8)
with sister_b200.Engine(96, 64, 40, n_slots=1) as eng:
    eng.test_sgm(vol)
print("ok")
