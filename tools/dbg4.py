import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests.test_gpu_fluids import _wca_pair
from tests.oracle_lib import OracleMC
cfg, eng, state = _wca_pair(64, 0.7, 4, method="samc")
print("created", flush=True)
for n in (1, 10, 100, 1000, 4000, 200, 5000):
    t = time.time(); eng.run(n); print("gpu", n, time.time() - t, eng.walker(0).energy, flush=True)
o = OracleMC(cfg, walker=0, system_state=state)
t = time.time(); o.run(10311); print("oracle", time.time() - t, o.energy(), flush=True)
