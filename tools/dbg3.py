import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from sad_monte_carlo_b200 import WalkerEngine, make_config, _abi
from tests.oracle_lib import OracleMC
from tests.test_gpu_lj import lj_cfg
fast = WalkerEngine(lj_cfg(n_walkers=4, lanes=1, flags=4))
exact = WalkerEngine(lj_cfg(n_walkers=4, lanes=1))
o = OracleMC(lj_cfg(n_walkers=4, lanes=1), walker=1)
rng = np.random.default_rng(1)
worst = 0
for step in range(600):
    st = o.walker()
    for eng in (fast, exact):
        eng.set_system(1, o.system())
        r = eng.rngs(); r[1] = (st.rng_s0, st.rng_s1); eng.set_rngs(r)
    scale = 0.05 if step % 3 else 0.3
    ef, ee, eo = fast.plan_move(1, scale), exact.plan_move(1, scale), o.plan_move(scale)
    if eo is not None:
        rel = abs(ef - eo) / max(1, abs(eo))
        if rel > 1e-13:
            print(step, "fast", repr(ef), "exact", repr(ee), "oracle", repr(eo), "rel", rel)
        if eo < o.energy() or rng.random() < 0.3:
            o.confirm()
print("done")
