"""EXPERIMENT (SADMC_FLAG_HELPER_WARPS): the LJ thread-per-walker kernel with a helper warp per bookkeeping warp for the
pair loop (csrc/sys_lj_paired.cuh) must stay in the tolerance tier: same generator stream, same decisions, energies
within 1e-12 of the reference-order oracle."""
import numpy as np
import pytest

from sad_monte_carlo_b200 import WalkerEngine, make_config, _abi
from tests.oracle_lib import OracleMC

pytestmark = pytest.mark.gpu
RTOL = 1e-12


def cfg(n_walkers, flags, N=31, R=2.5):
    return make_config("lj", "sad", N=N, lj_radius=R, max_allowed_energy=0.0, sad_min_T=0.01, energy_bin=0.01, n_walkers=n_walkers,
                       seed=0, init_mode=_abi.INIT_RANDOMIZE, lanes_per_walker=1, flags=flags, bin_window_lo=-8.8 * N, bin_window_hi=0.02)


@pytest.mark.parametrize("N,R,walkers", [(31, 2.5, 200), (38, 3.0, 70)])
def test_helper_warp_kernel_tracks_the_reference_trajectory(N, R, walkers):
    eng = WalkerEngine(cfg(walkers, _abi.FLAG_FAST_MATH | _abi.FLAG_HELPER_WARPS, N, R))
    oracles = {w: OracleMC(cfg(walkers, 0, N, R), walker=w) for w in (0, walkers // 2, walkers - 1)}
    eng.run(30000)
    eng.run(7)
    for w, o in oracles.items():
        o.run(30007)
        g, s = eng.walker(w), o.walker()
        assert g.status == 0
        assert abs(g.energy - s.energy) <= RTOL * abs(s.energy)
        assert (g.rng_s0, g.rng_s1, g.accepted_moves) == (s.rng_s0, s.rng_s1, s.accepted_moves)
        assert np.array_equal(eng.bins(w)["histogram"], o.bins()["histogram"])
        assert np.array_equal(eng.system(w)[:-2], o.system()[:-2])
        assert abs(eng.compute_energy(w) - g.energy) <= 1e-14 * N * N * abs(g.energy)
