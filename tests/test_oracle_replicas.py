"""CPU checks of the replicas oracle (oracle/oracle_replicas.hpp): the structure energy_replicas.rs maintains."""
import numpy as np

from sad_monte_carlo_b200 import _abi, make_config
from tests.oracle_lib import OracleReplicas


def test_setup_gives_an_unbounded_and_a_median_replica_sharing_one_generator():
    o = OracleReplicas(make_config("fake", fake_function=_abi.FAKE_QUADRATIC, N=3, seed=2), 0.001, 16, max_init=2048)
    r0, r1 = o.replicas()
    assert np.isinf(r0.max_energy) and r0.cutoff_energy == r1.max_energy > r1.cutoff_energy  # energies[len/2], [len/4]
    assert (r0.rng_s0, r0.rng_s1) == (r1.rng_s0, r1.rng_s1) != o.rng()                        # rng.clone() twice, then jump (374-383)
    assert r0.translation_scale == r1.translation_scale == 0.5 and r1.energy <= r1.max_energy  # max_size (fake.rs:145)
    assert list(o.median()) == [r1.cutoff_energy]


def test_ladder_grows_by_halving_the_volume_below_and_counts_moves_per_replica():
    o = OracleReplicas(make_config("fake", fake_function=_abi.FAKE_QUADRATIC, N=3, seed=2), 0.001, 16, max_init=4096)
    n_before, moves = o.num_replicas(), 0
    for _ in range(2000):
        moves += 3 * o.num_replicas()  # steps = min_moves_to_randomize = 3 per replica present at the start of the round
        o.run_once(1)
    assert o.moves() == moves and o.num_replicas() > n_before
    reps = o.replicas()
    for a, b in zip(reps, reps[1:]):
        assert a.cutoff_energy == b.max_energy  # the assert of energy_replicas.rs:537
    # D(E) ~ sqrt(E): a cutoff at the median of the energies below splits the volume in two, E -> E / 2^(2/3)
    ratios = [b.cutoff_energy / a.cutoff_energy for a, b in zip(reps[2:-2], reps[3:-1])]
    assert all(0.5 < r < 0.75 for r in ratios), ratios
    # new replicas shrink the step by 0.5^(1/dimensionality) (581-582)
    assert abs(reps[-1].translation_scale / reps[-2].translation_scale - 0.5 ** (1 / 3)) < 1e-12 or reps[-1].accepted_count > 128
