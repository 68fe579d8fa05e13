"""Differential test on random small systems: random connected-or-not graphs, mixed-sign and unequal couplings,
random fields of either sign, random initial cutoffs (down to 1, so capacity growth and full strings occur),
both cluster orders, both diagonal rules.  The CUDA path must equal the oracle bit for bit after every chunk of
sweeps.  (The reference's own tests use fixed small lattices -- tests/longitudinal_crash.rs, tests/convert_test.rs;
this widens them in the same spirit.)"""
import numpy as np
import pytest

from isingmontecarlo_b200 import MODE_FAST, MODE_STRICT
from oracle import pyoracle as po
from tests.test_gpu_sse_parity import assert_same

pytestmark = pytest.mark.gpu


def random_system(rng):
    nv = int(rng.integers(2, 14))
    ne = int(rng.integers(1, 3 * nv))
    edges = []
    for _ in range(ne):
        a, b = rng.choice(nv, size=2, replace=False)
        j = float(rng.choice([-1.0, 1.0, 0.5, -0.25, 2.0, 1.0 / 3.0]))
        edges.append(((int(a), int(b)), j))
    # the reference derives nvars from the largest index (qmc_ising.rs:92): make sure it is used
    edges.append(((0, nv - 1), float(rng.choice([-1.0, 1.0]))))
    gamma = float(rng.choice([0.3, 1.0, 2.5]))
    h = float(rng.choice([0.0, 0.0, 0.4, -0.7]))
    beta = float(rng.choice([0.2, 1.0, 3.0]))
    cutoff = int(rng.choice([1, 2, nv, 4 * nv]))
    return edges, nv, gamma, h, beta, cutoff


@pytest.mark.parametrize("seed", range(24))
def test_random_systems_bit_exact(seed):
    from isingmontecarlo_b200.sse import QmcIsingGraph

    rng = np.random.default_rng(1000 + seed)
    edges, nv, gamma, h, beta, cutoff = random_system(rng)
    mode = MODE_STRICT if seed % 2 else MODE_FAST
    heatbath = (seed // 2) % 2 == 1
    impl = 1 if seed % 8 == 7 else 0
    keys = [int(k) for k in rng.integers(1, 2**62, size=4)]
    g = QmcIsingGraph(edges, gamma, h, cutoff, keys, beta, mode=mode)
    g.set_option("impl", impl)
    g.set_enable_heatbath(heatbath)
    refs = [po.SseOracle(edges, gamma, h, cutoff, key=k) for k in keys]
    for ref in refs:
        ref.set_enable_heatbath(heatbath)
    assert g.nvars == nv
    for chunk in (1, 2, 5, 12):
        e = g.timesteps(chunk, beta)
        e_ref = np.array([ref.timesteps(chunk, beta, mode) for ref in refs])
        assert_same(g, refs, f"seed {seed} +{chunk}")
        assert np.array_equal(e, e_ref)
    assert g.verify() and all(ref.verify() and ref.error == 0 for ref in refs)
    # imaginary-time fold and bond counters agree as well
    m1, m2, mabs = g.imaginary_time_magnetization()
    for r, ref in enumerate(refs):
        slots, s1, s2, s3 = ref.itime_magnetization()
        assert m1[r] == s1 / slots / nv and m2[r] == s2 / slots / (nv * nv) and mabs[r] == s3 / slots / nv
        counts = g.get_bond_counts(r)
        assert [int(c) for c in counts] == [ref.bond_count(b) for b in range(len(counts))]
