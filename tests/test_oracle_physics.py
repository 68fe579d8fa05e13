"""Physics anchors for the oracle: energies from -<n>/beta + offset
(qmc_ising.rs:805-809) agree with exact diagonalisation within 3 sigma over independent chains,
in both cluster orders, and every string satisfies the reference invariant verify()."""
import numpy as np
import pytest

from isingmontecarlo_b200 import lattices
from oracle import pyoracle as po
from tests.ed import tfim_thermal

SYSTEMS = [
    ("small_qmc_ring", lattices.small_qmc_ring(), 1.0, 0.0, 1.0),
    ("ring_h", lattices.small_qmc_ring(), 1.0, 0.5, 2.0),
    ("ferro6", lattices.one_d_periodic(6, -1.0), 0.7, 0.0, 3.0),
    ("frustrated", [((0, 1), 1.0), ((1, 2), 1.0), ((2, 0), 1.0), ((2, 3), 1.0), ((3, 0), 1.0)], 0.5, -0.3, 2.0),
]


@pytest.mark.parametrize("mode,heatbath", [(po.MODE_STRICT, False), (po.MODE_FAST, False), (po.MODE_STRICT, True), (po.MODE_FAST, True),
                                           (po.MODE_COUNTER, False)])
@pytest.mark.parametrize("name,edges,gamma,h,beta", SYSTEMS)
def test_energy_matches_exact_diagonalisation(name, edges, gamma, h, beta, mode, heatbath):
    nvars = lattices.nvars_of(edges)
    exact = tfim_thermal(edges, nvars, gamma, h, beta)
    chains = 32
    reps = [po.SseOracle(edges, gamma, h, nvars, key=0xE0 + 1000 * mode + 7000 * heatbath + r) for r in range(chains)]
    for r in reps:
        r.set_enable_heatbath(heatbath)  # heatbath.rs:149-209 instead of diagonal.rs:142-191
    po.sse_batch_timesteps(reps, 500, [beta] * chains, mode)  # thermalise
    _, e = po.sse_batch_timesteps(reps, 6000, [beta] * chains, mode)
    assert all(r.error == 0 for r in reps)
    assert all(r.verify() for r in reps)
    mean, err = e.mean(), e.std(ddof=1) / np.sqrt(chains)
    assert abs(mean - exact["E"]) < 3.0 * err + 1e-9, (name, mean, err, exact["E"])
    # magnetisation moments from the p=0 states of further sweeps
    m2 = []
    for r in reps[:16]:
        s, _ = r.timesteps_sample(2000, beta, 1, mode)
        m = (2.0 * s.astype(np.float64) - 1.0).mean(axis=1)
        m2.append((m * m).mean())
    m2 = np.array(m2)
    assert abs(m2.mean() - exact["m2"]) < 3.5 * m2.std(ddof=1) / np.sqrt(len(m2)) + 1e-9, (name, m2.mean(), exact["m2"])


def test_longitudinal_crash_lattices_verify():
    # tests/longitudinal_crash.rs:39-178 (16 seeds, 1000 steps, verify)
    cases = [([((0, 1), 1.0)], 1.0, 1.0, 2, [0, 0]), (lattices.two_d_periodic_mixed(3), 1.0, 1.0, 9, None),
             (lattices.two_d_periodic_mixed(4), 1.0, 1.0, 16, None), (lattices.two_unit_cell(), 1.0, 1.0, 8, None)]
    for edges, g, h, cutoff, state in cases:
        for seed in range(8):
            for mode in (po.MODE_STRICT, po.MODE_FAST, po.MODE_COUNTER):
                q = po.SseOracle(edges, g, h, cutoff, key=seed, state=state)
                q.timesteps(300, 1.0, mode)
                assert q.error == 0 and q.verify()


def test_dump_load_round_trip():
    edges = lattices.two_d_periodic_mixed(4)
    a = po.SseOracle(edges, 1.0, 0.3, 16, key=5)
    a.timesteps(200, 2.0)
    b = po.SseOracle(edges, 1.0, 0.3, a.cutoff, key=5, state=a.state())
    b.load_ops(a.dump_ops(), a.state())
    b.set_cursor(a.cursor)
    assert b.n == a.n and b.verify()
    for mode in (po.MODE_STRICT, po.MODE_FAST, po.MODE_COUNTER):
        a.timestep(2.0, mode), b.timestep(2.0, mode)
        assert np.array_equal(a.dump_ops(), b.dump_ops()) and np.array_equal(a.state(), b.state())
        assert a.cursor == b.cursor
