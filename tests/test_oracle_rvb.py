"""RVB update of the oracle (oracle.c::orc_sse_rvb_update, restating rvb.rs:60-1221 and util/bondcontainer.rs).  The
reference holds no golden vectors for it (its RVB tests are crash tests: tests/check_rvb_crash.rs, the `rvb` cases of
tests/longitudinal_crash.rs), so the restatement is pinned the way the rest of the oracle is: the reference's own crash
tests with its invariant verify(), and exact diagonalisation of frustrated models -- a wrong acceptance ratio or a wrong
rotation breaks detailed balance and shows up in the energy."""
import numpy as np
import pytest

from isingmontecarlo_b200 import lattices
from oracle import pyoracle as po
from tests.ed import tfim_thermal

TRI2 = [((0, 1), 1.0), ((1, 2), 1.0), ((2, 0), 1.0), ((2, 3), 1.0), ((3, 4), 1.0), ((4, 2), 1.0)]
DIAMOND = [((0, 1), 1.0), ((1, 2), 1.0), ((2, 0), 1.0), ((2, 3), 1.0), ((3, 0), 1.0)]


@pytest.mark.parametrize("edges,gamma,h,nvars", [
    (lattices.two_d_periodic_mixed(3), 0.1, 0.0, 9),   # check_rvb_crash.rs:296-315 run_three
    (lattices.two_d_periodic_mixed(4), 0.1, 0.0, 16),  # :318-337 run_four
    (lattices.two_unit_cell(), 1.0, 0.0, 8),           # :340-359 run_two_unit_cell
    (lattices.two_d_periodic_mixed(3), 1.0, 1.0, 9),   # longitudinal_crash.rs rvb cases: h != 0 takes rvb_update_with_ising_weight
    (lattices.two_unit_cell(), 1.0, 1.0, 8),
])
@pytest.mark.parametrize("mode", [po.MODE_STRICT, po.MODE_FAST, po.MODE_COUNTER])
def test_rvb_crash_tests_keep_the_invariant(edges, gamma, h, nvars, mode):
    for seed in range(4):
        q = po.SseOracle(edges, gamma, h, nvars, key=seed, state=[0] * nvars)
        q.set_run_rvb(True)
        for _ in range(250):
            q.timestep(1.0, mode)
            assert q.error == 0
            assert q.verify()
        assert 0.0 < q.rvb_success_rate() < 1.0


def test_single_rvb_sweep_counts_and_keeps_the_invariant():
    q = po.SseOracle(lattices.two_unit_cell(), 1.0, 0.0, 8, key=3)
    q.timesteps(50, 2.0)
    n = q.n
    succ, att = q.single_rvb_sweep()
    assert att == (8 + 1) // 2 and 0 <= succ <= att  # qmc_ising.rs:375
    succ, att = q.single_rvb_sweep(40)
    assert att == 40 and 0 < succ <= 40
    assert q.error == 0 and q.verify() and q.n == n  # the update moves and flips ops, it never adds or removes one


@pytest.mark.parametrize("name,edges,gamma,h,beta", [
    ("two_triangles", TRI2, 0.3, 0.0, 3.0),
    ("diamond_mixed_J", [((0, 1), 1.0), ((1, 2), 0.7), ((2, 0), 1.3), ((2, 3), -1.0), ((3, 0), 1.0)], 0.4, 0.0, 2.5),
    ("diamond_h", DIAMOND, 0.4, 0.25, 2.0),
])
def test_rvb_energy_matches_exact_diagonalisation(name, edges, gamma, h, beta):
    nvars = lattices.nvars_of(edges)
    exact = tfim_thermal(edges, nvars, gamma, h, beta)
    chains = 32
    reps = [po.SseOracle(edges, gamma, h, nvars, key=0x7E0 + r) for r in range(chains)]
    for r in reps:
        r.set_run_rvb(True)
    po.sse_batch_timesteps(reps, 500, [beta] * chains, po.MODE_STRICT)
    _, e = po.sse_batch_timesteps(reps, 8000, [beta] * chains, po.MODE_STRICT)
    assert all(r.error == 0 and r.verify() for r in reps)
    assert all(r.rvb_success_rate() > 0.05 for r in reps)  # the move is doing work in these models
    mean, err = e.mean(), e.std(ddof=1) / np.sqrt(chains)
    assert abs(mean - exact["E"]) < 3.0 * err + 1e-9, (name, mean, err, exact["E"])
