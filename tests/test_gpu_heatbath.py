"""Heat-bath diagonal update (heatbath.rs:149-209, toggle qmc_ising.rs:444-486): the CUDA
path against the oracle, bit-exact, in both cluster orders and both kernel implementations."""
import numpy as np
import pytest

from isingmontecarlo_b200 import MODE_FAST, MODE_STRICT, lattices
from oracle import pyoracle as po
from tests.test_gpu_sse_parity import CASES, assert_same, make_pair
from tests.test_gpu_full_size import same, to_oracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("impl", [0, 1])
@pytest.mark.parametrize("mode", [MODE_STRICT, MODE_FAST])
@pytest.mark.parametrize("name,edges,gamma,h,cutoff,beta,sweeps", CASES)
def test_heatbath_sweeps_bit_exact(name, edges, gamma, h, cutoff, beta, sweeps, mode, impl):
    g, refs = make_pair(edges, gamma, h, cutoff, beta, mode, impl=impl)
    g.set_enable_heatbath(True)
    assert g.get_enable_heatbath()
    for ref in refs:
        ref.set_enable_heatbath(True)
    for chunk in (1, 1, 3, sweeps - 5):
        e_gpu = g.timesteps(chunk, beta)
        e_ref = [ref.timesteps(chunk, beta, mode) for ref in refs]
        assert_same(g, refs, f"{name} after +{chunk}")
        assert np.array_equal(e_gpu, np.array(e_ref)), name
    assert g.verify()
    # toggling back restores the Metropolis rule (qmc_ising.rs:483-485)
    g.set_enable_heatbath(False)
    for ref in refs:
        ref.set_enable_heatbath(False)
    g.timesteps(2, beta)
    for ref in refs:
        ref.timesteps(2, beta, mode)
    assert_same(g, refs, f"{name} metropolis again")


def test_heatbath_single_diagonal_step_matches():
    g, refs = make_pair(lattices.square_periodic(8, -1.0), 3.04, 0.0, 64, 4.0, MODE_FAST)
    g.set_enable_heatbath(True)
    for ref in refs:
        ref.set_enable_heatbath(True)
    for _ in range(6):
        g.single_diagonal_step()
        for ref in refs:
            ref.single_diagonal_step(4.0)
        assert_same(g, refs, "single diagonal step")


@pytest.mark.parametrize("name,mk,gamma,h,beta,cutoff,R,therm", [
    ("cfg3", lambda: lattices.square_periodic(32, -1.0), 3.04, 0.0, 16.0, 1024, 16, 60),
    ("cfg5", lambda: lattices.triangular_periodic(48, 1.0), 1.0, 0.2, 8.0, 2304, 8, 40),
])
def test_heatbath_full_size_parity_from_thermalised_state(name, mk, gamma, h, beta, cutoff, R, therm):
    from isingmontecarlo_b200.sse import QmcIsingGraph

    edges = mk()
    g = QmcIsingGraph(edges, gamma, h, cutoff, 0x4EA70000 + np.arange(R, dtype=np.uint64), beta, mode=MODE_FAST)
    g.set_enable_heatbath(True)
    e_therm = g.timesteps(therm, beta)
    assert g.verify()
    refs = {r: to_oracle(g, r, edges, gamma, h) for r in (0, R - 1)}
    for ref in refs.values():
        ref.set_enable_heatbath(True)
    e = g.timesteps(3, beta)
    for r, ref in refs.items():
        e_ref = ref.timesteps(3, beta, MODE_FAST)
        assert ref.error == 0 and same(g, r, ref), (name, r)
        assert e[r] == e_ref
    assert g.verify()


def test_heatbath_and_metropolis_agree_on_the_energy():
    # same Markov chain target: energies of the two diagonal rules agree within 3 sigma over chains
    from isingmontecarlo_b200.sse import QmcIsingGraph

    edges, beta, R = lattices.square_periodic(8, -1.0), 2.0, 64
    es = []
    for hb in (False, True):
        g = QmcIsingGraph(edges, 3.04, 0.0, 64, 0x77000000 + 1000 * hb + np.arange(R, dtype=np.uint64), beta, mode=MODE_FAST)
        g.set_enable_heatbath(hb)
        g.timesteps(300, beta)
        es.append(g.timesteps(1500, beta))
    d = es[0].mean() - es[1].mean()
    err = np.sqrt(es[0].var(ddof=1) / R + es[1].var(ddof=1) / R)
    assert abs(d) < 3.5 * err, (d, err)
