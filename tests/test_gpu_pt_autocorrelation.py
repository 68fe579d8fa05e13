"""ParallelTemperingAutocorrelations / ParallelTemperingBondAutoCorrelations of a TemperingContainer
(tempering_container.rs:484-630): qmcb_pt_variable_/spin_product_/bond_autocorrelation.  The series of a ladder slot are
the samples parallel_timesteps_sample (:411-453) takes at that slot; they must equal what a twin container's
timesteps_sample returns, and the device's exact bit-count correlation must agree with a literal numpy restatement of
fft_autocorrelation (autocorrelations.rs:99-133) fed with the reference's sample mappers."""
import numpy as np
import pytest

from isingmontecarlo_b200 import MODE_COUNTER, MODE_FAST, QmcbError, lattices
from tests.test_gpu_sse_parity import fft_autocorrelation

pytestmark = pytest.mark.gpu


def make(mode, n_chains=2, n_betas=4):
    from isingmontecarlo_b200.tempering import TemperingContainer

    edges = lattices.two_d_periodic_mixed(4)
    betas = np.linspace(0.8, 1.1, n_betas)  # close enough for swaps to be accepted often
    keys = 0x7A5E0000 + np.arange(n_chains * n_betas, dtype=np.uint64)
    tc = TemperingContainer(edges, 2.0, 0.0, 16, betas, n_chains=n_chains, rng_keys=keys, pt_key=0x5EED, mode=mode)
    tc.timesteps(30)
    return tc, edges


@pytest.mark.parametrize("mode", [MODE_FAST, MODE_COUNTER])
def test_variable_autocorrelation_per_slot(mode):
    tc, _ = make(mode)
    twin, _ = make(mode)
    T, swap, freq = 48, 3, 2
    ac, samples = tc.calculate_variable_autocorrelation(T * freq, swap, freq, return_samples=True)
    assert ac.shape == (tc.S, T) and samples.shape == (tc.S, T, 16)
    states, _ = twin.timesteps_sample(T * freq, swap, freq)  # the same trajectory through the reference-shaped call
    assert twin.get_total_swaps() == tc.get_total_swaps() > 0
    for slot in range(tc.S):
        assert np.array_equal(np.array(states[slot], dtype=np.uint8), samples[slot]), slot
        want = fft_autocorrelation(2.0 * samples[slot].astype(np.float64) - 1.0)
        assert np.allclose(ac[slot], want, rtol=0, atol=1e-10), np.abs(ac[slot] - want).max()
        assert abs(ac[slot][0] - 1.0) < 1e-12


def test_product_and_bond_autocorrelations_per_slot():
    tc, edges = make(MODE_FAST)
    prods = [[0, 1], [2, 5, 9], [3], [4, 8, 12, 15]]
    ac, samples = tc.calculate_spin_product_autocorrelation(96, 2, prods, 2, return_samples=True)
    for slot in range(tc.S):
        pm = 2.0 * samples[slot].astype(np.float64) - 1.0
        want = fft_autocorrelation(np.stack([pm[:, p].prod(axis=1) for p in prods], axis=1))
        assert np.allclose(ac[slot], want, rtol=0, atol=1e-10)
    ac, samples = tc.calculate_bond_autocorrelation(96, None, 2, return_samples=True)  # replica_swap_freq None = 1 (:590)
    for slot in range(tc.S):
        s = samples[slot].astype(bool)
        vals = []
        for (a, b), j in edges:
            even = ((s[:, a].astype(int) + s[:, b].astype(int)) % 2) == 0
            vals.append(np.where(even if j < 0.0 else ~even, 1.0, -1.0))
        want = fft_autocorrelation(np.stack(vals, axis=1))
        assert np.allclose(ac[slot], want, rtol=0, atol=1e-10, equal_nan=True)
    assert tc.verify()
