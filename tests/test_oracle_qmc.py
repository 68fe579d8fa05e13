"""The oracle's restatement of the generic `Qmc` runner (qmc_runner.rs:46-156, :363-377, :406-680): the reference's own
tests/convert_test.rs (an Ising graph and its `into_qmc()` image stay identical), exact diagonalisation of a model that is
NOT a transverse-field Ising model, and the constructor errors."""
import numpy as np
import pytest

from isingmontecarlo_b200 import lattices
from oracle import pyoracle as po


@pytest.mark.parametrize("mode", [po.MODE_STRICT, po.MODE_FAST, po.MODE_COUNTER])
def test_convert_and_run(mode):
    # tests/convert_test.rs:9-31
    edges = lattices.one_d_periodic(3)
    ising = po.SseOracle(edges, 1.0, 0.0, 3, key=1234, state=[1, 1, 1])
    qmc = po.into_qmc(ising, edges, 1.0, 0.0)
    assert qmc.has_cluster_edges and not qmc.breaks_ising_symmetry
    for _ in range(10):
        ising.timestep(1.0, mode), qmc.timestep(1.0, mode)
        assert np.array_equal(ising.dump_ops(), qmc.dump_ops())
    assert np.array_equal(ising.state(), qmc.state()) and ising.cursor == qmc.cursor and ising.cutoff == qmc.cutoff
    assert qmc.verify() and qmc.error == 0
    # energy offsets differ by the field terms: Qmc's offset only collects the minima of the *_and_offset interactions
    assert ising.offset == 3.0 + 3 * 1.0 and qmc.offset == 3.0
    with pytest.raises(ValueError, match="negative"):  # qmc_ising.rs:966-969: [h, 0, 0, -h] has a negative entry, the reference's unwrap() panics
        po.into_qmc(po.SseOracle(edges, 1.0, 0.4, 3, key=5), edges, 1.0, 0.4)


def dense_energy(nvars, inters, beta):
    """thermal energy of H = -sum_b M_b with M_b the interaction matrices (Interaction::at indexing: first variable most
    significant, outputs more significant than inputs)"""
    dim = 1 << nvars
    H = np.zeros((dim, dim))
    for mat, vs, diagonal in inters:
        n = len(vs)
        M = np.diag(mat) if diagonal else np.array(mat, dtype=float).reshape(1 << n, 1 << n)  # [out][in]
        for s_in in range(dim):
            sub_in = 0
            for v in vs:
                sub_in = (sub_in << 1) | ((s_in >> v) & 1)
            for sub_out in range(1 << n):
                s_out = s_in
                for k, v in enumerate(vs):
                    bit = (sub_out >> (n - 1 - k)) & 1
                    s_out = (s_out & ~(1 << v)) | (bit << v)
                H[s_out, s_in] -= M[sub_out, sub_in]
    w = np.linalg.eigvalsh(H)
    bw = np.exp(-beta * (w - w.min()))
    return float((w * bw).sum() / bw.sum())


@pytest.mark.parametrize("mode", [po.MODE_STRICT, po.MODE_COUNTER])
def test_generic_interactions_match_exact_diagonalisation(mode):
    inters = [([0.3, 1.0, 1.0, 0.3], [0, 1], True), ([2.0, 0.5, 0.5, 2.0], [1, 2], True), ([0.2, 0.9, 0.9, 0.2], [2, 3], True),
              ([0.6, 1.4, 1.4, 0.6], [3, 0], True)] + [([g] * 4, [v], False) for v, g in enumerate([0.5, 1.0, 1.5, 0.8])]
    beta, chains = 1.5, 256
    exact = dense_energy(4, inters, beta)
    reps = []
    for r in range(chains):
        q = po.QmcOracle(4, key=0xC0DE00 + 97 * mode + r)
        for mat, vs, diagonal in inters:
            (q.make_diagonal_interaction if diagonal else q.make_interaction)(mat, vs)
        assert q.has_cluster_edges and not q.breaks_ising_symmetry
        reps.append(q)
    po.sse_batch_timesteps(reps, 1000, [beta] * chains, mode)
    _, e = po.sse_batch_timesteps(reps, 8000, [beta] * chains, mode)
    assert all(q.error == 0 and q.verify() for q in reps)
    mean, err = e.mean(), e.std(ddof=1) / np.sqrt(chains)
    assert abs(mean - exact) < 3.5 * err + 1e-9, (mean, err, exact)


def test_interaction_constructor_errors_and_offsets():
    q = po.QmcOracle(3, key=1)
    with pytest.raises(ValueError, match="power of 2"):
        q.make_interaction([1.0, 1.0, 1.0], [0])
    with pytest.raises(ValueError, match="negative"):
        q.make_interaction([1.0, -0.1, 1.0, 1.0], [0])
    with pytest.raises(ValueError, match="vars"):
        q.make_diagonal_interaction([1.0, 2.0, 2.0, 1.0], [0])
    q.make_diagonal_interaction_and_offset([-1.0, 1.0, 1.0, -1.0], [0, 1])  # qmc_ising.rs:956-958 for J = 1
    assert q.offset == 1.0
    q.make_interaction_and_offset([0.5, 0.2, 0.2, 0.5], [2])  # full 2x2 matrix: the diagonal minimum 0.5 is removed
    assert q.offset == 0.5 and not q.has_cluster_edges  # [0, .2, .2, 0] is not constant: no cluster edge
    q.make_interaction([0.7] * 4, [2])
    assert q.has_cluster_edges
