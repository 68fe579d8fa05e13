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


# ---- directed-loop update (directed_loop.rs:103-301) -------------------------------------------------------------
# the Hamiltonian of tests/check_loop_crash.rs:19-27 as an interaction matrix: weight 1 when inputs == outputs or
# inputs == reversed outputs, index = (out0 out1 in0 in1), Interaction::at (qmc_runner.rs:560-600)
SWAP = [1.0 if ((i0, i1) == (o0, o1) or (i0, i1) == (o1, o0)) else 0.0 for o0 in (0, 1) for o1 in (0, 1) for i0 in (0, 1) for i1 in (0, 1)]


def op_word(bond, ins, outs):
    w = bond
    for r, (i, o) in enumerate(zip(ins, outs)):
        w |= int(i) << (24 + r)
        w |= int(o) << (26 + r)
    return w


def test_loop_update_known_answer():
    """Hand-derived: one diagonal bond op on [F, F].  Words: gen_range(0..1) -> 0; gen_range(0..2) -> var 0; gen::<bool>
    (sign bit of the upper half) -> Inputs; entrance (0, In) gives leg weights [1 (bounce), 0, 1, 1] (total 3);
    gen_range(0. ..3.) with value0_1 = 0.5 -> 1.5 -> leaves through (0, Out): the op becomes [T, F] -> [T, F]; no later op
    on variable 0, so state[0] = T and the walk re-enters the first op through (0, In) = where it started: done."""
    q = po.QmcOracle(2, key=0, state=[0, 0])
    q.make_interaction(SWAP, [0, 1])
    q.load_ops([op_word(0, [0, 0], [0, 0])], [0, 0])
    q.set_script([0, 0, 1 << 63, 1 << 63])
    q.loop_update()
    assert q.error == 0 and q.cursor == 4
    assert list(q.state()) == [1, 0]
    assert list(q.dump_ops()[:1]) == [op_word(0, [1, 0], [1, 0])]
    assert q.verify()


@pytest.mark.parametrize("nvars,ops", [(2, [(0, 1)]), (3, [(0, 1), (1, 2)])])
def test_check_loop_crash(nvars, ops):
    # tests/check_loop_crash.rs:6-74: run_single_bond / run_double_bond -- 100 loop updates, then verify
    q = po.QmcOracle(nvars, key=0xC4A5 + nvars, state=[0] * nvars)
    for a, b in ops:
        q.make_interaction(SWAP, [a, b])
    q.load_ops([op_word(k, [0, 0], [0, 0]) for k in range(len(ops))], [0] * nvars)
    seen = set()
    for _ in range(100):
        q.loop_update()
        assert q.error == 0 and q.verify()
        seen.add(tuple(q.state()))
    assert len(seen) > 1  # the loops do move the state


def test_loop_updates_match_exact_diagonalisation():
    """A model whose off-diagonal weight sits in two-variable exchange terms (no cluster edges at all): only the loop
    update can create off-diagonal ops, so the energy is right only if its weights and its walk are."""
    def xxz(d0, d1, x):  # diag(00, 01, 10, 11) = d0, d1, d1, d0; exchange 01 <-> 10 with weight x
        m = np.zeros((4, 4))
        m[0, 0] = m[3, 3] = d0
        m[1, 1] = m[2, 2] = d1
        m[1, 2] = m[2, 1] = x
        return list(m.reshape(-1))

    inters = [(xxz(0.4, 1.1, 0.8), [0, 1], False), (xxz(0.9, 0.5, 0.6), [1, 2], False), (xxz(0.3, 1.0, 1.0), [2, 3], False),
              (xxz(0.7, 0.7, 0.5), [3, 0], False)]
    beta, chains = 1.2, 256
    exact = dense_energy(4, inters, beta)
    reps = []
    for r in range(chains):
        q = po.QmcOracle(4, key=0x100B00 + r)
        for mat, vs, _ in inters:
            q.make_interaction(mat, vs)
        q.set_do_loop_updates(True)
        assert not q.has_cluster_edges and not q.breaks_ising_symmetry
        reps.append(q)
    po.sse_batch_timesteps(reps, 2000, [beta] * chains, po.MODE_STRICT)
    _, e = po.sse_batch_timesteps(reps, 20000, [beta] * chains, po.MODE_STRICT)
    assert all(q.error == 0 and q.verify() for q in reps)
    mean, err = e.mean(), e.std(ddof=1) / np.sqrt(chains)
    assert abs(mean - exact) < 3.5 * err + 1e-9, (mean, err, exact)
    # without the loop update the exchange terms never appear: the energy is the diagonal model's, far from `exact`
    diag_only = dense_energy(4, [(list(np.diag(np.array(m).reshape(4, 4))), vs, True) for m, vs, _ in inters], beta)
    assert abs(diag_only - exact) > 20 * err
