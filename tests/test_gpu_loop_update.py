"""Directed-loop update (directed_loop.rs:103-301) behind the C ABI: qmcb_create_qmc(do_loop_updates), qmcb_loop_update,
qmcb_set_do_loop_updates.  Bit-exact against the oracle's restatement (pinned by tests/test_oracle_qmc.py: the
reference's tests/check_loop_crash.rs, a hand-derived answer and exact diagonalisation)."""
import numpy as np
import pytest

from isingmontecarlo_b200 import MODE_COUNTER, MODE_FAST, MODE_STRICT, _lib, lattices
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu

SWAP = [1.0 if ((i0, i1) == (o0, o1) or (i0, i1) == (o1, o0)) else 0.0 for o0 in (0, 1) for o1 in (0, 1) for i0 in (0, 1) for i1 in (0, 1)]


def op_word(bond, ins, outs):
    w = bond
    for r, (i, o) in enumerate(zip(ins, outs)):
        w |= int(i) << (24 + r)
        w |= int(o) << (26 + r)
    return w


def xxz(d0, d1, x):
    m = np.zeros((4, 4))
    m[0, 0] = m[3, 3] = d0
    m[1, 1] = m[2, 2] = d1
    m[1, 2] = m[2, 1] = x
    return list(m.reshape(-1))


@pytest.mark.parametrize("nvars,pairs", [(2, [(0, 1)]), (3, [(0, 1), (1, 2)])])
def test_check_loop_crash_through_the_c_abi(nvars, pairs):
    # tests/check_loop_crash.rs:6-74 on a batch of replicas, every loop update compared with the oracle
    from isingmontecarlo_b200.sse import Qmc

    keys = [0xC4A500 + r for r in range(5)]
    g = Qmc(nvars, keys, 1.0, state=[0] * nvars, do_loop_updates=True)
    refs = []
    for k in keys:
        q = po.QmcOracle(nvars, key=k, state=[0] * nvars)
        refs.append(q)
    for a, b in pairs:
        g.make_interaction(SWAP, [a, b])
        for q in refs:
            q.make_interaction(SWAP, [a, b])
    ops = [op_word(k, [0, 0], [0, 0]) for k in range(len(pairs))]
    g.increase_cutoff_to(len(ops))
    for r, q in enumerate(refs):
        g.load_ops(r, np.array(ops, dtype=np.uint32))
        q.load_ops(ops, [0] * nvars)
    for it in range(100):
        g.loop_update()
        st, cur = g.state_ref(), g.rng_cursors()
        for r, q in enumerate(refs):
            q.loop_update()
            assert q.error == 0
            assert np.array_equal(g.dump_ops(r)[:len(ops)], q.dump_ops()[:len(ops)]), (it, r)
            assert np.array_equal(st[r], q.state()) and int(cur[r]) == q.cursor, (it, r)
    assert g.verify()
    g.close()


INTERS_LOOP_ONLY = [(xxz(0.4, 1.1, 0.8), [0, 1]), (xxz(0.9, 0.5, 0.6), [1, 2]), (xxz(0.3, 1.0, 1.0), [2, 3]), (xxz(0.7, 0.7, 0.5), [3, 0])]


def build_pair(inters, site_weights, nvars, keys, loop):
    from isingmontecarlo_b200.sse import Qmc

    g = Qmc(nvars, keys, 1.2, do_loop_updates=loop)
    refs = [po.QmcOracle(nvars, key=k) for k in keys]
    for mat, vs in inters:
        g.make_interaction(mat, vs)
        for q in refs:
            q.make_interaction(mat, vs)
    for v, w in enumerate(site_weights):
        g.make_interaction([w] * 4, [v])
        for q in refs:
            q.make_interaction([w] * 4, [v])
    for q in refs:
        q.set_do_loop_updates(loop)
    return g, refs


@pytest.mark.parametrize("with_sites", [False, True])
def test_timesteps_with_loop_updates_bit_exact(with_sites):
    """Qmc::timestep with do_loop_updates (qmc_runner.rs:363-377): diagonal update, loop update, cluster update (only when
    the model has cluster edges), free bits -- operator strings, states, stream positions and energies equal the
    oracle's; two-variable off-diagonal ops appear and are propagated by the next diagonal update."""
    keys = [0x100B00 + r for r in range(6)]
    g, refs = build_pair(INTERS_LOOP_ONLY, [0.5, 1.0, 1.5, 0.8] if with_sites else [], 4, keys, True)
    assert g.should_do_cluster_update() == with_sites
    e = g.timesteps(60, 1.2)
    offdiag = 0
    for r, q in enumerate(refs):
        er = q.timesteps(60, 1.2, MODE_STRICT)
        assert q.error == 0 and er == e[r], r
        w = g.dump_ops(r)
        assert np.array_equal(w, q.dump_ops()), r
        assert np.array_equal(g.state_ref()[r], q.state())
        assert int(g.rng_cursors()[r]) == q.cursor and int(g.get_cutoff()[r]) == q.cutoff
        occ = w[w != 0xFFFFFFFF]
        two = (occ & 0xFFFFFF) < 4
        offdiag += int(np.sum(two & (((occ >> 24) & 3) != ((occ >> 26) & 3))))
    assert offdiag > 0 and g.verify()
    # turning the loop updates off (set_do_loop_updates) keeps the batch in step with the oracle
    g.set_do_loop_updates(False)
    g.timesteps(5, 1.2)
    for r, q in enumerate(refs):
        q.set_do_loop_updates(False)
        q.timesteps(5, 1.2, MODE_STRICT)
        assert np.array_equal(g.dump_ops(r), q.dump_ops()), r
    g.close()


def test_loop_updates_sample_the_right_ensemble():
    # exact diagonalisation of the exchange model (see tests/test_oracle_qmc.py), 512 replicas on the device
    from tests.test_oracle_qmc import dense_energy

    from isingmontecarlo_b200.sse import Qmc

    beta, R = 1.2, 512
    exact = dense_energy(4, [(m, v, False) for m, v in INTERS_LOOP_ONLY], beta)
    g = Qmc(4, 0x200B00 + np.arange(R, dtype=np.uint64), beta, do_loop_updates=True)
    for mat, vs in INTERS_LOOP_ONLY:
        g.make_interaction(mat, vs)
    g.timesteps(2000, beta)
    e = g.timesteps(12000, beta)
    mean, err = e.mean(), e.std(ddof=1) / np.sqrt(R)
    assert abs(mean - exact) < 3.5 * err + 1e-9, (mean, err, exact)
    assert g.verify()
    g.close()


def test_loop_update_shapes_and_modes():
    from isingmontecarlo_b200.sse import Qmc, QmcIsingGraph

    g = Qmc(4, [1, 2], 1.0, do_loop_updates=True)
    for mat, vs in INTERS_LOOP_ONLY:
        g.make_interaction(mat, vs)
    g.timesteps(3, 1.0)
    for mode in (MODE_FAST, MODE_COUNTER):  # the walk needs the reference's link structure
        with pytest.raises(_lib.QmcbError, match="STRICT"):
            g.set_mode(mode)
    g.close()
    ising = QmcIsingGraph(lattices.one_d_periodic(4, 1.0), 1.0, 0.0, 4, [1], 1.0)
    with pytest.raises(_lib.QmcbError, match="qmcb_create_qmc"):
        _lib.check(ising._L.qmcb_loop_update(ising._h))
    ising.close()
    # an exchange term that is not symmetric under the global flip: the reference would run no cluster update
    bad = Qmc(2, [1], 1.0, do_loop_updates=True)
    m = xxz(0.5, 0.5, 0.3)
    m[1 * 4 + 0] = 0.2  # <01|M|00> without its mirror <10|M|11>
    bad.make_interaction(m, [0, 1])
    with pytest.raises(_lib.QmcbError, match="Ising symmetry"):
        bad.timesteps(1, 1.0)
