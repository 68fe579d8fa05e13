"""STRICT cluster step on both workspace layouts (qmcb_set_option "strict_layout": bit 0 = world-line arrays, the default (9 = with the bond-partner request);
0 = one 32-byte record per slot): the cluster NUMBERING (cluster.rs:57-97 discovery order) and everything that follows
from it must equal the oracle's literal walk on either."""
import numpy as np
import pytest

from isingmontecarlo_b200 import MODE_STRICT, lattices
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu

CASES = [
    # name, edges, gamma, h, cutoff0, beta, sweeps   (small gamma: clusters without a site op -> non-edge start ops, :82-91, :205-211)
    ("square8_crit", lattices.square_periodic(8, -1.0), 3.04, 0.0, 64, 4.0, 12),
    ("square6_low_gamma", lattices.square_periodic(6, -1.0), 0.05, 0.0, 36, 2.0, 14),
    ("mixed4x4_h_low_gamma", lattices.two_d_periodic_mixed(4), 0.1, 0.6, 16, 2.0, 14),
    ("tri6_frustrated_h", lattices.triangular_periodic(6, 1.0), 1.0, 0.2, 36, 2.0, 10),
    ("ring5_tiny_gamma", lattices.one_d_periodic(5, 1.0), 0.02, 0.3, 5, 3.0, 20),
    ("pair_h", [((0, 1), 1.0)], 1.0, 1.0, 2, 1.0, 20),
]


@pytest.mark.parametrize("layout", [9, 0, 1, 15])  # 9: the default; 15: world lines + both L1 requests + links in their own launch
@pytest.mark.parametrize("name,edges,gamma,h,cutoff,beta,sweeps", CASES)
def test_numbering_matches_reference_order(name, edges, gamma, h, cutoff, beta, sweeps, layout):
    from isingmontecarlo_b200.sse import QmcIsingGraph

    keys = [0x57A1C700 + r for r in range(6)]
    g = QmcIsingGraph(edges, gamma, h, cutoff, keys, beta, mode=MODE_STRICT)
    g.set_option("strict_layout", layout)
    refs = [po.SseOracle(edges, gamma, h, cutoff, key=k) for k in keys]
    for s in range(sweeps):
        g.single_diagonal_step(beta)
        ncl = g.single_cluster_step()
        for r, ref in enumerate(refs):
            ref.single_diagonal_step(beta)
            assert ref.single_cluster_step(MODE_STRICT) == int(ncl[r]), (name, s, r)
            m = ref.cutoff
            bi, bo = g.boundaries(r, m)
            ri, ro = ref.boundaries(m)
            has = ref.dump_ops()[:m] != 0xFFFFFFFF  # (an empty string leaves the previous step's boundaries in the oracle)
            assert np.all(ri[has] >= 0)
            assert np.array_equal(bi[has].astype(np.int64), ri[has]), (name, s, r)
            assert np.array_equal(bo[has].astype(np.int64), ro[has]), (name, s, r)
            assert np.array_equal(g.dump_ops(r), ref.dump_ops()), (name, s, r)
            assert np.array_equal(g.state_ref()[r], ref.state()), (name, s, r)
        # free spins + cutoff growth as timestep() does, so that the next diagonal step starts from the same place
    cur = g.rng_cursors()
    for r, ref in enumerate(refs):
        assert int(cur[r]) == ref.cursor and ref.error == 0
    assert g.verify()


def test_layouts_agree_on_a_thermalised_lattice():
    from isingmontecarlo_b200.sse import QmcIsingGraph

    edges = lattices.square_periodic(16, -1.0)
    keys = [0x55E00000 + r for r in range(12)]
    a = QmcIsingGraph(edges, 3.04, 0.0, 256, keys, 8.0, mode=MODE_STRICT)
    b = QmcIsingGraph(edges, 3.04, 0.0, 256, keys, 8.0, mode=MODE_STRICT)
    b.set_option("strict_layout", 0)
    ea, eb = a.timesteps(25, 8.0), b.timesteps(25, 8.0)
    assert np.array_equal(ea, eb)
    assert np.array_equal(a.rng_cursors(), b.rng_cursors()) and np.array_equal(a.get_n(), b.get_n())
    assert np.array_equal(a.state_ref(), b.state_ref())
    for r in range(len(keys)):
        assert np.array_equal(a.dump_ops(r), b.dump_ops(r))
    ref = po.SseOracle(edges, 3.04, 0.0, 256, key=keys[3])
    ref.timesteps(25, 8.0, MODE_STRICT)
    assert np.array_equal(a.dump_ops(3), ref.dump_ops())
    assert a.verify() and b.verify()
