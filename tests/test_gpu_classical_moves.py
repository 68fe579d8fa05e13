"""GPU parity: the reference's own classical schedule -- do_time_step with spin, edge and worm moves
(classical/graph.rs:91-406) -- replica-parallel on the device, against the CPU oracle under the same streams, plus
the reference's own worm tests (graph.rs:481-647), whose end states hold for every generator."""
import numpy as np
import pytest

from isingmontecarlo_b200 import lattices
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu

TRIANGLE = [((0, 1), 1.0), ((1, 2), 1.0), ((2, 0), 1.0)]


def make(edges, biases, betas, keys, state=None):
    from isingmontecarlo_b200.classical import GraphState

    g = GraphState(edges, biases, keys, betas, state=state)
    refs = [po.ClassicalOracle(edges, biases, key=int(k), state=state) for k in keys]
    return g, refs


def same(g, refs):
    st, cur = g.state_ref(), g.rng_cursors()
    for r, ref in enumerate(refs):
        assert np.array_equal(st[r], ref.state()), r
        assert int(cur[r]) == ref.cursor, r


def test_reference_worm_tests_on_the_device():
    keys = np.arange(64, dtype=np.uint64)
    z3 = np.zeros(3, dtype=np.uint8)
    for biases, beta, doubles, want in (([0.0] * 3, 1.0, False, "all"), ([-1.0] * 3, 1.0, False, "all"),
                                        ([1.0] * 3, 1000.0, False, "none"), ([0.0] * 3, 1.0, True, "equal")):
        g, refs = make(TRIANGLE, biases, beta, keys, state=z3)
        g.do_worm_flip(1, allow_doubles=doubles)
        st = g.state_ref()
        if want == "all":
            assert st.all()
        elif want == "none":
            assert not st.any()
        else:
            assert np.all(st.min(axis=1) == st.max(axis=1))
        for ref in refs:
            ref.worm_flips(beta, 1, allow_doubles=doubles)
        same(g, refs)
    # bounce (graph.rs:538-562)
    nvars = 20
    edges = [((x, x + 1), 1.0) for x in range(nvars - 1)]
    biases = [0.0] * nvars
    biases[0] = biases[-1] = 10.0
    g, refs = make(edges, biases, 1000.0, keys, state=np.zeros(nvars, dtype=np.uint8))
    g.do_worm_flip(1, allow_doubles=False)
    assert not g.state_ref().any()


@pytest.mark.parametrize("name", ["mixed4", "bathroom4", "triangular6", "random"])
def test_moves_match_oracle(name):
    rng = np.random.default_rng(5)
    if name == "mixed4":
        edges, n = lattices.two_d_periodic_mixed(4), 16
        biases = np.zeros(n)
    elif name == "bathroom4":
        edges, n = lattices.bathroom_unit_cells(4), 64
        biases = np.zeros(n)
    elif name == "triangular6":
        edges, n = lattices.triangular_periodic(6, 1.0), 36
        biases = np.full(n, 0.25)
    else:
        edges = [(e, float(rng.normal())) for e, _ in lattices.square_periodic(5, 1.0)]
        n = 25
        biases = rng.normal(size=n)
    keys = np.arange(100, 108, dtype=np.uint64)
    betas = np.linspace(0.3, 1.5, len(keys))
    g, refs = make(edges, biases, betas, keys)
    same(g, refs)
    g.do_spin_flip(7)
    for r, ref in enumerate(refs):
        ref.spin_flips(betas[r], 7)
    same(g, refs)
    g.do_edge_flip(5)
    for r, ref in enumerate(refs):
        ref.edge_flips(betas[r], 5)
    same(g, refs)
    for doubles in (False, True):
        g.do_worm_flip(3, allow_doubles=doubles)
        for r, ref in enumerate(refs):
            ref.worm_flips(betas[r], 3, allow_doubles=doubles)
        same(g, refs)
    for step in range(12):
        basic = step % 4 == 3
        ch = g.do_time_step(only_basic_moves=basic) if step % 2 else g.do_time_step(4, 3, 2, basic)
        for r, ref in enumerate(refs):
            want = ref.do_time_step(betas[r], only_basic_moves=basic) if step % 2 else ref.do_time_step(betas[r], 4, 3, 2, basic)
            assert int(ch[r]) == want
        same(g, refs)
    e = g.get_energy()
    for r, ref in enumerate(refs):
        assert abs(e[r] - ref.energy()) <= 1e-9 * max(1.0, abs(ref.energy()))


def test_edge_importance_sampling_and_errors():
    from isingmontecarlo_b200 import QmcbError

    edges = [((0, 1), 1.0), ((2, 3), 3.0), ((1, 2), 0.5), ((3, 0), 2.0)]
    keys = np.arange(16, dtype=np.uint64)
    g, refs = make(edges, [0.1, -0.2, 0.0, 0.3], 0.8, keys)
    g.enable_edge_importance_sampling(True)
    for ref in refs:
        ref.enable_edge_importance_sampling(True)
    g.do_edge_flip(9)
    for ref in refs:
        ref.edge_flips(0.8, 9)
    same(g, refs)
    g.enable_edge_importance_sampling(False)
    for ref in refs:
        ref.enable_edge_importance_sampling(False)
    g.do_edge_flip(4)
    for ref in refs:
        ref.edge_flips(0.8, 4)
    same(g, refs)
    # negative total weight: the reference panics on the empty range
    g2, _ = make([((0, 1), -1.0)], [0.0, 0.0], 1.0, keys[:2])
    g2.enable_edge_importance_sampling(True)
    with pytest.raises(QmcbError):
        g2.do_edge_flip(1)


def test_square_layout_takes_the_reference_schedule_too():
    # bit-packed layout: moves run on the unpacked bytes and are packed back; checkerboard sweeps in between
    L = 64
    edges = lattices.square_periodic(L, -1.0)
    keys = np.array([7, 8], dtype=np.uint64)
    g, refs = make(edges, np.zeros(L * L), 0.44, keys)
    assert g.is_bitpacked_square()
    colours, _ = g.colours()
    for _ in range(2):
        ch = g.do_time_step()
        for r, ref in enumerate(refs):
            assert ref.do_time_step(0.44) == int(ch[r])
        same(g, refs)
        g.sweeps(2)
        for ref in refs:
            ref.checkerboard_sweeps(0.44, colours, 2)
        same(g, refs)
