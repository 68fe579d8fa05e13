"""The reference's spin / edge / worm moves and its do_time_step schedule (classical/graph.rs:91-406) on the CPU
oracle.  The worm tests are the reference's own (graph.rs:481-647): their end states do not depend on the generator,
so they hold for every key of the injected stream."""
import itertools

import numpy as np
import pytest

from isingmontecarlo_b200 import lattices
from oracle import pyoracle as po

TRIANGLE = [((0, 1), 1.0), ((1, 2), 1.0), ((2, 0), 1.0)]
KEYS = range(24)


@pytest.mark.parametrize("key", KEYS)
def test_worm_flip(key):  # graph.rs:481-498
    g = po.ClassicalOracle(TRIANGLE, [0.0, 0.0, 0.0], key=key, state=[0, 0, 0])
    g.worm_flips(1.0, 1, allow_doubles=False)
    assert g.state().all()


@pytest.mark.parametrize("key", KEYS)
def test_worm_flip_bias(key):  # :500-517
    g = po.ClassicalOracle(TRIANGLE, [-1.0, -1.0, -1.0], key=key, state=[0, 0, 0])
    g.worm_flips(1.0, 1, allow_doubles=False)
    assert g.state().all()


@pytest.mark.parametrize("key", KEYS)
def test_worm_flip_bias_not(key):  # :519-536
    g = po.ClassicalOracle(TRIANGLE, [1.0, 1.0, 1.0], key=key, state=[0, 0, 0])
    g.worm_flips(1000.0, 1, allow_doubles=False)
    assert not g.state().any()


@pytest.mark.parametrize("key", KEYS)
def test_worm_flip_bounce(key):  # :538-562
    nvars = 20
    edges = [((x, x + 1), 1.0) for x in range(nvars - 1)]
    biases = [0.0] * nvars
    biases[0] = biases[-1] = 10.0
    g = po.ClassicalOracle(edges, biases, key=key, state=[0] * nvars)
    g.worm_flips(1000.0, 1, allow_doubles=False)
    assert not g.state().any()


@pytest.mark.parametrize("key", KEYS)
def test_worm_flip_doubles(key):  # :564-581
    g = po.ClassicalOracle(TRIANGLE, [0.0, 0.0, 0.0], key=key, state=[0, 0, 0])
    g.worm_flips(1.0, 1, allow_doubles=True)
    st = g.state()
    assert st.all() or not st.any()


@pytest.mark.parametrize("key", range(4))
def test_worm_2d_and_bathroom_terminate(key):  # :583-647 (no assertion in the reference: must not hang or crash)
    edges = lattices.two_d_periodic_mixed(4)
    g = po.ClassicalOracle(edges, [0.0] * 16, key=key, state=[0] * 16)
    e0 = g.energy()
    g.worm_flips(1000.0, 1, allow_doubles=True)
    assert g.energy() == e0  # a closed worm returns to the initial energy (or is undone)
    edges = lattices.bathroom_unit_cells(16)
    g = po.ClassicalOracle(edges, [0.0] * 1024, key=key, state=[0] * 1024)
    e0 = g.energy()
    g.worm_flips(1000.0, 1, allow_doubles=True)
    assert g.energy() == e0


def test_edge_flip_known_answer():
    # two aligned spins joined by J = -1 (aligned is favoured), no bias: flipping both costs nothing -> always flips,
    # one usize draw for the edge and no acceptance draw (graph.rs:140-152, :341)
    g = po.ClassicalOracle([((0, 1), -1.0), ((1, 2), -1.0)], [0.0, 0.0, 0.0], key=5, state=[1, 1, 0])
    c0 = g.cursor
    g.edge_flips(1.0, 1)
    st = list(g.state())
    assert st in ([0, 0, 0], [1, 0, 1])
    # edge (0,1): delta = d(0 omit 1) + d(1 omit 0) = 0 + (-2*-1*-1 = -2) <= 0 -> flip, no acceptance draw;
    # edge (1,2): d(1 omit 2) = -2*-1*+1 = +2, d(2 omit 1) = 0 -> +2 > 0 -> one f64 draw
    assert g.cursor - c0 in (1, 2)


def test_edge_importance_sampling_binary_search():
    # weights 1, 3 -> cumulative [1, 4]; p in [0, 4): edge 0 iff p <= 1 (binary_search_by: Err(0) for p < 1)
    edges = [((0, 1), 1.0), ((2, 3), 3.0)]
    hits = [0, 0]
    for key in range(400):
        g = po.ClassicalOracle(edges, [0.0] * 4, key=key, state=[0, 1, 0, 1])  # anti-aligned, J > 0: flipping both is free
        g.enable_edge_importance_sampling(True)
        g.edge_flips(1.0, 1)
        st = list(g.state())
        assert st in ([1, 0, 0, 1], [0, 1, 1, 0])
        hits[0 if st == [1, 0, 0, 1] else 1] += 1
    assert abs(hits[0] / 400 - 0.25) < 0.08


def test_do_time_step_defaults_and_draws():
    edges = lattices.square_periodic(4, -1.0)
    seen = set()
    for key in range(30):
        g = po.ClassicalOracle(edges, [0.1] * 16, key=key)
        assert g.cursor == 16  # make_random_spin_state: one word per spin
        seen.add(g.do_time_step(0.4))
        assert g.error == 0
    assert seen == {0, 1, 2}
    seen = {po.ClassicalOracle(edges, [0.1] * 16, key=k).do_time_step(0.4, only_basic_moves=True) for k in range(30)}
    assert seen == {0, 1}


def test_spin_and_edge_moves_sample_boltzmann():
    # 5 spins, frustrated couplings and biases: exact distribution by enumeration
    edges = [((0, 1), 1.0), ((1, 2), -0.5), ((2, 3), 1.0), ((3, 4), 0.7), ((4, 0), 1.0), ((1, 3), -0.3)]
    biases = [0.2, -0.1, 0.0, 0.3, -0.2]
    beta = 0.7
    g = po.ClassicalOracle(edges, biases, key=11)
    states = list(itertools.product([0, 1], repeat=5))
    en = []
    for st in states:
        g.set_state(st)
        en.append(g.energy())
    w = np.exp(-beta * np.array(en))
    p = w / w.sum()
    counts = np.zeros(32)
    nsamp = 60000
    for _ in range(nsamp):
        g.do_time_step(beta, nspinupdates=3, nedgeupdates=2, only_basic_moves=True)
        counts[int("".join(map(str, g.state())), 2)] += 1
    # samples are correlated; a loose 5-sigma band on every state with an inflation factor for autocorrelation
    sig = np.sqrt(p * (1 - p) / nsamp) * 3.0
    assert np.all(np.abs(counts / nsamp - p) < 5 * sig + 1e-3)
