"""GPU tests at BASELINE.json's full lattice sizes (fewer replicas so they run in seconds):
bit-exact parity with the oracle from a thermalised state, and the size-independent invariants
the domain offers -- the reference's verify() (op_container.rs:137-159), dump -> load -> continue
round trips, mode-independence of the diagonal update."""
import numpy as np
import pytest

from isingmontecarlo_b200 import MODE_FAST, MODE_STRICT, lattices
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu


def to_oracle(g, r, edges, gamma, h):
    """copy replica r of a GPU handle into a fresh oracle graph (ops, state, cutoff, stream)"""
    ref = po.SseOracle(edges, gamma, h, int(g.get_cutoff()[r]), key=int(g.rng_keys()[r]), state=g.state_ref()[r])
    ref.load_ops(g.dump_ops(r), g.state_ref()[r])
    ref.set_cursor(int(g.rng_cursors()[r]))
    return ref


def same(g, r, ref):
    return (int(g.get_n()[r]) == ref.n and int(g.get_cutoff()[r]) == ref.cutoff and int(g.rng_cursors()[r]) == ref.cursor
            and np.array_equal(g.state_ref()[r], ref.state()) and np.array_equal(g.dump_ops(r), ref.dump_ops()))


CONFIGS = [
    # BASELINE config #3: square L=32, J=-1, Gamma=3.04, beta=16
    ("cfg3", lambda: lattices.square_periodic(32, -1.0), 3.04, 0.0, 16.0, 1024, 24, 60),
    # config #4 lattice: square L=64 at one of the ladder's betas
    ("cfg4", lambda: lattices.square_periodic(64, -1.0), 3.04, 0.0, 4.0, 4096, 8, 40),
    # config #5: frustrated triangular L=48 with longitudinal field (Gamma=1.0, h=0.2)
    ("cfg5", lambda: lattices.triangular_periodic(48, 1.0), 1.0, 0.2, 8.0, 2304, 8, 40),
]


@pytest.mark.parametrize("mode", [MODE_STRICT, MODE_FAST])
@pytest.mark.parametrize("name,mk,gamma,h,beta,cutoff,R,therm", CONFIGS)
def test_full_size_parity_from_thermalised_state(name, mk, gamma, h, beta, cutoff, R, therm, mode):
    from isingmontecarlo_b200.sse import QmcIsingGraph

    edges = mk()
    g = QmcIsingGraph(edges, gamma, h, cutoff, 0x55E00000 + np.arange(R, dtype=np.uint64), beta, mode=MODE_FAST)
    g.timesteps(therm, beta)  # thermalise on the GPU
    assert g.verify()         # reference invariant on every replica at full size
    n0 = g.get_n()
    assert n0.min() > 0.5 * n0.max() > 1000
    g.set_mode(mode)
    refs = {r: to_oracle(g, r, edges, gamma, h) for r in (0, R - 1)}
    for r, ref in refs.items():
        assert ref.verify() and same(g, r, ref)
    e = g.timesteps(3, beta)
    for r, ref in refs.items():
        e_ref = ref.timesteps(3, beta, mode)
        assert ref.error == 0 and same(g, r, ref), (name, r)
        assert e[r] == e_ref
    assert g.verify()


def test_diagonal_update_does_not_depend_on_the_cluster_order():
    # the diagonal update is the reference rule in both modes: from equal states a single diagonal step
    # through the warp-parallel kernel (FAST) and the serial one (STRICT) must agree bit for bit
    from isingmontecarlo_b200.sse import QmcIsingGraph

    edges = lattices.square_periodic(32, -1.0)
    keys = 0x55E00000 + np.arange(6, dtype=np.uint64)
    a = QmcIsingGraph(edges, 3.04, 0.0, 1024, keys, 16.0, mode=MODE_FAST)
    b = QmcIsingGraph(edges, 3.04, 0.0, 1024, keys, 16.0, mode=MODE_FAST)
    a.timesteps(30, 16.0), b.timesteps(30, 16.0)
    b.set_mode(MODE_STRICT)
    for _ in range(3):
        a.single_diagonal_step(16.0), b.single_diagonal_step(16.0)
        assert np.array_equal(a.get_n(), b.get_n()) and np.array_equal(a.rng_cursors(), b.rng_cursors())
        for r in range(6):
            assert np.array_equal(a.dump_ops(r), b.dump_ops(r))


def test_small_beta_regime_takes_the_second_word_paths():
    # high temperature: num < den for every bond, so every empty slot reads a second word and most
    # diagonal ops are removed without a draw -- the opposite regime of config #3
    from isingmontecarlo_b200.sse import QmcIsingGraph

    edges = lattices.square_periodic(8, -1.0)
    for mode in (MODE_STRICT, MODE_FAST):
        g = QmcIsingGraph(edges, 0.8, 0.0, 64, [7, 8, 9], 0.05, mode=mode, capacity=4096)
        refs = [po.SseOracle(edges, 0.8, 0.0, 64, key=k) for k in (7, 8, 9)]
        for chunk in (1, 5, 20):
            g.timesteps(chunk, 0.05)
            for r, ref in enumerate(refs):
                ref.timesteps(chunk, 0.05, mode)
                assert same(g, r, ref), (mode, chunk, r)


@pytest.mark.parametrize("heatbath", [False, True])
@pytest.mark.parametrize("minblocks,shared_edges,pipeline", [(7, 1, 1), (7, 0, 1), (8, 1, 1), (6, 1, 1), (4, 0, 1), (4, 1, 1), (4, 0, 0), (4, 1, 0)])
def test_every_kernel_build_is_bit_exact(minblocks, shared_edges, pipeline, heatbath):
    """The launcher picks the register budget (72-register build for many replicas, 120-register build with the
    shared-memory edge table when few blocks are resident) from the batch shape; tests have few replicas, so force
    each compiled variant in turn and check it against the oracle from a thermalised config #3 state."""
    from isingmontecarlo_b200.sse import QmcIsingGraph

    edges = lattices.square_periodic(32, -1.0)
    R = 12
    g = QmcIsingGraph(edges, 3.04, 0.0, 1024, 0x0B100000 + np.arange(R, dtype=np.uint64), 8.0, mode=MODE_FAST)
    try:
        g.set_option("minblocks", minblocks)
        g.set_option("shared_edge_table", shared_edges)
        g.set_option("pipeline", pipeline)
        g.set_enable_heatbath(heatbath)
        g.timesteps(30, 8.0)
        refs = {r: to_oracle(g, r, edges, 3.04, 0.0) for r in (0, 5, R - 1)}
        for ref in refs.values():
            ref.set_enable_heatbath(heatbath)
        e = g.timesteps(3, 8.0)
        for r, ref in refs.items():
            e_ref = ref.timesteps(3, 8.0, MODE_FAST)
            assert ref.error == 0 and same(g, r, ref), (minblocks, shared_edges, r)
            assert e[r] == e_ref
        assert g.verify()
    finally:
        g.set_option("minblocks", 0)  # process-wide tuning knobs: back to automatic
        g.set_option("shared_edge_table", 1)
        g.set_option("pipeline", 1)
