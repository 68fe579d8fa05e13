"""Checkpoint round trips (replaces the reference's serde path, qmc_ising.rs:1001-1087 and
tempering_container.rs:671-793, whose own test is `serialize_test`: save, restore, compare).  The property:
a restored batch continues BIT-IDENTICALLY to the one that was saved."""
import numpy as np
import pytest

from isingmontecarlo_b200 import MODE_FAST, MODE_STRICT, QmcbError, lattices

pytestmark = pytest.mark.gpu


def snapshot(g):
    return (g.get_n().copy(), g.get_cutoff().copy(), g.rng_cursors().copy(), g.state_ref().copy(),
            [g.dump_ops(r).copy() for r in range(g.R)])


def same_snapshot(a, b):
    return all(np.array_equal(x, y) for x, y in zip(a[:4], b[:4])) and all(np.array_equal(x, y) for x, y in zip(a[4], b[4]))


@pytest.mark.parametrize("mode,heatbath", [(MODE_FAST, False), (MODE_STRICT, False), (MODE_FAST, True)])
def test_checkpoint_round_trip_continues_bit_identically(mode, heatbath):
    from isingmontecarlo_b200.sse import QmcIsingGraph

    edges = lattices.triangular_periodic(6, 1.0)
    betas = np.linspace(0.5, 3.0, 6)
    g = QmcIsingGraph(edges, 1.0, 0.2, 36, 0xC0FFEE00 + np.arange(6, dtype=np.uint64), betas, mode=mode)
    g.set_enable_heatbath(heatbath)
    g.timesteps(25)
    blob = g.save_checkpoint()
    e_a = g.timesteps(10)
    snap_a = snapshot(g)
    g2 = QmcIsingGraph.from_checkpoint(blob)
    assert g2.R == 6 and g2.nvars == 36 and g2.mode == mode and g2.get_enable_heatbath() == heatbath
    assert g2.get_edges() == [((int(a), int(b)), float(j)) for (a, b), j in edges]
    assert np.array_equal(g2.betas(), betas) and g2.verify()
    e_b = g2.timesteps(10)
    assert np.array_equal(e_a, e_b)
    assert same_snapshot(snap_a, snapshot(g2))
    assert g2.save_checkpoint() == g.save_checkpoint()  # same state -> same bytes


def test_checkpoint_carries_the_rvb_flag():
    # run_rvb_steps is part of the reference's serialised graph (qmc_ising.rs:1020, 1044): a restored batch keeps running RVB steps
    from isingmontecarlo_b200.sse import QmcIsingGraph

    edges = lattices.two_unit_cell()
    g = QmcIsingGraph(edges, 1.0, 0.0, 8, 0xC0FFEE40 + np.arange(4, dtype=np.uint64), 2.0, mode=MODE_FAST)
    g.set_run_rvb(True)
    g.timesteps(20)
    blob = g.save_checkpoint()
    e_a = g.timesteps(10)
    g2 = QmcIsingGraph.from_checkpoint(blob)
    e_b = g2.timesteps(10)
    assert np.array_equal(e_a, e_b) and same_snapshot(snapshot(g), snapshot(g2)) and g2.verify()
    assert (g2.rvb_success_rate() > 0).all()  # counted from the restore on


def test_checkpoint_rejects_corruption_and_truncation():
    from isingmontecarlo_b200.sse import QmcIsingGraph

    g = QmcIsingGraph(lattices.small_qmc_ring(), 1.0, 0.0, 4, [1, 2], 1.0, mode=MODE_FAST)
    g.timesteps(5)
    blob = bytearray(g.save_checkpoint())
    for bad in (bytes(blob[:-9]), bytes(blob[:40]), b"NOTACKPT" + bytes(blob[8:])):
        with pytest.raises(QmcbError):
            QmcIsingGraph.from_checkpoint(bad)
    blob[len(blob) // 2] ^= 0x40
    with pytest.raises(QmcbError):
        QmcIsingGraph.from_checkpoint(bytes(blob))


@pytest.mark.parametrize("unequal", [False, True])
def test_tempering_checkpoint_round_trip(unequal):
    import torch

    from isingmontecarlo_b200.tempering import TemperingContainer

    torch.cuda.set_device(0)
    edges = lattices.square_periodic(4, -1.0)
    betas = np.geomspace(0.3, 3.0, 6)
    hams = [(None, 2.0 + 0.1 * k, 0.0) for k in range(6)] if unequal else None
    tc = TemperingContainer(edges, 2.0, 0.0, 16, betas, n_chains=2, pt_key=0xABCD, mode=MODE_FAST, slot_hamiltonians=hams)
    for _ in range(12):
        tc.timesteps(2)
        tc.tempering_step()
    blob = tc.save_checkpoint()

    def cont(c):
        for _ in range(8):
            c.timesteps(2)
            c.tempering_step()
        return snapshot(c.graph), c.slots().copy(), c.get_total_swaps(), c.graph.betas().copy(), c.graph.hamiltonian_index().copy()

    a = cont(tc)
    tc2 = TemperingContainer.from_checkpoint(blob)
    assert tc2.S == 12 and tc2.n_betas == 6
    b = cont(tc2)
    assert a[2] == b[2] and a[2] > 0
    assert np.array_equal(a[1], b[1]) and np.array_equal(a[3], b[3]) and np.array_equal(a[4], b[4])
    assert unequal == (len(set(a[4])) > 1)
    assert same_snapshot(a[0], b[0])
