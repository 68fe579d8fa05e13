"""QMCB_MODE_COUNTER (FAST cluster order + one Philox block per slot for the diagonal update, sse_counter.cu) against the
oracle's COUNTER mode (oracle.c diagonal_update_counter): bit-exact operator strings, states, n, cutoff, cursor and
energies on the small cases, random systems, the full-size lattices and through tempering; every compiled kernel
variant; the serial on-device restatement ("impl" 1); production shapes (R = 4096 in one wave)."""
import numpy as np
import pytest

from isingmontecarlo_b200 import MODE_COUNTER, MODE_FAST, QmcbError, lattices
from oracle import pyoracle as po
from tests.test_gpu_full_size import CONFIGS, same, to_oracle
from tests.test_gpu_random_systems import random_system
from tests.test_gpu_sse_parity import CASES, assert_same, make_pair

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("impl", [0, 1])
@pytest.mark.parametrize("name,edges,gamma,h,cutoff,beta,sweeps", CASES)
def test_sweeps_bit_exact(name, edges, gamma, h, cutoff, beta, sweeps, impl):
    g, refs = make_pair(edges, gamma, h, cutoff, beta, MODE_COUNTER, impl=impl)
    for chunk in (1, 1, 3, sweeps - 5):
        e_gpu = g.timesteps(chunk, beta)
        e_ref = [ref.timesteps(chunk, beta, MODE_COUNTER) for ref in refs]
        assert_same(g, refs, f"{name} after +{chunk}")
        assert np.array_equal(e_gpu, np.array(e_ref)), name
    assert g.verify() and all(ref.verify() for ref in refs)


def test_single_steps_and_cluster_counts():
    edges = lattices.two_d_periodic_mixed(4)
    g, refs = make_pair(edges, 1.0, 0.5, 16, 1.5, MODE_COUNTER, R=4)
    for _ in range(6):
        g.single_diagonal_step(1.5)
        [ref.single_diagonal_step(1.5, MODE_COUNTER) for ref in refs]
        assert_same(g, refs, "diag")
        ncl = g.single_cluster_step()
        assert [int(x) for x in ncl] == [ref.single_cluster_step(MODE_COUNTER) for ref in refs]
        assert_same(g, refs, "cluster")


@pytest.mark.parametrize("seed", range(16))
def test_random_systems_bit_exact(seed):
    from isingmontecarlo_b200.sse import QmcIsingGraph

    rng = np.random.default_rng(7000 + seed)
    edges, nv, gamma, h, beta, cutoff = random_system(rng)
    keys = [int(k) for k in rng.integers(1, 2**62, size=4)]
    g = QmcIsingGraph(edges, gamma, h, cutoff, keys, beta, mode=MODE_COUNTER)
    g.set_option("impl", 1 if seed % 8 == 7 else 0)
    refs = [po.SseOracle(edges, gamma, h, cutoff, key=k) for k in keys]
    for chunk in (1, 2, 5, 12):
        e = g.timesteps(chunk, beta)
        e_ref = np.array([ref.timesteps(chunk, beta, MODE_COUNTER) for ref in refs])
        assert_same(g, refs, f"seed {seed} +{chunk}")
        assert np.array_equal(e, e_ref)
    assert g.verify() and all(ref.verify() and ref.error == 0 for ref in refs)


@pytest.mark.parametrize("name,mk,gamma,h,beta,cutoff,R,therm", CONFIGS)
def test_full_size_parity_from_thermalised_state(name, mk, gamma, h, beta, cutoff, R, therm):
    from isingmontecarlo_b200.sse import QmcIsingGraph

    edges = mk()
    g = QmcIsingGraph(edges, gamma, h, cutoff, 0x55E00000 + np.arange(R, dtype=np.uint64), beta, mode=MODE_COUNTER)
    g.timesteps(therm, beta)
    assert g.verify()
    n0 = g.get_n()
    assert n0.min() > 0.5 * n0.max() > 1000
    refs = {r: to_oracle(g, r, edges, gamma, h) for r in (0, R - 1)}
    e = g.timesteps(3, beta)
    for r, ref in refs.items():
        e_ref = ref.timesteps(3, beta, MODE_COUNTER)
        assert ref.error == 0 and same(g, r, ref), (name, r)
        assert e[r] == e_ref
    assert g.verify()


def test_small_beta_regime_takes_the_division_path():
    # high temperature: num < den for every bond, so every insertion goes through the reciprocal bounds
    from isingmontecarlo_b200.sse import QmcIsingGraph

    edges = lattices.square_periodic(8, -1.0)
    g = QmcIsingGraph(edges, 0.8, 0.0, 64, [7, 8, 9], 0.05, mode=MODE_COUNTER, capacity=4096)
    refs = [po.SseOracle(edges, 0.8, 0.0, 64, key=k) for k in (7, 8, 9)]
    for chunk in (1, 5, 20):
        g.timesteps(chunk, 0.05)
        for r, ref in enumerate(refs):
            ref.timesteps(chunk, 0.05, MODE_COUNTER)
            assert same(g, r, ref), (chunk, r)


@pytest.mark.parametrize("minblocks,shared_edges,pipeline", [(7, 1, 1), (7, 0, 1), (4, 0, 1), (4, 1, 1), (4, 0, 2), (4, 1, 2), (4, 0, 0), (4, 1, 0)])
def test_every_kernel_build_is_bit_exact(minblocks, shared_edges, pipeline):
    # minblocks 4 with the pipeline on: three warps per replica (1) or two (2); small batches pick it by themselves
    from isingmontecarlo_b200.sse import QmcIsingGraph

    edges = lattices.square_periodic(32, -1.0)
    R = 12
    g = QmcIsingGraph(edges, 3.04, 0.0, 1024, 0x0B100000 + np.arange(R, dtype=np.uint64), 8.0, mode=MODE_COUNTER)
    g.set_option("minblocks", minblocks)
    g.set_option("shared_edge_table", shared_edges)
    g.set_option("pipeline", pipeline)
    g.timesteps(30, 8.0)
    refs = {r: to_oracle(g, r, edges, 3.04, 0.0) for r in (0, 5, R - 1)}
    e = g.timesteps(3, 8.0)
    for r, ref in refs.items():
        e_ref = ref.timesteps(3, 8.0, MODE_COUNTER)
        assert ref.error == 0 and same(g, r, ref), (minblocks, shared_edges, r)
        assert e[r] == e_ref
    assert g.verify()


def test_knobs_are_per_handle():
    # qmcb_set_option on one handle must not change the kernel choice of another (process-wide globals)
    from isingmontecarlo_b200.sse import QmcIsingGraph

    edges = lattices.square_periodic(8, -1.0)
    a = QmcIsingGraph(edges, 3.04, 0.0, 64, [1, 2], 2.0, mode=MODE_COUNTER)
    b = QmcIsingGraph(edges, 3.04, 0.0, 64, [1, 2], 2.0, mode=MODE_COUNTER)
    a.set_option("minblocks", 4), a.set_option("shared_edge_table", 0), a.set_option("smem_pad", 10**9)  # pad is clamped
    a.timesteps(10, 2.0), b.timesteps(10, 2.0)
    for r in range(2):
        assert np.array_equal(a.dump_ops(r), b.dump_ops(r))
    with pytest.raises(QmcbError):
        a.set_option("minblocks", 5)


def test_heatbath_has_no_counter_contract():
    g, _ = make_pair(lattices.small_qmc_ring(), 1.0, 0.0, 3, 1.0, MODE_COUNTER, R=2)
    with pytest.raises(QmcbError, match="COUNTER"):
        g.set_enable_heatbath(True)
    g.set_mode(MODE_FAST)
    g.set_enable_heatbath(True)
    with pytest.raises(QmcbError, match="COUNTER"):
        g.set_mode(MODE_COUNTER)


def test_tempering_matches_reference_swaps():
    from isingmontecarlo_b200.tempering import TemperingContainer

    edges = lattices.two_d_periodic_mixed(4)
    n_chains, n_betas = 2, 5
    betas = np.linspace(0.5, 2.0, n_betas)
    S = n_chains * n_betas
    keys = 0x55E40000 + np.arange(S, dtype=np.uint64)
    tc = TemperingContainer(edges, 1.0, 0.0, 16, betas, n_chains=n_chains, rng_keys=keys, pt_key=0xABCDEF, mode=MODE_COUNTER)
    slots = [[po.SseOracle(edges, 1.0, 0.0, 16, key=int(keys[c * n_betas + k])) for k in range(n_betas)] for c in range(n_chains)]
    cursors, swaps_ref = [0] * n_chains, 0
    for step in range(10):
        tc.timesteps(3)
        for c in range(n_chains):
            for k in range(n_betas):
                slots[c][k].timesteps(3, float(betas[k]), MODE_COUNTER)
        tc.tempering_step()
        for c in range(n_chains):
            s, cursors[c] = po.pt_step(slots[c], betas, 0xABCDEF + c, cursors[c])
            swaps_ref += s
        g = tc.graph
        n, cut, cur, st = g.get_n(), g.get_cutoff(), g.rng_cursors(), g.state_ref()
        for s_local, slot in enumerate(tc.slots()):
            ref = slots[slot // n_betas][slot % n_betas]
            assert int(n[s_local]) == ref.n and int(cut[s_local]) == ref.cutoff and int(cur[s_local]) == ref.cursor
            assert np.array_equal(st[s_local], ref.state()) and np.array_equal(g.dump_ops(s_local), ref.dump_ops())
        assert tc.get_total_swaps() == swaps_ref
    assert swaps_ref > 0 and tc.verify()


def test_per_replica_hamiltonians():
    from isingmontecarlo_b200.sse import QmcIsingGraph

    edges = lattices.two_d_periodic_mixed(4)
    J0 = np.array([j for _, j in edges])
    for minblocks in (7, 4):
        g = QmcIsingGraph(edges, 1.0, 0.5, 16, [11, 12, 13], 1.0, mode=MODE_COUNTER)
        g.set_option("minblocks", minblocks)
        g.set_hamiltonians([J0, 2 * J0, 0.5 * J0], [1.0, 3.0, 0.7], [0.5, 0.25, 0.9], [1, 0, 2])
        rows = [(2 * J0, 3.0, 0.25), (J0, 1.0, 0.5), (0.5 * J0, 0.7, 0.9)]
        refs = [po.SseOracle([(e, float(j)) for (e, _), j in zip(edges, jr)], gam, hl, 16, key=k) for k, (jr, gam, hl) in zip([11, 12, 13], rows)]
        e = g.timesteps(25, 1.0)
        for r, ref in enumerate(refs):
            assert e[r] == ref.timesteps(25, 1.0, MODE_COUNTER)
            assert np.array_equal(g.dump_ops(r), ref.dump_ops()) and np.array_equal(g.state_ref()[r], ref.state())


@pytest.mark.parametrize("mode", [MODE_COUNTER, MODE_FAST])
def test_production_shape_one_wave(mode):
    """Config #3 at the bench's own shape: R = 4096 replicas (28 warps per SM, the 72-register build chosen by the
    launcher, one wave), 20 sweeps after thermalisation; 8 randomly chosen replicas are diffed against the oracle
    from the thermalised dump and the reference invariant is replayed on every 64th replica."""
    from isingmontecarlo_b200.sse import QmcIsingGraph

    edges = lattices.square_periodic(32, -1.0)
    R = 4096
    g = QmcIsingGraph(edges, 3.04, 0.0, 1024, 0x55E00000 + np.arange(R, dtype=np.uint64), 16.0, mode=mode)
    g.timesteps(40, 16.0)
    picks = sorted(int(x) for x in np.random.default_rng(42).choice(R, size=8, replace=False))
    refs = {r: to_oracle(g, r, edges, 3.04, 0.0) for r in picks}
    e = g.timesteps(20, 16.0)
    for r, ref in refs.items():
        e_ref = ref.timesteps(20, 16.0, mode)
        assert ref.error == 0 and same(g, r, ref), r
        assert e[r] == e_ref
    assert all(g.verify(r) for r in range(0, R, 64))


def test_production_shape_two_warp_build():
    """R = 512 replicas of config #3: at most one block of four replicas per SM is wanted, so the FAST launcher picks the
    two-warps-per-replica (PIPE) 120-register build; same check as above."""
    from isingmontecarlo_b200.sse import QmcIsingGraph

    edges = lattices.square_periodic(32, -1.0)
    R = 512
    g = QmcIsingGraph(edges, 3.04, 0.0, 1024, 0x55E00000 + np.arange(R, dtype=np.uint64), 16.0, mode=MODE_FAST)
    g.timesteps(40, 16.0)
    picks = sorted(int(x) for x in np.random.default_rng(43).choice(R, size=8, replace=False))
    refs = {r: to_oracle(g, r, edges, 3.04, 0.0) for r in picks}
    e = g.timesteps(20, 16.0)
    for r, ref in refs.items():
        e_ref = ref.timesteps(20, 16.0, MODE_FAST)
        assert ref.error == 0 and same(g, r, ref), r
        assert e[r] == e_ref
    assert all(g.verify(r) for r in range(0, R, 16))


@pytest.mark.parametrize("h,R,pipeline", [(0.0, 512, 1), (0.3, 300, 1), (0.0, 512, 2), (0.3, 300, 2)])
def test_production_shape_two_warp_counter_build(h, R, pipeline):
    """Few replicas (at most two blocks of four per SM): the COUNTER launcher picks the two-warps-per-replica build -- role A
    the diagonal update of step k, role B segments and unions of step k - 1, P3 split in halves -- with and without a
    longitudinal field; sweeps in one launch and one by one; against the oracle from the thermalised state, and against
    the one-warp build of the same kernel."""
    from isingmontecarlo_b200.sse import QmcIsingGraph

    edges = lattices.square_periodic(32, -1.0)
    keys = 0x55E10000 + np.arange(R, dtype=np.uint64)
    g = QmcIsingGraph(edges, 3.04, h, 1024, keys, 8.0, mode=MODE_COUNTER)
    g.set_option("pipeline", pipeline)  # 1: role C takes the union-find as well; 2: two roles
    one = QmcIsingGraph(edges, 3.04, h, 1024, keys, 8.0, mode=MODE_COUNTER)
    one.set_option("pipeline", 0)
    g.timesteps(30, 8.0), one.timesteps(30, 8.0)
    picks = sorted(int(x) for x in np.random.default_rng(44).choice(R, size=6, replace=False))
    refs = {r: to_oracle(g, r, edges, 3.04, h) for r in picks}
    e = g.timesteps(8, 8.0)
    for _ in range(3):
        g.timesteps(1, 8.0)
    eo = one.timesteps(11, 8.0)
    for r, ref in refs.items():
        e_ref = ref.timesteps(8, 8.0, MODE_COUNTER)
        assert e[r] == e_ref
        ref.timesteps(3, 8.0, MODE_COUNTER)
        assert ref.error == 0 and same(g, r, ref), r
    assert np.array_equal(g.state_ref(), one.state_ref()) and np.array_equal(g.get_n(), one.get_n())
    assert np.array_equal(g.rng_cursors(), one.rng_cursors()) and np.array_equal(g.get_cutoff(), one.get_cutoff())
    for r in picks:
        assert np.array_equal(g.dump_ops(r), one.dump_ops(r))
    assert all(g.verify(r) for r in range(0, R, 16))


@pytest.mark.parametrize("pipeline", [0, 1, 2])
def test_capacity_growth_in_every_build(pipeline):
    """A sweep that finds its cutoff above the capacity stops that replica (every role of the multi-warp builds leaves
    the sweep at the same place), the host re-lays the strings out and relaunches: results do not depend on where growth
    happened."""
    from isingmontecarlo_b200.sse import QmcIsingGraph

    edges = lattices.square_periodic(6, -1.0)
    keys = [0xCA9A0 + r for r in range(5)]
    g = QmcIsingGraph(edges, 2.5, 0.0, 36, keys, 4.0, mode=MODE_COUNTER, capacity=64)
    g.set_option("auto_capacity", 1)
    g.set_option("minblocks", 4)
    g.set_option("pipeline", pipeline)
    refs = [po.SseOracle(edges, 2.5, 0.0, 36, key=k) for k in keys]
    for chunk in (7, 1, 12):
        e = g.timesteps(chunk, 4.0)
        for r, ref in enumerate(refs):
            assert e[r] == ref.timesteps(chunk, 4.0, MODE_COUNTER)
            assert same(g, r, ref), (pipeline, chunk, r)
    assert g.get_capacity() > 64 and g.verify()
