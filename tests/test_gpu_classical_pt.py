"""GPU parity: classical checkerboard sweeps and the tempering swap step against the CPU oracle."""
import numpy as np
import pytest

from isingmontecarlo_b200 import MODE_FAST, MODE_STRICT, lattices
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu


def run_classical(edges, biases, betas, sweeps, keys, state=None):
    from isingmontecarlo_b200.classical import GraphState

    g = GraphState(edges, biases, keys, betas, state=state)
    colours, ncol = g.colours()
    refs = [po.ClassicalOracle(edges, biases, key=k, state=None if state is None else state) for k in keys]
    st0 = g.state_ref()
    for r, ref in enumerate(refs):
        assert np.array_equal(st0[r], ref.state()), "stream-drawn initial state"
    for chunk in (1, 2, sweeps - 3):
        g.sweeps(chunk)
        st = g.state_ref()
        e, m = g.get_energy(), g.magnetization()
        for r, ref in enumerate(refs):
            ref.checkerboard_sweeps(betas[r], colours, chunk)
            assert np.array_equal(st[r], ref.state()), (r, chunk)
            assert abs(e[r] - ref.energy()) <= 1e-9 * max(1.0, abs(ref.energy()))
            assert abs(m[r] - ref.magnetization()) <= 1e-12
    return g


def test_generic_graphs_bit_exact():
    rng = np.random.default_rng(3)
    tri = lattices.triangular_periodic(6, 1.0)
    g = run_classical(tri, np.zeros(36), [0.3, 0.6, 1.2], 8, [11, 12, 13])
    assert not g.is_bitpacked_square() and g.colours()[1] >= 3
    # random couplings and biases on a small square lattice (odd L => 3 colours)
    sq = [(e, float(rng.normal())) for e, _ in lattices.square_periodic(5, 1.0)]
    run_classical(sq, rng.normal(size=25), [0.5, 1.0], 8, [21, 22])
    # ring with a uniform bias, generic path (too small for the bit-packed layout)
    run_classical(lattices.one_d_periodic(10, -1.0), np.full(10, 0.25), [0.7, 0.7], 8, [31, 32])


@pytest.mark.parametrize("J,bias", [(-1.0, 0.0), (1.0, 0.0), (-1.0, 0.3), (0.75, -0.2)])
def test_square_bitpacked_bit_exact(J, bias):
    L = 64
    edges = lattices.square_periodic(L, J)
    g = run_classical(edges, np.full(L * L, bias), [0.4406868, 0.3, 0.8], 7, [0xB2000000, 0xB2000001, 0xB2000002])
    assert g.is_bitpacked_square() and g.colours()[1] == 2


def test_square_set_state_round_trip_and_energy_anchor():
    from isingmontecarlo_b200.classical import GraphState

    L = 64
    edges = lattices.square_periodic(L, -1.0)
    rng = np.random.default_rng(0)
    st = rng.integers(0, 2, size=(2, L * L), dtype=np.uint8)
    g = GraphState(edges, np.zeros(L * L), [1, 2], 0.44, state=st)
    assert np.array_equal(g.state_ref(), st)
    g.set_state(np.zeros((2, L * L), dtype=np.uint8))
    assert np.all(g.get_energy() == -2.0 * L * L)  # all aligned, J=-1: E = -2N (graph.rs:430-447)
    assert np.all(g.magnetization() == -1.0)


def onsager_energy(beta):
    """exact energy per site of the infinite square-lattice Ising ferromagnet (|J| = 1)"""
    from scipy.special import ellipk

    k = 2.0 * np.sinh(2 * beta) / np.cosh(2 * beta) ** 2
    return -1.0 / np.tanh(2 * beta) * (1.0 + (2.0 / np.pi) * (2.0 * np.tanh(2 * beta) ** 2 - 1.0) * ellipk(k * k))


@pytest.mark.parametrize("beta", [0.35, 0.55])
def test_classical_energy_matches_onsager(beta):
    # physics anchor away from T_c (correlation length << L, so finite-size effects are negligible):
    # mean energy over independent replicas within 3 sigma (+ a 1e-3 systematic allowance)
    from isingmontecarlo_b200.classical import GraphState

    L, R = 64, 64
    edges = lattices.square_periodic(L, -1.0)
    state = None if beta < 0.44 else np.zeros((R, L * L), dtype=np.uint8)
    g = GraphState(edges, np.zeros(L * L), 0xB2000000 + np.arange(R), beta, state=state)
    g.sweeps(500)
    es = []
    for _ in range(40):
        g.sweeps(10)
        es.append(g.get_energy() / (L * L))
    e = np.mean(es, axis=0)
    err = e.std(ddof=1) / np.sqrt(R)
    assert abs(e.mean() - onsager_energy(beta)) < 3.0 * err + 1e-3, (e.mean(), err, onsager_energy(beta))


@pytest.mark.parametrize("mode", [MODE_STRICT, MODE_FAST])
@pytest.mark.parametrize("n_chains,n_betas", [(1, 2), (2, 5), (3, 4)])
def test_tempering_matches_reference_swaps(n_chains, n_betas, mode):
    from isingmontecarlo_b200.tempering import TemperingContainer

    edges = lattices.two_d_periodic_mixed(4)
    betas = np.linspace(0.5, 2.0, n_betas)
    S = n_chains * n_betas
    keys = 0x55E00000 + np.arange(S, dtype=np.uint64)
    pt_key = 0xABCDEF
    tc = TemperingContainer(edges, 1.0, 0.0, 16, betas, n_chains=n_chains, rng_keys=keys, pt_key=pt_key, mode=mode)
    # reference: one TemperingContainer (one PT stream, key pt_key + chain) per ladder
    slots = [[po.SseOracle(edges, 1.0, 0.0, 16, key=int(keys[c * n_betas + k])) for k in range(n_betas)] for c in range(n_chains)]
    cursors = [0] * n_chains
    swaps_ref = 0
    for step in range(12):
        tc.timesteps(3)
        for c in range(n_chains):
            for k in range(n_betas):
                slots[c][k].timesteps(3, float(betas[k]), mode)
        tc.tempering_step()
        for c in range(n_chains):
            s, cursors[c] = po.pt_step(slots[c], betas, pt_key + c, cursors[c])
            swaps_ref += s
        # compare slot by slot: the configuration currently labelled `slot` on the GPU must equal
        # the reference graph sitting in that slot
        g = tc.graph
        slot_of = tc.slots()
        n, cut, cur, st = g.get_n(), g.get_cutoff(), g.rng_cursors(), g.state_ref()
        for s_local, slot in enumerate(slot_of):
            ref = slots[slot // n_betas][slot % n_betas]
            assert int(n[s_local]) == ref.n and int(cut[s_local]) == ref.cutoff and int(cur[s_local]) == ref.cursor
            assert np.array_equal(st[s_local], ref.state())
            assert np.array_equal(g.dump_ops(s_local), ref.dump_ops())
        assert tc.get_total_swaps() == swaps_ref
    assert (swaps_ref > 0 or n_betas <= 2) and tc.verify()


def test_tempering_over_nccl_two_ranks():
    # multi-GPU path for real: one process per GPU, all-gather of the slot records over NCCL
    import os
    import subprocess
    import sys

    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", "29541", os.path.join(root, "tests", "pt_worker.py"), root],
                         capture_output=True, text=True, timeout=600, env=dict(os.environ, MASTER_ADDR="127.0.0.1"))
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-3000:]
    assert res.stdout.count(" ok ") == 2


@pytest.mark.parametrize("mode", [MODE_STRICT, MODE_FAST])
@pytest.mark.parametrize("with_h,heatbath", [(False, False), (True, False), (True, True)])
def test_tempering_between_unequal_hamiltonians_matches_reference(with_h, heatbath, mode):
    """GraphWeights::relative_weight / ham_eq (tempering_traits.rs:122-154) in swap_on_chunks
    (tempering_container.rs:274-302): every ladder position has its own couplings."""
    from isingmontecarlo_b200.tempering import TemperingContainer

    edges = lattices.two_d_periodic_mixed(4)
    n_chains, n_betas = 2, 5
    betas = np.linspace(0.6, 1.4, n_betas)
    J0 = np.array([j for _, j in edges])
    # positions 0 and 1 share a Hamiltonian (ham_eq true -> factor 1); 1 -> 2 changes Gamma only; 2 -> 3 rescales
    # the couplings; 3 -> 4 changes the longitudinal field only (which ham_eq ignores, qmc_ising.rs:899-903)
    hl = (lambda x: x) if with_h else (lambda x: 0.0)
    hams = [(J0, 1.0, hl(0.5)), (J0, 1.0, hl(0.5)), (J0, 1.2, hl(0.5)), (1.1 * J0, 1.2, hl(0.45)), (1.1 * J0, 1.2, hl(0.6))]
    S = n_chains * n_betas
    keys = 0x55E10000 + np.arange(S, dtype=np.uint64)
    pt_key = 0xFEED
    tc = TemperingContainer(edges, hams[0][1], hams[0][2], 16, betas, n_chains=n_chains, rng_keys=keys, pt_key=pt_key, mode=mode,
                            slot_hamiltonians=hams)
    tc.graph.set_enable_heatbath(heatbath)

    def ref_graph(c, k):
        ek = [(e, float(j)) for (e, _), j in zip(edges, hams[k][0])]
        g = po.SseOracle(ek, hams[k][1], hams[k][2], 16, key=int(keys[c * n_betas + k]))
        g.set_enable_heatbath(heatbath)
        return g

    slots = [[ref_graph(c, k) for k in range(n_betas)] for c in range(n_chains)]
    assert np.array_equal(tc.graph.get_offsets(), [slots[s // n_betas][s % n_betas].offset for s in range(S)])
    cursors = [0] * n_chains
    swaps_ref = 0
    for step in range(14):
        e_gpu = tc.timesteps(3)
        slot_before = tc.slots()
        for c in range(n_chains):
            for k in range(n_betas):
                e_ref = slots[c][k].timesteps(3, float(betas[k]), mode)
                s_local = int(np.where(slot_before == c * n_betas + k)[0][0])
                assert e_gpu[s_local] == e_ref  # offsets follow the slot's Hamiltonian
        tc.tempering_step()
        for c in range(n_chains):
            s, cursors[c] = po.pt_step(slots[c], betas, pt_key + c, cursors[c])
            swaps_ref += s
        g = tc.graph
        n, cut, cur, st, hidx = g.get_n(), g.get_cutoff(), g.rng_cursors(), g.state_ref(), g.hamiltonian_index()
        for s_local, slot in enumerate(tc.slots()):
            ref = slots[slot // n_betas][slot % n_betas]
            assert hidx[s_local] == slot % n_betas
            assert int(n[s_local]) == ref.n and int(cut[s_local]) == ref.cutoff and int(cur[s_local]) == ref.cursor
            assert np.array_equal(st[s_local], ref.state())
            assert np.array_equal(g.dump_ops(s_local), ref.dump_ops())
        assert tc.get_total_swaps() == swaps_ref
    assert swaps_ref > 0 and tc.verify() and all(r.error == 0 for row in slots for r in row)


def test_can_swap_graphs_is_enforced():
    from isingmontecarlo_b200 import QmcbError
    from isingmontecarlo_b200.sse import QmcIsingGraph

    edges = lattices.two_d_periodic_mixed(4)
    J0 = np.array([j for _, j in edges])
    g = QmcIsingGraph(edges, 1.0, 0.5, 16, [1, 2], 1.0, mode=MODE_FAST)
    Jbad = J0.copy()
    Jbad[3] = -Jbad[3]
    with pytest.raises(QmcbError, match="same sign"):  # qmc_ising.rs:572-577
        g.set_hamiltonians([J0, Jbad], [1.0, 1.0], [0.5, 0.5], [0, 1])
    with pytest.raises(QmcbError, match="same sign"):  # qmc_ising.rs:581-586
        g.set_hamiltonians([J0, J0], [1.0, 1.0], [0.5, -0.5], [0, 1])
    a = po.SseOracle(edges, 1.0, 0.5, 16)
    b = po.SseOracle([(e, -j) for e, j in edges], 1.0, 0.5, 16)
    c = po.SseOracle(edges, 2.0, -0.5, 16)
    L = po.lib()
    assert L.orc_sse_can_swap(a._h, a._h) == 0 and L.orc_sse_can_swap(a._h, b._h) == 2 and L.orc_sse_can_swap(a._h, c._h) == 3
    g.set_hamiltonians([J0, 2 * J0], [1.0, 3.0], [0.5, 0.25], [1, 0])
    assert list(g.hamiltonian_index()) == [1, 0]
    refs = [po.SseOracle([(e, 2 * j) for e, j in edges], 3.0, 0.25, 16, key=1), po.SseOracle(edges, 1.0, 0.5, 16, key=2)]
    e = g.timesteps(20, 1.0)
    for r, ref in enumerate(refs):
        assert e[r] == ref.timesteps(20, 1.0, MODE_FAST)
        assert np.array_equal(g.dump_ops(r), ref.dump_ops()) and np.array_equal(g.state_ref()[r], ref.state())


@pytest.mark.parametrize("minblocks", [7, 4])
@pytest.mark.parametrize("heatbath", [False, True])
def test_per_replica_hamiltonians_in_every_kernel_build(minblocks, heatbath):
    """set_hamiltonians + both diagonal rules through the 72-register and the 120-register builds (the launcher would
    pick the latter for so few replicas)."""
    from isingmontecarlo_b200.sse import QmcIsingGraph

    edges = lattices.two_d_periodic_mixed(4)
    J0 = np.array([j for _, j in edges])
    g = QmcIsingGraph(edges, 1.0, 0.5, 16, [11, 12, 13], 1.0, mode=MODE_FAST)
    try:
        g.set_option("minblocks", minblocks)
        g.set_hamiltonians([J0, 2 * J0, 0.5 * J0], [1.0, 3.0, 0.7], [0.5, 0.25, 0.9], [1, 0, 2])
        g.set_enable_heatbath(heatbath)
        rows = [(2 * J0, 3.0, 0.25), (J0, 1.0, 0.5), (0.5 * J0, 0.7, 0.9)]
        refs = []
        for k, (jr, gam, hl) in zip([11, 12, 13], rows):
            ref = po.SseOracle([(e, float(j)) for (e, _), j in zip(edges, jr)], gam, hl, 16, key=k)
            ref.set_enable_heatbath(heatbath)
            refs.append(ref)
        e = g.timesteps(25, 1.0)
        for r, ref in enumerate(refs):
            assert e[r] == ref.timesteps(25, 1.0, MODE_FAST)
            assert np.array_equal(g.dump_ops(r), ref.dump_ops()) and np.array_equal(g.state_ref()[r], ref.state())
    finally:
        g.set_option("minblocks", 0)


def test_square_fused_launches_equal_per_pass_launches():
    """cmcb_set_option("fused"): both colours of many sweeps in one cooperative launch (per-replica barrier between
    colour passes) against one launch per colour pass, on a lattice whose row bands span several blocks; and more
    replicas than one launch can hold co-resident (the launcher splits them)."""
    from isingmontecarlo_b200.classical import GraphState

    for L, R, sweeps in ((256, 5, 9), (64, 310, 4)):
        edges = lattices.square_periodic(L, -1.0)
        keys = 0xB2100000 + np.arange(R, dtype=np.uint64)
        betas = np.linspace(0.3, 0.6, R)
        a = GraphState(edges, np.zeros(L * L), keys, betas)
        b = GraphState(edges, np.zeros(L * L), keys, betas)
        b.set_option("fused", 0)
        for chunk in (1, sweeps):
            a.sweeps(chunk), b.sweeps(chunk)
            assert np.array_equal(a.state_ref(), b.state_ref()), (L, chunk)
        assert np.array_equal(a.get_energy(), b.get_energy())
        assert a.launch_count() < b.launch_count()
        if L == 256:  # and against the oracle
            col, _ = a.colours()
            ref = po.ClassicalOracle(edges, np.zeros(L * L), key=int(keys[2]))
            ref.checkerboard_sweeps(float(betas[2]), col, 1 + sweeps)
            assert np.array_equal(a.state_ref()[2], ref.state())
        a.close(), b.close()


@pytest.mark.gpu
@pytest.mark.parametrize("L,R,sweeps", [(128, 3, 6), (192, 2, 5), (320, 2, 3), (1024, 1, 2)])
def test_square_row_widths_against_oracle(L, R, sweeps):
    """The fused kernel addresses a word's neighbours at fixed distances when a round of a block covers an even number of
    whole rows (L/64 a power of two up to 64) and generically otherwise (L = 192, 320): both against the oracle, and the
    deferred third Philox call (words with a tie after eight bit planes are parked and resolved 32 at a time)."""
    from isingmontecarlo_b200.classical import GraphState

    edges = lattices.square_periodic(L, -1.0)
    keys = 0xB2200000 + np.arange(R, dtype=np.uint64)
    betas = np.linspace(0.35, 0.5, R)
    g = GraphState(edges, np.full(L * L, 0.1 if L == 192 else 0.0), keys, betas)
    assert g.is_bitpacked_square()
    g.sweeps(1), g.sweeps(sweeps - 1)
    col, _ = g.colours()
    r = R - 1
    ref = po.ClassicalOracle(edges, np.full(L * L, 0.1 if L == 192 else 0.0), key=int(keys[r]))
    ref.checkerboard_sweeps(float(betas[r]), col, sweeps)
    assert np.array_equal(g.state_ref()[r], ref.state())
    g.close()
