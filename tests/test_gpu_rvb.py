"""RVB update on the device (sse_rvb.cu: qmcb_set_run_rvb, qmcb_single_rvb_sweep, qmcb_rvb_success_rate) against the
oracle's restatement of rvb.rs, bit-exact: spins, operator words, stream cursors, success counts -- in all three modes,
with and without a longitudinal field, through capacity growth, and as a stand-alone sweep."""
import numpy as np
import pytest

from isingmontecarlo_b200 import MODE_COUNTER, MODE_FAST, MODE_STRICT, QmcbError, lattices
from oracle import pyoracle as po
from tests.ed import tfim_thermal
from tests.test_gpu_sse_parity import assert_same, make_pair

pytestmark = pytest.mark.gpu

CASES = [
    # name, edges, gamma, h, cutoff0, beta, sweeps
    ("mixed3x3", lattices.two_d_periodic_mixed(3), 0.1, 0.0, 9, 1.0, 40),        # check_rvb_crash.rs:296-315
    ("mixed4x4", lattices.two_d_periodic_mixed(4), 0.1, 0.0, 16, 1.0, 30),       # :318-337
    ("two_unit_cell", lattices.two_unit_cell(), 1.0, 0.0, 8, 1.0, 40),           # :340-359
    ("two_unit_cell_h", lattices.two_unit_cell(), 1.0, 1.0, 8, 1.0, 40),         # longitudinal_crash.rs, rvb cases
    ("mixed4x4_h", lattices.two_d_periodic_mixed(4), 1.0, -0.4, 16, 2.0, 25),
    ("tri6_frustrated", lattices.triangular_periodic(6, 1.0), 0.6, 0.0, 36, 2.0, 12),
    ("tri6_frustrated_h", lattices.triangular_periodic(6, 1.0), 1.0, 0.2, 36, 2.0, 12),
]


@pytest.mark.parametrize("mode,impl", [(MODE_STRICT, 0), (MODE_STRICT, 1), (MODE_FAST, 0), (MODE_COUNTER, 0), (MODE_COUNTER, 1)])
@pytest.mark.parametrize("name,edges,gamma,h,cutoff,beta,sweeps", CASES)
def test_sweeps_with_rvb_steps_bit_exact(name, edges, gamma, h, cutoff, beta, sweeps, mode, impl):
    g, refs = make_pair(edges, gamma, h, cutoff, beta, mode, impl=impl)
    g.set_run_rvb(True)
    for ref in refs:
        ref.set_run_rvb(True)
    for chunk in (1, 1, 3, sweeps - 5):
        e_gpu = g.timesteps(chunk, beta)
        e_ref = [ref.timesteps(chunk, beta, mode) for ref in refs]
        assert_same(g, refs, f"{name} after +{chunk}")
        assert np.array_equal(e_gpu, np.array(e_ref)), name
    assert g.verify()
    rate = g.rvb_success_rate()
    assert np.array_equal(rate, np.array([ref.rvb_success_rate() for ref in refs]))
    assert (rate > 0).all() and (rate < 1).all()
    # off again: the plain sweep (qmc_ising.rs:705)
    g.set_run_rvb(False)
    for ref in refs:
        ref.set_run_rvb(False)
    g.timesteps(2, beta)
    for ref in refs:
        ref.timesteps(2, beta, mode)
    assert_same(g, refs, f"{name} rvb off")


def test_single_rvb_sweep_matches():
    edges = lattices.triangular_periodic(6, 1.0)
    g, refs = make_pair(edges, 0.8, 0.0, 36, 2.0, MODE_STRICT, R=6)
    g.timesteps(15, 2.0)
    for ref in refs:
        ref.timesteps(15, 2.0)
    for updates in (None, 1, 50):
        succ, att = g.single_rvb_sweep(updates)
        want = [ref.single_rvb_sweep(updates) for ref in refs]
        assert att == want[0][1]
        assert [int(s) for s in succ] == [w[0] for w in want]
        assert_same(g, refs, f"single_rvb_sweep({updates})")
    assert g.verify()


def test_rvb_through_capacity_growth():
    # cutoff 8 -> hundreds of slots: the workspace follows the re-layout of the operator strings
    edges = lattices.two_unit_cell()
    g, refs = make_pair(edges, 1.0, 0.0, 8, 6.0, MODE_FAST, R=3)
    g.set_run_rvb(True)
    for ref in refs:
        ref.set_run_rvb(True)
    g.timesteps(40, 6.0)
    for ref in refs:
        ref.timesteps(40, 6.0, MODE_FAST)
    assert_same(g, refs, "growth")
    assert g.verify()


def test_rvb_energy_matches_exact_diagonalisation_on_the_device():
    from isingmontecarlo_b200.sse import QmcIsingGraph

    edges = [((0, 1), 1.0), ((1, 2), 1.0), ((2, 0), 1.0), ((2, 3), 1.0), ((3, 4), 1.0), ((4, 2), 1.0)]
    gamma, beta, R = 0.3, 3.0, 256
    exact = tfim_thermal(edges, 5, gamma, 0.0, beta)
    g = QmcIsingGraph(edges, gamma, 0.0, 5, 0xEB0000 + np.arange(R, dtype=np.uint64), beta, mode=MODE_COUNTER)
    g.set_run_rvb(True)
    g.timesteps(300, beta)
    e = g.timesteps(3000, beta)
    assert g.verify()
    mean, err = e.mean(), e.std(ddof=1) / np.sqrt(R)
    assert abs(mean - exact["E"]) < 3.0 * err + 1e-9, (mean, err, exact["E"])
    assert (g.rvb_success_rate() > 0.2).all()


def test_generic_qmc_has_no_rvb():
    from isingmontecarlo_b200.sse import QmcIsingGraph

    g = QmcIsingGraph(lattices.small_qmc_ring(), 1.0, 0.0, 3, [1, 2], 1.0)
    q = g.into_qmc()
    with pytest.raises(QmcbError):
        q.set_run_rvb(True)  # Qmc::timestep has no RVB step (qmc_runner.rs:363-377)


@pytest.mark.parametrize("mode", [MODE_STRICT, MODE_FAST])
def test_rvb_with_heatbath_diagonal_update(mode):
    # longitudinal_crash.rs:39-178 runs its lattices with and without rvb, with and without the heat-bath rule
    g, refs = make_pair(lattices.two_d_periodic_mixed(4), 1.0, 1.0, 16, 1.0, mode)
    g.set_run_rvb(True)
    g.set_enable_heatbath(True)
    for ref in refs:
        ref.set_run_rvb(True)
        ref.set_enable_heatbath(True)
    g.timesteps(30, 1.0)
    for ref in refs:
        ref.timesteps(30, 1.0, mode)
    assert_same(g, refs, "rvb + heatbath")
    assert g.verify()


def test_tempering_ladder_with_rvb_steps():
    # tempering_container.rs:121-149 steps graphs that run their own timestep: with set_run_rvb the RVB updates are part of it
    from isingmontecarlo_b200.tempering import TemperingContainer

    edges = lattices.two_unit_cell()
    n_chains, n_betas, mode = 2, 4, MODE_COUNTER
    betas = np.linspace(0.5, 2.0, n_betas)
    keys = 0x55E00000 + np.arange(n_chains * n_betas, dtype=np.uint64)
    pt_key = 0xABCDEF
    tc = TemperingContainer(edges, 1.0, 0.0, 8, betas, n_chains=n_chains, rng_keys=keys, pt_key=pt_key, mode=mode)
    tc.graph.set_run_rvb(True)
    slots = [[po.SseOracle(edges, 1.0, 0.0, 8, key=int(keys[c * n_betas + k])) for k in range(n_betas)] for c in range(n_chains)]
    for ladder in slots:
        for ref in ladder:
            ref.set_run_rvb(True)
    cursors = [0] * n_chains
    swaps_ref = 0
    for _ in range(10):
        tc.timesteps(2)
        for c in range(n_chains):
            for k in range(n_betas):
                slots[c][k].timesteps(2, float(betas[k]), mode)
        tc.tempering_step()
        for c in range(n_chains):
            s, cursors[c] = po.pt_step(slots[c], betas, pt_key + c, cursors[c])
            swaps_ref += s
        g = tc.graph
        n, cut, cur, st = g.get_n(), g.get_cutoff(), g.rng_cursors(), g.state_ref()
        for s_local, slot in enumerate(tc.slots()):
            ref = slots[slot // n_betas][slot % n_betas]
            assert ref.error == 0
            assert int(n[s_local]) == ref.n and int(cut[s_local]) == ref.cutoff and int(cur[s_local]) == ref.cursor
            assert np.array_equal(st[s_local], ref.state())
            assert np.array_equal(g.dump_ops(s_local), ref.dump_ops())
        assert tc.get_total_swaps() == swaps_ref
    assert swaps_ref > 0 and tc.verify()


def tri_periodic(lx, ly, j=1.0):
    f = lambda i, jj: jj * lx + i  # noqa: E731
    e = []
    for i in range(lx):
        for jj in range(ly):
            e += [((f(i, jj), f((i + 1) % lx, jj)), j), ((f(i, jj), f(i, (jj + 1) % ly)), j), ((f(i, jj), f((i + 1) % lx, (jj + 1) % ly)), j)]
    return e


def test_rvb_energy_on_a_frustrated_torus_matches_exact_diagonalisation():
    # 3 x 4 periodic triangular antiferromagnet (12 sites, 36 bonds): the lattice family RVB was written for.  Without RVB
    # steps the same chain length is 2.7 sigma off with five times the error bar (the oracle, 64 chains); with them the
    # device reproduces the exact energy at 3e-3 precision
    from isingmontecarlo_b200.sse import QmcIsingGraph

    edges = tri_periodic(3, 4)
    gamma, beta, R = 0.7, 3.0, 512
    exact = tfim_thermal(edges, 12, gamma, 0.0, beta)
    g = QmcIsingGraph(edges, gamma, 0.0, 12, 0x50A40000 + np.arange(R, dtype=np.uint64), beta, mode=MODE_COUNTER)
    g.set_run_rvb(True)
    g.timesteps(500, beta)
    e = g.timesteps(3000, beta)
    assert g.verify()
    assert (g.rvb_success_rate() > 0.1).all()
    mean, err = e.mean(), e.std(ddof=1) / np.sqrt(R)
    assert abs(mean - exact["E"]) < 3.0 * err + 1e-9, (mean, err, exact["E"])


@pytest.mark.parametrize("mode", [MODE_STRICT, MODE_COUNTER])
def test_rvb_with_per_replica_hamiltonians(mode):
    # one batch, three Hamiltonians (couplings of the same signs, different magnitudes, own fields): the RVB weights
    # (bond_mag, the diagonal edge weights) are those of the replica's own row, as for one reference graph per Hamiltonian
    from isingmontecarlo_b200.sse import QmcIsingGraph

    edges = lattices.two_unit_cell()
    scales, gammas, fields = [1.0, 1.5, 0.6], [1.0, 0.8, 1.3], [0.0, 0.0, 0.0]
    keys = [0x4A770000 + r for r in range(3)]
    g = QmcIsingGraph(edges, 1.0, 0.0, 8, keys, 1.5, mode=mode)
    g.set_hamiltonians([[j * s for _, j in edges] for s in scales], gammas, fields, [0, 1, 2])
    g.set_run_rvb(True)
    refs = [po.SseOracle([(e, j * s) for e, j in edges], gam, h, 8, key=k) for s, gam, h, k in zip(scales, gammas, fields, keys)]
    for ref in refs:
        ref.set_run_rvb(True)
    g.timesteps(30, 1.5)
    for ref in refs:
        ref.timesteps(30, 1.5, mode)
    assert_same(g, refs, "per-replica Hamiltonians")
    assert g.verify()


def test_rvb_through_enqueue_sweeps():
    # the asynchronous entry point (qmcb_enqueue_sweeps + qmcb_synchronize) takes the same three launches per sweep
    g, refs = make_pair(lattices.two_d_periodic_mixed(3), 0.1, 0.0, 9, 1.0, MODE_COUNTER, R=4)
    g.set_run_rvb(True)
    for ref in refs:
        ref.set_run_rvb(True)
    g.enqueue_sweeps(7)
    g.enqueue_sweeps(5)
    g.synchronize()
    for ref in refs:
        ref.timesteps(12, 1.0, MODE_COUNTER)
    assert_same(g, refs, "enqueue_sweeps")
    assert g.verify()


@pytest.mark.parametrize("h", [0.0, 0.2])
def test_rvb_parity_on_long_strings_from_a_thermalised_batch(h):
    # long world lines (triangular L = 12, beta = 6: ~5000 ops, ~70 per variable), clusters that wrap through p = 0, dozens of
    # rotations per sweep: four replicas of a thermalised 128-replica batch are copied into oracle graphs and both sides run
    # sweeps with RVB steps
    from isingmontecarlo_b200.sse import QmcIsingGraph
    from tests.test_gpu_full_size import same, to_oracle

    edges = lattices.triangular_periodic(12, 1.0)
    gamma, beta, R = 1.0, 6.0, 128
    g = QmcIsingGraph(edges, gamma, h, 144, 0x7B1C0000 + np.arange(R, dtype=np.uint64), beta, mode=MODE_COUNTER)
    g.timesteps(40, beta)
    g.set_run_rvb(True)
    g.timesteps(3, beta)
    picks = (0, 37, 90, R - 1)
    refs = {r: to_oracle(g, r, edges, gamma, h) for r in picks}
    for ref in refs.values():
        ref.set_run_rvb(True)
    e = g.timesteps(4, beta)
    for r, ref in refs.items():
        e_ref = ref.timesteps(4, beta, MODE_COUNTER)
        assert ref.error == 0 and same(g, r, ref), r
        assert e[r] == e_ref
    assert g.verify()
    assert (g.rvb_success_rate() > 0.05).all()
