"""GPU parity tests proper: the CUDA path (through the C ABI, via the host mirror) against the CPU
oracle on the same seeded inputs.  Integer/bit work => bit-exact: operator words, spin states,
n, cutoff and stream cursor must be identical after every compared sweep; energies are the same
f64 expression of the same integers, so they must match exactly too."""
import numpy as np
import pytest

from isingmontecarlo_b200 import MODE_FAST, MODE_STRICT, QmcbError, lattices
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu

CASES = [
    # name, edges, gamma, h, cutoff0, beta, sweeps
    ("small_qmc", lattices.small_qmc_ring(), 1.0, 0.0, 3, 1.0, 40),
    ("pair_h", [((0, 1), 1.0)], 1.0, 1.0, 2, 1.0, 40),
    ("ring6_ferro", lattices.one_d_periodic(6, -1.0), 0.7, 0.0, 6, 3.0, 30),
    ("mixed4x4_h", lattices.two_d_periodic_mixed(4), 1.0, 1.0, 16, 1.0, 30),
    ("two_unit_cell_h", lattices.two_unit_cell(), 1.0, -0.4, 8, 2.0, 30),
    ("square8_crit", lattices.square_periodic(8, -1.0), 3.04, 0.0, 64, 4.0, 20),
    ("tri6_frustrated_h", lattices.triangular_periodic(6, 1.0), 1.0, 0.2, 36, 2.0, 15),
]


def make_pair(edges, gamma, h, cutoff, beta, mode, R=5, key0=0x55E00000, impl=0):
    from isingmontecarlo_b200.sse import QmcIsingGraph

    keys = [key0 + r for r in range(R)]
    g = QmcIsingGraph(edges, gamma, h, cutoff, keys, beta, mode=mode)
    g.set_option("impl", impl)
    refs = [po.SseOracle(edges, gamma, h, cutoff, key=k) for k in keys]
    return g, refs


def assert_same(g, refs, tag=""):
    n, cut, cur, st = g.get_n(), g.get_cutoff(), g.rng_cursors(), g.state_ref()
    for r, ref in enumerate(refs):
        assert ref.error == 0
        assert int(n[r]) == ref.n, (tag, r, "n")
        assert int(cut[r]) == ref.cutoff, (tag, r, "cutoff")
        assert int(cur[r]) == ref.cursor, (tag, r, "cursor")
        assert np.array_equal(st[r], ref.state()), (tag, r, "state")
        assert np.array_equal(g.dump_ops(r), ref.dump_ops()), (tag, r, "ops")


@pytest.mark.parametrize("mode", [MODE_STRICT, MODE_FAST])
@pytest.mark.parametrize("name,edges,gamma,h,cutoff,beta,sweeps", CASES)
def test_sweeps_bit_exact(name, edges, gamma, h, cutoff, beta, sweeps, mode):
    g, refs = make_pair(edges, gamma, h, cutoff, beta, mode)
    assert_same(g, refs, "init")  # stream-drawn initial states (classical/graph.rs:451-453)
    for chunk in (1, 1, 3, sweeps - 5):
        e_gpu = g.timesteps(chunk, beta)
        e_ref = [ref.timesteps(chunk, beta, mode) for ref in refs]
        assert_same(g, refs, f"{name} after +{chunk}")
        assert np.array_equal(e_gpu, np.array(e_ref)), name
    assert g.verify()
    assert all(ref.verify() for ref in refs)


@pytest.mark.parametrize("mode", [MODE_STRICT, MODE_FAST])
def test_serial_and_parallel_impl_agree(mode):
    # impl=1 forces the serial-order kernels; impl=0 picks the warp-parallel ones where available
    edges = lattices.square_periodic(8, -1.0)
    a, refs = make_pair(edges, 3.04, 0.0, 64, 4.0, mode, R=4, impl=0)
    b, _ = make_pair(edges, 3.04, 0.0, 64, 4.0, mode, R=4, impl=1)
    a.timesteps(12, 4.0), b.timesteps(12, 4.0)
    [ref.timesteps(12, 4.0, mode) for ref in refs]
    assert_same(a, refs, "impl0"), assert_same(b, refs, "impl1")


@pytest.mark.parametrize("mode", [MODE_STRICT, MODE_FAST])
def test_single_steps_and_cluster_counts(mode):
    edges = lattices.two_d_periodic_mixed(4)
    g, refs = make_pair(edges, 1.0, 0.5, 16, 1.5, mode, R=4)
    for _ in range(6):
        g.single_diagonal_step(1.5)
        [ref.single_diagonal_step(1.5) for ref in refs]
        assert_same(g, refs, "diag")
        ncl = g.single_cluster_step()
        ncl_ref = [ref.single_cluster_step(mode) for ref in refs]
        assert [int(x) for x in ncl] == ncl_ref
        assert_same(g, refs, "cluster")


def test_strict_cluster_numbering_matches_reference_order():
    # cluster ids per (slot, side) in discovery order (cluster.rs:57-97) for a thermalised string
    edges = lattices.two_d_periodic_mixed(4)
    g, refs = make_pair(edges, 1.0, 0.0, 16, 2.0, MODE_STRICT, R=3)
    g.timesteps(10, 2.0)
    [ref.timesteps(10, 2.0) for ref in refs]
    g.single_diagonal_step(2.0)
    [ref.single_diagonal_step(2.0) for ref in refs]
    g.single_cluster_step()
    for r, ref in enumerate(refs):
        ref.single_cluster_step()
        m = ref.cutoff
        bi, bo = g.boundaries(r, m)
        ri, ro = ref.boundaries(m)
        has_op = ri >= 0
        assert np.array_equal(bi[has_op].astype(np.int64), ri[has_op])
        assert np.array_equal(bo[has_op].astype(np.int64), ro[has_op])


def test_hand_built_strings_from_cluster_test_rs():
    # tests/cluster_test.rs:7-75 strings (B1-B3 of SURVEY Appendix B) through load_ops
    from isingmontecarlo_b200.sse import QmcIsingGraph

    def word(bond, i, o):
        return bond | (i << 24) | (o << 26)

    edges = [((0, 1), -1.0)]  # bond 0 two-site, bonds 1,2 transverse on var 0,1
    for ops, expect_ncl in (([word(1, 0, 0)], 1), ([word(1, 0, 0), word(1, 0, 0)], 2),
                            ([word(1, 0, 0), word(1, 0, 0), word(2, 0, 0), word(2, 0, 0)], 4),
                            ([word(1, 0, 0), word(0, 0, 0), word(2, 0, 0)], 1), ([word(0, 0, 0), word(0, 0, 0)], 1)):
        g = QmcIsingGraph(edges, 1.0, 0.0, len(ops), [77], 1.0, state=[0, 0])
        ref = po.SseOracle(edges, 1.0, 0.0, len(ops), key=77, state=[0, 0])
        g.load_ops(0, ops, [0, 0]), ref.load_ops(ops, [0, 0])
        assert int(g.get_n()[0]) == ref.n
        assert int(g.single_cluster_step()[0]) == ref.single_cluster_step() == expect_ncl
        assert np.array_equal(g.dump_ops(0), ref.dump_ops()) and np.array_equal(g.state_ref()[0], ref.state())
        assert g.verify()


def test_timesteps_sample_matches():
    edges = lattices.small_qmc_ring()
    g, refs = make_pair(edges, 1.0, 0.3, 3, 1.0, MODE_STRICT, R=3)
    s_gpu, e_gpu = g.timesteps_sample(21, 1.0, 4)
    for r, ref in enumerate(refs):
        s_ref, e_ref = ref.timesteps_sample(21, 1.0, 4)
        assert s_gpu.shape[1] == 5 and np.array_equal(s_gpu[r], s_ref)
        assert e_gpu[r] == e_ref


def test_capacity_error_and_growth():
    from isingmontecarlo_b200.sse import QmcIsingGraph

    edges = lattices.square_periodic(4, -1.0)
    g = QmcIsingGraph(edges, 2.0, 0.0, 16, [1, 2], 4.0, capacity=32)
    with pytest.raises(QmcbError) as ei:
        g.timesteps(20, 4.0)
    assert ei.value.code == -2  # QMCB_ERR_CAPACITY: status code instead of the reference's abort
    g2 = QmcIsingGraph(edges, 2.0, 0.0, 16, [1, 2], 4.0, capacity=32)
    g2.set_option("auto_capacity", 1)
    refs = [po.SseOracle(edges, 2.0, 0.0, 16, key=k) for k in (1, 2)]
    g2.timesteps(20, 4.0)
    [ref.timesteps(20, 4.0) for ref in refs]
    assert g2.get_capacity() > 32
    assert_same(g2, refs, "grown")


def test_bond_counts_and_dump_load_round_trip():
    edges = lattices.two_d_periodic_mixed(4)
    g, refs = make_pair(edges, 1.0, 0.3, 16, 2.0, MODE_STRICT, R=2)
    g.timesteps(15, 2.0)
    [ref.timesteps(15, 2.0) for ref in refs]
    counts = g.get_bond_counts(1)
    assert [int(c) for c in counts] == [refs[1].bond_count(b) for b in range(len(counts))]
    words, st = g.dump_ops(0), g.state_ref()[0]
    g.load_ops(1, words, st)  # replica 1 becomes a copy of replica 0's configuration
    assert g.verify(1) and int(g.get_n()[1]) == refs[0].n
    # a corrupted string must fail verify (op_container.rs:137-159)
    bad = words.copy()
    k = int(np.nonzero(bad != 0xFFFFFFFF)[0][0])
    bad[k] ^= 1 << 24
    g.load_ops(1, bad, st)
    assert not g.verify(1)


def test_cpp_host_mirror_runs_the_reference_integration_tests():
    # include/qmcb.hpp mirrors QmcIsingGraph / QmcStepper / GraphState in C++; tests/cpp/test_mirror.cpp is
    # tests/longitudinal_crash.rs + examples/small_qmc.rs + convert_test.rs re-read against it
    import os
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.run(["make", "-C", os.path.join(root, "isingmontecarlo_b200", "csrc"), "-s", "mirror"], check=True)
    res = subprocess.run([os.path.join(root, "isingmontecarlo_b200", "_build", "test_mirror")], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "cpp mirror ok" in res.stdout


def test_imaginary_time_fold_matches():
    # qmc_ising.rs:815-821 / fast_ops.rs:1296-1315: magnetisation fold on the device vs the literal fold
    edges = lattices.two_d_periodic_mixed(4)
    g, refs = make_pair(edges, 1.0, 0.3, 16, 2.0, MODE_FAST, R=4)
    g.timesteps(30, 2.0)
    for ref in refs:
        ref.timesteps(30, 2.0, MODE_FAST)
    assert_same(g, refs, "before fold")
    m1, m2, mabs = g.imaginary_time_magnetization()
    nv = 16.0
    for r, ref in enumerate(refs):
        slots, s1, s2, s3 = ref.itime_magnetization()
        assert slots == ref.cutoff
        assert m1[r] == s1 / slots / nv and m2[r] == s2 / slots / (nv * nv) and mabs[r] == s3 / slots / nv
        for p in (0, 1, ref.cutoff // 2, ref.cutoff - 1, ref.cutoff):
            st = np.zeros(16, dtype=np.uint8)
            from isingmontecarlo_b200._lib import check, ptr
            import ctypes as C
            check(g._L.qmcb_itime_state(g._h, r, p, ptr(st, C.c_uint8)))
            assert np.array_equal(st, ref.itime_state(p)), (r, p)
    # closure form on the host: count slices with positive magnetisation, against the oracle states
    pos = g.imaginary_time_fold(0, lambda acc, s: acc + (2 * int(s.sum()) > 16), 0)
    assert pos == sum(2 * int(refs[0].itime_state(p).sum()) > 16 for p in range(refs[0].cutoff))
    # the fold returns to the p = 0 state after the last slot (periodic in imaginary time)
    assert np.array_equal(refs[0].itime_state(refs[0].cutoff), refs[0].state())


def fft_autocorrelation(samples):
    """literal numpy restatement of fft_autocorrelation (autocorrelations.rs:99-133); samples [T][n] floats"""
    tmax, n = samples.shape
    x = samples - samples.mean(axis=0)
    x = x / np.sqrt((x * x).sum(axis=0))
    f = np.fft.fft(x.astype(np.complex128), axis=0)
    r = np.fft.ifft(np.abs(f) ** 2, axis=0).real * tmax  # rustfft's inverse is unnormalised
    return r.sum(axis=1) / (n * tmax)


def test_product_and_bond_autocorrelations_match_fft_restatement():
    """calculate_spin_product_autocorrelation / calculate_bond_autocorrelation (autocorrelations.rs:53-97) against the
    literal FFT restatement fed with the reference's own sample mappers (value_for_bond qmc_ising.rs:988-997)."""
    from isingmontecarlo_b200.sse import QmcIsingGraph

    edges = lattices.two_d_periodic_mixed(4)
    g = QmcIsingGraph(edges, 2.0, 0.0, 16, [5, 6, 7], 1.0, mode=MODE_FAST)
    g.timesteps(50, 1.0)
    prods = [[0, 1], [2, 5, 9], [3], [4, 8, 12, 15]]
    ac, samples = g.calculate_spin_product_autocorrelation(96, 1.0, prods, 2, return_samples=True)
    for r in range(3):
        pm = 2.0 * samples[r].astype(np.float64) - 1.0
        want = fft_autocorrelation(np.stack([pm[:, p].prod(axis=1) for p in prods], axis=1))
        assert np.allclose(ac[r], want, rtol=0, atol=1e-10)
    ac, samples = g.calculate_bond_autocorrelation(96, 1.0, 2, return_samples=True)
    for r in range(3):
        s = samples[r].astype(bool)
        vals = []
        for (a, b), j in edges:
            even = ((s[:, a].astype(int) + s[:, b].astype(int)) % 2) == 0
            vals.append(np.where(even if j < 0.0 else ~even, 1.0, -1.0))
        want = fft_autocorrelation(np.stack(vals, axis=1))
        assert np.allclose(ac[r], want, rtol=0, atol=1e-10, equal_nan=True)


@pytest.mark.parametrize("T,freq", [(256, 1), (100, 3), (37, 2)])
def test_variable_autocorrelation_matches_fft_restatement(T, freq):
    from isingmontecarlo_b200.sse import QmcIsingGraph

    edges = lattices.square_periodic(4, -1.0)
    g = QmcIsingGraph(edges, 2.0, 0.0, 16, [5, 6, 7], 1.0, mode=MODE_FAST)
    g.timesteps(50, 1.0)
    ac, samples = g.calculate_variable_autocorrelation(T * freq, 1.0, freq, return_samples=True)
    assert ac.shape == (3, T) and samples.shape == (3, T, 16)
    for r in range(3):
        want = fft_autocorrelation(2.0 * samples[r].astype(np.float64) - 1.0)
        assert np.allclose(ac[r], want, rtol=0, atol=1e-10), np.abs(ac[r] - want).max()  # tolerance: FFT rounding in the restatement
        assert abs(ac[r][0] - 1.0) < 1e-12


def test_stepper_iteration_helpers():
    # QmcStepper::timesteps_measure / timesteps_sample_iter / timesteps_sample_iter_zip (qmc_stepper.rs:43-103)
    edges = lattices.small_qmc_ring()
    g, refs = make_pair(edges, 1.0, 0.0, 3, 1.0, MODE_FAST, R=3)
    acc, e = g.timesteps_measure(12, 1.0, 0, lambda a, st: a + int(st.sum()), 3)
    seen = []
    e2 = g.timesteps_sample_iter(6, 1.0, 2, lambda st: seen.append(st.copy()))
    tagged = []
    g.timesteps_sample_iter_zip(6, 1.0, 1, ["a", "b"], lambda tag, st: tagged.append(tag))
    want = 0
    for ref in refs:
        s, _ = ref.timesteps_sample(12, 1.0, 3, MODE_FAST)
        want += int(s.sum())
    assert acc == want and len(seen) == 3 and seen[0].shape == (3, 4) and tagged == ["a", "b"]
    for r, ref in enumerate(refs):
        s, e_ref = ref.timesteps_sample(6, 1.0, 2, MODE_FAST)
        assert e2[r] == e_ref and all(np.array_equal(seen[k][r], s[k]) for k in range(3))


@pytest.mark.parametrize("mode", [MODE_STRICT, MODE_FAST, 2])
def test_convert_and_run(mode):
    """tests/convert_test.rs:9-31: `ising.clone().into_qmc()` and `ising` stepped 10 times from the same rng agree.  The
    two run DIFFERENT code paths here as they do in the reference: `ising` takes its weights from (J, Gamma) arithmetic
    (qmcb_create), `qmc` from the interaction tables (qmcb_create_qmc)."""
    from isingmontecarlo_b200 import QmcbError
    from isingmontecarlo_b200.sse import Qmc, QmcIsingGraph

    edges = lattices.one_d_periodic(3, 1.0)
    ising = QmcIsingGraph.new_with_rng(edges, 1.0, 0.0, 3, [1234, 1235], state=[1, 1, 1], betas=1.0, mode=mode)
    qmc = ising.clone().into_qmc()
    assert isinstance(qmc, Qmc) and qmc._h.value != ising._h.value
    for _ in range(10):
        ising.timestep(1.0)
        qmc.timestep(1.0)
        for r in range(2):
            assert np.array_equal(ising.dump_ops(r), qmc.dump_ops(r))
    assert np.array_equal(ising.state_ref(), qmc.state_ref())
    assert np.array_equal(ising.get_n(), qmc.get_n()) and np.array_equal(ising.rng_cursors(), qmc.rng_cursors()) and qmc.verify()
    # the Qmc energy offset carries the bond terms only (qmc_runner.rs:124-133)
    e_i, e_q = ising.timesteps(50, 1.0), qmc.timesteps(50, 1.0)
    assert np.allclose(e_i - e_q, 3 * 1.0) and qmc.get_offset() == 3.0
    assert len(qmc.get_bonds()) == 6 and qmc.should_do_cluster_update()
    with pytest.raises(QmcbError, match="fixed"):
        qmc.make_interaction([1.0, 1.0, 1.0, 1.0], [0])
    with pytest.raises(ValueError, match="negative weights"):  # qmc_ising.rs:966-969: the reference's unwrap() panics
        QmcIsingGraph.new_with_rng(edges, 1.0, 0.5, 3, [1], betas=1.0).into_qmc()
    # a bigger lattice, every kernel path the launcher can pick for the table-driven weights
    edges = lattices.two_d_periodic_mixed(4)
    ising = QmcIsingGraph.new_with_rng(edges, 0.8, 0.0, 16, [77, 78, 79], betas=1.5, mode=mode)
    ising.timesteps(20, 1.5)
    for impl, minblocks in ((0, 7), (0, 4), (1, 0)):
        qmc = ising.clone().into_qmc()
        qmc.set_option("impl", impl), qmc.set_option("minblocks", minblocks)
        a = ising.clone()
        a.timesteps(15, 1.5), qmc.timesteps(15, 1.5)
        for r in range(3):
            assert np.array_equal(a.dump_ops(r), qmc.dump_ops(r)), (impl, minblocks)
        assert np.array_equal(a.state_ref(), qmc.state_ref())


@pytest.mark.parametrize("mode", [MODE_STRICT, MODE_FAST, 2])
def test_generic_qmc_interactions_bit_exact(mode):
    """Qmc with interactions that are NOT a transverse-field Ising model (unequal two-variable weights, a full matrix with
    its offset, per-variable transverse weights), both diagonal rules, against the oracle's restatement of
    qmc_runner.rs; and the shapes the engine does not take."""
    from isingmontecarlo_b200 import QmcbError
    from isingmontecarlo_b200.sse import Qmc

    def build(target):
        target.make_diagonal_interaction_and_offset([0.3, 1.0, 1.0, 0.3], [0, 1])
        target.make_diagonal_interaction([2.0, 0.5, 0.5, 2.0], [1, 2])
        target.make_interaction_and_offset([0.6, 0, 0, 0, 0, 1.1, 0, 0, 0, 0, 1.1, 0, 0, 0, 0, 0.6], [2, 3])
        target.make_diagonal_interaction([0.6, 1.4, 1.4, 0.6], [3, 0])
        for v, gam in enumerate([0.5, 1.0, 1.5, 0.8]):
            target.make_interaction([gam] * 4, [v])

    keys = [0x9E0 + r for r in range(4)]
    for heatbath in ((False, True) if mode != 2 else (False,)):
        q = Qmc(4, keys, 1.2, mode=mode)
        build(q)
        q.set_enable_heatbath(heatbath)
        refs = []
        for k in keys:
            ref = po.QmcOracle(4, key=k)
            build(ref)
            ref.set_enable_heatbath(heatbath)
            refs.append(ref)
        assert q.get_offset() == refs[0].offset == -0.3 - 0.6
        for chunk in (1, 2, 10, 30):
            e = q.timesteps(chunk, 1.2)
            e_ref = [ref.timesteps(chunk, 1.2, mode) for ref in refs]
            assert_same(q, refs, f"generic +{chunk}")
            assert np.array_equal(e, np.array(e_ref))
        assert q.verify()
    bad = Qmc(3, [1], 1.0)
    bad.make_interaction([0.5] * 4, [0])  # one-variable interaction before the two-variable ones
    bad.make_diagonal_interaction([1.0, 0.2, 0.2, 1.0], [0, 1])
    with pytest.raises(QmcbError, match="interaction order|one constant"):
        bad.timestep(1.0)
    asym = Qmc(2, [1], 1.0)
    asym.make_diagonal_interaction([1.0, 0.2, 0.3, 1.0], [0, 1])
    asym.make_interaction([0.5] * 4, [0]), asym.make_interaction([0.5] * 4, [1])
    with pytest.raises(QmcbError, match="Ising symmetry"):
        asym.timestep(1.0)
    lp = Qmc(2, [1], 1.0)  # loop updates are offered (tests/test_gpu_loop_update.py): the flag is Qmc::set_do_loop_updates
    lp.set_do_loop_updates(True)
    assert lp.should_do_loop_update()


def test_small_accessors_and_new_from_graph(capsys):
    # get_transverse_field / get_longitudinal_field / clone_state / into_vec (qmc_ising.rs:497-529), print_debug (:489-494,
    # diagonal.rs:193-234), new_from_graph (:151-166), Qmc::diagonal_update / set_do_heatbath (qmc_runner.rs:159-202, 258-265)
    from isingmontecarlo_b200.classical import GraphState
    from isingmontecarlo_b200.sse import QmcIsingGraph

    edges = lattices.small_qmc_ring()
    c = GraphState(edges, np.zeros(4), [11, 12], 0.5)
    c.sweeps(3)
    g = QmcIsingGraph.new_from_graph(c, 1.0, 0.25, 4, betas=1.0, mode=MODE_FAST)
    assert np.array_equal(g.state_ref(), c.state_ref()) and list(g.rng_keys()) == [11, 12]
    assert g.get_transverse_field() == 1.0 and g.get_longitudinal_field() == 0.25 and g.get_edges() == list(edges)
    g.timesteps(10, 1.0)
    g.print_debug(1)
    lines = capsys.readouterr().out.splitlines()
    assert lines[0] == "====" and lines[1] == "".join(str(int(b)) for b in g.state_ref()[1])
    assert len(lines) == 2 + int(g.get_cutoff()[1]) and sum("\t" in ln and ":" in ln for ln in lines) == int(g.get_n()[1])
    st = g.clone_state()
    assert np.array_equal(g.into_vec(), st.astype(bool))
    with pytest.raises(QmcbError):  # biases must vanish (qmc_ising.rs:157)
        QmcIsingGraph.new_from_graph(GraphState(edges, np.ones(4), [1], 0.5), 1.0, 0.0, 4)
    q = QmcIsingGraph(edges, 1.0, 0.0, 4, [5, 6], 1.0, mode=MODE_FAST).into_qmc()
    q.set_do_heatbath(True)
    assert q.should_do_heatbath()
    q.diagonal_update(1.0)
    q.timesteps(5, 1.0)
    assert q.verify()
