"""CPU-side checks: the C-ABI library loads and exports every symbol include/qmcb.h declares (no
compute calls without a GPU), host logic of the mirrors, the oracle is not reachable from the
product package, and the multi-rank tempering plumbing over gloo (world_size 2)."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as ge

    ge.build()
    from isingmontecarlo_b200 import _lib

    return _lib


def test_library_exports_every_declared_symbol(built):
    header = open(os.path.join(ROOT, "include", "qmcb.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b((?:qmcb|cmcb)_[a-z0-9_]+)\s*\(", header))
    assert len(declared) > 50
    L = built.load()
    for name in declared:
        assert hasattr(L, name), name
    bound = set(built.SIGNATURES) | set(built.STRING_GETTERS)
    assert declared == bound, declared ^ bound


def test_no_gpu_fails_loudly_not_silently(built):
    import torch

    if torch.cuda.is_available():
        pytest.skip("has a GPU")
    from isingmontecarlo_b200 import QmcbError, lattices
    from isingmontecarlo_b200.sse import QmcIsingGraph

    with pytest.raises(QmcbError) as ei:
        QmcIsingGraph(lattices.small_qmc_ring(), 1.0, 0.0, 3, [1], 1.0)
    assert ei.value.code == -3  # QMCB_ERR_CUDA: there is no CPU fallback


def test_argument_validation_without_gpu(built):
    L = built.load()
    h = C.c_void_p()
    assert L.qmcb_create(None, 1, None, None, 1, 0, None, 0, C.byref(h)) == -1
    assert b"null" in L.qmcb_last_error()
    assert L.qmcb_destroy(None) == 0
    assert L.qmcb_get_n(None, None) == -1


def test_product_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "isingmontecarlo_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", "Makefile")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "pyoracle" not in src and "liboracle" not in src and "oracle/" not in src.replace("oracle.c cluster_update_fast", ""), f


def test_lattice_builders_match_reference_examples():
    from isingmontecarlo_b200 import lattices

    sq = lattices.square_periodic(24)  # examples/crash_check.rs:13-32
    assert len(sq) == 2 * 24 * 24 and lattices.nvars_of(sq) == 576
    assert sq[0] == ((0, 1), 1.0) and sq[1] == ((24, 25), 1.0)  # i-major enumeration, f(i,j) = j*L + i
    assert sq[576] == ((0, 24), 1.0)
    mixed = lattices.two_d_periodic_mixed(4)  # tests/longitudinal_crash.rs:5-23
    assert sum(1 for _, j in mixed if j > 0) == 8
    assert lattices.one_d_periodic(3) == [((0, 1), 1.0), ((1, 2), 1.0), ((2, 0), 1.0)]
    tri = lattices.triangular_periodic(48)
    assert len(tri) == 6912 and lattices.nvars_of(tri) == 2304


def test_partition_slots():
    from isingmontecarlo_b200.tempering import partition_slots

    assert [partition_slots(8192, 8, r) for r in (0, 1, 7)] == [(0, 1024), (1024, 1024), (7168, 1024)]
    with pytest.raises(ValueError):
        partition_slots(10, 4, 0)


WORKER = r"""
import os, sys
sys.path.insert(0, sys.argv[1])
import numpy as np, torch, torch.distributed as dist
from isingmontecarlo_b200.tempering import gather_records, partition_slots
from oracle import pyoracle as po
from isingmontecarlo_b200 import lattices
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
n_chains, n_betas = 2, 4
S = n_chains * n_betas
begin, R = partition_slots(S, world, rank)
betas = np.linspace(0.5, 2.0, n_betas)
edges = lattices.small_qmc_ring()
# every rank owns the configurations [begin, begin+R) and steps them with the CPU oracle (stand-in for
# the GPU handle in this plumbing test); records travel through the same gather_records() as on NCCL
mine = {g: po.SseOracle(edges, 1.0, 0.0, 4, key=100 + g) for g in range(begin, begin + R)}
slot_of = {g: g for g in mine}
full = [po.SseOracle(edges, 1.0, 0.0, 4, key=100 + g) for g in range(S)]  # single-process reference ladder
cursors = [0] * n_chains
for step in range(6):
    for g, q in mine.items():
        q.timesteps(2, float(betas[slot_of[g] % n_betas]))
    for c in range(n_chains):
        for k in range(n_betas):
            full[c * n_betas + k].timesteps(2, float(betas[k]))
    rec = torch.tensor([[slot_of[g], mine[g].n, mine[g].cursor, mine[g].cutoff] for g in sorted(mine)], dtype=torch.int64)
    allrec = gather_records(rec).numpy()
    assert allrec.shape == (S, 4)
    # reference decisions on the single-process ladder (configurations move between fixed slots)
    for c in range(n_chains):
        _, cursors[c] = po.pt_step(full[c * n_betas:(c + 1) * n_betas], betas, 7 + c, cursors[c])
    # labels-move bookkeeping from the gathered records: config g sits in the slot whose reference
    # graph now holds its n (ties broken by cursor-independent state comparison below)
    n_by_slot = {int(r[0]): int(r[1]) for r in allrec}
    assert sorted(n_by_slot) == list(range(S))  # every slot reported exactly once across ranks
    for g, q in mine.items():
        # find the new slot of config g: the reference slot whose state/ops equal q's
        new = [s for s in range(S) if s // n_betas == slot_of[g] // n_betas and full[s].n == q.n
               and np.array_equal(full[s].dump_ops()[:q.cutoff], q.dump_ops()) and np.array_equal(full[s].state(), q.state())]
        assert len(new) >= 1
        s_new = new[0]
        # relabel: take the slot's cursor (stays with the slot) and the ladder's max cutoff
        q.set_cursor(full[s_new].cursor)
        q.set_cutoff(full[s_new].cutoff)
        q.set_key(100 + s_new)  # the rng key is a slot label too
        slot_of[g] = s_new
print("rank", rank, "ok")
dist.destroy_process_group()
"""


def test_tempering_gather_plumbing_gloo_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", "29531", str(script), ROOT], capture_output=True, text=True, timeout=300, env=env)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert res.stdout.count("ok") == 2


def test_rust_sys_crate_binds_every_exported_symbol():
    """rust/qmcb-sys/src/lib.rs is generated from include/qmcb.h (tools/gen_rust_sys.py): it must be up to date, declare
    every function the header declares with the same arity, and the safe crate must only call symbols that exist."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("gen_rust_sys", os.path.join(ROOT, "tools", "gen_rust_sys.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    src, protos = gen.generate()
    assert open(gen.OUT).read() == src, "run python tools/gen_rust_sys.py"
    header = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "qmcb.h")).read(), flags=re.S)
    declared = set(re.findall(r"\b((?:qmcb|cmcb)_[a-z0-9_]+)\s*\(", header))
    bound = {m.group(1): m.group(2) for m in re.finditer(r"pub fn ((?:qmcb|cmcb)_\w+)\((.*?)\) ->", src)}
    assert set(bound) == declared
    for ret, name, params in protos:
        assert bound[name].count(":") == len(params), name
    safe = open(os.path.join(ROOT, "rust", "qmcb", "src", "lib.rs")).read()
    used = set(re.findall(r"sys::((?:qmcb|cmcb)_\w+)", safe))
    assert used <= declared and len(used) > 35, used - declared
    for t in ("parity.rs", "rand_kat.rs"):
        assert os.path.exists(os.path.join(ROOT, "rust", "qmcb", "tests", t))


def test_rand_known_answer_fixture_is_what_the_oracle_computes():
    """tests/golden/rand_kat.json freezes the oracle's restatement of rand 0.8; rust/qmcb/tests/rand_kat.rs checks the
    same file against rand itself on a machine with cargo.  Also re-derived here with Python integers, independently
    of oracle.c, for the integer mappings."""
    import ctypes as C
    import json

    from oracle import pyoracle as po

    with open(os.path.join(ROOT, "tests", "golden", "rand_kat.json")) as f:
        kat = json.load(f)
    L, key = po.lib(), kat["key"]
    words = [int(w, 16) for w in kat["words_hex"]]
    assert words == [int(L.orc_stream_word(key, c)) for c in range(64)]
    for case in kat["gen_bool"]:
        p = float.fromhex(case["p_hex"])
        thr = int(p * 2.0 ** 64)  # (p * 2^64) as u64: exact in Python for p < 1
        assert case["results"] == [int(w < thr) for w in words]
        assert case["results"] == [int(L.orc_gen_bool(key, C.byref(C.c_uint64(c)), p)) for c in range(64)]
    for case in kat["gen_range_usize"]:
        n = case["n"]
        zone = ((n << (64 - n.bit_length())) - 1) & (2 ** 64 - 1)
        cur, calls = 0, []
        while cur < 64:
            while cur < 64 and ((words[cur] * n) & (2 ** 64 - 1)) > zone:
                cur += 1
            if cur >= 64:
                break
            calls.append([(words[cur] * n) >> 64, cur + 1])
            cur += 1
        assert calls[:len(case["calls"])] == case["calls"] and len(calls) - len(case["calls"]) <= 1
    assert kat["gen_std_bool"] == [int((w >> 32) >= 2 ** 31) for w in words]
    assert kat["gen_f64"] == [float((w >> 11) * 2.0 ** -53).hex() for w in words]
    unit = [(float.fromhex("0x1." + "%013x" % (w >> 12) + "p+0") - 1.0).hex() for w in words]
    assert [v for v, _ in kat["gen_range_f64_unit"]] == unit


def test_host_mirrors_offer_the_reference_method_names():
    """The drop-in boundary seen from the host side: the Python mirrors carry the public method names of the reference's
    types (static lists, written down from qmc_ising.rs, qmc_stepper.rs, qmc_runner.rs, autocorrelations.rs,
    tempering_container.rs and classical/graph.rs; the reference tree is not read at test time)."""
    from isingmontecarlo_b200.classical import GraphState
    from isingmontecarlo_b200.sse import Qmc, QmcIsingGraph
    from isingmontecarlo_b200.tempering import TemperingContainer

    ising = """new_with_rng new_from_graph single_diagonal_step single_cluster_step single_rvb_sweep set_run_rvb set_enable_heatbath
        print_debug clone_state into_vec get_nvars get_edges get_transverse_field get_longitudinal_field get_cutoff set_cutoff
        get_offset rvb_success_rate into_qmc verify timestep timesteps timesteps_sample timesteps_measure timesteps_sample_iter
        timesteps_sample_iter_zip get_n state_ref get_bond_count get_energy_for_average_n imaginary_time_fold
        calculate_variable_autocorrelation calculate_spin_product_autocorrelation calculate_bond_autocorrelation""".split()
    qmc = """make_interaction make_interaction_and_offset make_diagonal_interaction make_diagonal_interaction_and_offset
        diagonal_update loop_update set_do_heatbath should_do_heatbath set_do_loop_updates should_do_loop_update
        should_do_cluster_update get_bonds get_offset increase_cutoff_to clone_state into_vec timestep timesteps""".split()
    tempering = """timesteps tempering_step timesteps_sample num_graphs get_total_swaps verify calculate_variable_autocorrelation
        calculate_spin_product_autocorrelation calculate_bond_autocorrelation""".split()
    classical = """new new_with_state_and_rng do_time_step enable_edge_importance_sampling get_state clone_state state_ref
        set_state get_energy""".split()
    for cls, names in ((QmcIsingGraph, ising), (Qmc, qmc), (TemperingContainer, tempering), (GraphState, classical)):
        missing = [n for n in names if not hasattr(cls, n)]
        assert not missing, (cls.__name__, missing)
