"""ctypes binding of the CPU oracle (oracle/oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module; the product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")

MODE_STRICT = 0
MODE_FAST = 1
MODE_COUNTER = 2
OP_EMPTY = 0xFFFFFFFF


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("oracle.c", "oracle.h", "Makefile")]
    stale = (not os.path.exists(_SO)) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    L = C.CDLL(build())
    u8p, u32p, u64p, i64p, f64p = (C.POINTER(t) for t in (C.c_uint8, C.c_uint32, C.c_uint64, C.c_int64, C.c_double))
    vp = C.c_void_p

    def sig(name, res, *args):
        f = getattr(L, name)
        f.restype, f.argtypes = res, list(args)

    sig("orc_philox4x32_10", None, u32p, u32p, u32p)
    sig("orc_stream_word", C.c_uint64, C.c_uint64, C.c_uint64)
    sig("orc_gen_bool", C.c_int, C.c_uint64, u64p, C.c_double)
    sig("orc_gen_range_usize", C.c_uint64, C.c_uint64, u64p, C.c_uint64)
    sig("orc_gen_range_u8", C.c_uint32, C.c_uint64, u64p, C.c_uint32)
    sig("orc_gen_range_f64_01", C.c_double, C.c_uint64, u64p)
    sig("orc_gen_range_f64", C.c_double, C.c_uint64, u64p, C.c_double, C.c_double)
    sig("orc_gen_f64", C.c_double, C.c_uint64, u64p)
    sig("orc_gen_std_bool", C.c_int, C.c_uint64, u64p)
    sig("orc_powi", C.c_double, C.c_double, C.c_int)
    sig("orc_bool_threshold", C.c_uint64, C.c_double)
    sig("orc_sse_create", vp, C.c_uint32, C.c_uint32, u32p, u32p, f64p, C.c_double, C.c_double, C.c_uint64, C.c_uint64, u8p)
    sig("orc_sse_destroy", None, vp)
    sig("orc_sse_set_script", None, vp, u64p, C.c_uint64)
    sig("orc_sse_error", C.c_int, vp)
    sig("orc_sse_set_enable_heatbath", None, vp, C.c_int)
    sig("orc_sse_get_enable_heatbath", C.c_int, vp)
    sig("orc_sse_timestep", None, vp, C.c_double, C.c_int)
    sig("orc_qmc_create", vp, C.c_uint32, C.c_uint64, u8p)
    sig("orc_qmc_make_interaction", C.c_int, vp, f64p, C.c_uint32, u32p, C.c_uint32, C.c_int, C.c_int)
    sig("orc_qmc_flags", C.c_int, vp)
    sig("orc_sse_use_small_rng", None, vp)
    sig("orc_qmc_loop_update", None, vp)
    sig("orc_sse_rvb_update", C.c_uint64, vp, C.c_uint64)
    sig("orc_sse_set_run_rvb", None, vp, C.c_int)
    sig("orc_sse_single_rvb_sweep", C.c_uint64, vp, C.c_int64, u64p)
    sig("orc_sse_rvb_success_rate", C.c_double, vp)
    sig("orc_qmc_set_do_loop_updates", None, vp, C.c_int)
    sig("orc_sse_single_diagonal_step", None, vp, C.c_double)
    sig("orc_sse_single_diagonal_step_mode", None, vp, C.c_double, C.c_int)
    sig("orc_sse_single_cluster_step", C.c_uint64, vp, C.c_int)
    sig("orc_sse_timesteps", C.c_double, vp, C.c_uint64, C.c_double, C.c_uint64, C.c_int, u8p)
    sig("orc_sse_nvars", C.c_uint32, vp)
    sig("orc_sse_get_n", C.c_uint64, vp)
    sig("orc_sse_get_cutoff", C.c_uint64, vp)
    sig("orc_sse_set_cutoff", None, vp, C.c_uint64)
    sig("orc_sse_get_cursor", C.c_uint64, vp)
    sig("orc_sse_set_cursor", None, vp, C.c_uint64)
    sig("orc_sse_set_key", None, vp, C.c_uint64)
    sig("orc_sse_get_key", C.c_uint64, vp)
    sig("orc_sse_get_offset", C.c_double, vp)
    sig("orc_sse_get_state", None, vp, u8p)
    sig("orc_sse_set_state", None, vp, u8p)
    sig("orc_sse_get_bond_count", C.c_uint64, vp, C.c_uint32)
    sig("orc_sse_itime_magnetization", C.c_uint64, vp, i64p)
    sig("orc_sse_itime_state", None, vp, C.c_uint64, u8p)
    sig("orc_sse_dump_ops", None, vp, u32p)
    sig("orc_sse_load_ops", C.c_int, vp, u32p, C.c_uint64, u8p)
    sig("orc_sse_verify", C.c_int, vp)
    sig("orc_sse_get_boundaries", None, vp, i64p, i64p, C.c_uint64)
    sig("orc_sse_batch_timesteps", C.c_uint64, C.POINTER(vp), C.c_uint32, C.c_uint64, f64p, C.c_int, f64p, C.c_int)
    sig("orc_sse_can_swap", C.c_int, vp, vp)
    sig("orc_pt_step", C.c_uint64, C.POINTER(vp), C.c_uint32, f64p, C.c_uint64, u64p)
    sig("orc_cls_create", vp, C.c_uint32, C.c_uint32, u32p, u32p, f64p, f64p, C.c_uint64, u8p)
    sig("orc_cls_destroy", None, vp)
    sig("orc_cls_spin_flips", None, vp, C.c_double, C.c_uint64)
    sig("orc_cls_enable_edge_importance_sampling", None, vp, C.c_int)
    sig("orc_cls_edge_flips", None, vp, C.c_double, C.c_uint64)
    sig("orc_cls_worm_flips", None, vp, C.c_double, C.c_uint64, C.c_int)
    sig("orc_cls_do_time_step", C.c_int, vp, C.c_double, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int)
    sig("orc_cls_get_error", C.c_int, vp)
    sig("orc_cls_set_cursor", None, vp, C.c_uint64)
    sig("orc_cls_checkerboard_sweeps", None, vp, C.c_double, u32p, C.c_uint32, C.c_uint64)
    sig("orc_cls_energy", C.c_double, vp)
    sig("orc_cls_magnetization", C.c_double, vp)
    sig("orc_cls_get_state", None, vp, u8p)
    sig("orc_cls_set_state", None, vp, u8p)
    sig("orc_cls_get_cursor", C.c_uint64, vp)
    sig("orc_cls_get_sweep", C.c_uint64, vp)
    sig("orc_cls_set_sweep", None, vp, C.c_uint64)
    sig("orc_cls_threshold", C.c_uint64, C.c_double, C.c_double)
    sig("orc_cls_batch_checkerboard", None, C.POINTER(vp), C.c_uint32, f64p, u32p, C.c_uint32, C.c_uint64, C.c_int)
    sig("orc_cls_batch_spin_flips", None, C.POINTER(vp), C.c_uint32, f64p, C.c_uint64, C.c_int)
    sig("orc_max_threads", C.c_int)
    _lib = L
    return L


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _edges(edges):
    ea = np.ascontiguousarray([e[0][0] for e in edges], dtype=np.uint32)
    eb = np.ascontiguousarray([e[0][1] for e in edges], dtype=np.uint32)
    J = np.ascontiguousarray([e[1] for e in edges], dtype=np.float64)
    return ea, eb, J


def philox(ctr, key):
    c = np.asarray(ctr, dtype=np.uint32)
    k = np.asarray(key, dtype=np.uint32)
    out = np.zeros(4, dtype=np.uint32)
    lib().orc_philox4x32_10(_p(c, C.c_uint32), _p(k, C.c_uint32), _p(out, C.c_uint32))
    return out


def stream_word(key: int, cursor: int) -> int:
    return lib().orc_stream_word(key, cursor)


class SseOracle:
    """One replica: QmcIsingGraph<R, FastOps> (qmc_ising.rs) driven by the Philox stream."""

    def __init__(self, edges, transverse, longitudinal, cutoff, key=0, state=None, nvars=None):
        ea, eb, J = _edges(edges)
        self.nvars = int(max(ea.max(), eb.max())) + 1 if nvars is None else nvars
        st = None if state is None else np.ascontiguousarray(state, dtype=np.uint8)
        self._h = lib().orc_sse_create(self.nvars, len(edges), _p(ea, C.c_uint32), _p(eb, C.c_uint32),
                                       _p(J, C.c_double), transverse, longitudinal, cutoff, key,
                                       None if st is None else _p(st, C.c_uint8))
        self._script = None

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.orc_sse_destroy(self._h)
            self._h = None

    def use_small_rng(self):
        """TIMING ONLY (bench.py's CPU arm): words from xoshiro256++ -- rand's SmallRng, which the reference's benches
        use (benches/end_to_end.rs:49) -- instead of Philox4x32-10; no GPU counterpart."""
        lib().orc_sse_use_small_rng(self._h)

    def set_script(self, words):
        self._script = np.ascontiguousarray(words, dtype=np.uint64)
        lib().orc_sse_set_script(self._h, _p(self._script, C.c_uint64), len(self._script))

    @property
    def error(self):
        return lib().orc_sse_error(self._h)

    def set_enable_heatbath(self, enable):
        """QmcIsingGraph::set_enable_heatbath (qmc_ising.rs:444-486)"""
        lib().orc_sse_set_enable_heatbath(self._h, int(bool(enable)))

    def set_run_rvb(self, run_rvb):  # qmc_ising.rs:434-441
        lib().orc_sse_set_run_rvb(self._h, int(bool(run_rvb)))

    def rvb_update(self, updates):  # rvb.rs:60-291 as qmc_ising.rs:705-752 calls it; returns the successes
        return int(lib().orc_sse_rvb_update(self._h, int(updates)))

    def single_rvb_sweep(self, updates_in_sweep=None):  # qmc_ising.rs:322-420 -> (successes, attempts)
        steps = C.c_uint64(0)
        succ = lib().orc_sse_single_rvb_sweep(self._h, -1 if updates_in_sweep is None else int(updates_in_sweep), C.byref(steps))
        return int(succ), int(steps.value)

    def rvb_success_rate(self):  # qmc_ising.rs:604-607
        return float(lib().orc_sse_rvb_success_rate(self._h))

    def timestep(self, beta, mode=MODE_STRICT):
        lib().orc_sse_timestep(self._h, beta, mode)

    def single_diagonal_step(self, beta, mode=MODE_STRICT):
        lib().orc_sse_single_diagonal_step_mode(self._h, beta, mode)

    def single_cluster_step(self, mode=MODE_STRICT):
        return lib().orc_sse_single_cluster_step(self._h, mode)

    def timesteps(self, t, beta, mode=MODE_STRICT):
        return lib().orc_sse_timesteps(self._h, t, beta, 1, mode, None)

    def timesteps_sample(self, t, beta, sampling_freq=1, mode=MODE_STRICT):
        k = t // sampling_freq
        buf = np.zeros((max(k, 1), self.nvars), dtype=np.uint8)
        e = lib().orc_sse_timesteps(self._h, t, beta, sampling_freq, mode, _p(buf, C.c_uint8))
        return buf[:k], e

    n = property(lambda s: lib().orc_sse_get_n(s._h))
    cutoff = property(lambda s: lib().orc_sse_get_cutoff(s._h))
    cursor = property(lambda s: lib().orc_sse_get_cursor(s._h))
    offset = property(lambda s: lib().orc_sse_get_offset(s._h))

    def set_cutoff(self, c):
        lib().orc_sse_set_cutoff(self._h, c)

    def set_cursor(self, c):
        lib().orc_sse_set_cursor(self._h, c)

    def set_key(self, k):
        lib().orc_sse_set_key(self._h, k)

    def state(self):
        out = np.zeros(self.nvars, dtype=np.uint8)
        lib().orc_sse_get_state(self._h, _p(out, C.c_uint8))
        return out

    def set_state(self, st):
        st = np.ascontiguousarray(st, dtype=np.uint8)
        lib().orc_sse_set_state(self._h, _p(st, C.c_uint8))

    def bond_count(self, b):
        return lib().orc_sse_get_bond_count(self._h, b)

    def itime_magnetization(self):
        """imaginary_time_fold with the magnetisation fold: (slots, sum m, sum m^2, sum |m|), m = sum_v (2 s_v - 1)"""
        sums = np.zeros(3, dtype=np.int64)
        slots = lib().orc_sse_itime_magnetization(self._h, _p(sums, C.c_int64))
        return slots, int(sums[0]), int(sums[1]), int(sums[2])

    def itime_state(self, p):
        out = np.zeros(self.nvars, dtype=np.uint8)
        lib().orc_sse_itime_state(self._h, int(p), _p(out, C.c_uint8))
        return out

    def dump_ops(self):
        out = np.zeros(self.cutoff, dtype=np.uint32)
        lib().orc_sse_dump_ops(self._h, _p(out, C.c_uint32))
        return out

    def load_ops(self, words, state=None):
        w = np.ascontiguousarray(words, dtype=np.uint32)
        st = None if state is None else np.ascontiguousarray(state, dtype=np.uint8)
        rc = lib().orc_sse_load_ops(self._h, _p(w, C.c_uint32), len(w), None if st is None else _p(st, C.c_uint8))
        if rc != 0:
            raise ValueError("bad op word")

    def verify(self):
        return bool(lib().orc_sse_verify(self._h))

    def boundaries(self, nslots=None):
        nslots = self.cutoff if nslots is None else nslots
        a = np.zeros(nslots, dtype=np.int64)
        b = np.zeros(nslots, dtype=np.int64)
        lib().orc_sse_get_boundaries(self._h, _p(a, C.c_int64), _p(b, C.c_int64), nslots)
        return a, b


def sse_batch_timesteps(reps, t, betas, mode=MODE_STRICT, nthreads=0):
    nthreads = nthreads or max_threads()
    arr = (C.c_void_p * len(reps))(*[r._h for r in reps])
    b = np.ascontiguousarray(betas, dtype=np.float64)
    e = np.zeros(len(reps), dtype=np.float64)
    tot = lib().orc_sse_batch_timesteps(arr, len(reps), t, _p(b, C.c_double), mode, _p(e, C.c_double), nthreads)
    return tot, e


def pt_step(slots, betas, pt_key, pt_cursor):
    """tempering_step over SseOracle slots; returns (swaps, new_cursor)."""
    arr = (C.c_void_p * len(slots))(*[r._h for r in slots])
    b = np.ascontiguousarray(betas, dtype=np.float64)
    cur = C.c_uint64(pt_cursor)
    swaps = lib().orc_pt_step(arr, len(slots), _p(b, C.c_double), pt_key, C.byref(cur))
    return swaps, cur.value


class ClassicalOracle:
    """GraphState (classical/graph.rs) with the reference rule and a checkerboard schedule."""

    def __init__(self, edges, biases, key=0, state=None):
        ea, eb, J = _edges(edges)
        bz = np.ascontiguousarray(biases, dtype=np.float64)
        self.nvars = len(bz)
        st = None if state is None else np.ascontiguousarray(state, dtype=np.uint8)
        self._h = lib().orc_cls_create(self.nvars, len(edges), _p(ea, C.c_uint32), _p(eb, C.c_uint32),
                                       _p(J, C.c_double), _p(bz, C.c_double), key,
                                       None if st is None else _p(st, C.c_uint8))

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.orc_cls_destroy(self._h)
            self._h = None

    def spin_flips(self, beta, count):
        lib().orc_cls_spin_flips(self._h, beta, count)

    # the reference's other moves and its own schedule (graph.rs:121-406)
    def enable_edge_importance_sampling(self, enable=True):
        lib().orc_cls_enable_edge_importance_sampling(self._h, int(bool(enable)))

    def edge_flips(self, beta, count):
        lib().orc_cls_edge_flips(self._h, beta, count)

    def worm_flips(self, beta, count=1, allow_doubles=True):
        lib().orc_cls_worm_flips(self._h, beta, count, int(bool(allow_doubles)))

    def do_time_step(self, beta, nspinupdates=None, nedgeupdates=None, nwormupdates=None, only_basic_moves=False):
        none = 2**64 - 1
        return lib().orc_cls_do_time_step(self._h, beta, none if nspinupdates is None else nspinupdates,
                                          none if nedgeupdates is None else nedgeupdates,
                                          none if nwormupdates is None else nwormupdates, int(bool(only_basic_moves)))

    error = property(lambda s: lib().orc_cls_get_error(s._h))

    def set_cursor(self, c):
        lib().orc_cls_set_cursor(self._h, int(c))

    def checkerboard_sweeps(self, beta, colours, nsweeps=1):
        col = np.ascontiguousarray(colours, dtype=np.uint32)
        lib().orc_cls_checkerboard_sweeps(self._h, beta, _p(col, C.c_uint32), int(col.max()) + 1, nsweeps)

    def energy(self):
        return lib().orc_cls_energy(self._h)

    def magnetization(self):
        return lib().orc_cls_magnetization(self._h)

    def state(self):
        out = np.zeros(self.nvars, dtype=np.uint8)
        lib().orc_cls_get_state(self._h, _p(out, C.c_uint8))
        return out

    def set_state(self, st):
        st = np.ascontiguousarray(st, dtype=np.uint8)
        lib().orc_cls_set_state(self._h, _p(st, C.c_uint8))

    cursor = property(lambda s: lib().orc_cls_get_cursor(s._h))
    sweep = property(lambda s: lib().orc_cls_get_sweep(s._h))

    def set_sweep(self, s):
        lib().orc_cls_set_sweep(self._h, s)


def cls_batch_checkerboard(reps, betas, colours, nsweeps, nthreads=0):
    nthreads = nthreads or max_threads()
    arr = (C.c_void_p * len(reps))(*[r._h for r in reps])
    b = np.ascontiguousarray(betas, dtype=np.float64)
    col = np.ascontiguousarray(colours, dtype=np.uint32)
    lib().orc_cls_batch_checkerboard(arr, len(reps), _p(b, C.c_double), _p(col, C.c_uint32), int(col.max()) + 1, nsweeps, nthreads)


def cls_batch_spin_flips(reps, betas, count, nthreads=0):
    nthreads = nthreads or max_threads()
    arr = (C.c_void_p * len(reps))(*[r._h for r in reps])
    b = np.ascontiguousarray(betas, dtype=np.float64)
    lib().orc_cls_batch_spin_flips(arr, len(reps), _p(b, C.c_double), count, nthreads)


def cls_threshold(beta, delta_e):
    return lib().orc_cls_threshold(beta, delta_e)


def max_threads():
    """Host cores this process may run on.  Not omp_get_max_threads(): torchrun exports OMP_NUM_THREADS=1 to its
    workers, which would silently turn the all-core CPU baseline into a one-core one."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


class QmcOracle(SseOracle):
    """Generic Qmc<R, FastOps> (qmc_runner.rs:22-403): interactions of one or two variables given as matrices;
    timestep = diagonal update, cluster update with Ising symmetry (when the interactions allow it), free-spin flips.
    Every SseOracle accessor works on it."""

    ERRORS = {1: "Matrix size must be power of 2", 2: "Given vars do not match the matrix size", 3: "Interaction contains negative weights",
              4: "interactions of more than two variables are not restated"}

    def __init__(self, nvars, key=0, state=None):
        self.nvars = int(nvars)
        st = None if state is None else np.ascontiguousarray(state, dtype=np.uint8)
        self._h = lib().orc_qmc_create(self.nvars, key, None if st is None else _p(st, C.c_uint8))
        self._script = None

    def _make(self, mat, vars_, diagonal, and_offset):
        m = np.ascontiguousarray(mat, dtype=np.float64)
        v = np.ascontiguousarray(list(vars_) + [0, 0], dtype=np.uint32)
        rc = lib().orc_qmc_make_interaction(self._h, _p(m, C.c_double), len(m), _p(v, C.c_uint32), len(vars_), int(diagonal), int(and_offset))
        if rc:
            raise ValueError(self.ERRORS.get(rc, str(rc)))

    def make_interaction(self, mat, vars_):  # qmc_runner.rs:113-122
        self._make(mat, vars_, False, False)

    def make_interaction_and_offset(self, mat, vars_):  # :125-135
        self._make(mat, vars_, False, True)

    def make_diagonal_interaction(self, mat, vars_):  # :138-146
        self._make(mat, vars_, True, False)

    def make_diagonal_interaction_and_offset(self, mat, vars_):  # :149-156
        self._make(mat, vars_, True, True)

    def loop_update(self):  # Qmc::loop_update, qmc_runner.rs:205-220
        lib().orc_qmc_loop_update(self._h)

    def set_do_loop_updates(self, enable):  # qmc_runner.rs:268-270
        lib().orc_qmc_set_do_loop_updates(self._h, int(bool(enable)))

    @property
    def has_cluster_edges(self):
        return bool(lib().orc_qmc_flags(self._h) & 1)

    @property
    def breaks_ising_symmetry(self):
        return bool(lib().orc_qmc_flags(self._h) & 2)


def into_qmc(ising, edges, transverse, longitudinal):
    """IntoQmc::into_qmc (qmc_ising.rs:943-976) for an SseOracle built from (edges, transverse, longitudinal): the same
    stream, state, cutoff and operator string, the Hamiltonian restated as generic interactions."""
    nvars = ising.nvars
    q = QmcOracle(nvars, key=lib().orc_sse_get_key(ising._h), state=ising.state())
    for (a, b), j in edges:
        q.make_diagonal_interaction_and_offset([-j, j, j, -j], [a, b])
    for v in range(nvars):
        q.make_interaction([transverse] * 4, [v])
    if abs(longitudinal) > np.finfo(np.float64).eps:
        for v in range(nvars):
            q.make_interaction([longitudinal, 0.0, 0.0, -longitudinal], [v])
    q.load_ops(ising.dump_ops(), ising.state())  # increase_cutoff_to(self.cutoff) + set_manager(self.op_manager)
    if q.cutoff < ising.cutoff:
        lib().orc_sse_set_cutoff(q._h, ising.cutoff)
    q.set_cursor(ising.cursor)
    return q
