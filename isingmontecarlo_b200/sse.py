"""Host-side mirror of `qmc::sse::QmcIsingGraph` + `QmcStepper` (qmc_ising.rs, qmc_stepper.rs) for
a BATCH of replicas.  Same method names and argument meaning as the reference; wherever the
reference returns one value per graph this returns one per replica.  All work happens in
libqmcb.so through the C ABI (include/qmcb.h)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import MODE_FAST, MODE_STRICT, Lattice, check, ptr


def _lattice(edges, transverse, longitudinal, nvars=None):
    va = np.ascontiguousarray([e[0][0] for e in edges], dtype=np.uint32)
    vb = np.ascontiguousarray([e[0][1] for e in edges], dtype=np.uint32)
    J = np.ascontiguousarray([e[1] for e in edges], dtype=np.float64)
    if nvars is None:
        nvars = int(max(va.max(), vb.max())) + 1  # qmc_ising.rs:92
    lat = Lattice(nvars, len(edges), ptr(va, C.c_uint32), ptr(vb, C.c_uint32), ptr(J, C.c_double),
                  float(transverse), float(longitudinal))
    return lat, (va, vb, J)


class QmcIsingGraph:
    """R replicas of one transverse-field Ising lattice resident on one GPU."""

    def __init__(self, edges, transverse, longitudinal, cutoff, rng_keys, betas, state=None, capacity=0,
                 device=0, mode=MODE_STRICT, nvars=None):
        L = _lib.load()
        self._L = L
        self._edges = list(edges)
        self.transverse, self.longitudinal = float(transverse), float(longitudinal)
        keys = np.ascontiguousarray(rng_keys, dtype=np.uint64)
        self.R = len(keys)
        self._betas = np.ascontiguousarray(np.broadcast_to(np.asarray(betas, dtype=np.float64), (self.R,)))
        lat, self._keep = _lattice(edges, transverse, longitudinal, nvars)
        self.nvars = lat.nvars
        st = None
        if state is not None:
            st = np.ascontiguousarray(np.broadcast_to(np.asarray(state, dtype=np.uint8), (self.R, self.nvars)))
        h = C.c_void_p()
        check(L.qmcb_create(C.byref(lat), self.R, ptr(self._betas, C.c_double), ptr(keys, C.c_uint64), int(cutoff),
                            int(capacity), None if st is None else ptr(st, C.c_uint8), device, C.byref(h)))
        self._h = h
        self.set_mode(mode)

    # -- construction helpers with the reference's names (qmc_ising.rs:131-148) ---------------
    @classmethod
    def new_with_rng(cls, edges, transverse, longitudinal, cutoff, rng_keys, state=None, betas=1.0, **kw):
        return cls(edges, transverse, longitudinal, cutoff, rng_keys, betas, state=state, **kw)

    # -- checkpoints (replaces SerializeQmcGraph, qmc_ising.rs:1001-1087) ----------------------
    def save_checkpoint(self) -> bytes:
        """One blob for the whole batch: lattice, spins, operator strings, cutoffs and the injected
        stream position of every replica (include/qmcb.h, qmcb_checkpoint_save)."""
        nbytes = C.c_uint64()
        check(self._L.qmcb_checkpoint_size(self._h, C.byref(nbytes)))
        buf = np.zeros(nbytes.value, dtype=np.uint8)
        check(self._L.qmcb_checkpoint_save(self._h, C.c_void_p(buf.ctypes.data), nbytes.value))
        return buf.tobytes()

    @classmethod
    def from_checkpoint(cls, blob, device=0):
        """SerializeQmcGraph::into_qmc: a batch that continues bit-identically to the saved one."""
        L = _lib.load()
        raw = np.frombuffer(blob, dtype=np.uint8)
        h = C.c_void_p()
        check(L.qmcb_checkpoint_load(C.c_void_p(raw.ctypes.data), len(raw), device, C.byref(h)))
        g = cls.__new__(cls)
        g._L, g._h = L, h
        n32 = C.c_uint32()
        check(L.qmcb_num_replicas(h, C.byref(n32)))
        g.R = n32.value
        check(L.qmcb_num_vars(h, C.byref(n32)))
        g.nvars = n32.value
        check(L.qmcb_num_edges(h, C.byref(n32)))
        va, vb, J = np.zeros(n32.value, np.uint32), np.zeros(n32.value, np.uint32), np.zeros(n32.value, np.float64)
        check(L.qmcb_get_edges(h, ptr(va, C.c_uint32), ptr(vb, C.c_uint32), ptr(J, C.c_double)))
        g._edges = [((int(a), int(b)), float(j)) for a, b, j in zip(va, vb, J)]
        t, l = C.c_double(), C.c_double()
        check(L.qmcb_get_fields(h, C.byref(t), C.byref(l)))
        g.transverse, g.longitudinal = t.value, l.value
        g._keep = None
        g._betas = g.betas()
        m = C.c_int()
        check(L.qmcb_get_mode(h, C.byref(m)))
        g.mode = m.value
        return g

    def close(self):
        if getattr(self, "_h", None):
            self._L.qmcb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- configuration -------------------------------------------------------------------------
    def set_mode(self, mode):
        check(self._L.qmcb_set_mode(self._h, mode))
        self.mode = mode

    def set_run_rvb(self, run_rvb):
        """QmcIsingGraph::set_run_rvb (qmc_ising.rs:434-441): (nvars + 1) / 2 RVB updates (rvb.rs:60-291) in every sweep, between
        the diagonal and the cluster update (:705-752)."""
        check(self._L.qmcb_set_run_rvb(self._h, int(bool(run_rvb))))

    def single_rvb_sweep(self, updates_in_sweep=None):
        """qmc_ising.rs:322-420 for every replica: (successes per replica, attempts)."""
        succ = np.zeros(self.R, dtype=np.uint64)
        att = C.c_uint64(0)
        check(self._L.qmcb_single_rvb_sweep(self._h, -1 if updates_in_sweep is None else int(updates_in_sweep),
                                            succ.ctypes.data_as(C.POINTER(C.c_uint64)), C.byref(att)))
        return succ, int(att.value)

    def rvb_success_rate(self):
        """qmc_ising.rs:604-607 for every replica (NaN before the first sweep with RVB steps)."""
        rate = np.zeros(self.R, dtype=np.float64)
        check(self._L.qmcb_rvb_success_rate(self._h, rate.ctypes.data_as(C.POINTER(C.c_double)), None, None))
        return rate

    def set_enable_heatbath(self, enable_heatbath):
        """qmc_ising.rs:444-486: heat-bath diagonal updates (heatbath.rs:149-209) for every replica."""
        check(self._L.qmcb_set_enable_heatbath(self._h, int(bool(enable_heatbath))))

    def get_enable_heatbath(self):
        out = C.c_int(0)
        check(self._L.qmcb_get_enable_heatbath(self._h, C.byref(out)))
        return bool(out.value)

    def set_hamiltonians(self, J_tab, transverse, longitudinal, ham_of_replica):
        """Replicas with their own couplings (the reference builds one graph per Hamiltonian): J_tab[H][E],
        transverse[H], longitudinal[H]; ham_of_replica[R] picks the row.  can_swap_managers
        (qmc_ising.rs:563-590) must hold between rows."""
        J = np.ascontiguousarray(J_tab, dtype=np.float64).reshape(len(transverse), -1)
        t = np.ascontiguousarray(transverse, dtype=np.float64)
        l = np.ascontiguousarray(longitudinal, dtype=np.float64)
        hr = np.ascontiguousarray(ham_of_replica, dtype=np.uint32)
        if J.shape[1] != len(self._edges) or len(l) != len(t) or len(hr) != self.R:
            raise ValueError("Hamiltonian table shapes do not match the batch")
        check(self._L.qmcb_set_hamiltonians(self._h, len(t), ptr(J, C.c_double), ptr(t, C.c_double), ptr(l, C.c_double),
                                            ptr(hr, C.c_uint32)))

    def hamiltonian_index(self):
        out = np.zeros(self.R, dtype=np.uint32)
        check(self._L.qmcb_get_hamiltonian_index(self._h, ptr(out, C.c_uint32)))
        return out

    def get_offsets(self):
        out = np.zeros(self.R, dtype=np.float64)
        check(self._L.qmcb_get_offsets(self._h, ptr(out, C.c_double)))
        return out

    def set_option(self, name, value):
        check(self._L.qmcb_set_option(self._h, name.encode(), int(value)))

    def debug_counters(self):
        out = np.zeros(64, dtype=np.uint64)
        check(self._L.qmcb_get_debug_counters(self._h, ptr(out, C.c_uint64)))
        return out

    def set_stream(self, cuda_stream_ptr):
        check(self._L.qmcb_set_stream(self._h, C.c_void_p(cuda_stream_ptr)))

    def _set_beta(self, beta):
        if beta is None:
            return
        b = np.ascontiguousarray(np.broadcast_to(np.asarray(beta, dtype=np.float64), (self.R,)))
        if not np.array_equal(b, self._betas):
            check(self._L.qmcb_set_betas(self._h, ptr(b, C.c_double)))
            self._betas = b.copy()

    def betas(self):
        out = np.zeros(self.R, dtype=np.float64)
        check(self._L.qmcb_get_betas(self._h, ptr(out, C.c_double)))
        return out

    # -- QmcStepper (qmc_stepper.rs) -----------------------------------------------------------
    def timestep(self, beta=None):
        self._set_beta(beta)
        check(self._L.qmcb_timesteps(self._h, 1, 1, None, None))
        return self.state_ref()

    def timesteps(self, t, beta=None):
        """qmc_stepper.rs:17-20: t sweeps, returns the average energy of every replica."""
        self._set_beta(beta)
        e = np.zeros(self.R, dtype=np.float64)
        check(self._L.qmcb_timesteps(self._h, int(t), 1, ptr(e, C.c_double), None))
        return e

    def timesteps_sample(self, t, beta=None, sampling_freq=None, out_samples=None, out_energies=None):
        """qmc_stepper.rs:23-40: returns (samples[R][t // freq][N], energies[R]).  out_samples / out_energies: optional
        caller-owned buffers of exactly those shapes (e.g. pinned host memory) that the results are written into."""
        self._set_beta(beta)
        freq = 1 if sampling_freq is None else int(sampling_freq)
        k = int(t) // freq
        e = np.zeros(self.R, dtype=np.float64) if out_energies is None else out_energies
        s = np.zeros((self.R, max(k, 1), self.nvars), dtype=np.uint8) if out_samples is None else out_samples
        if e.shape != (self.R,) or e.dtype != np.float64 or s.shape != (self.R, max(k, 1), self.nvars) or s.dtype != np.uint8 \
                or not (e.flags.c_contiguous and s.flags.c_contiguous):
            raise ValueError("output buffers must be C-contiguous float64[R] and uint8[R][t // freq][N]")
        check(self._L.qmcb_timesteps(self._h, int(t), freq, ptr(e, C.c_double), ptr(s, C.c_uint8)))
        return s[:, :k], e

    def timesteps_measure(self, timesteps, beta, init_t, state_fold, sampling_freq=None):
        """qmc_stepper.rs:85-103: fold `state_fold(acc, states)` over the sampled states; here `states` is the
        [R][N] array of all replicas at one sampling time.  Returns (acc, energies[R])."""
        samples, e = self.timesteps_sample(timesteps, beta, sampling_freq)
        acc = init_t
        for k in range(samples.shape[1]):
            acc = state_fold(acc, samples[:, k])
        return acc, e

    def timesteps_sample_iter(self, t, beta, sampling_freq, iter_fn):
        """qmc_stepper.rs:43-56: call iter_fn(states[R][N]) at every sampling time; returns the energies."""
        _, e = self.timesteps_measure(t, beta, None, lambda _acc, st: iter_fn(st), sampling_freq)
        return e

    def timesteps_sample_iter_zip(self, t, beta, sampling_freq, zip_with, iter_fn):
        """qmc_stepper.rs:59-77: iter_fn(item, states) for the items of zip_with, one per sampling time."""
        it = iter(zip_with)

        def step(_acc, st):
            try:
                iter_fn(next(it), st)
            except StopIteration:
                pass

        _, e = self.timesteps_measure(t, beta, None, step, sampling_freq)
        return e

    def enqueue_sweeps(self, t):
        check(self._L.qmcb_enqueue_sweeps(self._h, int(t)))

    def synchronize(self):
        check(self._L.qmcb_synchronize(self._h))

    def single_diagonal_step(self, beta=None):
        self._set_beta(beta)
        check(self._L.qmcb_single_diagonal_step(self._h))

    def single_cluster_step(self):
        out = np.zeros(self.R, dtype=np.uint64)
        check(self._L.qmcb_single_cluster_step(self._h, ptr(out, C.c_uint64)))
        return out

    def get_energy_for_average_n(self, average_n, beta):
        return -(np.asarray(average_n, dtype=np.float64) / beta) + self.get_offset()  # qmc_ising.rs:805-809

    # -- accessors -----------------------------------------------------------------------------
    def _u64(self, fn):
        out = np.zeros(self.R, dtype=np.uint64)
        check(fn(self._h, ptr(out, C.c_uint64)))
        return out

    def get_n(self):
        return self._u64(self._L.qmcb_get_n)

    def get_cutoff(self):
        return self._u64(self._L.qmcb_get_cutoffs)

    def set_cutoff(self, cutoff, replica=None):
        for r in (range(self.R) if replica is None else [replica]):
            check(self._L.qmcb_set_cutoff(self._h, r, int(cutoff)))

    def rng_cursors(self):
        return self._u64(self._L.qmcb_get_rng_cursors)

    def set_rng_cursor(self, r, cursor):
        check(self._L.qmcb_set_rng_cursor(self._h, r, int(cursor)))

    def rng_keys(self):
        return self._u64(self._L.qmcb_get_rng_keys)

    def get_capacity(self):
        c = C.c_uint64()
        check(self._L.qmcb_get_capacity(self._h, C.byref(c)))
        return c.value

    def get_nvars(self):
        return self.nvars

    def get_edges(self):
        return self._edges

    def get_offset(self):
        o = C.c_double()
        check(self._L.qmcb_get_offset(self._h, C.byref(o)))
        return o.value

    def num_bonds(self):
        n = C.c_uint32()
        check(self._L.qmcb_num_bonds(self._h, C.byref(n)))
        return n.value

    def state_ref(self):
        out = np.zeros((self.R, self.nvars), dtype=np.uint8)
        check(self._L.qmcb_get_states(self._h, ptr(out, C.c_uint8)))
        return out

    clone_state = state_ref  # qmc_ising.rs:502-504

    def into_vec(self):
        """qmc_ising.rs:507-509 / qmc_runner.rs:284-286: the p = 0 states; the batch is released"""
        st = self.state_ref().astype(bool)
        self.close()
        return st

    def get_transverse_field(self):  # qmc_ising.rs:522-524
        return self.transverse

    def get_longitudinal_field(self):  # qmc_ising.rs:527-529
        return self.longitudinal

    @classmethod
    def new_from_graph(cls, graph, transverse, longitudinal, cutoff, betas=1.0, mode=MODE_STRICT, device=0):
        """QmcIsingGraph::new_from_graph / new_qmc_from_graph (qmc_ising.rs:68-75, 151-166): the edges, the streams and the
        current states of a classical `GraphState` batch; its biases must be zero (the reference asserts it)."""
        if np.any(np.asarray(graph.biases) != 0.0):
            raise _lib.QmcbError(_lib.ERR_BAD_ARG, "new_from_graph: the classical graph has biases (qmc_ising.rs:157)")
        return cls(graph.get_edges(), transverse, longitudinal, cutoff, graph.rng_keys(), betas, state=graph.state_ref(), mode=mode, device=device)

    def print_debug(self, r=0, file=None):
        """qmc_ising.rs:489-494 -> debug_print_diagonal (diagonal.rs:193-234): the world lines of replica r, one row per slot"""
        import sys

        out = file or sys.stdout
        n = self.nvars
        E = len(self._edges)
        print("=" * n, file=out)
        print("".join("1" if b else "0" for b in self.state_ref()[r]), file=out)
        for p, w in enumerate(self.dump_ops(r)):
            if int(w) == _lib.OP_EMPTY:
                print("|" * n + f"\tp={p}", file=out)
                continue
            b, outs = int(w) & 0xFFFFFF, (int(w) >> 26) & 3
            vs = list(self._edges[b][0]) if b < E else [b - E if b < E + n else b - E - n]
            row, last = "", 0
            for var, bit in sorted((v, (outs >> k) & 1) for k, v in enumerate(vs)):
                row += "|" * (var - last) + str(bit)
                last = var + 1
            print(row + "|" * (n - last) + f"\tp={p}\t{b}: {vs}", file=out)

    def set_state(self, r, state):
        st = np.ascontiguousarray(state, dtype=np.uint8)
        check(self._L.qmcb_set_state(self._h, r, ptr(st, C.c_uint8)))

    def calculate_variable_autocorrelation(self, timesteps, beta=None, sampling_freq=None, return_samples=False):
        """QmcAutoCorrelations::calculate_variable_autocorrelation (autocorrelations.rs:48-61): [R][T] array, T =
        timesteps // sampling_freq; optionally also the samples it was computed from."""
        self._set_beta(beta)
        freq = 1 if sampling_freq is None else int(sampling_freq)
        T = int(timesteps) // freq
        out = np.zeros((self.R, T), dtype=np.float64)
        smp = np.zeros((self.R, T, self.nvars), dtype=np.uint8) if return_samples else None
        check(self._L.qmcb_variable_autocorrelation(self._h, int(timesteps), freq, ptr(out, C.c_double),
                                                    None if smp is None else ptr(smp, C.c_uint8), None))
        return (out, smp) if return_samples else out

    def calculate_spin_product_autocorrelation(self, timesteps, beta, var_products, sampling_freq=None, return_samples=False):
        """QmcAutoCorrelations::calculate_spin_product_autocorrelation (autocorrelations.rs:53-71): var_products is a list
        of variable lists; [R][T] array."""
        self._set_beta(beta)
        freq = 1 if sampling_freq is None else int(sampling_freq)
        T = int(timesteps) // freq
        off = np.zeros(len(var_products) + 1, dtype=np.uint32)
        off[1:] = np.cumsum([len(p) for p in var_products])
        vs = np.ascontiguousarray([v for p in var_products for v in p], dtype=np.uint32)
        out = np.zeros((self.R, T), dtype=np.float64)
        smp = np.zeros((self.R, T, self.nvars), dtype=np.uint8) if return_samples else None
        check(self._L.qmcb_spin_product_autocorrelation(self._h, int(timesteps), freq, len(var_products), ptr(off, C.c_uint32), ptr(vs, C.c_uint32),
                                                        ptr(out, C.c_double), None if smp is None else ptr(smp, C.c_uint8), None))
        return (out, smp) if return_samples else out

    def calculate_bond_autocorrelation(self, timesteps, beta=None, sampling_freq=None, return_samples=False):
        """QmcBondAutoCorrelations::calculate_bond_autocorrelation (autocorrelations.rs:80-97)"""
        self._set_beta(beta)
        freq = 1 if sampling_freq is None else int(sampling_freq)
        T = int(timesteps) // freq
        out = np.zeros((self.R, T), dtype=np.float64)
        smp = np.zeros((self.R, T, self.nvars), dtype=np.uint8) if return_samples else None
        check(self._L.qmcb_bond_autocorrelation(self._h, int(timesteps), freq, ptr(out, C.c_double), None if smp is None else ptr(smp, C.c_uint8), None))
        return (out, smp) if return_samples else out

    def imaginary_time_magnetization(self):
        """imaginary_time_fold (qmc_ising.rs:815-821) with the magnetisation fold on the device: per replica
        (<m>, <m^2>, <|m|>) of the per-site magnetisation over the M imaginary-time slices."""
        out = np.zeros((3, self.R), dtype=np.float64)
        check(self._L.qmcb_itime_magnetization(self._h, ptr(out[0], C.c_double), ptr(out[1], C.c_double), ptr(out[2], C.c_double)))
        return out[0], out[1], out[2]

    def imaginary_time_fold(self, r, fold_fn, init, stride=1):
        """The reference's closure form for one replica, evaluated on the host: fold_fn(acc, state) over the
        propagated states before slots 0, stride, 2*stride, ... (stride 1 = the reference's fold)."""
        acc = init
        st = np.zeros(self.nvars, dtype=np.uint8)
        for p in range(0, int(self.get_cutoff()[r]), stride):
            check(self._L.qmcb_itime_state(self._h, r, p, ptr(st, C.c_uint8)))
            acc = fold_fn(acc, st)
        return acc

    def get_bond_counts(self, r):
        out = np.zeros(self.num_bonds(), dtype=np.uint64)
        check(self._L.qmcb_get_bond_counts(self._h, r, ptr(out, C.c_uint64)))
        return out

    def get_bond_count(self, r, bond):
        return int(self.get_bond_counts(r)[bond])

    def dump_ops(self, r):
        m = int(self.get_cutoff()[r])
        out = np.zeros(max(m, 1), dtype=np.uint32)
        check(self._L.qmcb_dump_ops(self._h, r, ptr(out, C.c_uint32), len(out)))
        return out[:m]

    def load_ops(self, r, words, state=None):
        w = np.ascontiguousarray(words, dtype=np.uint32)
        st = None if state is None else np.ascontiguousarray(state, dtype=np.uint8)
        check(self._L.qmcb_load_ops(self._h, r, ptr(w, C.c_uint32), len(w), None if st is None else ptr(st, C.c_uint8)))

    def verify(self, r=None):
        ok = C.c_int()
        for k in (range(self.R) if r is None else [r]):
            check(self._L.qmcb_verify(self._h, k, C.byref(ok)))
            if not ok.value:
                return False
        return True

    def boundaries(self, r, nslots):
        a = np.zeros(nslots, dtype=np.uint32)
        b = np.zeros(nslots, dtype=np.uint32)
        check(self._L.qmcb_get_boundaries(self._h, r, ptr(a, C.c_uint32), ptr(b, C.c_uint32), nslots))
        return a, b

    def total_vertex_updates(self):
        t = C.c_uint64()
        check(self._L.qmcb_total_vertex_updates(self._h, C.byref(t)))
        return t.value

    def launch_count(self):
        t = C.c_uint64()
        check(self._L.qmcb_launch_count(self._h, C.byref(t)))
        return t.value


class Qmc(QmcIsingGraph):
    """`qmc::sse::Qmc` (qmc_runner.rs:22-403) for a batch of replicas: interactions are given as matrices
    (`make_interaction`, `make_diagonal_interaction`, the `*_and_offset` variants, :113-156) and the diagonal update takes
    its weights from those tables -- `qmcb_create_qmc`, a different code path from the (J, Gamma, h) arithmetic of
    QmcIsingGraph.  `timestep` is Qmc::timestep (:363-377): diagonal update, cluster update with Ising symmetry, free
    bits.  The handle is created when the first step (or accessor) needs it; the shape the engine takes is stated in
    include/qmcb.h (two-variable interactions first, then one constant one-variable interaction per variable, or none).
    `do_loop_updates` / `set_do_loop_updates` / `loop_update` are the reference's directed-loop update
    (directed_loop.rs:103-301; STRICT mode, one lane per replica).  Every QmcIsingGraph accessor works on it."""

    def __init__(self, nvars, rng_keys, betas=1.0, state=None, do_loop_updates=False, mode=MODE_STRICT, device=0, capacity=0):
        self._L = _lib.load()
        self._h = None
        self.nvars = int(nvars)
        self._keys = np.ascontiguousarray(rng_keys, dtype=np.uint64)
        self.R = len(self._keys)
        self._betas = np.ascontiguousarray(np.broadcast_to(np.asarray(betas, dtype=np.float64), (self.R,)))
        self._state0 = None if state is None else np.ascontiguousarray(np.broadcast_to(np.asarray(state, dtype=np.uint8), (self.R, self.nvars)))
        self._bonds = []  # (matrix, vars, diagonal)
        self._offset = 0.0
        self._cutoff0 = self.nvars  # Qmc::new_with_state: cutoff = nvars (qmc_runner.rs:77)
        self._mode0, self._device, self._capacity = mode, device, capacity
        self.mode = mode
        self.do_loop_updates = bool(do_loop_updates)
        self.transverse = self.longitudinal = None
        self._edges = []

    # -- interactions (qmc_runner.rs:113-156; Interaction::new / new_offset / new_diagonal / new_diagonal_offset :424-558)
    def _add(self, mat, variables, diagonal, and_offset):
        if self._h is not None:
            raise _lib.QmcbError(_lib.ERR_UNSUPPORTED, "interactions are fixed once the batch has been stepped")
        mat, variables = [float(x) for x in mat], [int(v) for v in variables]
        n = len(variables)
        if len(mat) != (1 << n if diagonal else 1 << (2 * n)):
            raise ValueError(f"Given {n} vars, matrix of {len(mat)} entries")
        if and_offset:  # subtract the smallest diagonal element and remember it
            idx = range(1 << n) if diagonal else [((1 << n) + 1) * k for k in range(1 << n)]
            lo = min(mat[i] for i in idx)
            for i in idx:
                mat[i] -= lo
            self._offset -= lo
        if not diagonal and any(x < 0.0 for x in mat):
            raise ValueError("Interaction contains negative weights")
        self._bonds.append((mat, variables, diagonal))

    def make_interaction(self, mat, variables):
        self._add(mat, variables, False, False)

    def make_interaction_and_offset(self, mat, variables):
        self._add(mat, variables, False, True)

    def make_diagonal_interaction(self, mat, variables):
        self._add(mat, variables, True, False)

    def make_diagonal_interaction_and_offset(self, mat, variables):
        self._add(mat, variables, True, True)

    def get_bonds(self):
        return [(list(v), list(m), "diagonal" if d else "full") for m, v, d in self._bonds]

    def get_offset(self):
        return self._offset

    def increase_cutoff_to(self, cutoff):  # qmc_runner.rs:307-309
        if self._h is None:
            self._cutoff0 = max(self._cutoff0, int(cutoff))
        else:
            cur = self.get_cutoff()
            for r in range(self.R):
                if cur[r] < cutoff:
                    self.set_cutoff(int(cutoff), r)

    def set_do_loop_updates(self, flag):  # qmc_runner.rs:268-270
        self.do_loop_updates = bool(flag)
        if self._h is not None:
            check(self._L.qmcb_set_do_loop_updates(self._h, int(bool(flag))))

    def set_do_heatbath(self, do_heatbath):  # qmc_runner.rs:258-260: heat-bath diagonal updates over the interactions' weights
        self._ensure()
        self.set_enable_heatbath(do_heatbath)

    def should_do_heatbath(self):  # :263-265
        return bool(self._h) and self.get_enable_heatbath()

    def diagonal_update(self, beta):
        """Qmc::diagonal_update (qmc_runner.rs:159-202): one diagonal update of every replica and the cutoff growth (:195)"""
        self._ensure()
        self._set_beta(beta)
        check(self._L.qmcb_single_diagonal_step(self._h))

    def should_do_loop_update(self):  # :273-275
        return self.do_loop_updates

    def loop_update(self):
        """Qmc::loop_update (qmc_runner.rs:205-220): one directed-loop update of every replica"""
        self._ensure()
        check(self._L.qmcb_loop_update(self._h))

    def should_do_cluster_update(self):  # :278-281; the shapes qmcb_create_qmc accepts keep the Ising symmetry
        return any(len(v) == 1 for _, v, _ in self._bonds)

    def _ensure(self):
        if self._h is not None:
            return
        n = len(self._bonds)
        nv = np.ascontiguousarray([len(v) for _, v, _ in self._bonds], dtype=np.uint32)
        vs = np.zeros(2 * max(n, 1), dtype=np.uint32)
        for b, (_, v, _) in enumerate(self._bonds):
            vs[2 * b:2 * b + len(v)] = v[:2]
        ml = np.ascontiguousarray([len(m) for m, _, _ in self._bonds], dtype=np.uint32)
        mats = np.ascontiguousarray([x for m, _, _ in self._bonds for x in m], dtype=np.float64)
        ints = _lib.Interactions(self.nvars, n, ptr(nv, C.c_uint32), ptr(vs, C.c_uint32), ptr(ml, C.c_uint32), ptr(mats, C.c_double),
                                 float(self._offset), int(self.do_loop_updates))
        h = C.c_void_p()
        check(self._L.qmcb_create_qmc(C.byref(ints), self.R, ptr(self._betas, C.c_double), ptr(self._keys, C.c_uint64), int(self._cutoff0),
                                      int(self._capacity), None if self._state0 is None else ptr(self._state0, C.c_uint8), self._device, C.byref(h)))
        self._h = h
        self._edges = [((int(v[0]), int(v[1])), 0.0) for _, v, _ in self._bonds if len(v) == 2]
        QmcIsingGraph.set_mode(self, self._mode0)

    def __getattribute__(self, name):
        # any call that reaches the C ABI needs the handle: build it on first use
        if name == "_L" and object.__getattribute__(self, "_h") is None and object.__getattribute__(self, "_bonds"):
            object.__getattribute__(self, "_ensure_guarded")()
        return object.__getattribute__(self, name)

    def _ensure_guarded(self):
        if not object.__getattribute__(self, "__dict__").get("_building"):
            self.__dict__["_building"] = True
            try:
                self._ensure()
            finally:
                self.__dict__["_building"] = False

    def set_mode(self, mode):
        self._mode0 = mode
        self.mode = mode
        if self._h is not None:
            QmcIsingGraph.set_mode(self, mode)


def _into_qmc(self):
    """IntoQmc::into_qmc (qmc_ising.rs:943-976): the same streams, states, cutoffs and operator strings; the
    Hamiltonian restated as interactions -- [-J, J, J, -J] (with its offset) per edge, [G, G, G, G] per variable.  The
    longitudinal interactions [h, 0, 0, -h] have a negative entry, which Interaction::new rejects (qmc_runner.rs:
    527-529): the reference's unwrap() panics there, this raises."""
    q = Qmc(self.nvars, self.rng_keys(), self.betas(), state=self.state_ref(), mode=self.mode)
    for (a, b), j in self.get_edges():
        q.make_diagonal_interaction_and_offset([-j, j, j, -j], [a, b])
    for v in range(self.nvars):
        q.make_interaction([self.transverse] * 4, [v])
    if abs(self.longitudinal) > np.finfo(np.float64).eps:
        for v in range(self.nvars):
            q.make_interaction([self.longitudinal, 0.0, 0.0, -self.longitudinal], [v])
    cut, cur = self.get_cutoff(), self.rng_cursors()
    q.increase_cutoff_to(int(cut.max()))
    q._ensure()
    for r in range(self.R):  # set_manager(self.op_manager) + the rng moves with the graph
        q.load_ops(r, self.dump_ops(r))
        q.set_cutoff(int(cut[r]), r)
        q.set_rng_cursor(r, int(cur[r]))
    return q


def _clone(self, device=0):
    """QmcIsingGraph::clone (qmc_ising.rs:909-933): an independent batch in the same state, stream positions included"""
    return QmcIsingGraph.from_checkpoint(self.save_checkpoint(), device=device)


QmcIsingGraph.into_qmc = _into_qmc
QmcIsingGraph.clone = _clone
DefaultQmcIsingGraph = QmcIsingGraph
