"""Host-side mirror of `qmc::classical::graph::GraphState` (classical/graph.rs:56-88, :350-447) for
a batch of replicas swept with the checkerboard schedule."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import check, ptr
from .sse import _lattice


class GraphState:
    def __init__(self, edges, biases, rng_keys, betas, state=None, device=0):
        L = _lib.load()
        self._L = L
        bz = np.ascontiguousarray(biases, dtype=np.float64)
        self.nvars = len(bz)
        keys = np.ascontiguousarray(rng_keys, dtype=np.uint64)
        self.R = len(keys)
        self._edges, self.biases, self._keys = list(edges), bz.copy(), keys.copy()  # what QmcIsingGraph::new_from_graph takes over
        b = np.ascontiguousarray(np.broadcast_to(np.asarray(betas, dtype=np.float64), (self.R,)))
        lat, self._keep = _lattice(edges, 0.0, 0.0, self.nvars)
        st = None
        if state is not None:
            st = np.ascontiguousarray(np.broadcast_to(np.asarray(state, dtype=np.uint8), (self.R, self.nvars)))
        h = C.c_void_p()
        check(L.cmcb_create(C.byref(lat), ptr(bz, C.c_double), self.R, ptr(b, C.c_double), ptr(keys, C.c_uint64),
                            None if st is None else ptr(st, C.c_uint8), device, C.byref(h)))
        self._h = h

    @classmethod
    def new(cls, edges, biases, rng_keys, betas, **kw):
        return cls(edges, biases, rng_keys, betas, **kw)

    @classmethod
    def new_with_state_and_rng(cls, state, edges, biases, rng_keys, betas, **kw):
        return cls(edges, biases, rng_keys, betas, state=state, **kw)

    def get_edges(self):  # graph.rs: the edge list the graph was built from
        return self._edges

    def rng_keys(self):
        return self._keys

    def close(self):
        if getattr(self, "_h", None):
            self._L.cmcb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream_ptr):
        check(self._L.cmcb_set_stream(self._h, C.c_void_p(cuda_stream_ptr)))

    def set_option(self, name, value):
        check(self._L.cmcb_set_option(self._h, name.encode(), int(value)))

    def sweeps(self, nsweeps=1):
        """Checkerboard sweeps = every site once per sweep (the throughput path; the schedule is builder-defined,
        the per-site rule is the reference's)."""
        check(self._L.cmcb_sweeps(self._h, int(nsweeps)))

    def do_time_step(self, nspinupdates=None, nedgeupdates=None, nwormupdates=None, only_basic_moves=None):
        """GraphState::do_time_step (graph.rs:350-406) for every replica under its own sequential stream, at the
        replica's beta: one draw picks spin flips, edge flips or (unless only_basic_moves) worm flips.  Returns the
        move each replica drew."""
        none = 2**64 - 1
        out = np.zeros(self.R, dtype=np.uint8)
        check(self._L.cmcb_do_time_step(self._h, none if nspinupdates is None else int(nspinupdates),
                                        none if nedgeupdates is None else int(nedgeupdates),
                                        none if nwormupdates is None else int(nwormupdates),
                                        int(bool(only_basic_moves)), ptr(out, C.c_uint8)))
        return out

    def do_spin_flip(self, count=1):  # graph.rs:91-119
        check(self._L.cmcb_spin_flips(self._h, int(count)))

    def do_edge_flip(self, count=1):  # graph.rs:122-153
        check(self._L.cmcb_edge_flips(self._h, int(count)))

    def do_worm_flip(self, count=1, allow_doubles=True):  # graph.rs:179-318
        check(self._L.cmcb_worm_flips(self._h, int(count), int(bool(allow_doubles))))

    def enable_edge_importance_sampling(self, enable=True):  # graph.rs:321-336
        check(self._L.cmcb_enable_edge_importance_sampling(self._h, int(bool(enable))))

    def rng_cursors(self):
        out = np.zeros(self.R, dtype=np.uint64)
        check(self._L.cmcb_get_rng_cursors(self._h, ptr(out, C.c_uint64)))
        return out

    def set_rng_cursors(self, cursors):
        c = np.ascontiguousarray(np.broadcast_to(np.asarray(cursors, dtype=np.uint64), (self.R,)))
        check(self._L.cmcb_set_rng_cursors(self._h, ptr(c, C.c_uint64)))

    def enqueue_sweeps(self, nsweeps):
        check(self._L.cmcb_enqueue_sweeps(self._h, int(nsweeps)))

    def synchronize(self):
        check(self._L.cmcb_synchronize(self._h))

    def state_ref(self):
        out = np.zeros((self.R, self.nvars), dtype=np.uint8)
        check(self._L.cmcb_get_states(self._h, ptr(out, C.c_uint8)))
        return out

    clone_state = get_state = state_ref

    def set_state(self, states):
        st = np.ascontiguousarray(np.broadcast_to(np.asarray(states, dtype=np.uint8), (self.R, self.nvars)))
        check(self._L.cmcb_set_states(self._h, ptr(st, C.c_uint8)))

    def get_energy(self):
        out = np.zeros(self.R, dtype=np.float64)
        check(self._L.cmcb_energy(self._h, ptr(out, C.c_double)))
        return out

    def magnetization(self):
        out = np.zeros(self.R, dtype=np.float64)
        check(self._L.cmcb_magnetization(self._h, ptr(out, C.c_double)))
        return out

    def colours(self):
        col = np.zeros(self.nvars, dtype=np.uint32)
        n = C.c_uint32()
        check(self._L.cmcb_get_colours(self._h, ptr(col, C.c_uint32), C.byref(n)))
        return col, n.value

    def sweep_count(self):
        s = C.c_uint64()
        check(self._L.cmcb_get_sweep_count(self._h, C.byref(s)))
        return s.value

    def is_bitpacked_square(self):
        s = C.c_int()
        check(self._L.cmcb_layout(self._h, C.byref(s)))
        return bool(s.value)

    def launch_count(self):
        t = C.c_uint64()
        check(self._L.cmcb_launch_count(self._h, C.byref(t)))
        return t.value
