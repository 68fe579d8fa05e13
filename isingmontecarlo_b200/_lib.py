"""ctypes loader for libqmcb.so (the C ABI of include/qmcb.h).  There is no Python or CPU
fallback: if the CUDA library has not been built, importing the engine fails loudly."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("QMCB_LIB") or os.path.join(_HERE, "_build", "libqmcb.so")  # QMCB_LIB: kernel experiments only

OK, ERR_BAD_ARG, ERR_CAPACITY, ERR_CUDA, ERR_UNSUPPORTED, ERR_INTERNAL, ERR_NCCL = 0, -1, -2, -3, -4, -5, -6
MODE_STRICT, MODE_FAST, MODE_COUNTER = 0, 1, 2
OP_EMPTY = 0xFFFFFFFF


class QmcbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"qmcb error {code}: {msg}")
        self.code = code


class Lattice(C.Structure):
    _fields_ = [("nvars", C.c_uint32), ("nedges", C.c_uint32), ("va", C.POINTER(C.c_uint32)),
                ("vb", C.POINTER(C.c_uint32)), ("J", C.POINTER(C.c_double)), ("transverse", C.c_double),
                ("longitudinal", C.c_double)]


u8p, u32p, u64p, f64p = (C.POINTER(t) for t in (C.c_uint8, C.c_uint32, C.c_uint64, C.c_double))


class Interactions(C.Structure):
    _fields_ = [("nvars", C.c_uint32), ("n_interactions", C.c_uint32), ("nv", u32p), ("vars", u32p), ("mat_len", u32p), ("mats", f64p),
                ("offset", C.c_double), ("do_loop_updates", C.c_int)]

vp, vpp = C.c_void_p, C.POINTER(C.c_void_p)

# name -> (argtypes); every function returns int except the two string getters
SIGNATURES = {
    "qmcb_create": [C.POINTER(Lattice), C.c_uint32, f64p, u64p, C.c_uint64, C.c_uint64, u8p, C.c_int, vpp],
    "qmcb_create_qmc": [C.POINTER(Interactions), C.c_uint32, f64p, u64p, C.c_uint64, C.c_uint64, u8p, C.c_int, vpp],
    "qmcb_destroy": [vp],
    "qmcb_set_stream": [vp, vp],
    "qmcb_set_mode": [vp, C.c_int],
    "qmcb_get_mode": [vp, C.POINTER(C.c_int)],
    "qmcb_set_enable_heatbath": [vp, C.c_int],
    "qmcb_get_enable_heatbath": [vp, C.POINTER(C.c_int)],
    "qmcb_set_hamiltonians": [vp, C.c_uint32, f64p, f64p, f64p, u32p],
    "qmcb_num_hamiltonians": [vp, u32p],
    "qmcb_get_hamiltonian_index": [vp, u32p],
    "qmcb_get_offsets": [vp, f64p],
    "qmcb_set_option": [vp, C.c_char_p, C.c_int64],
    "qmcb_get_debug_counters": [vp, u64p],
    "qmcb_set_betas": [vp, f64p],
    "qmcb_get_betas": [vp, f64p],
    "qmcb_num_replicas": [vp, u32p],
    "qmcb_num_vars": [vp, u32p],
    "qmcb_num_bonds": [vp, u32p],
    "qmcb_num_edges": [vp, u32p],
    "qmcb_get_edges": [vp, u32p, u32p, f64p],
    "qmcb_get_fields": [vp, f64p, f64p],
    "qmcb_timesteps": [vp, C.c_uint64, C.c_uint64, f64p, u8p],
    "qmcb_enqueue_sweeps": [vp, C.c_uint64],
    "qmcb_synchronize": [vp],
    "qmcb_single_diagonal_step": [vp],
    "qmcb_single_cluster_step": [vp, u64p],
    "qmcb_loop_update": [vp],
    "qmcb_set_do_loop_updates": [vp, C.c_int],
    "qmcb_get_do_loop_updates": [vp, C.POINTER(C.c_int)],
    "qmcb_pt_variable_autocorrelation": [vp, C.c_uint64, C.c_uint64, C.c_uint64, f64p, u8p, f64p],
    "qmcb_pt_spin_product_autocorrelation": [vp, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, u32p, u32p, f64p, u8p, f64p],
    "qmcb_pt_bond_autocorrelation": [vp, C.c_uint64, C.c_uint64, C.c_uint64, f64p, u8p, f64p],
    "qmcb_set_run_rvb": [vp, C.c_int],
    "qmcb_get_run_rvb": [vp, C.POINTER(C.c_int)],
    "qmcb_single_rvb_sweep": [vp, C.c_int64, u64p, u64p],
    "qmcb_rvb_success_rate": [vp, f64p, u64p, u64p],
    "qmcb_total_vertex_updates": [vp, u64p],
    "qmcb_launch_count": [vp, u64p],
    "qmcb_get_state": [vp, C.c_uint32, u8p],
    "qmcb_get_states": [vp, u8p],
    "qmcb_set_state": [vp, C.c_uint32, u8p],
    "qmcb_get_n": [vp, u64p],
    "qmcb_get_cutoffs": [vp, u64p],
    "qmcb_set_cutoff": [vp, C.c_uint32, C.c_uint64],
    "qmcb_get_capacity": [vp, u64p],
    "qmcb_get_offset": [vp, f64p],
    "qmcb_get_bond_counts": [vp, C.c_uint32, u64p],
    "qmcb_itime_magnetization": [vp, f64p, f64p, f64p],
    "qmcb_itime_state": [vp, C.c_uint32, C.c_uint64, u8p],
    "qmcb_variable_autocorrelation": [vp, C.c_uint64, C.c_uint64, f64p, u8p, f64p],
    "qmcb_spin_product_autocorrelation": [vp, C.c_uint64, C.c_uint64, C.c_uint32, u32p, u32p, f64p, u8p, f64p],
    "qmcb_bond_autocorrelation": [vp, C.c_uint64, C.c_uint64, f64p, u8p, f64p],
    "qmcb_get_rng_cursors": [vp, u64p],
    "qmcb_set_rng_cursor": [vp, C.c_uint32, C.c_uint64],
    "qmcb_get_rng_keys": [vp, u64p],
    "qmcb_dump_ops": [vp, C.c_uint32, u32p, C.c_uint64],
    "qmcb_load_ops": [vp, C.c_uint32, u32p, C.c_uint64, u8p],
    "qmcb_verify": [vp, C.c_uint32, C.POINTER(C.c_int)],
    "qmcb_get_boundaries": [vp, C.c_uint32, u32p, u32p, C.c_uint64],
    "qmcb_pt_configure": [vp, C.c_uint32, C.c_uint32, C.c_uint32, f64p, u64p, C.c_uint64],
    "qmcb_pt_set_slot_hamiltonians": [vp, u32p],
    "qmcb_pt_record_words": [vp, u32p],
    "qmcb_pt_export": [vp, vp],
    "qmcb_pt_apply": [vp, vp, C.c_uint64],
    "qmcb_pt_step_local": [vp],
    "qmcb_pt_comm_unique_id": [u8p],
    "qmcb_pt_comm_init": [vp, u8p, C.c_int, C.c_int],
    "qmcb_pt_comm_attach": [vp, vp, C.c_int, C.c_int],
    "qmcb_pt_step": [vp],
    "qmcb_pt_collective_bytes": [vp, u64p],
    "qmcb_pt_timesteps_sample": [vp, C.c_uint64, C.c_uint64, C.c_uint64, f64p, u8p, u32p],
    "qmcb_pt_total_swaps": [vp, u64p],
    "qmcb_pt_get_config": [vp, u32p, u32p, u32p],
    "qmcb_pt_get_slots": [vp, u32p],
    "qmcb_checkpoint_size": [vp, u64p],
    "qmcb_checkpoint_save": [vp, vp, C.c_uint64],
    "qmcb_checkpoint_load": [vp, C.c_uint64, C.c_int, vpp],
    "cmcb_create": [C.POINTER(Lattice), f64p, C.c_uint32, f64p, u64p, u8p, C.c_int, vpp],
    "cmcb_destroy": [vp],
    "cmcb_set_stream": [vp, vp],
    "cmcb_set_option": [vp, C.c_char_p, C.c_int64],
    "cmcb_sweeps": [vp, C.c_uint64],
    "cmcb_enqueue_sweeps": [vp, C.c_uint64],
    "cmcb_synchronize": [vp],
    "cmcb_get_state": [vp, C.c_uint32, u8p],
    "cmcb_get_states": [vp, u8p],
    "cmcb_set_states": [vp, u8p],
    "cmcb_energy": [vp, f64p],
    "cmcb_magnetization": [vp, f64p],
    "cmcb_do_time_step": [vp, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, u8p],
    "cmcb_spin_flips": [vp, C.c_uint64],
    "cmcb_edge_flips": [vp, C.c_uint64],
    "cmcb_worm_flips": [vp, C.c_uint64, C.c_int],
    "cmcb_enable_edge_importance_sampling": [vp, C.c_int],
    "cmcb_get_rng_cursors": [vp, u64p],
    "cmcb_set_rng_cursors": [vp, u64p],
    "cmcb_get_colours": [vp, u32p, u32p],
    "cmcb_get_sweep_count": [vp, u64p],
    "cmcb_set_sweep_count": [vp, C.c_uint64],
    "cmcb_layout": [vp, C.POINTER(C.c_int)],
    "cmcb_launch_count": [vp, u64p],
}
STRING_GETTERS = ("qmcb_last_error", "qmcb_version")

_lib = None


def load():
    """dlopen libqmcb.so and bind every symbol include/qmcb.h declares."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise ImportError(
            f"{SO_PATH} is missing: the CUDA library has not been built (run `python -c 'import __graft_entry__ as g; "
            "g.build()'` or `make -C isingmontecarlo_b200/csrc`). There is no CPU fallback.")
    L = C.CDLL(SO_PATH)
    for name, args in SIGNATURES.items():
        if not hasattr(L, name) and os.environ.get("QMCB_LIB"):
            continue  # kernel experiments against an older build
        f = getattr(L, name)
        f.restype, f.argtypes = C.c_int, args
    for name in STRING_GETTERS:
        getattr(L, name).restype = C.c_char_p
        getattr(L, name).argtypes = []
    _lib = L
    return L


def check(rc):
    if rc != OK:
        raise QmcbError(rc, load().qmcb_last_error().decode())


def ptr(arr, ctype):
    return arr.ctypes.data_as(C.POINTER(ctype))
