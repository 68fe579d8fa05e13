"""Host-side mirror of `qmc::sse::parallel_tempering::TemperingContainer`
(tempering_container.rs:19-302, rayon variants :316-478) over one batched GPU handle per rank.

Slots s = chain * n_betas + k are block-partitioned over the ranks of a torch.distributed group
(one process per GPU).  Operator strings never leave their GPU: a swap exchanges slot labels, and
the only inter-GPU traffic per tempering step is one all-gather of a 32-byte record per slot
(NCCL over NVLink on GPUs; gloo in the CPU tests of this plumbing)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import MODE_FAST, check, ptr

REC_WORDS = 4  # {slot, n, cursor, cutoff} as uint64 (int64 on the wire); 8 with per-slot Hamiltonians


def partition_slots(n_slots: int, world_size: int, rank: int):
    """Contiguous block of global slots owned by `rank` (weak scaling: fixed slots per GPU)."""
    if n_slots % world_size:
        raise ValueError("slots must divide evenly over the ranks")
    per = n_slots // world_size
    return rank * per, per


def gather_records(rec_local, group=None):
    """All-gather the per-configuration records of every rank, rank-major.  rec_local is a torch
    int64 tensor [R, 4] (CUDA with nccl, CPU with gloo).  Single process: returned unchanged."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return rec_local
    import torch

    world = dist.get_world_size(group)
    out = torch.empty((world * rec_local.shape[0], rec_local.shape[1]), dtype=rec_local.dtype, device=rec_local.device)
    dist.all_gather_into_tensor(out, rec_local.contiguous(), group=group)
    return out


class TemperingContainer:
    """n_chains independent ladders of n_betas slots each; add_qmc_stepper is replaced by giving
    the whole ladder at construction (all slots share one lattice, so can_swap_graphs holds)."""

    def __init__(self, edges, transverse, longitudinal, cutoff, betas, n_chains=1, rng_keys=None, pt_key=0x9E37,
                 mode=MODE_FAST, device=None, group=None, capacity=0, slot_hamiltonians=None, collective="library"):
        import torch
        import torch.distributed as dist

        from .sse import QmcIsingGraph

        self._torch = torch
        self.group = group
        dist_on = dist.is_available() and dist.is_initialized()
        self.rank = dist.get_rank(group) if dist_on else 0
        self.world = dist.get_world_size(group) if dist_on else 1
        betas = np.asarray(betas, dtype=np.float64)
        self.n_betas, self.n_chains = len(betas), int(n_chains)
        self.S = self.n_betas * self.n_chains
        self.betas_global = np.ascontiguousarray(np.tile(betas, self.n_chains))
        if rng_keys is None:
            rng_keys = 0x55E00000 + np.arange(self.S, dtype=np.uint64)
        self.keys_global = np.ascontiguousarray(rng_keys, dtype=np.uint64)
        self.slot_begin, self.R = partition_slots(self.S, self.world, self.rank)
        if device is None:
            device = torch.cuda.current_device()
        self.device = device
        sl = slice(self.slot_begin, self.slot_begin + self.R)
        self.graph = QmcIsingGraph(edges, transverse, longitudinal, cutoff, self.keys_global[sl], self.betas_global[sl],
                                   capacity=capacity, device=device, mode=mode)
        L = self.graph._L
        check(L.qmcb_pt_configure(self.graph._h, self.n_chains, self.n_betas, self.slot_begin,
                                  ptr(self.betas_global, C.c_double), ptr(self.keys_global, C.c_uint64), int(pt_key)))
        if slot_hamiltonians is not None:
            # one Hamiltonian per ladder position, as graphs with their own couplings handed to add_qmc_stepper
            # (tempering_container.rs:62-75); swaps use GraphWeights::relative_weight (tempering_traits.rs:126-154)
            if len(slot_hamiltonians) != self.n_betas:
                raise ValueError("one (J, transverse, longitudinal) per ladder position")
            J0 = np.array([e[1] for e in edges], dtype=np.float64)
            J_tab = np.stack([J0 if hj is None else np.asarray(hj, dtype=np.float64) for hj, _, _ in slot_hamiltonians])
            tr = [t for _, t, _ in slot_hamiltonians]
            lo = [l for _, _, l in slot_hamiltonians]
            ham_of_slot = np.ascontiguousarray(np.tile(np.arange(self.n_betas, dtype=np.uint32), self.n_chains))
            self.graph.set_hamiltonians(J_tab, tr, lo, ham_of_slot[sl])
            check(L.qmcb_pt_set_slot_hamiltonians(self.graph._h, ptr(ham_of_slot, C.c_uint32)))
        self._alloc_records()
        self._init_collective(collective)

    def _init_collective(self, collective):
        """collective="library": the all-gather runs inside libqmcb.so (ncclAllGather on the handle's stream, no host
        synchronisation: qmcb_pt_step); the ncclUniqueId of rank 0 is broadcast over the torch.distributed group.
        collective="torch": qmcb_pt_export + torch.distributed.all_gather_into_tensor + qmcb_pt_apply (the path a host
        with its own communication layer uses; also what runs over gloo)."""
        import torch.distributed as dist

        self.collective = collective if self.world > 1 else "local"
        if self.collective != "library":
            return
        if dist.get_backend(self.group) != "nccl":
            self.collective = "torch"
            return
        L, torch = self.graph._L, self._torch
        uid = torch.zeros(128, dtype=torch.uint8, device=f"cuda:{self.device}")
        if self.rank == 0:
            buf = np.zeros(128, dtype=np.uint8)
            check(L.qmcb_pt_comm_unique_id(ptr(buf, C.c_uint8)))
            uid.copy_(torch.from_numpy(buf))
        src = dist.get_global_rank(self.group, 0) if self.group is not None else 0
        dist.broadcast(uid, src=src, group=self.group)
        buf = np.ascontiguousarray(uid.cpu().numpy())
        check(L.qmcb_pt_comm_init(self.graph._h, ptr(buf, C.c_uint8), self.world, self.rank))

    def collective_bytes(self):
        b = C.c_uint64()
        check(self.graph._L.qmcb_pt_collective_bytes(self.graph._h, C.byref(b)))
        return b.value

    def _alloc_records(self):
        w = C.c_uint32()
        check(self.graph._L.qmcb_pt_record_words(self.graph._h, C.byref(w)))
        self._rec = self._torch.empty((self.R, w.value), dtype=self._torch.int64, device=f"cuda:{self.device}")

    # -- checkpoints (replaces SerializeTemperingContainer, tempering_container.rs:671-793) -----
    def save_checkpoint(self) -> bytes:
        """This rank's share of the container: its configurations with their current slot labels, the
        ladder, the PT stream position and the swap count."""
        return self.graph.save_checkpoint()

    @classmethod
    def from_checkpoint(cls, blob, device=None, group=None):
        import torch
        import torch.distributed as dist

        from .sse import QmcIsingGraph

        tc = cls.__new__(cls)
        tc._torch, tc.group = torch, group
        dist_on = dist.is_available() and dist.is_initialized()
        tc.rank = dist.get_rank(group) if dist_on else 0
        tc.world = dist.get_world_size(group) if dist_on else 1
        tc.device = torch.cuda.current_device() if device is None else device
        tc.graph = QmcIsingGraph.from_checkpoint(blob, device=tc.device)
        nc, nb, sb = C.c_uint32(), C.c_uint32(), C.c_uint32()
        check(tc.graph._L.qmcb_pt_get_config(tc.graph._h, C.byref(nc), C.byref(nb), C.byref(sb)))
        tc.n_chains, tc.n_betas, tc.slot_begin, tc.R = nc.value, nb.value, sb.value, tc.graph.R
        tc.S = tc.n_chains * tc.n_betas
        tc.betas_global = tc.keys_global = None  # live in the handle
        tc._alloc_records()
        tc._init_collective("library")
        return tc

    def num_graphs(self):
        return self.S

    def timesteps(self, t):
        """tempering_container.rs:76-81 / parallel_timesteps :366-371"""
        return self.graph.timesteps(t)

    def tempering_step(self):
        """tempering_container.rs:121-149 (parallel_tempering_step :373-402)"""
        L, g = self.graph._L, self.graph
        if self.collective in ("local", "library"):  # export, (ncclAllGather,) apply on the handle's stream: no host plumbing
            check(L.qmcb_pt_step(g._h))
            return
        check(L.qmcb_pt_export(g._h, C.c_void_p(self._rec.data_ptr())))
        check(L.qmcb_synchronize(g._h))
        allrec = gather_records(self._rec, self.group)
        self._allrec = allrec  # keep alive until the kernel ran
        # the collective is ordered on torch's current stream, the swap kernel runs on the handle's stream
        self._torch.cuda.current_stream().synchronize()
        check(L.qmcb_pt_apply(g._h, C.c_void_p(allrec.data_ptr()), allrec.shape[0]))
        check(L.qmcb_synchronize(g._h))

    def slots(self):
        out = np.zeros(self.R, dtype=np.uint32)
        check(self.graph._L.qmcb_pt_get_slots(self.graph._h, ptr(out, C.c_uint32)))
        return out

    def get_total_swaps(self):
        s = C.c_uint64()
        check(self.graph._L.qmcb_pt_total_swaps(self.graph._h, C.byref(s)))
        return s.value

    def timesteps_sample(self, timesteps, replica_swap_freq, sampling_freq):
        """tempering_container.rs:166-208: returns (states, energy_acc) indexed by GLOBAL slot.
        states[slot] is a list of sampled configurations (filled for the slots whose configuration
        lives on this rank at sampling time); energy_acc is summed over ranks."""
        if self.collective in ("local", "library"):  # the whole loop runs behind the C ABI
            T = int(timesteps) // int(sampling_freq)
            energy_acc = np.zeros(self.S, dtype=np.float64)
            samples = np.zeros((self.R, T, self.graph.nvars), dtype=np.uint8)
            sslots = np.zeros((self.R, T), dtype=np.uint32)
            check(self.graph._L.qmcb_pt_timesteps_sample(self.graph._h, int(timesteps), int(replica_swap_freq), int(sampling_freq),
                                                         ptr(energy_acc, C.c_double), ptr(samples, C.c_uint8), ptr(sslots, C.c_uint32)))
            states = [[] for _ in range(self.S)]
            for k in range(T):
                for r in range(self.R):
                    states[sslots[r, k]].append(samples[r, k].astype(bool))
            return states, energy_acc
        states = [[] for _ in range(self.S)]
        energy_acc = np.zeros(self.S, dtype=np.float64)
        remaining, to_swap, to_sample = int(timesteps), int(replica_swap_freq), int(sampling_freq)
        while remaining > 0:
            t = min(to_sample, to_swap, remaining)
            slots = self.slots()
            e = self.graph.timesteps(t)
            energy_acc[slots] += e * t
            to_sample -= t
            to_swap -= t
            remaining -= t
            if to_swap == 0:
                self.tempering_step()
                to_swap = int(replica_swap_freq)
            if to_sample == 0:
                st = self.graph.state_ref()
                for s, slot in enumerate(self.slots()):
                    states[slot].append(st[s])
                to_sample = int(sampling_freq)
        if self.world > 1:
            import torch.distributed as dist

            tacc = self._torch.from_numpy(energy_acc).to(f"cuda:{self.device}")
            dist.all_reduce(tacc, group=self.group)
            energy_acc = tacc.cpu().numpy()
        return states, energy_acc

    # -- ParallelTemperingAutocorrelations / ParallelTemperingBondAutoCorrelations (tempering_container.rs:484-630) --------
    def _autocorr(self, fn, timesteps, replica_swap_freq, sampling_freq, extra=(), return_samples=False):
        T = int(timesteps) // int(sampling_freq or 1)
        ac = np.zeros((self.S, T), dtype=np.float64)
        smp = np.zeros((self.S, T, self.graph.nvars), dtype=np.uint8) if return_samples else None
        check(fn(self.graph._h, int(timesteps), int(replica_swap_freq or 1), int(sampling_freq or 1), *extra, ptr(ac, C.c_double),
                 ptr(smp, C.c_uint8) if return_samples else None, None))
        return (ac, smp) if return_samples else ac

    def calculate_variable_autocorrelation(self, timesteps, replica_swap_freq=None, sampling_freq=None, return_samples=False):
        """:536-554: one autocorrelation per ladder slot, [S][T]"""
        return self._autocorr(self.graph._L.qmcb_pt_variable_autocorrelation, timesteps, replica_swap_freq, sampling_freq, (), return_samples)

    def calculate_spin_product_autocorrelation(self, timesteps, replica_swap_freq, var_products, sampling_freq=None, return_samples=False):
        """:556-578: series = products of the spins listed in each entry of var_products"""
        off = np.zeros(len(var_products) + 1, dtype=np.uint32)
        off[1:] = np.cumsum([len(p) for p in var_products])
        flat = np.ascontiguousarray(np.concatenate([np.asarray(p, dtype=np.uint32) for p in var_products]) if len(var_products) else np.zeros(0, np.uint32))
        return self._autocorr(self.graph._L.qmcb_pt_spin_product_autocorrelation, timesteps, replica_swap_freq, sampling_freq,
                              (len(var_products), ptr(off, C.c_uint32), ptr(flat, C.c_uint32)), return_samples)

    def calculate_bond_autocorrelation(self, timesteps, replica_swap_freq=None, sampling_freq=None, return_samples=False):
        """:608-630: one series per edge (value_for_bond, qmc_ising.rs:988-997)"""
        return self._autocorr(self.graph._L.qmcb_pt_bond_autocorrelation, timesteps, replica_swap_freq, sampling_freq, (), return_samples)

    def verify(self):
        return self.graph.verify()
