"""The multi-rank tempering code path on ONE GPU: two handles on the same device play rank 0 and rank 1 (slot_begin 0
and R), their qmcb_pt_export records are concatenated on the device and fed to both qmcb_pt_apply calls -- exactly what
the NCCL all-gather delivers on two GPUs -- and every configuration is compared with the oracle's literal restatement of
TemperingContainer::tempering_step (tempering_container.rs:121-149, :241-302).  Also: qmcb_pt_timesteps_sample
(tempering_container.rs:166-208) against the same loop over oracle graphs."""
import ctypes as C

import numpy as np
import pytest

from isingmontecarlo_b200 import MODE_FAST, MODE_STRICT, lattices
from isingmontecarlo_b200._lib import check, ptr
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu


def _oracle_ladders(edges, gamma, h, cutoff, keys, n_chains, n_betas):
    return [[po.SseOracle(edges, gamma, h, cutoff, key=int(keys[c * n_betas + k])) for k in range(n_betas)] for c in range(n_chains)]


@pytest.mark.parametrize("mode", [MODE_STRICT, MODE_FAST])
@pytest.mark.parametrize("n_chains,n_betas,with_h", [(2, 4, False), (4, 3, True), (2, 5, False)])
def test_two_handles_one_device_export_gather_apply(n_chains, n_betas, with_h, mode):
    import torch

    from isingmontecarlo_b200.sse import QmcIsingGraph

    edges = lattices.two_d_periodic_mixed(4)
    gamma, h = 1.0, (0.3 if with_h else 0.0)
    betas = np.linspace(0.5, 2.0, n_betas)
    S = n_chains * n_betas
    assert S % 2 == 0
    R = S // 2  # NB with n_betas odd a ladder straddles the two "ranks"
    keys = 0x55E20000 + np.arange(S, dtype=np.uint64)
    betas_global = np.ascontiguousarray(np.tile(betas, n_chains))
    pt_key = 0xC0FFEE
    ranks = []
    for rk in range(2):
        sl = slice(rk * R, (rk + 1) * R)
        g = QmcIsingGraph(edges, gamma, h, 16, keys[sl], betas_global[sl], mode=mode)
        check(g._L.qmcb_pt_configure(g._h, n_chains, n_betas, rk * R, ptr(betas_global, C.c_double), ptr(keys, C.c_uint64), pt_key))
        ranks.append(g)
    L = ranks[0]._L
    words = C.c_uint32()
    check(L.qmcb_pt_record_words(ranks[0]._h, C.byref(words)))
    rec = torch.zeros((S, words.value), dtype=torch.int64, device="cuda:0")
    ladder = _oracle_ladders(edges, gamma, h, 16, keys, n_chains, n_betas)
    cursors, swaps_ref = [0] * n_chains, 0
    for step in range(10):
        for g in ranks:
            g.timesteps(2)
        for c in range(n_chains):
            for k in range(n_betas):
                ladder[c][k].timesteps(2, float(betas[k]), mode)
        # "all-gather": rank rk's records land at rows [rk*R, (rk+1)*R) of the gathered buffer
        for rk, g in enumerate(ranks):
            check(L.qmcb_pt_export(g._h, C.c_void_p(rec.data_ptr() + rk * R * words.value * 8)))
            check(L.qmcb_synchronize(g._h))
        for g in ranks:
            check(L.qmcb_pt_apply(g._h, C.c_void_p(rec.data_ptr()), S))
            check(L.qmcb_synchronize(g._h))
        for c in range(n_chains):
            s, cursors[c] = po.pt_step(ladder[c], betas, pt_key + c, cursors[c])
            swaps_ref += s
        seen = []
        for g in ranks:
            slots = np.zeros(R, dtype=np.uint32)
            check(L.qmcb_pt_get_slots(g._h, ptr(slots, C.c_uint32)))
            n, cut, cur, st, bt = g.get_n(), g.get_cutoff(), g.rng_cursors(), g.state_ref(), g.betas()
            for s_local, slot in enumerate(slots):
                ref = ladder[slot // n_betas][slot % n_betas]
                assert bt[s_local] == betas[slot % n_betas]
                assert int(n[s_local]) == ref.n and int(cut[s_local]) == ref.cutoff and int(cur[s_local]) == ref.cursor, (step, slot)
                assert np.array_equal(st[s_local], ref.state()) and np.array_equal(g.dump_ops(s_local), ref.dump_ops())
            seen += list(slots)
            sw = C.c_uint64()
            check(L.qmcb_pt_total_swaps(g._h, C.byref(sw)))
            assert sw.value == swaps_ref  # every rank evaluates every swap
        assert sorted(seen) == list(range(S))
    assert swaps_ref > 0
    for g in ranks:
        assert g.verify()
        # a spread ladder has no local step and, without a communicator, no library-side step either
        assert L.qmcb_pt_step_local(g._h) != 0 and L.qmcb_pt_step(g._h) != 0
        b = C.c_uint64()
        check(L.qmcb_pt_collective_bytes(g._h, C.byref(b)))
        assert b.value == S * words.value * 8
        g.close()


@pytest.mark.parametrize("mode", [MODE_STRICT, MODE_FAST])
def test_pt_timesteps_sample_matches_reference_loop(mode):
    from isingmontecarlo_b200.tempering import TemperingContainer

    edges = lattices.two_d_periodic_mixed(4)
    n_chains, n_betas = 2, 4
    betas = np.linspace(0.5, 2.0, n_betas)
    S = n_chains * n_betas
    keys = 0x55E30000 + np.arange(S, dtype=np.uint64)
    pt_key = 0xBEEF
    tc = TemperingContainer(edges, 1.0, 0.0, 16, betas, n_chains=n_chains, rng_keys=keys, pt_key=pt_key, mode=mode)
    tc.timesteps(5)
    ladder = _oracle_ladders(edges, 1.0, 0.0, 16, keys, n_chains, n_betas)
    for row in ladder:
        for k, g in enumerate(row):
            g.timesteps(5, float(betas[k]), mode)
    timesteps, swap_freq, sample_freq = 14, 3, 4
    states, energy = tc.timesteps_sample(timesteps, swap_freq, sample_freq)
    # tempering_container.rs:166-208 over the oracle graphs (one container per ladder)
    e_ref = np.zeros(S)
    st_ref = [[] for _ in range(S)]
    cursors = [0] * n_chains
    remaining, to_swap, to_sample = timesteps, swap_freq, sample_freq
    while remaining > 0:
        t = min(to_sample, to_swap, remaining)
        for c in range(n_chains):
            for k in range(n_betas):
                e_ref[c * n_betas + k] += ladder[c][k].timesteps(t, float(betas[k]), mode) * t
        to_sample -= t
        to_swap -= t
        remaining -= t
        if to_swap == 0:
            for c in range(n_chains):
                _, cursors[c] = po.pt_step(ladder[c], betas, pt_key + c, cursors[c])
            to_swap = swap_freq
        if to_sample == 0:
            for c in range(n_chains):
                for k in range(n_betas):
                    st_ref[c * n_betas + k].append(ladder[c][k].state().astype(bool))
            to_sample = sample_freq
    assert np.array_equal(energy, e_ref)
    for s in range(S):
        assert len(states[s]) == timesteps // sample_freq == len(st_ref[s])
        for a, b in zip(states[s], st_ref[s]):
            assert np.array_equal(a, b)
