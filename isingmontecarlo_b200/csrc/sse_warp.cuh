// sse_warp.cuh -- helpers shared by the warp-per-replica sweep kernels (sse_fast.cu, sse_counter.cu): the lock-free
// min-root union-find over global memory, the asynchronous line prefetch, lattice helpers specialised on HAS_H.
#pragma once
#include "sse.cuh"

#define FULL 0xFFFFFFFFu

__device__ __forceinline__ uint32_t ld_cg(const uint32_t *p) { return __ldcg(p); }
__device__ __forceinline__ void st_cg(uint32_t *p, uint32_t v) { __stcg(p, v); }

// L2 eviction policies (createpolicy, carried in the load/store descriptor): the operator string and the per-slot
// records stream through once per pass (evict first); the union-find parents are revisited (evict last)
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint32_t ld_cg_pol(const uint32_t *p, uint64_t pol) {
    uint32_t v;
    asm volatile("ld.global.cg.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol) : "memory");
    return v;
}
__device__ __forceinline__ void st_cg_pol(uint32_t *p, uint32_t v, uint64_t pol) {
    asm volatile("st.global.cg.L2::cache_hint.u32 [%0], %1, %2;" ::"l"(p), "r"(v), "l"(pol) : "memory");
}

// lock-free union-find over global memory; roots are minima, parent[x] <= x always
__device__ __forceinline__ uint32_t uf_find_cg(uint32_t *P, uint32_t x) {
    uint32_t p = ld_cg(P + x);
    while (p != x) {
        uint32_t gp = ld_cg(P + p);
        if (gp == p) return p;
        st_cg(P + x, gp);  // path halving; a stale value is still an ancestor
        x = gp;
        p = ld_cg(P + x);
    }
    return x;
}
// The kernel is bound by the latency of dependent loads, so both chains are climbed in lockstep: the two parent
// loads of a round are in flight together (one L2 round trip per level instead of two).  The start nodes are
// pointed at the root they reached (compression of the nodes that are looked up again: S.rep entries).
template <bool MAXROOT>
__device__ __forceinline__ uint32_t uf_union_dir(uint32_t *P, uint32_t a, uint32_t b) {
    for (;;) {
        const uint32_t a0 = a, b0 = b;
        uint32_t pa = ld_cg(P + a), pb = ld_cg(P + b);
        const uint32_t pa0 = pa, pb0 = pb;
        while (pa != a || pb != b) {
            a = pa, b = pb;
            pa = ld_cg(P + a), pb = ld_cg(P + b);
        }
        if (pa0 != a) st_cg(P + a0, a);  // a stale value is still an ancestor
        if (pb0 != b) st_cg(P + b0, b);
        if (a == b) return a;
        if (MAXROOT ? a < b : a > b) {  // b is hooked under a
            uint32_t t = a;
            a = b, b = t;
        }
        if (atomicCAS(P + b, b, a) == b) return a;
    }
}
// roots are minima (FAST contract: a cluster is keyed by its smallest segment id)
__device__ __forceinline__ uint32_t uf_union_cg(uint32_t *P, uint32_t a, uint32_t b) { return uf_union_dir<false>(P, a, b); }
// roots are maxima: measured and dropped (profiles/README.md, round 2): a maximum changes whenever a newer segment joins
// the cluster, so cached roots go stale and the climbs get longer (2.5 parent loads per step instead of 0.6)
__device__ __forceinline__ uint32_t uf_union_max(uint32_t *P, uint32_t a, uint32_t b) { return uf_union_dir<true>(P, a, b); }
// min-root union with an L2 eviction policy on the parent loads / stores
__device__ __forceinline__ uint32_t uf_union_pol(uint32_t *P, uint32_t a, uint32_t b, uint64_t pol) {
    for (;;) {
        const uint32_t a0 = a, b0 = b;
        uint32_t pa = ld_cg_pol(P + a, pol), pb = ld_cg_pol(P + b, pol);
        const uint32_t pa0 = pa, pb0 = pb;
        while (pa != a || pb != b) {
            a = pa, b = pb;
            pa = ld_cg_pol(P + a, pol), pb = ld_cg_pol(P + b, pol);
        }
        if (pa0 != a) st_cg_pol(P + a0, a, pol);
        if (pb0 != b) st_cg_pol(P + b0, b, pol);
        if (a == b) return a;
        if (a > b) {  // b is hooked under a: roots are minima
            uint32_t t = a;
            a = b, b = t;
        }
        if (atomicCAS(P + b, b, a) == b) return a;
    }
}

__device__ __forceinline__ uint32_t warp_excl_scan(uint32_t v, int lane) {
    uint32_t x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t y = __shfl_up_sync(FULL, x, d);
        if (lane >= d) x += y;
    }
    return x - v;
}

// Software prefetch of the next 128-byte line of the operator string WITHOUT a register: held in a register across a
// whole step, the value was spilled right after the load was issued (register cap), and the spill store waited for the
// DRAM round trip at the top of every step.  Lanes 0..7 copy 16 bytes each straight into shared memory (cp.async.cg:
// through L2 like ld.cg); the line is picked up at the top of the next step.  Rows are 128-byte aligned (cap % 32 == 0).
__device__ __forceinline__ void fetch_line(uint32_t *line_smem, const uint32_t *src, int lane) {
    if (lane < 8) {
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(line_smem + 4 * lane);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src + 4 * lane) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __forceinline__ void fetch_line_pol(uint32_t *line_smem, const uint32_t *src, int lane, uint64_t pol) {
    if (lane < 8) {
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(line_smem + 4 * lane);
        asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(dst), "l"(src + 4 * lane), "l"(pol) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __forceinline__ uint32_t take_line(const uint32_t *line_smem, int lane) {
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    const uint32_t w = line_smem[lane];
    __syncwarp();  // everyone has read the line before the next copy may land in it
    return w;
}

// ---- bulk-asynchronous staging of the operator string (TMA bulk copy, no tensor map: the rows are contiguous) -----------
// One lane issues cp.async.bulk.shared::cluster.global for a tile of several 128-byte lines; completion is signalled on an
// mbarrier in shared memory (complete_tx), the consumers wait on its phase parity.
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void bulk_fetch(void *dst_smem, const void *src, uint32_t bytes, uint64_t *bar, uint64_t pol) {
    const uint32_t b = (uint32_t)__cvta_generic_to_shared(bar), d = (uint32_t)__cvta_generic_to_shared(dst_smem);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy reads of the buffer are done before the async proxy overwrites it
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(d), "l"(src),
                 "r"(bytes), "r"(b), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    const uint32_t b = (uint32_t)__cvta_generic_to_shared(bar);
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "MBAR_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra MBAR_DONE;\n"
        "bra MBAR_WAIT;\n"
        "MBAR_DONE:\n"
        "}\n" ::"r"(b),
        "r"(parity)
        : "memory");
}

// lattice helpers specialised on HAS_H: without a longitudinal field there are no KIND_LONG bonds
template <bool HAS_H>
__device__ __forceinline__ int bkind(const SseDev &D, uint32_t b) {
    return b < D.E ? KIND_BOND : ((!HAS_H || b < D.E + D.N) ? KIND_SITE : KIND_LONG);
}
template <bool HAS_H>
__device__ __forceinline__ double bweight(const Ham &Hm, uint32_t b, int kind, uint32_t s0, uint32_t s1) {
    if (Hm.gw2) return kind == KIND_BOND ? __ldg(Hm.gw2 + 4 * (size_t)b + (s0 | (s1 << 1))) : __ldg(Hm.ggam + b);  // generic interactions
    if (kind == KIND_BOND) {
        const double j = __ldg(Hm.J + b);
        return fabs(j) + (s0 == s1 ? -j : j);
    }
    if (!HAS_H || kind == KIND_SITE) return Hm.gamma;
    return fabs(Hm.h) + (s0 ? Hm.h : -Hm.h);
}

