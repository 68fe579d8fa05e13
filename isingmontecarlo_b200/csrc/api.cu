// api.cu -- the C ABI of include/qmcb.h: handle management, host<->device staging, launches.
// No CPU fallback: every entry point needs a CUDA device and fails with QMCB_ERR_CUDA otherwise.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <dlfcn.h>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/qmcb.h"
#include "classical.cuh"
#include "pt.cuh"
#include "sse.cuh"

// kernels (sse_serial.cu, sse_fast.cu, classical.cu, pt.cu)
int launch_sse_serial(const SseDev &D, int mode, uint64_t target, uint32_t phases, uint64_t sample_freq,
                      uint64_t sample_origin, uint8_t *samples, uint64_t samples_per_rep, int layout, cudaStream_t st);
void launch_sse_verify(const SseDev &D, uint32_t r, int *ok_dev, uint32_t *scratch_dev, cudaStream_t st);
void launch_sse_bond_counts(const SseDev &D, uint32_t r, unsigned long long *counts_dev, cudaStream_t st);
void launch_sse_recount(const SseDev &D, uint32_t r, cudaStream_t st);
void launch_sse_itime_magnetization(const SseDev &D, long long *sums_dev, cudaStream_t st);
void launch_sse_itime_state(const SseDev &D, uint32_t r, uint64_t p_at, uint32_t *out_dev, cudaStream_t st);
void launch_autocorrelation(const uint8_t *samples, uint32_t R, uint32_t N, uint32_t T, uint32_t *bits, uint32_t *ones, double *out,
                            cudaStream_t st);
void launch_spin_products(const uint8_t *samples, uint32_t R, uint32_t N, uint32_t T, uint32_t K, const uint32_t *offsets, const uint32_t *vars,
                          uint8_t *derived, cudaStream_t st);
void launch_sse_init_state(const SseDev &D, cudaStream_t st);
int launch_sse_fast(const SseDev &D, const SseTuning &T, uint64_t target, uint32_t phases, uint64_t sample_freq, uint64_t sample_origin,
                    uint8_t *samples, uint64_t samples_per_rep, cudaStream_t st);  // returns #launches, <0 unsupported
int launch_sse_counter(const SseDev &D, const SseTuning &T, uint64_t target, uint32_t phases, uint64_t sample_freq, uint64_t sample_origin,
                       uint8_t *samples, uint64_t samples_per_rep, cudaStream_t st);  // returns #launches, <0 unsupported
int launch_sse_rvb(const SseDev &D, const RvbDev &W, uint64_t target, long long updates, unsigned long long *succ_out, cudaStream_t st);
int launch_pt_export(const SseDev &D, const PtDev &P, uint64_t *rec, cudaStream_t st);
void launch_pt_apply(const SseDev &D, const PtDev &P, const uint64_t *rec, uint32_t S, cudaStream_t st);
void launch_cls_generic(const ClsDev &D, uint32_t colour, uint32_t cstart, uint32_t ccount, uint64_t sweep, cudaStream_t st);
void launch_cls_square(const ClsDev &D, uint32_t colour, uint64_t sweep, cudaStream_t st);
int launch_cls_square_fused(const ClsDev &D, uint64_t sweep0, uint32_t nsweeps, unsigned int *bar, unsigned int *bar_count, int nsm,
                            cudaStream_t st);  // returns #launches, <0 if the blocks of a replica cannot be co-resident
void launch_cls_square_measure(const ClsDev &D, unsigned long long *unsat, unsigned long long *up, cudaStream_t st);
void launch_cls_square_pack(const ClsDev &D, const uint8_t *bytes, cudaStream_t st);
void launch_cls_square_unpack(const ClsDev &D, uint8_t *bytes, cudaStream_t st);
void launch_cls_init_bytes(const ClsDev &D, uint8_t *bytes, cudaStream_t st);
void launch_cls_generic_energy(const ClsDev &D, const double *adj_j, const double *biases, double *energy, double *mag, cudaStream_t st);
void launch_cls_ref_moves(const ClsRefDev &D, int move, uint64_t nspin, uint64_t nedge, uint64_t nworm, int only_basic,
                          int allow_doubles, uint8_t *choice_out, cudaStream_t st);

static thread_local std::string g_err;

static int fail(int code, const std::string &msg) {
    g_err = msg;
    return code;
}
static int fail_cuda(cudaError_t e, const char *expr, const char *file, int line) {
    char buf[512];
    snprintf(buf, sizeof buf, "CUDA error %s (%s) at %s:%d: %s", cudaGetErrorName(e), cudaGetErrorString(e), file, line, expr);
    g_err = buf;
    cudaGetLastError();
    return QMCB_ERR_CUDA;
}

extern "C" const char *qmcb_last_error(void) { return g_err.c_str(); }
extern "C" const char *qmcb_version(void) { return "isingmontecarlo_b200 0.1 (sm_100a)"; }

struct Pool {  // owns device allocations of one handle
    std::vector<void *> ptrs;
    template <typename T>
    cudaError_t alloc(T **out, size_t count) {
        void *p = nullptr;
        cudaError_t e = cudaMalloc(&p, std::max<size_t>(count, 1) * sizeof(T));
        if (e == cudaSuccess) ptrs.push_back(p);
        *out = (T *)p;
        return e;
    }
    void release(void *p) {
        if (!p) return;
        auto it = std::find(ptrs.begin(), ptrs.end(), p);
        if (it != ptrs.end()) ptrs.erase(it);
        cudaFree(p);
    }
    void release_all() {
        for (void *p : ptrs) cudaFree(p);
        ptrs.clear();
    }
};

// =============================================================================================
// SSE handle
// =============================================================================================
struct QmcbHandle {
    int device = 0;
    cudaStream_t stream = nullptr, own_stream = nullptr;
    SseDev D{};
    Pool pool;
    int mode = QMCB_MODE_STRICT;
    int impl = 0;  // 0 auto (warp-parallel FAST kernels where supported), 1 serial kernels only
    int strict_layout = 9;  // STRICT cluster step: 1 world-line arrays | 8 other leg of a bond op requested in L1 (default), 0 one record per slot (qmcb_set_option "strict_layout")
    bool strict_wl_last = false;  // which layout the last STRICT cluster step left in the workspace (qmcb_get_boundaries)
    double offset = 0.0;
    uint64_t target = 0;  // sweeps requested so far
    uint64_t launches = 0;
    bool strict_ws = false, fast_ws = false, counter_ws = false;
    bool auto_capacity = false;
    double beta_max = 0.0;
    // tempering
    bool pt_on = false;
    PtDev P{};
    uint32_t pt_S = 0;
    uint8_t *samples_dev = nullptr;  // reused between qmcb_timesteps calls
    size_t samples_cap = 0;
    // host copy of the lattice (heat-bath tables, checkpoints)
    std::vector<uint32_t> va_h, vb_h;
    std::vector<double> J_h;
    double *hb_cum_dev = nullptr, *hb_maxw_dev = nullptr, *hb_total_dev = nullptr;
    // per-replica Hamiltonians (qmcb_set_hamiltonians): row 0 of the tables is the lattice given at creation
    uint32_t H = 1;
    std::vector<double> Jtab_h, gam_h, hl_h, offset_h;  // [H][E], [H], [H], [H]
    std::vector<uint32_t> ham_slot_h;                   // [S] tempering: Hamiltonian row of each slot
    uint64_t *pt_rec_dev = nullptr;                     // record buffer of qmcb_pt_step_local
    // autocorrelation work buffers (reused between calls)
    uint32_t *ac_bits = nullptr, *ac_ones = nullptr, *ac_lists = nullptr;
    double *ac_out = nullptr;
    uint8_t *ac_derived = nullptr;
    size_t ac_bits_cap = 0, ac_ones_cap = 0, ac_out_cap = 0, ac_derived_cap = 0, ac_lists_cap = 0;
    // multi-GPU tempering: NCCL communicator (qmcb_pt_comm_init / qmcb_pt_comm_attach)
    void *comm = nullptr;
    bool own_comm = false;
    int comm_rank = 0, comm_nranks = 1;
    uint64_t *pt_rec_all_dev = nullptr;  // [S] gathered records
    double *pt_energy_dev = nullptr;     // [S] per-segment energies of qmcb_pt_timesteps_sample
    bool generic = false;  // created by qmcb_create_qmc: weights from interaction tables
    bool offdiag2 = false; // ... some two-variable interaction has off-diagonal elements (only loop updates produce such ops)
    std::vector<double> gw2_h, ggam_h;
    // kernel-selection knobs of the warp-parallel sweep (qmcb_set_option), per handle
    SseTuning tune{};
    // RVB update (qmc_ising.rs:39-43: run_rvb_steps, classical_bonds, the two success counters)
    bool rvb_on = false, rvb_ws = false;
    RvbDev W{};
    unsigned long long *rvb_succ_last = nullptr;  // [R] successes of the last qmcb_single_rvb_sweep
};

static void pt_comm_release(QmcbHandle *h);
#define CHECK_H(h)                                              \
    if (!(h)) return fail(QMCB_ERR_BAD_ARG, "null handle");     \
    CUDA_TRY(cudaSetDevice((h)->device))

static size_t bits_stride(const SseDev &D) { return (size_t)(D.cap / 32 + 2 + D.N / 32); }

static int alloc_strict_ws(QmcbHandle *h) {
    if (h->strict_ws) return QMCB_OK;
    SseDev &D = h->D;
    CUDA_TRY(h->pool.alloc(&D.rec, (size_t)D.R * strict_rec_stride(D)));
    CUDA_TRY(h->pool.alloc(&D.ent, (size_t)D.R * D.cap));
    CUDA_TRY(h->pool.alloc(&D.frontier, (size_t)D.R * (2 * D.cap + 16)));
    CUDA_TRY(h->pool.alloc(&D.interior, (size_t)D.R * (4 * D.cap + 16)));
    h->strict_ws = true;
    return QMCB_OK;
}
static int alloc_fast_ws(QmcbHandle *h) {
    if (h->fast_ws) return QMCB_OK;
    SseDev &D = h->D;
    CUDA_TRY(h->pool.alloc(&D.parent, (size_t)D.R * (D.N + D.cap + 1)));
    h->fast_ws = true;
    return QMCB_OK;
}

static int alloc_counter_ws(QmcbHandle *h) {
    if (h->counter_ws) return QMCB_OK;
    SseDev &D = h->D;
    CUDA_TRY(h->pool.alloc(&D.sid, (size_t)D.R * D.cap));
    h->counter_ws = true;
    return QMCB_OK;
}

// RVB workspace: scratch only (every launch starts from an empty workspace), sized by the capacity; the lattice's
// classical_bonds table and the two counters are made once
static int alloc_rvb_ws(QmcbHandle *h) {
    if (h->rvb_ws) return QMCB_OK;
    SseDev &D = h->D;
    RvbDev &W = h->W;
    if (!W.vb_start) {
        std::vector<uint32_t> start(D.N + 1, 0), list(2 * (size_t)D.E + 1), fill(D.N, 0);
        for (uint32_t b = 0; b < D.E; b++) start[h->va_h[b] + 1]++, start[h->vb_h[b] + 1]++;
        for (uint32_t v = 0; v < D.N; v++) start[v + 1] += start[v];
        for (uint32_t b = 0; b < D.E; b++) {  // bond order: edge_lookup[a].push(bond); edge_lookup[b].push(bond)
            list[start[h->va_h[b]] + fill[h->va_h[b]]++] = b;
            list[start[h->vb_h[b]] + fill[h->vb_h[b]]++] = b;
        }
        uint32_t *sd = nullptr, *ld = nullptr;
        cudaError_t e0 = h->pool.alloc(&sd, start.size());
        if (e0 == cudaSuccess) e0 = h->pool.alloc(&ld, list.size());
        if (e0 == cudaSuccess) e0 = cudaMemcpy(sd, start.data(), sizeof(uint32_t) * start.size(), cudaMemcpyHostToDevice);
        if (e0 == cudaSuccess) e0 = cudaMemcpy(ld, list.data(), sizeof(uint32_t) * list.size(), cudaMemcpyHostToDevice);
        if (e0 != cudaSuccess) {
            h->pool.release(sd), h->pool.release(ld);
            return fail_cuda(e0, "RVB bond table", __FILE__, __LINE__);
        }
        W.vb_start = sd, W.vb_list = ld;
    }
    if (!h->rvb_succ_last) {  // the counters: all three or none
        unsigned long long *a = nullptr, *b = nullptr, *c = nullptr;
        cudaError_t e0 = h->pool.alloc(&a, D.R);
        if (e0 == cudaSuccess) e0 = h->pool.alloc(&b, D.R);
        if (e0 == cudaSuccess) e0 = h->pool.alloc(&c, D.R);
        if (e0 == cudaSuccess) e0 = cudaMemset(a, 0, sizeof(unsigned long long) * D.R);
        if (e0 == cudaSuccess) e0 = cudaMemset(b, 0, sizeof(unsigned long long) * D.R);
        if (e0 == cudaSuccess) e0 = cudaMemset(c, 0, sizeof(unsigned long long) * D.R);
        if (e0 != cudaSuccess) {
            h->pool.release(a), h->pool.release(b), h->pool.release(c);
            return fail_cuda(e0, "RVB counters", __FILE__, __LINE__);
        }
        W.succ = a, W.count = b, h->rvb_succ_last = c;
    }
    W.stride32 = rvb_stride32(D), W.stride64 = rvb_stride64(D), W.stride8 = rvb_stride8(D);
    cudaError_t e = h->pool.alloc(&W.u32, (size_t)D.R * W.stride32);
    if (e == cudaSuccess) e = h->pool.alloc(&W.f64, (size_t)D.R * W.stride64);
    if (e == cudaSuccess) e = h->pool.alloc(&W.u8, (size_t)D.R * W.stride8);
    if (e != cudaSuccess) {
        cudaGetLastError();
        h->pool.release(W.u32), h->pool.release(W.f64), h->pool.release(W.u8);
        W.u32 = nullptr, W.f64 = nullptr, W.u8 = nullptr;
        return fail(QMCB_ERR_CAPACITY, "out of device memory for the RVB workspace");
    }
    h->rvb_ws = true;
    return QMCB_OK;
}
static void release_rvb_ws(QmcbHandle *h) {  // the capacity changed: the next RVB launch allocates for the new one
    if (!h->rvb_ws) return;
    RvbDev &W = h->W;
    h->pool.release(W.u32), h->pool.release(W.f64), h->pool.release(W.u8);
    W.u32 = nullptr, W.f64 = nullptr, W.u8 = nullptr;
    h->rvb_ws = false;
}

// re-layout every per-slot array for a larger per-replica capacity
static int grow(QmcbHandle *h, uint64_t newcap) {
    SseDev &D = h->D;
    newcap = (newcap + 31) / 32 * 32;  // rows stay 128-byte aligned (the sweep kernels copy whole lines)
    if (newcap <= D.cap) return QMCB_OK;
    if (newcap >= (1ull << 29)) return fail(QMCB_ERR_CAPACITY, "operator string capacity above 2^29 slots per replica");
    // every new buffer is allocated before the old layout is touched: a failed growth (it typically happens near the
    // end of device memory) leaves the handle exactly as it was
    const bool s = h->strict_ws, f = h->fast_ws, c = h->counter_ws;
    SseDev Dn = D;
    Dn.cap = newcap;
    uint32_t *nops = nullptr, *nbits = nullptr, *nfrozen = nullptr, *nrec = nullptr, *nfrontier = nullptr, *ninterior = nullptr, *nparent = nullptr, *nsid = nullptr;
    uint32_t *nent = nullptr;
    cudaError_t e = h->pool.alloc(&nops, (size_t)D.R * newcap);
    if (e == cudaSuccess) e = h->pool.alloc(&nbits, (size_t)D.R * bits_stride(Dn));
    if (e == cudaSuccess) e = h->pool.alloc(&nfrozen, (size_t)D.R * bits_stride(Dn));
    if (e == cudaSuccess && s) e = h->pool.alloc(&nrec, (size_t)D.R * strict_rec_stride(Dn));
    if (e == cudaSuccess && s) e = h->pool.alloc(&nent, (size_t)D.R * newcap);
    if (e == cudaSuccess && s) e = h->pool.alloc(&nfrontier, (size_t)D.R * (2 * newcap + 16));
    if (e == cudaSuccess && s) e = h->pool.alloc(&ninterior, (size_t)D.R * (4 * newcap + 16));
    if (e == cudaSuccess && f) e = h->pool.alloc(&nparent, (size_t)D.R * (D.N + newcap + 1));
    if (e == cudaSuccess && c) e = h->pool.alloc(&nsid, (size_t)D.R * newcap);
    if (e == cudaSuccess) e = cudaMemsetAsync(nops, 0xFF, (size_t)D.R * newcap * 4, h->stream);
    if (e == cudaSuccess) e = cudaMemcpy2DAsync(nops, newcap * 4, D.ops, D.cap * 4, D.cap * 4, D.R, cudaMemcpyDeviceToDevice, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    if (e != cudaSuccess) {
        cudaGetLastError();
        h->pool.release(nops), h->pool.release(nbits), h->pool.release(nfrozen), h->pool.release(nrec);
        h->pool.release(nfrontier), h->pool.release(ninterior), h->pool.release(nparent), h->pool.release(nsid), h->pool.release(nent);
        return fail(QMCB_ERR_CAPACITY, "out of device memory while growing the operator strings");
    }
    h->pool.release(D.ops), h->pool.release(D.rec), h->pool.release(D.frontier), h->pool.release(D.interior);
    h->pool.release(D.parent), h->pool.release(D.bits), h->pool.release(D.frozen), h->pool.release(D.sid), h->pool.release(D.ent);
    D.ent = nent;
    D.ops = nops, D.bits = nbits, D.frozen = nfrozen, D.rec = nrec, D.frontier = nfrontier, D.interior = ninterior, D.parent = nparent, D.sid = nsid;
    D.cap = newcap;
    release_rvb_ws(h);
    return QMCB_OK;
}

static int check_status(QmcbHandle *h, int *status_out) {
    int st = 0;
    CUDA_TRY(cudaMemcpyAsync(&st, h->D.status, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    *status_out = st;
    return QMCB_OK;
}

static int status_to_error(int st) {
    if (st & DEV_ERR_INVARIANT) return fail(QMCB_ERR_INTERNAL, "device invariant failed in the cluster update (reference: unreachable!)");
    if (st & DEV_ERR_STACK) return fail(QMCB_ERR_INTERNAL, "cluster stack overflow");
    if (st & DEV_ERR_PROB) return fail(QMCB_ERR_INTERNAL, "acceptance probability outside [0,1] (reference: gen_bool panics)");
    return QMCB_OK;
}

static int launch_sweeps(QmcbHandle *h, uint32_t phases, uint64_t freq, uint64_t origin, uint8_t *samples_dev, uint64_t spr) {
    int rc;
    if (h->rvb_on && (phases & 0xFu) == 0xFu) {
        // QmcIsingGraph::timestep with run_rvb_steps (qmc_ising.rs:644-795): diagonal update, RVB update (:705-752), cluster
        // update, in three launches per sweep; every launch takes only the replicas whose next sweep is `tgt`
        if ((rc = alloc_rvb_ws(h))) return rc;
        if (h->mode == QMCB_MODE_STRICT && (rc = alloc_strict_ws(h))) return rc;
        if ((h->mode != QMCB_MODE_STRICT || h->impl != 1) && (rc = alloc_fast_ws(h))) return rc;
        if (h->mode == QMCB_MODE_COUNTER && (rc = alloc_counter_ws(h))) return rc;
        auto part = [&](uint64_t tgt, uint32_t ph, uint8_t *smp, uint64_t sp) {
            int nl = -1;
            if (h->impl != 1) {
                if (h->mode == QMCB_MODE_COUNTER) nl = launch_sse_counter(h->D, h->tune, tgt, ph, freq, origin, smp, sp, h->stream);
                else if (h->mode == QMCB_MODE_FAST || ph == (1u | 16u)) nl = launch_sse_fast(h->D, h->tune, tgt, ph, freq, origin, smp, sp, h->stream);
            }
            if (nl < 0) {
                const int wl = launch_sse_serial(h->D, h->mode == QMCB_MODE_STRICT ? 0 : (h->mode == QMCB_MODE_FAST ? 1 : 2), tgt, ph, freq, origin, smp, sp,
                                                 h->mode == QMCB_MODE_STRICT && h->impl != 1 ? h->strict_layout : 0, h->stream);
                if (ph & 2u) h->strict_wl_last = wl != 0;
                nl = wl == 2 ? 2 : 1;
            }
            h->launches += (uint64_t)nl;
        };
        for (uint64_t tgt = origin + 1; tgt <= h->target; tgt++) {
            part(tgt, 1u | 16u, nullptr, 0);
            h->launches += (uint64_t)launch_sse_rvb(h->D, h->W, tgt, -1, nullptr, h->stream);
            part(tgt, 2u | 4u | 8u | 16u, samples_dev, spr);
        }
        CUDA_TRY(cudaGetLastError());
        return QMCB_OK;
    }
    if (h->mode == QMCB_MODE_STRICT) {
        if ((rc = alloc_strict_ws(h))) return rc;
        const bool full = (phases & 0xFu) == 0xFu;
        if (full && h->impl != 1 && !h->D.loop_path && h->target > origin) {
            // the diagonal update is order-exact in the warp-parallel kernel, so STRICT sweeps use it too:
            // per sweep one launch for the diagonal update, one for links + reference-order clusters
            if ((rc = alloc_fast_ws(h))) return rc;
            bool ok = true;
            for (uint64_t tgt = origin + 1; tgt <= h->target && ok; tgt++) {
                int nl = launch_sse_fast(h->D, h->tune, tgt, 1u | 16u, freq, origin, nullptr, 0, h->stream);
                if (nl < 0) { ok = false; break; }
                const int ns = launch_sse_serial(h->D, 0, tgt, 2u | 4u | 8u | 16u, freq, origin, samples_dev, spr, h->strict_layout, h->stream);
                h->strict_wl_last = ns != 0;
                h->launches += (uint64_t)nl + (ns == 2 ? 2 : 1);
            }
            if (ok) {
                CUDA_TRY(cudaGetLastError());
                return QMCB_OK;
            }
        }
        const int wl = launch_sse_serial(h->D, 0, h->target, phases, freq, origin, samples_dev, spr, h->impl == 1 || h->D.loop_path ? 0 : h->strict_layout, h->stream);
        if (phases & 2u) h->strict_wl_last = wl != 0;
        h->launches += 1;
    } else if (h->mode == QMCB_MODE_COUNTER) {
        if ((rc = alloc_fast_ws(h)) || (rc = alloc_counter_ws(h))) return rc;
        int nl = h->impl == 1 ? -1 : launch_sse_counter(h->D, h->tune, h->target, phases, freq, origin, samples_dev, spr, h->stream);
        if (nl < 0) {
            launch_sse_serial(h->D, 2, h->target, phases, freq, origin, samples_dev, spr, 0, h->stream);
            nl = 1;
        }
        h->launches += (uint64_t)nl;
    } else {
        if ((rc = alloc_fast_ws(h))) return rc;
        int nl = h->impl == 1 ? -1 : launch_sse_fast(h->D, h->tune, h->target, phases, freq, origin, samples_dev, spr, h->stream);
        if (nl < 0) {
            launch_sse_serial(h->D, 1, h->target, phases, freq, origin, samples_dev, spr, 0, h->stream);
            nl = 1;
        }
        h->launches += (uint64_t)nl;
    }
    CUDA_TRY(cudaGetLastError());
    return QMCB_OK;
}

// run until every replica reached h->target, growing the arrays when a cutoff outgrows them
static int run_to_target(QmcbHandle *h, uint32_t phases, uint64_t freq, uint64_t origin, uint8_t *samples_dev, uint64_t spr) {
    for (int attempt = 0; attempt < 64; attempt++) {
        int rc = launch_sweeps(h, phases, freq, origin, samples_dev, spr);
        if (rc) return rc;
        int st = 0;
        if ((rc = check_status(h, &st))) return rc;
        if ((rc = status_to_error(st))) return rc;
        if (!(st & DEV_ERR_CAPACITY)) return QMCB_OK;
        if (!h->auto_capacity) return fail(QMCB_ERR_CAPACITY, "cutoff outgrew the capacity given to qmcb_create");
        std::vector<uint32_t> M(h->D.R);
        CUDA_TRY(cudaMemcpy(M.data(), h->D.M, sizeof(uint32_t) * h->D.R, cudaMemcpyDeviceToHost));
        uint64_t need = *std::max_element(M.begin(), M.end());
        CUDA_TRY(cudaMemsetAsync(h->D.status, 0, sizeof(int), h->stream));
        if ((rc = grow(h, need + need / 2 + 1024))) return rc;
    }
    return fail(QMCB_ERR_INTERNAL, "capacity growth did not converge");
}

extern "C" int qmcb_create(const QmcbLattice *lat, uint32_t R, const double *betas, const uint64_t *keys, uint64_t cutoff0,
                           uint64_t capacity, const uint8_t *init_state, int device, QmcbHandle **out) {
    if (!lat || !betas || !keys || !out || R == 0) return fail(QMCB_ERR_BAD_ARG, "null argument or zero replicas");
    if (R > 65535) return fail(QMCB_ERR_UNSUPPORTED, "at most 65535 replicas per handle (the per-replica kernels index them with gridDim.y)");
    if (lat->nvars == 0) return fail(QMCB_ERR_BAD_ARG, "lattice without variables");
    if (lat->nedges && (!lat->va || !lat->vb || !lat->J)) return fail(QMCB_ERR_BAD_ARG, "edge arrays missing");
    if (!(lat->transverse >= 0.0)) return fail(QMCB_ERR_BAD_ARG, "transverse field must be >= 0");
    for (uint32_t e = 0; e < lat->nedges; e++)
        if (lat->va[e] >= lat->nvars || lat->vb[e] >= lat->nvars || lat->va[e] == lat->vb[e])
            return fail(QMCB_ERR_BAD_ARG, "edge endpoint out of range or self-loop");
    if ((uint64_t)lat->nedges + 2ull * lat->nvars >= (1u << 24)) return fail(QMCB_ERR_BAD_ARG, "bond index does not fit 24 bits");
    int ndev = 0;
    CUDA_TRY(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(QMCB_ERR_CUDA, "no such CUDA device");
    CUDA_TRY(cudaSetDevice(device));
    QmcbHandle *h = new QmcbHandle();
    h->device = device;
    cudaDeviceGetAttribute(&h->tune.nsm, cudaDevAttrMultiProcessorCount, device);
    SseDev &D = h->D;
    D.N = lat->nvars, D.E = lat->nedges, D.Nw = (lat->nvars + 31) / 32;
    D.has_h = std::fabs(lat->longitudinal) > DBL_EPSILON;
    D.Nb = D.E + D.N + (D.has_h ? D.N : 0);
    D.gamma = lat->transverse, D.h = lat->longitudinal;
    D.R = R;
    double edge_offset = 0.0;  // qmc_ising.rs:97-99
    for (uint32_t e = 0; e < lat->nedges; e++) edge_offset += std::fabs(lat->J[e]);
    h->offset = edge_offset + (double)lat->nvars * (lat->transverse + std::fabs(lat->longitudinal));
    for (uint32_t r = 0; r < R; r++) h->beta_max = std::max(h->beta_max, betas[r]);
    h->auto_capacity = capacity == 0;
    if (capacity == 0) capacity = std::max<uint64_t>(cutoff0, (uint64_t)(3.2 * h->beta_max * h->offset) + 1024);
    if (capacity < cutoff0) {
        delete h;
        return fail(QMCB_ERR_BAD_ARG, "capacity smaller than the initial cutoff");
    }
    capacity = (capacity + 31) / 32 * 32;
    D.cap = capacity;
#define TRYC(expr)                                   \
    do {                                             \
        cudaError_t e_ = (expr);                     \
        if (e_ != cudaSuccess) {                     \
            h->pool.release_all();                   \
            delete h;                                \
            return fail_cuda(e_, #expr, __FILE__, __LINE__); \
        }                                            \
    } while (0)
    TRYC(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
    h->stream = h->own_stream;
    uint32_t *va, *vb;
    double *J;
    TRYC(h->pool.alloc(&va, D.E));
    TRYC(h->pool.alloc(&vb, D.E));
    TRYC(h->pool.alloc(&J, D.E));
    TRYC(cudaMemcpy(va, lat->va, sizeof(uint32_t) * D.E, cudaMemcpyHostToDevice));
    TRYC(cudaMemcpy(vb, lat->vb, sizeof(uint32_t) * D.E, cudaMemcpyHostToDevice));
    TRYC(cudaMemcpy(J, lat->J, sizeof(double) * D.E, cudaMemcpyHostToDevice));
    D.va = va, D.vb = vb, D.J = J;
    {
        std::vector<uint2> vab(D.E);
        for (uint32_t e = 0; e < D.E; e++) vab[e] = make_uint2(lat->va[e], lat->vb[e]);
        uint2 *vab_dev;
        TRYC(h->pool.alloc(&vab_dev, D.E));
        TRYC(cudaMemcpy(vab_dev, vab.data(), sizeof(uint2) * D.E, cudaMemcpyHostToDevice));
        D.vab = vab_dev;
        D.zone = ((uint64_t)D.Nb << __builtin_clzll((uint64_t)D.Nb)) - 1ull;
        // packed edge table (see sse.cuh): only when it can represent the lattice exactly
        std::vector<double> dict;
        std::vector<uint32_t> epk(D.E);
        bool ok = D.N <= 16384;
        for (uint32_t e = 0; e < D.E && ok; e++) {
            size_t c = 0;
            while (c < dict.size() && memcmp(&dict[c], &lat->J[e], sizeof(double)) != 0) c++;  // bitwise: -0.0 != 0.0
            if (c == dict.size()) {
                if (dict.size() == 16) ok = false;
                else dict.push_back(lat->J[e]);
            }
            epk[e] = lat->va[e] | (lat->vb[e] << 14) | ((uint32_t)c << 28);
        }
        if (ok && D.E) {
            uint32_t *epk_dev;
            TRYC(h->pool.alloc(&epk_dev, D.E));
            TRYC(cudaMemcpy(epk_dev, epk.data(), sizeof(uint32_t) * D.E, cudaMemcpyHostToDevice));
            D.epk = epk_dev;
            for (size_t c = 0; c < dict.size(); c++) D.jdict[c] = dict[c];
        }
    }
    h->va_h.assign(lat->va, lat->va + D.E), h->vb_h.assign(lat->vb, lat->vb + D.E), h->J_h.assign(lat->J, lat->J + D.E);
    h->Jtab_h = h->J_h, h->gam_h = {lat->transverse}, h->hl_h = {lat->longitudinal}, h->offset_h = {h->offset};
    TRYC(h->pool.alloc(&D.ops, (size_t)R * D.cap));
    TRYC(cudaMemset(D.ops, 0xFF, (size_t)R * D.cap * 4));
    TRYC(h->pool.alloc(&D.state, (size_t)R * D.Nw));
    TRYC(cudaMemset(D.state, 0, (size_t)R * D.Nw * 4));
    TRYC(h->pool.alloc(&D.n, R));
    TRYC(cudaMemset(D.n, 0, sizeof(uint32_t) * R));
    TRYC(h->pool.alloc(&D.M, R));
    TRYC(h->pool.alloc(&D.cursor, R));
    TRYC(cudaMemset(D.cursor, 0, sizeof(uint64_t) * R));
    TRYC(h->pool.alloc(&D.key, R));
    TRYC(h->pool.alloc(&D.beta, R));
    TRYC(h->pool.alloc(&D.done, R));
    TRYC(cudaMemset(D.done, 0, sizeof(uint64_t) * R));
    TRYC(h->pool.alloc(&D.sum_n, R));
    TRYC(cudaMemset(D.sum_n, 0, sizeof(uint64_t) * R));
    TRYC(h->pool.alloc(&D.vupd, R));
    TRYC(cudaMemset(D.vupd, 0, sizeof(uint64_t) * R));
    TRYC(h->pool.alloc(&D.ncl, R));
    TRYC(cudaMemset(D.ncl, 0, sizeof(uint32_t) * R));
    TRYC(h->pool.alloc(&D.ends, (size_t)R * 4));
    TRYC(cudaMemset(D.ends, 0xFF, sizeof(uint32_t) * R * 4));
    TRYC(h->pool.alloc(&D.status, 1));
    TRYC(cudaMemset(D.status, 0, sizeof(int)));
    TRYC(h->pool.alloc(&D.vfirst, (size_t)R * D.N));
    TRYC(h->pool.alloc(&D.vlast, (size_t)R * D.N));
    TRYC(h->pool.alloc(&D.cur, (size_t)R * D.N));
    TRYC(h->pool.alloc(&D.bits, (size_t)R * bits_stride(D)));
    TRYC(h->pool.alloc(&D.frozen, (size_t)R * bits_stride(D)));
    std::vector<uint32_t> M(R, (uint32_t)cutoff0);
    TRYC(cudaMemcpy(D.M, M.data(), sizeof(uint32_t) * R, cudaMemcpyHostToDevice));
    TRYC(cudaMemcpy(D.key, keys, sizeof(uint64_t) * R, cudaMemcpyHostToDevice));
    TRYC(cudaMemcpy(D.beta, betas, sizeof(double) * R, cudaMemcpyHostToDevice));
    if (init_state) {
        std::vector<uint32_t> packed((size_t)R * D.Nw, 0u);
        for (uint32_t r = 0; r < R; r++)
            for (uint32_t v = 0; v < D.N; v++)
                if (init_state[(size_t)r * D.N + v]) packed[(size_t)r * D.Nw + (v >> 5)] |= 1u << (v & 31);
        TRYC(cudaMemcpy(D.state, packed.data(), packed.size() * 4, cudaMemcpyHostToDevice));
    } else {
        launch_sse_init_state(D, h->stream);
        h->launches++;
        TRYC(cudaGetLastError());
        TRYC(cudaStreamSynchronize(h->stream));
    }
#undef TRYC
    *out = h;
    return QMCB_OK;
}

// Qmc with generic interactions (qmc_runner.rs:46-156, :406-680; into_qmc qmc_ising.rs:943-976)
extern "C" int qmcb_create_qmc(const QmcbInteractions *I, uint32_t R, const double *betas, const uint64_t *keys, uint64_t cutoff0,
                               uint64_t capacity, const uint8_t *init_state, int device, QmcbHandle **out) {
    if (!I || !out || !I->nv || !I->vars || !I->mat_len || !I->mats || I->nvars == 0) return fail(QMCB_ERR_BAD_ARG, "null argument or no variables");
    const uint32_t N = I->nvars, n = I->n_interactions;
    // shape taken: E two-variable interactions, then either one constant one-variable interaction per variable (cluster
    // edges) or none at all (a model that moves by loop updates only)
    uint32_t E = 0;
    while (E < n && I->nv[E] == 2) E++;
    const bool no_site = E == n;
    if (!no_site && n - E != N) return fail(QMCB_ERR_UNSUPPORTED, "need one constant one-variable interaction per variable (cluster edges) after the two-variable ones, or none");
    bool offdiag2 = false;
    std::vector<double> full(16 * (size_t)E + 4 * (size_t)N, 0.0);
    std::vector<uint32_t> va(E), vb(E);
    std::vector<double> w2(4 * (size_t)E), gam((size_t)n, 0.0), Jz(E, 0.0);
    const double *m = I->mats;
    const double EPS = 2.220446049250313e-16;
    for (uint32_t b = 0; b < n; b++) {
        const uint32_t nv = I->nv[b], len = I->mat_len[b];
        if (nv != 1 && nv != 2) return fail(QMCB_ERR_UNSUPPORTED, "interactions of more than two variables are not offered");
        const bool diagonal = len == (1u << nv);
        if (!diagonal && len != (1u << (2 * nv))) return fail(QMCB_ERR_BAD_ARG, "Matrix size must be power of 2 matching the variables");  // get_mat_var_size :666-680
        if (!diagonal)
            for (uint32_t k = 0; k < len; k++)
                if (m[k] < 0.0) return fail(QMCB_ERR_BAD_ARG, "Interaction contains negative weights");  // Interaction::new :527-529
        for (uint32_t k = 0; k < nv; k++)
            if (I->vars[2 * b + k] >= N) return fail(QMCB_ERR_BAD_ARG, "interaction variable out of range");
        if (b < E) {
            if (nv != 2) return fail(QMCB_ERR_UNSUPPORTED, "interaction order: the two-variable interactions come first, then one constant one-variable interaction per variable");
            if (I->vars[2 * b] == I->vars[2 * b + 1]) return fail(QMCB_ERR_BAD_ARG, "interaction with a repeated variable");
            double d[4];  // diagonal elements by Interaction::at index (in0 << 1 | in1)
            for (uint32_t x = 0; x < 4; x++) d[x] = diagonal ? m[x] : m[(x << 2) + x];
            for (uint32_t x = 0; x < 4; x++)
                if (d[x] < 0.0) return fail(QMCB_ERR_BAD_ARG, "negative diagonal weight (use the *_and_offset constructor)");
            // sym_under_ising :626-648 on the diagonal, and no off-diagonal elements (they only matter for loop updates)
            if (!(std::fabs(d[0] - d[3]) < EPS && std::fabs(d[1] - d[2]) < EPS))
                return fail(QMCB_ERR_UNSUPPORTED, "a two-variable interaction breaks the Ising symmetry: the reference then runs no cluster update (qmc_runner.rs:278-281); not offered");
            if (!diagonal) {  // off-diagonal elements matter to loop updates only; the cluster update flips all four legs, so
                              // the whole matrix has to be symmetric under the global flip (sym_under_ising :626-648)
                for (uint32_t x = 0; x < 16; x++)
                    if (!(std::fabs(m[x] - m[(~x) & 15u]) < EPS))
                        return fail(QMCB_ERR_UNSUPPORTED, "a two-variable interaction breaks the Ising symmetry: the reference then runs no cluster update (qmc_runner.rs:278-281); not offered");
                for (uint32_t o = 0; o < 4; o++)
                    for (uint32_t i = 0; i < 4; i++)
                        if (o != i && m[(o << 2) + i] != 0.0) offdiag2 = true;
            }
            for (uint32_t x = 0; x < 16; x++)  // Interaction::at (:560-600): a Diagonal interaction is 0 off the diagonal
                full[16 * (size_t)b + x] = diagonal ? ((x >> 2) == (x & 3u) ? m[x & 3u] : 0.0) : m[x];
            va[b] = I->vars[2 * b], vb[b] = I->vars[2 * b + 1];
            for (uint32_t s0 = 0; s0 < 2; s0++)
                for (uint32_t s1 = 0; s1 < 2; s1++) w2[4 * (size_t)b + (s0 | (s1 << 1))] = d[(s0 << 1) | s1];
        } else {
            const uint32_t v = b - E;
            if (nv != 1 || I->vars[2 * b] != v)
                return fail(QMCB_ERR_UNSUPPORTED, "interaction order: after the two-variable interactions, exactly one one-variable interaction per variable, in variable order");
            bool constant = !diagonal;  // InteractionType::Full(true): every matrix element equal
            for (uint32_t k = 1; constant && k < len; k++) constant = std::fabs(m[k - 1] - m[k]) < EPS;
            if (!constant) return fail(QMCB_ERR_UNSUPPORTED, "one-variable interactions must be constant (cluster edges, cluster.rs:284-286)");
            gam[b] = m[0];
            for (uint32_t x = 0; x < 4; x++) full[16 * (size_t)E + 4 * (size_t)v + x] = m[x];
        }
        m += len;
    }
    if (no_site) gam.resize((size_t)E + N, 0.0);
    QmcbLattice lat{N, E, va.data(), vb.data(), Jz.data(), gam[E], 0.0};
    QmcbHandle *h = nullptr;
    int rc = qmcb_create(&lat, R, betas, keys, cutoff0, capacity, init_state, device, &h);
    if (rc) return rc;
    SseDev &D = h->D;
    double *w2_dev = nullptr, *gam_dev = nullptr;
    cudaError_t e = h->pool.alloc(&w2_dev, w2.size());
    if (e == cudaSuccess) e = h->pool.alloc(&gam_dev, gam.size());
    if (e == cudaSuccess) e = cudaMemcpy(w2_dev, w2.data(), sizeof(double) * w2.size(), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(gam_dev, gam.data(), sizeof(double) * gam.size(), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        qmcb_destroy(h);
        return fail_cuda(e, "interaction tables", __FILE__, __LINE__);
    }
    double *full_dev = nullptr;
    e = h->pool.alloc(&full_dev, full.size());
    if (e == cudaSuccess) e = cudaMemcpy(full_dev, full.data(), sizeof(double) * full.size(), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        qmcb_destroy(h);
        return fail_cuda(e, "interaction tables", __FILE__, __LINE__);
    }
    D.g_w2 = w2_dev, D.g_gam = gam_dev, D.g_full = full_dev;
    D.loop_updates = I->do_loop_updates ? 1 : 0, D.no_site = no_site ? 1 : 0;
    h->offdiag2 = offdiag2;
    D.loop_path = D.loop_updates || D.no_site || offdiag2;
    if (no_site) {  // bond indices are the two-variable interactions only (qmc_runner.rs:177)
        D.Nb = E;
        D.zone = E ? (((uint64_t)E << __builtin_clzll((unsigned long long)E)) - 1ull) : 0;
        if (E == 0) {
            qmcb_destroy(h);
            return fail(QMCB_ERR_BAD_ARG, "no interactions");
        }
    }
    D.epk = nullptr;  // the packed (J, Gamma) tables of the Ising path do not describe this handle
    h->generic = true, h->gw2_h = w2, h->ggam_h = gam;
    h->offset = I->offset, h->offset_h = {I->offset};
    *out = h;
    return QMCB_OK;
}

extern "C" int qmcb_destroy(QmcbHandle *h) {
    if (!h) return QMCB_OK;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    pt_comm_release(h);
    h->pool.release_all();
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    delete h;
    return QMCB_OK;
}

extern "C" int qmcb_set_stream(QmcbHandle *h, void *s) {
    CHECK_H(h);
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    h->stream = s ? (cudaStream_t)s : h->own_stream;
    return QMCB_OK;
}
extern "C" int qmcb_set_mode(QmcbHandle *h, int mode) {
    CHECK_H(h);
    if (mode != QMCB_MODE_STRICT && mode != QMCB_MODE_FAST && mode != QMCB_MODE_COUNTER) return fail(QMCB_ERR_BAD_ARG, "unknown mode");
    if (mode != QMCB_MODE_STRICT && h->D.loop_path) return fail(QMCB_ERR_UNSUPPORTED, "models with loop updates, off-diagonal two-variable interactions or no cluster edges run in QMCB_MODE_STRICT only");
    if (mode == QMCB_MODE_COUNTER && h->D.hb_cum) return fail(QMCB_ERR_UNSUPPORTED, "the heat-bath diagonal update has no COUNTER-mode contract (use FAST or STRICT)");
    h->mode = mode;
    return QMCB_OK;
}
// QmcIsingGraph::set_enable_heatbath (qmc_ising.rs:444-486): BondWeights from make_bond_weights
// (heatbath.rs:130-146: maximum diagonal weight per bond; BondWeights::new :17-35: running sum)
extern "C" int qmcb_set_enable_heatbath(QmcbHandle *h, int enable) {
    CHECK_H(h);
    SseDev &D = h->D;
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    if (!enable) {
        D.hb_cum = D.hb_maxw = nullptr;
        return QMCB_OK;
    }
    if (h->mode == QMCB_MODE_COUNTER) return fail(QMCB_ERR_UNSUPPORTED, "the heat-bath diagonal update has no COUNTER-mode contract (use FAST or STRICT)");
    if (!h->hb_cum_dev) {
        const uint32_t H = h->H;
        std::vector<double> maxw((size_t)H * D.Nb), cum((size_t)H * D.Nb), tot(H);
        for (uint32_t hi = 0; hi < H; hi++) {
            double *mw = maxw.data() + (size_t)hi * D.Nb, *cm = cum.data() + (size_t)hi * D.Nb;
            const double gam = h->gam_h[hi], hl = h->hl_h[hi];
            for (uint32_t b = 0; b < D.Nb; b++) {
                double acc = 0.0;
                if (h->generic) {  // make_bond_weights (heatbath.rs:130-146) over the interaction's diagonal substates
                    if (b < D.E) {
                        for (int x = 0; x < 4; x++) acc = std::max(acc, h->gw2_h[4 * (size_t)b + x]);
                    } else acc = h->ggam_h[b];
                } else if (b < D.E) {  // two_site_hamiltonian, qmc_ising.rs:863-875, over the four diagonal substates
                    const double j = h->Jtab_h[(size_t)hi * D.E + b], cand[2] = {std::fabs(j) - j, std::fabs(j) + j};
                    for (double w : cand)
                        if (w > acc) acc = w;
                } else if (b < D.E + D.N) {
                    if (gam > acc) acc = gam;
                } else {
                    const double cand[2] = {std::fabs(hl) - hl, std::fabs(hl) + hl};
                    for (double w : cand)
                        if (w > acc) acc = w;
                }
                mw[b] = acc;
                cm[b] = b == 0 ? acc : acc + cm[b - 1];
            }
            tot[hi] = cm[D.Nb - 1];
            if (!(tot[hi] > 0.0)) return fail(QMCB_ERR_BAD_ARG, "heat-bath update needs a non-zero total bond weight");
        }
        CUDA_TRY(h->pool.alloc(&h->hb_cum_dev, (size_t)H * D.Nb));
        CUDA_TRY(h->pool.alloc(&h->hb_maxw_dev, (size_t)H * D.Nb));
        CUDA_TRY(h->pool.alloc(&h->hb_total_dev, H));
        CUDA_TRY(cudaMemcpy(h->hb_cum_dev, cum.data(), sizeof(double) * cum.size(), cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemcpy(h->hb_maxw_dev, maxw.data(), sizeof(double) * maxw.size(), cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemcpy(h->hb_total_dev, tot.data(), sizeof(double) * H, cudaMemcpyHostToDevice));
        D.hb_total = tot[0];
    }
    D.hb_cum_tab = h->hb_cum_dev, D.hb_maxw_tab = h->hb_maxw_dev, D.hb_total_tab = h->hb_total_dev;
    D.hb_cum = h->hb_cum_dev, D.hb_maxw = h->hb_maxw_dev;
    return QMCB_OK;
}
// f64::signum
static double signum(double x) { return x != x ? x : (std::signbit(x) ? -1.0 : 1.0); }

// Replicas of one batch with unequal Hamiltonians: what the reference expresses as graphs built with
// different couplings in one TemperingContainer.  The rows must pass SwapManagers::can_swap_graphs
// (qmc_ising.rs:563-590: same edges -- given, one lattice per handle -- couplings of the same sign per edge,
// longitudinal fields of the same sign); the bond-index space additionally needs |h| > eps in all rows or none.
extern "C" int qmcb_set_hamiltonians(QmcbHandle *h, uint32_t n_ham, const double *J_tab, const double *transverse,
                                     const double *longitudinal, const uint32_t *ham_of_replica) {
    CHECK_H(h);
    SseDev &D = h->D;
    if (h->generic) return fail(QMCB_ERR_UNSUPPORTED, "per-replica Hamiltonians are (J, Gamma, h) rows: not available for a handle with generic interactions");
    if (!n_ham || (D.E && !J_tab) || !transverse || !longitudinal || !ham_of_replica) return fail(QMCB_ERR_BAD_ARG, "null argument");
    for (uint32_t hi = 0; hi < n_ham; hi++) {
        if (!(transverse[hi] >= 0.0)) return fail(QMCB_ERR_BAD_ARG, "transverse field must be >= 0");
        if ((std::fabs(longitudinal[hi]) > DBL_EPSILON) != (D.has_h != 0))
            return fail(QMCB_ERR_BAD_ARG, "longitudinal field must be non-zero in all Hamiltonians of a batch or in none");
        if (signum(longitudinal[hi]) != signum(longitudinal[0]))
            return fail(QMCB_ERR_BAD_ARG, "Longitudinal fields are not of the same sign");  // qmc_ising.rs:581-586
        for (uint32_t e = 0; e < D.E; e++)
            if (signum(J_tab[(size_t)hi * D.E + e]) != signum(J_tab[e]))
                return fail(QMCB_ERR_BAD_ARG, "bonds must be of same sign");  // qmc_ising.rs:572-577
    }
    for (uint32_t r = 0; r < D.R; r++)
        if (ham_of_replica[r] >= n_ham) return fail(QMCB_ERR_BAD_ARG, "Hamiltonian index out of range");
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    const bool hb = D.hb_cum != nullptr;
    h->H = n_ham;
    h->Jtab_h.assign(J_tab, J_tab + (size_t)n_ham * D.E);
    h->gam_h.assign(transverse, transverse + n_ham), h->hl_h.assign(longitudinal, longitudinal + n_ham);
    h->offset_h.resize(n_ham);
    for (uint32_t hi = 0; hi < n_ham; hi++) {  // qmc_ising.rs:97-99
        double edge_offset = 0.0;
        for (uint32_t e = 0; e < D.E; e++) edge_offset += std::fabs(J_tab[(size_t)hi * D.E + e]);
        h->offset_h[hi] = edge_offset + (double)D.N * (transverse[hi] + std::fabs(longitudinal[hi]));
    }
    double *jt, *gt, *ht;
    uint32_t *hr;
    CUDA_TRY(h->pool.alloc(&jt, (size_t)n_ham * D.E));
    CUDA_TRY(h->pool.alloc(&gt, n_ham));
    CUDA_TRY(h->pool.alloc(&ht, n_ham));
    CUDA_TRY(h->pool.alloc(&hr, D.R));
    CUDA_TRY(cudaMemcpy(jt, J_tab, sizeof(double) * (size_t)n_ham * D.E, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(gt, transverse, sizeof(double) * n_ham, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(ht, longitudinal, sizeof(double) * n_ham, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(hr, ham_of_replica, sizeof(uint32_t) * D.R, cudaMemcpyHostToDevice));
    h->pool.release((void *)D.J_tab), h->pool.release((void *)D.gam_tab), h->pool.release((void *)D.h_tab), h->pool.release((void *)D.ham);
    D.J_tab = jt, D.gam_tab = gt, D.h_tab = ht, D.ham = hr;
    D.epk = nullptr;  // couplings differ between replicas: no shared packed table
    // the heat-bath tables are per Hamiltonian: rebuild them
    h->pool.release(h->hb_cum_dev), h->pool.release(h->hb_maxw_dev), h->pool.release(h->hb_total_dev);
    h->hb_cum_dev = h->hb_maxw_dev = h->hb_total_dev = nullptr;
    D.hb_cum = D.hb_maxw = nullptr;
    return hb ? qmcb_set_enable_heatbath(h, 1) : QMCB_OK;
}
extern "C" int qmcb_num_hamiltonians(const QmcbHandle *h, uint32_t *n_ham) {
    if (!h || !n_ham) return fail(QMCB_ERR_BAD_ARG, "null argument");
    *n_ham = h->H;
    return QMCB_OK;
}
// Hamiltonian row of every replica (moves with the slot label under tempering)
extern "C" int qmcb_get_hamiltonian_index(QmcbHandle *h, uint32_t *ham_of_replica) {
    CHECK_H(h);
    if (!ham_of_replica) return fail(QMCB_ERR_BAD_ARG, "null argument");
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    if (!h->D.ham) std::fill(ham_of_replica, ham_of_replica + h->D.R, 0u);
    else CUDA_TRY(cudaMemcpy(ham_of_replica, h->D.ham, sizeof(uint32_t) * h->D.R, cudaMemcpyDeviceToHost));
    return QMCB_OK;
}
// get_offset of every replica's current Hamiltonian (qmc_ising.rs:97-99, :556-559)
extern "C" int qmcb_get_offsets(QmcbHandle *h, double *offsets) {
    CHECK_H(h);
    if (!offsets) return fail(QMCB_ERR_BAD_ARG, "null argument");
    std::vector<uint32_t> hr(h->D.R);
    int rc = qmcb_get_hamiltonian_index(h, hr.data());
    if (rc) return rc;
    for (uint32_t r = 0; r < h->D.R; r++) offsets[r] = h->offset_h[hr[r]];
    return QMCB_OK;
}
extern "C" int qmcb_get_enable_heatbath(const QmcbHandle *h, int *enabled) {
    if (!h || !enabled) return fail(QMCB_ERR_BAD_ARG, "null argument");
    *enabled = h->D.hb_cum != nullptr;
    return QMCB_OK;
}
extern "C" int qmcb_get_mode(const QmcbHandle *h, int *mode) {
    if (!h || !mode) return fail(QMCB_ERR_BAD_ARG, "null argument");
    *mode = h->mode;
    return QMCB_OK;
}
extern "C" int qmcb_set_option(QmcbHandle *h, const char *name, int64_t value) {
    if (!h || !name) return fail(QMCB_ERR_BAD_ARG, "null argument");
    if (!strcmp(name, "strict_layout")) {
        if (value < 0 || value > 15 || (value && !(value & 1))) return fail(QMCB_ERR_BAD_ARG, "strict_layout is 0 (one record per slot) or 1 (world-line arrays) [| 2 next interior entry requested in L1 | 4 links in their own launch | 8 other leg of a bond op requested in L1]");
        h->strict_layout = (int)value;
        return QMCB_OK;
    }
    if (!strcmp(name, "l2_fetch_granularity")) {  // device-wide: bytes the L2 fetches from DRAM per miss (32, 64 or 128)
        CUDA_TRY(cudaSetDevice(h->device));
        CUDA_TRY(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)value));
        return QMCB_OK;
    }
    if (!strcmp(name, "impl")) {
        h->impl = (int)value;
        return QMCB_OK;
    }
    if (!strcmp(name, "smem_pad")) {
        h->tune.pad = (int)std::min<int64_t>(std::max<int64_t>(value, 0), 227 * 1024);
        return QMCB_OK;
    }
    if (!strcmp(name, "smem_carveout")) {
        h->tune.carveout = (int)value;
        return QMCB_OK;
    }
    if (!strcmp(name, "pipeline")) {
        h->tune.pipe = (int)value;
        return QMCB_OK;
    }
    if (!strcmp(name, "shared_edge_table")) {
        h->tune.epk = (int)value;
        return QMCB_OK;
    }
    if (!strcmp(name, "minblocks")) {
        if (value != 0 && value != 4 && value != 6 && value != 7 && value != 8) return fail(QMCB_ERR_BAD_ARG, "minblocks must be 0, 4, 6, 7 or 8");
        h->tune.minblocks = (int)value;
        return QMCB_OK;
    }
    if (!strcmp(name, "debug_counters")) {
        if (value && !h->D.dbg) {
            if (h->pool.alloc(&h->D.dbg, 64) != cudaSuccess) return fail(QMCB_ERR_CUDA, "alloc debug counters");
            cudaMemset(h->D.dbg, 0, 64 * sizeof(unsigned long long));
        }
        return QMCB_OK;
    }
    if (!strcmp(name, "auto_capacity")) {  // grow the strings on demand even if a capacity was given
        h->auto_capacity = value != 0;
        return QMCB_OK;
    }
    return fail(QMCB_ERR_BAD_ARG, "unknown option");
}
extern "C" int qmcb_get_debug_counters(QmcbHandle *h, uint64_t *out16) {
    CHECK_H(h);
    if (!out16 || !h->D.dbg) return fail(QMCB_ERR_BAD_ARG, "debug counters are off");
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    CUDA_TRY(cudaMemcpy(out16, h->D.dbg, 64 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    return QMCB_OK;
}
extern "C" int qmcb_set_betas(QmcbHandle *h, const double *betas) {
    CHECK_H(h);
    if (!betas) return fail(QMCB_ERR_BAD_ARG, "null betas");
    CUDA_TRY(cudaMemcpyAsync(h->D.beta, betas, sizeof(double) * h->D.R, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return QMCB_OK;
}
extern "C" int qmcb_get_betas(const QmcbHandle *h, double *betas) {
    if (!h || !betas) return fail(QMCB_ERR_BAD_ARG, "null argument");
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    CUDA_TRY(cudaMemcpy(betas, h->D.beta, sizeof(double) * h->D.R, cudaMemcpyDeviceToHost));
    return QMCB_OK;
}
extern "C" int qmcb_num_replicas(const QmcbHandle *h, uint32_t *r) {
    if (!h || !r) return fail(QMCB_ERR_BAD_ARG, "null argument");
    *r = h->D.R;
    return QMCB_OK;
}
extern "C" int qmcb_num_vars(const QmcbHandle *h, uint32_t *n) {
    if (!h || !n) return fail(QMCB_ERR_BAD_ARG, "null argument");
    *n = h->D.N;
    return QMCB_OK;
}
extern "C" int qmcb_num_edges(const QmcbHandle *h, uint32_t *ne) {
    if (!h || !ne) return fail(QMCB_ERR_BAD_ARG, "null argument");
    *ne = h->D.E;
    return QMCB_OK;
}
extern "C" int qmcb_get_edges(const QmcbHandle *h, uint32_t *va, uint32_t *vb, double *J) {
    if (!h || !va || !vb || !J) return fail(QMCB_ERR_BAD_ARG, "null argument");
    std::copy(h->va_h.begin(), h->va_h.end(), va);
    std::copy(h->vb_h.begin(), h->vb_h.end(), vb);
    std::copy(h->J_h.begin(), h->J_h.end(), J);
    return QMCB_OK;
}
extern "C" int qmcb_get_fields(const QmcbHandle *h, double *transverse, double *longitudinal) {
    if (!h || !transverse || !longitudinal) return fail(QMCB_ERR_BAD_ARG, "null argument");
    *transverse = h->D.gamma, *longitudinal = h->D.h;
    return QMCB_OK;
}
extern "C" int qmcb_num_bonds(const QmcbHandle *h, uint32_t *nb) {
    if (!h || !nb) return fail(QMCB_ERR_BAD_ARG, "null argument");
    *nb = h->D.Nb;
    return QMCB_OK;
}

static int timesteps_impl(QmcbHandle *h, uint64_t t, uint64_t freq, double *energy_out, uint8_t *samples_out, bool keep_device_samples) {
    if (freq == 0) freq = 1;
    const SseDev &D = h->D;
    const uint64_t spr = t / freq;
    uint8_t *samples_dev = nullptr;
    int rc = QMCB_OK;
    if ((samples_out || keep_device_samples) && spr) {
        const size_t need = (size_t)D.R * spr * D.N;
        if (need > h->samples_cap) {
            if (h->samples_dev) h->pool.release(h->samples_dev);
            h->samples_dev = nullptr, h->samples_cap = 0;
            cudaError_t e = h->pool.alloc(&h->samples_dev, need);
            if (e != cudaSuccess) return fail_cuda(e, "cudaMalloc(samples)", __FILE__, __LINE__);
            h->samples_cap = need;
        }
        samples_dev = h->samples_dev;
    }
    CUDA_TRY(cudaMemsetAsync(D.sum_n, 0, sizeof(uint64_t) * D.R, h->stream));
    const uint64_t origin = h->target;
    h->target += t;
    if (t) rc = run_to_target(h, 1u | 2u | 4u | 8u, freq, origin, samples_dev, spr);
    if (rc == QMCB_OK && energy_out) {
        std::vector<unsigned long long> sn(D.R);
        std::vector<double> beta(D.R);
        std::vector<uint32_t> hr(D.R, 0u);
        cudaError_t e = cudaMemcpyAsync(sn.data(), D.sum_n, sizeof(uint64_t) * D.R, cudaMemcpyDeviceToHost, h->stream);
        if (e == cudaSuccess && D.ham) e = cudaMemcpyAsync(hr.data(), D.ham, sizeof(uint32_t) * D.R, cudaMemcpyDeviceToHost, h->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(beta.data(), D.beta, sizeof(double) * D.R, cudaMemcpyDeviceToHost, h->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
        if (e != cudaSuccess) rc = fail_cuda(e, "copy estimators", __FILE__, __LINE__);
        else
            for (uint32_t r = 0; r < D.R; r++) {  // qmc_stepper.rs:160-161, qmc_ising.rs:805-809
                double average_n = (double)sn[r] / (double)spr;
                energy_out[r] = -(average_n / beta[r]) + h->offset_h[hr[r]];
            }
    }
    if (rc == QMCB_OK && samples_dev && samples_out) {
        cudaError_t e = cudaMemcpyAsync(samples_out, samples_dev, (size_t)D.R * spr * D.N, cudaMemcpyDeviceToHost, h->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
        if (e != cudaSuccess) rc = fail_cuda(e, "copy samples", __FILE__, __LINE__);
    }
    return rc;
}
extern "C" int qmcb_timesteps(QmcbHandle *h, uint64_t t, uint64_t freq, double *energy_out, uint8_t *samples_out) {
    CHECK_H(h);
    return timesteps_impl(h, t, freq, energy_out, samples_out, false);
}

// QmcAutoCorrelations (autocorrelations.rs:48-97; fft_autocorrelation :99-133): t sweeps sampled every sampling_freq, every
// series mean-removed and normalised, circular autocorrelation averaged over the series.  The reference goes through an
// FFT; here the circular correlation of a two-valued series is counted exactly with XOR + popcount on bit-packed time
// series (C[tau] = T - 2 * mismatches).  Series: the variables themselves (n_products = 0), or products of spins
// (calculate_spin_product_autocorrelation :53-71; a product of +-1 values is the parity of the spins, and the
// normalisation makes the sign convention irrelevant), of which calculate_bond_autocorrelation (:80-97, value_for_bond
// qmc_ising.rs:988-997) is the instance "one product per edge".  Work buffers belong to the handle and are reused.
static int autocorrelation_impl(QmcbHandle *h, uint64_t t, uint64_t freq, uint32_t n_products, const uint32_t *offsets, const uint32_t *vars,
                                double *autocorr_out, uint8_t *samples_out, double *energy_out, uint64_t pt_swap_freq = 0) {
    if (!autocorr_out) return fail(QMCB_ERR_BAD_ARG, "null argument");
    if (freq == 0) freq = 1;
    const SseDev &D = h->D;
    const uint64_t T = t / freq;
    if (T == 0 || T > (1u << 20)) return fail(QMCB_ERR_BAD_ARG, "need between 1 and 2^20 samples");
    uint32_t nv_total = 0;
    if (n_products) {
        if (!offsets || !vars || offsets[0] != 0) return fail(QMCB_ERR_BAD_ARG, "product lists missing");
        for (uint32_t k = 0; k < n_products; k++)
            if (offsets[k + 1] < offsets[k]) return fail(QMCB_ERR_BAD_ARG, "product offsets must not decrease");
        nv_total = offsets[n_products];
        for (uint32_t i = 0; i < nv_total; i++)
            if (vars[i] >= D.N) return fail(QMCB_ERR_BAD_ARG, "product variable out of range");
    }
    int rc;
    if (pt_swap_freq) {
        // ParallelTemperingAutocorrelations::calculate_autocorrelation (tempering_container.rs:580-606): the series of SLOT s
        // are the samples parallel_timesteps_sample (:411-453) collected for the graph at ladder position s, whichever
        // configuration sat there; the series are put in slot order and handed to the same device kernels
        const uint32_t R = D.R, S = h->pt_S;
        if (!h->pt_on) return fail(QMCB_ERR_BAD_ARG, "tempering not configured");
        if (!(h->P.cfg_begin == 0 && R == S))
            return fail(QMCB_ERR_UNSUPPORTED, "tempering autocorrelations need the whole ladder in one handle (a slot's series would be spread over ranks)");
        std::vector<uint8_t> smp((size_t)R * T * D.N), by_slot((size_t)R * T * D.N);
        std::vector<uint32_t> sslots((size_t)R * T);
        std::vector<double> eacc(S);
        if ((rc = qmcb_pt_timesteps_sample(h, t, pt_swap_freq, freq, eacc.data(), smp.data(), sslots.data()))) return rc;
        for (uint32_t r = 0; r < R; r++)
            for (uint64_t k = 0; k < T; k++)
                memcpy(&by_slot[((size_t)sslots[(size_t)r * T + k] * T + k) * D.N], &smp[((size_t)r * T + k) * D.N], D.N);
        if (by_slot.size() > h->samples_cap) {
            if (h->samples_dev) h->pool.release(h->samples_dev);
            h->samples_dev = nullptr, h->samples_cap = 0;
            cudaError_t ea = h->pool.alloc(&h->samples_dev, by_slot.size());
            if (ea != cudaSuccess) return fail_cuda(ea, "sample buffer", __FILE__, __LINE__);
            h->samples_cap = by_slot.size();
        }
        CUDA_TRY(cudaMemcpyAsync(h->samples_dev, by_slot.data(), by_slot.size(), cudaMemcpyHostToDevice, h->stream));
        CUDA_TRY(cudaStreamSynchronize(h->stream));  // by_slot goes out of scope
        if (samples_out) memcpy(samples_out, by_slot.data(), by_slot.size());
        if (energy_out)
            for (uint32_t sl = 0; sl < S; sl++) energy_out[sl] = eacc[sl];  // energy_acc of parallel_timesteps_sample: sum of te * t, as the reference leaves it
    } else if ((rc = timesteps_impl(h, t, freq, energy_out, samples_out, true))) {
        return rc;
    }
    const uint32_t K = n_products ? n_products : D.N;  // number of series per replica
    const uint32_t Tw = (uint32_t)((T + 31) / 32);
    const size_t need_bits = (size_t)D.R * K * (2 * Tw + 1), need_ones = (size_t)D.R * K, need_out = (size_t)D.R * T;
    const size_t need_der = n_products ? (size_t)D.R * T * K : 0, need_lists = n_products ? (size_t)n_products + 1 + nv_total : 0;
    auto ensure = [&](auto **ptr, size_t &cap, size_t need) -> cudaError_t {
        if (need <= cap) return cudaSuccess;
        if (*ptr) h->pool.release(*ptr);
        *ptr = nullptr, cap = 0;
        cudaError_t e = h->pool.alloc(ptr, need);
        if (e == cudaSuccess) cap = need;
        return e;
    };
    cudaError_t e = ensure(&h->ac_bits, h->ac_bits_cap, need_bits);
    if (e == cudaSuccess) e = ensure(&h->ac_ones, h->ac_ones_cap, need_ones);
    if (e == cudaSuccess) e = ensure(&h->ac_out, h->ac_out_cap, need_out);
    if (e == cudaSuccess) e = ensure(&h->ac_derived, h->ac_derived_cap, need_der);
    if (e == cudaSuccess) e = ensure(&h->ac_lists, h->ac_lists_cap, need_lists);
    if (e != cudaSuccess) return fail_cuda(e, "autocorrelation buffers", __FILE__, __LINE__);
    const uint8_t *series = h->samples_dev;
    if (n_products) {
        e = cudaMemcpyAsync(h->ac_lists, offsets, sizeof(uint32_t) * (n_products + 1), cudaMemcpyHostToDevice, h->stream);
        if (e == cudaSuccess && nv_total)
            e = cudaMemcpyAsync(h->ac_lists + n_products + 1, vars, sizeof(uint32_t) * nv_total, cudaMemcpyHostToDevice, h->stream);
        if (e != cudaSuccess) return fail_cuda(e, "copy product lists", __FILE__, __LINE__);
        launch_spin_products(h->samples_dev, D.R, D.N, (uint32_t)T, n_products, h->ac_lists, h->ac_lists + n_products + 1, h->ac_derived, h->stream);
        h->launches += 1;
        series = h->ac_derived;
    }
    launch_autocorrelation(series, D.R, K, (uint32_t)T, h->ac_bits, h->ac_ones, h->ac_out, h->stream);
    h->launches += 2;
    e = cudaMemcpyAsync(autocorr_out, h->ac_out, sizeof(double) * (size_t)D.R * T, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    if (e != cudaSuccess) return fail_cuda(e, "autocorrelation", __FILE__, __LINE__);
    return QMCB_OK;
}
extern "C" int qmcb_variable_autocorrelation(QmcbHandle *h, uint64_t t, uint64_t freq, double *autocorr_out, uint8_t *samples_out,
                                             double *energy_out) {
    CHECK_H(h);
    return autocorrelation_impl(h, t, freq, 0, nullptr, nullptr, autocorr_out, samples_out, energy_out);
}
extern "C" int qmcb_spin_product_autocorrelation(QmcbHandle *h, uint64_t t, uint64_t freq, uint32_t n_products, const uint32_t *product_offsets,
                                                 const uint32_t *product_vars, double *autocorr_out, uint8_t *samples_out, double *energy_out) {
    CHECK_H(h);
    if (n_products == 0) return fail(QMCB_ERR_BAD_ARG, "no products given");
    return autocorrelation_impl(h, t, freq, n_products, product_offsets, product_vars, autocorr_out, samples_out, energy_out);
}
extern "C" int qmcb_bond_autocorrelation(QmcbHandle *h, uint64_t t, uint64_t freq, double *autocorr_out, uint8_t *samples_out, double *energy_out) {
    CHECK_H(h);
    const uint32_t E = h->D.E;  // n_bonds() = edges.len() (qmc_ising.rs:984-986)
    if (E == 0) return fail(QMCB_ERR_BAD_ARG, "lattice without edges");
    std::vector<uint32_t> off(E + 1), vars(2 * (size_t)E);
    for (uint32_t b = 0; b < E; b++) off[b] = 2 * b, vars[2 * b] = h->va_h[b], vars[2 * b + 1] = h->vb_h[b];
    off[E] = 2 * E;
    return autocorrelation_impl(h, t, freq, E, off.data(), vars.data(), autocorr_out, samples_out, energy_out);
}

// ParallelTemperingAutocorrelations / ParallelTemperingBondAutoCorrelations for TemperingContainer
// (tempering_container.rs:484-630): one autocorrelation per ladder slot, autocorr_out [S][T]
extern "C" int qmcb_pt_variable_autocorrelation(QmcbHandle *h, uint64_t timesteps, uint64_t replica_swap_freq, uint64_t sampling_freq,
                                                double *autocorr_out, uint8_t *samples_out, double *energy_out) {
    CHECK_H(h);
    if (replica_swap_freq == 0) replica_swap_freq = 1;  // Option::unwrap_or(1) :590
    return autocorrelation_impl(h, timesteps, sampling_freq, 0, nullptr, nullptr, autocorr_out, samples_out, energy_out, replica_swap_freq);
}
extern "C" int qmcb_pt_spin_product_autocorrelation(QmcbHandle *h, uint64_t timesteps, uint64_t replica_swap_freq, uint64_t sampling_freq,
                                                    uint32_t n_products, const uint32_t *product_offsets, const uint32_t *product_vars,
                                                    double *autocorr_out, uint8_t *samples_out, double *energy_out) {
    CHECK_H(h);
    if (n_products == 0) return fail(QMCB_ERR_BAD_ARG, "no products given");
    if (replica_swap_freq == 0) replica_swap_freq = 1;
    return autocorrelation_impl(h, timesteps, sampling_freq, n_products, product_offsets, product_vars, autocorr_out, samples_out, energy_out, replica_swap_freq);
}
extern "C" int qmcb_pt_bond_autocorrelation(QmcbHandle *h, uint64_t timesteps, uint64_t replica_swap_freq, uint64_t sampling_freq,
                                            double *autocorr_out, uint8_t *samples_out, double *energy_out) {
    CHECK_H(h);
    const uint32_t E = h->D.E;
    if (E == 0) return fail(QMCB_ERR_BAD_ARG, "lattice without edges");
    if (replica_swap_freq == 0) replica_swap_freq = 1;
    std::vector<uint32_t> off(E + 1), vars(2 * (size_t)E);
    for (uint32_t b = 0; b < E; b++) off[b] = 2 * b, vars[2 * b] = h->va_h[b], vars[2 * b + 1] = h->vb_h[b];
    off[E] = 2 * E;
    return autocorrelation_impl(h, timesteps, sampling_freq, E, off.data(), vars.data(), autocorr_out, samples_out, energy_out, replica_swap_freq);
}

extern "C" int qmcb_enqueue_sweeps(QmcbHandle *h, uint64_t t) {
    CHECK_H(h);
    const uint64_t origin = h->target;
    h->target += t;
    return launch_sweeps(h, 1u | 2u | 4u | 8u, ~0ull, origin, nullptr, 0);
}
extern "C" int qmcb_synchronize(QmcbHandle *h) {
    CHECK_H(h);
    int st = 0, rc;
    if ((rc = check_status(h, &st))) return rc;
    if ((rc = status_to_error(st))) return rc;
    if (st & DEV_ERR_CAPACITY) {  // finish the sweeps that were cut short
        CUDA_TRY(cudaMemsetAsync(h->D.status, 0, sizeof(int), h->stream));
        if (!h->auto_capacity) return fail(QMCB_ERR_CAPACITY, "cutoff outgrew the capacity given to qmcb_create");
        std::vector<uint32_t> M(h->D.R);
        CUDA_TRY(cudaMemcpy(M.data(), h->D.M, sizeof(uint32_t) * h->D.R, cudaMemcpyDeviceToHost));
        uint64_t need = *std::max_element(M.begin(), M.end());
        if ((rc = grow(h, need + need / 2 + 1024))) return rc;
        return run_to_target(h, 1u | 2u | 4u | 8u, ~0ull, h->target, nullptr, 0);
    }
    return QMCB_OK;
}
// single steps are not resumable, so make room before launching them
static int ensure_capacity(QmcbHandle *h) {
    std::vector<uint32_t> M(h->D.R);
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    CUDA_TRY(cudaMemcpy(M.data(), h->D.M, sizeof(uint32_t) * h->D.R, cudaMemcpyDeviceToHost));
    uint64_t need = *std::max_element(M.begin(), M.end());
    if (need <= h->D.cap) return QMCB_OK;
    if (!h->auto_capacity) return fail(QMCB_ERR_CAPACITY, "cutoff outgrew the capacity given to qmcb_create");
    return grow(h, need + need / 2 + 1024);
}
extern "C" int qmcb_single_diagonal_step(QmcbHandle *h) {
    CHECK_H(h);
    int rc = ensure_capacity(h);
    return rc ? rc : run_to_target(h, 1u | 4u, 1, 0, nullptr, 0);
}
extern "C" int qmcb_single_cluster_step(QmcbHandle *h, uint64_t *ncl_out) {
    CHECK_H(h);
    int rc = ensure_capacity(h);
    if (rc) return rc;
    if ((rc = run_to_target(h, 2u, 1, 0, nullptr, 0))) return rc;
    if (ncl_out) {
        std::vector<uint32_t> ncl(h->D.R);
        CUDA_TRY(cudaMemcpy(ncl.data(), h->D.ncl, sizeof(uint32_t) * h->D.R, cudaMemcpyDeviceToHost));
        for (uint32_t r = 0; r < h->D.R; r++) ncl_out[r] = ncl[r];
    }
    return QMCB_OK;
}
// QmcIsingGraph::set_run_rvb (qmc_ising.rs:434-441)
extern "C" int qmcb_set_run_rvb(QmcbHandle *h, int run_rvb) {
    CHECK_H(h);
    if (run_rvb && h->generic) return fail(QMCB_ERR_UNSUPPORTED, "the RVB update belongs to QmcIsingGraph (qmc_ising.rs:705-752); Qmc has none");
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    if (run_rvb) {
        int rc = alloc_rvb_ws(h);
        if (rc) return rc;
    }
    h->rvb_on = run_rvb != 0;
    return QMCB_OK;
}
extern "C" int qmcb_get_run_rvb(const QmcbHandle *h, int *run_rvb) {
    if (!h || !run_rvb) return fail(QMCB_ERR_BAD_ARG, "null argument");
    *run_rvb = h->rvb_on ? 1 : 0;
    return QMCB_OK;
}
// QmcIsingGraph::single_rvb_sweep (qmc_ising.rs:322-420): updates_in_sweep < 0 = None = (nvars + 1) / 2
extern "C" int qmcb_single_rvb_sweep(QmcbHandle *h, int64_t updates_in_sweep, uint64_t *successes_out, uint64_t *attempts_out) {
    CHECK_H(h);
    if (h->generic) return fail(QMCB_ERR_UNSUPPORTED, "the RVB update belongs to QmcIsingGraph (qmc_ising.rs:322-420); Qmc has none");
    int rc = ensure_capacity(h);
    if (rc || (rc = alloc_rvb_ws(h))) return rc;
    h->launches += (uint64_t)launch_sse_rvb(h->D, h->W, 0, updates_in_sweep < 0 ? -1 : (long long)updates_in_sweep, h->rvb_succ_last, h->stream);
    CUDA_TRY(cudaGetLastError());
    int st = 0;
    if ((rc = check_status(h, &st)) || (rc = status_to_error(st))) return rc;
    if (successes_out) {
        static_assert(sizeof(unsigned long long) == sizeof(uint64_t), "");
        CUDA_TRY(cudaMemcpy(successes_out, h->rvb_succ_last, sizeof(uint64_t) * h->D.R, cudaMemcpyDeviceToHost));
    }
    if (attempts_out) *attempts_out = updates_in_sweep < 0 ? ((uint64_t)h->D.N + 1) / 2 : (uint64_t)updates_in_sweep;
    return QMCB_OK;
}
// QmcIsingGraph::rvb_success_rate (qmc_ising.rs:604-607) of every replica: total_rvb_successes / rvb_clusters_counted
// (NaN before the first sweep with RVB steps, as 0 / 0 is in the reference); the two totals on request
extern "C" int qmcb_rvb_success_rate(QmcbHandle *h, double *rate_out, uint64_t *successes_out, uint64_t *counted_out) {
    CHECK_H(h);
    if (!rate_out && !successes_out && !counted_out) return fail(QMCB_ERR_BAD_ARG, "null argument");
    const uint32_t R = h->D.R;
    std::vector<uint64_t> s(R, 0), c(R, 0);
    if (h->W.succ) {
        CUDA_TRY(cudaStreamSynchronize(h->stream));
        CUDA_TRY(cudaMemcpy(s.data(), h->W.succ, sizeof(uint64_t) * R, cudaMemcpyDeviceToHost));
        CUDA_TRY(cudaMemcpy(c.data(), h->W.count, sizeof(uint64_t) * R, cudaMemcpyDeviceToHost));
    }
    for (uint32_t r = 0; r < R; r++) {
        if (rate_out) rate_out[r] = (double)s[r] / (double)c[r];
        if (successes_out) successes_out[r] = s[r];
        if (counted_out) counted_out[r] = c[r];
    }
    return QMCB_OK;
}
// Qmc::loop_update (qmc_runner.rs:205-220): one directed-loop update of every replica
extern "C" int qmcb_loop_update(QmcbHandle *h) {
    CHECK_H(h);
    if (!h->generic) return fail(QMCB_ERR_UNSUPPORTED, "loop updates belong to handles made by qmcb_create_qmc (QmcIsingGraph has none, qmc_ising.rs:644-795)");
    if (h->mode != QMCB_MODE_STRICT) return fail(QMCB_ERR_UNSUPPORTED, "loop updates run in QMCB_MODE_STRICT");
    int rc = ensure_capacity(h);
    return rc ? rc : run_to_target(h, 32u, 1, 0, nullptr, 0);
}
// Qmc::set_do_loop_updates (qmc_runner.rs:268-270)
extern "C" int qmcb_set_do_loop_updates(QmcbHandle *h, int enable) {
    CHECK_H(h);
    if (!h->generic) return fail(QMCB_ERR_UNSUPPORTED, "loop updates belong to handles made by qmcb_create_qmc");
    if (enable && h->mode != QMCB_MODE_STRICT) return fail(QMCB_ERR_UNSUPPORTED, "loop updates run in QMCB_MODE_STRICT: set the mode first");
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    h->D.loop_updates = enable ? 1 : 0;
    h->D.loop_path = h->D.loop_updates || h->D.no_site || h->offdiag2;
    return QMCB_OK;
}
extern "C" int qmcb_get_do_loop_updates(const QmcbHandle *h, int *enabled) {
    if (!h || !enabled) return fail(QMCB_ERR_BAD_ARG, "null argument");
    *enabled = h->D.loop_updates;
    return QMCB_OK;
}
extern "C" int qmcb_total_vertex_updates(QmcbHandle *h, uint64_t *total) {
    CHECK_H(h);
    if (!total) return fail(QMCB_ERR_BAD_ARG, "null argument");
    std::vector<unsigned long long> v(h->D.R);
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    CUDA_TRY(cudaMemcpy(v.data(), h->D.vupd, sizeof(uint64_t) * h->D.R, cudaMemcpyDeviceToHost));
    uint64_t s = 0;
    for (auto x : v) s += x;
    *total = s;
    return QMCB_OK;
}
extern "C" int qmcb_launch_count(const QmcbHandle *h, uint64_t *launches) {
    if (!h || !launches) return fail(QMCB_ERR_BAD_ARG, "null argument");
    *launches = h->launches;
    return QMCB_OK;
}

static int copy_u32_as_u64(QmcbHandle *h, const uint32_t *dev, uint64_t *out) {
    std::vector<uint32_t> v(h->D.R);
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    CUDA_TRY(cudaMemcpy(v.data(), dev, sizeof(uint32_t) * h->D.R, cudaMemcpyDeviceToHost));
    for (uint32_t r = 0; r < h->D.R; r++) out[r] = v[r];
    return QMCB_OK;
}

extern "C" int qmcb_get_states(QmcbHandle *h, uint8_t *states) {
    CHECK_H(h);
    if (!states) return fail(QMCB_ERR_BAD_ARG, "null argument");
    const SseDev &D = h->D;
    std::vector<uint32_t> packed((size_t)D.R * D.Nw);
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    CUDA_TRY(cudaMemcpy(packed.data(), D.state, packed.size() * 4, cudaMemcpyDeviceToHost));
    for (uint32_t r = 0; r < D.R; r++)
        for (uint32_t v = 0; v < D.N; v++) states[(size_t)r * D.N + v] = (packed[(size_t)r * D.Nw + (v >> 5)] >> (v & 31)) & 1u;
    return QMCB_OK;
}
extern "C" int qmcb_get_state(QmcbHandle *h, uint32_t r, uint8_t *state) {
    CHECK_H(h);
    const SseDev &D = h->D;
    if (!state || r >= D.R) return fail(QMCB_ERR_BAD_ARG, "bad replica index");
    std::vector<uint32_t> packed(D.Nw);
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    CUDA_TRY(cudaMemcpy(packed.data(), D.state + (size_t)r * D.Nw, D.Nw * 4, cudaMemcpyDeviceToHost));
    for (uint32_t v = 0; v < D.N; v++) state[v] = (packed[v >> 5] >> (v & 31)) & 1u;
    return QMCB_OK;
}
extern "C" int qmcb_set_state(QmcbHandle *h, uint32_t r, const uint8_t *state) {
    CHECK_H(h);
    const SseDev &D = h->D;
    if (!state || r >= D.R) return fail(QMCB_ERR_BAD_ARG, "bad replica index");
    std::vector<uint32_t> packed(D.Nw, 0u);
    for (uint32_t v = 0; v < D.N; v++)
        if (state[v]) packed[v >> 5] |= 1u << (v & 31);
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    CUDA_TRY(cudaMemcpy(D.state + (size_t)r * D.Nw, packed.data(), D.Nw * 4, cudaMemcpyHostToDevice));
    return QMCB_OK;
}
extern "C" int qmcb_get_n(QmcbHandle *h, uint64_t *n) {
    CHECK_H(h);
    return n ? copy_u32_as_u64(h, h->D.n, n) : fail(QMCB_ERR_BAD_ARG, "null argument");
}
extern "C" int qmcb_get_cutoffs(QmcbHandle *h, uint64_t *c) {
    CHECK_H(h);
    return c ? copy_u32_as_u64(h, h->D.M, c) : fail(QMCB_ERR_BAD_ARG, "null argument");
}
extern "C" int qmcb_set_cutoff(QmcbHandle *h, uint32_t r, uint64_t cutoff) {
    CHECK_H(h);
    if (r >= h->D.R) return fail(QMCB_ERR_BAD_ARG, "bad replica index");
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    if (cutoff > h->D.cap) {
        if (!h->auto_capacity) return fail(QMCB_ERR_CAPACITY, "cutoff above capacity");
        int rc = grow(h, cutoff + cutoff / 2);
        if (rc) return rc;
    }
    // QmcIsingGraph::set_cutoff (qmc_ising.rs:537-540) sets the cutoff unconditionally and the reference then panics
    // on the usize underflow of M - n.  Here shrinking below the last occupied slot is refused: the diagonal update
    // would compute M - n on wrapped integers and ops above the cutoff would be orphaned.
    {
        uint32_t Mold = 0;
        CUDA_TRY(cudaMemcpy(&Mold, h->D.M + r, sizeof(uint32_t), cudaMemcpyDeviceToHost));
        if (cutoff < Mold) {
            const uint64_t lim = std::min<uint64_t>(Mold, h->D.cap);
            std::vector<uint32_t> tail(lim - cutoff);
            CUDA_TRY(cudaMemcpy(tail.data(), h->D.ops + (size_t)r * h->D.cap + cutoff, 4 * (lim - cutoff), cudaMemcpyDeviceToHost));
            for (uint32_t w : tail)
                if (w != QMCB_OP_EMPTY) return fail(QMCB_ERR_BAD_ARG, "cutoff below the last occupied slot of the operator string");
        }
    }
    uint32_t c = (uint32_t)cutoff;
    CUDA_TRY(cudaMemcpy(h->D.M + r, &c, sizeof(uint32_t), cudaMemcpyHostToDevice));
    return QMCB_OK;
}
extern "C" int qmcb_get_capacity(const QmcbHandle *h, uint64_t *cap) {
    if (!h || !cap) return fail(QMCB_ERR_BAD_ARG, "null argument");
    *cap = h->D.cap;
    return QMCB_OK;
}
extern "C" int qmcb_get_offset(const QmcbHandle *h, double *offset) {
    if (!h || !offset) return fail(QMCB_ERR_BAD_ARG, "null argument");
    *offset = h->offset;
    return QMCB_OK;
}
extern "C" int qmcb_get_bond_counts(QmcbHandle *h, uint32_t r, uint64_t *counts) {
    CHECK_H(h);
    const SseDev &D = h->D;
    if (!counts || r >= D.R) return fail(QMCB_ERR_BAD_ARG, "bad replica index");
    unsigned long long *dev = nullptr;
    const uint32_t nb = D.E + 2 * D.N;
    CUDA_TRY(cudaMalloc(&dev, sizeof(uint64_t) * nb));
    cudaMemsetAsync(dev, 0, sizeof(uint64_t) * nb, h->stream);
    launch_sse_bond_counts(D, r, dev, h->stream);
    h->launches++;
    std::vector<unsigned long long> host(nb);
    cudaError_t e = cudaMemcpyAsync(host.data(), dev, sizeof(uint64_t) * nb, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(dev);
    if (e != cudaSuccess) return fail_cuda(e, "bond counts", __FILE__, __LINE__);
    for (uint32_t b = 0; b < D.Nb; b++) counts[b] = host[b];
    return QMCB_OK;
}
// imaginary_time_fold (qmc_stepper.rs:165-168, qmc_ising.rs:815-821) with the magnetisation fold, all replicas
extern "C" int qmcb_itime_magnetization(QmcbHandle *h, double *m_mean, double *m_sq, double *m_abs) {
    CHECK_H(h);
    const SseDev &D = h->D;
    long long *dev = nullptr;
    CUDA_TRY(cudaMalloc(&dev, sizeof(long long) * 3 * D.R));
    launch_sse_itime_magnetization(D, dev, h->stream);
    h->launches++;
    std::vector<long long> s(3 * (size_t)D.R);
    std::vector<uint32_t> M(D.R);
    cudaError_t e = cudaMemcpyAsync(s.data(), dev, sizeof(long long) * s.size(), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(M.data(), D.M, sizeof(uint32_t) * D.R, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(dev);
    if (e != cudaSuccess) return fail_cuda(e, "itime magnetization", __FILE__, __LINE__);
    for (uint32_t r = 0; r < D.R; r++) {  // per-site magnetisation averaged over the M imaginary-time slices
        const double slots = (double)M[r], nv = (double)D.N;
        if (m_mean) m_mean[r] = (double)s[3 * (size_t)r] / slots / nv;
        if (m_sq) m_sq[r] = (double)s[3 * (size_t)r + 1] / slots / (nv * nv);
        if (m_abs) m_abs[r] = (double)s[3 * (size_t)r + 2] / slots / nv;
    }
    return QMCB_OK;
}
// the state the fold closure of imaginary_time_fold is handed at slot p (fast_ops.rs:1300-1311)
extern "C" int qmcb_itime_state(QmcbHandle *h, uint32_t r, uint64_t p, uint8_t *state) {
    CHECK_H(h);
    const SseDev &D = h->D;
    if (!state || r >= D.R) return fail(QMCB_ERR_BAD_ARG, "bad replica index");
    uint32_t *dev = nullptr;
    CUDA_TRY(cudaMalloc(&dev, sizeof(uint32_t) * D.Nw));
    launch_sse_itime_state(D, r, p, dev, h->stream);
    h->launches++;
    std::vector<uint32_t> packed(D.Nw);
    cudaError_t e = cudaMemcpyAsync(packed.data(), dev, sizeof(uint32_t) * D.Nw, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(dev);
    if (e != cudaSuccess) return fail_cuda(e, "itime state", __FILE__, __LINE__);
    for (uint32_t v = 0; v < D.N; v++) state[v] = (packed[v >> 5] >> (v & 31)) & 1u;
    return QMCB_OK;
}
extern "C" int qmcb_get_rng_cursors(QmcbHandle *h, uint64_t *cursors) {
    CHECK_H(h);
    if (!cursors) return fail(QMCB_ERR_BAD_ARG, "null argument");
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    CUDA_TRY(cudaMemcpy(cursors, h->D.cursor, sizeof(uint64_t) * h->D.R, cudaMemcpyDeviceToHost));
    return QMCB_OK;
}
extern "C" int qmcb_set_rng_cursor(QmcbHandle *h, uint32_t r, uint64_t cursor) {
    CHECK_H(h);
    if (r >= h->D.R) return fail(QMCB_ERR_BAD_ARG, "bad replica index");
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    CUDA_TRY(cudaMemcpy(h->D.cursor + r, &cursor, sizeof(uint64_t), cudaMemcpyHostToDevice));
    return QMCB_OK;
}
extern "C" int qmcb_get_rng_keys(QmcbHandle *h, uint64_t *keys) {
    CHECK_H(h);
    if (!keys) return fail(QMCB_ERR_BAD_ARG, "null argument");
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    CUDA_TRY(cudaMemcpy(keys, h->D.key, sizeof(uint64_t) * h->D.R, cudaMemcpyDeviceToHost));
    return QMCB_OK;
}
extern "C" int qmcb_dump_ops(QmcbHandle *h, uint32_t r, uint32_t *words, uint64_t nwords) {
    CHECK_H(h);
    const SseDev &D = h->D;
    if (!words || r >= D.R) return fail(QMCB_ERR_BAD_ARG, "bad replica index");
    uint32_t M = 0;
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    CUDA_TRY(cudaMemcpy(&M, D.M + r, sizeof(uint32_t), cudaMemcpyDeviceToHost));
    if (nwords < M) return fail(QMCB_ERR_BAD_ARG, "buffer shorter than the cutoff");
    const uint64_t have = std::min<uint64_t>(M, D.cap);
    CUDA_TRY(cudaMemcpy(words, D.ops + (size_t)r * D.cap, have * 4, cudaMemcpyDeviceToHost));
    for (uint64_t p = have; p < nwords; p++) words[p] = QMCB_OP_EMPTY;
    return QMCB_OK;
}
// a well-formed non-identity operator word: bond index in range, no stray bits, no second-leg bits on a one-variable op
static bool op_word_ok(const SseDev &D, uint32_t w) {
    const uint32_t b = w & 0xFFFFFFu;
    if ((w >> 28) || b >= D.Nb) return false;
    if (b >= D.E && (((w >> 25) & 1u) || ((w >> 27) & 1u))) return false;
    return true;
}
extern "C" int qmcb_load_ops(QmcbHandle *h, uint32_t r, const uint32_t *words, uint64_t nwords, const uint8_t *state) {
    CHECK_H(h);
    SseDev &D = h->D;
    if (!words || r >= D.R) return fail(QMCB_ERR_BAD_ARG, "bad replica index");
    for (uint64_t p = 0; p < nwords; p++) {
        if (words[p] == QMCB_OP_EMPTY) continue;
        if (!op_word_ok(D, words[p])) return fail(QMCB_ERR_BAD_ARG, "malformed operator word (bond out of range, stray bits, or second-leg bits on a one-variable op)");
    }
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    if (nwords > D.cap) {
        if (!h->auto_capacity) return fail(QMCB_ERR_CAPACITY, "string longer than capacity");
        int rc = grow(h, nwords + nwords / 2);
        if (rc) return rc;
    }
    CUDA_TRY(cudaMemset(D.ops + (size_t)r * D.cap, 0xFF, D.cap * 4));
    CUDA_TRY(cudaMemcpy(D.ops + (size_t)r * D.cap, words, nwords * 4, cudaMemcpyHostToDevice));
    uint32_t M = 0;
    CUDA_TRY(cudaMemcpy(&M, D.M + r, sizeof(uint32_t), cudaMemcpyDeviceToHost));
    if (nwords > M) {
        M = (uint32_t)nwords;
        CUDA_TRY(cudaMemcpy(D.M + r, &M, sizeof(uint32_t), cudaMemcpyHostToDevice));
    }
    launch_sse_recount(D, r, h->stream);
    h->launches++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    if (state) return qmcb_set_state(h, r, state);
    return QMCB_OK;
}
extern "C" int qmcb_verify(QmcbHandle *h, uint32_t r, int *ok) {
    CHECK_H(h);
    if (!ok || r >= h->D.R) return fail(QMCB_ERR_BAD_ARG, "bad replica index");
    int *ok_dev = nullptr;
    uint32_t *scratch = nullptr;
    CUDA_TRY(cudaMalloc(&ok_dev, sizeof(int)));
    cudaError_t e = cudaMalloc(&scratch, sizeof(uint32_t) * h->D.Nw);
    if (e == cudaSuccess) {
        launch_sse_verify(h->D, r, ok_dev, scratch, h->stream);
        h->launches++;
        e = cudaMemcpyAsync(ok, ok_dev, sizeof(int), cudaMemcpyDeviceToHost, h->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    }
    cudaFree(ok_dev), cudaFree(scratch);
    if (e != cudaSuccess) return fail_cuda(e, "verify", __FILE__, __LINE__);
    return QMCB_OK;
}
extern "C" int qmcb_get_boundaries(QmcbHandle *h, uint32_t r, uint32_t *b_in, uint32_t *b_out, uint64_t nslots) {
    CHECK_H(h);
    const SseDev &D = h->D;
    if (!b_in || !b_out || r >= D.R) return fail(QMCB_ERR_BAD_ARG, "bad replica index");
    if (!h->strict_ws) return fail(QMCB_ERR_BAD_ARG, "no STRICT cluster step has run yet");
    uint64_t k = std::min<uint64_t>(nslots, D.cap);
    std::vector<uint32_t> ow(k);
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    CUDA_TRY(cudaMemcpy(ow.data(), D.ops + (size_t)r * D.cap, 4 * k, cudaMemcpyDeviceToHost));
    if (h->strict_wl_last) {  // world-line layout: the boundaries sit in the entry of leg 0, ent[p]
        std::vector<uint32_t> wl(strict_rec_stride(D)), ent(k);
        CUDA_TRY(cudaMemcpy(wl.data(), D.rec + (size_t)r * strict_rec_stride(D), 4 * wl.size(), cudaMemcpyDeviceToHost));
        CUDA_TRY(cudaMemcpy(ent.data(), D.ent + (size_t)r * D.cap, 4 * k, cudaMemcpyDeviceToHost));
        for (uint64_t p = 0; p < nslots; p++) {
            const bool has = p < k && ow[p] != QMCB_OP_EMPTY;
            b_in[p] = has ? wl[4 * (size_t)ent[p] + 2] : NONE32, b_out[p] = has ? wl[4 * (size_t)ent[p] + 3] : NONE32;
        }
        return QMCB_OK;
    }
    std::vector<uint32_t> b(8 * k);
    CUDA_TRY(cudaMemcpy(b.data(), D.rec + (size_t)r * strict_rec_stride(D), 32 * k, cudaMemcpyDeviceToHost));
    for (uint64_t p = 0; p < nslots; p++) {
        const bool has = p < k && ow[p] != QMCB_OP_EMPTY;  // records of empty slots are stale
        b_in[p] = has ? b[8 * p + 5] : NONE32, b_out[p] = has ? b[8 * p + 6] : NONE32;
    }
    return QMCB_OK;
}

// ---- tempering --------------------------------------------------------------------------
extern "C" int qmcb_pt_configure(QmcbHandle *h, uint32_t n_chains, uint32_t n_betas, uint32_t slot_begin,
                                 const double *betas_global, const uint64_t *keys_global, uint64_t pt_key) {
    CHECK_H(h);
    SseDev &D = h->D;
    const uint64_t S = (uint64_t)n_chains * n_betas;
    if (!betas_global || !keys_global || S == 0 || slot_begin + (uint64_t)D.R > S)
        return fail(QMCB_ERR_BAD_ARG, "tempering ladder does not cover this handle's replicas");
    // SwapManagers::can_swap_graphs (qmc_ising.rs:563-590) holds trivially: one lattice per handle
    PtDev &P = h->P;
    P.n_chains = n_chains, P.n_betas = n_betas, P.cfg_begin = slot_begin, P.pt_key = pt_key;
    double *bs;
    uint64_t *ks;
    CUDA_TRY(h->pool.alloc(&bs, S));
    CUDA_TRY(h->pool.alloc(&ks, S));
    CUDA_TRY(cudaMemcpy(bs, betas_global, sizeof(double) * S, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(ks, keys_global, sizeof(uint64_t) * S, cudaMemcpyHostToDevice));
    P.beta_slot = bs, P.key_slot = ks;
    CUDA_TRY(h->pool.alloc(&P.pt_cursor, 1));
    CUDA_TRY(cudaMemset(P.pt_cursor, 0, sizeof(uint64_t)));
    CUDA_TRY(h->pool.alloc(&P.swaps, 1));
    CUDA_TRY(cudaMemset(P.swaps, 0, sizeof(uint64_t)));
    CUDA_TRY(h->pool.alloc(&P.slot_of_local, D.R));
    std::vector<uint32_t> slots(D.R);
    for (uint32_t s = 0; s < D.R; s++) slots[s] = slot_begin + s;
    CUDA_TRY(cudaMemcpy(P.slot_of_local, slots.data(), sizeof(uint32_t) * D.R, cudaMemcpyHostToDevice));
    CUDA_TRY(h->pool.alloc(&P.n_slot, S));
    CUDA_TRY(h->pool.alloc(&P.cursor_slot, S));
    CUDA_TRY(h->pool.alloc(&P.cfg_slot, S));
    CUDA_TRY(h->pool.alloc(&P.maxM_chain, n_chains));
    // the local configurations start in slots [slot_begin, slot_begin + R): take those labels
    CUDA_TRY(cudaMemcpy(D.beta, betas_global + slot_begin, sizeof(double) * D.R, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(D.key, keys_global + slot_begin, sizeof(uint64_t) * D.R, cudaMemcpyHostToDevice));
    h->pt_on = true;
    h->pt_S = (uint32_t)S;
    return QMCB_OK;
}
// Hamiltonian of every slot of the ladder (rows of qmcb_set_hamiltonians' tables): like beta it is a label of
// the SLOT, configurations move between slots (swap_manager_and_state, qmc_ising.rs:593-602, leaves edges and
// fields where they are).  Swaps between slots whose rows are not ham_eq (tempering_traits.rs:122-124; NB it
// compares edges and transverse field only) get the relative_weight factors of tempering_container.rs:286-292.
extern "C" int qmcb_pt_set_slot_hamiltonians(QmcbHandle *h, const uint32_t *ham_of_slot) {
    CHECK_H(h);
    SseDev &D = h->D;
    PtDev &P = h->P;
    if (!h->pt_on || !ham_of_slot) return fail(QMCB_ERR_BAD_ARG, "tempering not configured");
    if (!D.ham) return fail(QMCB_ERR_BAD_ARG, "call qmcb_set_hamiltonians first");
    const uint32_t S = h->pt_S, H = h->H;
    for (uint32_t s = 0; s < S; s++)
        if (ham_of_slot[s] >= H) return fail(QMCB_ERR_BAD_ARG, "Hamiltonian index out of range");
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    std::vector<uint8_t> eq((size_t)H * H);
    for (uint32_t a = 0; a < H; a++)
        for (uint32_t b = 0; b < H; b++)  // HamInfo::eq, qmc_ising.rs:899-903
            eq[(size_t)a * H + b] = h->gam_h[a] == h->gam_h[b] &&
                                    std::equal(h->Jtab_h.begin() + (size_t)a * D.E, h->Jtab_h.begin() + (size_t)(a + 1) * D.E, h->Jtab_h.begin() + (size_t)b * D.E);
    uint32_t *hs;
    uint8_t *he;
    CUDA_TRY(h->pool.alloc(&hs, S));
    CUDA_TRY(h->pool.alloc(&he, (size_t)H * H));
    CUDA_TRY(cudaMemcpy(hs, ham_of_slot, sizeof(uint32_t) * S, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(he, eq.data(), eq.size(), cudaMemcpyHostToDevice));
    h->pool.release((void *)P.ham_slot), h->pool.release((void *)P.ham_eq), h->pool.release(P.counts), h->pool.release(P.oslot_cfg);
    P.ham_slot = hs, P.ham_eq = he, P.H = H;
    CUDA_TRY(h->pool.alloc(&P.counts, (size_t)D.R * D.Nb));
    CUDA_TRY(h->pool.alloc(&P.oslot_cfg, S));
    h->ham_slot_h.assign(ham_of_slot, ham_of_slot + S);
    // the local configurations take the label of the slot they are in
    std::vector<uint32_t> slots(D.R), hr(D.R);
    CUDA_TRY(cudaMemcpy(slots.data(), P.slot_of_local, sizeof(uint32_t) * D.R, cudaMemcpyDeviceToHost));
    for (uint32_t r = 0; r < D.R; r++) hr[r] = ham_of_slot[slots[r]];
    CUDA_TRY(cudaMemcpy((void *)D.ham, hr.data(), sizeof(uint32_t) * D.R, cudaMemcpyHostToDevice));
    return QMCB_OK;
}
extern "C" int qmcb_pt_record_words(const QmcbHandle *h, uint32_t *words) {
    if (!h || !words) return fail(QMCB_ERR_BAD_ARG, "null argument");
    *words = h->P.ham_slot ? PT_REC_WORDS_MH : PT_REC_WORDS_EQ;
    return QMCB_OK;
}
extern "C" int qmcb_pt_export(QmcbHandle *h, uint64_t *rec_dev) {
    CHECK_H(h);
    if (!h->pt_on || !rec_dev) return fail(QMCB_ERR_BAD_ARG, "tempering not configured");
    h->launches += (uint64_t)launch_pt_export(h->D, h->P, rec_dev, h->stream) - 1;
    h->launches++;
    CUDA_TRY(cudaGetLastError());
    return QMCB_OK;
}
extern "C" int qmcb_pt_apply(QmcbHandle *h, const uint64_t *all_rec_dev, uint64_t n_records) {
    CHECK_H(h);
    if (!h->pt_on || !all_rec_dev || n_records != h->pt_S) return fail(QMCB_ERR_BAD_ARG, "record count does not match the ladder");
    if (h->D.ham && !h->P.ham_slot) return fail(QMCB_ERR_BAD_ARG, "replicas have their own Hamiltonians: call qmcb_pt_set_slot_hamiltonians before tempering");
    launch_pt_apply(h->D, h->P, all_rec_dev, h->pt_S, h->stream);
    h->launches++;
    CUDA_TRY(cudaGetLastError());
    return QMCB_OK;
}
// tempering_step for a container that lives on ONE handle (every slot of every ladder is local): export + apply with
// a record buffer owned by the handle, no host plumbing in between
extern "C" int qmcb_pt_step_local(QmcbHandle *h) {
    CHECK_H(h);
    if (!h->pt_on) return fail(QMCB_ERR_BAD_ARG, "tempering not configured");
    if (h->P.cfg_begin != 0 || h->D.R != h->pt_S) return fail(QMCB_ERR_BAD_ARG, "the ladder is spread over several handles: use qmcb_pt_export + all-gather + qmcb_pt_apply");
    if (!h->pt_rec_dev) CUDA_TRY(h->pool.alloc(&h->pt_rec_dev, (size_t)h->pt_S * PT_REC_WORDS_MH));
    int rc = qmcb_pt_export(h, h->pt_rec_dev);
    return rc ? rc : qmcb_pt_apply(h, h->pt_rec_dev, h->pt_S);
}
// ---- multi-GPU tempering: the all-gather of tempering_step inside the library ----------------------------
// NCCL is bound at run time (dlopen): libqmcb.so has no link-time dependency on it, and inside a process that
// already carries an NCCL (torch) the same copy is used.  Only the five entry points below are needed.
namespace {
struct NcclUid { char internal[128]; };  // ncclUniqueId
typedef int (*nccl_get_uid_t)(NcclUid *);
typedef int (*nccl_init_rank_t)(void **, int, NcclUid, int);
typedef int (*nccl_allgather_t)(const void *, void *, size_t, int, void *, cudaStream_t);
typedef int (*nccl_allreduce_t)(const void *, void *, size_t, int, int, void *, cudaStream_t);
typedef int (*nccl_destroy_t)(void *);
typedef const char *(*nccl_errstr_t)(int);
struct NcclApi {
    void *lib = nullptr;
    nccl_get_uid_t get_uid = nullptr;
    nccl_init_rank_t init_rank = nullptr;
    nccl_allgather_t allgather = nullptr;
    nccl_allreduce_t allreduce = nullptr;
    nccl_destroy_t destroy = nullptr;
    nccl_errstr_t errstr = nullptr;
};
const int NCCL_UINT64 = 5, NCCL_FLOAT64 = 8, NCCL_SUM = 0;  // ncclDataType_t / ncclRedOp_t values (nccl.h)
NcclApi *nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char *names[] = {getenv("QMCB_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char *nm : names) {
            if (!nm || !*nm) continue;
            void *l = dlopen(nm, RTLD_NOW | RTLD_NOLOAD);  // the copy the process already has (torch's), if any
            if (!l) l = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (l) { api.lib = l; break; }
        }
        if (!api.lib) return;
        api.get_uid = (nccl_get_uid_t)dlsym(api.lib, "ncclGetUniqueId");
        api.init_rank = (nccl_init_rank_t)dlsym(api.lib, "ncclCommInitRank");
        api.allgather = (nccl_allgather_t)dlsym(api.lib, "ncclAllGather");
        api.allreduce = (nccl_allreduce_t)dlsym(api.lib, "ncclAllReduce");
        api.destroy = (nccl_destroy_t)dlsym(api.lib, "ncclCommDestroy");
        api.errstr = (nccl_errstr_t)dlsym(api.lib, "ncclGetErrorString");
        if (!api.get_uid || !api.init_rank || !api.allgather || !api.allreduce || !api.destroy) api.lib = nullptr;
    });
    return api.lib ? &api : nullptr;
}
int fail_nccl(NcclApi *A, int rc, const char *what) {
    return fail(QMCB_ERR_NCCL, std::string("NCCL error in ") + what + ": " + (A && A->errstr ? A->errstr(rc) : "?"));
}
}  // namespace

static void pt_comm_release(QmcbHandle *h) {
    if (h->comm && h->own_comm && nccl_api()) nccl_api()->destroy(h->comm);
    h->comm = nullptr, h->own_comm = false;
}
extern "C" int qmcb_pt_comm_unique_id(uint8_t *id128) {
    if (!id128) return fail(QMCB_ERR_BAD_ARG, "null argument");
    NcclApi *A = nccl_api();
    if (!A) return fail(QMCB_ERR_NCCL, "libnccl.so.2 not found (set QMCB_NCCL_LIB)");
    NcclUid u;
    int rc = A->get_uid(&u);
    if (rc) return fail_nccl(A, rc, "ncclGetUniqueId");
    memcpy(id128, u.internal, 128);
    return QMCB_OK;
}
extern "C" int qmcb_pt_comm_init(QmcbHandle *h, const uint8_t *id128, int nranks, int rank) {
    CHECK_H(h);
    if (!id128 || nranks < 1 || rank < 0 || rank >= nranks) return fail(QMCB_ERR_BAD_ARG, "bad communicator shape");
    if (h->comm) return fail(QMCB_ERR_BAD_ARG, "a communicator is already attached");
    NcclApi *A = nccl_api();
    if (!A) return fail(QMCB_ERR_NCCL, "libnccl.so.2 not found (set QMCB_NCCL_LIB)");
    NcclUid u;
    memcpy(u.internal, id128, 128);
    void *comm = nullptr;
    int rc = A->init_rank(&comm, nranks, u, rank);
    if (rc) return fail_nccl(A, rc, "ncclCommInitRank");
    h->comm = comm, h->own_comm = true, h->comm_rank = rank, h->comm_nranks = nranks;
    return QMCB_OK;
}
extern "C" int qmcb_pt_comm_attach(QmcbHandle *h, void *nccl_comm, int nranks, int rank) {
    CHECK_H(h);
    if (h->comm && h->own_comm) return fail(QMCB_ERR_BAD_ARG, "the handle owns a communicator already");
    if (nccl_comm && !nccl_api()) return fail(QMCB_ERR_NCCL, "libnccl.so.2 not found (set QMCB_NCCL_LIB)");
    h->comm = nccl_comm, h->own_comm = false, h->comm_rank = rank, h->comm_nranks = nccl_comm ? nranks : 1;
    return QMCB_OK;
}
// tempering_step / parallel_tempering_step (tempering_container.rs:121-149, :373-402) wherever the ladder lives: on this
// handle alone (export + apply), or spread over the ranks of the attached communicator (export, ONE ncclAllGather of
// the per-configuration records, apply).  Everything is enqueued on the handle's stream; nothing synchronises the host.
extern "C" int qmcb_pt_step(QmcbHandle *h) {
    CHECK_H(h);
    if (!h->pt_on) return fail(QMCB_ERR_BAD_ARG, "tempering not configured");
    if (h->P.cfg_begin == 0 && h->D.R == h->pt_S) return qmcb_pt_step_local(h);
    if (!h->comm) return fail(QMCB_ERR_BAD_ARG, "the ladder is spread over several handles: attach a communicator (qmcb_pt_comm_init) or use qmcb_pt_export + all-gather + qmcb_pt_apply");
    if ((uint64_t)h->D.R * (uint64_t)h->comm_nranks != h->pt_S || h->P.cfg_begin != (uint32_t)h->comm_rank * h->D.R)
        return fail(QMCB_ERR_BAD_ARG, "slots must be block-partitioned evenly over the ranks of the communicator");
    NcclApi *A = nccl_api();
    uint32_t words = 0;
    qmcb_pt_record_words(h, &words);
    if (!h->pt_rec_dev) CUDA_TRY(h->pool.alloc(&h->pt_rec_dev, (size_t)h->pt_S * PT_REC_WORDS_MH));
    if (!h->pt_rec_all_dev) CUDA_TRY(h->pool.alloc(&h->pt_rec_all_dev, (size_t)h->pt_S * PT_REC_WORDS_MH));
    int rc = qmcb_pt_export(h, h->pt_rec_dev);
    if (rc) return rc;
    int nrc = A->allgather(h->pt_rec_dev, h->pt_rec_all_dev, (size_t)h->D.R * words, NCCL_UINT64, h->comm, h->stream);
    if (nrc) return fail_nccl(A, nrc, "ncclAllGather");
    return qmcb_pt_apply(h, h->pt_rec_all_dev, h->pt_S);
}
extern "C" int qmcb_pt_collective_bytes(const QmcbHandle *h, uint64_t *bytes_per_step) {
    if (!h || !bytes_per_step) return fail(QMCB_ERR_BAD_ARG, "null argument");
    if (!h->pt_on) return fail(QMCB_ERR_BAD_ARG, "tempering not configured");
    uint32_t words = 0;
    qmcb_pt_record_words(h, &words);
    *bytes_per_step = (h->P.cfg_begin == 0 && h->D.R == h->pt_S) ? 0 : (uint64_t)h->pt_S * words * 8;
    return QMCB_OK;
}

static int timesteps_impl(QmcbHandle *h, uint64_t t, uint64_t freq, double *energy_out, uint8_t *samples_out, bool keep_device_samples);
// TemperingContainer::timesteps_sample / parallel_timesteps_sample (tempering_container.rs:166-208, :411-453)
extern "C" int qmcb_pt_timesteps_sample(QmcbHandle *h, uint64_t timesteps, uint64_t replica_swap_freq, uint64_t sampling_freq,
                                        double *energy_acc, uint8_t *samples_out, uint32_t *sample_slots_out) {
    CHECK_H(h);
    if (!h->pt_on) return fail(QMCB_ERR_BAD_ARG, "tempering not configured");
    if (!energy_acc || replica_swap_freq == 0 || sampling_freq == 0) return fail(QMCB_ERR_BAD_ARG, "null energies or zero frequency");
    const SseDev &D = h->D;
    const uint32_t S = h->pt_S, R = D.R;
    const uint64_t T = timesteps / sampling_freq;
    const bool spread = !(h->P.cfg_begin == 0 && R == S);
    if (spread && !h->comm) return fail(QMCB_ERR_BAD_ARG, "the ladder is spread over several handles: attach a communicator first");
    NcclApi *A = spread ? nccl_api() : nullptr;
    std::vector<double> e(R), eseg(S);
    std::vector<uint32_t> slots(R);
    std::fill(energy_acc, energy_acc + S, 0.0);
    uint64_t remaining = timesteps, to_swap = replica_swap_freq, to_sample = sampling_freq, k = 0;
    int rc;
    while (remaining > 0) {
        const uint64_t t = std::min(std::min(to_sample, to_swap), remaining);
        if ((rc = qmcb_pt_get_slots(h, slots.data()))) return rc;
        if ((rc = timesteps_impl(h, t, 1, e.data(), nullptr, false))) return rc;  // g.timesteps(t, beta): mean energy of the t sweeps
        std::fill(eseg.begin(), eseg.end(), 0.0);
        for (uint32_t r = 0; r < R; r++) eseg[slots[r]] = e[r];
        if (spread) {  // every slot has exactly one owner: the sum is exact, and the accumulation below stays in slot order on every rank
            if (!h->pt_energy_dev) CUDA_TRY(h->pool.alloc(&h->pt_energy_dev, (size_t)S));
            CUDA_TRY(cudaMemcpyAsync(h->pt_energy_dev, eseg.data(), sizeof(double) * S, cudaMemcpyHostToDevice, h->stream));
            int nrc = A->allreduce(h->pt_energy_dev, h->pt_energy_dev, S, NCCL_FLOAT64, NCCL_SUM, h->comm, h->stream);
            if (nrc) return fail_nccl(A, nrc, "ncclAllReduce");
            CUDA_TRY(cudaMemcpyAsync(eseg.data(), h->pt_energy_dev, sizeof(double) * S, cudaMemcpyDeviceToHost, h->stream));
            CUDA_TRY(cudaStreamSynchronize(h->stream));
        }
        for (uint32_t s = 0; s < S; s++) energy_acc[s] += eseg[s] * (double)t;  // *e += te * t as f64
        to_sample -= t, to_swap -= t, remaining -= t;
        if (to_swap == 0) {
            if ((rc = qmcb_pt_step(h))) return rc;
            to_swap = replica_swap_freq;
        }
        if (to_sample == 0) {
            if (k < T) {
                if (sample_slots_out) {
                    if ((rc = qmcb_pt_get_slots(h, slots.data()))) return rc;
                    for (uint32_t r = 0; r < R; r++) sample_slots_out[(size_t)r * T + k] = slots[r];
                }
                if (samples_out) {
                    std::vector<uint32_t> packed((size_t)R * D.Nw);
                    CUDA_TRY(cudaStreamSynchronize(h->stream));
                    CUDA_TRY(cudaMemcpy(packed.data(), D.state, packed.size() * 4, cudaMemcpyDeviceToHost));
                    for (uint32_t r = 0; r < R; r++)
                        for (uint32_t v = 0; v < D.N; v++)
                            samples_out[((size_t)r * T + k) * D.N + v] = (packed[(size_t)r * D.Nw + (v >> 5)] >> (v & 31)) & 1u;
                }
                k++;
            }
            to_sample = sampling_freq;
        }
    }
    return QMCB_OK;
}
extern "C" int qmcb_pt_total_swaps(QmcbHandle *h, uint64_t *swaps) {
    CHECK_H(h);
    if (!h->pt_on || !swaps) return fail(QMCB_ERR_BAD_ARG, "tempering not configured");
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    CUDA_TRY(cudaMemcpy(swaps, h->P.swaps, sizeof(uint64_t), cudaMemcpyDeviceToHost));
    return QMCB_OK;
}
extern "C" int qmcb_pt_get_config(const QmcbHandle *h, uint32_t *n_chains, uint32_t *n_betas, uint32_t *slot_begin) {
    if (!h || !n_chains || !n_betas || !slot_begin) return fail(QMCB_ERR_BAD_ARG, "null argument");
    if (!h->pt_on) return fail(QMCB_ERR_BAD_ARG, "tempering not configured");
    *n_chains = h->P.n_chains, *n_betas = h->P.n_betas, *slot_begin = h->P.cfg_begin;
    return QMCB_OK;
}
extern "C" int qmcb_pt_get_slots(QmcbHandle *h, uint32_t *slots) {
    CHECK_H(h);
    if (!h->pt_on || !slots) return fail(QMCB_ERR_BAD_ARG, "tempering not configured");
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    CUDA_TRY(cudaMemcpy(slots, h->P.slot_of_local, sizeof(uint32_t) * h->D.R, cudaMemcpyDeviceToHost));
    return QMCB_OK;
}

// ---- checkpoints ---------------------------------------------------------
// The reference serialises a graph without its rng (SerializeQmcGraph, qmc_ising.rs:1001-1087) and a
// tempering container as (graph, beta) pairs + total_swaps (tempering_container.rs:671-793).  Here the
// injected stream is (key, cursor), 16 bytes, so it is part of the record and a restored batch continues
// bit-identically.  Layout (little-endian), all arrays replica-major:
//   "QMCBCKP1" | u32 version, N, E, R, mode, flags(bit0 heat-bath, bit1 tempering, bit2 Hamiltonian table, bit3 run_rvb_steps) | u64 target |
//   f64 transverse, longitudinal | va[E] u32 | vb[E] u32 | pad to 8 | J[E] f64 |
//   beta[R] f64 | key[R] u64 | cursor[R] u64 | done[R] u64 | vupd[R] u64 | M[R] u32 | n[R] u32 |
//   state[R][Nw] u32 | pad to 8 | ops of replica 0 (M[0] words), replica 1, ... | pad to 8 |
//   tempering block (if flagged): u32 n_chains, n_betas, cfg_begin, pad | u64 pt_key, pt_cursor, swaps |
//   beta_slot[S] f64 | key_slot[S] u64 | slot_of_local[R] u32 | pad to 8 |
//   Hamiltonian block (if flagged): u32 H, has_slot_table | J_tab[H][E] f64 | transverse[H] | longitudinal[H] |
//   ham[R] u32 | ham_slot[S] u32 (if has_slot_table) | pad to 8 |  u64 FNV-1a of all bytes before.
namespace {
struct CkWriter {
    uint8_t *p;
    uint64_t off = 0, cap;
    bool dry;
    CkWriter(void *buf, uint64_t cap_) : p((uint8_t *)buf), cap(cap_), dry(buf == nullptr) {}
    void put(const void *src, uint64_t nbytes) {
        if (!dry && off + nbytes <= cap) memcpy(p + off, src, nbytes);
        off += nbytes;
    }
    template <typename T>
    void val(T v) { put(&v, sizeof(T)); }
    void pad8() {
        const uint64_t z = 0;
        if (off % 8) put(&z, 8 - off % 8);
    }
};
struct CkReader {
    const uint8_t *p;
    uint64_t off = 0, cap;
    bool ok = true;
    CkReader(const void *buf, uint64_t cap_) : p((const uint8_t *)buf), cap(cap_) {}
    void get(void *dst, uint64_t nbytes) {
        if (!ok || off + nbytes > cap) { ok = false; return; }
        memcpy(dst, p + off, nbytes);
        off += nbytes;
    }
    template <typename T>
    T val() { T v{}; get(&v, sizeof(T)); return v; }
    template <typename T>
    std::vector<T> vec(uint64_t count) {
        std::vector<T> v;
        if (!ok || count > (cap - off) / sizeof(T)) { ok = false; return v; }
        v.resize(count);
        get(v.data(), count * sizeof(T));
        return v;
    }
    void pad8() { if (off % 8) off += 8 - off % 8; }
};
uint64_t fnv1a(const uint8_t *p, uint64_t nbytes) {
    uint64_t hsh = 1469598103934665603ull;
    for (uint64_t i = 0; i < nbytes; i++) hsh = (hsh ^ p[i]) * 1099511628211ull;
    return hsh;
}
template <typename T>
cudaError_t fetch(std::vector<T> &dst, const T *dev, size_t count) {
    dst.resize(count);
    return cudaMemcpy(dst.data(), dev, sizeof(T) * count, cudaMemcpyDeviceToHost);
}
}  // namespace

static int checkpoint_write(QmcbHandle *h, CkWriter &W) {
    const SseDev &D = h->D;
    if (h->generic) return fail(QMCB_ERR_UNSUPPORTED, "checkpoints of handles with generic interactions are not offered");
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    std::vector<double> beta;
    std::vector<uint64_t> key, cursor, done;
    std::vector<unsigned long long> vupd;
    std::vector<uint32_t> M, n, state;
    CUDA_TRY(fetch(beta, D.beta, D.R));
    CUDA_TRY(fetch(key, D.key, D.R));
    CUDA_TRY(fetch(cursor, D.cursor, D.R));
    CUDA_TRY(fetch(done, D.done, D.R));
    CUDA_TRY(fetch(vupd, D.vupd, D.R));
    CUDA_TRY(fetch(M, D.M, D.R));
    CUDA_TRY(fetch(n, D.n, D.R));
    CUDA_TRY(fetch(state, D.state, (size_t)D.R * D.Nw));
    W.put("QMCBCKP1", 8);
    W.val<uint32_t>(1), W.val<uint32_t>(D.N), W.val<uint32_t>(D.E), W.val<uint32_t>(D.R);
    W.val<uint32_t>((uint32_t)h->mode), W.val<uint32_t>((D.hb_cum ? 1u : 0u) | (h->pt_on ? 2u : 0u) | (D.ham ? 4u : 0u) | (h->rvb_on ? 8u : 0u));
    W.val<uint64_t>(h->target);
    W.val<double>(D.gamma), W.val<double>(D.h);
    W.put(h->va_h.data(), 4ull * D.E), W.put(h->vb_h.data(), 4ull * D.E), W.pad8();
    W.put(h->J_h.data(), 8ull * D.E);
    W.put(beta.data(), 8ull * D.R), W.put(key.data(), 8ull * D.R), W.put(cursor.data(), 8ull * D.R);
    W.put(done.data(), 8ull * D.R), W.put(vupd.data(), 8ull * D.R);
    W.put(M.data(), 4ull * D.R), W.put(n.data(), 4ull * D.R);
    W.put(state.data(), 4ull * D.R * D.Nw), W.pad8();
    for (uint32_t r = 0; r < D.R; r++) {
        const uint64_t have = std::min<uint64_t>(M[r], D.cap);
        if (W.dry || W.off + 4ull * M[r] > W.cap) {
            W.off += 4ull * M[r];
            continue;
        }
        CUDA_TRY(cudaMemcpy(W.p + W.off, D.ops + (size_t)r * D.cap, have * 4, cudaMemcpyDeviceToHost));
        for (uint64_t q = have; q < M[r]; q++) ((uint32_t *)(W.p + W.off))[q] = QMCB_OP_EMPTY;  // cutoff raised, not yet grown
        W.off += 4ull * M[r];
    }
    W.pad8();
    if (h->pt_on) {
        const PtDev &P = h->P;
        std::vector<double> bs;
        std::vector<uint64_t> ks, one;
        std::vector<unsigned long long> sw;
        std::vector<uint32_t> sl;
        CUDA_TRY(fetch(bs, P.beta_slot, h->pt_S));
        CUDA_TRY(fetch(ks, P.key_slot, h->pt_S));
        CUDA_TRY(fetch(one, (const uint64_t *)P.pt_cursor, 1));
        CUDA_TRY(fetch(sw, (const unsigned long long *)P.swaps, 1));
        CUDA_TRY(fetch(sl, (const uint32_t *)P.slot_of_local, D.R));
        W.val<uint32_t>(P.n_chains), W.val<uint32_t>(P.n_betas), W.val<uint32_t>(P.cfg_begin), W.val<uint32_t>(0);
        W.val<uint64_t>(P.pt_key), W.val<uint64_t>(one[0]), W.val<uint64_t>(sw[0]);
        W.put(bs.data(), 8ull * h->pt_S), W.put(ks.data(), 8ull * h->pt_S);
        W.put(sl.data(), 4ull * D.R), W.pad8();
    }
    if (D.ham) {
        std::vector<uint32_t> hr;
        CUDA_TRY(fetch(hr, D.ham, D.R));
        W.val<uint32_t>(h->H), W.val<uint32_t>(h->ham_slot_h.empty() ? 0u : 1u);
        W.put(h->Jtab_h.data(), 8ull * h->H * D.E), W.put(h->gam_h.data(), 8ull * h->H), W.put(h->hl_h.data(), 8ull * h->H);
        W.put(hr.data(), 4ull * D.R);
        if (!h->ham_slot_h.empty()) W.put(h->ham_slot_h.data(), 4ull * h->ham_slot_h.size());
        W.pad8();
    }
    return QMCB_OK;
}

extern "C" int qmcb_checkpoint_size(QmcbHandle *h, uint64_t *bytes) {
    CHECK_H(h);
    if (!bytes) return fail(QMCB_ERR_BAD_ARG, "null argument");
    CkWriter W(nullptr, 0);
    int rc = checkpoint_write(h, W);
    if (rc) return rc;
    *bytes = W.off + 8;
    return QMCB_OK;
}
extern "C" int qmcb_checkpoint_save(QmcbHandle *h, void *buf, uint64_t bytes) {
    CHECK_H(h);
    if (!buf) return fail(QMCB_ERR_BAD_ARG, "null buffer");
    CkWriter W(buf, bytes);
    int rc = checkpoint_write(h, W);
    if (rc) return rc;
    if (W.off + 8 > bytes) return fail(QMCB_ERR_BAD_ARG, "checkpoint buffer too small (see qmcb_checkpoint_size)");
    const uint64_t sum = fnv1a((const uint8_t *)buf, W.off);
    memcpy((uint8_t *)buf + W.off, &sum, 8);
    return QMCB_OK;
}
extern "C" int qmcb_checkpoint_load(const void *buf, uint64_t bytes, int device, QmcbHandle **out) {
    if (!buf || !out || bytes < 64) return fail(QMCB_ERR_BAD_ARG, "null or truncated checkpoint");
    CkReader Rd(buf, bytes);
    char magic[8];
    Rd.get(magic, 8);
    if (memcmp(magic, "QMCBCKP1", 8) != 0) return fail(QMCB_ERR_BAD_ARG, "not a qmcb checkpoint");
    const uint32_t version = Rd.val<uint32_t>(), N = Rd.val<uint32_t>(), E = Rd.val<uint32_t>(), R = Rd.val<uint32_t>();
    const uint32_t mode = Rd.val<uint32_t>(), flags = Rd.val<uint32_t>();
    const uint64_t target = Rd.val<uint64_t>();
    const double gamma = Rd.val<double>(), hl = Rd.val<double>();
    if (version != 1 || R == 0 || N == 0) return fail(QMCB_ERR_BAD_ARG, "unsupported checkpoint version or empty batch");
    auto va = Rd.vec<uint32_t>(E), vb = Rd.vec<uint32_t>(E);
    Rd.pad8();
    auto J = Rd.vec<double>(E);
    auto beta = Rd.vec<double>(R);
    auto key = Rd.vec<uint64_t>(R), cursor = Rd.vec<uint64_t>(R), done = Rd.vec<uint64_t>(R), vupd = Rd.vec<uint64_t>(R);
    auto M = Rd.vec<uint32_t>(R), n = Rd.vec<uint32_t>(R);
    const uint32_t Nw = (N + 31) / 32;
    auto state = Rd.vec<uint32_t>((uint64_t)R * Nw);
    Rd.pad8();
    if (!Rd.ok) return fail(QMCB_ERR_BAD_ARG, "truncated checkpoint");
    const uint64_t ops_off = Rd.off;
    uint64_t total_ops = 0, maxM = 0;
    for (uint32_t r = 0; r < R; r++) total_ops += M[r], maxM = std::max<uint64_t>(maxM, M[r]);
    if (total_ops > (bytes - Rd.off) / 4) return fail(QMCB_ERR_BAD_ARG, "truncated checkpoint");
    Rd.off += 4 * total_ops;
    Rd.pad8();
    // tempering block
    uint32_t n_chains = 0, n_betas = 0, cfg_begin = 0;
    uint64_t pt_key = 0, pt_cursor = 0, swaps = 0;
    std::vector<double> bs;
    std::vector<uint64_t> ks;
    std::vector<uint32_t> sl;
    if (flags & 2u) {
        n_chains = Rd.val<uint32_t>(), n_betas = Rd.val<uint32_t>(), cfg_begin = Rd.val<uint32_t>();
        Rd.val<uint32_t>();
        pt_key = Rd.val<uint64_t>(), pt_cursor = Rd.val<uint64_t>(), swaps = Rd.val<uint64_t>();
        const uint64_t S = (uint64_t)n_chains * n_betas;
        bs = Rd.vec<double>(S), ks = Rd.vec<uint64_t>(S), sl = Rd.vec<uint32_t>(R);
        Rd.pad8();
    }
    uint32_t H = 0, has_slot_tab = 0;
    std::vector<double> Jtab, gtab, htab;
    std::vector<uint32_t> hrep, hslot;
    if (flags & 4u) {
        H = Rd.val<uint32_t>(), has_slot_tab = Rd.val<uint32_t>();
        Jtab = Rd.vec<double>((uint64_t)H * E), gtab = Rd.vec<double>(H), htab = Rd.vec<double>(H);
        hrep = Rd.vec<uint32_t>(R);
        if (has_slot_tab) hslot = Rd.vec<uint32_t>((uint64_t)n_chains * n_betas);
        Rd.pad8();
    }
    if (!Rd.ok || Rd.off + 8 > bytes) return fail(QMCB_ERR_BAD_ARG, "truncated checkpoint");
    uint64_t sum = 0;
    memcpy(&sum, (const uint8_t *)buf + Rd.off, 8);
    if (sum != fnv1a((const uint8_t *)buf, Rd.off)) return fail(QMCB_ERR_BAD_ARG, "checkpoint checksum mismatch");

    QmcbLattice lat{N, E, va.data(), vb.data(), J.data(), gamma, hl};
    QmcbHandle *h = nullptr;
    const uint64_t cutoff0 = std::max<uint64_t>(maxM, 1);
    int rc = qmcb_create(&lat, R, beta.data(), key.data(), cutoff0, 0, nullptr, device, &h);
    if (rc) return rc;
    SseDev &D = h->D;
    auto bail = [&](int code) {
        qmcb_destroy(h);
        return code;
    };
#define TRYL(expr)                                                              \
    do {                                                                        \
        cudaError_t e_ = (expr);                                                \
        if (e_ != cudaSuccess) return bail(fail_cuda(e_, #expr, __FILE__, __LINE__)); \
    } while (0)
    const uint8_t *op_src = (const uint8_t *)buf + ops_off;
    for (uint32_t r = 0; r < R; r++) {
        const uint32_t *wr = (const uint32_t *)op_src;
        uint32_t cnt = 0;
        for (uint32_t q = 0; q < M[r]; q++) {
            if (wr[q] == QMCB_OP_EMPTY) continue;
            if (!op_word_ok(D, wr[q])) return bail(fail(QMCB_ERR_BAD_ARG, "malformed operator word in checkpoint"));
            cnt++;
        }
        if (cnt != n[r]) return bail(fail(QMCB_ERR_BAD_ARG, "operator count in checkpoint does not match its string"));
        TRYL(cudaMemcpy(D.ops + (size_t)r * D.cap, op_src, 4ull * M[r], cudaMemcpyHostToDevice));
        op_src += 4ull * M[r];
    }
    TRYL(cudaMemcpy(D.state, state.data(), 4ull * R * Nw, cudaMemcpyHostToDevice));
    TRYL(cudaMemcpy(D.cursor, cursor.data(), 8ull * R, cudaMemcpyHostToDevice));
    TRYL(cudaMemcpy(D.done, done.data(), 8ull * R, cudaMemcpyHostToDevice));
    TRYL(cudaMemcpy(D.vupd, vupd.data(), 8ull * R, cudaMemcpyHostToDevice));
    TRYL(cudaMemcpy(D.M, M.data(), 4ull * R, cudaMemcpyHostToDevice));
    TRYL(cudaMemcpy(D.n, n.data(), 4ull * R, cudaMemcpyHostToDevice));
    h->target = target;
    if ((flags & 4u) && (rc = qmcb_set_hamiltonians(h, H, Jtab.data(), gtab.data(), htab.data(), hrep.data()))) return bail(rc);
    if ((rc = qmcb_set_mode(h, (int)mode))) return bail(rc);
    if ((flags & 1u) && (rc = qmcb_set_enable_heatbath(h, 1))) return bail(rc);
    if ((flags & 8u) && (rc = qmcb_set_run_rvb(h, 1))) return bail(rc);  // run_rvb_steps is part of the reference's serde mirror (qmc_ising.rs:1020)
    if (flags & 2u) {
        // slot labels index the per-slot tables of the swap kernels: every one must be a slot of the ladder, no two alike
        const uint64_t S = (uint64_t)n_chains * n_betas;
        if ((uint64_t)cfg_begin + R > S) return bail(fail(QMCB_ERR_BAD_ARG, "checkpoint: configurations do not fit the ladder"));
        std::vector<uint8_t> taken(S, 0);
        for (uint32_t r = 0; r < R; r++) {
            if (sl[r] >= S || taken[sl[r]]) return bail(fail(QMCB_ERR_BAD_ARG, "checkpoint: slot label out of range or duplicated"));
            taken[sl[r]] = 1;
        }
        if (has_slot_tab)
            for (uint32_t r = 0; r < R; r++)
                if (hrep[r] >= H) return bail(fail(QMCB_ERR_BAD_ARG, "checkpoint: Hamiltonian index out of range"));
        if ((rc = qmcb_pt_configure(h, n_chains, n_betas, cfg_begin, bs.data(), ks.data(), pt_key))) return bail(rc);
        // pt_configure relabelled the replicas with the initial slots: restore the saved labels
        TRYL(cudaMemcpy(D.beta, beta.data(), 8ull * R, cudaMemcpyHostToDevice));
        TRYL(cudaMemcpy(D.key, key.data(), 8ull * R, cudaMemcpyHostToDevice));
        TRYL(cudaMemcpy(h->P.slot_of_local, sl.data(), 4ull * R, cudaMemcpyHostToDevice));
        TRYL(cudaMemcpy(h->P.pt_cursor, &pt_cursor, 8, cudaMemcpyHostToDevice));
        TRYL(cudaMemcpy(h->P.swaps, &swaps, 8, cudaMemcpyHostToDevice));
        if (has_slot_tab) {
            if ((rc = qmcb_pt_set_slot_hamiltonians(h, hslot.data()))) return bail(rc);
            TRYL(cudaMemcpy((void *)D.ham, hrep.data(), 4ull * R, cudaMemcpyHostToDevice));
        }
    }
#undef TRYL
    *out = h;
    return QMCB_OK;
}

// =============================================================================================
// classical handle
// =============================================================================================
struct CmcbHandle {
    int device = 0;
    cudaStream_t stream = nullptr, own_stream = nullptr;
    ClsDev D{};
    Pool pool;
    bool square = false;
    uint64_t sweeps = 0, launches = 0;
    uint32_t ncolours = 0;
    std::vector<uint32_t> colours, colour_start;
    double J_uniform = 0.0, bias_uniform = 0.0;
    double *adj_j_dev = nullptr, *biases_dev = nullptr;
    uint8_t *bytes_dev = nullptr;  // staging for the square layout
    // fused multi-sweep launches of the square layout: per-replica barrier counters and how far they have advanced
    unsigned int *bar_dev = nullptr;
    unsigned int bar_count = 0;
    int nsm = 148;
    int fused = 1;  // cmcb_set_option("fused", 0): one launch per colour pass (round-1 behaviour, A/B measurements)
    // reference-schedule moves (classical_ref.cu): host copies of what the reference's GraphState holds, uploaded on first use
    std::vector<uint32_t> h_start, h_idx, h_ea, h_eb;
    std::vector<double> h_aj, h_bias, h_ej, h_beta;
    uint64_t cursor0 = 0;  // words the constructor drew (make_random_spin_state: one per spin)
    ClsRefDev Rf{};
    bool ref_ready = false;
    double *cum_dev = nullptr;
    uint8_t *choice_dev = nullptr;
};

#define CHECK_C(h)                                              \
    if (!(h)) return fail(QMCB_ERR_BAD_ARG, "null handle");     \
    CUDA_TRY(cudaSetDevice((h)->device))

// #{d in [0, 2^32) : d * 2^-32 < exp(-beta * dE)}, 2^32 when dE <= 0 (should_flip, graph.rs:339-347)
static uint64_t metropolis_threshold(double beta, double delta_e) {
    if (!(delta_e > 0.0)) return 4294967296ull;
    double chance = std::exp(-beta * delta_e);
    double scaled = std::ceil(chance * 4294967296.0);
    if (scaled >= 4294967296.0) return 4294967296ull;
    return (uint64_t)scaled;
}

extern "C" int cmcb_create(const QmcbLattice *lat, const double *biases, uint32_t R, const double *betas, const uint64_t *keys,
                           const uint8_t *init_state, int device, CmcbHandle **out) {
    if (!lat || !biases || !betas || !keys || !out || R == 0 || lat->nvars == 0) return fail(QMCB_ERR_BAD_ARG, "null argument");
    if (R > 65535) return fail(QMCB_ERR_UNSUPPORTED, "at most 65535 replicas per handle (the per-replica kernels index them with gridDim.y)");
    const uint32_t N = lat->nvars, E = lat->nedges;
    for (uint32_t e = 0; e < E; e++)
        if (lat->va[e] >= N || lat->vb[e] >= N || lat->va[e] == lat->vb[e]) return fail(QMCB_ERR_BAD_ARG, "edge endpoint out of range or self-loop");
    int ndev = 0;
    CUDA_TRY(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(QMCB_ERR_CUDA, "no such CUDA device");
    CUDA_TRY(cudaSetDevice(device));
    // binding_mat (graph.rs:62-78): push order, then stable sort by neighbour index
    std::vector<std::vector<std::pair<uint32_t, double>>> adj(N);
    for (uint32_t e = 0; e < E; e++) {
        adj[lat->va[e]].push_back({lat->vb[e], lat->J[e]});
        adj[lat->vb[e]].push_back({lat->va[e], lat->J[e]});
    }
    uint32_t maxdeg = 0;
    for (auto &v : adj) {
        std::stable_sort(v.begin(), v.end(), [](const std::pair<uint32_t, double> &a, const std::pair<uint32_t, double> &b) { return a.first < b.first; });
        maxdeg = std::max<uint32_t>(maxdeg, (uint32_t)v.size());
    }
    if (maxdeg > 12) return fail(QMCB_ERR_UNSUPPORTED, "checkerboard kernel supports vertex degree <= 12");
    // greedy colouring in index order: smallest colour unused by lower-indexed neighbours
    std::vector<uint32_t> colour(N, 0);
    uint32_t ncol = 0;
    for (uint32_t i = 0; i < N; i++) {
        uint32_t used = 0;
        for (auto &nb : adj[i])
            if (nb.first < i) used |= 1u << colour[nb.first];
        uint32_t c = 0;
        while (used & (1u << c)) c++;
        colour[i] = c;
        ncol = std::max(ncol, c + 1);
    }
    CmcbHandle *h = new CmcbHandle();
    h->device = device;
    h->colours = colour, h->ncolours = ncol;
    h->h_start.assign(N + 1, 0);
    for (uint32_t i = 0; i < N; i++) {
        h->h_start[i + 1] = h->h_start[i] + (uint32_t)adj[i].size();
        for (auto &nb : adj[i]) h->h_idx.push_back(nb.first), h->h_aj.push_back(nb.second);
    }
    h->h_ea.assign(lat->va, lat->va + E), h->h_eb.assign(lat->vb, lat->vb + E), h->h_ej.assign(lat->J, lat->J + E);
    h->h_bias.assign(biases, biases + N), h->h_beta.assign(betas, betas + R);
    h->cursor0 = init_state ? 0 : N;
    ClsDev &D = h->D;
    D.N = N, D.R = R;
#define TRYC(expr)                                   \
    do {                                             \
        cudaError_t e_ = (expr);                     \
        if (e_ != cudaSuccess) {                     \
            h->pool.release_all();                   \
            delete h;                                \
            return fail_cuda(e_, #expr, __FILE__, __LINE__); \
        }                                            \
    } while (0)
    TRYC(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
    h->stream = h->own_stream;
    uint64_t *key_dev;
    TRYC(h->pool.alloc(&key_dev, R));
    TRYC(cudaMemcpy(key_dev, keys, sizeof(uint64_t) * R, cudaMemcpyHostToDevice));
    D.key = key_dev;

    // ---- is this the L x L periodic square lattice with uniform J and bias? -----------------
    uint32_t L = (uint32_t)std::llround(std::sqrt((double)N));
    bool square = (uint64_t)L * L == N && L % 64 == 0 && E == 2 * N && ncol == 2;
    if (square) {
        for (uint32_t e = 1; e < E && square; e++) square = lat->J[e] == lat->J[0];
        for (uint32_t i = 1; i < N && square; i++) square = biases[i] == biases[0];
        std::vector<uint8_t> seen(2 * (size_t)N, 0);
        for (uint32_t e = 0; e < E && square; e++) {
            uint32_t a = lat->va[e], b = lat->vb[e];
            bool hit = false;
            for (int swap = 0; swap < 2 && !hit; swap++) {
                uint32_t x = a % L, y = a / L;
                uint32_t right = y * L + (x + 1) % L, down = ((y + 1) % L) * L + x;
                if (b == right && !seen[2 * (size_t)a]) seen[2 * (size_t)a] = 1, hit = true;
                else if (b == down && !seen[2 * (size_t)a + 1]) seen[2 * (size_t)a + 1] = 1, hit = true;
                std::swap(a, b);
            }
            square = hit;
        }
        for (uint32_t i = 0; i < N && square; i++) square = colour[i] == ((i % L + i / L) & 1u);
    }
    // delta_e must depend only on (#anti-aligned neighbours, own spin): check every sign sequence
    std::vector<double> sq_de(16, 0.0);
    if (square) {
        const double J = lat->J[0], b = biases[0];
        if (square) {
            for (int own = 0; own < 2; own++)
                for (int cnt = 0; cnt <= 4; cnt++) {
                    int mask = (0xF << cnt) & 0xF;
                    double de = 0.0;
                    for (int k = 0; k < 4; k++) de += -2.0 * J * (((mask >> k) & 1) ? 1.0 : -1.0);
                    sq_de[own * 8 + cnt] = de + (2.0 * b * (own ? 1.0 : -1.0));
                }
            for (int own = 0; own < 2 && square; own++)
                for (int mask = 0; mask < 16 && square; mask++) {
                    double de = 0.0;
                    for (int k = 0; k < 4; k++) de += -2.0 * J * (((mask >> k) & 1) ? 1.0 : -1.0);
                    de = de + (2.0 * b * (own ? 1.0 : -1.0));
                    square = de == sq_de[own * 8 + (4 - __builtin_popcount(mask))];
                }
        }
    }
    h->square = square;
    h->J_uniform = E ? lat->J[0] : 0.0, h->bias_uniform = biases[0];

    std::vector<uint8_t> init_bytes;
    if (square) {
        D.L = L;
        TRYC(h->pool.alloc(&D.planes, (size_t)R * 2 * L * (L >> 6)));
        std::vector<uint32_t> gT((size_t)R * 10, 0u), gmem((size_t)R * 10, 0u), alw(R, 0u);
        uint32_t maxg = 0;
        for (uint32_t r = 0; r < R; r++) {
            uint32_t ng = 0;
            for (int own = 0; own < 2; own++)
                for (int cnt = 0; cnt <= 4; cnt++) {
                    const uint32_t idx = own * 8 + cnt;
                    const uint64_t T = metropolis_threshold(betas[r], sq_de[idx]);
                    if (T >= 4294967296ull) alw[r] |= 1u << idx;
                    else if (T > 0) {  // classes with the same threshold share a group
                        uint32_t g = 0;
                        while (g < ng && gT[(size_t)r * 10 + g] != (uint32_t)T) g++;
                        if (g == ng) gT[(size_t)r * 10 + ng++] = (uint32_t)T;
                        gmem[(size_t)r * 10 + g] |= 1u << idx;
                    }
                }
            maxg = std::max(maxg, ng);
        }
        uint32_t *gT_dev, *gmem_dev, *alw_dev;
        TRYC(h->pool.alloc(&gT_dev, gT.size()));
        TRYC(h->pool.alloc(&gmem_dev, gmem.size()));
        TRYC(h->pool.alloc(&alw_dev, alw.size()));
        TRYC(cudaMemcpy(gT_dev, gT.data(), gT.size() * 4, cudaMemcpyHostToDevice));
        TRYC(cudaMemcpy(gmem_dev, gmem.data(), gmem.size() * 4, cudaMemcpyHostToDevice));
        TRYC(cudaMemcpy(alw_dev, alw.data(), alw.size() * 4, cudaMemcpyHostToDevice));
        D.sq_gT = gT_dev, D.sq_gmem = gmem_dev, D.sq_always = alw_dev, D.sq_ngroups = maxg;
        D.sq_wpr_shift = -1;
        for (int sh = 0; sh < 20; sh++)
            if ((L >> 6) == (1u << sh)) D.sq_wpr_shift = sh;
        TRYC(h->pool.alloc(&h->bytes_dev, (size_t)R * N));
    } else {
        // CSR + site classes + per-replica threshold tables
        std::vector<uint32_t> start(N + 1, 0), idx;
        std::vector<double> aj;
        for (uint32_t i = 0; i < N; i++) {
            start[i + 1] = start[i] + (uint32_t)adj[i].size();
            for (auto &nb : adj[i]) idx.push_back(nb.first), aj.push_back(nb.second);
        }
        std::map<std::pair<std::vector<double>, double>, uint32_t> classes;
        std::vector<uint32_t> site_class(N), class_off;
        std::vector<std::pair<std::vector<double>, double>> class_def;
        uint32_t stride = 0;
        for (uint32_t i = 0; i < N; i++) {
            std::vector<double> js;
            for (auto &nb : adj[i]) js.push_back(nb.second);
            auto keyc = std::make_pair(js, biases[i]);
            auto it = classes.find(keyc);
            if (it == classes.end()) {
                it = classes.insert({keyc, (uint32_t)class_def.size()}).first;
                class_def.push_back(keyc);
                class_off.push_back(stride);
                stride += 2u << js.size();
            }
            site_class[i] = it->second;
        }
        if ((uint64_t)stride * R > (1ull << 26)) {
            h->pool.release_all();
            delete h;
            return fail(QMCB_ERR_UNSUPPORTED, "too many distinct site classes for the checkerboard threshold tables");
        }
        std::vector<unsigned long long> thr((size_t)stride * R);
        for (uint32_t r = 0; r < R; r++)
            for (size_t c = 0; c < class_def.size(); c++) {
                const auto &js = class_def[c].first;
                const uint32_t deg = (uint32_t)js.size();
                for (uint32_t own = 0; own < 2; own++)
                    for (uint32_t mask = 0; mask < (1u << deg); mask++) {
                        double de = 0.0;  // graph.rs:101-113
                        for (uint32_t k = 0; k < deg; k++) de += -2.0 * js[k] * (((mask >> k) & 1u) ? 1.0 : -1.0);
                        de = de + (2.0 * class_def[c].second * (own ? 1.0 : -1.0));
                        thr[(size_t)r * stride + class_off[c] + ((own << deg) | mask)] = metropolis_threshold(betas[r], de);
                    }
            }
        std::vector<uint32_t> csites, cstart(ncol + 1, 0);
        for (uint32_t c = 0; c < ncol; c++) {
            for (uint32_t i = 0; i < N; i++)
                if (colour[i] == c) csites.push_back(i);
            cstart[c + 1] = (uint32_t)csites.size();
        }
        h->colour_start = cstart;
        uint32_t *d_start, *d_idx, *d_class, *d_off, *d_csites;
        unsigned long long *d_thr;
        TRYC(h->pool.alloc(&d_start, start.size()));
        TRYC(h->pool.alloc(&d_idx, idx.size()));
        TRYC(h->pool.alloc(&d_class, site_class.size()));
        TRYC(h->pool.alloc(&d_off, class_off.size()));
        TRYC(h->pool.alloc(&d_csites, csites.size()));
        TRYC(h->pool.alloc(&d_thr, thr.size()));
        TRYC(h->pool.alloc(&h->adj_j_dev, aj.size()));
        TRYC(h->pool.alloc(&h->biases_dev, N));
        TRYC(cudaMemcpy(d_start, start.data(), start.size() * 4, cudaMemcpyHostToDevice));
        TRYC(cudaMemcpy(d_idx, idx.data(), idx.size() * 4, cudaMemcpyHostToDevice));
        TRYC(cudaMemcpy(d_class, site_class.data(), site_class.size() * 4, cudaMemcpyHostToDevice));
        TRYC(cudaMemcpy(d_off, class_off.data(), class_off.size() * 4, cudaMemcpyHostToDevice));
        TRYC(cudaMemcpy(d_csites, csites.data(), csites.size() * 4, cudaMemcpyHostToDevice));
        TRYC(cudaMemcpy(d_thr, thr.data(), thr.size() * 8, cudaMemcpyHostToDevice));
        TRYC(cudaMemcpy(h->adj_j_dev, aj.data(), aj.size() * 8, cudaMemcpyHostToDevice));
        TRYC(cudaMemcpy(h->biases_dev, biases, N * 8, cudaMemcpyHostToDevice));
        D.adj_start = d_start, D.adj_idx = d_idx, D.site_class = d_class, D.class_off = d_off;
        D.colour_sites = d_csites, D.thr = d_thr, D.thr_stride = stride;
        TRYC(h->pool.alloc(&D.spins, (size_t)R * N));
    }
    // initial spins
    uint8_t *bytes = square ? h->bytes_dev : D.spins;
    if (init_state) TRYC(cudaMemcpy(bytes, init_state, (size_t)R * N, cudaMemcpyHostToDevice));
    else {
        launch_cls_init_bytes(D, bytes, h->stream);
        h->launches++;
    }
    if (square) {
        launch_cls_square_pack(D, bytes, h->stream);
        h->launches++;
    }
    TRYC(cudaGetLastError());
    TRYC(cudaStreamSynchronize(h->stream));
#undef TRYC
    *out = h;
    return QMCB_OK;
}

extern "C" int cmcb_destroy(CmcbHandle *h) {
    if (!h) return QMCB_OK;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    h->pool.release_all();
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    delete h;
    return QMCB_OK;
}
extern "C" int cmcb_set_stream(CmcbHandle *h, void *s) {
    CHECK_C(h);
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    h->stream = s ? (cudaStream_t)s : h->own_stream;
    return QMCB_OK;
}
extern "C" int cmcb_set_option(CmcbHandle *h, const char *name, int64_t value) {
    if (!h || !name) return fail(QMCB_ERR_BAD_ARG, "null argument");
    if (!strcmp(name, "fused")) {
        h->fused = value != 0;
        return QMCB_OK;
    }
    return fail(QMCB_ERR_BAD_ARG, "unknown option");
}
extern "C" int cmcb_enqueue_sweeps(CmcbHandle *h, uint64_t nsweeps) {
    CHECK_C(h);
    if (h->square && h->fused && nsweeps) {
        // both colours of up to 4096 sweeps per launch; the blocks of a replica meet at a barrier between colour passes
        if (!h->bar_dev) {
            CUDA_TRY(h->pool.alloc(&h->bar_dev, h->D.R));
            CUDA_TRY(cudaMemsetAsync(h->bar_dev, 0, sizeof(unsigned int) * h->D.R, h->stream));
            cudaDeviceGetAttribute(&h->nsm, cudaDevAttrMultiProcessorCount, h->device);
        }
        uint64_t left = nsweeps;
        while (left) {
            const uint32_t chunk = (uint32_t)std::min<uint64_t>(left, 4096);
            const int nl = launch_cls_square_fused(h->D, h->sweeps, chunk, h->bar_dev, &h->bar_count, h->nsm, h->stream);
            if (nl < 0) break;  // not co-resident on this device: fall through to the per-pass launches
            h->launches += (uint64_t)nl, h->sweeps += chunk, left -= chunk;
        }
        CUDA_TRY(cudaGetLastError());
        if (!left) return QMCB_OK;
        nsweeps = left;
    }
    for (uint64_t s = 0; s < nsweeps; s++, h->sweeps++) {
        for (uint32_t c = 0; c < h->ncolours; c++) {
            if (h->square) launch_cls_square(h->D, c, h->sweeps, h->stream);
            else launch_cls_generic(h->D, c, h->colour_start[c], h->colour_start[c + 1] - h->colour_start[c], h->sweeps, h->stream);
            h->launches++;
        }
    }
    CUDA_TRY(cudaGetLastError());
    return QMCB_OK;
}
extern "C" int cmcb_synchronize(CmcbHandle *h) {
    CHECK_C(h);
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return QMCB_OK;
}
extern "C" int cmcb_sweeps(CmcbHandle *h, uint64_t nsweeps) {
    int rc = cmcb_enqueue_sweeps(h, nsweeps);
    return rc ? rc : cmcb_synchronize(h);
}
extern "C" int cmcb_get_states(CmcbHandle *h, uint8_t *states) {
    CHECK_C(h);
    if (!states) return fail(QMCB_ERR_BAD_ARG, "null argument");
    const uint8_t *src = h->D.spins;
    if (h->square) {
        launch_cls_square_unpack(h->D, h->bytes_dev, h->stream);
        h->launches++;
        src = h->bytes_dev;
    }
    CUDA_TRY(cudaMemcpyAsync(states, src, (size_t)h->D.R * h->D.N, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return QMCB_OK;
}
extern "C" int cmcb_get_state(CmcbHandle *h, uint32_t r, uint8_t *state) {
    CHECK_C(h);
    if (!state || r >= h->D.R) return fail(QMCB_ERR_BAD_ARG, "bad replica index");
    const uint8_t *src = h->D.spins;
    if (h->square) {
        launch_cls_square_unpack(h->D, h->bytes_dev, h->stream);
        h->launches++;
        src = h->bytes_dev;
    }
    CUDA_TRY(cudaMemcpyAsync(state, src + (size_t)r * h->D.N, h->D.N, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return QMCB_OK;
}
extern "C" int cmcb_set_states(CmcbHandle *h, const uint8_t *states) {
    CHECK_C(h);
    if (!states) return fail(QMCB_ERR_BAD_ARG, "null argument");
    uint8_t *dst = h->square ? h->bytes_dev : h->D.spins;
    CUDA_TRY(cudaMemcpyAsync(dst, states, (size_t)h->D.R * h->D.N, cudaMemcpyHostToDevice, h->stream));
    if (h->square) {
        launch_cls_square_pack(h->D, dst, h->stream);
        h->launches++;
    }
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return QMCB_OK;
}
static int cls_measure(CmcbHandle *h, double *energy, double *mag) {
    const ClsDev &D = h->D;
    if (h->square) {
        unsigned long long *dev = nullptr;
        CUDA_TRY(cudaMalloc(&dev, sizeof(uint64_t) * 2 * D.R));
        cudaMemsetAsync(dev, 0, sizeof(uint64_t) * 2 * D.R, h->stream);
        launch_cls_square_measure(D, dev, dev + D.R, h->stream);
        h->launches++;
        std::vector<unsigned long long> host(2 * (size_t)D.R);
        cudaError_t e = cudaMemcpyAsync(host.data(), dev, sizeof(uint64_t) * 2 * D.R, cudaMemcpyDeviceToHost, h->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
        cudaFree(dev);
        if (e != cudaSuccess) return fail_cuda(e, "measure", __FILE__, __LINE__);
        const double N = (double)D.N;
        for (uint32_t r = 0; r < D.R; r++) {
            double unsat = (double)host[r], up = (double)host[D.R + r];
            if (energy) energy[r] = h->J_uniform * (2.0 * N - 2.0 * unsat) + h->bias_uniform * (N - 2.0 * up);
            if (mag) mag[r] = (2.0 * up - N) / N;
        }
        return QMCB_OK;
    }
    double *dev = nullptr;
    CUDA_TRY(cudaMalloc(&dev, sizeof(double) * 2 * D.R));
    launch_cls_generic_energy(D, h->adj_j_dev, h->biases_dev, dev, dev + D.R, h->stream);
    h->launches++;
    std::vector<double> host(2 * (size_t)D.R);
    cudaError_t e = cudaMemcpyAsync(host.data(), dev, sizeof(double) * 2 * D.R, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(dev);
    if (e != cudaSuccess) return fail_cuda(e, "measure", __FILE__, __LINE__);
    for (uint32_t r = 0; r < D.R; r++) {
        if (energy) energy[r] = host[r];
        if (mag) mag[r] = host[D.R + r];
    }
    return QMCB_OK;
}
extern "C" int cmcb_energy(CmcbHandle *h, double *energy) {
    CHECK_C(h);
    return energy ? cls_measure(h, energy, nullptr) : fail(QMCB_ERR_BAD_ARG, "null argument");
}
extern "C" int cmcb_magnetization(CmcbHandle *h, double *m) {
    CHECK_C(h);
    return m ? cls_measure(h, nullptr, m) : fail(QMCB_ERR_BAD_ARG, "null argument");
}
// ---- the reference's own schedule: do_time_step with spin / edge / worm moves (graph.rs:91-406) --------------
static int cls_ref_ensure(CmcbHandle *h) {
    if (h->ref_ready) return QMCB_OK;
    ClsRefDev &F = h->Rf;
    const uint32_t N = h->D.N, R = h->D.R, E = (uint32_t)h->h_ea.size();
    F.N = N, F.R = R, F.E = E, F.key = h->D.key;
    uint32_t *d_start, *d_idx, *d_ea, *d_eb;
    double *d_aj, *d_bias, *d_beta;
    CUDA_TRY(h->pool.alloc(&d_start, h->h_start.size()));
    CUDA_TRY(h->pool.alloc(&d_idx, std::max<size_t>(h->h_idx.size(), 1)));
    CUDA_TRY(h->pool.alloc(&d_aj, std::max<size_t>(h->h_aj.size(), 1)));
    CUDA_TRY(h->pool.alloc(&d_ea, std::max<size_t>(E, 1)));
    CUDA_TRY(h->pool.alloc(&d_eb, std::max<size_t>(E, 1)));
    CUDA_TRY(h->pool.alloc(&d_bias, N));
    CUDA_TRY(h->pool.alloc(&d_beta, R));
    CUDA_TRY(h->pool.alloc(&F.cursor, R));
    CUDA_TRY(h->pool.alloc(&F.path, (size_t)R * 2 * ((size_t)N + 2)));
    CUDA_TRY(h->pool.alloc(&F.status, 1));
    CUDA_TRY(h->pool.alloc(&h->choice_dev, R));
    CUDA_TRY(cudaMemcpy(d_start, h->h_start.data(), h->h_start.size() * 4, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(d_idx, h->h_idx.data(), h->h_idx.size() * 4, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(d_aj, h->h_aj.data(), h->h_aj.size() * 8, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(d_ea, h->h_ea.data(), (size_t)E * 4, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(d_eb, h->h_eb.data(), (size_t)E * 4, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(d_bias, h->h_bias.data(), (size_t)N * 8, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(d_beta, h->h_beta.data(), (size_t)R * 8, cudaMemcpyHostToDevice));
    std::vector<uint64_t> cur(R, h->cursor0);
    CUDA_TRY(cudaMemcpy(F.cursor, cur.data(), (size_t)R * 8, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemset(F.status, 0, sizeof(int)));
    F.adj_start = d_start, F.adj_idx = d_idx, F.adj_j = d_aj, F.ea = d_ea, F.eb = d_eb, F.biases = d_bias, F.beta = d_beta;
    F.cum_w = nullptr, F.total_w = 0.0;
    F.spins = h->square ? h->bytes_dev : h->D.spins;
    h->ref_ready = true;
    return QMCB_OK;
}
static int cls_ref_run(CmcbHandle *h, int move, uint64_t nspin, uint64_t nedge, uint64_t nworm, int only_basic, int allow_doubles,
                       uint8_t *choices_out) {
    int rc = cls_ref_ensure(h);
    if (rc) return rc;
    const uint64_t none = ~0ull;  // None: the reference's defaults, graph.rs:361-363
    if (nspin == none) nspin = std::max<uint64_t>(1, h->D.N / 2);
    if (nedge == none) nedge = std::max<uint64_t>(1, h->h_ea.size() / 2);
    if (nworm == none) nworm = 1;
    if (h->square) {
        launch_cls_square_unpack(h->D, h->bytes_dev, h->stream);
        h->launches++;
    }
    launch_cls_ref_moves(h->Rf, move, nspin, nedge, nworm, only_basic, allow_doubles, h->choice_dev, h->stream);
    h->launches++;
    if (h->square) {
        launch_cls_square_pack(h->D, h->bytes_dev, h->stream);
        h->launches++;
    }
    CUDA_TRY(cudaGetLastError());
    int status = 0;
    CUDA_TRY(cudaMemcpyAsync(&status, h->Rf.status, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    if (choices_out) CUDA_TRY(cudaMemcpyAsync(choices_out, h->choice_dev, h->D.R, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    if (status) {
        cudaMemset(h->Rf.status, 0, sizeof(int));
        return fail(QMCB_ERR_BAD_ARG, "edge move without edges or with a non-positive total edge weight (the reference panics: empty range)");
    }
    return QMCB_OK;
}
extern "C" int cmcb_do_time_step(CmcbHandle *h, uint64_t nspinupdates, uint64_t nedgeupdates, uint64_t nwormupdates, int only_basic_moves,
                                 uint8_t *choices_out) {
    CHECK_C(h);
    return cls_ref_run(h, 3, nspinupdates, nedgeupdates, nwormupdates, only_basic_moves != 0, 1, choices_out);
}
extern "C" int cmcb_spin_flips(CmcbHandle *h, uint64_t count) {
    CHECK_C(h);
    return cls_ref_run(h, 0, count, 0, 0, 0, 0, nullptr);
}
extern "C" int cmcb_edge_flips(CmcbHandle *h, uint64_t count) {
    CHECK_C(h);
    return cls_ref_run(h, 1, 0, count, 0, 0, 0, nullptr);
}
extern "C" int cmcb_worm_flips(CmcbHandle *h, uint64_t count, int allow_doubles) {
    CHECK_C(h);
    return cls_ref_run(h, 2, 0, 0, count, 0, allow_doubles != 0, nullptr);
}
extern "C" int cmcb_enable_edge_importance_sampling(CmcbHandle *h, int enable) {
    CHECK_C(h);
    int rc = cls_ref_ensure(h);
    if (rc) return rc;
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    if (!enable) {
        h->Rf.cum_w = nullptr, h->Rf.total_w = 0.0;
        return QMCB_OK;
    }
    const size_t E = h->h_ej.size();
    std::vector<double> cum(std::max<size_t>(E, 1), 0.0);
    double acc = 0.0;  // graph.rs:323-331: running sums in edge order
    for (size_t e = 0; e < E; e++) acc = acc + h->h_ej[e], cum[e] = acc;
    if (!h->cum_dev) CUDA_TRY(h->pool.alloc(&h->cum_dev, cum.size()));
    CUDA_TRY(cudaMemcpy(h->cum_dev, cum.data(), cum.size() * 8, cudaMemcpyHostToDevice));
    h->Rf.cum_w = h->cum_dev, h->Rf.total_w = acc;
    return QMCB_OK;
}
extern "C" int cmcb_get_rng_cursors(CmcbHandle *h, uint64_t *cursors) {
    CHECK_C(h);
    if (!cursors) return fail(QMCB_ERR_BAD_ARG, "null argument");
    int rc = cls_ref_ensure(h);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(cursors, h->Rf.cursor, (size_t)h->D.R * 8, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return QMCB_OK;
}
extern "C" int cmcb_set_rng_cursors(CmcbHandle *h, const uint64_t *cursors) {
    CHECK_C(h);
    if (!cursors) return fail(QMCB_ERR_BAD_ARG, "null argument");
    int rc = cls_ref_ensure(h);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(h->Rf.cursor, cursors, (size_t)h->D.R * 8, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return QMCB_OK;
}
extern "C" int cmcb_get_colours(const CmcbHandle *h, uint32_t *colours, uint32_t *ncolours) {
    if (!h) return fail(QMCB_ERR_BAD_ARG, "null handle");
    if (colours) memcpy(colours, h->colours.data(), sizeof(uint32_t) * h->colours.size());
    if (ncolours) *ncolours = h->ncolours;
    return QMCB_OK;
}
extern "C" int cmcb_get_sweep_count(const CmcbHandle *h, uint64_t *sweeps) {
    if (!h || !sweeps) return fail(QMCB_ERR_BAD_ARG, "null argument");
    *sweeps = h->sweeps;
    return QMCB_OK;
}
extern "C" int cmcb_set_sweep_count(CmcbHandle *h, uint64_t sweeps) {
    if (!h) return fail(QMCB_ERR_BAD_ARG, "null handle");
    h->sweeps = sweeps;
    return QMCB_OK;
}
extern "C" int cmcb_layout(const CmcbHandle *h, int *sq) {
    if (!h || !sq) return fail(QMCB_ERR_BAD_ARG, "null argument");
    *sq = h->square ? 1 : 0;
    return QMCB_OK;
}
extern "C" int cmcb_launch_count(const CmcbHandle *h, uint64_t *launches) {
    if (!h || !launches) return fail(QMCB_ERR_BAD_ARG, "null argument");
    *launches = h->launches;
    return QMCB_OK;
}
