// pt.cu -- parallel-tempering swap step (TemperingContainer::tempering_step,
// tempering_container.rs:121-149; perform_swaps :241-260; swap_on_chunks :274-302).
//
// Configurations never move: a swap exchanges the slot LABELS (beta, rng key, rng cursor) of two
// configurations, which is the reference's swap_manager_and_state (qmc_ising.rs:593-602) seen from
// the configuration's side.  Every rank evaluates the swaps of every ladder from the gathered
// (slot, n, cursor, cutoff) records and the shared PT stream, so all ranks reach the same
// permutation without further communication; each rank then relabels its own configurations.
#include "pt.cuh"


// compiler-rt __powidf2 (what f64::powi lowers to)
__device__ __forceinline__ double powi_rt(double a, int b) {
    const bool recip = b < 0;
    double r = 1.0;
    for (;;) {
        if (b & 1) r = __dmul_rn(r, a);
        b /= 2;
        if (b == 0) break;
        a = __dmul_rn(a, a);
    }
    return recip ? __ddiv_rn(1.0, r) : r;
}

__global__ void k_pt_export(SseDev D, PtDev P, uint64_t *rec, uint32_t words) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= D.R) return;
    rec[words * (size_t)s + 0] = P.slot_of_local[s];
    rec[words * (size_t)s + 1] = D.n[s];
    rec[words * (size_t)s + 2] = D.cursor[s];
    rec[words * (size_t)s + 3] = D.M[s];
}

// partner of ladder position k in the pass over make_first_subgraphs (type 0: pairs (0,1),(2,3),..) or
// make_second_subgraphs (type 1: pairs (1,2),(3,4),..), tempering_container.rs:83-99; nb = ladder length
__device__ __forceinline__ int pt_partner(int type, uint32_t k, uint32_t nb) {
    if (type == 0) {
        const uint32_t a_pairs = nb / 2;
        return k < 2 * a_pairs ? (int)(k ^ 1u) : -1;
    }
    const uint32_t b_pairs = (nb % 2 == 1) ? (nb - 1) / 2 : (nb - 2) / 2;
    return (k >= 1 && k < 1 + 2 * b_pairs) ? (int)(((k - 1) ^ 1u) + 1) : -1;
}

// per-bond operator counts of the local configurations (fast_ops.rs:1281-1294)
__global__ void k_pt_counts(SseDev D, PtDev P) {
    const uint32_t r = blockIdx.y;
    const uint32_t *ops = D.ops + (size_t)r * D.cap;
    uint32_t *cnt = P.counts + (size_t)r * D.Nb;
    const uint32_t M = min((uint64_t)D.M[r], D.cap);
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < M; p += gridDim.x * blockDim.x) {
        const uint32_t w = ops[p];
        if (w != OP_EMPTY) atomicAdd(&cnt[op_bond(w)], 1u);
    }
}

// GraphWeights::relative_weight (tempering_traits.rs:126-154) of a configuration with bond counters cnt,
// held by a graph with Hamiltonian row hs, evaluated for the graph with row ho
__device__ double pt_relative_weight(const SseDev &D, const uint32_t *cnt, uint32_t hs, uint32_t ho) {
    const double *Jo = D.J_tab + (size_t)ho * D.E, *Js = D.J_tab + (size_t)hs * D.E;
    double bond_ratio = 1.0;  // Iterator::product: fold from 1.0, in bond order
    for (uint32_t b = 0; b < D.E; b++) bond_ratio = __dmul_rn(bond_ratio, powi_rt(__ddiv_rn(Jo[b], Js[b]), (int)cnt[b]));
    uint32_t t_count = 0;
    for (uint32_t v = 0; v < D.N; v++) t_count += cnt[D.E + v];
    const double transverse_ratio = powi_rt(__ddiv_rn(D.gam_tab[ho], D.gam_tab[hs]), (int)t_count);
    if (fabs(D.h_tab[hs]) > 2.220446049250313e-16) {
        uint32_t l_count = 0;
        for (uint32_t v = 0; v < D.N; v++) l_count += cnt[D.E + D.N + v];
        const double longitudinal_ratio = powi_rt(__ddiv_rn(D.h_tab[ho], D.h_tab[hs]), (int)l_count);
        return __dmul_rn(__dmul_rn(bond_ratio, transverse_ratio), longitudinal_ratio);
    }
    return __dmul_rn(bond_ratio, transverse_ratio);
}

// the three relative weights a configuration can need in one tempering step: against its partner in the
// first pass; against its partner in the second pass if it stayed; against the second-pass partner of the
// slot it moved to if it was swapped.  One thread per (configuration, case).
__global__ void k_pt_export_weights(SseDev D, PtDev P, uint64_t *rec) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t s = t / 3, j = t % 3;
    if (s >= D.R) return;
    const uint32_t nb = P.n_betas, slot = P.slot_of_local[s], chain = slot / nb, k = slot % nb;
    const bool a_first = stream_word(P.pt_key + chain, *P.pt_cursor) < 0x8000000000000000ull;  // :140
    const int t1 = a_first ? 0 : 1, t2 = 1 - t1;
    int from = (int)k, to;
    if (j == 0) to = pt_partner(t1, k, nb);
    else if (j == 1) to = pt_partner(t2, k, nb);
    else {
        from = pt_partner(t1, k, nb);
        to = from < 0 ? -1 : pt_partner(t2, (uint32_t)from, nb);
    }
    double wgt = 1.0;
    if (from >= 0 && to >= 0) {
        const uint32_t hs = P.ham_slot[chain * nb + (uint32_t)from], ho = P.ham_slot[chain * nb + (uint32_t)to];
        if (!P.ham_eq[hs * P.H + ho]) wgt = pt_relative_weight(D, P.counts + (size_t)s * D.Nb, hs, ho);
    }
    rec[PT_REC_WORDS_MH * (size_t)s + 4 + j] = (uint64_t)__double_as_longlong(wgt);
    if (j == 0) rec[PT_REC_WORDS_MH * (size_t)s + 7] = 0;
}

// one block; records are indexed by global configuration id
__global__ void __launch_bounds__(1024) k_pt_apply(SseDev D, PtDev P, const uint64_t *rec, uint32_t S, uint32_t words) {
    __shared__ unsigned int s_swaps;
    const uint32_t nb = P.n_betas;
    if (nb <= 1) return;  // tempering_container.rs:122-124
    if (threadIdx.x == 0) s_swaps = 0;
    for (uint32_t c = threadIdx.x; c < P.n_chains; c += blockDim.x) P.maxM_chain[c] = 0;
    __syncthreads();
    for (uint32_t g = threadIdx.x; g < S; g += blockDim.x) {
        uint32_t slot = (uint32_t)rec[words * (size_t)g];
        P.n_slot[slot] = (uint32_t)rec[words * (size_t)g + 1];
        P.cursor_slot[slot] = rec[words * (size_t)g + 2];
        P.cfg_slot[slot] = g;
        if (P.ham_slot) P.oslot_cfg[g] = slot;
        atomicMax(&P.maxM_chain[slot / nb], (uint32_t)rec[words * (size_t)g + 3]);  // :129-137, per ladder
    }
    __syncthreads();
    const uint64_t cur0 = *P.pt_cursor;
    {
        const uint32_t a_pairs = nb / 2;                                // make_first_subgraphs :83-90
        const uint32_t b_pairs = (nb % 2 == 1) ? (nb - 1) / 2 : (nb - 2) / 2;  // make_second_subgraphs :92-99
        for (int pass = 0; pass < 2; pass++) {
            for (uint32_t t = threadIdx.x; t < P.n_chains * max(a_pairs, b_pairs); t += blockDim.x) {
                const uint32_t chain = t / max(a_pairs, b_pairs), j = t % max(a_pairs, b_pairs);
                const uint64_t ckey = P.pt_key + chain;  // one TemperingContainer (own rng) per ladder
                const bool a_first = stream_word(ckey, cur0) < 0x8000000000000000ull;  // gen_bool(0.5) :140
                const bool is_a = (pass == 0) == a_first;
                const uint32_t npairs = is_a ? a_pairs : b_pairs;
                if (j >= npairs) continue;
                // draws: word 0 = order; then the pairs of the first pass, then of the second (:255)
                const uint64_t widx = cur0 + 1 + (pass == 0 ? 0 : (a_first ? a_pairs : b_pairs)) + j;
                const uint64_t v = stream_word(ckey, widx);
                const double u = __longlong_as_double((long long)((v >> 12) | 0x3FF0000000000000ull)) - 1.0;
                const uint32_t i = chain * nb + (is_a ? 0u : 1u) + 2 * j;
                const double ba = P.beta_slot[i], bb = P.beta_slot[i + 1];
                const int dn = (int)P.n_slot[i + 1] - (int)P.n_slot[i];
                double rel_h_weight = 1.0;  // :286-292
                if (P.ham_slot && !P.ham_eq[P.ham_slot[i] * P.H + P.ham_slot[i + 1]]) {
                    const uint32_t ga = P.cfg_slot[i], gb = P.cfg_slot[i + 1];
                    const uint32_t ja = pass == 0 ? 0u : (P.oslot_cfg[ga] == i ? 1u : 2u), jb = pass == 0 ? 0u : (P.oslot_cfg[gb] == i + 1 ? 1u : 2u);
                    const double rel_bstate = __longlong_as_double((long long)rec[words * (size_t)ga + 4 + ja]);  // ga.relative_weight(gb)
                    const double rel_astate = __longlong_as_double((long long)rec[words * (size_t)gb + 4 + jb]);  // gb.relative_weight(ga)
                    rel_h_weight = __dmul_rn(rel_bstate, rel_astate);
                }
                const double p_swap = __dmul_rn(powi_rt(__ddiv_rn(ba, bb), dn), rel_h_weight);  // :294-295
                if (p_swap > u) {  // :296-301
                    uint32_t tn = P.n_slot[i];
                    P.n_slot[i] = P.n_slot[i + 1], P.n_slot[i + 1] = tn;
                    uint32_t tc = P.cfg_slot[i];
                    P.cfg_slot[i] = P.cfg_slot[i + 1], P.cfg_slot[i + 1] = tc;
                    atomicAdd(&s_swaps, 1u);
                }
            }
            __syncthreads();
        }
    }
    // relabel the local configurations
    for (uint32_t slot = threadIdx.x; slot < S; slot += blockDim.x) {
        uint32_t g = P.cfg_slot[slot];
        if (g >= P.cfg_begin && g < P.cfg_begin + D.R) {
            uint32_t s = g - P.cfg_begin;
            P.slot_of_local[s] = slot;
            D.beta[s] = P.beta_slot[slot];
            D.key[s] = P.key_slot[slot];
            D.cursor[s] = P.cursor_slot[slot];
            D.M[s] = P.maxM_chain[slot / nb];
            if (P.ham_slot) ((uint32_t *)D.ham)[s] = P.ham_slot[slot];
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        *P.pt_cursor = cur0 + 1 + nb / 2 + ((nb % 2 == 1) ? (nb - 1) / 2 : (nb - 2) / 2);
        *P.swaps += s_swaps;
    }
}

// returns the number of kernels launched
int launch_pt_export(const SseDev &D, const PtDev &P, uint64_t *rec, cudaStream_t st) {
    const uint32_t words = P.ham_slot ? PT_REC_WORDS_MH : PT_REC_WORDS_EQ;
    k_pt_export<<<(D.R + 255) / 256, 256, 0, st>>>(D, P, rec, words);
    if (!P.ham_slot) return 1;
    cudaMemsetAsync(P.counts, 0, sizeof(uint32_t) * (size_t)D.R * D.Nb, st);
    k_pt_counts<<<dim3(16, D.R), 256, 0, st>>>(D, P);
    k_pt_export_weights<<<(3 * D.R + 127) / 128, 128, 0, st>>>(D, P, rec);
    return 3;
}
void launch_pt_apply(const SseDev &D, const PtDev &P, const uint64_t *rec, uint32_t S, cudaStream_t st) {
    k_pt_apply<<<1, 1024, 0, st>>>(D, P, rec, S, P.ham_slot ? PT_REC_WORDS_MH : PT_REC_WORDS_EQ);
}
