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

__global__ void k_pt_export(SseDev D, PtDev P, uint64_t *rec) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= D.R) return;
    rec[4 * (size_t)s + 0] = P.slot_of_local[s];
    rec[4 * (size_t)s + 1] = D.n[s];
    rec[4 * (size_t)s + 2] = D.cursor[s];
    rec[4 * (size_t)s + 3] = D.M[s];
}

// one block; records are indexed by global configuration id
__global__ void __launch_bounds__(1024) k_pt_apply(SseDev D, PtDev P, const uint64_t *rec, uint32_t S) {
    __shared__ unsigned int s_swaps;
    const uint32_t nb = P.n_betas;
    if (nb <= 1) return;  // tempering_container.rs:122-124
    if (threadIdx.x == 0) s_swaps = 0;
    for (uint32_t c = threadIdx.x; c < P.n_chains; c += blockDim.x) P.maxM_chain[c] = 0;
    __syncthreads();
    for (uint32_t g = threadIdx.x; g < S; g += blockDim.x) {
        uint32_t slot = (uint32_t)rec[4 * (size_t)g];
        P.n_slot[slot] = (uint32_t)rec[4 * (size_t)g + 1];
        P.cursor_slot[slot] = rec[4 * (size_t)g + 2];
        P.cfg_slot[slot] = g;
        atomicMax(&P.maxM_chain[slot / nb], (uint32_t)rec[4 * (size_t)g + 3]);  // :129-137, per ladder
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
                const double p_swap = powi_rt(__ddiv_rn(ba, bb), dn) * 1.0;  // :294 (equal Hamiltonians)
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
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        *P.pt_cursor = cur0 + 1 + nb / 2 + ((nb % 2 == 1) ? (nb - 1) / 2 : (nb - 2) / 2);
        *P.swaps += s_swaps;
    }
}

void launch_pt_export(const SseDev &D, const PtDev &P, uint64_t *rec, cudaStream_t st) {
    k_pt_export<<<(D.R + 255) / 256, 256, 0, st>>>(D, P, rec);
}
void launch_pt_apply(const SseDev &D, const PtDev &P, const uint64_t *rec, uint32_t S, cudaStream_t st) {
    k_pt_apply<<<1, 1024, 0, st>>>(D, P, rec, S);
}
