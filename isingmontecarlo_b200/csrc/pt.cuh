// pt.cuh -- device-side view of the parallel-tempering state of one SSE handle.
#pragma once
#include "sse.cuh"

struct PtDev {
    uint32_t n_chains, n_betas;   // global ladder shape; slot s = chain * n_betas + k
    uint32_t cfg_begin;           // global id of this rank's first configuration
    const double *beta_slot;      // [S] static tables
    const uint64_t *key_slot;     // [S]
    uint64_t pt_key;
    uint64_t *pt_cursor;          // [1]
    unsigned long long *swaps;    // [1]
    uint32_t *slot_of_local;      // [R] current slot of each local configuration
    // scratch [S]
    uint32_t *n_slot;
    uint64_t *cursor_slot;
    uint32_t *cfg_slot;
    uint32_t *maxM_chain;         // [n_chains]
    // unequal Hamiltonians (GraphWeights, tempering_traits.rs:122-154); ham_slot == NULL: all slots equal
    const uint32_t *ham_slot;     // [S] Hamiltonian row of each slot (a label, like beta)
    const uint8_t *ham_eq;        // [H][H] GraphWeights::ham_eq of two rows
    uint32_t H;
    uint32_t *counts;             // [R][Nb] bond counters of the local configurations (scratch)
    uint32_t *oslot_cfg;          // [S] slot each configuration held when the step began (scratch)
};
#define PT_REC_WORDS_EQ 4  // {slot, n, cursor, cutoff}
#define PT_REC_WORDS_MH 8  // ... + three relative weights (f64 bits) + pad
