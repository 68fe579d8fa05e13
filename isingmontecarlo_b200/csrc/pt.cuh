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
};
