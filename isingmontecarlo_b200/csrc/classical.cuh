// classical.cuh -- device-side view of one batched classical handle.
#pragma once
#include "common.cuh"

struct ClsDev {
    uint32_t N, R;
    const uint64_t *key;  // [R]
    // generic layout
    uint8_t *spins;              // [R][N] one byte per spin (reference Vec<bool>)
    const uint32_t *adj_start;   // [N+1] CSR, neighbours sorted by index (graph.rs:69-78)
    const uint32_t *adj_idx;
    const uint32_t *site_class;  // [N]
    const uint32_t *class_off;   // [nclasses] offset of the class table inside a replica's table
    const unsigned long long *thr;  // [R][thr_stride] thresholds, index (own << deg) | aligned-mask
    uint32_t thr_stride;
    const uint32_t *colour_sites;  // sites sorted by (colour, index)
    // square layout (L % 64 == 0, uniform J and bias)
    uint32_t L;
    uint32_t *planes;          // [R][2][L][L/64] bit-packed colour planes
    const uint32_t *sq_thr;    // [R][16] thresholds, index own << 3 | #anti-aligned neighbours
    const uint32_t *sq_always; // [R] bit idx set: always flip (threshold 2^32)
    const uint32_t *sq_prob;   // [R] bit idx set: 0 < threshold < 2^32
};
