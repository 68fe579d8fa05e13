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
    // classes are indexed own << 3 | #anti-aligned neighbours; classes with equal thresholds form a group
    const uint32_t *sq_always; // [R] class bits: always flip (threshold 2^32)
    const uint32_t *sq_gT;     // [R][10] threshold of each probabilistic group (0 < T < 2^32)
    const uint32_t *sq_gmem;   // [R][10] class bits of each group (0 = unused)
    uint32_t sq_ngroups;       // max number of groups over the replicas
    int sq_wpr_shift;          // log2(L / 64) if that is a power of two, else -1
};

// reference-schedule moves (classical_ref.cu): GraphState::do_time_step with spin, edge and worm flips
struct ClsRefDev {
    uint32_t N, R, E;
    const uint64_t *key;   // [R]
    uint64_t *cursor;      // [R] position in the replica's sequential stream
    const double *beta;    // [R]
    uint8_t *spins;        // [R][N] bytes
    const uint32_t *adj_start, *adj_idx;  // binding_mat as CSR, neighbours sorted by index (graph.rs:69-78)
    const double *adj_j, *biases;
    const uint32_t *ea, *eb;  // edges in construction order
    const double *cum_w;      // enable_edge_importance_sampling (graph.rs:321-336), nullptr when off
    double total_w;
    uint32_t *path;           // [R][2 (N + 2)] visit_path of the worm move
    int *status;
};
