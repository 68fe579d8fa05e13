// common.cuh -- shared device helpers: Philox4x32-10, the injected word stream and the rand-0.8
// draw->value mappings, operator-word helpers, error plumbing.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define QMCB_TAG_CLUS 0x434C5553u  // counter word 3 of the FAST-mode cluster bits
#define QMCB_TAG_DIAG 0x44494147u  // counter word 3 of the COUNTER-mode diagonal update: one block per slot
#define QMCB_TAG_CB 0x43420000u    // counter word 3 (| colour) of the checkerboard draws, high 16 bits
#define QMCB_TAG_CB2 0x43430000u   // ... low 16 bits
#define OP_EMPTY 0xFFFFFFFFu
#define NONE32 0xFFFFFFFFu

// device status bits (SseDev::status / ClsDev::status)
#define DEV_ERR_CAPACITY 1
#define DEV_ERR_INVARIANT 2
#define DEV_ERR_STACK 4
#define DEV_ERR_PROB 8

struct Philox4 {
    uint32_t x, y, z, w;
};

__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2,
                                                           uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
#ifdef __CUDA_ARCH__
        uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
#else
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t h0 = (uint32_t)(p0 >> 32), l0 = (uint32_t)p0, h1 = (uint32_t)(p1 >> 32), l1 = (uint32_t)p1;
#endif
        uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
        c0 = n0, c1 = l1, c2 = n2, c3 = l0;
        k0 += 0x9E3779B9u, k1 += 0xBB67AE85u;
    }
    Philox4 o;
    o.x = c0, o.y = c1, o.z = c2, o.w = c3;
    return o;
}

// W[c] of the sequential stream: counter = c >> 1, word = (x[2(c&1)+1] << 32) | x[2(c&1)]
__host__ __device__ __forceinline__ uint64_t stream_word(uint64_t key, uint64_t c) {
    uint64_t blk = c >> 1;
    Philox4 o = philox4x32_10((uint32_t)blk, (uint32_t)(blk >> 32), 0u, 0u, (uint32_t)key, (uint32_t)(key >> 32));
    return (c & 1) ? (((uint64_t)o.w << 32) | o.z) : (((uint64_t)o.y << 32) | o.x);
}

// Bernoulli p_int of rand 0.8: (p * 2^64) as u64 for p in [0,1)
__device__ __forceinline__ uint64_t bool_threshold(double p) {
    return __double2ull_rz(p * 18446744073709551616.0);
}

// ---- operator word ---------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t op_bond(uint32_t w) { return w & 0xFFFFFFu; }
__host__ __device__ __forceinline__ uint32_t op_in(uint32_t w) { return (w >> 24) & 3u; }
__host__ __device__ __forceinline__ uint32_t op_out(uint32_t w) { return (w >> 26) & 3u; }
__host__ __device__ __forceinline__ bool op_is_diag(uint32_t w) { return op_in(w) == op_out(w); }
__host__ __device__ __forceinline__ uint32_t make_op(uint32_t bond, uint32_t in, uint32_t out) {
    return bond | (in << 24) | (out << 26);
}

#define CUDA_TRY(expr)                                                         \
    do {                                                                       \
        cudaError_t e_ = (expr);                                               \
        if (e_ != cudaSuccess) return fail_cuda(e_, #expr, __FILE__, __LINE__); \
    } while (0)
