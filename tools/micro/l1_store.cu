// Does a global store to a sector that sits in L1 leave the sector there (update) or drop it (the next load goes to L2)?
// One thread, dependent loads timed with clock64.  Build: nvcc -arch=sm_100a -o l1_store l1_store.cu
#include <cstdio>
#include <cstdint>
__global__ void k(uint32_t *a, long long *out) {
    uint32_t acc = 0;
    long long t[8];
    // warm: bring the line in
    acc += a[0];
    acc += a[acc & 1];  // dependent
    long long c0 = clock64();
    acc += a[acc & 1];  // L1 hit expected
    long long c1 = clock64();
    t[0] = c1 - c0;
    a[2 + (acc & 1)] = acc;  // store to the SAME 32-byte sector (words 2/3)
    __threadfence_block();
    c0 = clock64();
    acc += a[acc & 1];  // same sector after a store into it
    c1 = clock64();
    t[1] = c1 - c0;
    c0 = clock64();
    acc += a[acc & 1];
    c1 = clock64();
    t[2] = c1 - c0;
    a[9 + (acc & 1)] = acc;  // store to ANOTHER sector of the same 128-byte line (words 8..15)
    __threadfence_block();
    c0 = clock64();
    acc += a[acc & 1];
    c1 = clock64();
    t[3] = c1 - c0;
    // a line never touched: L2 / DRAM reference latencies
    c0 = clock64();
    acc += a[4096 + (acc & 1)];
    c1 = clock64();
    t[4] = c1 - c0;
    // store first to an uncached line, then load it
    a[8192 + (acc & 1)] = acc;
    __threadfence_block();
    c0 = clock64();
    acc += a[8192 + 2 + (acc & 1)];
    c1 = clock64();
    t[5] = c1 - c0;
    c0 = clock64();
    acc += a[8192 + 2 + (acc & 1)];
    c1 = clock64();
    t[6] = c1 - c0;
    for (int i = 0; i < 7; i++) out[i] = t[i];
    out[7] = acc;
}
int main() {
    uint32_t *a;
    long long *o, h[8];
    cudaMalloc(&a, 1 << 20);
    cudaMemset(a, 0, 1 << 20);
    cudaMalloc(&o, 64);
    for (int rep = 0; rep < 2; rep++) {
        k<<<1, 1>>>(a, o);
        cudaMemcpy(h, o, 64, cudaMemcpyDeviceToHost);
        printf("L1 hit %lld | after store to same sector %lld, again %lld | after store to other sector of the line %lld | cold line %lld | load after store to cold line %lld, again %lld\n",
               h[0], h[1], h[2], h[3], h[4], h[5], h[6]);
    }
    return 0;
}
