// Two questions about global stores and the L1 (one thread, dependent loads timed with clock64):
//  (a) does a store into a sector that sits in L1 leave the sector there (update) or drop it (the next load goes to L2)?
//  (b) what does a load cost that follows a store to the same sector / the same line WITHOUT a fence in between (the store
//      still on its way to the L2)?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o l1_store l1_store.cu
#include <cstdio>
#include <cstdint>
__global__ void k(uint32_t *a, long long *out) {
    uint32_t acc = 0;
    long long t[12];
    for (int i = 0; i < 12; i++) t[i] = 0;
    acc += a[0];
    acc += a[acc & 1];
    long long c0 = clock64();
    acc += a[acc & 1];  // L1 hit
    long long c1 = clock64();
    t[0] = c1 - c0;
    a[2 + (acc & 1)] = acc;  // store to the same sector, then a fence
    __threadfence_block();
    c0 = clock64();
    acc += a[acc & 1];
    c1 = clock64();
    t[1] = c1 - c0;
    a[9 + (acc & 1)] = acc;  // other sector of the line, fence
    __threadfence_block();
    c0 = clock64();
    acc += a[acc & 1];
    c1 = clock64();
    t[2] = c1 - c0;
    c0 = clock64();
    acc += a[4096 + (acc & 1)];  // a line only the L2 holds
    c1 = clock64();
    t[3] = c1 - c0;
    // (b) no fence: store, then load the same word / same sector / another sector of the line / another cached line
    uint32_t *b = a + 16384;
    acc += b[0];
    acc += b[64 + (acc & 1)];
    acc += b[acc & 1];
    c0 = clock64();
    b[2 + (acc & 1)] = acc;
    acc += b[2 + (acc & 1)];  // same word
    c1 = clock64();
    t[4] = c1 - c0;
    c0 = clock64();
    b[4 + (acc & 1)] = acc;
    acc += b[acc & 1];  // same sector, other word
    c1 = clock64();
    t[5] = c1 - c0;
    c0 = clock64();
    b[6 + (acc & 1)] = acc;
    acc += b[8 + (acc & 1)];  // next sector of the line (already in L1? no: first touch -> L2)
    c1 = clock64();
    t[6] = c1 - c0;
    c0 = clock64();
    b[6 + (acc & 1)] = acc;
    acc += b[8 + (acc & 1)];  // next sector again (now in L1), store pending to the sector before it
    c1 = clock64();
    t[7] = c1 - c0;
    c0 = clock64();
    b[6 + (acc & 1)] = acc;
    acc += b[64 + (acc & 1)];  // another cached line, store pending elsewhere
    c1 = clock64();
    t[8] = c1 - c0;
    c0 = clock64();
    acc += b[64 + (acc & 1)];  // plain hit for reference (includes clock overhead)
    c1 = clock64();
    t[9] = c1 - c0;
    for (int i = 0; i < 10; i++) out[i] = t[i];
    out[10] = acc;
}
int main() {
    uint32_t *a;
    long long *o, h[12];
    cudaMalloc(&a, 1 << 20);
    cudaMemset(a, 0, 1 << 20);
    cudaMalloc(&o, 128);
    for (int rep = 0; rep < 2; rep++) {
        k<<<1, 1>>>(a, o);
        cudaMemcpy(h, o, 96, cudaMemcpyDeviceToHost);
        printf("(a) L1 hit %lld | after store+fence same sector %lld | other sector of the line %lld | L2 hit %lld\n", h[0], h[1], h[2], h[3]);
        printf("(b) no fence: store+load same word %lld | same sector %lld | next sector (first touch) %lld | next sector (cached) %lld | other cached line %lld | plain hit %lld\n",
               h[4], h[5], h[6], h[7], h[8], h[9]);
    }
    return 0;
}
