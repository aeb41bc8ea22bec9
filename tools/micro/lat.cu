// micro-benchmarks for warp primitives on sm_100a (dev tool)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define FULL 0xFFFFFFFFu
__global__ void k(unsigned long long *out, int nwarps_active)
{
    __shared__ uint32_t sm[4096];
    __shared__ uint16_t s16[8192];
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = 0;
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) s16[i] = i;
    __syncthreads();
    if (warp >= nwarps_active) return;
    uint32_t acc = lane * 2654435761u + warp;
    long long t[12];
    const int N = 256;
    // (0) ballot chain: 8 ballots per iter
    t[0] = clock64();
    for (int i = 0; i < N; i++) {
        uint32_t d = acc >> 20, peers = FULL;
#pragma unroll
        for (int b = 0; b < 7; b++) { uint32_t bal = __ballot_sync(FULL, (d >> b) & 1); peers &= ((d >> b) & 1) ? bal : ~bal; }
        acc = acc * 1664525u + __popc(peers);
    }
    t[1] = clock64();
    // (1) match_any distinct values
    for (int i = 0; i < N; i++) { uint32_t m = __match_any_sync(FULL, acc >> 8); acc = acc * 1664525u + m + lane; }
    t[2] = clock64();
    // (2) match_any on 4 distinct values
    for (int i = 0; i < N; i++) { uint32_t m = __match_any_sync(FULL, (acc >> 8) & 3); acc = acc * 1664525u + m + lane; }
    t[3] = clock64();
    // (3) smem atomicAdd spread addresses
    for (int i = 0; i < N; i++) { atomicAdd(&sm[(acc >> 10) & 4095], 1u); acc = acc * 1664525u + 1013904223u; }
    t[4] = clock64();
    // (4) smem atomicAdd with return, dependent
    for (int i = 0; i < N; i++) { uint32_t o = atomicAdd(&sm[(acc >> 10) & 4095], 1u); acc = acc * 1664525u + o; }
    t[5] = clock64();
    // (5) volatile LDS.U16 -> STS.U16 dependent chain
    volatile uint16_t *v16 = s16;
    for (int i = 0; i < N; i++) { uint32_t a = (acc >> 10) & 8191; uint32_t x = v16[a]; v16[a] = (uint16_t)(x + 1); acc = acc * 1664525u + x; }
    t[6] = clock64();
    // (6) plain LDS.U16 dependent pointer chase
    { uint32_t a = lane; for (int i = 0; i < N; i++) { a = s16[(a * 37 + i) & 8191]; } acc += a; }
    t[7] = clock64();
    // (7) shfl chain
    for (int i = 0; i < N; i++) { acc = __shfl_sync(FULL, acc, (acc >> 5) & 31) + 1; }
    t[8] = clock64();
    // (8) redux
    for (int i = 0; i < N; i++) { acc = __reduce_or_sync(FULL, acc) + lane; }
    t[9] = clock64();
    if (lane == 0 && warp == 0) { for (int j = 0; j < 9; j++) out[j] = (unsigned long long)(t[j + 1] - t[j]); }
    if (acc == 0x12345678) out[20] = acc;
}
int main()
{
    unsigned long long *d, h[32];
    cudaMalloc(&d, 256);
    const char *names[9] = {"8 ballots+popc", "match_any 32 distinct", "match_any 4 distinct", "ATOMS spread (no ret)", "ATOMS spread (ret, dep)",
                            "volatile LDS16->STS16 dep", "LDS16 pointer chase", "shfl chain", "redux.or chain"};
    for (int nw : {1, 4, 16, 32}) {
        k<<<1, 1024>>>(d, nw);
        cudaMemcpy(h, d, 256, cudaMemcpyDeviceToHost);
        printf("active warps %d (cycles per iteration, warp 0):\n", nw);
        for (int j = 0; j < 9; j++) printf("  %-28s %8.1f\n", names[j], h[j] / 256.0);
    }
    return 0;
}
