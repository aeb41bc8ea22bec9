// Does a shared-memory atomicAdd hand out tickets in ascending lane order among
// the lanes of one warp instruction that hit the same address?  (dev tool)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void k(unsigned long long *viol, unsigned long long *tot, int mode)
{
    __shared__ uint32_t cur[256];
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t acc = (blockIdx.x * 1024 + threadIdx.x) * 2654435761u + 12345;
    unsigned long long v = 0, t = 0;
    for (int it = 0; it < 2000; it++) {
        for (int i = threadIdx.x; i < 256; i += blockDim.x) cur[i] = 0;
        __syncthreads();
        acc = acc * 1664525u + 1013904223u;
        uint32_t nd = mode == 0 ? 4 : mode == 1 ? 32 : 128;
        uint32_t d = (acc >> 13) % nd;
        // every warp uses its own counters (packed u16 pairs when mode==3)
        uint32_t tk;
        if (mode == 3) { d = (acc >> 13) % 64; tk = (atomicAdd(&cur[(d * 4) + (warp >> 3) % 4], 1u << ((warp & 1) * 16)) >> ((warp & 1) * 16)) & 0xFFFF; }
        else tk = atomicAdd(&cur[(d + warp * 7) & 255], 1u);
        // lanes with same address: tickets must ascend with lane (only comparable within one warp if counters private)
        if (mode != 3 ) {
            for (int j = 0; j < 32; j++) {
                uint32_t od = __shfl_sync(0xFFFFFFFFu, d, j), otk = __shfl_sync(0xFFFFFFFFu, tk, j);
                if (j < lane && od == d) { t++; if (otk > tk) v++; }
            }
        }
        __syncthreads();
    }
    atomicAdd(viol, v); atomicAdd(tot, t);
}
int main()
{
    unsigned long long *d, h[2];
    cudaMalloc(&d, 16);
    for (int mode = 0; mode < 3; mode++) {
        for (int nthreads : {32, 1024}) {
            cudaMemset(d, 0, 16);
            k<<<148, nthreads>>>(d, d + 1, mode);
            cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
            printf("mode %d threads %d: same-address lane pairs %llu, order violations %llu\n", mode, nthreads, h[1], h[0]);
        }
    }
    return 0;
}
