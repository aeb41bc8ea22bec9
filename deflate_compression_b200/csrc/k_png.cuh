/* k_png.cuh -- PNG scanline filtering: k_png_filter, k_png_filter4 (included by dfl_kernels.cu). */
#pragma once
/* ------------------------------------------------------------------------- */
/* k_png_filter                                                              */
/* ------------------------------------------------------------------------- */

/*
 * PNG scanline filtering ahead of the compressor: the forward direction of
 * the reference's defilter() (src/png/png_decoder.c:819-857; filter ids src/png/include/png_decoder.h:57-61;
 * the Paeth predictor is the PNG specification's -- the reference's, :803-817, never sets pb / pc).
 * One warp per row, two sweeps: the sums of |residual as a signed byte| of the five filter types, then the
 * row filtered with the first type of minimal sum (mode 0) or with type mode - 1.  Rows only read the
 * unfiltered row above, so all rows run in parallel.  HBM: the row and the row above read (the second
 * sweep and the neighbouring warp hit L1/L2), the row written: ~3 B per pixel byte.
 */
#define PNGF_WARPS 8

__device__ __forceinline__ int png_paeth(int a, int b, int c)
{
    const int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

__global__ void __launch_bounds__(PNGF_WARPS * 32)
k_png_filter(const uint8_t *__restrict__ pix, long long row_bytes, long long rows, int bpp, int mode, uint8_t *__restrict__ out)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long y = (long long)blockIdx.x * PNGF_WARPS + warp;
    if (y >= rows) return;
    const uint8_t *cur = pix + y * row_bytes, *prev = cur - row_bytes;
    const bool has_prev = y > 0;
    uint8_t *o = out + y * (row_bytes + 1);
    int type = mode ? mode - 1 : 0;
    if (!mode) {
        uint32_t s0 = 0, s1 = 0, s2 = 0, s3 = 0, s4 = 0;
        for (long long x = lane; x < row_bytes; x += 32) {
            const int v = cur[x], a = x >= bpp ? cur[x - bpp] : 0, b = has_prev ? prev[x] : 0, c = (has_prev && x >= bpp) ? prev[x - bpp] : 0;
            const uint32_t r0 = (uint8_t)v, r1 = (uint8_t)(v - a), r2 = (uint8_t)(v - b), r3 = (uint8_t)(v - ((a + b) >> 1)),
                           r4 = (uint8_t)(v - png_paeth(a, b, c));
            s0 += r0 < 128 ? r0 : 256 - r0;
            s1 += r1 < 128 ? r1 : 256 - r1;
            s2 += r2 < 128 ? r2 : 256 - r2;
            s3 += r3 < 128 ? r3 : 256 - r3;
            s4 += r4 < 128 ? r4 : 256 - r4;
        }
        s0 = __reduce_add_sync(FULL, s0); s1 = __reduce_add_sync(FULL, s1); s2 = __reduce_add_sync(FULL, s2);
        s3 = __reduce_add_sync(FULL, s3); s4 = __reduce_add_sync(FULL, s4);
        uint32_t best = s0;
        if (s1 < best) { best = s1; type = 1; }
        if (s2 < best) { best = s2; type = 2; }
        if (s3 < best) { best = s3; type = 3; }
        if (s4 < best) { best = s4; type = 4; }
    }
    if (lane == 0) o[0] = (uint8_t)type;
    for (long long x = lane; x < row_bytes; x += 32) {
        const int v = cur[x], a = x >= bpp ? cur[x - bpp] : 0, b = has_prev ? prev[x] : 0, c = (has_prev && x >= bpp) ? prev[x - bpp] : 0;
        const int pred = type == 0 ? 0 : type == 1 ? a : type == 2 ? b : type == 3 ? (a + b) >> 1 : png_paeth(a, b, c);
        o[1 + x] = (uint8_t)(v - pred);
    }
}

/* Four pixels bytes at a time (rows that start 4-byte aligned and are a multiple of 4 bytes long): one
 * 32-bit load per row word, byte-wise SIMD arithmetic, the Paeth predictor on 16-bit lanes. */
__device__ __forceinline__ uint32_t png_paeth4(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t r = 0;
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const uint32_t A = (a >> (8 * h)) & 0x00FF00FFu, B = (b >> (8 * h)) & 0x00FF00FFu, C = (c >> (8 * h)) & 0x00FF00FFu;
        const uint32_t pa = __vabsdiffu2(B, C), pb = __vabsdiffu2(A, C), pc = __vabsdiffu2(__vadd2(A, B), __vadd2(C, C));
        const uint32_t m1 = __vcmpleu2(pa, pb) & __vcmpleu2(pa, pc), m2 = __vcmpleu2(pb, pc);
        r |= ((A & m1) | (~m1 & ((B & m2) | (C & ~m2)))) << (8 * h);
    }
    return r;
}

__global__ void __launch_bounds__(PNGF_WARPS * 32)
k_png_filter4(const uint8_t *__restrict__ pix, long long row_bytes, long long rows, int bpp, int mode, uint8_t *__restrict__ out)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long y = (long long)blockIdx.x * PNGF_WARPS + warp;
    if (y >= rows) return;
    const uint32_t *cur = reinterpret_cast<const uint32_t *>(pix + y * row_bytes);
    const uint32_t *prev = reinterpret_cast<const uint32_t *>(pix + (y - 1) * row_bytes);
    const bool has_prev = y > 0;
    const int nw = (int)(row_bytes >> 2), sh = 8 * (4 - bpp); /* bpp 1..4 */
    uint8_t *o = out + y * (row_bytes + 1);
    int type = mode ? mode - 1 : 0;
    if (!mode) {
        uint32_t s0 = 0, s1 = 0, s2 = 0, s3 = 0, s4 = 0;
        for (int k0 = 0; k0 < nw; k0 += 32) {
            const int k = k0 + lane;
            const bool in = k < nw;
            const uint32_t wc = in ? __ldg(cur + k) : 0u, wp = (in && has_prev) ? __ldg(prev + k) : 0u;
            uint32_t wcl = __shfl_up_sync(FULL, wc, 1), wpl = __shfl_up_sync(FULL, wp, 1);
            if (lane == 0) {
                wcl = k > 0 ? __ldg(cur + k - 1) : 0u;
                wpl = (k > 0 && has_prev) ? __ldg(prev + k - 1) : 0u;
            }
            if (in) {
                const uint32_t a = __funnelshift_r(wcl, wc, sh), b = wp, c = __funnelshift_r(wpl, wp, sh);
                s0 += __vsadu4(__vabs4(wc), 0u);
                s1 += __vsadu4(__vabs4(__vsub4(wc, a)), 0u);
                s2 += __vsadu4(__vabs4(__vsub4(wc, b)), 0u);
                s3 += __vsadu4(__vabs4(__vsub4(wc, __vhaddu4(a, b))), 0u);
                s4 += __vsadu4(__vabs4(__vsub4(wc, png_paeth4(a, b, c))), 0u);
            }
        }
        s0 = __reduce_add_sync(FULL, s0); s1 = __reduce_add_sync(FULL, s1); s2 = __reduce_add_sync(FULL, s2);
        s3 = __reduce_add_sync(FULL, s3); s4 = __reduce_add_sync(FULL, s4);
        uint32_t best = s0;
        if (s1 < best) { best = s1; type = 1; }
        if (s2 < best) { best = s2; type = 2; }
        if (s3 < best) { best = s3; type = 3; }
        if (s4 < best) { best = s4; type = 4; }
    }
    if (lane == 0) o[0] = (uint8_t)type;
    for (int k0 = 0; k0 < nw; k0 += 32) {
        const int k = k0 + lane;
        const bool in = k < nw;
        const uint32_t wc = in ? __ldg(cur + k) : 0u, wp = (in && has_prev) ? __ldg(prev + k) : 0u;
        uint32_t wcl = __shfl_up_sync(FULL, wc, 1), wpl = __shfl_up_sync(FULL, wp, 1);
        if (lane == 0) {
            wcl = k > 0 ? __ldg(cur + k - 1) : 0u;
            wpl = (k > 0 && has_prev) ? __ldg(prev + k - 1) : 0u;
        }
        if (in) {
            const uint32_t a = __funnelshift_r(wcl, wc, sh), b = wp, c = __funnelshift_r(wpl, wp, sh);
            const uint32_t pred = type == 0 ? 0u : type == 1 ? a : type == 2 ? b : type == 3 ? __vhaddu4(a, b) : png_paeth4(a, b, c);
            const uint32_t r = __vsub4(wc, pred);
            uint8_t *d = o + 1 + 4 * k;
            d[0] = (uint8_t)r; d[1] = (uint8_t)(r >> 8); d[2] = (uint8_t)(r >> 16); d[3] = (uint8_t)(r >> 24);
        }
    }
}
