/* k_parse.cuh -- the greedy decision over the evaluated heads: k_parse, k_parse_warp (included by dfl_kernels.cu). */
#pragma once
/* ------------------------------------------------------------------------- */
/* k_parse                                                                   */
/* ------------------------------------------------------------------------- */

#define PARSE_THREADS 1024
#define PARSE_WARP_MIN_SEGS 3072              /* from here on a chunk goes to k_parse_warp */
#define PARSE_ROUNDS 4                      /* pointer-doubling rounds */
#define PARSE_STRIDE (1 << PARSE_ROUNDS)    /* the walker visits every 16th token */
#define PARSE_PER_THREAD (DFL_MAX_SEG / PARSE_THREADS)

struct ParseSmem {
    uint16_t np[DFL_MAX_SEG + 2];                        /* position -> position PARSE_STRIDE tokens further on (seglen = end) */
    uint16_t visited[DFL_MAX_SEG / PARSE_STRIDE + 2];    /* positions of tokens 0, 16, 32, ... */
    uint32_t hist[LLS + DS];
    uint32_t nvis, extra;
};

__device__ __forceinline__ int tok_advance(uint32_t v)
{
    const int a = (v >> 16) ? (int)(v & 0xFFFF) : 1;
    return a ? a : 1; /* positions no parser evaluated hold stale words; they are never on the chain, but must not stall it */
}

/*
 * One CTA per segment: the greedy token decision of process_loop
 * (src/deflate_compress.c:265-288) -- from the segment start take the candidate
 * at the current head, advance by its length (or 1 for a literal) -- without
 * walking the ~7000 tokens of a segment one dependent load after the other:
 *   1. np[p] = p + advance(candidate at p) for every position, in shared memory;
 *   2. PARSE_ROUNDS rounds of pointer doubling, np[p] <- np[np[p]] (all reads of a
 *      round before its writes), so np[p] jumps PARSE_STRIDE tokens;
 *   3. one thread follows np from position 0: that is every 16th token of the parse;
 *   4. thread j re-walks the 16 tokens from visited[j] with the candidates themselves,
 *      writes them to tokens[16 j ...] (token order = chain order, no scan needed) and
 *      adds them to the lit/len and distance histograms (what aht_insert feeds,
 *      :270,:277-278, plus the one EOB of :234).
 */
__global__ void __launch_bounds__(PARSE_THREADS, 2)
k_parse(long long n, long long seg0, dfl_params prm, const uint32_t *__restrict__ tokcand, uint32_t *__restrict__ tokens,
        SegArrays A)
{
    extern __shared__ __align__(16) unsigned char parse_raw[];
    ParseSmem &S = *reinterpret_cast<ParseSmem *>(parse_raw);
    const int tid = threadIdx.x;
    const int sl = blockIdx.x;
    const long long seg_lo = (seg0 + sl) * prm.seg_bytes;
    const int seglen = (int)(((seg_lo + prm.seg_bytes < n) ? seg_lo + prm.seg_bytes : n) - seg_lo);
    const uint32_t *tc = tokcand + (size_t)sl * prm.seg_bytes;
    uint32_t *to = tokens + (size_t)sl * prm.seg_bytes;

    for (int i = tid; i < LLS + DS; i += PARSE_THREADS) S.hist[i] = 0;
    if (tid == 0) S.extra = 0;
#pragma unroll
    for (int j = 0; j < PARSE_PER_THREAD; j++) {
        const int pos = j * PARSE_THREADS + tid;
        if (pos < seglen) {
            const int nx = pos + tok_advance(__ldg(tc + pos));
            S.np[pos] = (uint16_t)(nx < seglen ? nx : seglen);
        }
    }
    if (tid == 0) S.np[seglen] = (uint16_t)seglen;
    __syncthreads();
    for (int r = 0; r < PARSE_ROUNDS; r++) {
        uint32_t t[PARSE_PER_THREAD / 2]; /* two 16-bit results per register */
#pragma unroll
        for (int j = 0; j < PARSE_PER_THREAD; j += 2) {
            const int p0 = j * PARSE_THREADS + tid, p1 = p0 + PARSE_THREADS;
            const uint32_t a0 = p0 < seglen ? S.np[S.np[p0]] : 0u;
            const uint32_t a1 = p1 < seglen ? S.np[S.np[p1]] : 0u;
            t[j / 2] = a0 | (a1 << 16);
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < PARSE_PER_THREAD; j += 2) {
            const int p0 = j * PARSE_THREADS + tid, p1 = p0 + PARSE_THREADS;
            if (p0 < seglen) S.np[p0] = (uint16_t)(t[j / 2] & 0xFFFF);
            if (p1 < seglen) S.np[p1] = (uint16_t)(t[j / 2] >> 16);
        }
        __syncthreads();
    }
    if (tid == 0) {
        int pos = 0, nv = 0;
        while (pos < seglen) {
            S.visited[nv++] = (uint16_t)pos;
            pos = S.np[pos];
        }
        S.nvis = (uint32_t)nv;
    }
    __syncthreads();
    const int nvis = (int)S.nvis;
    uint32_t extra = 0;
    for (int j = tid; j < nvis; j += PARSE_THREADS) {
        int h = S.visited[j];
        const int end = j + 1 < nvis ? (int)S.visited[j + 1] : seglen;
        int k = j * PARSE_STRIDE;
        while (h < end) {
            const uint32_t v = __ldg(tc + h);
            to[k++] = v;
            if ((v >> 16) == 0) {
                atomicAdd(&S.hist[v & 0xFF], 1u);
            } else {
                int eb, ev;
                atomicAdd(&S.hist[len_sym((int)(v & 0xFFFF), eb, ev)], 1u);
                extra += eb;
                atomicAdd(&S.hist[LLS + dist_sym((int)(v >> 16), eb, ev)], 1u);
                extra += eb;
            }
            h += tok_advance(v);
        }
        if (j == nvis - 1) A.ntok[sl] = (uint32_t)k;
    }
    if (nvis == 0 && tid == 0) A.ntok[sl] = 0;
    extra = __reduce_add_sync(FULL, extra);
    if ((tid & 31) == 0 && extra) atomicAdd(&S.extra, extra);
    __syncthreads();
    if (tid == 0) {
        S.hist[256] += 1; /* end of block */
        A.extra[sl] = S.extra;
    }
    __syncthreads();
    for (int i = tid; i < LLS; i += PARSE_THREADS) A.ll_freq[(size_t)sl * LLS + i] = S.hist[i];
    if (tid < DS) A.d_freq[(size_t)sl * DS + tid] = S.hist[LLS + tid];
}

#define PARSE_WARPS 8

/*
 * The same decision with one WARP per segment: 32 candidates per step, the heads among
 * them resolved with shuffles.  Far more instructions per token are avoided, but a
 * segment is ~1000 dependent steps: this is the better kernel when there are enough
 * segments to fill every warp slot of the chip (big device-resident chunks), k_parse
 * above when there are not (the 32 MiB pieces of the host pipeline).
 */
__global__ void __launch_bounds__(PARSE_WARPS * 32)
k_parse_warp(long long n, long long seg0, int nseg, dfl_params prm, const uint32_t *__restrict__ tokcand,
             uint32_t *__restrict__ tokens, SegArrays A)
{
    __shared__ uint32_t s_hist[PARSE_WARPS][LLS + DS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sl = blockIdx.x * PARSE_WARPS + warp;
    if (sl >= nseg) return;
    const long long seg_lo = (seg0 + sl) * prm.seg_bytes;
    const int seglen = (int)(((seg_lo + prm.seg_bytes < n) ? seg_lo + prm.seg_bytes : n) - seg_lo);
    const uint32_t *tc = tokcand + (size_t)sl * prm.seg_bytes;
    uint32_t *to = tokens + (size_t)sl * prm.seg_bytes;
    uint32_t *hist = s_hist[warp];
    for (int i = lane; i < LLS + DS; i += 32) hist[i] = 0;
    __syncwarp();

    int cur = 0, ntok = 0;
    uint32_t extra = 0;
    /* two windows of candidates kept in flight */
    uint32_t v1 = lane < seglen ? tc[lane] : 0, v2 = 32 + lane < seglen ? tc[32 + lane] : 0;
    for (int base = 0; base < seglen; base += 32) {
        int i = base + lane;
        const uint32_t v = v1;
        v1 = v2;
        v2 = i + 64 < seglen ? tc[i + 64] : 0;
        int adv = tok_advance(v);
        int lim = seglen - base < 32 ? seglen - base : 32;
        uint32_t heads = 0;
        int c = cur - base;
        if (c < lim && __all_sync(FULL, adv == 1 || lane >= lim)) {
            /* a window of literals (incompressible data): every position from c on is a head */
            heads = (lim == 32 ? 0xFFFFFFFFu : (1u << lim) - 1u) & ~((1u << c) - 1u);
            c = lim;
        } else {
            while (c < lim) {
                heads |= 1u << c;
                c += __shfl_sync(FULL, adv, c);
            }
        }
        cur = base + c;
        __syncwarp();
        if ((heads >> lane) & 1) {
            to[ntok + __popc(heads & lanemask_lt())] = v;
            if ((v >> 16) == 0) {
                atomicAdd(&hist[v & 0xFF], 1u);
            } else {
                int eb, ev;
                atomicAdd(&hist[len_sym((int)(v & 0xFFFF), eb, ev)], 1u);
                extra += eb;
                atomicAdd(&hist[LLS + dist_sym((int)(v >> 16), eb, ev)], 1u);
                extra += eb;
            }
        }
        ntok += __popc(heads);
    }
    for (int o = 16; o; o >>= 1) extra += __shfl_xor_sync(FULL, extra, o);
    __syncwarp();
    if (lane == 0) {
        hist[256] += 1; /* end of block */
        A.ntok[sl] = ntok;
        A.extra[sl] = extra;
    }
    __syncwarp();
    for (int i = lane; i < LLS; i += 32) A.ll_freq[(size_t)sl * LLS + i] = hist[i];
    A.d_freq[(size_t)sl * DS + lane] = hist[LLS + lane];
}
