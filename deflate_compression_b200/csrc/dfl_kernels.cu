/*
 * dfl_kernels.cu -- sm_100a DEFLATE compress kernels + the thin C-ABI
 * (include/deflate_b200.h).  B200-native replacement for the hot path of
 * mparker97/deflate_compression: process_loop() in src/deflate_compress.c:219-347
 * (hash-chain match search, greedy parse) and the Huffman accounting of
 * src/h_tree.c:75-302, completed with the encode side the reference leaves as
 * TODOs (src/deflate_compress.c:269,:279).
 *
 * Pipeline, per chunk of <= DFL_CHUNK_BYTES input (all on one stream, no host sync):
 *   k_match : one CTA per segment.  [window | segment] staged into shared memory
 *             with TMA bulk copies; one warp links every position into its hash
 *             bucket chain (16-bit backward deltas); all warps then find, for
 *             every position of the segment, the longest / nearest match by
 *             walking the chain newest -> oldest (persistent lanes, refilled from
 *             a block-wide counter).  Also the segment's Adler-32 partial sums.
 *   k_parse : one warp per segment; greedy resolution of the per-position
 *             candidates into the token stream (in place), lit/len + dist
 *             histograms with shared-memory atomics.
 *   k_huff  : one warp per segment; length-limited code lengths (15/15/7),
 *             canonical codes, RFC 1951 3.2.7 header, exact stored/fixed/dynamic
 *             cost and choice.
 *   k_scan  : exclusive scan of segment sizes -> byte offsets in the stream.
 *   k_emit  : one CTA per segment; prefix-sum bit packer, 128-bit stores.
 * No tensor cores: nothing here is a dense contraction.
 */
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "deflate_b200.h"

#define FULL 0xFFFFFFFFu
#define NLL 286
#define ND 30
#define NCL 19
#define LLS 288          /* stride of per-segment lit/len tables */
#define DS 32            /* stride of per-segment distance tables */
#define HDR_WORDS 160    /* 640 B >= 3+14+57+316*14 bits */
#define NPOS_MAX 65536
#define HASH_SIZE DFL_HASH_BUCKETS
#define MATCH_THREADS 1024
#ifndef REFILL_MIN_IDLE
#define REFILL_MIN_IDLE 8 
#endif
#define DFL_CHUNK_BYTES (256u << 20)
#define ADLER_MOD 65521u

enum { K_MATCH = 0, K_PARSE, K_HUFF, K_SCAN, K_EMIT, K_FINAL, K_INFLATE, K_INF_RESOLVE, K_INF_FINISH, K_PNG_FILTER, K_NUM };
static const char *k_names[K_NUM] = {"k_match", "k_parse", "k_huff", "k_scan", "k_emit", "k_final", "k_inflate", "k_inf_resolve", "k_inf_finish", "k_png_filter"};

struct DevState {
    unsigned long long total_out;  /* bytes of deflate body written so far */
    unsigned long long adl_a;      /* sum of per-segment byte sums (mod 65521 each) */
    unsigned long long adl_b;      /* sum of per-segment weighted sums */
    unsigned int error;            /* 1 = output overflow */
    unsigned int adler;            /* final Adler-32 */
    unsigned long long t_phase[24]; /* k_match SM cycles summed over CTAs: stage-in, build, search, total */
    unsigned int sort_fallbacks;   /* segments whose index had to be rebuilt with the ballot-ranked scatter */
    unsigned int inf_pending;      /* k_inf_resolve: some word is still a pointer */
};

struct SegArrays {
    uint32_t *ll_freq;   /* [nseg][LLS] */
    uint32_t *d_freq;    /* [nseg][DS] */
    uint32_t *ntok;      /* [nseg] */
    uint32_t *extra;     /* [nseg] sum of length/distance extra bits */
    uint32_t *lltab;     /* [nseg][LLS]  code_rev | len << 16 of the chosen block type */
    uint32_t *dtab;      /* [nseg][DS] */
    uint32_t *hdr;       /* [nseg][HDR_WORDS] packed block header bits */
    uint32_t *hdr_bits;  /* [nseg] */
    uint32_t *btype;     /* [nseg] */
    uint32_t *out_bytes; /* [nseg] */
    unsigned long long *block_bits; /* [nseg] */
    unsigned long long *out_off;    /* [nseg] */
};

/* ------------------------------------------------------------------------- */
/* small device helpers                                                      */
/* ------------------------------------------------------------------------- */

__device__ __forceinline__ uint32_t lanemask_lt()
{
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

/* unaligned little-endian 4-byte load from a word-aligned shared array */
__device__ __forceinline__ uint32_t ld4(const uint32_t *w, uint32_t byte_idx)
{
    uint32_t i = byte_idx >> 2;
    uint32_t lo = w[i], hi = w[i + 1];
    return __funnelshift_r(lo, hi, (byte_idx & 3) * 8);
}

/* Explicit shared-window addressing for the hot loops: a 32-bit shared address computed once
 * (ptxas otherwise re-derives the window base with S2R SR_CgaCtaId inside the loops). */
__device__ __forceinline__ uint32_t lds_u32(uint32_t a)
{
    uint32_t v;
    asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t a)
{
    uint32_t v;
    asm("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t lds_u8(uint32_t a)
{
    uint32_t v;
    asm("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts_u16(uint32_t a, uint32_t v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t atoms_add(uint32_t a, uint32_t v)
{
    uint32_t o;
    asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(o) : "r"(a), "r"(v) : "memory");
    return o;
}
__device__ __forceinline__ void reds_add(uint32_t a, uint32_t v) { asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
/* unaligned little-endian 4-byte load at shared address a */
__device__ __forceinline__ uint32_t ld4s(uint32_t a)
{
    const uint32_t w = a & ~3u;
    return __funnelshift_r(lds_u32(w), lds_u32(w + 4), a << 3);
}

/* DFL_HASH4 with the group width chosen at run time: sb = 0 gives plain trigram buckets (DFL_HASH3),
 * sb = DFL_SUB_BITS the 4-byte buckets.  ngrp = DFL_HASH_BUCKETS >> sb. */
__device__ __forceinline__ uint32_t hash4(uint32_t v, uint32_t ngrp, int sb)
{
    const uint32_t grp = __umulhi((v & 0xFFFFFFu) * 0x9E3779B1u, ngrp);
    return (grp << sb) | __funnelshift_l((v >> 24) * 0x9E3779B1u, 0u, sb);
}

/* RFC 1951 3.2.5: length 3..258 -> symbol 257..285 (same numbers as
 * get_len_code, src/deflate_compress.c:182-200; extra bits per the RFC). */
__device__ __forceinline__ int len_sym(int len, int &eb, int &ev)
{
    int t = len - 3;
    if (len == 258) { eb = 0; ev = 0; return 285; }
    if (t < 8) { eb = 0; ev = 0; return 257 + t; }
    int k = 31 - __clz(t);
    eb = k - 2;
    ev = t & ((1 << eb) - 1);
    return 261 + 4 * eb + ((t >> eb) & 3);
}

/* distance 1..32768 -> symbol 0..29 (get_dist_code, src/deflate_compress.c:202-217) */
__device__ __forceinline__ int dist_sym(int dist, int &eb, int &ev)
{
    int t = dist - 1;
    if (t < 4) { eb = 0; ev = 0; return t; }
    int k = 31 - __clz(t);
    eb = k - 1;
    ev = t & ((1 << eb) - 1);
    return 2 * k + ((t >> eb) & 1);
}

__device__ __forceinline__ int fixed_ll_len(int s) { return s < 144 ? 8 : s < 256 ? 9 : s < 280 ? 7 : 8; }

/* mbarrier + TMA bulk copy (global -> shared), sm_90+ PTX */
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    uint32_t a = (uint32_t)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    uint32_t a = (uint32_t)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
    uint32_t d = (uint32_t)__cvta_generic_to_shared(dst_smem);
    uint32_t b = (uint32_t)__cvta_generic_to_shared(bar);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(d), "l"(src_gmem), "r"(bytes), "r"(b)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t a = (uint32_t)__cvta_generic_to_shared(bar);
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(a), "r"(parity)
        : "memory");
}

/* ------------------------------------------------------------------------- */
/* k_match                                                                   */
/* ------------------------------------------------------------------------- */

#ifndef MATCH_CHUNK
#define MATCH_CHUNK 32 /* positions per speculative parse chunk */
#endif
#ifndef HOP_BATCH
#define HOP_BATCH 8 /* bucket entries a lane visits before the warp re-converges */
#endif
#ifndef COOP_MIN
#define COOP_MIN 64 /* buckets longer than this are walked by the whole warp */
#endif
#ifndef SPLIT_MIN_SQ
#define SPLIT_MIN_SQ 3u /* split the trigram buckets when the sampled sum of squares exceeds this many times the sample count */
#endif
#ifndef SEARCH_WARPS
#define SEARCH_WARPS 32  /* warps that run the chunk parsers (the index build uses all 32) */
#endif
#ifndef SKIP_OVERRUN
#define SKIP_OVERRUN 1   /* a chunk some chain has already entered is not started on its own */
#endif
#define SORT_WARPS 32    /* every warp sorts one stripe */
#define SORT_BINS 128    /* 7-bit digits: bucket = digit2 * 128 + digit1 */
#define SCRATCH_SLOTS 256

struct MatchSmem {
    uint32_t data[(NPOS_MAX + 64) / 4]; /* [window | segment] bytes, 16 B of slack either side */
    uint16_t list[NPOS_MAX];            /* positions sorted by (bucket, position) */
    uint16_t off[HASH_SIZE + 2];        /* bucket b owns list[off[b], off[b+1]); radix-sort counters during the build */
    uint32_t claimed[DFL_MAX_SEG / 32]; /* positions some lane has taken as a parse head */
    uint64_t bar;
    int next_chunk;
    uint32_t red_a[32];
    unsigned long long red_b[32];
    uint32_t scan_w[32];
    uint32_t scan_total;
};

/* exclusive block scan of one value per thread (all MATCH_THREADS threads call) */
__device__ __forceinline__ uint32_t block_excl_scan(MatchSmem &S, uint32_t v, int lane, int warp)
{
    uint32_t x = v;
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(FULL, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) S.scan_w[warp] = x;
    __syncthreads();
    if (warp == 0) {
        uint32_t t = S.scan_w[lane], u = t;
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t y = __shfl_up_sync(FULL, u, o);
            if (lane >= o) u += y;
        }
        S.scan_w[lane] = u - t;
    }
    __syncthreads();
    uint32_t r = S.scan_w[warp] + x - v;
    __syncthreads();
    return r;
}

/* (digit, stripe) counter of the radix passes: u16, two per 32-bit word, parked in S.off
 * (free until the offsets are written).  Stripe-major layout so that the lanes of a warp
 * (one stripe, many digits) spread over the banks. */
__device__ __forceinline__ uint32_t cnt_word(uint32_t stripe, uint32_t d) { return (stripe * SORT_BINS + d) >> 1; }
__device__ __forceinline__ uint32_t cnt_shift(uint32_t d) { return (d & 1) * 16; }
#define CNT_WORDS (SORT_BINS * SORT_WARPS / 2) /* 32-bit words of one counter array (8 KB) */

/*
 * Scatter sweep of a sort warp over its stripe, 32 elements per step, 4 steps per
 * round, in order.  PASS 1 reads positions in natural order, PASS 2 the pass-1
 * order parked in gbuf (one round of loads kept in flight).  The (digit, stripe)
 * counters are running cursors.
 *   FAST:  slot = atomicAdd(cursor, 1).  Lanes of one warp instruction that hit
 *          the same shared-memory word are served in ascending lane order on this
 *          chip (measured: 0 violations in 1.3e9 pairs, tools/micro/atom_order.cu),
 *          which makes the scatter stable.  It is not an architectural guarantee,
 *          so the result is verified afterwards and, if ever wrong, redone with
 *   !FAST: lanes with the same digit found with 7 ballots; cursor + rank in the
 *          group.  Stable by construction; several times slower (a vote costs ~3
 *          cycles of a unit shared by the whole SM).
 */
template <int PASS, bool FAST>
__device__ __forceinline__ void scatter_sweep(MatchSmem &S, uint32_t soff, int k_lo, int k_hi, int SS, const uint16_t *gbuf,
                                              int lane, int warp, int sb)
{
    const uint32_t ss_inv = SS ? 0xFFFFFFFFu / (uint32_t)SS + 1u : 0u; /* x / SS == umulhi(x, ss_inv) for x * SS < 2^32 */
    const uint32_t ngrp = HASH_SIZE >> sb;
    /* pass-1 cursors at S.off[0..), pass-2 counters / cursors behind them */
    uint32_t *cnt32 = reinterpret_cast<uint32_t *>(S.off) + (PASS == 1 ? 0 : CNT_WORDS);
    volatile uint16_t *cnt16 = reinterpret_cast<volatile uint16_t *>(cnt32);
    uint32_t dbase = (uint32_t)__cvta_generic_to_shared(S.data) + soff;
    uint32_t lbase = (uint32_t)__cvta_generic_to_shared(S.list);
    uint32_t cbase = (uint32_t)__cvta_generic_to_shared(S.off) + (PASS == 1 ? 0 : CNT_WORDS * 4);
    asm volatile("" : "+r"(dbase), "+r"(lbase), "+r"(cbase)); /* keep the window addresses in registers */
    const uint32_t c2base = cbase + CNT_WORDS * 4; /* only used by pass 1 */
    const uint32_t lt = lanemask_lt();
    uint32_t eq[4] = {0, 0, 0, 0};
    if (PASS == 2) {
#pragma unroll
        for (int j = 0; j < 4; j++)
            if (k_lo + j * 32 + lane < k_hi) eq[j] = __ldcg(gbuf + k_lo + j * 32 + lane);
    }
    for (int k0 = k_lo; k0 < k_hi; k0 += 128) {
        uint32_t e[4], d[4], d2[4];
#pragma unroll
        for (int j = 0; j < 4; j++) { /* bytes + bucket digit of the round's 4 steps: independent work first */
            const int k = k0 + j * 32 + lane;
            e[j] = PASS == 1 ? (uint32_t)k : eq[j];
            if (PASS == 2 && k + 128 < k_hi) eq[j] = __ldcg(gbuf + k + 128);
            const uint32_t b = hash4(ld4s(dbase + (k < k_hi ? e[j] : 0)), ngrp, sb);
            d[j] = PASS == 1 ? (b & (SORT_BINS - 1)) : (b / SORT_BINS);
            d2[j] = b / SORT_BINS;
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int k = k0 + j * 32 + lane;
            if (k0 + j * 32 >= k_hi) break;
            const bool valid = k < k_hi;
            if (FAST) {
                if (valid) {
                    const uint32_t sh = cnt_shift(d[j]);
                    const uint32_t slot = (atoms_add(cbase + 4 * cnt_word(warp, d[j]), 1u << sh) >> sh) & 0xFFFF;
                    sts_u16(lbase + 2 * slot, e[j]);
                    /* pass 1 knows where the element lands: count pass 2's digit for that stripe now */
                    if (PASS == 1) reds_add(c2base + 4 * cnt_word(__umulhi(slot, ss_inv), d2[j]), 1u << cnt_shift(d2[j]));
                }
            } else {
                uint32_t peers = __ballot_sync(FULL, valid);
#pragma unroll
                for (int bit = 0; bit < 7; bit++) {
                    const uint32_t bal = __ballot_sync(FULL, (d[j] >> bit) & 1);
                    peers &= ((d[j] >> bit) & 1) ? bal : ~bal;
                }
                if (valid) {
                    const uint32_t ci = warp * SORT_BINS + d[j];
                    const uint32_t base = cnt16[ci];
                    if ((peers >> lane) == 1u) cnt16[ci] = (uint16_t)(base + __popc(peers)); /* top lane of the group */
                    const uint32_t slot = base + __popc(peers & lt);
                    sts_u16(lbase + 2 * slot, e[j]);
                    if (PASS == 1) reds_add(c2base + 4 * cnt_word(__umulhi(slot, ss_inv), d2[j]), 1u << cnt_shift(d2[j]));
                }
            }
        }
    }
}

/*
 * One stable counting-sort pass over the NI indexed positions, by a 7-bit digit
 * of their bucket.  PASS 1: natural order -> S.list by the low digit.  PASS 2:
 * the pass-1 order (parked linearly in gbuf, global scratch, L2 resident) ->
 * S.list by the high digit.  After both passes S.list is ordered by
 * (bucket, position).  Both scatters stay in shared memory.  The input is cut
 * into one contiguous stripe per warp; counting is order free (shared memory
 * atomics), the scatter of a stripe is done in order by its warp.
 */
template <int PASS, bool FAST>
__device__ __forceinline__ void sort_pass(MatchSmem &S, uint32_t soff, int NI, const uint16_t *gbuf, int tid, int lane, int warp,
                                          DevState *state, int sb)
{
    const uint32_t ngrp = HASH_SIZE >> sb;
    long long c0 = clock64(), c1;
#define STAMP2(slot) do { c1 = clock64(); if (tid == 0 && PASS == 1) atomicAdd(&state->t_phase[slot], (unsigned long long)(c1 - c0)); c0 = c1; } while (0)
    uint32_t *cnt32 = reinterpret_cast<uint32_t *>(S.off) + (PASS == 1 ? 0 : CNT_WORDS);
    uint16_t *cnt16 = reinterpret_cast<uint16_t *>(cnt32);
    uint32_t dbase = (uint32_t)__cvta_generic_to_shared(S.data) + soff;
    uint32_t cbase = (uint32_t)__cvta_generic_to_shared(S.off) + (PASS == 1 ? 0 : CNT_WORDS * 4);
    asm volatile("" : "+r"(dbase), "+r"(cbase));
    const int SS = ((NI + SORT_WARPS * 32 - 1) / (SORT_WARPS * 32)) * 32; /* stripe length, multiple of 32 */
    const uint32_t ss_inv = SS ? 0xFFFFFFFFu / (uint32_t)SS + 1u : 0u;     /* x / SS == umulhi(x, ss_inv) for x * SS < 2^32 */

    if (PASS == 1) {
        for (int i = tid; i < 2 * CNT_WORDS; i += MATCH_THREADS) reinterpret_cast<uint32_t *>(S.off)[i] = 0;
        __syncthreads();
        /* count the low digits: every thread takes 8 consecutive positions; their 8 trigrams come
         * out of three unaligned word reads */
        for (int k8 = tid * 8; k8 < NI; k8 += MATCH_THREADS * 8) {
            const uint32_t w0 = ld4s(dbase + k8), w1 = ld4s(dbase + k8 + 4), w2 = ld4s(dbase + k8 + 8);
            const uint32_t sw = __umulhi((uint32_t)k8, ss_inv); /* 8 | SS: the 8 positions share a stripe */
#pragma unroll
            for (int j = 0; j < 8; j++) {
                if (k8 + j < NI) {
                    const uint32_t t = j < 4 ? __funnelshift_r(w0, w1, (j & 3) * 8) : __funnelshift_r(w1, w2, (j & 3) * 8);
                    const uint32_t d = hash4(t, ngrp, sb) & (SORT_BINS - 1);
                    reds_add(cbase + 4 * cnt_word(sw, d), 1u << cnt_shift(d));
                }
            }
        }
        __syncthreads();
    }
    STAMP2(8);
    /* (pass 2's digits were counted by pass 1's scatter) */
    /* exclusive scan in (digit, stripe) order: SORT_BINS * SORT_WARPS = 4096 u16 entries, 4 per thread */
    {
        const uint32_t d = tid / (SORT_WARPS / 4), s0 = (tid % (SORT_WARPS / 4)) * 4;
        uint32_t c[4], sum = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) { c[j] = cnt16[(s0 + j) * SORT_BINS + d]; sum += c[j]; }
        uint32_t base = block_excl_scan(S, sum, lane, warp);
#pragma unroll
        for (int j = 0; j < 4; j++) { cnt16[(s0 + j) * SORT_BINS + d] = (uint16_t)base; base += c[j]; }
    }
    __syncthreads();
    STAMP2(9);
    {
        const int k_lo = warp * SS, k_hi = (k_lo + SS < NI) ? k_lo + SS : NI;
        scatter_sweep<PASS, FAST>(S, soff, k_lo, k_hi, SS, gbuf, lane, warp, sb);
    }
    __syncthreads();
    STAMP2(10);
#undef STAMP2
}

/* Sort the NI indexed positions by (bucket, position) into S.list and fill the
 * bucket offsets.  Returns (to every thread) whether the order verified. */
template <bool FAST>
__device__ __forceinline__ bool build_index(MatchSmem &S, uint32_t soff, int NI, uint16_t *gbuf, int tid, int lane, int warp,
                                            DevState *state, int sb)
{
    const uint32_t ngrp = HASH_SIZE >> sb;
    uint32_t dbase = (uint32_t)__cvta_generic_to_shared(S.data) + soff;
    uint32_t lbase = (uint32_t)__cvta_generic_to_shared(S.list);
    uint32_t obase = (uint32_t)__cvta_generic_to_shared(S.off);
    asm volatile("" : "+r"(dbase), "+r"(lbase), "+r"(obase));
    long long c0 = clock64(), c1;
#define STAMP(slot) do { c1 = clock64(); if (tid == 0) atomicAdd(&state->t_phase[slot], (unsigned long long)(c1 - c0)); c0 = c1; } while (0)
    sort_pass<1, FAST>(S, soff, NI, gbuf, tid, lane, warp, state, sb);
    STAMP(4);
    /* park the pass-1 order in global scratch: linear, 16 B per thread */
    for (int k8 = tid * 8; k8 < NI; k8 += MATCH_THREADS * 8)
        __stcg(reinterpret_cast<uint4 *>(gbuf + k8), *reinterpret_cast<const uint4 *>(&S.list[k8]));
    STAMP(5);
    sort_pass<2, FAST>(S, soff, NI, gbuf, tid, lane, warp, state, sb);
    STAMP(6);
    /* bucket offsets from the sorted order (off[b] = first slot whose bucket is >= b), and
     * the check that every bucket is in ascending position order.  Slot kk is compared with
     * slot kk-1; consecutive threads take consecutive slots. */
    bool bad = false;
    if (NI == 0) {
        for (int i = tid; i < HASH_SIZE + 2; i += MATCH_THREADS) S.off[i] = 0;
        return true;
    }
    /* every offset starts as "unset"; the first slot of each non-empty bucket writes its own; the
     * empty buckets then take the offset of the next non-empty one (suffix minimum) */
    for (int i = tid; i < (HASH_SIZE + 2) / 2; i += MATCH_THREADS) reinterpret_cast<uint32_t *>(S.off)[i] = 0xFFFFFFFFu;
    __syncthreads();
    for (int k0 = 0; k0 < NI + 1; k0 += 4 * MATCH_THREADS) {
        /* 4 consecutive slots per thread (one 8-byte read), the slot before them from the lane below */
        const int kk0 = k0 + 4 * tid;
        int q[4], b[4];
        {
            uint32_t lo = 0, hi = 0;
            if (kk0 < NI) asm("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(lo), "=r"(hi) : "r"(lbase + 2 * kk0));
            q[0] = lo & 0xFFFF; q[1] = lo >> 16; q[2] = hi & 0xFFFF; q[3] = hi >> 16;
        }
#pragma unroll
        for (int j = 0; j < 4; j++) b[j] = kk0 + j < NI ? (int)hash4(ld4s(dbase + q[j]), ngrp, sb) : HASH_SIZE;
        int pq = __shfl_up_sync(FULL, q[3], 1), pb_ = __shfl_up_sync(FULL, b[3], 1);
        if (lane == 0) {
            pq = kk0 > 0 && kk0 <= NI ? (int)lds_u16(lbase + 2 * (kk0 - 1)) : -1;
            pb_ = kk0 > 0 && kk0 <= NI ? (int)hash4(ld4s(dbase + pq), ngrp, sb) : -1;
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int kk = kk0 + j;
            if (kk <= NI) {
                bad |= kk < NI && ((b[j] < pb_) || (b[j] == pb_ && q[j] <= pq));
                if (b[j] != pb_) sts_u16(obase + 2 * b[j], kk); /* slot NI carries bucket HASH_SIZE: the end marker */
            }
            pq = q[j];
            pb_ = b[j];
        }
    }
    __syncthreads();
    {
        /* suffix minimum over the HASH_SIZE + 1 offsets: 16 consecutive entries per thread */
        const int e0 = tid * 16;
        uint32_t v[16], m = 0xFFFFu;
#pragma unroll
        for (int j = 15; j >= 0; j--) {
            v[j] = e0 + j <= HASH_SIZE ? lds_u16(obase + 2 * (e0 + j)) : 0xFFFFu;
            m = v[j] < m ? v[j] : m;
        }
        /* minimum over all threads to the right: inclusive suffix scan in the warp, then across warps */
        uint32_t x = m;
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_down_sync(FULL, x, o);
            if (lane + o < 32 && y < x) x = y;
        }
        uint32_t carry = __shfl_down_sync(FULL, x, 1);
        if (lane == 31) carry = 0xFFFFu;
        if (lane == 0) S.scan_w[warp] = x;
        __syncthreads();
        for (int w = warp + 1; w < MATCH_THREADS / 32; w++) {
            const uint32_t y = S.scan_w[w];
            carry = y < carry ? y : carry;
        }
#pragma unroll
        for (int j = 15; j >= 0; j--) {
            carry = v[j] < carry ? v[j] : carry;
            if (e0 + j <= HASH_SIZE && v[j] == 0xFFFFu) sts_u16(obase + 2 * (e0 + j), carry);
        }
        if (tid == 0) S.off[HASH_SIZE + 1] = (uint16_t)NI;
    }
    const bool ok = __syncthreads_or(bad) == 0;
    STAMP(7);
#undef STAMP
    return ok;
}

/*
 * One CTA per segment: process_loop's match search AND its greedy token
 * decision (src/deflate_compress.c:243-288), parallel yet exact.
 *
 * Index.  The reference threads every position onto the chain of its hash bucket
 * (dup_ht / dup_entries, :64-67, :313-318) and walks a chain newest -> oldest
 * (:249-263).  Here the same buckets are laid out as one array: all positions of
 * [window | segment] sorted by (bucket, position) with a two-pass stable radix
 * sort done by the whole CTA, plus bucket offsets.  Walking a bucket backwards
 * from a position's own slot visits exactly the chain the reference walks, but
 * the next candidate's address never depends on the previous load.
 *
 * Parse.  The longest / nearest match M(p) = argmax_{q in [max(lo, p-window), p-1]}
 * lcp(q, p) (ties -> nearest, capped at min(258, seg_hi - p)) is a pure function of
 * the bytes, because every position is indexed and buckets are walked to the end.
 * Only parse heads need it.  The segment is cut into 32-byte chunks; every lane
 * takes a chunk, parses greedily from the chunk start as if it were a head
 * (:265-288: len >= 3 -> advance len, else literal, advance 1) and KEEPS GOING past
 * the chunk end until it steps on a position another lane has already claimed
 * (shared bitmap, atomicOr) or leaves the segment.  By induction from position 0
 * (claimed by chunk 0's lane) every head of the true greedy parse is claimed and
 * evaluated; speculative parses re-synchronise with the true one within a few
 * tokens, so the extra work is small.  k_parse then follows the evaluated heads.
 */
template <bool HAS_K> /* HAS_K: a chain limit is in force (the hop counters vanish otherwise) */
__global__ void __launch_bounds__(MATCH_THREADS, 1)
k_match(const uint8_t *__restrict__ d_in, long long n, long long avail_hist, long long seg0,
        dfl_params prm, uint32_t *__restrict__ tokcand, uint16_t *__restrict__ scratch, DevState *state)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    MatchSmem &S = *reinterpret_cast<MatchSmem *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long seg = seg0 + blockIdx.x;
    const long long seg_lo = seg * prm.seg_bytes;
    const long long seg_hi = (seg_lo + prm.seg_bytes < n) ? seg_lo + prm.seg_bytes : n;
    const int seglen = (int)(seg_hi - seg_lo);
    long long hist_lo = seg_lo;
    if (prm.use_dict) {
        hist_lo = seg_lo - prm.window;
        if (hist_lo < -avail_hist) hist_lo = -avail_hist;
    }
    const int hist = (int)(seg_lo - hist_lo);
    const int npos = hist + seglen;
    const uint8_t *g0 = d_in + hist_lo;
    const uint32_t soff = 16 + (uint32_t)((uintptr_t)g0 & 15); /* data byte i lives at smem byte soff + i */
    uint8_t *sd = reinterpret_cast<uint8_t *>(S.data);

    const long long t0 = clock64();
    /* ---- stage [window | segment] into shared memory (TMA bulk copy for the aligned interior) ---- */
    {
        int head_n = (int)((16 - ((uintptr_t)g0 & 15)) & 15);
        if (head_n > npos) head_n = npos;
        int bulk_n = (npos - head_n) & ~15;
        int tail0 = head_n + bulk_n;
        if (tid == 0) {
            mbar_init(&S.bar, 1);
            S.next_chunk = 0;
        }
        for (int i = tid; i < DFL_MAX_SEG / 32; i += MATCH_THREADS) S.claimed[i] = 0;
        if (tid < 4) S.data[tid] = 0;
        __syncthreads();
        if (tid == 0 && bulk_n > 0) {
            mbar_expect_tx(&S.bar, (uint32_t)bulk_n);
            for (int o = 0; o < bulk_n; o += 16384) {
                int c = bulk_n - o < 16384 ? bulk_n - o : 16384;
                tma_bulk_g2s(sd + soff + head_n + o, g0 + head_n + o, (uint32_t)c, &S.bar);
            }
        }
        if (tid < head_n) sd[soff + tid] = g0[tid];
        for (int i = tail0 + tid; i < npos; i += MATCH_THREADS) sd[soff + i] = g0[i];
        if (tid < 32) sd[soff + npos + tid] = 0; /* zero slack: over-reads past the data are deterministic */
        if (bulk_n > 0) mbar_wait(&S.bar, 0);
        __syncthreads();
    }
    const long long t1 = clock64();

    /* ---- Adler-32 partial sums of the segment bytes ---- */
    {
        uint32_t a = 0;
        unsigned long long b = 0;
        for (int i = tid; i < seglen; i += MATCH_THREADS) {
            uint32_t v = sd[soff + hist + i];
            a += v;
            b += (unsigned long long)v * (uint32_t)(seglen - i);
        }
        for (int o = 16; o; o >>= 1) {
            a += __shfl_xor_sync(FULL, a, o);
            b += __shfl_xor_sync(FULL, b, o);
        }
        if (lane == 0) { S.red_a[warp] = a; S.red_b[warp] = b; }
    }

    /* ---- index: positions with 3 bytes available, sorted by (bucket, position) ---- */
    const int NI = npos >= 3 ? npos - 2 : 0;
    /* Bucket width.  Splitting the trigram buckets by the fourth byte pays when the trigram buckets are
     * long (text, images, logs).  On data with hardly any repetition (ciphertext, compressed payloads)
     * nearly every head ends in the nearest-trigram sweep over a whole group, and plain trigram buckets
     * -- searched in a single pass -- are cheaper.  Decided per segment from a 1-in-8 sample of the
     * positions: sum of squared trigram-bucket counts / samples = 1 + (candidates a position would see) / 8.
     * The token stream does not depend on it (every chain is walked to the end either way); with a chain
     * limit the buckets are always split so that the limit means one thing. */
    int sb = DFL_SUB_BITS;
    if (!HAS_K) {
        uint32_t *off32 = reinterpret_cast<uint32_t *>(S.off);
        uint32_t obase_ = (uint32_t)__cvta_generic_to_shared(S.off), dbase_ = (uint32_t)__cvta_generic_to_shared(S.data) + soff;
        for (int i2 = tid; i2 < (HASH_SIZE + 2) / 2; i2 += MATCH_THREADS) off32[i2] = 0;
        __syncthreads();
        for (int k8 = tid * 8; k8 < NI; k8 += MATCH_THREADS * 8) {
            const uint32_t b = DFL_HASH3(ld4s(dbase_ + k8));
            reds_add(obase_ + 4 * (b >> 1), 1u << ((b & 1) * 16));
        }
        __syncthreads();
        uint32_t sq = 0;
        for (int i2 = tid; i2 < HASH_SIZE / 2; i2 += MATCH_THREADS) {
            const uint32_t w = off32[i2], lo_ = w & 0xFFFF, hi_ = w >> 16;
            sq += lo_ * lo_ + hi_ * hi_;
        }
        sq = __reduce_add_sync(FULL, sq);
        if (lane == 0) S.scan_w[warp] = sq;
        __syncthreads();
        sq = __reduce_add_sync(FULL, S.scan_w[lane]);
        __syncthreads();
        const uint32_t nsamp = (uint32_t)(NI + 7) / 8;
        if (sq <= SPLIT_MIN_SQ * nsamp) sb = 0;
        if (tid == 0 && sb == 0) atomicAdd(&state->t_phase[20], 1ull); /* segments indexed by plain trigram buckets */
    }
    {
        unsigned int smid;
        asm("mov.u32 %0, %%smid;" : "=r"(smid));
        uint16_t *gbuf = scratch + (size_t)(smid % SCRATCH_SLOTS) * NPOS_MAX; /* one CTA per SM: private slot */
        if (!build_index<true>(S, soff, NI, gbuf, tid, lane, warp, state, sb)) {
            /* the atomics did not hand out slots in lane order: redo with the ballot-ranked scatter */
            if (tid == 0) atomicAdd(&state->sort_fallbacks, 1u);
            build_index<false>(S, soff, NI, gbuf, tid, lane, warp, state, sb);
        }
    }
    const long long t2 = clock64();

    /* ---- speculative chunk parsers ----
     * Run by the first SEARCH_WARPS warps only: chains die as they merge, so with one chunk per lane
     * half the lanes of a warp sit idle behind its slowest chain; fewer lanes that each take several
     * chunks from the shared counter keep the warps' lanes occupied (the SM is issue bound: what counts
     * is useful lanes per instruction, not resident warps). */
    if (warp < SEARCH_WARPS) {
        enum { ST_IDLE = 0, ST_SEEK = 1, ST_WALK = 2, ST_CAND = 3, ST_DONE = 4, ST_FIN = 5, ST_HEAVY = 6 };
        const int W = prm.window, K = HAS_K ? prm.chain_limit : 0;
        const uint32_t ngrp = HASH_SIZE >> sb;
        /* first filter of a walk: 4 bytes when the buckets are split (the walk only looks for matches
         * of >= 4, the trigram sweep does the rest), the trigram otherwise */
        const uint32_t fmask0 = sb ? 0xFFFFFFFFu : 0xFFFFFFu;
        /* shared-window addresses held in registers (the empty asm keeps ptxas from re-deriving them
         * from special registers inside the loops) */
        uint32_t dbase = (uint32_t)__cvta_generic_to_shared(S.data) + soff; /* data byte 0 */
        uint32_t lbase = (uint32_t)__cvta_generic_to_shared(S.list);
        uint32_t obase = (uint32_t)__cvta_generic_to_shared(S.off);
        asm volatile("" : "+r"(dbase), "+r"(lbase), "+r"(obase));
        const int nchunks = (seglen + MATCH_CHUNK - 1) / MATCH_CHUNK;
        uint32_t *out = tokcand + (size_t)blockIdx.x * prm.seg_bytes;
        int st = ST_IDLE;
        int i = 0, p = 0, q = 0, k = 0, k_lo = 0, k_hi = 0, best_len = 2, best_dist = 0, maxlen = 0, hops = 0;
        uint32_t pw0 = 0, pw1 = 0, fword = 0, fmask = 0;
        int foff = 0;

        /* take position i_ as a parse head (caller has claimed it) */
#define INIT_HEAD(i_)                                                                   \
        do {                                                                            \
            i = (i_);                                                                   \
            p = hist + i;                                                               \
            maxlen = seglen - i < DFL_MAX_MATCH ? seglen - i : DFL_MAX_MATCH;           \
            best_len = 2; best_dist = 0; hops = 0;                                      \
            if (maxlen < DFL_MIN_MATCH) { st = ST_DONE; }                               \
            else {                                                                      \
                pw0 = ld4s(dbase + p);                                                  \
                if (sb && maxlen == DFL_MIN_MATCH) { st = ST_DONE; } /* split buckets: only the trigram sweep applies */ \
                else {                                                                  \
                    st = ST_SEEK;                                                       \
                    pw1 = ld4s(dbase + p + 4);                                          \
                    foff = 0; fmask = fmask0; fword = pw0 & fmask0;                     \
                    const uint32_t b_ = hash4(pw0, ngrp, sb);                           \
                    k_lo = (int)lds_u16(obase + 2 * b_);                                \
                    k = k_lo;                                                           \
                    k_hi = (int)lds_u16(obase + 2 * b_ + 2);                            \
                }                                                                       \
            }                                                                           \
        } while (0)

#ifdef DFL_SEARCH_STATS
        unsigned long long sx_iter = 0, sx_walk = 0, sx_fin = 0, sx_done = 0, sx_cand = 0, sx_heads = 0, sx_hops = 0, sx_seek = 0;
#endif
        for (;;) {
#ifdef DFL_SEARCH_STATS
            {
                const uint32_t c1 = __reduce_add_sync(FULL, (st == ST_WALK ? 1u : 0u) | (st == ST_FIN ? 0x100u : 0u) | (st == ST_DONE ? 0x10000u : 0u));
                sx_iter++; sx_walk += c1 & 0xFF; sx_fin += (c1 >> 8) & 0xFF; sx_done += c1 >> 16;
            }
#endif
            /* phase A: walk the bucket backwards until a candidate passes the 4-byte filter.  Without a chain
             * limit two entries are fetched per step: their list and data loads overlap (the loop is bound by
             * the dependent shared-memory latency of a hop, not by issue slots). */
            if (st == ST_WALK) {
                if (HAS_K) {
#pragma unroll 1
                    for (int t = 0; t < HOP_BATCH; t++) {
                        if (k <= k_lo || hops >= K) { st = ST_DONE; break; }
                        q = (int)lds_u16(lbase + 2 * (--k));
                        if (p - q > W) { st = ST_DONE; break; }
                        hops++;
                        /* filter: the 4 bytes ending at offset best_len (fmask0 while nothing is found
                         * yet) must match for q to beat the best so far */
                        if ((ld4s(dbase + q + foff) & fmask) == fword) { st = ST_CAND; break; }
                    }
                } else {
#pragma unroll 1
                    for (int t = 0; t < HOP_BATCH; t += 2) {
                        if (k <= k_lo) { st = ST_DONE; break; }
                        const int q1 = (int)lds_u16(lbase + 2 * (k - 1));
                        const int q2 = k - 1 > k_lo ? (int)lds_u16(lbase + 2 * (k - 2)) : p - W - 1; /* no second entry: reads as "outside the window" */
                        const uint32_t w1 = ld4s(dbase + q1 + foff), w2 = ld4s(dbase + (q2 < 0 ? 0 : q2) + foff);
#ifdef DFL_SEARCH_STATS
                        sx_hops += 2;
#endif
                        const bool o1 = p - q1 > W, h1 = (w1 & fmask) == fword, o2 = p - q2 > W, h2 = (w2 & fmask) == fword;
                        if (o1 | h1 | o2 | h2) { /* one exit; which of the four it was is sorted out off the hot path */
                            if (o1) { st = ST_DONE; k -= 1; }
                            else if (h1) { st = ST_CAND; q = q1; k -= 1; }
                            else if (o2) { st = ST_DONE; k -= 2; }
                            else { st = ST_CAND; q = q2; k -= 2; }
                            break;
                        }
                        k -= 2;
                    }
                }
            }
#ifdef DFL_SEARCH_STATS
            sx_cand += __reduce_add_sync(FULL, st == ST_CAND ? 1u : 0u);
#endif
            /* a long remainder of the bucket is walked by the whole warp together (phase C) */
            if (st == ST_WALK) {
                int rem = k - k_lo;
                if (HAS_K && K - hops < rem) rem = K - hops;
                if (rem > COOP_MIN) st = ST_HEAVY;
            }
            /* phase B: full compare (check_dup_str) for the lanes holding a candidate */
            if (st == ST_CAND) {
                int len;
                const uint32_t x0 = ld4s(dbase + q) ^ pw0, x1 = ld4s(dbase + q + 4) ^ pw1; /* both loads in flight together */
                if (x0) {
                    len = (__ffs(x0) - 1) >> 3;
                } else if (x1) {
                    len = 4 + ((__ffs(x1) - 1) >> 3);
                } else {
                    len = 8;
                    while (len < maxlen) {
                        const uint32_t x = ld4s(dbase + p + len) ^ ld4s(dbase + q + len);
                        if (x) { len += (__ffs(x) - 1) >> 3; break; }
                        len += 4;
                    }
                }
                if (len > maxlen) len = maxlen;
                st = ST_WALK;
                if (len > best_len) {
                    best_len = len;
                    best_dist = p - q;
                    if (len == maxlen) st = ST_DONE;
                    else { foff = len - 3; fmask = 0xFFFFFFFFu; fword = ld4s(dbase + p + foff); }
                }
            }
            /* phase C: finished heads emit their token and move on; idle lanes take new
             * chunks.  One warp reduction per round steers it (votes and shuffles share a
             * slow unit on this chip). */
            const uint32_t census = __reduce_add_sync(FULL, (st == ST_WALK ? 1u : 0u) | (st == ST_FIN ? 0x100u : 0u) |
                                                                (st == ST_HEAVY ? 0x10000u : 0u));
            const int n_busy = census & 0xFF, n_fin = (census >> 8) & 0xFF, n_heavy = census >> 16;
            if (n_fin == 32) break;
            if (n_heavy) {
                /* cooperative walk: the 32 lanes take 32 consecutive bucket entries per pass,
                 * nearest first; one warp reduction per pass picks the longest (then nearest)
                 * improvement and tightens the filter for the next pass.  Same result as the
                 * sequential walk: strict improvement only, nearer passes first. */
                uint32_t heavy = __ballot_sync(FULL, st == ST_HEAVY);
                while (heavy) {
                    const int L = __ffs(heavy) - 1;
                    heavy &= heavy - 1;
                    const int hp = __shfl_sync(FULL, p, L), hk = __shfl_sync(FULL, k, L);
                    int hlo = __shfl_sync(FULL, k_lo, L);
                    const int hmax = __shfl_sync(FULL, maxlen, L);
                    const uint32_t hw0 = __shfl_sync(FULL, pw0, L), hw1 = __shfl_sync(FULL, pw1, L);
                    const int hhops = HAS_K ? __shfl_sync(FULL, hops, L) : 0;
                    if (HAS_K && hk - (K - hhops) > hlo) hlo = hk - (K - hhops);
                    int hbest = __shfl_sync(FULL, best_len, L), hdist = __shfl_sync(FULL, best_dist, L);
                    int hfoff = hbest == 2 ? 0 : hbest - 3;
                    uint32_t hfmask = hbest == 2 ? fmask0 : 0xFFFFFFFFu;
                    uint32_t hfword = ld4s(dbase + hp + hfoff) & hfmask;
                    for (int top = hk; top > hlo; top -= 32) {
                        const int kk = top - 1 - lane;
                        bool valid = kk >= hlo;
                        const int cq = valid ? (int)lds_u16(lbase + 2 * kk) : 0;
                        valid = valid && (hp - cq <= W);
                        uint32_t key = valid ? 0x80000000u : 0u;
                        if (valid && (ld4s(dbase + cq + hfoff) & hfmask) == hfword) {
                            int len;
                            uint32_t x = ld4s(dbase + cq) ^ hw0;
                            if (x) {
                                len = (__ffs(x) - 1) >> 3;
                            } else {
                                x = ld4s(dbase + cq + 4) ^ hw1;
                                if (x) {
                                    len = 4 + ((__ffs(x) - 1) >> 3);
                                } else {
                                    len = 8;
                                    while (len < hmax) {
                                        x = ld4s(dbase + hp + len) ^ ld4s(dbase + cq + len);
                                        if (x) { len += (__ffs(x) - 1) >> 3; break; }
                                        len += 4;
                                    }
                                }
                            }
                            if (len > hmax) len = hmax;
                            if (len > hbest) key |= ((uint32_t)len << 16) | (uint32_t)(0xFFFF - (hp - cq));
                        }
                        const uint32_t kmax = __reduce_max_sync(FULL, key);
                        if (!(kmax >> 31)) break; /* the whole pass is outside the window */
                        if (kmax & 0x7FFFFFFFu) {
                            hbest = (kmax >> 16) & 0x7FFF;
                            hdist = 0xFFFF - (kmax & 0xFFFF);
                            if (hbest == hmax) break;
                            hfoff = hbest - 3; hfmask = 0xFFFFFFFFu; hfword = ld4s(dbase + hp + hfoff);
                        }
                    }
                    if (lane == L) { best_len = hbest; best_dist = hdist; st = ST_DONE; }
                }
            }
            if (32 - n_busy - n_fin >= REFILL_MIN_IDLE || n_busy == 0) {
                /* Heads whose own bucket held no match of >= 4 bytes: the answer is the NEAREST earlier
                 * occurrence of the trigram (all remaining matches have length exactly 3, :258 keeps the
                 * first found), which lives somewhere in the 2^DFL_SUB_BITS buckets of the trigram's group.
                 * The group is one contiguous range of the list: the warp sweeps it flat, one head at a
                 * time, and keeps the largest matching position inside the window. */
                uint32_t need = sb ? __ballot_sync(FULL, st == ST_DONE && best_len == 2 && maxlen >= DFL_MIN_MATCH) : 0u;
                while (need) {
                    const int L = __ffs(need) - 1;
                    need &= need - 1;
                    const int hp = __shfl_sync(FULL, p, L);
                    const uint32_t hw = __shfl_sync(FULL, pw0, L) & 0xFFFFFFu;
                    const uint32_t g_ = hash4(hw, ngrp, sb); /* fourth byte 0 -> first bucket of the group */
                    const int glo = (int)lds_u16(obase + 2 * g_), ghi = (int)lds_u16(obase + 2 * (g_ + (1 << DFL_SUB_BITS)));
                    uint32_t loc = 0;
                    for (int kk = ghi - 1 - lane; kk >= glo; kk -= 64) { /* two steps' loads in flight together */
                        const int kb = kk - 32;
                        const int cq = (int)lds_u16(lbase + 2 * kk), cq2 = kb >= glo ? (int)lds_u16(lbase + 2 * kb) : hp;
                        const int d = hp - cq, d2 = hp - cq2;
                        const bool in1 = (unsigned)(d - 1) < (unsigned)W, in2 = (unsigned)(d2 - 1) < (unsigned)W;
                        const uint32_t t1 = in1 ? ld4s(dbase + cq) & 0xFFFFFFu : ~hw, t2 = in2 ? ld4s(dbase + cq2) & 0xFFFFFFu : ~hw;
                        if (t1 == hw && (uint32_t)cq + 1 > loc) loc = (uint32_t)cq + 1;
                        if (t2 == hw && (uint32_t)cq2 + 1 > loc) loc = (uint32_t)cq2 + 1;
                    }
                    loc = __reduce_max_sync(FULL, loc);
                    if (lane == L && loc) { best_len = DFL_MIN_MATCH; best_dist = p - (int)(loc - 1); }
                }
                if (st == ST_DONE) {
                    int adv = 1;
                    uint32_t t = lds_u8(dbase + p);
                    if (best_len >= DFL_MIN_MATCH) {
                        t = ((uint32_t)best_dist << 16) | (uint32_t)best_len;
                        adv = best_len;
                    }
                    out[i] = t;
#ifdef DFL_SEARCH_STATS
                    sx_heads++;
#endif
                    st = ST_IDLE;
                    const int ni = i + adv;
                    if (ni < seglen) {
                        const uint32_t bit = 1u << (ni & 31);
                        if (!(atomicOr(&S.claimed[ni >> 5], bit) & bit)) INIT_HEAD(ni);
                    }
                }
                if (st == ST_IDLE) {
                    const int c = atomicAdd(&S.next_chunk, 1);
                    if (c >= nchunks) {
                        st = ST_FIN;
                    } else {
                        const int ci = c * MATCH_CHUNK;
                        const uint32_t bit = 1u << (ci & 31);
                        /* a chain that has already claimed a head inside this chunk covers it (it runs until
                         * it meets a claimed head); starting here too would only add speculative heads */
                        const bool entered = SKIP_OVERRUN && MATCH_CHUNK == 32 && S.claimed[ci >> 5] != 0;
                        if (!entered && !(atomicOr(&S.claimed[ci >> 5], bit) & bit)) INIT_HEAD(ci);
                    }
                }
                /* the new heads find their own slot in their bucket together: lower bound of p in
                 * list[k, k_hi); candidates are then k-1 ... k_lo */
                if (st == ST_SEEK) {
#ifdef DFL_SEARCH_STATS
                    sx_seek++;
#endif
                    while (k < k_hi) {
                        const int mid = (k + k_hi) >> 1;
                        if ((int)lds_u16(lbase + 2 * mid) < p) k = mid + 1; else k_hi = mid;
                    }
                    st = ST_WALK;
                }
            }
        }
#ifdef DFL_SEARCH_STATS
        sx_heads = __reduce_add_sync(FULL, (uint32_t)sx_heads);
        sx_hops = __reduce_add_sync(FULL, (uint32_t)sx_hops);
        sx_seek = __reduce_add_sync(FULL, (uint32_t)sx_seek);
        if (lane == 0) {
            atomicAdd(&state->t_phase[12], sx_iter); atomicAdd(&state->t_phase[13], sx_walk);
            atomicAdd(&state->t_phase[14], sx_fin); atomicAdd(&state->t_phase[15], sx_done);
            atomicAdd(&state->t_phase[16], sx_cand); atomicAdd(&state->t_phase[17], sx_heads);
            atomicAdd(&state->t_phase[18], sx_hops); atomicAdd(&state->t_phase[19], sx_seek);
            S.scan_w[warp] = (uint32_t)(clock64() - t2); /* when this warp ran out of work */
        }
#endif
#undef INIT_HEAD
    }
    __syncthreads();
#ifdef DFL_SEARCH_STATS
    if (warp == 0 && warp < SEARCH_WARPS) {
        const uint32_t e = lane < SEARCH_WARPS ? S.scan_w[lane] : 0u;
        const uint32_t mx = __reduce_max_sync(FULL, e), sm = __reduce_add_sync(FULL, e);
        if (lane == 0) { atomicAdd(&state->t_phase[21], (unsigned long long)(sm / SEARCH_WARPS)); atomicAdd(&state->t_phase[22], (unsigned long long)mx); }
    }
#endif
    if (warp == 0) {
        uint32_t a = S.red_a[lane];
        unsigned long long b = S.red_b[lane];
        for (int o = 16; o; o >>= 1) {
            a += __shfl_xor_sync(FULL, a, o);
            b += __shfl_xor_sync(FULL, b, o);
        }
        if (lane == 0) {
            /* s1 += A ; s2 += B + A * (bytes after this segment)   (mod 65521) */
            unsigned long long am = a % ADLER_MOD;
            unsigned long long rest = (unsigned long long)(n - seg_hi) % ADLER_MOD;
            unsigned long long bm = (b % ADLER_MOD + am * rest) % ADLER_MOD;
            atomicAdd(&state->adl_a, am);
            atomicAdd(&state->adl_b, bm);
            const long long t3 = clock64();
            atomicAdd(&state->t_phase[0], (unsigned long long)(t1 - t0));
            atomicAdd(&state->t_phase[1], (unsigned long long)(t2 - t1));
            atomicAdd(&state->t_phase[2], (unsigned long long)(t3 - t2));
            atomicAdd(&state->t_phase[3], (unsigned long long)(t3 - t0));
        }
    }
}

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
        while (c < lim) {
            heads |= 1u << c;
            c += __shfl_sync(FULL, adv, c);
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

/* ------------------------------------------------------------------------- */
/* k_huff                                                                    */
/* ------------------------------------------------------------------------- */

#define HUFF_WARPS 4

struct HuffScratch {
    uint32_t w[LLS];       /* padded weights, later internal node weights */
    uint32_t sw[LLS];      /* sorted weights */
    uint16_t ssym[LLS];    /* sorted symbols */
    uint16_t lpar[LLS], ipar[LLS], idepth[LLS];
    uint32_t iw[LLS];
};

struct HuffSmem {
    uint32_t ll_freq[LLS], d_freq[DS], cl_freq[32];
    uint8_t ll_len[LLS], d_len[DS], cl_len[32];
    uint16_t ll_code[LLS], d_code[DS], cl_code[32];
    uint8_t seq[LLS + DS + 8];
    uint8_t item_sym[LLS + DS + 8], item_extra[LLS + DS + 8];
    uint32_t hdr[HDR_WORDS];
    HuffScratch sc;
};

__constant__ uint8_t c_cl_order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

/*
 * Length-limited Huffman code lengths, executed by one warp.  Bit-exact twin of
 * orc_huff_lengths (oracle/deflate_oracle.c): used symbols padded to two, sort by
 * (weight, symbol), two-queue merge with ties to the leaf queue (the
 * construction of h_tree_builder_build, src/h_tree.c:242-280), leaf count per
 * depth clipped at `limit` with Kraft repair, lengths handed out along the
 * sorted order.
 */
__device__ void huff_lengths_warp(const uint32_t *freq, int n, int limit, uint8_t *len, HuffScratch &sc, int lane)
{
    int m = 0;
    for (int s = lane; s < n; s += 32) {
        uint32_t f = freq[s];
        sc.w[s] = f;
        len[s] = 0;
        m += f != 0;
    }
    for (int o = 16; o; o >>= 1) m += __shfl_xor_sync(FULL, m, o);
    __syncwarp();
    if (m < 2) {
        if (lane == 0) {
            for (int s = 0; m < 2 && s < n; s++)
                if (!sc.w[s]) { sc.w[s] = 1; m++; }
        }
        m = 2;
        __syncwarp();
    }
    /* rank sort by key = weight << 9 | symbol (keys are distinct; weights are token counts < 2^16).
     * The used symbols are compacted first (sc.iw is free until the merge): m, not n, keys to rank. */
    {
        const uint32_t lt = lanemask_lt();
        int base = 0;
        for (int s0 = 0; s0 < n; s0 += 32) {
            const int s = s0 + lane;
            const uint32_t ws = s < n ? sc.w[s] : 0u;
            const uint32_t mask = __ballot_sync(FULL, ws != 0);
            if (ws) sc.iw[base + __popc(mask & lt)] = (ws << 9) | (uint32_t)s;
            base += __popc(mask);
        }
        __syncwarp();
        for (int r0 = lane; r0 < m; r0 += 32) {
            const uint32_t key = sc.iw[r0];
            int r = 0;
            for (int j = 0; j < m; j++) r += sc.iw[j] < key;
            sc.sw[r] = key >> 9;
            sc.ssym[r] = (uint16_t)(key & 511u);
        }
    }
    __syncwarp();
    if (lane == 0) {
        int li = 0, ih = 0, it = 0;
        while (it < m - 1) {
            uint32_t sum = 0;
#pragma unroll
            for (int k = 0; k < 2; k++) {
                if (li < m && (ih >= it || sc.sw[li] <= sc.iw[ih])) { sum += sc.sw[li]; sc.lpar[li] = (uint16_t)it; li++; }
                else { sum += sc.iw[ih]; sc.ipar[ih] = (uint16_t)it; ih++; }
            }
            sc.iw[it++] = sum;
        }
        sc.idepth[m - 2] = 0;
        for (int i = m - 3; i >= 0; i--) sc.idepth[i] = sc.idepth[sc.ipar[i]] + 1;
        int bl[16];
#pragma unroll
        for (int i = 0; i < 16; i++) bl[i] = 0;
        for (int i = 0; i < m; i++) {
            int d = sc.idepth[sc.lpar[i]] + 1;
            bl[d > limit ? limit : d]++;
        }
        /* Kraft repair in units of 2^-limit (same steps as the oracle) */
        unsigned int kraft = 0;
        for (int bits = 1; bits <= limit; bits++) kraft += (unsigned int)bl[bits] << (limit - bits);
        while (kraft > (1u << limit)) {
            bl[limit]--;
            for (int bits = limit - 1; bits > 0; bits--) {
                if (bl[bits]) {
                    bl[bits]--;
                    bl[bits + 1] += 2;
                    break;
                }
            }
            kraft--;
        }
        int j = 0;
        for (int bits = limit; bits >= 1; bits--)
            for (int i = 0; i < bl[bits]; i++) len[sc.ssym[j++]] = (uint8_t)bits;
    }
    __syncwarp();
}

/* Canonical codes (RFC 1951 3.2.2), bit-reversed for the LSB-first packer; lane 0. */
__device__ void canonical_codes_lane0(const uint8_t *len, int n, uint16_t *code)
{
    int bl[16], next[16];
#pragma unroll
    for (int i = 0; i < 16; i++) bl[i] = 0;
    for (int s = 0; s < n; s++) bl[len[s]]++;
    bl[0] = 0;
    int c = 0;
    for (int b = 1; b <= 15; b++) {
        c = (c + bl[b - 1]) << 1;
        next[b] = c;
    }
    for (int s = 0; s < n; s++) {
        int l = len[s];
        code[s] = l ? (uint16_t)(__brev((uint32_t)next[l]++) >> (32 - l)) : 0;
    }
}

struct BitW {
    uint32_t *w;
    uint32_t nbits;
    __device__ void put(uint32_t v, int n)
    {
        uint32_t i = nbits >> 5, sh = nbits & 31;
        w[i] |= v << sh;
        if (sh + n > 32) w[i + 1] |= v >> (32 - sh);
        nbits += n;
    }
};

/*
 * One warp per segment: code construction + exact block costs + choice.
 * Twin of orc_emit_segment's front half (oracle/deflate_oracle.c); what the
 * reference accounts per token in h_tree_d_lens / h_tree_builder_*
 * (src/h_tree.c:75-148, :242-302) is done once per block here.
 */
__global__ void __launch_bounds__(HUFF_WARPS * 32)
k_huff(long long n, long long seg0, int nseg, dfl_params prm, SegArrays A)
{
    __shared__ HuffSmem s_all[HUFF_WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sl = blockIdx.x * HUFF_WARPS + warp;
    if (sl >= nseg) return;
    HuffSmem &S = s_all[warp];
    const long long seg_lo = (seg0 + sl) * prm.seg_bytes;
    const long long seglen = ((seg_lo + prm.seg_bytes < n) ? seg_lo + prm.seg_bytes : n) - seg_lo;

    for (int i = lane; i < LLS; i += 32) S.ll_freq[i] = i < NLL ? A.ll_freq[(size_t)sl * LLS + i] : 0;
    S.d_freq[lane] = lane < ND ? A.d_freq[(size_t)sl * DS + lane] : 0;
    S.cl_freq[lane] = 0;
    for (int i = lane; i < HDR_WORDS; i += 32) S.hdr[i] = 0;
    __syncwarp();

    huff_lengths_warp(S.ll_freq, NLL, 15, S.ll_len, S.sc, lane);
    huff_lengths_warp(S.d_freq, ND, 15, S.d_len, S.sc, lane);

    int hlit = 0, hdist = 0, hclen = 0, nitems = 0;
    if (lane == 0) {
        for (hlit = NLL; hlit > 257 && S.ll_len[hlit - 1] == 0; hlit--) {}
        for (hdist = ND; hdist > 1 && S.d_len[hdist - 1] == 0; hdist--) {}
        int nseq = hlit + hdist;
        for (int i = 0; i < hlit; i++) S.seq[i] = S.ll_len[i];
        for (int i = 0; i < hdist; i++) S.seq[hlit + i] = S.d_len[i];
        /* RFC 1951 3.2.7 run-length coding (what h_tree_d_lens counts, src/h_tree.c:87-136) */
        int i = 0, k = 0;
        while (i < nseq) {
            int v = S.seq[i], j = i + 1;
            while (j < nseq && S.seq[j] == v) j++;
            int run = j - i;
            if (v == 0) {
                while (run >= 11) {
                    int r = run < 138 ? run : 138;
                    S.item_sym[k] = 18; S.item_extra[k] = (uint8_t)(r - 11); k++;
                    run -= r;
                }
                if (run >= 3) { S.item_sym[k] = 17; S.item_extra[k] = (uint8_t)(run - 3); k++; run = 0; }
                while (run-- > 0) { S.item_sym[k] = 0; S.item_extra[k] = 0; k++; }
            } else {
                S.item_sym[k] = (uint8_t)v; S.item_extra[k] = 0; k++;
                run--;
                while (run >= 3) {
                    int r = run < 6 ? run : 6;
                    S.item_sym[k] = 16; S.item_extra[k] = (uint8_t)(r - 3); k++;
                    run -= r;
                }
                while (run-- > 0) { S.item_sym[k] = (uint8_t)v; S.item_extra[k] = 0; k++; }
            }
            i = j;
        }
        nitems = k;
        for (i = 0; i < nitems; i++) S.cl_freq[S.item_sym[i]]++;
    }
    __syncwarp();
    huff_lengths_warp(S.cl_freq, NCL, 7, S.cl_len, S.sc, lane);

    /* exact costs */
    unsigned long long dyn = 0, fix = 0;
    for (int s = lane; s < NLL; s += 32) {
        dyn += (unsigned long long)S.ll_freq[s] * S.ll_len[s];
        fix += (unsigned long long)S.ll_freq[s] * fixed_ll_len(s);
    }
    if (lane < ND) {
        dyn += (unsigned long long)S.d_freq[lane] * S.d_len[lane];
        fix += (unsigned long long)S.d_freq[lane] * 5;
    }
    if (lane < NCL) {
        dyn += (unsigned long long)S.cl_freq[lane] * S.cl_len[lane];
        if (lane == 16) dyn += 2ull * S.cl_freq[16];
        if (lane == 17) dyn += 3ull * S.cl_freq[17];
        if (lane == 18) dyn += 7ull * S.cl_freq[18];
    }
    for (int o = 16; o; o >>= 1) {
        dyn += __shfl_xor_sync(FULL, dyn, o);
        fix += __shfl_xor_sync(FULL, fix, o);
    }
    int btype = 0;
    unsigned long long block_bits = 0;
    if (lane == 0) {
        for (hclen = 19; hclen > 4 && S.cl_len[c_cl_order[hclen - 1]] == 0; hclen--) {}
        unsigned long long extra = A.extra[sl];
        dyn += 3 + 5 + 5 + 4 + 3 * hclen + extra;
        fix += 3 + extra;
        unsigned long long nstored = (unsigned long long)((seglen + 65534) / 65535);
        if (nstored < 1) nstored = 1;
        unsigned long long stored = 8ull * ((unsigned long long)seglen + 5 * nstored);
        if (prm.block_mode == 1) btype = 0;
        else if (prm.block_mode == 2) btype = 1;
        else if (prm.block_mode == 3) btype = 2;
        else if (stored <= fix && stored <= dyn) btype = 0;
        else if (fix <= dyn) btype = 1;
        else btype = 2;
        block_bits = btype == 0 ? stored : btype == 1 ? fix : dyn;

        BitW bw{S.hdr, 0};
        if (btype == 2) {
            canonical_codes_lane0(S.ll_len, NLL, S.ll_code);
            canonical_codes_lane0(S.d_len, ND, S.d_code);
            canonical_codes_lane0(S.cl_len, NCL, S.cl_code);
            bw.put(0, 1);
            bw.put(2, 2);
            bw.put(hlit - 257, 5);
            bw.put(hdist - 1, 5);
            bw.put(hclen - 4, 4);
            for (int i = 0; i < hclen; i++) bw.put(S.cl_len[c_cl_order[i]], 3);
            for (int i = 0; i < nitems; i++) {
                int cs = S.item_sym[i];
                bw.put(S.cl_code[cs], S.cl_len[cs]);
                if (cs == 16) bw.put(S.item_extra[i], 2);
                else if (cs == 17) bw.put(S.item_extra[i], 3);
                else if (cs == 18) bw.put(S.item_extra[i], 7);
            }
        } else if (btype == 1) {
            bw.put(0, 1);
            bw.put(1, 2);
        }
        A.hdr_bits[sl] = bw.nbits;
        A.btype[sl] = btype;
        A.block_bits[sl] = block_bits;
        A.out_bytes[sl] = btype == 0 ? (uint32_t)(seglen + 5 * nstored + 5)
                                     : (uint32_t)((block_bits + 3 + 7) / 8 + 4);
    }
    btype = __shfl_sync(FULL, btype, 0);
    __syncwarp();
    if (btype == 2) {
        for (int s = lane; s < LLS; s += 32)
            A.lltab[(size_t)sl * LLS + s] = s < NLL ? (uint32_t)S.ll_code[s] | ((uint32_t)S.ll_len[s] << 16) : 0;
        A.dtab[(size_t)sl * DS + lane] = lane < ND ? (uint32_t)S.d_code[lane] | ((uint32_t)S.d_len[lane] << 16) : 0;
        for (int i = lane; i < HDR_WORDS; i += 32) A.hdr[(size_t)sl * HDR_WORDS + i] = S.hdr[i];
    } else if (btype == 1) {
        /* fixed code of RFC 1951 3.2.6 */
        for (int s = lane; s < LLS; s += 32) {
            int l = fixed_ll_len(s);
            uint32_t code = s < 144 ? 0x30 + s : s < 256 ? 0x190 + (s - 144) : s < 280 ? (s - 256) : 0xC0 + (s - 280);
            A.lltab[(size_t)sl * LLS + s] = (__brev(code) >> (32 - l)) | ((uint32_t)l << 16);
        }
        A.dtab[(size_t)sl * DS + lane] = (__brev((uint32_t)lane) >> 27) | (5u << 16);
        if (lane == 0) A.hdr[(size_t)sl * HDR_WORDS] = S.hdr[0];
    }
}

/* ------------------------------------------------------------------------- */
/* k_scan                                                                    */
/* ------------------------------------------------------------------------- */

#define SCAN_THREADS 1024

/* Exclusive scan of the segment sizes of one chunk -> byte offsets in the body. */
__global__ void __launch_bounds__(SCAN_THREADS)
k_scan(int nseg, SegArrays A, DevState *state)
{
    __shared__ unsigned long long s_warp[32];
    __shared__ unsigned long long s_carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_carry = state->total_out;
    __syncthreads();
    for (int base = 0; base < nseg; base += SCAN_THREADS) {
        int i = base + tid;
        unsigned long long v = i < nseg ? A.out_bytes[i] : 0, x = v;
        for (int o = 1; o < 32; o <<= 1) {
            unsigned long long y = __shfl_up_sync(FULL, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) s_warp[warp] = x;
        __syncthreads();
        if (warp == 0) {
            unsigned long long t = s_warp[lane], u = t;
            for (int o = 1; o < 32; o <<= 1) {
                unsigned long long y = __shfl_up_sync(FULL, u, o);
                if (lane >= o) u += y;
            }
            s_warp[lane] = u - t;
        }
        __syncthreads();
        unsigned long long excl = s_carry + s_warp[warp] + x - v;
        if (i < nseg) A.out_off[i] = excl;
        __syncthreads();
        if (tid == SCAN_THREADS - 1) s_carry = excl + v;
        __syncthreads();
    }
    if (tid == 0) state->total_out = s_carry;
}

/* ------------------------------------------------------------------------- */
/* k_emit                                                                    */
/* ------------------------------------------------------------------------- */

#define EMIT_THREADS 256
#define EMIT_TPT 4
#define EMIT_ROUND (EMIT_THREADS * EMIT_TPT)
#define STG_WORDS 2048 /* 8 KB: 16 B shift + 640 B header + 1024 tokens * 48 bits + EOB/sync */

__device__ __forceinline__ void stg_put(uint32_t *stg, uint32_t bitpos, unsigned long long v, int nb)
{
    if (nb == 0) return;
    uint32_t i = bitpos >> 5, sh = bitpos & 31;
    unsigned long long lo = v << sh;
    atomicOr(&stg[i], (uint32_t)lo);
    if (sh + nb > 32) atomicOr(&stg[i + 1], (uint32_t)(lo >> 32));
    if (sh + nb > 64) atomicOr(&stg[i + 2], (uint32_t)(v >> (64 - sh)));
}

/*
 * One CTA per segment: the bit packer the reference never wrote
 * (src/deflate_compress.c:269,:279 "TODO: write lit / len/dist pair").
 * Per round of 1024 tokens: each thread maps 4 tokens to (bits, nbits), a block
 * exclusive scan gives every token its bit offset, codes are OR-ed into a shared
 * staging buffer whose origin is congruent (mod 16 B) with the output address,
 * and whole 16-byte words are stored to HBM.
 */
__global__ void __launch_bounds__(EMIT_THREADS)
k_emit(const uint8_t *__restrict__ d_in, long long n, long long seg0, long long last_seg, int is_final,
       dfl_params prm, const uint32_t *__restrict__ tokcand, SegArrays A, uint8_t *__restrict__ d_out,
       unsigned long long out_cap, DevState *state)
{
    __shared__ __align__(16) uint32_t stg[STG_WORDS];
    __shared__ uint32_t s_ll[LLS], s_d[DS];
    __shared__ uint32_t s_wsum[EMIT_THREADS / 32];
    __shared__ uint32_t s_total;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int sl = blockIdx.x;
    const long long seg = seg0 + sl;
    const long long seg_lo = seg * prm.seg_bytes;
    const long long seglen = ((seg_lo + prm.seg_bytes < n) ? seg_lo + prm.seg_bytes : n) - seg_lo;
    const unsigned long long off = A.out_off[sl];
    const uint32_t out_bytes = A.out_bytes[sl];
    const int btype = (int)A.btype[sl];
    const uint32_t final_bit = (is_final && seg == last_seg) ? 1u : 0u;
    if (off + out_bytes > out_cap) {
        if (tid == 0) atomicExch(&state->error, 1u);
        return;
    }
    uint8_t *dst = d_out + off;

    if (btype == 0) {
        /* stored block(s), RFC 1951 3.2.4, then the sync marker */
        const uint8_t *src = d_in + seg_lo;
        long long nst = (seglen + 65534) / 65535;
        if (nst < 1) nst = 1;
        for (long long b = 0; b < nst; b++) {
            long long o = b * 65535, len = seglen - o < 65535 ? seglen - o : 65535;
            uint8_t *q = dst + o + 5 * b;
            if (tid == 0) {
                q[0] = 0;
                q[1] = (uint8_t)len; q[2] = (uint8_t)(len >> 8);
                q[3] = (uint8_t)~len; q[4] = (uint8_t)(~len >> 8);
            }
            for (long long i = tid; i < len; i += EMIT_THREADS) q[5 + i] = src[o + i];
        }
        if (tid == 0) {
            uint8_t *q = dst + seglen + 5 * nst;
            q[0] = (uint8_t)final_bit; q[1] = 0; q[2] = 0; q[3] = 0xFF; q[4] = 0xFF;
        }
        return;
    }

    for (int i = tid; i < LLS; i += EMIT_THREADS) s_ll[i] = A.lltab[(size_t)sl * LLS + i];
    if (tid < DS) s_d[tid] = A.dtab[(size_t)sl * DS + tid];
    const uint32_t *tc = tokcand + (size_t)sl * prm.seg_bytes;
    const int ntok = (int)A.ntok[sl];
    const uint32_t hdr_bits = A.hdr_bits[sl];
    const uint32_t *hdr = A.hdr + (size_t)sl * HDR_WORDS;

    unsigned long long gpos = 0; /* bytes of this segment already stored */
    uint32_t carry_n = 0, carry_v = 0;
    const int nround = ntok == 0 ? 1 : (ntok + EMIT_ROUND - 1) / EMIT_ROUND;
    for (int r = 0; r < nround; r++) {
        const uint32_t sh = (uint32_t)((uintptr_t)(dst + gpos) & 15);
        for (int i = tid; i < STG_WORDS; i += EMIT_THREADS) stg[i] = 0;
        __syncthreads();
        uint32_t bit0 = sh * 8 + carry_n;
        if (tid == 0 && carry_n) atomicOr(&stg[sh >> 2], carry_v << ((sh & 3) * 8));
        if (r == 0) {
            for (uint32_t i = tid; i * 32 < hdr_bits; i += EMIT_THREADS) {
                uint32_t nb = hdr_bits - i * 32 < 32 ? hdr_bits - i * 32 : 32;
                stg_put(stg, bit0 + i * 32, hdr[i], (int)nb);
            }
            bit0 += hdr_bits;
        }
        /* map this thread's tokens */
        unsigned long long cb[EMIT_TPT];
        int cn[EMIT_TPT];
        uint32_t tsum = 0;
        const int t0 = r * EMIT_ROUND + tid * EMIT_TPT;
#pragma unroll
        for (int j = 0; j < EMIT_TPT; j++) {
            cb[j] = 0; cn[j] = 0;
            if (t0 + j < ntok) {
                uint32_t t = tc[t0 + j];
                if ((t >> 16) == 0) {
                    uint32_t e = s_ll[t & 0xFF];
                    cb[j] = e & 0xFFFF; cn[j] = e >> 16;
                } else {
                    int eb, ev;
                    uint32_t e = s_ll[len_sym((int)(t & 0xFFFF), eb, ev)];
                    unsigned long long bits = e & 0xFFFF;
                    int nb = e >> 16;
                    bits |= (unsigned long long)ev << nb; nb += eb;
                    e = s_d[dist_sym((int)(t >> 16), eb, ev)];
                    bits |= (unsigned long long)(e & 0xFFFF) << nb; nb += e >> 16;
                    bits |= (unsigned long long)ev << nb; nb += eb;
                    cb[j] = bits; cn[j] = nb;
                }
                tsum += cn[j];
            }
        }
        /* block exclusive scan of tsum */
        uint32_t x = tsum;
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t y = __shfl_up_sync(FULL, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) s_wsum[warp] = x;
        __syncthreads();
        if (tid == 0) {
            uint32_t acc = 0;
            for (int w = 0; w < EMIT_THREADS / 32; w++) { uint32_t t = s_wsum[w]; s_wsum[w] = acc; acc += t; }
            s_total = acc;
        }
        __syncthreads();
        uint32_t bp = bit0 + s_wsum[warp] + x - tsum;
#pragma unroll
        for (int j = 0; j < EMIT_TPT; j++) {
            stg_put(stg, bp, cb[j], cn[j]);
            bp += cn[j];
        }
        uint32_t endbit = bit0 + s_total;
        const bool last = (r == nround - 1);
        if (last && tid == 0) {
            /* end of block, then the empty stored block that byte-aligns the segment */
            uint32_t e = s_ll[256];
            stg_put(stg, endbit, e & 0xFFFF, (int)(e >> 16));
            uint32_t eb2 = endbit + (e >> 16);
            stg_put(stg, eb2, final_bit, 3);
            uint32_t bytepos = (eb2 + 3 + 7) >> 3;
            /* LEN = 0x0000, NLEN = 0xFFFF */
            stg_put(stg, bytepos * 8 + 16, 0xFFFF, 16);
        }
        if (last) {
            uint32_t e = s_ll[256];
            endbit = (((endbit + (e >> 16) + 3 + 7) >> 3) + 4) * 8;
        }
        __syncthreads();
        /* flush whole bytes [sh, e) */
        const uint32_t e_byte = endbit >> 3;
        const uint8_t *sb = reinterpret_cast<const uint8_t *>(stg);
        uint8_t *gbase = dst + gpos - sh; /* 16 B aligned */
        {
            const uint32_t a0 = (sh + 15) & ~15u; /* first 16 B boundary >= sh */
            const uint32_t a1 = e_byte & ~15u;    /* last 16 B boundary <= e */
            if (a1 < a0) {
                for (uint32_t i = sh + tid; i < e_byte; i += EMIT_THREADS) gbase[i] = sb[i];
            } else {
                for (uint32_t i = sh + tid; i < a0; i += EMIT_THREADS) gbase[i] = sb[i];
                for (uint32_t i = a0 / 16 + tid; i < a1 / 16; i += EMIT_THREADS)
                    reinterpret_cast<uint4 *>(gbase)[i] = reinterpret_cast<const uint4 *>(stg)[i];
                for (uint32_t i = a1 + tid; i < e_byte; i += EMIT_THREADS) gbase[i] = sb[i];
            }
        }
        carry_n = endbit & 7;
        carry_v = carry_n ? (sb[e_byte] & ((1u << carry_n) - 1)) : 0;
        gpos += e_byte - sh;
        __syncthreads();
    }
}

/* ------------------------------------------------------------------------- */
/* k_final + synth                                                           */
/* ------------------------------------------------------------------------- */

__global__ void k_final(long long n, DevState *state)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        unsigned long long s1 = (1 + state->adl_a) % ADLER_MOD;
        unsigned long long s2 = ((unsigned long long)n % ADLER_MOD + state->adl_b) % ADLER_MOD;
        state->adler = (unsigned int)((s2 << 16) | s1);
    }
}

__device__ __host__ __forceinline__ unsigned long long splitmix64_at(unsigned long long seed, unsigned long long i)
{
    unsigned long long z = seed + (i + 1) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__global__ void k_synth_random(uint8_t *out, unsigned long long n, unsigned long long seed)
{
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (; i * 8 < n; i += stride) {
        unsigned long long z = splitmix64_at(seed, i);
        if (i * 8 + 8 <= n && (((uintptr_t)out & 7) == 0)) {
            reinterpret_cast<unsigned long long *>(out)[i] = z;
        } else {
            for (int k = 0; k < 8 && i * 8 + k < n; k++) out[i * 8 + k] = (uint8_t)(z >> (8 * k));
        }
    }
}

/* ------------------------------------------------------------------------- */
/* k_inflate                                                                 */
/* ------------------------------------------------------------------------- */

/*
 * Segment-parallel inflate of a stream this library produced (the CPU
 * twin is host/deflate_decompress.c, the reference's is src/deflate_decompress.c:146-409).  Every
 * segment starts byte aligned at a known offset (the segment index k_scan computed), so the Huffman
 * decoding of all segments runs in parallel, one warp per segment.  The LZ77 copies do not: with
 * dictionary priming the head of every segment copies from the tail of the one before, a dependency
 * chain through the whole stream.  So decoding and copying are separated (the multi-round resolution
 * of Sitaridi et al., "Massively-parallel lossless data decompression"):
 *   k_inflate     : per segment, lane 0 decodes (RFC 1951 3.2.3-3.2.7: stored, fixed, dynamic
 *                   blocks); every output byte gets a 32-bit word -- 0x80000000 | value when it is
 *                   known, else the distance back to the byte it copies.  Copies inside the segment
 *                   are followed on the spot, so an unknown byte always points BEFORE its segment.
 *   k_inf_resolve : pointer jumping, w[p] <- w[p - w[p]] (value, or the sum of the two distances);
 *                   every round at least doubles the number of segments a pointer spans.
 *   k_inf_finish  : words -> bytes, Adler-32 partial sums per segment.
 */
#define INF_WARPS 4
#define INF_LL_BITS 10
#define INF_D_BITS 8

struct InfTables {
    uint16_t ll[1 << INF_LL_BITS]; /* (symbol << 4) | length, 0 = longer than INF_LL_BITS */
    uint16_t ds[1 << INF_D_BITS];
    uint16_t sorted[288 + 32];     /* symbols ordered by (length, symbol): lit/len then distance */
    uint16_t count[2][16], first[2][16], offs[2][16]; /* canonical decode per code length */
    uint8_t lens[288 + 32];
    uint8_t cl[128];               /* code-length code: (symbol << 3) | length */
};

__constant__ uint16_t c_len_base[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
__constant__ uint8_t c_len_extra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
__constant__ uint16_t c_dist_base[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
__constant__ uint8_t c_dist_extra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};

/* LSB-first bit reader over the (4-byte aligned view of the) body; lane 0 only */
struct BitReader {
    const uint32_t *words;
    unsigned long long nwords, wi;
    unsigned long long buf;
    int cnt;
    __device__ __forceinline__ uint32_t ldw(unsigned long long i) const { return i < nwords ? __ldg(words + i) : 0u; }
    __device__ __forceinline__ void init(unsigned long long byte_pos)
    {
        wi = byte_pos >> 2;
        const int sh = (int)(byte_pos & 3) * 8;
        buf = (unsigned long long)(ldw(wi++) >> sh);
        cnt = 32 - sh;
        refill();
    }
    __device__ __forceinline__ void refill()
    {
        if (cnt <= 32) { buf |= (unsigned long long)ldw(wi++) << cnt; cnt += 32; }
    }
    __device__ __forceinline__ uint32_t peek(int nb) const { return (uint32_t)buf & ((1u << nb) - 1u); }
    __device__ __forceinline__ void skip(int nb) { buf >>= nb; cnt -= nb; }
    __device__ __forceinline__ uint32_t get(int nb) { const uint32_t v = peek(nb); skip(nb); return v; }
    __device__ __forceinline__ unsigned long long byte_pos() const { return (wi * 32 - (unsigned long long)cnt) >> 3; }
};

/* canonical decode, one bit at a time (codes longer than the table): returns symbol or -1 */
__device__ __forceinline__ int inf_slow(BitReader &br, const InfTables &T, int which)
{
    int code = 0, index_off = 0;
    for (int len = 1; len <= 15; len++) {
        code |= (int)((br.buf >> (len - 1)) & 1);
        const int cnt = T.count[which][len], first = T.first[which][len];
        if (code - first < cnt && code >= first) {
            br.skip(len);
            return T.sorted[(which ? 288 : 0) + T.offs[which][len] + (code - first)];
        }
        code <<= 1;
        (void)index_off;
    }
    return -1;
}

/* Build the decode structures of one alphabet from T.lens[base .. base + n) (executed by the whole warp).
 * Returns false if the lengths oversubscribe the code space. */
__device__ bool inf_build(InfTables &T, int which, int base, int n, int lane)
{
    uint16_t *tab = which ? T.ds : T.ll;
    const int bits = which ? INF_D_BITS : INF_LL_BITS;
    for (int i = lane; i < (1 << bits); i += 32) tab[i] = 0;
    if (lane < 16) T.count[which][lane] = 0;
    __syncwarp();
    if (lane == 0) {
        for (int s = 0; s < n; s++) T.count[which][T.lens[base + s]]++;
        T.count[which][0] = 0;
        int code = 0, off = 0, left = 1;
        bool ok = true;
        for (int len = 1; len <= 15; len++) {
            left <<= 1;
            left -= T.count[which][len];
            if (left < 0) ok = false;
            code = (code + T.count[which][len - 1]) << 1;
            T.first[which][len] = (uint16_t)code;
            T.offs[which][len] = (uint16_t)off;
            off += T.count[which][len];
        }
        T.first[which][0] = ok ? 1 : 0;
        /* symbols in (length, symbol) order; next[] reuses first[] + running counters kept in registers */
        int nxt[16];
#pragma unroll
        for (int len = 0; len < 16; len++) nxt[len] = T.offs[which][len];
        for (int s = 0; s < n; s++) {
            const int l = T.lens[base + s];
            if (l) {
                int slot = 0;
#pragma unroll
                for (int len = 1; len < 16; len++)
                    if (len == l) slot = nxt[len]++;
                T.sorted[(which ? 288 : 0) + slot] = (uint16_t)s;
            }
        }
    }
    __syncwarp();
    const bool ok = T.first[which][0] != 0;
    /* table: every code of <= bits bits, replicated over the unused high index bits */
    int total = 0;
    for (int len = 1; len <= 15; len++) total += T.count[which][len];
    for (int r = lane; r < total; r += 32) {
        const int s = T.sorted[(which ? 288 : 0) + r];
        const int l = T.lens[base + s];
        if (l <= bits) {
            const int code = T.first[which][l] + (r - T.offs[which][l]);
            const uint32_t rev = __brev((uint32_t)code) >> (32 - l);
            for (uint32_t i = rev; i < (1u << bits); i += 1u << l) tab[i] = (uint16_t)((s << 4) | l);
        }
    }
    __syncwarp();
    return ok;
}

#define INF_KNOWN 0x80000000u

__global__ void __launch_bounds__(INF_WARPS * 32)
k_inflate(const uint8_t *__restrict__ body, unsigned long long body_len, const unsigned long long *__restrict__ seg_off,
          long long seg0, int ns, int seg_bytes, long long n, long long hist_len, uint32_t *w, DevState *state)
{
    __shared__ InfTables s_tab[INF_WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    InfTables &T = s_tab[warp];
    const int sl = blockIdx.x * INF_WARPS + warp;
    if (sl >= ns) return;
    const long long s = seg0 + sl;
    const long long seg_lo = s * seg_bytes;
    const int seglen = (int)(((seg_lo + seg_bytes < n) ? seg_lo + seg_bytes : n) - seg_lo);
    uint32_t *out = w + (size_t)sl * seg_bytes;
    const unsigned long long a0 = (unsigned long long)((uintptr_t)body & 3);
    BitReader br;
    br.words = reinterpret_cast<const uint32_t *>(body - a0);
    br.nwords = (body_len + a0 + 3) >> 2;
    const unsigned long long seg_end = seg_off[s + 1] + a0;
    if (lane == 0) br.init(seg_off[s] + a0);
    int pos = 0;
    int err = 0;

    for (;;) {
        /* ---- block header ---- */
        int hdr = 0;
        if (lane == 0) {
            br.refill();
            hdr = (int)br.get(3);
        }
        hdr = __shfl_sync(FULL, hdr, 0);
        const int bfinal = hdr & 1, btype = hdr >> 1;
        if (btype == 0) {
            /* stored: LEN, NLEN at the next byte boundary, then LEN raw bytes */
            unsigned long long src = 0;
            int len = 0;
            if (lane == 0) {
                br.skip(br.cnt & 7);
                br.refill();
                const uint32_t l = br.get(16);
                br.refill();
                const uint32_t nl = br.get(16);
                if ((l ^ nl) != 0xFFFFu) err = 1;
                len = (int)l;
                src = br.byte_pos();
            }
            len = __shfl_sync(FULL, len, 0);
            src = __shfl_sync(FULL, src, 0);
            err = __shfl_sync(FULL, err, 0);
            if (err || pos + len > seglen || src + (unsigned long long)len > body_len + a0) { err = 1; break; }
            const uint8_t *sb = reinterpret_cast<const uint8_t *>(br.words) + src;
            for (int t = lane; t < len; t += 32) out[pos + t] = INF_KNOWN | __ldg(sb + t);
            pos += len;
            __syncwarp();
            if (lane == 0) br.init(src + (unsigned long long)len);
        } else if (btype == 1 || btype == 2) {
            int nll = 288, nd = 30;
            if (btype == 1) {
                for (int i = lane; i < 288; i += 32) T.lens[i] = (uint8_t)fixed_ll_len(i);
                T.lens[288 + lane] = 5;
                nd = 32;
            } else {
                /* dynamic: HLIT, HDIST, HCLEN, the code-length code, then the run-length coded lengths */
                if (lane == 0) {
                    br.refill();
                    nll = (int)br.get(5) + 257;
                    nd = (int)br.get(5) + 1;
                    const int ncl = (int)br.get(4) + 4;
                    uint8_t cl_len[19];
#pragma unroll
                    for (int i = 0; i < 19; i++) cl_len[i] = 0;
                    for (int i = 0; i < ncl; i++) {
                        br.refill();
                        const int v = (int)br.get(3);
#pragma unroll
                        for (int j = 0; j < 19; j++)
                            if (j == c_cl_order[i]) cl_len[j] = (uint8_t)v;
                    }
                    /* canonical code of the 19 symbols -> 7-bit table */
                    int cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0}, nxt[8];
#pragma unroll
                    for (int j = 0; j < 19; j++) {
#pragma unroll
                        for (int l = 1; l < 8; l++)
                            if (cl_len[j] == l) cnt[l]++;
                    }
                    int code = 0;
                    nxt[0] = 0;
#pragma unroll
                    for (int l = 1; l < 8; l++) { code = (code + (l > 1 ? cnt[l - 1] : 0)) << 1; nxt[l] = code; }
                    for (int i = 0; i < 128; i++) T.cl[i] = 0;
#pragma unroll
                    for (int j = 0; j < 19; j++) {
                        const int l = cl_len[j];
                        if (l) {
                            int c = 0;
#pragma unroll
                            for (int q = 1; q < 8; q++)
                                if (q == l) c = nxt[q]++;
                            const uint32_t rev = __brev((uint32_t)c) >> (32 - l);
                            for (uint32_t i = rev; i < 128u; i += 1u << l) T.cl[i] = (uint8_t)((j << 3) | l);
                        }
                    }
                    if (nll > 286 || nd > 30) err = 1;
                    int i = 0, prev = 0;
                    while (!err && i < nll + nd) {
                        br.refill();
                        const uint32_t e = T.cl[br.peek(7)];
                        if (!e) { err = 1; break; }
                        br.skip((int)(e & 7));
                        const int sym = (int)(e >> 3);
                        int rep = 1, val = sym;
                        if (sym == 16) { rep = 3 + (int)br.get(2); val = prev; if (i == 0) err = 1; }
                        else if (sym == 17) { rep = 3 + (int)br.get(3); val = 0; }
                        else if (sym == 18) { rep = 11 + (int)br.get(7); val = 0; }
                        if (i + rep > nll + nd) { err = 1; break; }
                        for (int r = 0; r < rep; r++, i++) T.lens[i < nll ? i : 288 + (i - nll)] = (uint8_t)val;
                        prev = val;
                    }
                    for (int k = nll; k < 288; k++) T.lens[k] = 0;
                    for (int k = nd; k < 32; k++) T.lens[288 + k] = 0;
                }
                err = __shfl_sync(FULL, err, 0);
                if (err) break;
                nll = 288; nd = 32; /* the unused tails were zeroed */
            }
            __syncwarp();
            const bool ok1 = inf_build(T, 0, 0, nll, lane);
            const bool ok2 = inf_build(T, 1, 288, nd, lane);
            if (!ok1 || !ok2) { err = 1; break; }
            /* ---- tokens ---- */
            for (;;) {
                int mlen = 0, mdist = 0, npos = pos;
                if (lane == 0) {
                    for (;;) {
                        br.refill();
                        int sym;
                        const uint32_t e = T.ll[br.peek(INF_LL_BITS)];
                        if (e) { br.skip((int)(e & 15)); sym = (int)(e >> 4); }
                        else sym = inf_slow(br, T, 0);
                        if (sym < 0) { mlen = -2; break; }
                        if (sym < 256) {
                            if (npos >= seglen) { mlen = -2; break; }
                            out[npos++] = INF_KNOWN | (uint32_t)sym;
                            continue;
                        }
                        if (sym == 256) { mlen = -1; break; }
                        if (sym > 285) { mlen = -2; break; }
                        mlen = c_len_base[sym - 257] + (int)br.get(c_len_extra[sym - 257]);
                        br.refill();
                        int ds;
                        const uint32_t e2 = T.ds[br.peek(INF_D_BITS)];
                        if (e2) { br.skip((int)(e2 & 15)); ds = (int)(e2 >> 4); }
                        else ds = inf_slow(br, T, 1);
                        if (ds < 0 || ds > 29) { mlen = -2; break; }
                        mdist = c_dist_base[ds] + (int)br.get(c_dist_extra[ds]);
                        if (seg_lo + npos - mdist < -hist_len) mlen = -2; /* reaches before the history the caller supplied */
                        break;
                    }
                }
                mlen = __shfl_sync(FULL, mlen, 0);
                mdist = __shfl_sync(FULL, mdist, 0);
                pos = __shfl_sync(FULL, npos, 0);
                if (mlen == -1) break;                 /* end of block */
                if (mlen < 0 || pos + mlen > seglen) { err = 1; break; }
                __syncwarp(); /* lane 0's literals are ordered before the copy's reads */
                /* copy: the source repeats with period mdist when the match overlaps itself.  A source inside
                 * the segment is already final-or-pointing-before-the-segment: follow it now. */
                for (int t = lane; t < mlen; t += 32) {
                    const int so = mdist >= mlen ? t : t % mdist;
                    const int sp = pos - mdist + so;
                    uint32_t v = (uint32_t)(pos + t - sp);
                    if (sp >= 0) {
                        const uint32_t a = __ldcg(out + sp);
                        v = (a & INF_KNOWN) ? a : v + a;
                    }
                    out[pos + t] = v;
                }
                pos += mlen;
                __syncwarp();
            }
            if (err) break;
        } else {
            err = 1;
            break;
        }
        unsigned long long bp = 0;
        if (lane == 0) bp = br.byte_pos();
        bp = __shfl_sync(FULL, bp, 0);
        if (bfinal || bp >= seg_end) break;
    }
    if (!err && pos != seglen) err = 1;
    if (err && lane == 0) atomicExch(&state->error, 2u);
}

/* One round of pointer jumping over the cnt words of a chunk (chunk position 0 = output byte chunk_lo).
 * Four words per thread and step; groups that are already all values cost one 16-byte read. */
__device__ __forceinline__ uint32_t inf_jump(const uint32_t *w, long long p, uint32_t v, long long chunk_lo, const uint8_t *d_out)
{
    const long long q = p - (long long)v;
    if (q < 0) return INF_KNOWN | d_out[chunk_lo + q]; /* an earlier chunk (or the caller's history): final bytes */
    const uint32_t a = __ldcg(w + q);
    return (a & INF_KNOWN) ? a : v + a;
}

__global__ void __launch_bounds__(256)
k_inf_resolve(uint32_t *w, long long cnt, long long chunk_lo, const uint8_t *__restrict__ d_out, unsigned int *pending)
{
    unsigned int left = 0;
    const long long quads = cnt >> 2;
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < quads; g += (long long)gridDim.x * blockDim.x) {
        uint4 v = __ldcg(reinterpret_cast<const uint4 *>(w) + g);
        if (v.x & v.y & v.z & v.w & INF_KNOWN) continue;
        const long long p = g << 2;
        if (!(v.x & INF_KNOWN)) v.x = inf_jump(w, p, v.x, chunk_lo, d_out);
        if (!(v.y & INF_KNOWN)) v.y = inf_jump(w, p + 1, v.y, chunk_lo, d_out);
        if (!(v.z & INF_KNOWN)) v.z = inf_jump(w, p + 2, v.z, chunk_lo, d_out);
        if (!(v.w & INF_KNOWN)) v.w = inf_jump(w, p + 3, v.w, chunk_lo, d_out);
        reinterpret_cast<uint4 *>(w)[g] = v;
        left |= ~(v.x & v.y & v.z & v.w) & INF_KNOWN;
    }
    if (blockIdx.x == 0 && threadIdx.x < (cnt & 3)) {
        const long long p = (quads << 2) + threadIdx.x;
        const uint32_t v = __ldcg(w + p);
        if (!(v & INF_KNOWN)) {
            const uint32_t nv = inf_jump(w, p, v, chunk_lo, d_out);
            w[p] = nv;
            left |= ~nv & INF_KNOWN;
        }
    }
    if (__any_sync(__activemask(), left != 0) && (threadIdx.x & 31) == 0) atomicOr(pending, 1u);
}

/* words -> bytes; Adler-32 partial sums of the segment (same bookkeeping as k_match) */
__global__ void __launch_bounds__(256)
k_inf_finish(const uint32_t *__restrict__ w, long long seg0, int seg_bytes, long long n, uint8_t *d_out, DevState *state)
{
    __shared__ uint32_t s_a[8];
    __shared__ unsigned long long s_b[8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long s = seg0 + blockIdx.x;
    const long long seg_lo = s * seg_bytes;
    const int seglen = (int)(((seg_lo + seg_bytes < n) ? seg_lo + seg_bytes : n) - seg_lo);
    const uint32_t *ws = w + (size_t)blockIdx.x * seg_bytes;
    uint32_t a = 0, bad = 0;
    unsigned long long b = 0;
    for (int i = tid; i < seglen; i += 256) {
        const uint32_t v = ws[i];
        bad |= !(v & INF_KNOWN);
        const uint32_t byte = v & 0xFF;
        d_out[seg_lo + i] = (uint8_t)byte;
        a += byte;
        b += (unsigned long long)byte * (uint32_t)(seglen - i);
    }
    for (int o = 16; o; o >>= 1) {
        a += __shfl_xor_sync(FULL, a, o);
        b += __shfl_xor_sync(FULL, b, o);
    }
    if (lane == 0) { s_a[warp] = a; s_b[warp] = b; }
    if (__syncthreads_or((int)bad) && tid == 0) atomicExch(&state->error, 3u);
    if (tid == 0) {
        for (int k = 1; k < 8; k++) { a += s_a[k]; b += s_b[k]; }
        const unsigned long long am = a % ADLER_MOD;
        const unsigned long long rest = (unsigned long long)(n - (seg_lo + seglen)) % ADLER_MOD;
        atomicAdd(&state->adl_a, am);
        atomicAdd(&state->adl_b, (b % ADLER_MOD + am * rest) % ADLER_MOD);
    }
}

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

/* ------------------------------------------------------------------------- */
/* host side: context + C-ABI                                                */
/* ------------------------------------------------------------------------- */

struct dfl_ctx {
    int device;
    dfl_params prm;
    /* chunk workspace */
    size_t ws_segs, ws_seg_bytes;
    uint32_t *tokcand; /* k_match: candidate token at every evaluated position */
    uint32_t *tokens;  /* k_parse: the segment's tokens, compacted */
    uint16_t *scratch; /* k_match radix-sort ping buffer, one slot per SM */
    SegArrays A;
    void *seg_blob;
    DevState *state;
    DevState *h_state; /* pinned */
    /* host-buffer path */
    uint8_t *d_in, *d_out;
    size_t d_in_cap, d_out_cap;
    cudaStream_t s_in, s_run, s_out; /* host path: copy-in, kernels, copy-out */
    cudaEvent_t *e_in, *e_done;
    unsigned long long *h_totals;     /* pinned: end offset of every piece */
    int pipe_cap;
    int parse_mode; /* 0 = by segment count, 1 = k_parse_warp, 2 = k_parse */
    /* segment index of the last compress call (byte offset of every segment in the body, + the end) and inflate state */
    unsigned long long *seg_index;
    size_t seg_index_cap, seg_index_n;
    uint8_t *d_png; /* PNG encode: the filtered scanlines */
    size_t d_png_cap;
    uint32_t *inf_w; /* inflate: one word per output byte of a chunk */
    size_t inf_cap;
    int inf_rounds;
    /* bookkeeping */
    long launches;
    long long last_n;
    int pending;
    /* token keeping */
    int keep;
    uint32_t *h_tokens;
    size_t h_tokens_cap;
    long nseg_kept;
    uint32_t *h_ntok, *h_btype, *h_out_bytes;
    unsigned long long *h_block_bits, *h_out_off;
    size_t h_seg_cap;
    /* timing */
    int timing;
    cudaEvent_t *ev;
    int *ev_class;
    int ev_n, ev_cap;
    double k_ms[K_NUM];
    char errbuf[256];
};

#define CK(call)                                                                          \
    do {                                                                                  \
        cudaError_t e_ = (call);                                                          \
        if (e_ != cudaSuccess) {                                                          \
            if (ctx) snprintf(ctx->errbuf, sizeof(ctx->errbuf), "%s: %s", #call, cudaGetErrorString(e_)); \
            return DFL_E_CUDA;                                                            \
        }                                                                                 \
    } while (0)

extern "C" void dfl_b200_default_params(dfl_params *p)
{
    p->window = DFL_MAX_WINDOW;
    p->seg_bytes = DFL_MAX_SEG;
    p->chain_limit = 0;
    p->use_dict = 1;
    p->block_mode = 0;
}

extern "C" const char *dfl_b200_strerror(int err)
{
    switch (err) {
    case DFL_OK: return "ok";
    case DFL_E_MALLOC: return "E_MALLOC: allocation failed";
    case DFL_E_SZ: return "E_SZ: output buffer too small";
    case DFL_E_RANGE: return "E_RANGE: parameter out of range";
    case DFL_E_INVAL: return "E_INVAL: invalid argument";
    case DFL_E_CUDA: return "CUDA failure (no sm_100 device, driver error or kernel fault)";
    case DFL_E_DATA: return "E_ZINV: corrupt stream, wrong segment index or Adler-32 mismatch";
    default: return "unknown error";
    }
}

static int check_params(const dfl_params *p)
{
    if (p->window < 1 || p->window > DFL_MAX_WINDOW) return DFL_E_RANGE;
    if (p->seg_bytes < 1 || p->seg_bytes > DFL_MAX_SEG) return DFL_E_RANGE;
    if (p->chain_limit < 0) return DFL_E_RANGE;
    if (p->block_mode < 0 || p->block_mode > 3) return DFL_E_RANGE;
    return DFL_OK;
}

extern "C" size_t dfl_b200_bound(const dfl_params *prm, size_t n)
{
    size_t nseg = n ? (n + prm->seg_bytes - 1) / prm->seg_bytes : 0;
    size_t b = n + nseg * (5 * ((prm->seg_bytes + 65534) / 65535) + 5 + 8) + 64;
    /* a forced fixed/dynamic block may expand: <= 15 bits per literal + header */
    if (prm->block_mode >= 2) b = 2 * n + nseg * 640 + 64;
    return b;
}

static void free_ws(dfl_ctx *ctx)
{
    if (ctx->tokcand) cudaFree(ctx->tokcand);
    if (ctx->tokens) cudaFree(ctx->tokens);
    ctx->tokens = NULL;
    if (ctx->seg_blob) cudaFree(ctx->seg_blob);
    ctx->tokcand = NULL;
    ctx->seg_blob = NULL;
    ctx->ws_segs = 0;
}

static int ensure_index(dfl_ctx *ctx, size_t nseg)
{
    if (ctx->seg_index_cap < nseg + 1) {
        if (ctx->seg_index) cudaFree(ctx->seg_index);
        ctx->seg_index = NULL;
        ctx->seg_index_cap = 0;
        CK(cudaMalloc(&ctx->seg_index, (nseg + 1) * sizeof(unsigned long long)));
        ctx->seg_index_cap = nseg + 1;
    }
    ctx->seg_index_n = nseg;
    return DFL_OK;
}

static int ensure_ws(dfl_ctx *ctx, size_t nseg_chunk)
{
    if (ctx->ws_segs >= nseg_chunk && ctx->ws_seg_bytes == (size_t)ctx->prm.seg_bytes) return DFL_OK;
    free_ws(ctx);
    size_t ns = nseg_chunk;
    CK(cudaMalloc(&ctx->tokcand, ns * ctx->prm.seg_bytes * sizeof(uint32_t)));
    CK(cudaMalloc(&ctx->tokens, ns * ctx->prm.seg_bytes * sizeof(uint32_t)));
    size_t per = (LLS + DS + LLS + DS + HDR_WORDS + 6) * sizeof(uint32_t) + 2 * sizeof(unsigned long long);
    CK(cudaMalloc(&ctx->seg_blob, ns * per + 256));
    unsigned char *p = (unsigned char *)ctx->seg_blob;
    SegArrays &A = ctx->A;
    A.block_bits = (unsigned long long *)p; p += ns * 8;
    A.out_off = (unsigned long long *)p; p += ns * 8;
    A.ll_freq = (uint32_t *)p; p += ns * LLS * 4;
    A.d_freq = (uint32_t *)p; p += ns * DS * 4;
    A.lltab = (uint32_t *)p; p += ns * LLS * 4;
    A.dtab = (uint32_t *)p; p += ns * DS * 4;
    A.hdr = (uint32_t *)p; p += ns * HDR_WORDS * 4;
    A.ntok = (uint32_t *)p; p += ns * 4;
    A.extra = (uint32_t *)p; p += ns * 4;
    A.hdr_bits = (uint32_t *)p; p += ns * 4;
    A.btype = (uint32_t *)p; p += ns * 4;
    A.out_bytes = (uint32_t *)p; p += ns * 4;
    ctx->ws_segs = ns;
    ctx->ws_seg_bytes = ctx->prm.seg_bytes;
    return DFL_OK;
}

extern "C" int dfl_b200_create(dfl_ctx **out, int device, const dfl_params *prm)
{
    dfl_ctx *ctx = NULL;
    int ndev = 0;
    if (!out) return DFL_E_INVAL;
    *out = NULL;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return DFL_E_CUDA; /* no CPU fallback */
    if (device < 0 || device >= ndev) return DFL_E_RANGE;
    ctx = (dfl_ctx *)calloc(1, sizeof(dfl_ctx));
    if (!ctx) return DFL_E_MALLOC;
    ctx->device = device;
    if (prm) ctx->prm = *prm; else dfl_b200_default_params(&ctx->prm);
    int rc = check_params(&ctx->prm);
    if (rc) { free(ctx); return rc; }
    cudaDeviceProp prop;
    if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&prop, device) != cudaSuccess ||
        prop.major < 10) {
        free(ctx);
        return DFL_E_CUDA;
    }
    if (cudaFuncSetAttribute(k_match<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(MatchSmem)) != cudaSuccess ||
        cudaFuncSetAttribute(k_match<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(MatchSmem)) != cudaSuccess ||
        cudaFuncSetAttribute(k_parse, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ParseSmem)) != cudaSuccess ||
        cudaMalloc(&ctx->state, sizeof(DevState)) != cudaSuccess ||
        cudaMalloc(&ctx->scratch, (size_t)SCRATCH_SLOTS * NPOS_MAX * sizeof(uint16_t)) != cudaSuccess ||
        cudaMallocHost(&ctx->h_state, sizeof(DevState)) != cudaSuccess) {
        free(ctx);
        return DFL_E_CUDA;
    }
    *out = ctx;
    return DFL_OK;
}

extern "C" void dfl_b200_destroy(dfl_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    free_ws(ctx);
    if (ctx->state) cudaFree(ctx->state);
    if (ctx->scratch) cudaFree(ctx->scratch);
    if (ctx->h_state) cudaFreeHost(ctx->h_state);
    if (ctx->d_in) cudaFree(ctx->d_in);
    if (ctx->d_out) cudaFree(ctx->d_out);
    if (ctx->h_tokens) cudaFreeHost(ctx->h_tokens);
    if (ctx->seg_index) cudaFree(ctx->seg_index);
    if (ctx->inf_w) cudaFree(ctx->inf_w);
    if (ctx->d_png) cudaFree(ctx->d_png);
    for (int i = 0; i < ctx->pipe_cap; i++) { cudaEventDestroy(ctx->e_in[i]); cudaEventDestroy(ctx->e_done[i]); }
    free(ctx->e_in); free(ctx->e_done);
    if (ctx->h_totals) cudaFreeHost(ctx->h_totals);
    if (ctx->s_in) { cudaStreamDestroy(ctx->s_in); cudaStreamDestroy(ctx->s_run); cudaStreamDestroy(ctx->s_out); }
    free(ctx->h_ntok); free(ctx->h_btype); free(ctx->h_out_bytes); free(ctx->h_block_bits); free(ctx->h_out_off);
    for (int i = 0; i < ctx->ev_cap; i++) cudaEventDestroy(ctx->ev[i]);
    free(ctx->ev); free(ctx->ev_class);
    free(ctx);
}

extern "C" int dfl_b200_set_params(dfl_ctx *ctx, const dfl_params *prm)
{
    if (!ctx || !prm) return DFL_E_INVAL;
    int rc = check_params(prm);
    if (rc) return rc;
    ctx->prm = *prm;
    return DFL_OK;
}

extern "C" int dfl_b200_set_parse_mode(dfl_ctx *ctx, int mode)
{
    if (!ctx || mode < 0 || mode > 2) return DFL_E_INVAL;
    ctx->parse_mode = mode;
    return DFL_OK;
}

extern "C" int dfl_b200_keep_tokens(dfl_ctx *ctx, int enable)
{
    if (!ctx) return DFL_E_INVAL;
    ctx->keep = enable != 0;
    return DFL_OK;
}

extern "C" int dfl_b200_enable_timing(dfl_ctx *ctx, int enable)
{
    if (!ctx) return DFL_E_INVAL;
    ctx->timing = enable != 0;
    return DFL_OK;
}

extern "C" int dfl_b200_num_kernels(void) { return K_NUM; }
extern "C" const char *dfl_b200_kernel_name(int k) { return (k >= 0 && k < K_NUM) ? k_names[k] : ""; }
extern "C" double dfl_b200_kernel_ms(dfl_ctx *ctx, int k) { return (ctx && k >= 0 && k < K_NUM) ? ctx->k_ms[k] : 0.0; }
extern "C" long dfl_b200_kernel_launches(dfl_ctx *ctx) { return ctx ? ctx->launches : 0; }

static int ev_mark(dfl_ctx *ctx, int cls, cudaStream_t st)
{
    /* records one event; events come in (start, stop) pairs per launch */
    if (!ctx->timing) return DFL_OK;
    if (ctx->ev_n == ctx->ev_cap) {
        int ncap = ctx->ev_cap ? ctx->ev_cap * 2 : 256;
        cudaEvent_t *ne = (cudaEvent_t *)realloc(ctx->ev, ncap * sizeof(cudaEvent_t));
        int *nc = (int *)realloc(ctx->ev_class, ncap * sizeof(int));
        if (!ne || !nc) return DFL_E_MALLOC;
        ctx->ev = ne; ctx->ev_class = nc;
        for (int i = ctx->ev_cap; i < ncap; i++) CK(cudaEventCreate(&ctx->ev[i]));
        ctx->ev_cap = ncap;
    }
    ctx->ev_class[ctx->ev_n] = cls;
    CK(cudaEventRecord(ctx->ev[ctx->ev_n], st));
    ctx->ev_n++;
    return DFL_OK;
}

#define LAUNCH(cls, kern, grid, block, smem, st, ...)                 \
    do {                                                              \
        int r_ = ev_mark(ctx, cls, st);                               \
        if (r_) return r_;                                            \
        kern<<<grid, block, smem, st>>>(__VA_ARGS__);                 \
        ctx->launches++;                                              \
        r_ = ev_mark(ctx, cls, st);                                   \
        if (r_) return r_;                                            \
    } while (0)

static int keep_reserve(dfl_ctx *ctx, size_t n, size_t nseg)
{
    if (ctx->h_tokens_cap < n) {
        if (ctx->h_tokens) cudaFreeHost(ctx->h_tokens);
        ctx->h_tokens = NULL;
        ctx->h_tokens_cap = 0;
        CK(cudaMallocHost(&ctx->h_tokens, (n ? n : 1) * sizeof(uint32_t)));
        ctx->h_tokens_cap = n;
    }
    if (ctx->h_seg_cap < nseg) {
        free(ctx->h_ntok); free(ctx->h_btype); free(ctx->h_out_bytes); free(ctx->h_block_bits); free(ctx->h_out_off);
        ctx->h_ntok = (uint32_t *)malloc(nseg * 4);
        ctx->h_btype = (uint32_t *)malloc(nseg * 4);
        ctx->h_out_bytes = (uint32_t *)malloc(nseg * 4);
        ctx->h_block_bits = (unsigned long long *)malloc(nseg * 8);
        ctx->h_out_off = (unsigned long long *)malloc(nseg * 8);
        if (!ctx->h_ntok || !ctx->h_btype || !ctx->h_out_bytes || !ctx->h_block_bits || !ctx->h_out_off) return DFL_E_MALLOC;
        ctx->h_seg_cap = nseg;
    }
    return DFL_OK;
}

/* The five kernels of one chunk (segments [s0, s0 + ns) of a stream of nseg segments), on stream st. */
static int launch_chunk(dfl_ctx *ctx, const uint8_t *d_in, size_t n, size_t hist_len, int is_final, size_t s0, int ns,
                        size_t nseg, uint8_t *d_out, size_t out_cap, cudaStream_t st)
{
    const dfl_params prm = ctx->prm;
    if (prm.chain_limit)
        LAUNCH(K_MATCH, k_match<true>, ns, MATCH_THREADS, sizeof(MatchSmem), st, d_in, (long long)n, (long long)hist_len,
               (long long)s0, prm, ctx->tokcand, ctx->scratch, ctx->state);
    else
        LAUNCH(K_MATCH, k_match<false>, ns, MATCH_THREADS, sizeof(MatchSmem), st, d_in, (long long)n, (long long)hist_len,
               (long long)s0, prm, ctx->tokcand, ctx->scratch, ctx->state);
    /* few segments: a CTA per segment (short dependent chain); enough segments to fill every warp slot
     * of the chip: a warp per segment (fewer instructions per token) */
    const int per_warp = ctx->parse_mode ? ctx->parse_mode == 1 : ns >= PARSE_WARP_MIN_SEGS;
    if (per_warp)
        LAUNCH(K_PARSE, k_parse_warp, (ns + PARSE_WARPS - 1) / PARSE_WARPS, PARSE_WARPS * 32, 0, st, (long long)n,
               (long long)s0, ns, prm, ctx->tokcand, ctx->tokens, ctx->A);
    else
        LAUNCH(K_PARSE, k_parse, ns, PARSE_THREADS, sizeof(ParseSmem), st, (long long)n, (long long)s0, prm, ctx->tokcand,
               ctx->tokens, ctx->A);
    LAUNCH(K_HUFF, k_huff, (ns + HUFF_WARPS - 1) / HUFF_WARPS, HUFF_WARPS * 32, 0, st, (long long)n,
           (long long)s0, ns, prm, ctx->A);
    LAUNCH(K_SCAN, k_scan, 1, SCAN_THREADS, 0, st, ns, ctx->A, ctx->state);
    /* segment index: where each segment starts in the body, and (entry nseg) where the body ends so far */
    CK(cudaMemcpyAsync(ctx->seg_index + s0, ctx->A.out_off, (size_t)ns * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, st));
    CK(cudaMemcpyAsync(ctx->seg_index + s0 + ns, &ctx->state->total_out, sizeof(unsigned long long), cudaMemcpyDeviceToDevice, st));
    LAUNCH(K_EMIT, k_emit, ns, EMIT_THREADS, 0, st, d_in, (long long)n, (long long)s0, (long long)(nseg - 1),
           is_final, prm, ctx->tokens, ctx->A, d_out, (unsigned long long)out_cap, ctx->state);
    return DFL_OK;
}

extern "C" int dfl_b200_deflate_enqueue(dfl_ctx *ctx, const uint8_t *d_in, size_t n, size_t hist_len,
                                        int is_final, uint8_t *d_out, size_t out_cap, void *stream)
{
    if (!ctx || (!d_in && n) || !d_out) return DFL_E_INVAL;
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaSetDevice(ctx->device));
    const dfl_params prm = ctx->prm;
    const size_t nseg = n ? (n + prm.seg_bytes - 1) / prm.seg_bytes : 0;
    size_t chunk_segs = DFL_CHUNK_BYTES / (size_t)prm.seg_bytes;
    if (chunk_segs > nseg) chunk_segs = nseg;
    if (chunk_segs < 1) chunk_segs = 1;
    int rc = ensure_ws(ctx, chunk_segs);
    if (rc) return rc;
    rc = ensure_index(ctx, nseg);
    if (rc) return rc;
    ctx->ev_n = 0;
    ctx->last_n = (long long)n;
    ctx->nseg_kept = 0;
    if (ctx->keep) {
        rc = keep_reserve(ctx, n, nseg);
        if (rc) return rc;
    }
    CK(cudaMemsetAsync(ctx->state, 0, sizeof(DevState), st));
    if (n == 0) {
        /* empty input: a lone empty stored block */
        static const uint8_t empty_final[5] = {1, 0, 0, 0xFF, 0xFF}, empty_more[5] = {0, 0, 0, 0xFF, 0xFF};
        if (out_cap < 5) return DFL_E_SZ;
        CK(cudaMemcpyAsync(d_out, is_final ? empty_final : empty_more, 5, cudaMemcpyHostToDevice, st));
        unsigned long long five = 5;
        CK(cudaMemcpyAsync(&ctx->state->total_out, &five, 8, cudaMemcpyHostToDevice, st));
    }
    for (size_t s0 = 0; s0 < nseg; s0 += chunk_segs) {
        int ns = (int)(nseg - s0 < chunk_segs ? nseg - s0 : chunk_segs);
        rc = launch_chunk(ctx, d_in, n, hist_len, is_final, s0, ns, nseg, d_out, out_cap, st);
        if (rc) return rc;
        if (ctx->keep) {
            /* debug / fd_stats path: pull this chunk's tokens and segment records */
            size_t pos0 = s0 * (size_t)prm.seg_bytes;
            size_t cnt = (size_t)ns * prm.seg_bytes;
            if (pos0 + cnt > n) cnt = n - pos0;
            CK(cudaMemcpyAsync(ctx->h_tokens + pos0, ctx->tokens, cnt * 4, cudaMemcpyDeviceToHost, st));
            CK(cudaMemcpyAsync(ctx->h_ntok + s0, ctx->A.ntok, ns * 4, cudaMemcpyDeviceToHost, st));
            CK(cudaMemcpyAsync(ctx->h_btype + s0, ctx->A.btype, ns * 4, cudaMemcpyDeviceToHost, st));
            CK(cudaMemcpyAsync(ctx->h_out_bytes + s0, ctx->A.out_bytes, ns * 4, cudaMemcpyDeviceToHost, st));
            CK(cudaMemcpyAsync(ctx->h_block_bits + s0, ctx->A.block_bits, ns * 8, cudaMemcpyDeviceToHost, st));
            CK(cudaMemcpyAsync(ctx->h_out_off + s0, ctx->A.out_off, ns * 8, cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            ctx->nseg_kept = (long)(s0 + ns);
        }
    }
    LAUNCH(K_FINAL, k_final, 1, 32, 0, st, (long long)n, ctx->state);
    CK(cudaGetLastError());
    ctx->pending = 1;
    return DFL_OK;
}

extern "C" int dfl_b200_deflate_finish(dfl_ctx *ctx, size_t *out_len, uint32_t *adler, void *stream)
{
    if (!ctx || !ctx->pending) return DFL_E_INVAL;
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaMemcpyAsync(ctx->h_state, ctx->state, sizeof(DevState), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    ctx->pending = 0;
    if (ctx->timing) {
        for (int k = 0; k < K_NUM; k++) ctx->k_ms[k] = 0;
        for (int i = 0; i + 1 < ctx->ev_n; i += 2) {
            float ms = 0;
            CK(cudaEventElapsedTime(&ms, ctx->ev[i], ctx->ev[i + 1]));
            ctx->k_ms[ctx->ev_class[i]] += ms;
        }
    }
    if (ctx->h_state->error) return DFL_E_SZ;
    if (out_len) *out_len = (size_t)ctx->h_state->total_out;
    if (adler) *adler = ctx->h_state->adler;
    return DFL_OK;
}

extern "C" int dfl_b200_deflate_device(dfl_ctx *ctx, const uint8_t *d_in, size_t n, size_t hist_len,
                                       int is_final, uint8_t *d_out, size_t out_cap, size_t *out_len,
                                       uint32_t *adler, void *stream)
{
    int rc = dfl_b200_deflate_enqueue(ctx, d_in, n, hist_len, is_final, d_out, out_cap, stream);
    if (rc) return rc;
    return dfl_b200_deflate_finish(ctx, out_len, adler, stream);
}

#ifndef HOST_PIPE_BYTES
#define HOST_PIPE_BYTES (64u << 20) /* piece size of the host path: H2D, compute and D2H of neighbouring pieces overlap */
#endif

static int host_pipe_init(dfl_ctx *ctx, int npieces)
{
    if (!ctx->s_in) {
        CK(cudaStreamCreateWithFlags(&ctx->s_in, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&ctx->s_run, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&ctx->s_out, cudaStreamNonBlocking));
    }
    if (ctx->pipe_cap < npieces) {
        for (int i = 0; i < ctx->pipe_cap; i++) { cudaEventDestroy(ctx->e_in[i]); cudaEventDestroy(ctx->e_done[i]); }
        free(ctx->e_in); free(ctx->e_done);
        if (ctx->h_totals) cudaFreeHost(ctx->h_totals);
        ctx->pipe_cap = 0;
        ctx->e_in = (cudaEvent_t *)calloc(npieces, sizeof(cudaEvent_t));
        ctx->e_done = (cudaEvent_t *)calloc(npieces, sizeof(cudaEvent_t));
        if (!ctx->e_in || !ctx->e_done) return DFL_E_MALLOC;
        CK(cudaMallocHost(&ctx->h_totals, (size_t)npieces * sizeof(unsigned long long)));
        for (int i = 0; i < npieces; i++) {
            CK(cudaEventCreateWithFlags(&ctx->e_in[i], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&ctx->e_done[i], cudaEventDisableTiming));
        }
        ctx->pipe_cap = npieces;
    }
    return DFL_OK;
}

extern "C" int dfl_b200_compress_host(dfl_ctx *ctx, const uint8_t *in, size_t n, uint8_t *out, size_t cap,
                                      size_t *out_len)
{
    if (!ctx || (!in && n) || !out || !out_len) return DFL_E_INVAL;
    CK(cudaSetDevice(ctx->device));
    const dfl_params prm = ctx->prm;
    size_t bound = dfl_b200_bound(&prm, n);
    if (ctx->d_in_cap < n + 16) {
        if (ctx->d_in) cudaFree(ctx->d_in);
        ctx->d_in = NULL; ctx->d_in_cap = 0;
        CK(cudaMalloc(&ctx->d_in, n + 16));
        ctx->d_in_cap = n + 16;
    }
    if (ctx->d_out_cap < bound + 16) {
        if (ctx->d_out) cudaFree(ctx->d_out);
        ctx->d_out = NULL; ctx->d_out_cap = 0;
        CK(cudaMalloc(&ctx->d_out, bound + 16));
        ctx->d_out_cap = bound + 16;
    }
    uint8_t *d_body = ctx->d_out + 2; /* body at +2 so the zlib header can be prepended without a second copy */
    size_t body = 0;
    uint32_t adler = 1;
    const size_t nseg = n ? (n + prm.seg_bytes - 1) / prm.seg_bytes : 0;
    size_t piece_segs = HOST_PIPE_BYTES / (size_t)prm.seg_bytes;
    if (piece_segs < 1) piece_segs = 1;
    /* piece boundaries (in segments): a quarter piece first so that the kernels start early, then full pieces
     * (measured on 256 MiB: 64 MiB pieces 16.0 GB/s, 32 MiB 15.6, 16 MiB 13.4; an extra small last piece 15.8) */
    const size_t first_segs = piece_segs / 4 ? piece_segs / 4 : 1;
    std::vector<size_t> bnd(1, 0);
    if (nseg) {
        size_t b = first_segs < nseg ? first_segs : nseg;
        bnd.push_back(b);
        while (nseg - b > piece_segs) { b += piece_segs; bnd.push_back(b); }
        if (b < nseg) bnd.push_back(nseg);
    }
    const int npieces = (int)bnd.size() - 1;
#define PIECE_S0(i) (bnd[(i)])
#define PIECE_S1(i) (bnd[(i) + 1])
    if (ctx->keep || npieces < 2) {
        /* small input, or the token-keeping debug path: one copy in, one call, one copy out */
        cudaStream_t st = 0;
        if (n) CK(cudaMemcpyAsync(ctx->d_in, in, n, cudaMemcpyHostToDevice, st));
        int rc = dfl_b200_deflate_device(ctx, ctx->d_in, n, 0, 1, d_body, bound, &body, &adler, st);
        if (rc) return rc;
        if (cap < body + 6) return DFL_E_SZ;
        CK(cudaMemcpyAsync(out + 2, d_body, body, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
    } else {
        /* pipeline over pieces: copy-in of piece i+1, kernels of piece i and copy-out of piece i-1 overlap
         * on three streams; output offsets stay on the device, each piece's end offset comes back with it */
        int rc = ensure_ws(ctx, piece_segs);
        if (rc) return rc;
        rc = ensure_index(ctx, nseg);
        if (rc) return rc;
        ctx->last_n = (long long)n;
        rc = host_pipe_init(ctx, npieces);
        if (rc) return rc;
        ctx->ev_n = 0;
        ctx->nseg_kept = 0;
        CK(cudaMemsetAsync(ctx->state, 0, sizeof(DevState), ctx->s_run));
        for (int i = 0; i < npieces; i++) {
            const size_t lo = PIECE_S0(i) * prm.seg_bytes;
            const size_t hi = (PIECE_S1(i) * prm.seg_bytes < n) ? PIECE_S1(i) * prm.seg_bytes : n;
            CK(cudaMemcpyAsync(ctx->d_in + lo, in + lo, hi - lo, cudaMemcpyHostToDevice, ctx->s_in));
            CK(cudaEventRecord(ctx->e_in[i], ctx->s_in));
        }
        size_t copied = 0;
        for (int i = 0; i <= npieces; i++) {
            if (i < npieces) {
                const size_t s0 = PIECE_S0(i);
                const int ns = (int)(PIECE_S1(i) - s0);
                CK(cudaStreamWaitEvent(ctx->s_run, ctx->e_in[i], 0));
                rc = launch_chunk(ctx, ctx->d_in, n, 0, 1, s0, ns, nseg, d_body, bound, ctx->s_run);
                if (rc) return rc;
                if (i == npieces - 1) {
                    LAUNCH(K_FINAL, k_final, 1, 32, 0, ctx->s_run, (long long)n, ctx->state);
                    CK(cudaMemcpyAsync(ctx->h_state, ctx->state, sizeof(DevState), cudaMemcpyDeviceToHost, ctx->s_run));
                }
                CK(cudaMemcpyAsync(&ctx->h_totals[i], &ctx->state->total_out, sizeof(unsigned long long),
                                   cudaMemcpyDeviceToHost, ctx->s_run));
                CK(cudaEventRecord(ctx->e_done[i], ctx->s_run));
            }
            if (i >= 1) {
                /* piece i-1 is complete: its bytes [copied, h_totals[i-1]) can leave while piece i runs */
                CK(cudaEventSynchronize(ctx->e_done[i - 1]));
                const size_t end = (size_t)ctx->h_totals[i - 1];
                if (cap < end + 6) { /* let everything in flight land before the caller's buffers go away */
                    cudaStreamSynchronize(ctx->s_run);
                    cudaStreamSynchronize(ctx->s_out);
                    return DFL_E_SZ;
                }
                CK(cudaMemcpyAsync(out + 2 + copied, d_body + copied, end - copied, cudaMemcpyDeviceToHost, ctx->s_out));
                copied = end;
            }
        }
        CK(cudaStreamSynchronize(ctx->s_out));
        CK(cudaGetLastError());
        if (ctx->h_state->error) return DFL_E_SZ;
        body = (size_t)ctx->h_state->total_out;
        adler = ctx->h_state->adler;
    }
#undef PIECE_S0
#undef PIECE_S1
    out[0] = 0x78; /* CM = 8, CINFO = 7: header the decoder checks at src/deflate_decompress.c:352-360 */
    out[1] = 0x9C;
    out[2 + body + 0] = (uint8_t)(adler >> 24); /* RFC 1950: Adler-32 MSB first */
    out[2 + body + 1] = (uint8_t)(adler >> 16);
    out[2 + body + 2] = (uint8_t)(adler >> 8);
    out[2 + body + 3] = (uint8_t)adler;
    *out_len = body + 6;
    return DFL_OK;
}

extern "C" long dfl_b200_num_segments(dfl_ctx *ctx) { return ctx ? ctx->nseg_kept : 0; }

/* segments of the last call whose index was rebuilt with the ballot-ranked scatter (expected 0) */
extern "C" unsigned int dfl_b200_sort_fallbacks(dfl_ctx *ctx) { return ctx ? ctx->h_state->sort_fallbacks : 0; }

/* k_match SM cycles of the last call, summed over CTAs: 0 stage-in, 1 build, 2 search, 3 total */
extern "C" unsigned long long dfl_b200_phase_cycles(dfl_ctx *ctx, int k)
{
    return (ctx && k >= 0 && k < 24) ? ctx->h_state->t_phase[k] : 0;
}

extern "C" int dfl_b200_seg_info(dfl_ctx *ctx, long seg, dfl_seg_info *info)
{
    if (!ctx || !info || seg < 0 || seg >= ctx->nseg_kept) return DFL_E_RANGE;
    info->btype = (int32_t)ctx->h_btype[seg];
    info->ntokens = (int32_t)ctx->h_ntok[seg];
    info->block_bits = (int64_t)ctx->h_block_bits[seg];
    info->out_bytes = (int64_t)ctx->h_out_bytes[seg];
    info->out_offset = (int64_t)ctx->h_out_off[seg];
    return DFL_OK;
}

extern "C" long dfl_b200_tokens(dfl_ctx *ctx, long seg, uint32_t *tokens, long cap)
{
    if (!ctx || seg < 0 || seg >= ctx->nseg_kept) return DFL_E_RANGE;
    long nt = (long)ctx->h_ntok[seg];
    if (tokens) {
        if (cap < nt) return DFL_E_SZ;
        memcpy(tokens, ctx->h_tokens + (size_t)seg * ctx->prm.seg_bytes, (size_t)nt * 4);
    }
    return nt;
}

extern "C" int dfl_b200_synth_random(uint8_t *d_out, size_t n, uint64_t seed, void *stream)
{
    dfl_ctx *ctx = NULL;
    k_synth_random<<<148 * 8, 256, 0, (cudaStream_t)stream>>>(d_out, (unsigned long long)n, (unsigned long long)seed);
    CK(cudaGetLastError());
    return DFL_OK;
}

/* ------------------------------------------------------------------------- */
/* segment index + inflate                                                   */
/* ------------------------------------------------------------------------- */

extern "C" long dfl_b200_seg_index(dfl_ctx *ctx, uint64_t *offsets, long cap)
{
    if (!ctx) return DFL_E_INVAL;
    const long cnt = (long)ctx->seg_index_n + 1;
    if (!ctx->seg_index) return DFL_E_INVAL;
    if (offsets) {
        if (cap < cnt) return DFL_E_SZ;
        if (cudaSetDevice(ctx->device) != cudaSuccess ||
            cudaMemcpy(offsets, ctx->seg_index, (size_t)cnt * sizeof(uint64_t), cudaMemcpyDeviceToHost) != cudaSuccess)
            return DFL_E_CUDA;
        if (ctx->seg_index_n == 0) offsets[0] = 0;
    }
    return cnt;
}

extern "C" const uint64_t *dfl_b200_seg_index_device(dfl_ctx *ctx) { return ctx ? (const uint64_t *)ctx->seg_index : NULL; }

extern "C" int dfl_b200_inflate_device(dfl_ctx *ctx, const uint8_t *d_body, size_t body_len, const uint64_t *d_seg_off,
                                       size_t nseg, uint32_t seg_bytes, size_t hist_len, uint8_t *d_out, size_t n,
                                       uint32_t *adler, void *stream)
{
    if (!ctx || (!d_body && body_len) || (!d_out && n) || (!d_seg_off && nseg)) return DFL_E_INVAL;
    if (seg_bytes < 1 || seg_bytes > DFL_MAX_SEG) return DFL_E_RANGE;
    if (nseg != (n ? (n + seg_bytes - 1) / seg_bytes : 0)) return DFL_E_INVAL;
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaSetDevice(ctx->device));
    size_t chunk_segs = DFL_CHUNK_BYTES / (size_t)seg_bytes;
    if (chunk_segs > nseg) chunk_segs = nseg;
    if (chunk_segs < 1) chunk_segs = 1;
    const size_t words = chunk_segs * seg_bytes;
    if (ctx->inf_cap < words) {
        if (ctx->inf_w) cudaFree(ctx->inf_w);
    if (ctx->d_png) cudaFree(ctx->d_png);
        ctx->inf_w = NULL;
        ctx->inf_cap = 0;
        CK(cudaMalloc(&ctx->inf_w, words * sizeof(uint32_t) + 16));
        ctx->inf_cap = words;
    }
    ctx->ev_n = 0;
    CK(cudaMemsetAsync(ctx->state, 0, sizeof(DevState), st));
    unsigned int *pending = &ctx->state->inf_pending;
    int rounds_total = 0;
    for (size_t s0 = 0; s0 < nseg; s0 += chunk_segs) {
        const int ns = (int)(nseg - s0 < chunk_segs ? nseg - s0 : chunk_segs);
        const long long chunk_lo = (long long)(s0 * seg_bytes);
        const long long cnt = (long long)((s0 + ns) * (size_t)seg_bytes < n ? (size_t)ns * seg_bytes : n - s0 * seg_bytes);
        LAUNCH(K_INFLATE, k_inflate, (ns + INF_WARPS - 1) / INF_WARPS, INF_WARPS * 32, 0, st, d_body, (unsigned long long)body_len,
               (const unsigned long long *)d_seg_off, (long long)s0, ns, (int)seg_bytes, (long long)n, (long long)hist_len,
               ctx->inf_w, ctx->state);
        /* pointer jumping until no word is a pointer any more: every round at least doubles the number of
         * segments a pointer spans, so ceil(log2(ns)) + 2 rounds always suffice */
        int max_rounds = 3;
        for (int v = ns; v > 1; v >>= 1) max_rounds++;
        for (int r = 0;; r++) {
            CK(cudaMemsetAsync(pending, 0, sizeof(unsigned int), st));
            LAUNCH(K_INF_RESOLVE, k_inf_resolve, 148 * 16, 256, 0, st, ctx->inf_w, cnt, chunk_lo, (const uint8_t *)d_out, pending);
            unsigned int left = 0;
            CK(cudaMemcpyAsync(&ctx->h_state->inf_pending, pending, sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            left = ctx->h_state->inf_pending;
            rounds_total++;
            if (!left) break;
            if (r >= max_rounds) return DFL_E_DATA; /* a pointer cycle: not a stream this library wrote */
        }
        LAUNCH(K_INF_FINISH, k_inf_finish, ns, 256, 0, st, (const uint32_t *)ctx->inf_w, (long long)s0, (int)seg_bytes, (long long)n,
               d_out, ctx->state);
    }
    LAUNCH(K_FINAL, k_final, 1, 32, 0, st, (long long)n, ctx->state);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(ctx->h_state, ctx->state, sizeof(DevState), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    ctx->inf_rounds = rounds_total;
    if (ctx->timing) {
        for (int k = 0; k < K_NUM; k++) ctx->k_ms[k] = 0;
        for (int i = 0; i + 1 < ctx->ev_n; i += 2) {
            float ms = 0;
            CK(cudaEventElapsedTime(&ms, ctx->ev[i], ctx->ev[i + 1]));
            ctx->k_ms[ctx->ev_class[i]] += ms;
        }
    }
    if (ctx->h_state->error) return DFL_E_DATA;
    if (adler) *adler = ctx->h_state->adler;
    return DFL_OK;
}

/* pointer-jumping rounds of the last inflate call (profiling aid) */
extern "C" int dfl_b200_inflate_rounds(dfl_ctx *ctx) { return ctx ? ctx->inf_rounds : 0; }

extern "C" int dfl_b200_decompress_host(dfl_ctx *ctx, const uint8_t *in, size_t in_len, const uint64_t *seg_off, size_t nseg,
                                        uint32_t seg_bytes, uint8_t *out, size_t n)
{
    if (!ctx || !in || in_len < 6 || (!out && n) || (!seg_off && nseg)) return DFL_E_INVAL;
    /* zlib wrapper: the checks of src/deflate_decompress.c:347-368 (CM = 8, window <= 32 KiB, FCHECK, no preset dictionary) */
    if ((in[0] & 0x0F) != 8 || (in[0] >> 4) > 7 || ((in[0] << 8) | in[1]) % 31 != 0 || (in[1] & 0x20)) return DFL_E_DATA;
    CK(cudaSetDevice(ctx->device));
    const size_t body_len = in_len - 6;
    if (ctx->d_in_cap < in_len + 16) {
        if (ctx->d_in) cudaFree(ctx->d_in);
        ctx->d_in = NULL; ctx->d_in_cap = 0;
        CK(cudaMalloc(&ctx->d_in, in_len + 16));
        ctx->d_in_cap = in_len + 16;
    }
    if (ctx->d_out_cap < n + 16) {
        if (ctx->d_out) cudaFree(ctx->d_out);
        ctx->d_out = NULL; ctx->d_out_cap = 0;
        CK(cudaMalloc(&ctx->d_out, n + 16));
        ctx->d_out_cap = n + 16;
    }
    uint64_t *d_off = NULL;
    CK(cudaMalloc(&d_off, (nseg + 1) * sizeof(uint64_t)));
    cudaStream_t st = 0;
    uint32_t adler = 1;
    int rc = DFL_OK;
    if (cudaMemcpyAsync(ctx->d_in, in, in_len, cudaMemcpyHostToDevice, st) != cudaSuccess ||
        cudaMemcpyAsync(d_off, seg_off, (nseg + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, st) != cudaSuccess)
        rc = DFL_E_CUDA;
    if (!rc) rc = dfl_b200_inflate_device(ctx, ctx->d_in + 2, body_len, d_off, nseg, seg_bytes, 0, ctx->d_out, n, &adler, st);
    if (!rc && n && cudaMemcpy(out, ctx->d_out, n, cudaMemcpyDeviceToHost) != cudaSuccess) rc = DFL_E_CUDA;
    cudaFree(d_off);
    if (rc) return rc;
    const uint32_t want = ((uint32_t)in[in_len - 4] << 24) | ((uint32_t)in[in_len - 3] << 16) | ((uint32_t)in[in_len - 2] << 8) | in[in_len - 1];
    return want == adler ? DFL_OK : DFL_E_DATA; /* RFC 1950: Adler-32 of the uncompressed data, MSB first */
}

/* ------------------------------------------------------------------------- */
/* PNG: scanline filter + container                                          */
/* ------------------------------------------------------------------------- */

extern "C" int dfl_b200_png_filter_device(dfl_ctx *ctx, const uint8_t *d_pix, size_t row_bytes, size_t rows, int bpp, int mode,
                                          uint8_t *d_out, void *stream)
{
    if (!ctx || ((!d_pix || !d_out) && rows && row_bytes)) return DFL_E_INVAL;
    if (bpp < 1 || bpp > 8 || mode < 0 || mode > 5) return DFL_E_RANGE;
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaSetDevice(ctx->device));
    ctx->ev_n = 0;
    if (rows && row_bytes) {
        const unsigned int grid = (unsigned int)((rows + PNGF_WARPS - 1) / PNGF_WARPS);
        if (bpp <= 4 && row_bytes % 4 == 0 && ((uintptr_t)d_pix & 3) == 0)
            LAUNCH(K_PNG_FILTER, k_png_filter4, grid, PNGF_WARPS * 32, 0, st, d_pix, (long long)row_bytes, (long long)rows, bpp, mode, d_out);
        else
            LAUNCH(K_PNG_FILTER, k_png_filter, grid, PNGF_WARPS * 32, 0, st, d_pix, (long long)row_bytes, (long long)rows, bpp, mode, d_out);
    }
    CK(cudaGetLastError());
    if (ctx->timing) {
        CK(cudaStreamSynchronize(st));
        for (int k = 0; k < K_NUM; k++) ctx->k_ms[k] = 0;
        for (int i = 0; i + 1 < ctx->ev_n; i += 2) {
            float ms = 0;
            CK(cudaEventElapsedTime(&ms, ctx->ev[i], ctx->ev[i + 1]));
            ctx->k_ms[ctx->ev_class[i]] += ms;
        }
    }
    return DFL_OK;
}

/* CRC-32 of a PNG chunk (calc_crc, src/include/crc.h:22-30) */
static uint32_t png_crc(const uint8_t *p, size_t n)
{
    static uint32_t table[256];
    static int ready = 0;
    if (!ready) {
        for (uint32_t i = 0; i < 256; i++) {
            uint32_t c = i;
            for (int k = 0; k < 8; k++) c = (c >> 1) ^ (0xEDB88320u & (0u - (c & 1u)));
            table[i] = c;
        }
        ready = 1;
    }
    uint32_t c = 0xFFFFFFFFu;
    for (size_t i = 0; i < n; i++) c = table[(c ^ p[i]) & 0xFF] ^ (c >> 8);
    return ~c;
}

static void put_be32(uint8_t *p, uint32_t v) { p[0] = (uint8_t)(v >> 24); p[1] = (uint8_t)(v >> 16); p[2] = (uint8_t)(v >> 8); p[3] = (uint8_t)v; }

extern "C" size_t dfl_b200_png_bound(const dfl_params *prm, uint32_t width, uint32_t height, uint32_t channels)
{
    const size_t filtered = (size_t)height * ((size_t)width * channels + 1);
    return dfl_b200_bound(prm, filtered) + 6 + 8 + 25 + 12 + 12;
}

/* 8-bit grey (1), grey + alpha (2), RGB (3) or RGBA (4) pixels -> a PNG file: signature, IHDR, one IDAT
 * (filter on the GPU, compress on the GPU, zlib wrapper), IEND -- what src/png/png_decoder.c reads. */
extern "C" int dfl_b200_png_encode_host(dfl_ctx *ctx, const uint8_t *pix, uint32_t width, uint32_t height, uint32_t channels,
                                        int filter_mode, uint8_t *out, size_t cap, size_t *out_len)
{
    static const uint8_t sig[8] = {0x89, 0x50, 0x4E, 0x47, 0x0D, 0x0A, 0x1A, 0x0A}; /* PNG_HEAD, src/png/include/png_decoder.h:63 */
    static const uint8_t ctype[5] = {0, 0, 4, 2, 6};
    if (!ctx || !pix || !out || !out_len) return DFL_E_INVAL;
    if (!width || !height || channels < 1 || channels > 4) return DFL_E_RANGE;
    const size_t row_bytes = (size_t)width * channels, npix = row_bytes * height, nf = (size_t)height * (row_bytes + 1);
    if (nf + 64 > 0x7FFFFFFFull) return DFL_E_RANGE; /* one IDAT chunk */
    CK(cudaSetDevice(ctx->device));
    const dfl_params prm = ctx->prm;
    const size_t bound = dfl_b200_bound(&prm, nf);
    if (cap < 8 + 25 + 12 + 12) return DFL_E_SZ;
    if (ctx->d_in_cap < npix + 16) {
        if (ctx->d_in) cudaFree(ctx->d_in);
        ctx->d_in = NULL; ctx->d_in_cap = 0;
        CK(cudaMalloc(&ctx->d_in, npix + 16));
        ctx->d_in_cap = npix + 16;
    }
    if (ctx->d_png_cap < nf + 16) {
        if (ctx->d_png) cudaFree(ctx->d_png);
        ctx->d_png = NULL; ctx->d_png_cap = 0;
        CK(cudaMalloc(&ctx->d_png, nf + 16));
        ctx->d_png_cap = nf + 16;
    }
    if (ctx->d_out_cap < bound + 16) {
        if (ctx->d_out) cudaFree(ctx->d_out);
        ctx->d_out = NULL; ctx->d_out_cap = 0;
        CK(cudaMalloc(&ctx->d_out, bound + 16));
        ctx->d_out_cap = bound + 16;
    }
    cudaStream_t st = 0;
    CK(cudaMemcpyAsync(ctx->d_in, pix, npix, cudaMemcpyHostToDevice, st));
    int rc = dfl_b200_png_filter_device(ctx, ctx->d_in, row_bytes, height, (int)channels, filter_mode, ctx->d_png, st);
    if (rc) return rc;
    size_t body = 0;
    uint32_t adler = 1;
    rc = dfl_b200_deflate_device(ctx, ctx->d_png, nf, 0, 1, ctx->d_out, bound, &body, &adler, st);
    if (rc) return rc;
    const size_t idat = 2 + body + 4, total = 8 + 25 + 12 + idat + 12;
    if (cap < total) return DFL_E_SZ;
    uint8_t *p = out;
    memcpy(p, sig, 8); p += 8;
    put_be32(p, 13); memcpy(p + 4, "IHDR", 4);
    put_be32(p + 8, width); put_be32(p + 12, height);
    p[16] = 8; p[17] = ctype[channels]; p[18] = 0; p[19] = 0; p[20] = 0; /* depth, colour type, deflate, adaptive filtering, no interlace */
    put_be32(p + 21, png_crc(p + 4, 17)); p += 25;
    put_be32(p, (uint32_t)idat); memcpy(p + 4, "IDAT", 4);
    p[8] = 0x78; p[9] = 0x9C;
    CK(cudaMemcpy(p + 10, ctx->d_out, body, cudaMemcpyDeviceToHost));
    put_be32(p + 10 + body, adler);
    put_be32(p + 8 + idat, png_crc(p + 4, 4 + idat)); p += 12 + idat;
    put_be32(p, 0); memcpy(p + 4, "IEND", 4); put_be32(p + 8, png_crc(p + 4, 4));
    *out_len = total;
    return DFL_OK;
}
