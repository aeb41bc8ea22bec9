/* k_match.cuh -- index build + match search + speculative greedy parse (included by dfl_kernels.cu). */
#pragma once
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
#ifndef LONG_CMP_FROM
#define LONG_CMP_FROM 16 /* bytes a lane compares alone before the warp takes a long match over */
#endif
#ifndef SKIP_OVERRUN
#define SKIP_OVERRUN 1   /* a chunk some chain has already entered is not started on its own */
#endif
#ifndef START_AFTER_DELIM
#define START_AFTER_DELIM 1 /* a speculative chain starts after the first delimiter byte of its chunk's first 8, if there is one */
#endif
#ifndef DELIM_MAX
#define DELIM_MAX 0x20
#endif
#define SORT_WARPS 32    /* every warp sorts one stripe */
#define SORT_BINS 128    /* 7-bit digits: bucket = digit2 * 128 + digit1 */
#define SCRATCH_SLOTS 256
/* u16 units of global scratch per slot (a slot belongs to one resident CTA): the radix park (64 Ki positions for
 * k_match; k_match_run uses the first 32 Ki) and k_match_run's two saved blocks (sorted positions, their buckets,
 * bucket offsets), which take turns as "the segment's block" and "the window's block" */
#define SAVED_BLK (32768 + 32768 + 16384)
#define SCRATCH_STRIDE (NPOS_MAX + 2 * SAVED_BLK)

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
template <int PASS, bool FAST, bool BKC>
__device__ __forceinline__ void scatter_sweep(MatchSmem &S, uint32_t soff, int pos0, int lofs, int k_lo, int k_hi, int SS,
                                              const uint16_t *gbuf, int lane, int warp, int sb)
{
    /* BKC: the bucket of position pos0 + i is cached at S.list[DFL_MAX_SEG + i] (see build_index) */
    uint32_t kbase = (uint32_t)__cvta_generic_to_shared(S.list) + 2 * DFL_MAX_SEG - 2 * (uint32_t)pos0;
    asm volatile("" : "+r"(kbase));
    const uint32_t ss_inv = SS ? 0xFFFFFFFFu / (uint32_t)SS + 1u : 0u; /* x / SS == umulhi(x, ss_inv) for x * SS < 2^32 */
    const uint32_t ngrp = HASH_SIZE >> sb;
    /* pass-1 cursors at S.off[0..), pass-2 counters / cursors behind them */
    uint32_t *cnt32 = reinterpret_cast<uint32_t *>(S.off) + (PASS == 1 ? 0 : CNT_WORDS);
    volatile uint16_t *cnt16 = reinterpret_cast<volatile uint16_t *>(cnt32);
    uint32_t dbase = (uint32_t)__cvta_generic_to_shared(S.data) + soff;
    uint32_t lbase = (uint32_t)__cvta_generic_to_shared(S.list) + 2 * lofs;
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
            e[j] = PASS == 1 ? (uint32_t)(pos0 + k) : eq[j];
            if (PASS == 2 && k + 128 < k_hi) eq[j] = __ldcg(gbuf + k + 128);
            const uint32_t b = BKC ? lds_u16(kbase + 2 * (k < k_hi ? e[j] : (uint32_t)pos0))
                                   : hash4(ld4s(dbase + (k < k_hi ? e[j] : (uint32_t)pos0)), ngrp, sb);
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
template <int PASS, bool FAST, bool BKC>
__device__ __forceinline__ void sort_pass(MatchSmem &S, uint32_t soff, int pos0, int lofs, int NI, const uint16_t *gbuf, int tid,
                                          int lane, int warp, DevState *state, int sb)
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
            const uint32_t w0 = ld4s(dbase + pos0 + k8), w1 = ld4s(dbase + pos0 + k8 + 4), w2 = ld4s(dbase + pos0 + k8 + 8);
            const uint32_t sw = __umulhi((uint32_t)k8, ss_inv); /* 8 | SS: the 8 positions share a stripe */
            uint32_t bk[8];
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const uint32_t t = j < 4 ? __funnelshift_r(w0, w1, (j & 3) * 8) : __funnelshift_r(w1, w2, (j & 3) * 8);
                bk[j] = hash4(t, ngrp, sb);
                if (k8 + j < NI) {
                    const uint32_t d = bk[j] & (SORT_BINS - 1);
                    reds_add(cbase + 4 * cnt_word(sw, d), 1u << cnt_shift(d));
                }
            }
            if (BKC) /* the position's bucket is hashed once: both scatters and the offsets sweep read it back */
                *reinterpret_cast<uint4 *>(&S.list[DFL_MAX_SEG + k8]) =
                    make_uint4(bk[0] | (bk[1] << 16), bk[2] | (bk[3] << 16), bk[4] | (bk[5] << 16), bk[6] | (bk[7] << 16));
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
        scatter_sweep<PASS, FAST, BKC>(S, soff, pos0, lofs, k_lo, k_hi, SS, gbuf, lane, warp, sb);
    }
    __syncthreads();
    STAMP2(10);
#undef STAMP2
}

/* Sort the NI positions [pos0, pos0 + NI) by (bucket, position) into S.list[lofs, lofs + NI) and fill the
 * bucket offsets (relative to lofs); bkt_out (global, optional) receives the bucket of every slot.  Returns (to
 * every thread) whether the order verified. */
template <bool FAST, bool BKC = false>
__device__ __noinline__ bool build_index(MatchSmem &S, uint32_t soff, int pos0, int lofs, int NI, uint16_t *gbuf, int tid, int lane,
                                            int warp, DevState *state, int sb, uint16_t *bkt_out = nullptr)
{
    const uint32_t ngrp = HASH_SIZE >> sb;
    uint32_t dbase = (uint32_t)__cvta_generic_to_shared(S.data) + soff;
    uint32_t lbase = (uint32_t)__cvta_generic_to_shared(S.list) + 2 * lofs;
    uint32_t obase = (uint32_t)__cvta_generic_to_shared(S.off);
    /* BKC (k_match_run: lofs == 0 and NI <= DFL_MAX_SEG, the upper half of the list is free while a block is sorted):
     * bucket of position pos0 + i cached at S.list[DFL_MAX_SEG + i] by pass 1's counting sweep */
    uint32_t kbase = (uint32_t)__cvta_generic_to_shared(S.list) + 2 * DFL_MAX_SEG - 2 * (uint32_t)pos0;
    asm volatile("" : "+r"(dbase), "+r"(lbase), "+r"(obase), "+r"(kbase));
    long long c0 = clock64(), c1;
#define STAMP(slot) do { c1 = clock64(); if (tid == 0) atomicAdd(&state->t_phase[slot], (unsigned long long)(c1 - c0)); c0 = c1; } while (0)
    sort_pass<1, FAST, BKC>(S, soff, pos0, lofs, NI, gbuf, tid, lane, warp, state, sb);
    STAMP(4);
    /* park the pass-1 order in global scratch: linear, 16 B per thread */
    for (int k8 = tid * 8; k8 < NI; k8 += MATCH_THREADS * 8)
        __stcg(reinterpret_cast<uint4 *>(gbuf + k8), *reinterpret_cast<const uint4 *>(&S.list[lofs + k8]));
    STAMP(5);
    sort_pass<2, FAST, BKC>(S, soff, pos0, lofs, NI, gbuf, tid, lane, warp, state, sb);
    STAMP(6);
    /* bucket offsets from the sorted order (off[b] = first slot whose bucket is >= b), and
     * the check that every bucket is in ascending position order.  Slot kk is compared with
     * slot kk-1; consecutive threads take consecutive slots. */
    bool bad = false;
    if (NI == 0) {
        for (int i = tid; i < HASH_SIZE + 2; i += MATCH_THREADS) S.off[i] = 0;
        __syncthreads();
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
        for (int j = 0; j < 4; j++)
            b[j] = kk0 + j < NI ? (int)(BKC ? lds_u16(kbase + 2 * q[j]) : hash4(ld4s(dbase + q[j]), ngrp, sb)) : HASH_SIZE;
        /* k_match_run keeps every slot's bucket beside the list it saves for the next segment's merge */
        if (bkt_out && kk0 < NI) __stcg(reinterpret_cast<uint2 *>(bkt_out + kk0), make_uint2((uint32_t)b[0] | ((uint32_t)b[1] << 16), (uint32_t)b[2] | ((uint32_t)b[3] << 16)));
        int pq = __shfl_up_sync(FULL, q[3], 1), pb_ = __shfl_up_sync(FULL, b[3], 1);
        if (lane == 0) {
            pq = kk0 > 0 && kk0 <= NI ? (int)lds_u16(lbase + 2 * (kk0 - 1)) : -1;
            pb_ = kk0 > 0 && kk0 <= NI ? (int)(BKC ? lds_u16(kbase + 2 * pq) : hash4(ld4s(dbase + pq), ngrp, sb)) : -1;
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

#include "k_search.cuh"

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
        uint16_t *gbuf = scratch + (size_t)(smid % SCRATCH_SLOTS) * SCRATCH_STRIDE; /* one CTA per SM: private slot */
        if (!build_index<true>(S, soff, 0, 0, NI, gbuf, tid, lane, warp, state, sb)) {
            /* the atomics did not hand out slots in lane order: redo with the ballot-ranked scatter */
            if (tid == 0) atomicAdd(&state->sort_fallbacks, 1u);
            if (!build_index<false>(S, soff, 0, 0, NI, gbuf, tid, lane, warp, state, sb) && tid == 0)
                atomicExch(&state->error, 4u); /* the index is not in order either way (a scratch slot shared with another launch?): the call fails */
        }
    }
    const long long t2 = clock64();

    /* ---- speculative chunk parsers (k_search.cuh) ---- */
    if (!HAS_K && sb == 0) /* hardly any repetition: every position's match, flat over the list */
        search_flat(S, soff, hist, seglen, prm.window, NI, NI, tokcand + (size_t)blockIdx.x * prm.seg_bytes, tid);
    else if (warp < SEARCH_WARPS)
        search_segment<HAS_K>(S, soff, hist, seglen, prm.window, HAS_K ? prm.chain_limit : 0, sb,
                              tokcand + (size_t)blockIdx.x * prm.seg_bytes, state, lane, warp, t2);
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
