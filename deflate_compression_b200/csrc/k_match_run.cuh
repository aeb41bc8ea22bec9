/* k_match_run.cuh -- k_match for runs of consecutive full-size segments: the window's index is inherited
 * (included by dfl_kernels.cu after k_match.cuh). */
#pragma once

#ifndef RUN_SEGS
#define RUN_SEGS 8 /* consecutive segments one CTA takes in a row (the first one sorts its window itself) */
#endif
#define RUN_LOOKAHEAD 4 /* bytes past the segment staged with it: its last positions are hashed over their real 4 bytes */

/*
 * Same job as k_match (process_loop's index, match search and greedy token decision,
 * src/deflate_compress.c:243-338) for the default geometry -- segments of 32768 bytes, window 32768,
 * dictionary priming on, no chain limit -- where the window of a segment IS the segment before it.
 *
 * k_match sorts all 65536 positions of [window | segment] for every segment, so every byte is sorted twice
 * (once as segment, once as the next segment's window).  Here a CTA takes RUN_SEGS consecutive segments.
 * All positions of the window precede all positions of the segment, so a bucket of the combined index is
 * exactly [window block's entries | segment block's entries]: the two blocks keep separate lists, only the
 * segment block is sorted, and the list it leaves behind becomes the window block's list of the next segment
 * (positions rebased by 32768; bucket offsets packed two per word into global scratch, which stays in L2).
 * search_segment<_, TWO = true> walks a bucket through the segment block's list first and then through the
 * window block's.  Only the first segment of a run has to sort its window itself.
 *
 * Frame: position P of the shared-memory frame is input byte seg_lo - 32768 + P; the window block is
 * [32768 - hist, 32768), the segment block [32768, 32768 + seglen).  The RUN_LOOKAHEAD bytes after the
 * segment are staged too, so that the last positions of a block are hashed over the same four bytes now and
 * when the block has become a window.
 */
__global__ void __launch_bounds__(MATCH_THREADS, 1)
k_match_run(const uint8_t *__restrict__ d_in, long long n, long long avail_hist, long long seg0, int ns, int run_segs,
            uint32_t *__restrict__ tokcand, uint16_t *__restrict__ scratch, DevState *state)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    MatchSmem &S = *reinterpret_cast<MatchSmem *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int B = RUN_BLK; /* 32768: block size, frame position of the segment, list slot of the window block */
    const uint32_t soff = 16 + (uint32_t)((uintptr_t)d_in & 15); /* frame byte P lives at smem byte soff + P (seg_lo is a multiple of 16) */
    uint8_t *sd = reinterpret_cast<uint8_t *>(S.data);
    uint16_t *gbuf = scratch + (size_t)blockIdx.x * SCRATCH_STRIDE;       /* radix park: 32 Ki positions */
    uint32_t *oo = reinterpret_cast<uint32_t *>(gbuf + B);                 /* window block's offsets: oo[b] = off[b] | off[b + 1] << 16 */
    const int nruns = (ns + run_segs - 1) / run_segs;
    uint32_t bar_phase = 0;
    if (tid == 0) mbar_init(&S.bar, 1);
    __syncthreads();

    for (int run = blockIdx.x; run < nruns; run += gridDim.x) {
        const int sA = run * run_segs, sB = sA + run_segs < ns ? sA + run_segs : ns; /* segments of this launch */
        int sb = DFL_SUB_BITS;
        int ni_new = 0; /* entries of the segment block's list (of the previous segment when the loop body starts) */
        for (int sl = sA; sl < sB; sl++) {
            const long long t0 = clock64();
            const bool first = sl == sA;
            const long long seg_lo = (seg0 + sl) * B;
            const long long seg_hi = seg_lo + B < n ? seg_lo + B : n;
            const int seglen = (int)(seg_hi - seg_lo);
            const long long before = seg_lo + avail_hist; /* bytes that exist before the segment */
            const int hist = before < B ? (int)before : B;
            const int la = n - seg_hi < RUN_LOOKAHEAD ? (int)(n - seg_hi) : RUN_LOOKAHEAD;
            const int P0 = B - hist, Pend = B + seglen + la; /* staged frame positions */
            const uint8_t *g0 = d_in + (seg_lo - hist);      /* = frame position P0 */

            /* ---- stage [window | segment | lookahead] (TMA bulk copy for the aligned interior) ---- */
            {
                const int npos = Pend - P0;
                int head_n = (int)((16 - ((uintptr_t)g0 & 15)) & 15);
                if (head_n > npos) head_n = npos;
                const int bulk_n = (npos - head_n) & ~15;
                const int tail0 = head_n + bulk_n;
                if (tid == 0) S.next_chunk = 0;
                for (int i = tid; i < DFL_MAX_SEG / 32; i += MATCH_THREADS) S.claimed[i] = 0;
                if (tid == 0 && bulk_n > 0) {
                    /* the previous segment's reads of this memory (generic proxy) are ordered before the bulk writes */
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    mbar_expect_tx(&S.bar, (uint32_t)bulk_n);
                    for (int o = 0; o < bulk_n; o += 16384) {
                        const int c = bulk_n - o < 16384 ? bulk_n - o : 16384;
                        tma_bulk_g2s(sd + soff + P0 + head_n + o, g0 + head_n + o, (uint32_t)c, &S.bar);
                    }
                }
                if (tid < head_n) sd[soff + P0 + tid] = g0[tid];
                for (int i = tail0 + tid; i < npos; i += MATCH_THREADS) sd[soff + P0 + i] = g0[i];
                if (tid < 24) sd[soff + Pend + tid] = 0; /* zero slack: over-reads past the data are deterministic */
                if (bulk_n > 0) mbar_wait(&S.bar, bar_phase);
                if (bulk_n > 0) bar_phase ^= 1;
                __syncthreads();
            }
            const long long t1 = clock64();

            /* ---- Adler-32 partial sums of the segment bytes ---- */
            {
                uint32_t a = 0;
                unsigned long long b = 0;
                for (int i = tid; i < seglen; i += MATCH_THREADS) {
                    const uint32_t v = sd[soff + B + i];
                    a += v;
                    b += (unsigned long long)v * (uint32_t)(seglen - i);
                }
                for (int o = 16; o; o >>= 1) {
                    a += __shfl_xor_sync(FULL, a, o);
                    b += __shfl_xor_sync(FULL, b, o);
                }
                if (lane == 0) { S.red_a[warp] = a; S.red_b[warp] = b; }
            }

            /* positions with a whole trigram inside the data: P + 3 <= Pend */
            const int lim = Pend - 2;
            const int ni_old_fresh = (lim < B ? lim : B) - P0 > 0 ? (lim < B ? lim : B) - P0 : 0;
            const int ni_seg = (lim < B + seglen ? lim : B + seglen) - B > 0 ? (lim < B + seglen ? lim : B + seglen) - B : 0;

#ifdef RUN_SINGLE /* experiment: k_match's own algorithm (one list over window + segment) inside this kernel's loop */
            {
                const int NI = lim - P0 > 0 ? lim - P0 : 0;
                if (!build_index<true>(S, soff, P0, 0, NI, gbuf, tid, lane, warp, state, sb)) build_index<false>(S, soff, P0, 0, NI, gbuf, tid, lane, warp, state, sb);
                const long long t2 = clock64();
                if (tid == 0) { atomicAdd(&state->t_phase[0], (unsigned long long)(t1 - t0)); atomicAdd(&state->t_phase[1], (unsigned long long)(t2 - t1)); }
                if (warp < SEARCH_WARPS) search_segment<false, false>(S, soff, B, seglen, B, 0, sb, tokcand + (size_t)sl * B, oo, state, lane, warp, t2);
                __syncthreads();
                if (tid == 0) atomicAdd(&state->t_phase[2], (unsigned long long)(clock64() - t2));
                __syncthreads();
                continue;
            }
#endif
            if (first) {
                /* bucket width for the run (k_match decides per segment): plain trigram buckets for data with hardly
                 * any repetition, 4-byte buckets otherwise; sum of squared trigram-bucket counts of a 1-in-8 sample */
                uint32_t *off32 = reinterpret_cast<uint32_t *>(S.off);
                const uint32_t obase_ = (uint32_t)__cvta_generic_to_shared(S.off), dbase_ = (uint32_t)__cvta_generic_to_shared(S.data) + soff + P0;
                const int NI = lim - P0 > 0 ? lim - P0 : 0;
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
                sb = sq <= SPLIT_MIN_SQ * ((uint32_t)(NI + 7) / 8) ? 0 : DFL_SUB_BITS;
                /* the window block's list, sorted here for the first segment of the run only */
                if (!build_index<true>(S, soff, P0, B, ni_old_fresh, gbuf, tid, lane, warp, state, sb)) {
                    if (tid == 0) atomicAdd(&state->sort_fallbacks, 1u);
                    build_index<false>(S, soff, P0, B, ni_old_fresh, gbuf, tid, lane, warp, state, sb);
                }
            } else {
                /* the previous segment's block has become the window: its list moves to the window block's slots,
                 * rebased by one block (its entries are 32768 + position) */
                for (int k8 = tid * 8; k8 < ni_new; k8 += MATCH_THREADS * 8) {
                    uint4 v = *reinterpret_cast<const uint4 *>(&S.list[k8]);
                    v.x &= 0x7FFF7FFFu; v.y &= 0x7FFF7FFFu; v.z &= 0x7FFF7FFFu; v.w &= 0x7FFF7FFFu;
                    *reinterpret_cast<uint4 *>(&S.list[B + k8]) = v;
                }
            }
            /* the window block's bucket offsets go to global scratch, two per word (S.off still holds them) */
            for (int b = tid; b <= HASH_SIZE; b += MATCH_THREADS) __stcg(oo + b, (uint32_t)S.off[b] | ((uint32_t)S.off[b + 1] << 16));
            if (tid == 0 && sb == 0) atomicAdd(&state->t_phase[20], 1ull); /* segments indexed by plain trigram buckets */
            __threadfence(); /* gpu scope: also drops this SM's L1 lines of the previous segment's offsets (the search reads them through L1) */
            __syncthreads();

            /* ---- the segment block's list ---- */
            if (!build_index<true>(S, soff, B, 0, ni_seg, gbuf, tid, lane, warp, state, sb)) {
                if (tid == 0) atomicAdd(&state->sort_fallbacks, 1u);
                build_index<false>(S, soff, B, 0, ni_seg, gbuf, tid, lane, warp, state, sb);
            }
            ni_new = ni_seg;
            const long long t2 = clock64();
            if (tid == 0) {
                atomicAdd(&state->t_phase[0], (unsigned long long)(t1 - t0));
                atomicAdd(&state->t_phase[1], (unsigned long long)(t2 - t1));
            }

            /* ---- speculative chunk parsers (k_search.cuh) ---- */
            if (sb == 0) /* hardly any repetition: every position's match, flat over the list */
                search_flat<true>(S, soff, B, seglen, B, ni_seg, 0, tokcand + (size_t)sl * B, oo, tid);
            else if (warp < SEARCH_WARPS)
                search_segment<false, true>(S, soff, B, seglen, B, 0, sb, tokcand + (size_t)sl * B, oo, state, lane, warp, t2);
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
                    const unsigned long long am = a % ADLER_MOD;
                    const long long hi_ = (seg0 + sl + 1) * B;
                    const unsigned long long rest = (unsigned long long)(n - (hi_ < n ? hi_ : n)) % ADLER_MOD;
                    const unsigned long long bm = (b % ADLER_MOD + am * rest) % ADLER_MOD;
                    atomicAdd(&state->adl_a, am);
                    atomicAdd(&state->adl_b, bm);
                    atomicAdd(&state->t_phase[2], (unsigned long long)(clock64() - t2));
                }
            }
            __syncthreads(); /* red_a / red_b / scan_w are reused by the next segment */
        }
    }
}
