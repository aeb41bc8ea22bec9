/* k_match_run.cuh -- k_match for runs of consecutive full-size segments: the window's index is inherited
 * (included by dfl_kernels.cu after k_match.cuh). */
#pragma once

#define RUN_SEGS_MAX 64 /* consecutive segments one CTA takes in a row at most (the first one sorts its window itself); the host picks the length */
#define RUN_BLK 32768   /* block size = frame position of the segment */
#define RUN_LOOKAHEAD 4 /* bytes past the segment staged with it: its last positions are hashed over their real 4 bytes */

/*
 * Same job as k_match (process_loop's index, match search and greedy token decision,
 * src/deflate_compress.c:243-338) for the default geometry -- segments of 32768 bytes, window 32768,
 * dictionary priming on, no chain limit -- where the window of a segment IS the segment before it.
 *
 * k_match sorts all 65536 positions of [window | segment] for every segment, so every byte is sorted twice
 * (once as segment, once as the next segment's window).  Here a CTA takes RUN_SEGS consecutive segments and
 * sorts only the segment's own 32768 positions.  All positions of the window precede all positions of the
 * segment, so a bucket of the combined index is exactly [window block's entries | segment block's entries]:
 * the combined list is the bucket-wise concatenation of the two blocks' sorted lists, and its offsets are the
 * sums of theirs.  Every sorted block is saved to global scratch (positions, the bucket of every slot, bucket
 * offsets; 160 KB per block, two slots per CTA, L2 resident) and is merged twice: now as the segment's block,
 * and by the next segment as its window's block.  The merge is two linear sweeps with one shared-memory lookup
 * per entry (the other block's offset of that bucket).  The search is k_match's (search_segment), over a list laid
 * out exactly as k_match's.  Only the first segment of a run has to sort its window itself.
 * (Keeping the two lists apart and handing a bucket walk over from one to the other was measured first: the same
 * walk steps, but 24 % more time per step of the search loop -- more than the saved sorting.)
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
    constexpr int B = RUN_BLK; /* 32768: block size, frame position of the segment */
    const uint32_t soff = 16 + (uint32_t)((uintptr_t)d_in & 15); /* frame byte P lives at smem byte soff + P (seg_lo is a multiple of 16) */
    uint8_t *sd = reinterpret_cast<uint8_t *>(S.data);
    uint16_t *gbuf = scratch + (size_t)blockIdx.x * SCRATCH_STRIDE; /* radix park: 32 Ki positions */
    uint16_t *saved[2] = {gbuf + NPOS_MAX, gbuf + NPOS_MAX + SAVED_BLK}; /* a saved block: pos[32768] | bucket[32768] | off[HASH_SIZE + 2] */
    uint32_t lbase = (uint32_t)__cvta_generic_to_shared(S.list), obase = (uint32_t)__cvta_generic_to_shared(S.off);
    const int nruns = (ns + run_segs - 1) / run_segs;
    uint32_t bar_phase = 0;
    if (tid == 0) mbar_init(&S.bar, 1);
    __syncthreads();

    /* sort the ni positions from pos0 on, save the block (sorted positions, buckets, offsets) to slot w */
    auto sort_and_save = [&](int pos0, int ni, int sb, int w) {
        if (!build_index<true, true>(S, soff, pos0, 0, ni, gbuf, tid, lane, warp, state, sb, saved[w] + B)) {
            /* the atomics did not hand out slots in lane order: redo with the ballot-ranked scatter */
            if (tid == 0) atomicAdd(&state->sort_fallbacks, 1u);
            if (!build_index<false, true>(S, soff, pos0, 0, ni, gbuf, tid, lane, warp, state, sb, saved[w] + B) && tid == 0)
                atomicExch(&state->error, 4u); /* not in order either way: the call fails rather than emit tokens from a wrong index */
        }
        for (int k8 = tid * 8; k8 < ni; k8 += MATCH_THREADS * 8)
            __stcg(reinterpret_cast<uint4 *>(saved[w] + k8), *reinterpret_cast<const uint4 *>(&S.list[k8]));
        for (int b2 = tid; b2 < (HASH_SIZE + 2) / 2; b2 += MATCH_THREADS)
            __stcg(reinterpret_cast<uint32_t *>(saved[w] + 2 * B) + b2, reinterpret_cast<const uint32_t *>(S.off)[b2]);
        __syncthreads();
    };

    for (;;) {
        /* runs are handed out by a counter: a CTA that was placed late (or drew cheap segments) does not keep a fixed share */
        if (tid == 0) S.next_chunk = (int)atomicAdd(&state->run_next, 1u);
        __syncthreads();
        const int run = S.next_chunk;
        __syncthreads();
        if (run >= nruns) break;
        const int sA = run * run_segs, sB = sA + run_segs < ns ? sA + run_segs : ns; /* segments of this launch */
        int sb = DFL_SUB_BITS;
        int ni_win = 0, w_win = 0; /* the window's block: entries, slot it is saved in */
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
            const int ni_win_fresh = (lim < B ? lim : B) - P0 > 0 ? (lim < B ? lim : B) - P0 : 0;
            const int ni_seg = (lim < B + seglen ? lim : B + seglen) - B > 0 ? (lim < B + seglen ? lim : B + seglen) - B : 0;

            if (first) {
                /* bucket width for the run (k_match decides per segment): plain trigram buckets for data with hardly
                 * any repetition, 4-byte buckets otherwise; sum of squared trigram-bucket counts of a 1-in-8 sample */
                uint32_t *off32 = reinterpret_cast<uint32_t *>(S.off);
                const uint32_t dbase_ = (uint32_t)__cvta_generic_to_shared(S.data) + soff + P0;
                const int NI = lim - P0 > 0 ? lim - P0 : 0;
                for (int i2 = tid; i2 < (HASH_SIZE + 2) / 2; i2 += MATCH_THREADS) off32[i2] = 0;
                __syncthreads();
                for (int k8 = tid * 8; k8 < NI; k8 += MATCH_THREADS * 8) {
                    const uint32_t b = DFL_HASH3(ld4s(dbase_ + k8));
                    reds_add(obase + 4 * (b >> 1), 1u << ((b & 1) * 16));
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
                /* the window's block, sorted here for the first segment of the run only */
                w_win = 0;
                ni_win = ni_win_fresh;
                sort_and_save(P0, ni_win, sb, w_win);
            }
            if (tid == 0 && sb == 0) atomicAdd(&state->t_phase[20], 1ull); /* segments indexed by plain trigram buckets */

            /* ---- the segment's block: sorted and saved (the next segment merges it again, as its window) ---- */
            const int w_seg = w_win ^ 1;
            sort_and_save(B, ni_seg, sb, w_seg); /* S.off = the segment block's offsets */

            /* The combined list has 16-bit offsets: 65536 entries (a full window and a full segment) are one too many.
             * The segment's last position is no candidate for any head of this segment (it is the largest one) and is
             * left out of the combined list; it stays in the saved block, the next segment needs it.  It is the last
             * entry of its bucket's part of the segment block's list. */
            int k_drop = 0x7FFFFFFF, b_drop = 0x7FFFFFFF;
            if (ni_win + ni_seg > 65535) {
                b_drop = (int)hash4(ld4s((uint32_t)__cvta_generic_to_shared(S.data) + soff + B + ni_seg - 1), HASH_SIZE >> sb, sb);
                k_drop = (int)lds_u16(obase + 2 * b_drop + 2) - 1;
            }
            /* ---- merge: bucket b of the combined list = [window block's entries of b | segment block's entries of b] ---- */
            {
                const uint16_t *wp = saved[w_win], *wb = wp + B, *sp = saved[w_seg], *sbk = sp + B;
                /* a window entry at slot k of its list, bucket b, lands at k + (segment block's offset of b), less the
                 * position left out if that one's bucket comes before b */
                /* eight entries per thread and step (two 16-byte loads), the next step's loads issued before this step's stores */
                {
                    int k8 = tid * 8;
                    uint4 pp = make_uint4(0, 0, 0, 0), bb = pp;
                    if (k8 < ni_win) { pp = __ldcg(reinterpret_cast<const uint4 *>(wp + k8)); bb = __ldcg(reinterpret_cast<const uint4 *>(wb + k8)); }
                    while (k8 < ni_win) {
                        const int kn = k8 + MATCH_THREADS * 8;
                        uint4 pn = make_uint4(0, 0, 0, 0), bn = pn;
                        if (kn < ni_win) { pn = __ldcg(reinterpret_cast<const uint4 *>(wp + kn)); bn = __ldcg(reinterpret_cast<const uint4 *>(wb + kn)); }
                        const uint32_t pw_[4] = {pp.x, pp.y, pp.z, pp.w}, bw_[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
                        for (int j = 0; j < 8; j++) {
                            const uint32_t pos = (pw_[j >> 1] >> ((j & 1) * 16)) & 0xFFFF, bk = (bw_[j >> 1] >> ((j & 1) * 16)) & 0xFFFF;
                            if (k8 + j < ni_win) /* (a block saved as the segment's holds 32768 + position) */
                                sts_u16(lbase + 2 * (k8 + j + lds_u16(obase + 2 * bk) - ((int)bk > b_drop ? 1 : 0)), pos & (B - 1));
                        }
                        pp = pn; bb = bn; k8 = kn;
                    }
                }
                __syncthreads();
                /* S.off <- the window block's offsets; a segment entry at slot k, bucket b, lands at k + (window block's offset of b + 1) */
                for (int b2 = tid; b2 < (HASH_SIZE + 2) / 2; b2 += MATCH_THREADS)
                    reinterpret_cast<uint32_t *>(S.off)[b2] = __ldcg(reinterpret_cast<const uint32_t *>(wp + 2 * B) + b2);
                __syncthreads();
                {
                    int k8 = tid * 8;
                    uint4 pp = make_uint4(0, 0, 0, 0), bb = pp;
                    if (k8 < ni_seg) { pp = __ldcg(reinterpret_cast<const uint4 *>(sp + k8)); bb = __ldcg(reinterpret_cast<const uint4 *>(sbk + k8)); }
                    while (k8 < ni_seg) {
                        const int kn = k8 + MATCH_THREADS * 8;
                        uint4 pn = make_uint4(0, 0, 0, 0), bn = pn;
                        if (kn < ni_seg) { pn = __ldcg(reinterpret_cast<const uint4 *>(sp + kn)); bn = __ldcg(reinterpret_cast<const uint4 *>(sbk + kn)); }
                        const uint32_t pw_[4] = {pp.x, pp.y, pp.z, pp.w}, bw_[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
                        for (int j = 0; j < 8; j++) {
                            const uint32_t pos = (pw_[j >> 1] >> ((j & 1) * 16)) & 0xFFFF, bk = (bw_[j >> 1] >> ((j & 1) * 16)) & 0xFFFF;
                            const int kk = k8 + j;
                            if (kk < ni_seg && kk != k_drop) sts_u16(lbase + 2 * (kk - (kk > k_drop ? 1 : 0) + lds_u16(obase + 2 * bk + 2)), pos);
                        }
                        pp = pn; bb = bn; k8 = kn;
                    }
                }
                __syncthreads();
                /* combined offsets: the sum of the two blocks' (entry HASH_SIZE + 1 = the total) */
                for (int b2 = tid; b2 < (HASH_SIZE + 2) / 2; b2 += MATCH_THREADS) { /* two offsets per word; no sum exceeds 65535 */
                    const uint32_t dec = (2 * b2 > b_drop ? 1u : 0u) | (2 * b2 + 1 > b_drop ? 0x10000u : 0u);
                    reinterpret_cast<uint32_t *>(S.off)[b2] += __ldcg(reinterpret_cast<const uint32_t *>(sp + 2 * B) + b2) - dec;
                }
                __syncthreads();
            }
            const int ni_all = ni_win + ni_seg - (k_drop != 0x7FFFFFFF ? 1 : 0);
            const long long t2 = clock64();
            if (tid == 0) {
                atomicAdd(&state->t_phase[0], (unsigned long long)(t1 - t0));
                atomicAdd(&state->t_phase[1], (unsigned long long)(t2 - t1));
            }

            /* ---- the search (k_search.cuh), over a list laid out as k_match's ---- */
            if (sb == 0) { /* hardly any repetition: every position's match, flat over the list */
                search_flat(S, soff, B, seglen, B, ni_all, lim, tokcand + (size_t)sl * B, tid);
                if (k_drop != 0x7FFFFFFF && tid == 0) tokcand[(size_t)sl * B + seglen - 1] = sd[soff + B + seglen - 1]; /* the position left out of the list */
            } else if (warp < SEARCH_WARPS)
                search_segment<false>(S, soff, B, seglen, B, 0, sb, tokcand + (size_t)sl * B, state, lane, warp, t2);
            __syncthreads();
            ni_win = ni_seg;
            w_win = w_seg;
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
    if (tid == 0 && atomicAdd(&state->run_done, 1u) == gridDim.x - 1) { /* the last CTA out resets the counters for the next launch */
        state->run_next = 0;
        state->run_done = 0;
    }
}
