/* k_search.cuh -- the match search + speculative greedy parse of one segment (included by k_match.cuh;
 * shared by k_match and k_match_run). */
#pragma once

/*
 * process_loop's match search AND its greedy token decision (src/deflate_compress.c:243-288) for one
 * segment whose bytes and index are in shared memory: S.list holds every position of [window | segment]
 * sorted by (bucket, position), S.off the bucket offsets; position `hist` of that frame is the segment's first byte.
 *
 * The longest / nearest match M(p) = argmax_{q in [max(lo, p-window), p-1]} lcp(q, p) (ties ->
 * nearest, capped at min(258, seg_hi - p)) is a pure function of the bytes, because every position is
 * indexed and buckets are walked to the end.  Only parse heads need it.  The segment is cut into 32-byte
 * chunks; every lane takes a chunk, parses greedily from the chunk start as if it were a head (:265-288:
 * len >= 3 -> advance len, else literal, advance 1) and KEEPS GOING past the chunk end until it steps on a
 * position another lane has already claimed (shared bitmap, atomicOr) or leaves the segment.  By induction
 * from position 0 (claimed by chunk 0's lane) every head of the true greedy parse is claimed and evaluated;
 * speculative parses re-synchronise with the true one within a few tokens.  k_parse then follows the evaluated heads.
 */
template <bool HAS_K>
__device__ __forceinline__ void search_segment(MatchSmem &S, const uint32_t soff, const int hist, const int seglen, const int W,
                                               const int K, const int sb, uint32_t *__restrict__ out, DevState *state,
                                               const int lane, const int warp, const long long t2)
{
    enum { ST_IDLE = 0, ST_SEEK = 1, ST_WALK = 2, ST_CAND = 3, ST_DONE = 4, ST_FIN = 5, ST_HEAVY = 6, ST_LONG = 7 };
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
        /* phase B: full compare (check_dup_str) for the lanes holding a candidate.  A lane compares up
         * to LONG_CMP_FROM bytes on its own; a match still running then (logs, runs, images: up to 258
         * bytes = 64 dependent steps for one lane) is finished by the whole warp, 128 bytes per step. */
#define APPLY_LEN(len_)                                                                               \
        do {                                                                                      \
            int l_ = (len_);                                                                      \
            if (l_ > maxlen) l_ = maxlen;                                                         \
            st = ST_WALK;                                                                         \
            if (l_ > best_len) {                                                                  \
                best_len = l_;                                                                    \
                best_dist = p - q;                                                                \
                if (l_ == maxlen) st = ST_DONE;                                                   \
                else { foff = l_ - 3; fmask = 0xFFFFFFFFu; fword = ld4s(dbase + p + foff); }      \
            }                                                                                     \
        } while (0)
        int long_from = 0;
        if (st == ST_CAND) {
            int len;
            const uint32_t x0 = ld4s(dbase + q) ^ pw0, x1 = ld4s(dbase + q + 4) ^ pw1; /* both loads in flight together */
            if (x0) {
                len = (__ffs(x0) - 1) >> 3;
            } else if (x1) {
                len = 4 + ((__ffs(x1) - 1) >> 3);
            } else {
                len = 8;
                while (len < maxlen && len < LONG_CMP_FROM) {
                    const uint32_t x = ld4s(dbase + p + len) ^ ld4s(dbase + q + len);
                    if (x) { len += (__ffs(x) - 1) >> 3; long_from = -1; break; }
                    len += 4;
                }
                if (long_from == 0 && len < maxlen) long_from = len; /* still equal at LONG_CMP_FROM bytes */
                else long_from = 0;
            }
            if (long_from) st = ST_LONG;
            else APPLY_LEN(len);
        }
        /* phase C: finished heads emit their token and move on; idle lanes take new
         * chunks.  One warp reduction per round steers it (votes and shuffles share a
         * slow unit on this chip). */
        const uint32_t census = __reduce_add_sync(FULL, (st == ST_WALK ? 1u : 0u) | (st == ST_FIN ? 0x100u : 0u) |
                                                            (st == ST_HEAVY ? 0x10000u : 0u) | (st == ST_LONG ? 0x1000000u : 0u));
        const int n_busy = census & 0xFF, n_fin = (census >> 8) & 0xFF, n_heavy = (census >> 16) & 0xFF, n_long = census >> 24;
        if (n_fin == 32) break;
        if (n_long)
        {
            uint32_t lng = __ballot_sync(FULL, st == ST_LONG);
            while (lng) {
                const int L = __ffs(lng) - 1;
                lng &= lng - 1;
                const int hp = __shfl_sync(FULL, p, L), hq = __shfl_sync(FULL, q, L), hmax = __shfl_sync(FULL, maxlen, L);
                const int l0 = __shfl_sync(FULL, long_from, L);
                int flen = hmax;
                for (int base = l0; base < hmax; base += 128) {
                    const int off_ = base + 4 * lane;
                    const uint32_t x = off_ < hmax ? ld4s(dbase + hp + off_) ^ ld4s(dbase + hq + off_) : 1u; /* the end reads as a mismatch */
                    const uint32_t m = __ballot_sync(FULL, x != 0);
                    if (m) {
                        const int f = __ffs(m) - 1;
                        const uint32_t xf = __shfl_sync(FULL, x, f);
                        const int fo = base + 4 * f;
                        flen = fo < hmax ? fo + ((__ffs(xf) - 1) >> 3) : hmax;
                        break;
                    }
                }
                if (lane == L) APPLY_LEN(flen);
            }
        }
#undef APPLY_LEN
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
                    int ci = c * MATCH_CHUNK;
#if START_AFTER_DELIM
                    /* A chain may start anywhere in its chunk (chains only add parallelism; the one from position 0 finds
                     * every head by itself).  Right after a delimiter -- a byte <= ' ' within the chunk's first 8 -- it is more
                     * likely to start on a head of the true parse and merges with it a head sooner: 2 % fewer search cycles
                     * on text and log lines, nothing on data without such bytes. */
                    if (c && ci + 16 <= seglen) {
                        const uint32_t w0 = ld4s(dbase + hist + ci - 1), w1 = ld4s(dbase + hist + ci + 3);
#pragma unroll
                        for (int j = 0; j < 8; j++)
                            if ((((j < 4 ? w0 : w1) >> ((j & 3) * 8)) & 0xFF) <= DELIM_MAX) { ci += j; break; }
                    }
#endif
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
#ifndef DFL_SEARCH_STATS
    (void)state; (void)t2; (void)warp;
#endif
#undef INIT_HEAD
}

/*
 * The same search for segments indexed by plain trigram buckets (sb == 0: data with hardly any repetition --
 * ciphertext, compressed payloads; a bucket holds a handful of positions): M(p) for EVERY position of the segment, a
 * thread per list slot, no parse heads, no chains, no claims.  Slot k holds position p; its candidates are the slots
 * before it in its bucket, nearest first, down to the window's edge.  Longest wins, the nearest on ties (strict '>',
 * src/deflate_compress.c:258); k_parse then follows the chain of tokens from position 0.  Positions the index does not
 * hold (the last ones of the input, which have no whole trigram) are literals.
 */
__device__ __forceinline__ void search_flat(MatchSmem &S, const uint32_t soff, const int hist, const int seglen, const int W,
                                            const int ni, const int idx_end, uint32_t *__restrict__ out, const int tid)
{
    uint32_t dbase = (uint32_t)__cvta_generic_to_shared(S.data) + soff;
    uint32_t lbase = (uint32_t)__cvta_generic_to_shared(S.list);
    uint32_t obase = (uint32_t)__cvta_generic_to_shared(S.off);
    asm volatile("" : "+r"(dbase), "+r"(lbase), "+r"(obase));
    /* four slots per thread and round: their list, data and offset loads are in flight together (a thread's own
     * chain of loads -- slot -> position -> bytes -> bucket offset -> candidate -- is what bounds this loop) */
    for (int k0 = tid; k0 < ni; k0 += 4 * MATCH_THREADS) {
        int p[4], lo[4];
        uint32_t pw[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int k = k0 + j * MATCH_THREADS;
            p[j] = k < ni ? (int)lds_u16(lbase + 2 * k) : -1;
        }
#pragma unroll
        for (int j = 0; j < 4; j++) pw[j] = ld4s(dbase + (p[j] < 0 ? 0 : p[j]));
#pragma unroll
        for (int j = 0; j < 4; j++) lo[j] = (int)lds_u16(obase + 2 * hash4(pw[j], HASH_SIZE, 0));
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int k = k0 + j * MATCH_THREADS, i = p[j] - hist;
            if (i < 0) continue; /* no slot, or a window position: indexed, but gets no token */
            const int maxlen = seglen - i < DFL_MAX_MATCH ? seglen - i : DFL_MAX_MATCH;
            uint32_t tok = pw[j] & 0xFF;
            if (maxlen >= DFL_MIN_MATCH) {
                int best_len = 2, best_q = 0;
                for (int kk = k - 1; kk >= lo[j]; kk--) {
                    const int q = (int)lds_u16(lbase + 2 * kk);
                    if (p[j] - q > W) break; /* everything older is outside the window */
                    const uint32_t x0 = ld4s(dbase + q) ^ pw[j];
                    if (x0 & 0xFFFFFFu) continue; /* another trigram of the bucket */
                    int len = 3;
                    if (!(x0 >> 24)) {
                        len = 4;
                        while (len < maxlen) {
                            const uint32_t x = ld4s(dbase + p[j] + len) ^ ld4s(dbase + q + len);
                            if (x) { len += (__ffs(x) - 1) >> 3; break; }
                            len += 4;
                        }
                    }
                    if (len > maxlen) len = maxlen;
                    if (len > best_len) { best_len = len; best_q = q; }
                    if (best_len == maxlen) break;
                }
                if (best_len >= DFL_MIN_MATCH) tok = ((uint32_t)(p[j] - best_q) << 16) | (uint32_t)best_len;
            }
            out[i] = tok;
        }
    }
    /* frame positions from idx_end on have no whole trigram in the data: not in the list, literals */
    for (int p = (idx_end > hist ? idx_end : hist) + tid; p < hist + seglen; p += MATCH_THREADS) out[p - hist] = lds_u8(dbase + p);
}
