/* k_huff.cuh -- Huffman construction, block header, block-type choice (included by dfl_kernels.cu). */
#pragma once
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

/* what huff_plan decided for one block (uniform across the warp) */
struct HuffPlan {
    int btype;                       /* 0 stored, 1 fixed, 2 dynamic */
    unsigned long long block_bits;   /* bits of the chosen block incl. its 3 header bits (stored: 8 * (bytes + 5 per 65535)) */
    uint32_t hdr_bits;               /* bits packed into S.hdr (0 for stored) */
    uint32_t tree_bits, d_bits;      /* cost split of the chosen block: code description, distance symbols + extra bits */
};

/*
 * Code construction + exact block costs + choice for ONE block whose lit/len and distance counts (EOB included) are in
 * S.ll_freq / S.d_freq, by one warp.  Leaves the chosen block's packed header in S.hdr and, for a dynamic block, its
 * codes in S.ll_code / S.d_code (unless cost_only).  Twin of orc_plan_block (oracle/deflate_oracle.c); what the reference accounts per
 * token in h_tree_d_lens / h_tree_builder_* (src/h_tree.c:75-148, :242-302) is done once per block here.
 */
__device__ HuffPlan huff_plan(HuffSmem &S, const int lane, const unsigned long long extra, const long long nbytes, const int block_mode,
                              const bool cost_only = false)
{
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

    /* exact costs; the distance part (symbols + extra bits, RFC 1951 3.2.5: dist symbol s >= 4 carries s/2 - 1 extra
     * bits) is kept apart for the cost split of struct compress_stats (src/include/deflate_ext.h:19-31) */
    unsigned long long dyn = 0, fix = 0;
    uint32_t d_dyn = 0, d_fix = 0;
    for (int s = lane; s < NLL; s += 32) {
        dyn += (unsigned long long)S.ll_freq[s] * S.ll_len[s];
        fix += (unsigned long long)S.ll_freq[s] * fixed_ll_len(s);
    }
    if (lane < ND) {
        const uint32_t dx = lane >= 4 ? (uint32_t)(lane / 2 - 1) : 0u;
        d_dyn = S.d_freq[lane] * (S.d_len[lane] + dx);
        d_fix = S.d_freq[lane] * (5 + dx);
        dyn += (unsigned long long)S.d_freq[lane] * S.d_len[lane];
        fix += (unsigned long long)S.d_freq[lane] * 5;
    }
    d_dyn = __reduce_add_sync(FULL, d_dyn);
    d_fix = __reduce_add_sync(FULL, d_fix);
    unsigned long long tree = 0; /* code-length code part of the dynamic header */
    if (lane < NCL) {
        tree = (unsigned long long)S.cl_freq[lane] * S.cl_len[lane];
        if (lane == 16) tree += 2ull * S.cl_freq[16];
        if (lane == 17) tree += 3ull * S.cl_freq[17];
        if (lane == 18) tree += 7ull * S.cl_freq[18];
        dyn += tree;
    }
    for (int o = 16; o; o >>= 1) {
        dyn += __shfl_xor_sync(FULL, dyn, o);
        fix += __shfl_xor_sync(FULL, fix, o);
        tree += __shfl_xor_sync(FULL, tree, o);
    }
    HuffPlan P;
    P.btype = 0; P.block_bits = 0; P.hdr_bits = 0; P.tree_bits = 0; P.d_bits = 0;
    if (lane == 0) {
        int btype;
        for (hclen = 19; hclen > 4 && S.cl_len[c_cl_order[hclen - 1]] == 0; hclen--) {}
        dyn += 3 + 5 + 5 + 4 + 3 * hclen + extra;
        fix += 3 + extra;
        unsigned long long nstored = (unsigned long long)((nbytes + 65534) / 65535);
        if (nstored < 1) nstored = 1;
        unsigned long long stored = 8ull * ((unsigned long long)nbytes + 5 * nstored);
        if (block_mode == 1) btype = 0;
        else if (block_mode == 2) btype = 1;
        else if (block_mode == 3) btype = 2;
        else if (stored <= fix && stored <= dyn) btype = 0;
        else if (fix <= dyn) btype = 1;
        else btype = 2;
        P.btype = btype;
        P.block_bits = btype == 0 ? stored : btype == 1 ? fix : dyn;

        BitW bw{S.hdr, 0};
        if (cost_only) {
            /* only the decision is wanted (k_huff_split weighing a split): no codes, no header */
        } else if (btype == 2) {
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
        P.hdr_bits = bw.nbits;
        P.tree_bits = btype == 2 ? (uint32_t)(tree + 5 + 5 + 4 + 3 * hclen) : 0u;
        P.d_bits = btype == 2 ? d_dyn : btype == 1 ? d_fix : 0u;
    }
    P.btype = __shfl_sync(FULL, P.btype, 0);
    P.block_bits = __shfl_sync(FULL, P.block_bits, 0);
    P.hdr_bits = __shfl_sync(FULL, P.hdr_bits, 0);
    P.tree_bits = __shfl_sync(FULL, P.tree_bits, 0);
    P.d_bits = __shfl_sync(FULL, P.d_bits, 0);
    __syncwarp();
    return P;
}

/* the planned block's code tables (code_rev | len << 16) and packed header, to the arrays k_emit reads */
__device__ void huff_store(const HuffSmem &S, const int lane, const int btype, uint32_t *lltab, uint32_t *dtab, uint32_t *hdr)
{
    if (btype == 2) {
        for (int s = lane; s < LLS; s += 32)
            lltab[s] = s < NLL ? (uint32_t)S.ll_code[s] | ((uint32_t)S.ll_len[s] << 16) : 0;
        dtab[lane] = lane < ND ? (uint32_t)S.d_code[lane] | ((uint32_t)S.d_len[lane] << 16) : 0;
        for (int i = lane; i < HDR_WORDS; i += 32) hdr[i] = S.hdr[i];
    } else if (btype == 1) {
        /* fixed code of RFC 1951 3.2.6 */
        for (int s = lane; s < LLS; s += 32) {
            int l = fixed_ll_len(s);
            uint32_t code = s < 144 ? 0x30 + s : s < 256 ? 0x190 + (s - 144) : s < 280 ? (s - 256) : 0xC0 + (s - 280);
            lltab[s] = (__brev(code) >> (32 - l)) | ((uint32_t)l << 16);
        }
        dtab[lane] = (__brev((uint32_t)lane) >> 27) | (5u << 16);
        if (lane == 0) hdr[0] = S.hdr[0];
    }
}

/* One warp per segment, one block per segment (block_mode 0..3). */
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
    __syncwarp();
    const HuffPlan P = huff_plan(S, lane, A.extra[sl], seglen, prm.block_mode);
    if (lane == 0) {
        unsigned long long nstored = (unsigned long long)((seglen + 65534) / 65535);
        if (nstored < 1) nstored = 1;
        A.hdr_bits[sl] = P.hdr_bits;
        A.cost[2 * (size_t)sl] = P.tree_bits;
        A.cost[2 * (size_t)sl + 1] = P.d_bits;
        A.btype[sl] = P.btype;
        A.block_bits[sl] = P.block_bits;
        A.out_bytes[sl] = P.btype == 0 ? (uint32_t)(seglen + 5 * nstored + 5)
                                       : (uint32_t)((P.block_bits + 3 + 7) / 8 + 4);
    }
    huff_store(S, lane, P.btype, A.lltab + (size_t)sl * LLS, A.dtab + (size_t)sl * DS, A.hdr + (size_t)sl * HDR_WORDS);
}

/* ------------------------------------------------------------------------- */
/* k_huff_split: adaptive block splitting (block_mode 4)                     */
/* ------------------------------------------------------------------------- */

#define SPLIT_WARPS 2

/* k_huff_split weighs and builds blocks at four places: one copy of the planner, called (inlined four times it is
 * 62 K SASS lines and the kernel spends its time fetching instructions: 9.6 -> 6 ms on data that splits everywhere) */
__device__ __noinline__ HuffPlan huff_plan_call(HuffSmem &S, const int lane, const unsigned long long extra, const long long nbytes,
                                                const int block_mode, const bool cost_only)
{
    return huff_plan(S, lane, extra, nbytes, block_mode, cost_only);
}

/* per-block arrays of a chunk in block_mode 4 (block b of segment sl at index sl * DFL_SUB_MAX + b) */
struct SplitArrays {
    uint32_t *nblk;                 /* [nseg] blocks of the segment; NULL = not splitting */
    uint32_t *btype, *hdr_bits;     /* [nseg * DFL_SUB_MAX] */
    uint32_t *tok0, *ntok;          /* the block's tokens: first index in the segment's token array, count */
    uint32_t *pos0, *nbytes;        /* the input bytes it covers (stored blocks copy them) */
    uint32_t *lltab, *dtab, *hdr;   /* [nseg * DFL_SUB_MAX][LLS | DS | HDR_WORDS] */
};

struct SplitSmem {
    HuffSmem H;
    uint32_t sub_ll[DFL_SUB_MAX][LLS], sub_d[DFL_SUB_MAX][DS];
    uint32_t sub_extra[DFL_SUB_MAX];
    uint32_t tok_start[DFL_SUB_MAX + 1], pos_start[DFL_SUB_MAX + 1];
};

/*
 * One warp per segment.  Adaptive block splitting, the reference's stated aim (README.md:15-20: "how to divide the data
 * up into chunks, each able to have its own Huffman coding"; its running-cost machinery, src/aht.c:239-277, never
 * reaches a decision).  Twin of orc_emit_segment / orc_split_range (oracle/deflate_oracle.c): the segment's tokens are
 * cut where they cross multiples of DFL_SUB_BYTES input bytes; the warp histograms the sub-blocks in shared memory,
 * then bisects a range of sub-blocks whenever its two halves, each as its own cheapest block (exact costs, huff_plan),
 * take strictly fewer bits than the range as one block.  The token stream is untouched.
 */
__global__ void __launch_bounds__(SPLIT_WARPS * 32)
k_huff_split(long long n, long long seg0, int nseg, dfl_params prm, const uint32_t *__restrict__ tokens, SegArrays A, SplitArrays X)
{
    __shared__ SplitSmem s_all[SPLIT_WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sl = blockIdx.x * SPLIT_WARPS + warp;
    if (sl >= nseg) return;
    SplitSmem &S = s_all[warp];
    const long long seg_lo = (seg0 + sl) * prm.seg_bytes;
    const int seglen = (int)(((seg_lo + prm.seg_bytes < n) ? seg_lo + prm.seg_bytes : n) - seg_lo);
    const int ntok = (int)A.ntok[sl];
    const uint32_t *tc = tokens + (size_t)sl * prm.seg_bytes;
    const int nsub = seglen > DFL_SUB_BYTES ? (seglen + DFL_SUB_BYTES - 1) / DFL_SUB_BYTES : 1;

    /* ---- sub-block histograms, first token and first byte of every sub-block ---- */
    for (int i = lane; i < DFL_SUB_MAX * LLS; i += 32) (&S.sub_ll[0][0])[i] = 0;
    for (int i = lane; i < DFL_SUB_MAX * DS; i += 32) (&S.sub_d[0][0])[i] = 0;
    if (lane < DFL_SUB_MAX) S.sub_extra[lane] = 0;
    if (lane <= DFL_SUB_MAX) { S.tok_start[lane] = lane ? (uint32_t)ntok : 0u; S.pos_start[lane] = lane ? (uint32_t)seglen : 0u; }
    __syncwarp();
    int pos = 0, prev_sub = 0;
    for (int base = 0; base < ntok; base += 32) {
        const int k = base + lane;
        const uint32_t t = k < ntok ? tc[k] : 0u;
        const int adv = k < ntok ? ((t >> 16) ? (int)(t & 0xFFFF) : 1) : 0;
        int x = adv;
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(FULL, x, o);
            if (lane >= o) x += y;
        }
        const int mypos = pos + x - adv;
        const int sub = nsub > 1 ? mypos / DFL_SUB_BYTES : 0;
        int before = __shfl_up_sync(FULL, sub, 1);
        if (lane == 0) before = prev_sub;
        if (k < ntok) {
            if (sub != before) { S.tok_start[sub] = (uint32_t)k; S.pos_start[sub] = (uint32_t)mypos; } /* (a token is shorter than a sub-block) */
            if ((t >> 16) == 0) {
                atomicAdd(&S.sub_ll[sub][t & 0xFF], 1u);
            } else {
                int eb, ev, e2;
                atomicAdd(&S.sub_ll[sub][len_sym((int)(t & 0xFFFF), eb, ev)], 1u);
                atomicAdd(&S.sub_d[sub][dist_sym((int)(t >> 16), e2, ev)], 1u);
                atomicAdd(&S.sub_extra[sub], (uint32_t)(eb + e2));
            }
        }
        pos += __shfl_sync(FULL, x, 31);
        const int nlast = ntok - base < 32 ? ntok - base : 32;
        prev_sub = __shfl_sync(FULL, sub, nlast - 1);
    }
    __syncwarp();

    /* a block over sub-blocks [a, b): counts summed into the plan's tables, one EOB */
    auto plan_range = [&](int a, int b, bool cost_only) -> HuffPlan {
        for (int s = lane; s < LLS; s += 32) {
            uint32_t f = s == 256 ? 1u : 0u;
            for (int j = a; j < b; j++) f += S.sub_ll[j][s];
            S.H.ll_freq[s] = s < NLL ? f : 0u;
        }
        uint32_t fd = 0, ex = lane >= a && lane < b ? S.sub_extra[lane] : 0u;
        for (int j = a; j < b; j++) fd += S.sub_d[j][lane];
        S.H.d_freq[lane] = lane < ND ? fd : 0u;
        ex = __reduce_add_sync(FULL, ex);
        __syncwarp();
        return huff_plan_call(S.H, lane, ex, (long long)S.pos_start[b] - (long long)S.pos_start[a], prm.block_mode == 4 ? 0 : prm.block_mode, cost_only);
    };

    /* ---- top-down bisection (explicit stack: ranges still to be looked at, leftmost on top; a range that came out of
     * a split brings its cost along) ---- */
    int st_a[4], st_b[4], sp = 0;
    unsigned long long st_c[4];
    st_a[0] = 0; st_b[0] = nsub; st_c[0] = 0; sp = 1;
    uint32_t nblk = 0, tree_sum = 0, d_sum = 0, first_type = 0;
    unsigned long long bits_sum = 0, run_bits = 0; /* costs as planned; exact bits of the segment's stream so far */
    while (sp > 0) {
        sp--;
        const int a = st_a[sp], b = st_b[sp];
        HuffPlan P;
        bool have = false; /* the plan's tables are this range's, in full */
        if (b - a >= 2) {
            const int mid = (a + b) / 2;
            const HuffPlan L = plan_range(a, mid, true), R = plan_range(mid, b, true);
            unsigned long long w = st_c[sp];
            if (!w) { P = plan_range(a, b, false); w = P.block_bits; have = true; } /* the whole segment: planned in full, it is the likely outcome */
            if (L.block_bits + R.block_bits < w) {
                st_a[sp] = mid; st_b[sp] = b; st_c[sp] = R.block_bits; sp++;
                st_a[sp] = a; st_b[sp] = mid; st_c[sp] = L.block_bits; sp++;
                continue;
            }
        }
        if (!have) P = plan_range(a, b, false); /* the block: codes, header */
        const size_t gb = (size_t)sl * DFL_SUB_MAX + nblk;
        const uint32_t nb = S.pos_start[b] - S.pos_start[a];
        if (lane == 0) {
            X.btype[gb] = (uint32_t)P.btype;
            X.hdr_bits[gb] = P.hdr_bits;
            X.tok0[gb] = S.tok_start[a];
            X.ntok[gb] = S.tok_start[b] - S.tok_start[a];
            X.pos0[gb] = S.pos_start[a];
            X.nbytes[gb] = nb;
        }
        huff_store(S.H, lane, P.btype, X.lltab + gb * LLS, X.dtab + gb * DS, X.hdr + gb * HDR_WORDS);
        if (nblk == 0) first_type = (uint32_t)P.btype;
        if (P.btype == 0) run_bits = ((run_bits + 3 + 7) & ~7ull) + 32 + 8ull * nb; /* header bits, padding, LEN / NLEN, the bytes */
        else run_bits += P.block_bits;
        bits_sum += P.block_bits;
        tree_sum += P.tree_bits;
        d_sum += P.d_bits;
        nblk++;
        __syncwarp();
    }
    if (lane == 0) {
        X.nblk[sl] = nblk;
        A.btype[sl] = first_type | (nblk << 8); /* (the host reads the block count from the upper bits) */
        A.hdr_bits[sl] = 0;
        A.cost[2 * (size_t)sl] = tree_sum;
        A.cost[2 * (size_t)sl + 1] = d_sum;
        A.block_bits[sl] = bits_sum;
        A.out_bytes[sl] = (uint32_t)(((run_bits + 3 + 7) >> 3) + 4); /* + the sync marker */
    }
}
