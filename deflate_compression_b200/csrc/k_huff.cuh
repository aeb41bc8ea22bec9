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
        A.cost[2 * (size_t)sl] = btype == 2 ? (uint32_t)(tree + 5 + 5 + 4 + 3 * hclen) : 0u;
        A.cost[2 * (size_t)sl + 1] = btype == 2 ? d_dyn : btype == 1 ? d_fix : 0u;
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
