/* k_emit.cuh -- k_scan, k_emit (bit packer), k_final, device random generator (included by dfl_kernels.cu). */
#pragma once
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

/* copy len bytes (any alignment either side): whole 32-bit words of the destination, each assembled from the two
 * source words that cover it; the unaligned head and tail byte by byte.  Aligned 4-byte source reads may cover up to 3
 * bytes either side of [s0, s0 + len) inside the same aligned word -- never another word than the data's own. */
__device__ __forceinline__ void emit_copy(uint8_t *d0, const uint8_t *s0, long long len, int tid)
{
    long long head = (long long)((4 - ((uintptr_t)d0 & 3)) & 3);
    if (head > len) head = len;
    const long long nw = (len - head) >> 2;
    if (tid < head) d0[tid] = s0[tid];
    const uint8_t *sw = s0 + head;                     /* source of the first whole word */
    const int sh = (int)((uintptr_t)sw & 3) * 8;
    const uint32_t *sa = reinterpret_cast<const uint32_t *>(sw - ((uintptr_t)sw & 3));
    uint32_t *da = reinterpret_cast<uint32_t *>(d0 + head);
    /* the second source word is only read when the shift needs it */
    for (long long w = tid; w < nw; w += EMIT_THREADS)
        da[w] = sh ? __funnelshift_r(__ldg(sa + w), __ldg(sa + w + 1), sh) : __ldg(sa + w);
    for (long long i = head + 4 * nw + tid; i < len; i += EMIT_THREADS) d0[i] = s0[i];
}

/*
 * One CTA per segment: the bit packer the reference never wrote
 * (src/deflate_compress.c:269,:279 "TODO: write lit / len/dist pair").
 * Per round of 1024 tokens: each thread maps 4 tokens to (bits, nbits), a block
 * exclusive scan gives every token its bit offset, codes are OR-ed into a shared
 * staging buffer whose origin is congruent (mod 16 B) with the output address,
 * and whole 16-byte words are stored to HBM.
 * block_mode 4 (X.nblk set): the segment is the blocks k_huff_split decided, one after the other in the same bit
 * stream (a stored block among them is byte aligned as RFC 1951 3.2.4 says), then the one sync marker.
 */
__global__ void __launch_bounds__(EMIT_THREADS)
k_emit(const uint8_t *__restrict__ d_in, long long n, long long seg0, long long last_seg, int is_final,
       dfl_params prm, const uint32_t *__restrict__ tokcand, SegArrays A, SplitArrays X, uint8_t *__restrict__ d_out,
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
    const int nblk = X.nblk ? (int)X.nblk[sl] : 1;
    const uint32_t final_bit = (is_final && seg == last_seg) ? 1u : 0u;
    if (off + out_bytes > out_cap) {
        if (tid == 0) atomicExch(&state->error, 1u);
        return;
    }
    uint8_t *dst = d_out + off;

    if (nblk == 1 && (A.btype[sl] & 0xFF) == 0) {
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
            emit_copy(q + 5, src + o, len, tid);
        }
        if (tid == 0) {
            uint8_t *q = dst + seglen + 5 * nst;
            q[0] = (uint8_t)final_bit; q[1] = 0; q[2] = 0; q[3] = 0xFF; q[4] = 0xFF;
        }
        return;
    }

    unsigned long long gpos = 0; /* bytes of this segment already stored */
    uint32_t carry_n = 0, carry_v = 0; /* bits of the byte at dst + gpos that are already decided */
    for (int blk = 0; blk < nblk; blk++) {
        const size_t gb = (size_t)sl * DFL_SUB_MAX + blk;
        const bool split = X.nblk != nullptr;
        const int btype = split ? (int)X.btype[gb] : (int)(A.btype[sl] & 0xFF);
        const bool last_blk = blk == nblk - 1;
        if (btype == 0) {
            /* a stored block inside a split segment (never the only block: that case returned above): the pending bits
             * and the three header bits (all zero) fill one or two bytes, then LEN, NLEN and the bytes */
            const uint32_t hb = (carry_n + 3 + 7) >> 3;
            const uint32_t len = X.nbytes[gb];
            uint8_t *q = dst + gpos;
            if (tid == 0) {
                q[0] = (uint8_t)carry_v;
                if (hb == 2) q[1] = 0;
                q[hb] = (uint8_t)len; q[hb + 1] = (uint8_t)(len >> 8);
                q[hb + 2] = (uint8_t)~len; q[hb + 3] = (uint8_t)(~len >> 8);
            }
            emit_copy(q + hb + 4, d_in + seg_lo + X.pos0[gb], len, tid);
            gpos += hb + 4 + len;
            carry_n = 0; carry_v = 0;
            if (last_blk && tid == 0) {
                uint8_t *m = dst + gpos;
                m[0] = (uint8_t)final_bit; m[1] = 0; m[2] = 0; m[3] = 0xFF; m[4] = 0xFF;
            }
            continue;
        }
        __syncthreads(); /* the previous block's rounds are done with the tables */
        for (int i = tid; i < LLS; i += EMIT_THREADS) s_ll[i] = split ? X.lltab[gb * LLS + i] : A.lltab[(size_t)sl * LLS + i];
        if (tid < DS) s_d[tid] = split ? X.dtab[gb * DS + tid] : A.dtab[(size_t)sl * DS + tid];
        const uint32_t *tc = tokcand + (size_t)sl * prm.seg_bytes + (split ? X.tok0[gb] : 0u);
        const int ntok = split ? (int)X.ntok[gb] : (int)A.ntok[sl];
        const uint32_t hdr_bits = split ? X.hdr_bits[gb] : A.hdr_bits[sl];
        const uint32_t *hdr = split ? X.hdr + gb * HDR_WORDS : A.hdr + (size_t)sl * HDR_WORDS;

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
            if (last) {
                /* end of block; after the segment's last block the empty stored block that byte-aligns the segment */
                const uint32_t e = s_ll[256];
                if (tid == 0) stg_put(stg, endbit, e & 0xFFFF, (int)(e >> 16));
                endbit += e >> 16;
                if (last_blk) {
                    if (tid == 0) {
                        stg_put(stg, endbit, final_bit, 3);
                        /* LEN = 0x0000, NLEN = 0xFFFF */
                        stg_put(stg, ((endbit + 3 + 7) >> 3) * 8 + 16, 0xFFFF, 16);
                    }
                    endbit = (((endbit + 3 + 7) >> 3) + 4) * 8;
                }
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
