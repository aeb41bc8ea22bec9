/* k_inflate.cuh -- segment-parallel inflate: k_inflate, k_inf_resolve, k_inf_finish (included by dfl_kernels.cu). */
#pragma once
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
    if (err) {
        /* a corrupt segment leaves no stale word behind: everything it did not decode reads as a value, so
         * that k_inf_resolve never follows a "distance" this kernel did not validate */
        pos = __shfl_sync(FULL, pos, 0);
        if (pos < 0) pos = 0;
        for (int t = pos + lane; t < seglen; t += 32) out[t] = INF_KNOWN;
        if (lane == 0) atomicExch(&state->error, 2u);
    }
}

/* One round of pointer jumping over the cnt words of a chunk (chunk position 0 = output byte chunk_lo).
 * Four words per thread and step; groups that are already all values cost one 16-byte read. */
__device__ __forceinline__ uint32_t inf_jump(const uint32_t *w, long long p, uint32_t v, long long chunk_lo, long long q_min,
                                             const uint8_t *d_out, unsigned int *error)
{
    const long long q = p - (long long)v;
    if (q < q_min) { /* before the history the caller supplied: not a word k_inflate wrote */
        atomicExch(error, 2u);
        return INF_KNOWN;
    }
    if (q < 0) return INF_KNOWN | d_out[chunk_lo + q]; /* an earlier chunk (or the caller's history): final bytes */
    const uint32_t a = __ldcg(w + q);
    return (a & INF_KNOWN) ? a : v + a;
}

#define INF_BLK 4096 /* words per resolve block: a block whose words are all values is skipped from then on */

__global__ void __launch_bounds__(256)
k_inf_resolve(uint32_t *w, long long cnt, long long chunk_lo, long long hist_len, const uint8_t *__restrict__ d_out,
              unsigned char *blk_done, unsigned int *pending, unsigned int *error)
{
    const long long q_min = -(chunk_lo + hist_len); /* chunk position of the first byte that exists */
    const long long nblk = (cnt + INF_BLK - 1) / INF_BLK;
    for (long long blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
        if (blk_done[blk]) continue; /* uniform for the CTA */
        unsigned int left = 0;
        const long long p0 = blk * INF_BLK;
        const long long pend = p0 + INF_BLK < cnt ? p0 + INF_BLK : cnt;
        const long long quads = (pend - p0) >> 2;
        for (long long g = threadIdx.x; g < quads; g += blockDim.x) {
            uint4 *q4 = reinterpret_cast<uint4 *>(w + p0) + g;
            uint4 v = __ldcg(q4);
            if (v.x & v.y & v.z & v.w & INF_KNOWN) continue;
            const long long p = p0 + (g << 2);
            if (!(v.x & INF_KNOWN)) v.x = inf_jump(w, p, v.x, chunk_lo, q_min, d_out, error);
            if (!(v.y & INF_KNOWN)) v.y = inf_jump(w, p + 1, v.y, chunk_lo, q_min, d_out, error);
            if (!(v.z & INF_KNOWN)) v.z = inf_jump(w, p + 2, v.z, chunk_lo, q_min, d_out, error);
            if (!(v.w & INF_KNOWN)) v.w = inf_jump(w, p + 3, v.w, chunk_lo, q_min, d_out, error);
            *q4 = v;
            left |= ~(v.x & v.y & v.z & v.w) & INF_KNOWN;
        }
        if (threadIdx.x < ((pend - p0) & 3)) { /* tail of the last block */
            const long long p = p0 + (quads << 2) + threadIdx.x;
            const uint32_t v = __ldcg(w + p);
            if (!(v & INF_KNOWN)) {
                const uint32_t nv = inf_jump(w, p, v, chunk_lo, q_min, d_out, error);
                w[p] = nv;
                left |= ~nv & INF_KNOWN;
            }
        }
        const int any = __syncthreads_or((int)(left != 0));
        if (threadIdx.x == 0) {
            if (any) atomicOr(pending, 1u);
            else blk_done[blk] = 1;
        }
    }
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
