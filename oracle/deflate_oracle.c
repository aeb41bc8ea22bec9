/*
 * oracle/deflate_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see oracle.h).
 *
 * Plain-C restatement of the reference compress path
 * (mparker97/deflate_compression, src/deflate_compress.c + src/h_tree.c) plus the
 * encode side the reference lacks (RFC 1951 / RFC 1950).  Every function cites
 * the reference lines it follows or the RFC clause that pins it.
 *
 * Token stream: PINNED against the real reference (oracle/_ref, tests/golden).
 * Bitstream:    parity unpinned by the reference (it emits no bytes); pinned by
 *               RFC 1951/1950 and zlib inflate.
 */
#include "oracle.h"

#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------- */
/* parameters                                                                */
/* ------------------------------------------------------------------------- */

void orc_default_params(orc_params *p)
{
    p->window = 32768;      /* tests/check_lld.c:18, README.md:6 */
    p->seg_bytes = 32768;   /* largest input the reference survives */
    p->chain_limit = 0;     /* reference walks every chain entry: src/deflate_compress.c:249 */
    p->use_dict = 1;
    p->block_mode = 0;
}

/* Bucket hash of the 3 bytes at p.  The reference's dup_hash
 * (src/deflate_compress.c:115-135) is a 10-bit bit-interleave; with an
 * unlimited chain the hash is unobservable ("key design
 * observation"), so the restatement uses a wider multiplicative hash.  It only
 * becomes observable with chain_limit > 0, where the CUDA path uses the same
 * function (include/deflate_b200.h DFL_HASH3). */
uint32_t orc_hash3(const uint8_t *p)
{
    uint32_t v = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16);
    return (uint32_t)(((uint64_t)(v * 0x9E3779B1u) * ORC_HASH_BUCKETS) >> 32);
}

/* Bucket of the 4 bytes (b3 = the fourth byte, 0 past the end of the segment): the CUDA path's
 * index (include/deflate_b200.h DFL_HASH4) -- the trigram picks a group of 2^ORC_SUB_BITS
 * neighbouring buckets, the fourth byte one of them.  Observable only with chain_limit > 0. */
uint32_t orc_hash4(const uint8_t *p, uint32_t b3)
{
    uint32_t v = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16);
    uint32_t grp = (uint32_t)(((uint64_t)(v * 0x9E3779B1u) * (ORC_HASH_BUCKETS >> ORC_SUB_BITS)) >> 32);
    return (grp << ORC_SUB_BITS) | ((b3 * 0x9E3779B1u) >> (32 - ORC_SUB_BITS));
}

/* ------------------------------------------------------------------------- */
/* LZ77: hash-chain longest match + greedy parse                             */
/* ------------------------------------------------------------------------- */

/* Common-prefix length, capped.  Follows check_dup_str
 * (src/deflate_compress.c:164-180): byte-wise compare, stop at MAXLEN or at the
 * end of the available bytes; overlap (dist < len) is allowed. */
static int orc_lcp(const uint8_t *a, const uint8_t *b, int cap)
{
    int n = 0;
    while (n < cap && a[n] == b[n]) n++;
    return n;
}

/*
 * Greedy parse of one segment.  Follows process_loop
 * (src/deflate_compress.c:243-338):
 *   - candidates are walked newest -> oldest (:249-263), every earlier position
 *     of the bucket that is within the window;
 *   - the longest wins, the first found (= nearest) on ties (strict '>' at :258);
 *   - max_len starts at 2 so only len >= 3 is a match (:247,:266);
 *   - len >= 3 -> (len, dist) and advance len, else literal and advance 1
 *     (:266-288);
 *   - every covered position enters the index (:313-318).
 * Whole-buffer semantics replace the reference's broken two-window bookkeeping
 * (SURVEY.md F4, Appendix A): history is the preceding <= window bytes, matches
 * are not truncated at window edges, and positions with < 3 bytes left are
 * neither matched nor hashed.
 */
long orc_lz77_segment(const uint8_t *buf, long hist_lo, long seg_lo, long seg_hi,
                      const orc_params *prm, uint32_t *tokens)
{
    const long W = prm->window;
    const int K = prm->chain_limit;
    long p, ntok = 0, span;
    int32_t *head, *prev, *head4, *prev4;
    long i;

    if (!prm->use_dict) hist_lo = seg_lo;
    if (hist_lo < seg_lo - W) hist_lo = seg_lo - W;
    if (hist_lo < 0) hist_lo = 0;
    span = seg_hi - hist_lo;
    if (span <= 0) return 0;

    head = (int32_t *)malloc(sizeof(int32_t) * ORC_HASH_BUCKETS);
    prev = (int32_t *)malloc(sizeof(int32_t) * (size_t)span);
    head4 = (int32_t *)malloc(sizeof(int32_t) * ORC_HASH_BUCKETS);
    prev4 = (int32_t *)malloc(sizeof(int32_t) * (size_t)span);
    if (!head || !prev || !head4 || !prev4) { free(head); free(prev); free(head4); free(prev4); return -1; }
    for (i = 0; i < ORC_HASH_BUCKETS; i++) head[i] = head4[i] = -1;

    /* every position with 3 bytes left enters the trigram chain (and, for chain_limit > 0, the
     * chain of its 4-byte bucket) */
#define ORC_INSERT(pos)                                                                        \
    do {                                                                                       \
        if ((pos) + 3 <= seg_hi) {                                                             \
            uint32_t h_ = orc_hash3(buf + (pos));                                              \
            prev[(pos) - hist_lo] = head[h_];                                                  \
            head[h_] = (int32_t)((pos) - hist_lo);                                             \
            if (K) {                                                                           \
                h_ = orc_hash4(buf + (pos), (pos) + 3 < seg_hi ? buf[(pos) + 3] : 0);          \
                prev4[(pos) - hist_lo] = head4[h_];                                            \
                head4[h_] = (int32_t)((pos) - hist_lo);                                        \
            }                                                                                  \
        }                                                                                      \
    } while (0)

    for (p = hist_lo; p < seg_lo; p++) ORC_INSERT(p); /* dictionary priming: indexed, no tokens */

    p = seg_lo;
    while (p < seg_hi) {
        int best_len = 2, best_dist = 0, adv;
        int maxlen = (seg_hi - p) < ORC_MAX_MATCH ? (int)(seg_hi - p) : ORC_MAX_MATCH;
        if (maxlen >= ORC_MIN_MATCH && !K) {
            /* the reference's search: the whole chain */
            int32_t q = head[orc_hash3(buf + p)];
            while (q >= 0) {
                long dist = (p - hist_lo) - q;
                int len;
                if (dist > W) break;
                len = orc_lcp(buf + hist_lo + q, buf + p, maxlen);
                if (len > best_len) {
                    best_len = len;
                    best_dist = (int)dist;
                    if (len == maxlen) break; /* nothing nearer or longer can follow */
                }
                q = prev[q];
            }
        } else if (maxlen >= ORC_MIN_MATCH) {
            /* chain_limit K (not in the reference; the semantics are the CUDA path's):
             * (1) the newest K entries of the 4-byte bucket, looking for matches of >= 4 bytes;
             * (2) if there is none, the nearest earlier occurrence of the trigram in the window. */
            if (maxlen > ORC_MIN_MATCH) {
                int32_t q = head4[orc_hash4(buf + p, buf[p + 3])];
                int hops = 0;
                while (q >= 0) {
                    long dist = (p - hist_lo) - q;
                    int len;
                    if (dist > W || hops >= K) break;
                    hops++;
                    len = orc_lcp(buf + hist_lo + q, buf + p, maxlen);
                    if (len > ORC_MIN_MATCH && len > best_len) {
                        best_len = len;
                        best_dist = (int)dist;
                        if (len == maxlen) break;
                    }
                    q = prev4[q];
                }
            }
            if (best_len == 2) {
                int32_t q = head[orc_hash3(buf + p)];
                while (q >= 0) {
                    long dist = (p - hist_lo) - q;
                    if (dist > W) break;
                    if (orc_lcp(buf + hist_lo + q, buf + p, ORC_MIN_MATCH) == ORC_MIN_MATCH) {
                        best_len = ORC_MIN_MATCH;
                        best_dist = (int)dist;
                        break;
                    }
                    q = prev[q];
                }
            }
        }
        if (best_len < ORC_MIN_MATCH) {
            tokens[ntok++] = buf[p];
            adv = 1;
        } else {
            tokens[ntok++] = ((uint32_t)best_dist << 16) | (uint32_t)best_len;
            adv = best_len;
        }
        for (i = 0; i < adv; i++) ORC_INSERT(p + i);
        p += adv;
    }
#undef ORC_INSERT
    free(head);
    free(prev);
    free(head4);
    free(prev4);
    return ntok;
}

/* check_lld method (tests/check_lld.c:20-39): re-expand tokens. `hist` holds the
 * hist_len bytes preceding the segment. */
long orc_tokens_expand(const uint32_t *tokens, long ntok, const uint8_t *hist, long hist_len,
                       uint8_t *out, long cap)
{
    long n = 0, k;
    for (k = 0; k < ntok; k++) {
        uint32_t t = tokens[k];
        uint32_t d = ORC_TOK_DIST(t), l = ORC_TOK_LEN(t);
        if (d == 0) {
            if (n >= cap) return -1;
            out[n++] = (uint8_t)l;
        } else {
            uint32_t i;
            if (n + (long)l > cap) return -1;
            if ((long)d > n + hist_len) return -2;
            for (i = 0; i < l; i++, n++) {
                long src = n - (long)d;
                out[n] = src >= 0 ? out[src] : hist[hist_len + src];
            }
        }
    }
    return n;
}

/* ------------------------------------------------------------------------- */
/* RFC 1951 section 3.2.5 symbol mapping                                     */
/* ------------------------------------------------------------------------- */

static int ilog2_u32(uint32_t v) { return 31 - __builtin_clz(v); }

/* length 3..258 -> code 257..285; same code numbers as get_len_code
 * (src/deflate_compress.c:182-200).  Extra-bit count/value follow RFC 1951
 * 3.2.5 (the reference's extra-bit outputs are off by one). */
static int orc_len_sym(int len, int *ebits, int *eval)
{
    int t = len - 3, k, eb;
    if (len == 258) { *ebits = 0; *eval = 0; return 285; }
    if (t < 8) { *ebits = 0; *eval = 0; return 257 + t; }
    k = ilog2_u32((uint32_t)t);
    eb = k - 2;
    *ebits = eb;
    *eval = t & ((1 << eb) - 1);
    return 261 + 4 * eb + ((t >> eb) & 3);
}

/* distance 1..32768 -> code 0..29; same code numbers as get_dist_code
 * (src/deflate_compress.c:202-217). */
static int orc_dist_sym(int dist, int *ebits, int *eval)
{
    int t = dist - 1, k, eb;
    if (t < 4) { *ebits = 0; *eval = 0; return t; }
    k = ilog2_u32((uint32_t)t);
    eb = k - 1;
    *ebits = eb;
    *eval = t & ((1 << eb) - 1);
    return 2 * k + ((t >> eb) & 1);
}

/* ------------------------------------------------------------------------- */
/* Huffman construction                                                      */
/* ------------------------------------------------------------------------- */

/*
 * Length-limited code lengths.  The reference only has the two-queue builder
 * for the 19-symbol alphabet (h_tree_builder_build, src/h_tree.c:242-280: sort
 * by (weight, value) :182-189, then repeatedly join the two lightest of
 * {leaf queue head, node queue head}) and no length limiting.
 * This restatement keeps that construction and adds what DEFLATE needs:
 *   1. used symbols only; an alphabet with < 2 used symbols is padded with the
 *      lowest-numbered unused symbols at weight 1 (single-leaf case is a TODO in
 *      the reference, src/h_tree.c:257-259);
 *   2. sort by (weight asc, symbol asc);
 *   3. two-queue merge, ties -> leaf queue first;
 *   4. count leaves per depth, clip at `limit`, repair the Kraft sum one unit of
 *      2^-limit at a time by moving a leaf of the deepest shorter level down;
 *   5. hand lengths out along the sorted order, longest to the lightest.
 * Deterministic: the CUDA path must reproduce it bit for bit.
 */
void orc_huff_lengths(const uint32_t *freq, int n, int limit, uint8_t *len)
{
    int sym[ORC_NUM_LL + 2];
    uint32_t w[ORC_NUM_LL + 2];
    uint32_t iw[ORC_NUM_LL + 2];
    int lpar[ORC_NUM_LL + 2], ipar[ORC_NUM_LL + 2], idepth[ORC_NUM_LL + 2];
    int bl_count[64];
    int m = 0, s, i, j, li, ih, it, bits;
    unsigned long kraft;

    for (s = 0; s < n; s++) {
        len[s] = 0;
        if (freq[s]) { sym[m] = s; w[m] = freq[s]; m++; }
    }
    for (s = 0; m < 2 && s < n; s++) {
        if (!freq[s]) { sym[m] = s; w[m] = 1; m++; }
    }
    /* insertion sort by (w, sym) */
    for (i = 1; i < m; i++) {
        int cs = sym[i];
        uint32_t cw = w[i];
        for (j = i - 1; j >= 0 && (w[j] > cw || (w[j] == cw && sym[j] > cs)); j--) {
            w[j + 1] = w[j];
            sym[j + 1] = sym[j];
        }
        w[j + 1] = cw;
        sym[j + 1] = cs;
    }
    /* two-queue merge */
    li = 0; ih = 0; it = 0;
    while (it < m - 1) {
        uint32_t sum = 0;
        int k;
        for (k = 0; k < 2; k++) {
            if (li < m && (ih >= it || w[li] <= iw[ih])) { sum += w[li]; lpar[li] = it; li++; }
            else { sum += iw[ih]; ipar[ih] = it; ih++; }
        }
        iw[it++] = sum;
    }
    /* depths: root is node m-2 */
    idepth[m - 2] = 0;
    for (i = m - 3; i >= 0; i--) idepth[i] = idepth[ipar[i]] + 1;
    memset(bl_count, 0, sizeof(bl_count));
    for (i = 0; i < m; i++) {
        int d = idepth[lpar[i]] + 1;
        bl_count[d > limit ? limit : d]++;
    }
    /* Kraft repair in units of 2^-limit: while the clipped code is over-subscribed, drop one
     * code of the longest length a level... i.e. take one leaf from the deepest shorter level,
     * turn it into two leaves one level down, and remove one leaf slot at `limit` (net: same
     * number of codes, Kraft sum minus one unit). */
    kraft = 0;
    for (bits = 1; bits <= limit; bits++) kraft += (unsigned long)bl_count[bits] << (limit - bits);
    while (kraft > (1ul << limit)) {
        bl_count[limit]--;
        for (bits = limit - 1; bits > 0; bits--) {
            if (bl_count[bits]) {
                bl_count[bits]--;
                bl_count[bits + 1] += 2;
                break;
            }
        }
        kraft--;
    }
    j = 0;
    for (bits = limit; bits >= 1; bits--)
        for (i = 0; i < bl_count[bits]; i++) len[sym[j++]] = (uint8_t)bits;
}

static uint16_t orc_bitrev(uint16_t c, int n)
{
    uint16_t r = 0;
    int i;
    for (i = 0; i < n; i++) { r = (uint16_t)((r << 1) | (c & 1)); c >>= 1; }
    return r;
}

/* Canonical code assignment, RFC 1951 3.2.2 (docs/rfc1951.txt:372); the
 * reference has the decode-side twin in form_h_tree
 * (src/deflate_decompress.c:146-161).  Codes are returned bit-reversed, ready
 * for the LSB-first packer. */
void orc_canonical_codes(const uint8_t *len, int n, uint16_t *code_rev)
{
    int bl_count[16] = {0}, next_code[16], bits, s, code = 0;
    for (s = 0; s < n; s++) bl_count[len[s]]++;
    bl_count[0] = 0;
    for (bits = 1; bits <= 15; bits++) {
        code = (code + bl_count[bits - 1]) << 1;
        next_code[bits] = code;
    }
    for (s = 0; s < n; s++) {
        code_rev[s] = 0;
        if (len[s]) code_rev[s] = orc_bitrev((uint16_t)next_code[len[s]]++, len[s]);
    }
}

/* ------------------------------------------------------------------------- */
/* bit writer (RFC 1951 3.1.1: LSB first)                                    */
/* ------------------------------------------------------------------------- */

typedef struct {
    uint8_t *out;
    long cap, pos;
    uint64_t acc;
    int nacc;
    int err;
} orc_bw;

static void bw_put(orc_bw *b, uint32_t v, int n)
{
    b->acc |= (uint64_t)v << b->nacc;
    b->nacc += n;
    while (b->nacc >= 8) {
        if (b->pos >= b->cap) { b->err = 1; b->nacc = 0; return; }
        b->out[b->pos++] = (uint8_t)b->acc;
        b->acc >>= 8;
        b->nacc -= 8;
    }
}

static void bw_align(orc_bw *b)
{
    if (b->nacc) bw_put(b, 0, 8 - b->nacc);
}

static void bw_bytes(orc_bw *b, const uint8_t *src, long n)
{
    if (b->pos + n > b->cap) { b->err = 1; return; }
    memcpy(b->out + b->pos, src, (size_t)n);
    b->pos += n;
}

/* ------------------------------------------------------------------------- */
/* block encoder                                                             */
/* ------------------------------------------------------------------------- */

static const uint8_t orc_cl_order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
/* same permutation as D1_INIT_CODE_LENS, src/deflate_decompress.c:23 */

static int orc_fixed_ll_len(int s) { return s < 144 ? 8 : s < 256 ? 9 : s < 280 ? 7 : 8; }

typedef struct {
    uint8_t sym;   /* 0..18 */
    uint8_t extra; /* value of the extra bits */
} orc_cl_item;

/*
 * Run-length code the HLIT+HDIST code lengths as ONE sequence (RFC 1951 3.2.7,
 * docs/rfc1951.txt:690; runs may cross from the literal/length lengths into
 * the distance lengths, as h_tree_d_lens does at src/h_tree.c:90-97).
 * Behaviourally what h_tree_d_lens counts (src/h_tree.c:75-148), minus its two
 * defects (zero run of 2 costed as a 17, :111-113; depth > 15, :120):
 *   zero run : 18 (11..138) while run >= 11, then 17 (3..10) if run >= 3, then 0s;
 *   other    : the length once, then 16 (3..6) while run >= 3, then the length.
 */
static int orc_rle_lengths(const uint8_t *seq, int n, orc_cl_item *items)
{
    int i = 0, k = 0;
    while (i < n) {
        int v = seq[i], j = i + 1, run;
        while (j < n && seq[j] == v) j++;
        run = j - i;
        if (v == 0) {
            while (run >= 11) {
                int r = run < 138 ? run : 138;
                items[k].sym = 18; items[k].extra = (uint8_t)(r - 11); k++;
                run -= r;
            }
            if (run >= 3) { items[k].sym = 17; items[k].extra = (uint8_t)(run - 3); k++; run = 0; }
            while (run-- > 0) { items[k].sym = 0; items[k].extra = 0; k++; }
        } else {
            items[k].sym = (uint8_t)v; items[k].extra = 0; k++;
            run--;
            while (run >= 3) {
                int r = run < 6 ? run : 6;
                items[k].sym = 16; items[k].extra = (uint8_t)(r - 3); k++;
                run -= r;
            }
            while (run-- > 0) { items[k].sym = (uint8_t)v; items[k].extra = 0; k++; }
        }
        i = j;
    }
    return k;
}

typedef struct {
    uint8_t ll_len[ORC_NUM_LL + 2], d_len[ORC_NUM_D + 2], cl_len[ORC_NUM_CL];
    uint16_t ll_code[ORC_NUM_LL + 2], d_code[ORC_NUM_D + 2], cl_code[ORC_NUM_CL];
    orc_cl_item items[ORC_NUM_LL + ORC_NUM_D + 4];
    int nitems, hlit, hdist, hclen;
} orc_dyn;

/* What one block of `ntok` tokens over `n` raw bytes would be: histograms, the dynamic code, exact costs, the choice. */
typedef struct {
    uint32_t ll_freq[ORC_NUM_LL + 2], d_freq[ORC_NUM_D + 2], cl_freq[ORC_NUM_CL];
    orc_dyn dy;
    int64_t extra_bits, d_extra_bits, fixed_bits, dyn_bits, stored_bits;
    int btype;
} orc_plan;

static int64_t plan_bits(const orc_plan *pl) { return pl->btype == 0 ? pl->stored_bits : pl->btype == 1 ? pl->fixed_bits : pl->dyn_bits; }

static void orc_plan_block(orc_plan *pl, long n, const uint32_t *tokens, long ntok, int block_mode)
{
    uint8_t seq[ORC_NUM_LL + ORC_NUM_D + 4];
    orc_dyn *dy = &pl->dy;
    long k, nstored;
    int s, eb, ev, i;

    memset(pl->ll_freq, 0, sizeof(pl->ll_freq));
    memset(pl->d_freq, 0, sizeof(pl->d_freq));
    memset(pl->cl_freq, 0, sizeof(pl->cl_freq));
    pl->extra_bits = pl->d_extra_bits = 0;
    for (k = 0; k < ntok; k++) {
        uint32_t t = tokens[k];
        if (ORC_TOK_DIST(t) == 0) {
            pl->ll_freq[t & 0xFF]++;
        } else {
            pl->ll_freq[orc_len_sym((int)ORC_TOK_LEN(t), &eb, &ev)]++;
            pl->extra_bits += eb;
            pl->d_freq[orc_dist_sym((int)ORC_TOK_DIST(t), &eb, &ev)]++;
            pl->extra_bits += eb;
            pl->d_extra_bits += eb;
        }
    }
    pl->ll_freq[256]++; /* end of block, counted once: src/deflate_compress.c:233-234 */

    /* dynamic code (RFC 1951 3.2.7) */
    orc_huff_lengths(pl->ll_freq, ORC_NUM_LL, 15, dy->ll_len);
    orc_huff_lengths(pl->d_freq, ORC_NUM_D, 15, dy->d_len);
    for (dy->hlit = ORC_NUM_LL; dy->hlit > 257 && dy->ll_len[dy->hlit - 1] == 0; dy->hlit--) {}
    for (dy->hdist = ORC_NUM_D; dy->hdist > 1 && dy->d_len[dy->hdist - 1] == 0; dy->hdist--) {}
    memcpy(seq, dy->ll_len, (size_t)dy->hlit);
    memcpy(seq + dy->hlit, dy->d_len, (size_t)dy->hdist);
    dy->nitems = orc_rle_lengths(seq, dy->hlit + dy->hdist, dy->items);
    for (i = 0; i < dy->nitems; i++) pl->cl_freq[dy->items[i].sym]++;
    orc_huff_lengths(pl->cl_freq, ORC_NUM_CL, 7, dy->cl_len);
    for (dy->hclen = 19; dy->hclen > 4 && dy->cl_len[orc_cl_order[dy->hclen - 1]] == 0; dy->hclen--) {}
    orc_canonical_codes(dy->ll_len, ORC_NUM_LL, dy->ll_code);
    orc_canonical_codes(dy->d_len, ORC_NUM_D, dy->d_code);
    orc_canonical_codes(dy->cl_len, ORC_NUM_CL, dy->cl_code);

    /* exact costs (README.md:16,19: "calculate the compression ratio for each case") */
    pl->dyn_bits = 3 + 5 + 5 + 4 + 3 * dy->hclen + pl->extra_bits;
    for (s = 0; s < ORC_NUM_CL; s++) pl->dyn_bits += (int64_t)pl->cl_freq[s] * dy->cl_len[s];
    pl->dyn_bits += 2 * (int64_t)pl->cl_freq[16] + 3 * (int64_t)pl->cl_freq[17] + 7 * (int64_t)pl->cl_freq[18];
    for (s = 0; s < ORC_NUM_LL; s++) pl->dyn_bits += (int64_t)pl->ll_freq[s] * dy->ll_len[s];
    for (s = 0; s < ORC_NUM_D; s++) pl->dyn_bits += (int64_t)pl->d_freq[s] * dy->d_len[s];
    pl->fixed_bits = 3 + pl->extra_bits;
    for (s = 0; s < ORC_NUM_LL; s++) pl->fixed_bits += (int64_t)pl->ll_freq[s] * orc_fixed_ll_len(s);
    for (s = 0; s < ORC_NUM_D; s++) pl->fixed_bits += (int64_t)pl->d_freq[s] * 5;
    nstored = (n + 65534) / 65535;
    if (nstored < 1) nstored = 1;
    /* a segment always starts byte aligned; a stored block inside a split segment is costed the same way (its
     * 3 header bits plus the padding count as one byte: the decision's yardstick, the emitted size is exact) */
    pl->stored_bits = 8 * (n + 5 * nstored);

    if (block_mode == 1) pl->btype = 0;
    else if (block_mode == 2) pl->btype = 1;
    else if (block_mode == 3) pl->btype = 2;
    else if (pl->stored_bits <= pl->fixed_bits && pl->stored_bits <= pl->dyn_bits) pl->btype = 0;
    else if (pl->fixed_bits <= pl->dyn_bits) pl->btype = 1;
    else pl->btype = 2;
}

/* The block itself (not final: the sync marker that follows a segment carries BFINAL). */
static void orc_emit_block(orc_bw *bw, const orc_plan *pl, const uint8_t *raw, long n, const uint32_t *tokens, long ntok)
{
    const orc_dyn *dy = &pl->dy;
    const int btype = pl->btype;
    long k;
    int s, eb, ev, i;
    if (btype == 0) {
        long off = 0;
        do {
            long len = n - off < 65535 ? n - off : 65535;
            bw_put(bw, 0, 3); /* BFINAL=0, BTYPE=00 (RFC 1951 3.2.4) */
            bw_align(bw);
            bw_put(bw, (uint32_t)len, 16);
            bw_put(bw, (uint32_t)(~len & 0xFFFF), 16);
            bw_bytes(bw, raw + off, len);
            off += len;
        } while (off < n);
    } else {
        const uint8_t *ll_len = dy->ll_len, *d_len = dy->d_len;
        const uint16_t *ll_code = dy->ll_code, *d_code = dy->d_code;
        uint8_t fl_len[288], fd_len[32];
        uint16_t fl_code[288], fd_code[32];
        bw_put(bw, 0, 1);
        bw_put(bw, (uint32_t)btype, 2);
        if (btype == 1) {
            /* fixed code of RFC 1951 3.2.6 (decode twins: code_fl/code_fd,
             * src/deflate_decompress.c:98-131) */
            for (s = 0; s < 288; s++) fl_len[s] = (uint8_t)orc_fixed_ll_len(s);
            for (s = 0; s < 32; s++) fd_len[s] = 5;
            orc_canonical_codes(fl_len, 288, fl_code);
            orc_canonical_codes(fd_len, 32, fd_code);
            ll_len = fl_len; d_len = fd_len; ll_code = fl_code; d_code = fd_code;
        } else {
            bw_put(bw, (uint32_t)(dy->hlit - 257), 5);
            bw_put(bw, (uint32_t)(dy->hdist - 1), 5);
            bw_put(bw, (uint32_t)(dy->hclen - 4), 4);
            for (i = 0; i < dy->hclen; i++) bw_put(bw, dy->cl_len[orc_cl_order[i]], 3);
            for (i = 0; i < dy->nitems; i++) {
                int cs = dy->items[i].sym;
                bw_put(bw, dy->cl_code[cs], dy->cl_len[cs]);
                if (cs == 16) bw_put(bw, dy->items[i].extra, 2);
                else if (cs == 17) bw_put(bw, dy->items[i].extra, 3);
                else if (cs == 18) bw_put(bw, dy->items[i].extra, 7);
            }
        }
        for (k = 0; k < ntok; k++) {
            uint32_t t = tokens[k];
            if (ORC_TOK_DIST(t) == 0) {
                bw_put(bw, ll_code[t & 0xFF], ll_len[t & 0xFF]);
            } else {
                s = orc_len_sym((int)ORC_TOK_LEN(t), &eb, &ev);
                bw_put(bw, ll_code[s], ll_len[s]);
                if (eb) bw_put(bw, (uint32_t)ev, eb);
                s = orc_dist_sym((int)ORC_TOK_DIST(t), &eb, &ev);
                bw_put(bw, d_code[s], d_len[s]);
                if (eb) bw_put(bw, (uint32_t)ev, eb);
            }
        }
        bw_put(bw, ll_code[256], ll_len[256]);
    }
}

/* cost split of a planned block (what the reference's per-token record estimates, src/deflate_compress.c:291-309) */
static void plan_split_costs(const orc_plan *pl, int64_t *tree, int64_t *ll, int64_t *d)
{
    int s;
    *tree = 0; *d = 0;
    if (pl->btype == 0) {
        *ll = pl->stored_bits - 3;
    } else {
        int64_t llb = pl->extra_bits - pl->d_extra_bits, db = pl->d_extra_bits;
        for (s = 0; s < ORC_NUM_LL; s++) llb += (int64_t)pl->ll_freq[s] * (pl->btype == 2 ? pl->dy.ll_len[s] : orc_fixed_ll_len(s));
        for (s = 0; s < ORC_NUM_D; s++) db += (int64_t)pl->d_freq[s] * (pl->btype == 2 ? pl->dy.d_len[s] : 5);
        *ll = llb;
        *d = db;
        *tree = plan_bits(pl) - 3 - llb - db;
    }
}

/*
 * Adaptive block splitting (block_mode 4; the reference's stated aim, README.md:15-20 "how to divide the data up into
 * chunks, each able to have its own Huffman coding", which it approaches with the running cost of an adaptive tree,
 * src/aht.c:239-277 -- never wired to a decision).  Decided here from exact block costs: the segment's tokens are cut
 * where they cross multiples of ORC_SUB_BYTES input bytes (sub-block j = the tokens that START in
 * [j * ORC_SUB_BYTES, (j + 1) * ORC_SUB_BYTES)); a range of sub-blocks is bisected when its two halves, each with its own
 * cheapest block, cost strictly fewer bits than the range as one block; the halves are then looked at the same way.
 * The token stream is untouched.
 */
static void orc_split_range(orc_bw *bw, const uint8_t *raw, const uint32_t *tokens, const long *tok_start, const long *pos_start,
                            int a, int b, const orc_plan *whole, orc_seg_stats *st)
{
    orc_plan *L = NULL, *R = NULL;
    if (b - a >= 2) {
        const int mid = (a + b) / 2;
        L = (orc_plan *)malloc(sizeof(orc_plan));
        R = (orc_plan *)malloc(sizeof(orc_plan));
        orc_plan_block(L, pos_start[mid] - pos_start[a], tokens + tok_start[a], tok_start[mid] - tok_start[a], 0);
        orc_plan_block(R, pos_start[b] - pos_start[mid], tokens + tok_start[mid], tok_start[b] - tok_start[mid], 0);
        if (plan_bits(L) + plan_bits(R) < plan_bits(whole)) {
            orc_split_range(bw, raw, tokens, tok_start, pos_start, a, mid, L, st);
            orc_split_range(bw, raw, tokens, tok_start, pos_start, mid, b, R, st);
            free(L); free(R);
            return;
        }
        free(L); free(R);
    }
    orc_emit_block(bw, whole, raw + pos_start[a], pos_start[b] - pos_start[a], tokens + tok_start[a], tok_start[b] - tok_start[a]);
    if (st) {
        int64_t t, l, d;
        plan_split_costs(whole, &t, &l, &d);
        if (st->nblocks == 0) st->btype = whole->btype;
        st->nblocks++;
        st->block_bits += plan_bits(whole);
        st->tree_bits += t; st->ll_bits += l; st->d_bits += d;
    }
}

/* Emit one segment: its block(s) then the byte-aligning empty stored block. */
static void orc_emit_segment(orc_bw *bw, const uint8_t *raw, long n, const uint32_t *tokens,
                             long ntok, const orc_params *prm, int final_seg, orc_seg_stats *st)
{
    orc_plan *pl = (orc_plan *)malloc(sizeof(orc_plan));
    long start_pos = bw->pos;
    orc_plan_block(pl, n, tokens, ntok, prm->block_mode == 4 ? 0 : prm->block_mode);
    if (st) {
        memset(st, 0, sizeof(*st));
        st->ntokens = (int32_t)ntok;
        st->stored_bits = pl->stored_bits;
        st->fixed_bits = pl->fixed_bits;
        st->dynamic_bits = pl->dyn_bits;
    }
    if (prm->block_mode == 4 && n > ORC_SUB_BYTES) {
        long tok_start[ORC_SUB_MAX + 1], pos_start[ORC_SUB_MAX + 1], k, pos = 0;
        const int nsub = (int)((n + ORC_SUB_BYTES - 1) / ORC_SUB_BYTES);
        int j = 0;
        for (k = 0; k < ntok; k++) {
            while (j <= pos / ORC_SUB_BYTES) { tok_start[j] = k; pos_start[j] = pos; j++; } /* (a token is shorter than a sub-block) */
            pos += ORC_TOK_DIST(tokens[k]) ? (long)ORC_TOK_LEN(tokens[k]) : 1;
        }
        for (; j <= nsub; j++) { tok_start[j] = ntok; pos_start[j] = n; }
        tok_start[nsub] = ntok; pos_start[nsub] = n;
        orc_split_range(bw, raw, tokens, tok_start, pos_start, 0, nsub, pl, st);
    } else {
        orc_emit_block(bw, pl, raw, n, tokens, ntok);
        if (st) {
            st->btype = pl->btype;
            st->nblocks = 1;
            st->block_bits = plan_bits(pl);
            plan_split_costs(pl, &st->tree_bits, &st->ll_bits, &st->d_bits);
        }
    }
    /* sync marker: empty stored block, byte aligns the segment
     * (RFC 1951 3.2.4, docs/rfc1951.txt:576-587) */
    bw_put(bw, final_seg ? 1u : 0u, 1);
    bw_put(bw, 0, 2);
    bw_align(bw);
    bw_put(bw, 0x0000, 16);
    bw_put(bw, 0xFFFF, 16);
    if (st) st->out_bytes = bw->pos - start_pos;
    free(pl);
}

long orc_num_segments(long n, const orc_params *prm)
{
    return n <= 0 ? 0 : (n + prm->seg_bytes - 1) / prm->seg_bytes;
}

long orc_bound(long n, const orc_params *prm)
{
    long nseg = orc_num_segments(n, prm);
    long per = 5 * ((prm->seg_bytes + 65534) / 65535) + 5 + 8 + (prm->block_mode == 4 ? 6 * ORC_SUB_MAX : 0);
    if (prm->block_mode >= 2) return 2 * n + nseg * 640 + 64; /* forced fixed/dynamic may expand */
    return n + nseg * per + 64;
}

long orc_deflate_body(const uint8_t *buf, long data_lo, long data_hi, const orc_params *prm,
                      int is_final, uint8_t *out, long cap, orc_seg_stats *stats, long stats_cap)
{
    orc_bw bw;
    uint32_t *tokens;
    long s, nseg = orc_num_segments(data_hi - data_lo, prm);

    memset(&bw, 0, sizeof(bw));
    bw.out = out;
    bw.cap = cap;
    if (nseg == 0) {
        /* empty input: a lone (final) empty stored block */
        bw_put(&bw, is_final ? 1u : 0u, 1);
        bw_put(&bw, 0, 2);
        bw_align(&bw);
        bw_put(&bw, 0x0000, 16);
        bw_put(&bw, 0xFFFF, 16);
        return bw.err ? -1 : bw.pos;
    }
    tokens = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)prm->seg_bytes);
    if (!tokens) return -1;
    for (s = 0; s < nseg; s++) {
        long lo = data_lo + s * prm->seg_bytes;
        long hi = lo + prm->seg_bytes < data_hi ? lo + prm->seg_bytes : data_hi;
        long ntok = orc_lz77_segment(buf, 0, lo, hi, prm, tokens);
        if (ntok < 0) { free(tokens); return -1; }
        orc_emit_segment(&bw, buf + lo, hi - lo, tokens, ntok, prm, is_final && s == nseg - 1,
                         (stats && s < stats_cap) ? &stats[s] : NULL);
        if (bw.err) { free(tokens); return -2; }
    }
    free(tokens);
    return bw.pos;
}

/* ------------------------------------------------------------------------- */
/* zlib container (RFC 1950)                                                 */
/* ------------------------------------------------------------------------- */

/* Same recurrence as adler32, src/deflate_decompress.c:60-68. */
uint32_t orc_adler32(uint32_t adler, const uint8_t *b, size_t len)
{
    uint32_t s1 = adler & 0xFFFF, s2 = adler >> 16;
    size_t i;
    for (i = 0; i < len; i++) {
        s1 = (s1 + b[i]) % 65521;
        s2 = (s2 + s1) % 65521;
    }
    return (s2 << 16) | s1;
}

/* Adler-32 of a concatenation from the parts' checksums. */
uint32_t orc_adler32_combine(uint32_t a1, uint32_t a2, uint64_t len2)
{
    const uint64_t M = 65521;
    uint64_t s1a = a1 & 0xFFFF, s2a = a1 >> 16, s1b = a2 & 0xFFFF, s2b = a2 >> 16;
    uint64_t s1 = (s1a + s1b + M - 1) % M;
    uint64_t s2 = (s2a + s2b + (len2 % M) * ((s1a + M - 1) % M)) % M;
    return (uint32_t)((s2 << 16) | s1);
}

long orc_zlib_compress(const uint8_t *in, long n, const orc_params *prm, uint8_t *out, long cap)
{
    long body;
    uint32_t a;
    if (cap < 6) return -1;
    out[0] = 0x78; /* CM=8, CINFO=7 (RFC 1950 2.2; checked at src/deflate_decompress.c:352-360) */
    out[1] = 0x9C; /* FLEVEL=2, FDICT=0, FCHECK so that (CMF*256+FLG) % 31 == 0 */
    body = orc_deflate_body(in, 0, n, prm, 1, out + 2, cap - 6, NULL, 0);
    if (body < 0) return body;
    a = orc_adler32(1, in, (size_t)n);
    out[2 + body + 0] = (uint8_t)(a >> 24); /* MSB first, RFC 1950 2.2 (docs/rfc1950.txt:328-329) */
    out[2 + body + 1] = (uint8_t)(a >> 16);
    out[2 + body + 2] = (uint8_t)(a >> 8);
    out[2 + body + 3] = (uint8_t)a;
    return body + 6;
}

/* ------------------------------------------------------------------------- */
/* inflate: repaired restatement of src/deflate_decompress.c                 */
/* ------------------------------------------------------------------------- */

typedef struct {
    const uint8_t *in;
    long n, pos;
    uint32_t acc;
    int nacc;
    int err;
} orc_br;

/* LSB-first bit reader (reference: _bits32/read_bits32, src/include/globals.h:52-64) */
static uint32_t br_bits(orc_br *b, int n)
{
    uint32_t v;
    while (b->nacc < n) {
        if (b->pos >= b->n) { b->err = 1; return 0; }
        b->acc |= (uint32_t)b->in[b->pos++] << b->nacc;
        b->nacc += 8;
    }
    v = b->acc & ((n == 32) ? 0xFFFFFFFFu : ((1u << n) - 1));
    b->acc >>= n;
    b->nacc -= n;
    return v;
}

typedef struct {
    uint16_t count[16];
    uint16_t symbol[288];
} orc_hdec;

/* Canonical decode table from code lengths: the RFC 1951 3.2.2 procedure that
 * form_h_tree implements (src/deflate_decompress.c:146-161).  Returns 0 for a
 * complete code, >0 incomplete, <0 over-subscribed. */
static int hdec_build(orc_hdec *h, const uint8_t *len, int n)
{
    uint16_t offs[16];
    int s, l, left = 1;
    memset(h->count, 0, sizeof(h->count));
    for (s = 0; s < n; s++) h->count[len[s]]++;
    if (h->count[0] == n) return 0;
    for (l = 1; l <= 15; l++) {
        left <<= 1;
        left -= h->count[l];
        if (left < 0) return left;
    }
    offs[1] = 0;
    for (l = 1; l < 15; l++) offs[l + 1] = (uint16_t)(offs[l] + h->count[l]);
    for (s = 0; s < n; s++)
        if (len[s]) h->symbol[offs[len[s]]++] = (uint16_t)s;
    return left;
}

/* Bit-serial symbol decode (reference: h_tree_lookup, src/h_tree.c:24-40). */
static int hdec_sym(orc_br *b, const orc_hdec *h)
{
    int code = 0, first = 0, index = 0, l;
    for (l = 1; l <= 15; l++) {
        int count = h->count[l];
        code |= (int)br_bits(b, 1);
        if (b->err) return -1;
        if (code - count < first) return h->symbol[index + (code - first)];
        index += count;
        first += count;
        first <<= 1;
        code <<= 1;
    }
    return -1; /* E_HUFINV */
}

static const uint16_t orc_lbase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31,
                                       35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
static const uint8_t orc_lext[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2,
                                     3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
static const uint16_t orc_dbase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193,
                                       257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145,
                                       8193, 12289, 16385, 24577};
static const uint8_t orc_dext[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6,
                                     7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};

/* Symbol loop: do_decompress, src/deflate_decompress.c:238-288. */
static int orc_inflate_codes(orc_br *b, const orc_hdec *hl, const orc_hdec *hd, uint8_t *out,
                             long cap, long *pn)
{
    long n = *pn;
    for (;;) {
        int s = hdec_sym(b, hl);
        if (s < 0) return -1;
        if (s == 256) break;
        if (s < 256) {
            if (n >= cap) return -2;
            out[n++] = (uint8_t)s;
        } else {
            int len, dist, ds;
            long i;
            s -= 257;
            if (s >= 29) return -3;
            len = orc_lbase[s] + (int)br_bits(b, orc_lext[s]);
            ds = hdec_sym(b, hd);
            if (ds < 0 || ds >= 30) return -4;
            dist = orc_dbase[ds] + (int)br_bits(b, orc_dext[ds]);
            if (b->err) return -1;
            if (dist > n) return -5; /* E_HUFDIS */
            if (n + len > cap) return -2;
            for (i = 0; i < len; i++, n++) out[n] = out[n - dist];
        }
    }
    *pn = n;
    return 0;
}

/* Block loop: deflate_block, src/deflate_decompress.c:291-345 (stored :303-314,
 * dynamic :315-324 with form_d1 :164-177 / form_d2 :180-235, fixed :325-330). */
long orc_inflate_raw(const uint8_t *in, long n, uint8_t *out, long cap, long *consumed)
{
    orc_br b;
    long outn = 0;
    int bfinal;
    memset(&b, 0, sizeof(b));
    b.in = in;
    b.n = n;
    do {
        int btype;
        bfinal = (int)br_bits(&b, 1);
        btype = (int)br_bits(&b, 2);
        if (b.err) return -1;
        if (btype == 0) {
            uint32_t len, nlen;
            b.acc = 0; b.nacc = 0; /* skip to byte boundary */
            if (b.pos + 4 > b.n) return -1;
            len = in[b.pos] | ((uint32_t)in[b.pos + 1] << 8);
            nlen = in[b.pos + 2] | ((uint32_t)in[b.pos + 3] << 8);
            b.pos += 4;
            if (len != (~nlen & 0xFFFF)) return -6; /* E_ZNLEN */
            if (b.pos + (long)len > b.n) return -1;
            if (outn + (long)len > cap) return -2;
            memcpy(out + outn, in + b.pos, len);
            b.pos += len;
            outn += len;
        } else if (btype == 1 || btype == 2) {
            orc_hdec hl, hd;
            uint8_t lens[320];
            int rc;
            if (btype == 1) {
                int s;
                for (s = 0; s < 288; s++) lens[s] = (uint8_t)orc_fixed_ll_len(s);
                hdec_build(&hl, lens, 288);
                for (s = 0; s < 30; s++) lens[s] = 5;
                hdec_build(&hd, lens, 30);
            } else {
                orc_hdec hc;
                uint8_t cl[19];
                int hlit = (int)br_bits(&b, 5) + 257;
                int hdist = (int)br_bits(&b, 5) + 1;
                int hclen = (int)br_bits(&b, 4) + 4;
                int i = 0;
                if (b.err) return -1;
                if (hlit > 286 || hdist > 30) return -7; /* E_ZINV */
                memset(cl, 0, sizeof(cl));
                for (i = 0; i < hclen; i++) cl[orc_cl_order[i]] = (uint8_t)br_bits(&b, 3);
                if (hdec_build(&hc, cl, 19) != 0) return -8;
                i = 0;
                while (i < hlit + hdist) {
                    int s = hdec_sym(&b, &hc), rep, v = 0;
                    if (s < 0) return -1;
                    if (s < 16) { lens[i++] = (uint8_t)s; continue; }
                    if (s == 16) {
                        if (i == 0) return -9; /* E_HUFINV: no previous length */
                        v = lens[i - 1];
                        rep = 3 + (int)br_bits(&b, 2);
                    } else if (s == 17) rep = 3 + (int)br_bits(&b, 3);
                    else rep = 11 + (int)br_bits(&b, 7);
                    if (i + rep > hlit + hdist) return -10;
                    while (rep--) lens[i++] = (uint8_t)v;
                }
                if (b.err) return -1;
                if (lens[256] == 0) return -11;
                rc = hdec_build(&hl, lens, hlit);
                if (rc < 0 || (rc > 0 && hlit - hl.count[0] != 1)) return -12;
                rc = hdec_build(&hd, lens + hlit, hdist);
                if (rc < 0 || (rc > 0 && hdist - hd.count[0] != 1)) return -13;
            }
            rc = orc_inflate_codes(&b, &hl, &hd, out, cap, &outn);
            if (rc) return rc;
        } else {
            return -14; /* E_ZBTYPE */
        }
    } while (!bfinal);
    if (consumed) *consumed = b.pos;
    return outn;
}

/* deflate_decompress, src/deflate_decompress.c:371-409: header check
 * (deflate_decompress_header :347-368), blocks, Adler-32 (MSB first). */
long orc_zlib_decompress(const uint8_t *in, long n, uint8_t *out, long cap)
{
    long used = 0, outn;
    uint32_t a;
    if (n < 6) return -20;                                   /* E_ZHEAD */
    if ((((unsigned)in[0] << 8) | in[1]) % 31) return -21;   /* E_ZFCHCK */
    if ((in[0] & 0x0F) != 8) return -22;                     /* E_ZCMPMT */
    if ((in[0] >> 4) > 7) return -23;                        /* E_ZSLWIN */
    if (in[1] & 0x20) return -24;                            /* E_ZPDICT */
    outn = orc_inflate_raw(in + 2, n - 6, out, cap, &used);
    if (outn < 0) return outn;
    a = ((uint32_t)in[2 + used] << 24) | ((uint32_t)in[3 + used] << 16) |
        ((uint32_t)in[4 + used] << 8) | in[5 + used];
    if (a != orc_adler32(1, out, (size_t)outn)) return -25;  /* E_ZADL32 */
    return outn;
}

/* ------------------------------------------------------------------------- */
/* PNG scanline filtering                        */
/* ------------------------------------------------------------------------- */

/* Paeth predictor of the PNG specification (the reference's Paeth, src/png/png_decoder.c:803-817,
 * assigns pa three times and never sets pb / pc: SURVEY 8(d) says to use the specification's). */
static int orc_paeth(int a, int b, int c)
{
    int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

/* The forward direction of defilter() (src/png/png_decoder.c:819-857; filter ids
 * src/png/include/png_decoder.h:57-61).  mode 0: per row the type with the smallest sum of
 * |residual as a signed byte| (first minimum in the order None, Sub, Up, Average, Paeth);
 * mode 1..5: type mode - 1 for every row.  out: rows * (row_bytes + 1) bytes. */
void orc_png_filter(const uint8_t *pix, long row_bytes, long rows, int bpp, int mode, uint8_t *out)
{
    long y, x;
    for (y = 0; y < rows; y++) {
        const uint8_t *cur = pix + y * row_bytes, *prev = y ? cur - row_bytes : NULL;
        uint8_t *o = out + y * (row_bytes + 1);
        unsigned long sum[5] = {0, 0, 0, 0, 0};
        int type = mode ? mode - 1 : 0, k;
        for (k = 0; k < 5; k++) {
            if (mode && k != type) continue;
            for (x = 0; x < row_bytes; x++) {
                int v = cur[x], a = x >= bpp ? cur[x - bpp] : 0, b = prev ? prev[x] : 0, c = (prev && x >= bpp) ? prev[x - bpp] : 0;
                int pred = k == 0 ? 0 : k == 1 ? a : k == 2 ? b : k == 3 ? (a + b) >> 1 : orc_paeth(a, b, c);
                uint8_t r = (uint8_t)(v - pred);
                sum[k] += r < 128 ? r : 256 - r;
            }
        }
        if (!mode)
            for (k = 1; k < 5; k++)
                if (sum[k] < sum[type]) type = k;
        o[0] = (uint8_t)type;
        for (x = 0; x < row_bytes; x++) {
            int v = cur[x], a = x >= bpp ? cur[x - bpp] : 0, b = prev ? prev[x] : 0, c = (prev && x >= bpp) ? prev[x - bpp] : 0;
            int pred = type == 0 ? 0 : type == 1 ? a : type == 2 ? b : type == 3 ? (a + b) >> 1 : orc_paeth(a, b, c);
            o[1 + x] = (uint8_t)(v - pred);
        }
    }
}

/* CRC-32 of PNG chunks (calc_crc, src/include/crc.h:22-30: reflected 0xEDB88320, init and final ~0). */
uint32_t orc_crc32(const uint8_t *data, size_t len)
{
    uint32_t c = 0xFFFFFFFFu;
    size_t n;
    int k;
    for (n = 0; n < len; n++) {
        c ^= data[n];
        for (k = 0; k < 8; k++) c = (c >> 1) ^ (0xEDB88320u & (0u - (c & 1u)));
    }
    return ~c;
}
