/*
 * host/deflate_decompress.c -- deflate_decompress() of the drop-in API: a
 * working RFC 1950/1951 decoder under the reference's signature
 * (src/include/deflate_ext.h:16).  The reference's own decoder
 * (src/deflate_decompress.c) does not compile and is wrong in >= 15 places
 * (SURVEY.md section 8(c)); this keeps its structure -- header check (:347-368),
 * block loop (:291-345), dynamic-tree parse (:164-235), symbol loop (:238-288),
 * Adler-32 (:60-68, stored MSB first per RFC 1950) -- with table-driven decoding.
 * CPU code, like the reference's; the GPU path is compress only.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "deflate_errors.h"
#include "deflate_ext.h"

#define FAST_BITS 10

typedef struct {
    uint16_t fast[1 << FAST_BITS]; /* len << 12 | symbol, 0 = not a short code */
    uint16_t first[17];            /* first canonical code of each length */
    uint16_t offs[17];             /* index of that code's symbol in sorted[] */
    uint16_t count[17];
    uint16_t sorted[288];
    int used;                      /* number of symbols with a code */
} hufftab;

typedef struct {
    const unsigned char *p, *end;
    uint64_t acc;
    int nacc;
} bitsrc;

typedef struct {
    unsigned char *d;
    size_t n, cap;
} outbuf;

static inline void refill(bitsrc *b)
{
    while (b->nacc <= 56 && b->p < b->end) {
        b->acc |= (uint64_t)*b->p++ << b->nacc;
        b->nacc += 8;
    }
}

static inline uint32_t getbits(bitsrc *b, int n, int *err)
{
    uint32_t v;
    if (b->nacc < n) {
        refill(b);
        if (b->nacc < n) { *err = 1; return 0; }
    }
    v = (uint32_t)(b->acc & ((1ull << n) - 1));
    b->acc >>= n;
    b->nacc -= n;
    return v;
}

static uint32_t rev16(uint32_t v, int n)
{
    uint32_t r = 0;
    int i;
    for (i = 0; i < n; i++) { r = (r << 1) | (v & 1); v >>= 1; }
    return r;
}

/* RFC 1951 3.2.2; 0 ok, <0 over-subscribed, >0 incomplete */
static int build(hufftab *h, const unsigned char *len, int n)
{
    int s, l, left = 1, code = 0;
    uint16_t next[17];
    memset(h->fast, 0, sizeof(h->fast));
    memset(h->count, 0, sizeof(h->count));
    for (s = 0; s < n; s++) h->count[len[s]]++;
    h->used = n - h->count[0];
    h->count[0] = 0;
    for (l = 1; l <= 15; l++) {
        left = (left << 1) - h->count[l];
        if (left < 0) return -1;
    }
    h->offs[1] = 0;
    for (l = 1; l <= 15; l++) {
        h->first[l] = (uint16_t)code;
        code = (code + h->count[l]) << 1;
        if (l < 15) h->offs[l + 1] = (uint16_t)(h->offs[l] + h->count[l]);
        next[l] = h->offs[l];
    }
    for (s = 0; s < n; s++) {
        l = len[s];
        if (!l) continue;
        {
            int idx = next[l]++;
            int c = h->first[l] + (idx - h->offs[l]);
            h->sorted[idx] = (uint16_t)s;
            if (l <= FAST_BITS) {
                uint32_t r = rev16((uint32_t)c, l), step = 1u << l;
                for (; r < (1u << FAST_BITS); r += step) h->fast[r] = (uint16_t)((l << 12) | s);
            }
        }
    }
    return left;
}

static inline int decode(bitsrc *b, const hufftab *h, int *err)
{
    uint32_t e;
    int l, code;
    if (b->nacc < 15) refill(b);
    e = h->fast[b->acc & ((1u << FAST_BITS) - 1)];
    if (e) {
        l = e >> 12;
        if (l > b->nacc) { *err = 1; return 0; }
        b->acc >>= l;
        b->nacc -= l;
        return e & 0xFFF;
    }
    /* long code: canonical walk (h_tree_lookup's role, src/h_tree.c:24-40) */
    code = (int)rev16((uint32_t)(b->acc & ((1u << FAST_BITS) - 1)), FAST_BITS);
    for (l = FAST_BITS + 1; l <= 15; l++) {
        if (l > b->nacc) { *err = 1; return 0; }
        code = (code << 1) | (int)((b->acc >> (l - 1)) & 1);
        if (code - (int)h->first[l] < (int)h->count[l] && code >= (int)h->first[l]) {
            b->acc >>= l;
            b->nacc -= l;
            return h->sorted[h->offs[l] + (code - h->first[l])];
        }
    }
    *err = 2;
    return 0;
}

static int out_reserve(outbuf *o, size_t extra)
{
    if (o->n + extra > o->cap) {
        size_t nc = o->cap ? o->cap : 4096;
        unsigned char *nd;
        while (nc < o->n + extra) nc *= 2;
        nd = (unsigned char *)realloc(o->d, nc);
        if (!nd) return -1;
        o->d = nd;
        o->cap = nc;
    }
    return 0;
}

static const uint16_t lbase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31,
                                   35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
static const uint8_t lext[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
static const uint16_t dbase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769,
                                   1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
static const uint8_t dext[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
static const uint8_t clorder[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15}; /* src/deflate_decompress.c:23 */

/* do_decompress, src/deflate_decompress.c:238-288 */
static int inflate_codes(bitsrc *b, outbuf *o, const hufftab *hl, const hufftab *hd, size_t window)
{
    int err = 0;
    for (;;) {
        int s = decode(b, hl, &err);
        if (err) return -(int)E_HUFINV;
        if (s < 256) {
            if (out_reserve(o, 1)) return -E_MALLOC;
            o->d[o->n++] = (unsigned char)s;
        } else if (s == 256) {
            return 0;
        } else {
            uint32_t len, dist;
            int ds;
            size_t i;
            s -= 257;
            if (s >= 29) return -(int)E_HUFVAL;
            len = lbase[s] + getbits(b, lext[s], &err);
            ds = decode(b, hd, &err);
            if (err) return -(int)E_HUFINV;
            if (ds >= 30) return -(int)E_HUFVAL;
            dist = dbase[ds] + getbits(b, dext[ds], &err);
            if (err) return -(int)E_HUFINV;
            if (dist > o->n || dist > window) return -(int)E_HUFDIS; /* :278 */
            if (out_reserve(o, len)) return -E_MALLOC;
            if (dist >= len) {
                memcpy(o->d + o->n, o->d + o->n - dist, len);
                o->n += len;
            } else {
                for (i = 0; i < len; i++, o->n++) o->d[o->n] = o->d[o->n - dist];
            }
        }
    }
}

static uint32_t adler32_buf(const unsigned char *p, size_t n)
{
    uint32_t s1 = 1, s2 = 0;
    while (n) {
        size_t k = n < 5552 ? n : 5552, i;
        for (i = 0; i < k; i++) { s1 += p[i]; s2 += s1; }
        s1 %= 65521;
        s2 %= 65521;
        p += k;
        n -= k;
    }
    return (s2 << 16) | s1;
}

int deflate_decompress(struct string_len *decompr_dat, struct string_len *compr_dat, int ops)
{
    bitsrc b;
    outbuf o = {NULL, 0, 0};
    hufftab *hl = NULL, *hd = NULL, *hc = NULL;
    const unsigned char *in;
    size_t n, window;
    int rc = 0, bfinal = 0, err = 0;
    uint32_t a32;

    if (!decompr_dat || !compr_dat) return -E_INVAL;
    decompr_dat->str = NULL; /* poison values if error, src/deflate_decompress.c:377-378 */
    decompr_dat->len = 0;
    in = compr_dat->str;
    n = compr_dat->len;
    if (n == 0) return 0; /* no data, skip (:380) */
    /* deflate_decompress_header, :347-368 */
    if (n < 2 + 4) return -(int)E_ZHEAD;
    if ((((unsigned)in[0] << 8) | in[1]) % 31) return -(int)E_ZFCHCK;
    if ((in[0] & 0x0F) != 8) return -(int)E_ZCMPMT;
    if ((in[0] >> 4) > 7) return -(int)E_ZSLWIN;
    if (in[1] & 0x20) return -(int)E_ZPDICT;
    window = (size_t)1 << ((in[0] >> 4) + 8);

    hl = (hufftab *)malloc(sizeof(hufftab));
    hd = (hufftab *)malloc(sizeof(hufftab));
    hc = (hufftab *)malloc(sizeof(hufftab));
    if (!hl || !hd || !hc) { rc = -E_MALLOC; goto fail; }
    b.p = in + 2;
    b.end = in + n - 4; /* take off adler32 (:386) */
    b.acc = 0;
    b.nacc = 0;

    /* block loop, :391-397 / deflate_block :291-345 */
    while (!bfinal) {
        int btype;
        bfinal = (int)getbits(&b, 1, &err);
        btype = (int)getbits(&b, 2, &err);
        if (err) { rc = -(int)E_ZBSZ; goto fail; }
        if (btype == 0) {
            uint32_t len, nlen;
            int drop = b.nacc & 7;
            b.acc >>= drop;
            b.nacc -= drop;
            len = getbits(&b, 16, &err);
            nlen = getbits(&b, 16, &err);
            if (err) { rc = -(int)E_ZBSZ; goto fail; }
            if (len != (~nlen & 0xFFFF)) { rc = -(int)E_ZNLEN; goto fail; }
            if (out_reserve(&o, len)) { rc = -E_MALLOC; goto fail; }
            while (len && b.nacc >= 8) { /* drain buffered bytes first */
                o.d[o.n++] = (unsigned char)b.acc;
                b.acc >>= 8;
                b.nacc -= 8;
                len--;
            }
            if ((size_t)(b.end - b.p) < len) { rc = -(int)E_ZBSZ; goto fail; }
            memcpy(o.d + o.n, b.p, len);
            o.n += len;
            b.p += len;
        } else if (btype == 1) {
            unsigned char lens[288];
            int s;
            for (s = 0; s < 288; s++) lens[s] = (unsigned char)(s < 144 ? 8 : s < 256 ? 9 : s < 280 ? 7 : 8);
            build(hl, lens, 288);
            for (s = 0; s < 30; s++) lens[s] = 5;
            build(hd, lens, 30);
            rc = inflate_codes(&b, &o, hl, hd, window);
            if (rc) goto fail;
        } else if (btype == 2) {
            unsigned char lens[320], cl[19];
            int hlit = (int)getbits(&b, 5, &err) + 257;
            int hdist = (int)getbits(&b, 5, &err) + 1;
            int hclen = (int)getbits(&b, 4, &err) + 4;
            int i, r;
            if (err) { rc = -(int)E_ZBSZ; goto fail; }
            if (hlit > 286 || hdist > 30) { rc = -(int)E_ZINV; goto fail; } /* :320-321 */
            memset(cl, 0, sizeof(cl));
            for (i = 0; i < hclen; i++) cl[clorder[i]] = (unsigned char)getbits(&b, 3, &err); /* form_d1 :164-177 */
            if (err || build(hc, cl, 19) < 0) { rc = -(int)E_HUFINV; goto fail; }
            for (i = 0; i < hlit + hdist;) { /* form_d2 :180-235 */
                int s = decode(&b, hc, &err), rep, v = 0;
                if (err) { rc = -(int)E_HUFINV; goto fail; }
                if (s < 16) { lens[i++] = (unsigned char)s; continue; }
                if (s == 16) {
                    if (i == 0) { rc = -(int)E_HUFINV; goto fail; } /* no current code length, :200-201 */
                    v = lens[i - 1];
                    rep = 3 + (int)getbits(&b, 2, &err);
                } else if (s == 17) rep = 3 + (int)getbits(&b, 3, &err);
                else rep = 11 + (int)getbits(&b, 7, &err);
                if (err || i + rep > hlit + hdist) { rc = -(int)E_HUFVAL; goto fail; }
                while (rep--) lens[i++] = (unsigned char)v;
            }
            if (lens[256] == 0) { rc = -(int)E_HUFINV; goto fail; }
            r = build(hl, lens, hlit);
            if (r < 0 || (r > 0 && hl->used != 1)) { rc = -(int)E_HUFAMB; goto fail; }
            r = build(hd, lens + hlit, hdist);
            if (r < 0 || (r > 0 && hd->used > 1)) { rc = -(int)E_HUFAMB; goto fail; }
            rc = inflate_codes(&b, &o, hl, hd, window);
            if (rc) goto fail;
        } else {
            rc = -(int)E_ZBTYPE; /* 3 is reserved, :336 */
            goto fail;
        }
    }
    /* footer, :398-406 */
    a32 = ((uint32_t)b.end[0] << 24) | ((uint32_t)b.end[1] << 16) | ((uint32_t)b.end[2] << 8) | b.end[3];
    if (adler32_buf(o.d, o.n) != a32) { rc = -(int)E_ZADL32; goto fail; }
    if ((ops & DEFLATE_NULLTERM) && (o.n == 0 || o.d[o.n - 1] != 0)) {
        if (out_reserve(&o, 1)) { rc = -E_MALLOC; goto fail; }
        o.d[o.n++] = 0;
    }
    if (!o.d) {
        o.d = (unsigned char *)malloc(1);
        if (!o.d) { rc = -E_MALLOC; goto fail; }
    }
    decompr_dat->str = o.d; /* owned by the caller, :405-406 */
    decompr_dat->len = o.n;
    free(hl); free(hd); free(hc);
    return 0;
fail:
    free(o.d);
    free(hl); free(hd); free(hc);
    return rc;
}
