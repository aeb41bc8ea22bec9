/*
 * oracle/oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C) of the reference's DEFLATE compress hot path
 * (mparker97/deflate_compression), used ONLY as the checker for the CUDA path:
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load it; nothing under deflate_compression_b200/ may.
 *
 * Parity pinning:
 *   - token stream: PINNED.  orc_lz77_segment() is checked against the real
 *     reference (oracle/_ref/libref.so, built from /root/reference by
 *     oracle/Makefile) on every input size the reference survives (3..32768 B),
 *     and against committed fixtures generated from it (tests/golden/).
 *   - code lengths / block type / bitstream: "parity unpinned" by the reference
 *     (it never emits a compressed byte: src/deflate_compress.c:269,:279 are
 *     TODOs).  Pinned instead by RFC 1951/1950 and by zlib inflate as an
 *     independent decoder; the reference's cost accounting
 *     (src/h_tree.c:75-148, src/aht.c score) is cross-checked where defined.
 */
#ifndef DEFLATE_ORACLE_H
#define DEFLATE_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MIN_MATCH 3     /* src/deflate_compress.c:247 (max_len starts at 2) */
#define ORC_MAX_MATCH 258   /* src/include/deflate.h:4 MAXLEN */
#define ORC_NUM_LL 286      /* src/include/deflate.h:5 */
#define ORC_NUM_D 30        /* src/include/deflate.h:6 */
#define ORC_NUM_CL 19
#define ORC_HASH_BUCKETS 15360 /* bucket hash range; only observable when chain_limit > 0 */
#define ORC_SUB_BITS 3          /* buckets per trigram group = 2^3 (include/deflate_b200.h DFL_SUB_BITS) */

/* Token encoding shared with the CUDA path: literal -> byte value (dist field 0);
 * match -> (dist << 16) | len, len 3..258, dist 1..32768. */
#define ORC_TOK_LEN(t) ((t) & 0xFFFFu)
#define ORC_TOK_DIST(t) ((t) >> 16)

typedef struct orc_params {
    int32_t window;      /* sliding window (reference arg `sw`), 1..32768 */
    int32_t seg_bytes;   /* independent segment size; parse restarts, one block + sync marker each */
    int32_t chain_limit; /* 0 = unlimited (reference behaviour); K = newest K entries of the 4-byte bucket, then the nearest 3-byte match */
    int32_t use_dict;    /* 1: a segment may match into the <=window bytes before it */
    int32_t block_mode;  /* 0 = choose cheapest of stored/fixed/dynamic; 1/2/3 = force stored/fixed/dynamic;
                          * 4 = cheapest, and a segment may be split into up to ORC_SUB_MAX blocks (adaptive block splitting) */
} orc_params;

#define ORC_SUB_BYTES 4096 /* block_mode 4: a segment is cut where its tokens cross multiples of this many input bytes */
#define ORC_SUB_MAX 8      /* 32768 / ORC_SUB_BYTES */

typedef struct orc_seg_stats {
    int32_t btype;       /* 0 stored, 1 fixed, 2 dynamic (block_mode 4: of the segment's first block) */
    int32_t ntokens;
    int64_t block_bits;  /* bits of the block(s) incl. 3-bit header(s), excl. sync marker (a stored block inside a split
                          * segment counts 8 * (bytes + 5)) */
    int64_t out_bytes;   /* bytes of the segment incl. sync marker */
    int64_t stored_bits, fixed_bits, dynamic_bits;
    /* the chosen block's cost split the way struct compress_stats reports it (src/include/deflate_ext.h:19-31):
     * tree_bits + ll_bits + d_bits + 3 == block_bits */
    int64_t tree_bits; /* code description: HLIT/HDIST/HCLEN, code-length code, run-length coded lengths (0 unless dynamic) */
    int64_t ll_bits;   /* lit/len symbols + length extra bits (stored: everything but the 3 header bits) */
    int64_t d_bits;    /* distance symbols + distance extra bits */
    int64_t nblocks;   /* blocks of the segment (1 unless block_mode 4 split it); tree + ll + d + 3 * nblocks == block_bits */
} orc_seg_stats;

void orc_default_params(orc_params *p);
uint32_t orc_hash3(const uint8_t *p);
uint32_t orc_hash4(const uint8_t *p, uint32_t b3);

/* Greedy LZ77 parse of buf[seg_lo, seg_hi) with history buf[hist_lo, seg_lo).
 * Writes at most (seg_hi - seg_lo) tokens; returns the token count. */
long orc_lz77_segment(const uint8_t *buf, long hist_lo, long seg_lo, long seg_hi,
                      const orc_params *prm, uint32_t *tokens);

/* Length-limited Huffman code lengths (deterministic; the CUDA path must match). */
void orc_huff_lengths(const uint32_t *freq, int n, int limit, uint8_t *len);
void orc_canonical_codes(const uint8_t *len, int n, uint16_t *code_rev);

/* Raw DEFLATE body for buf[data_lo, data_hi): segments, each ending in an
 * empty stored block; the last one carries BFINAL iff is_final.
 * stats (optional) receives one record per segment.  Returns bytes or <0. */
long orc_deflate_body(const uint8_t *buf, long data_lo, long data_hi, const orc_params *prm,
                      int is_final, uint8_t *out, long cap, orc_seg_stats *stats, long stats_cap);
long orc_num_segments(long n, const orc_params *prm);
long orc_bound(long n, const orc_params *prm);

/* zlib (RFC 1950) stream of in[0,n). */
long orc_zlib_compress(const uint8_t *in, long n, const orc_params *prm, uint8_t *out, long cap);

uint32_t orc_adler32(uint32_t adler, const uint8_t *b, size_t len);
uint32_t orc_adler32_combine(uint32_t a1, uint32_t a2, uint64_t len2);

/* Repaired restatement of deflate_decompress (src/deflate_decompress.c:371). */
long orc_inflate_raw(const uint8_t *in, long n, uint8_t *out, long cap, long *consumed);
long orc_zlib_decompress(const uint8_t *in, long n, uint8_t *out, long cap);

/* PNG scanline filtering, the forward direction of defilter() (src/png/png_decoder.c:819-857), and chunk CRC. */
void orc_png_filter(const uint8_t *pix, long row_bytes, long rows, int bpp, int mode, uint8_t *out);
uint32_t orc_crc32(const uint8_t *data, size_t len);

/* Token helpers for tests. */
long orc_tokens_expand(const uint32_t *tokens, long ntok, const uint8_t *hist, long hist_len,
                       uint8_t *out, long cap);

#ifdef __cplusplus
}
#endif
#endif
