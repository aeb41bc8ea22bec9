/*
 * include/deflate_ext.h -- the drop-in boundary.
 *
 * Public surface of mparker97/deflate_compression, restated: same names,
 * argument meaning and return convention as the reference's
 * src/include/deflate_ext.h:6-31 (and struct string_len of
 * src/include/globals.h:42-46), so that a caller built against the reference
 * header (tests/check_lld.c:74, src/png/png_decoder.c:249,533,614,709) links
 * against libdeflate_b200.so unchanged.  The compress side runs on a B200
 * through include/deflate_b200.h; there is no CPU fallback for it.
 */
#ifndef DEFLATE_EXT_H
#define DEFLATE_EXT_H

#include <stddef.h>
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DEFLATE_NULLTERM 1 /* src/include/deflate_ext.h:6 */

typedef unsigned short swi; /* sliding window index, src/include/deflate_ext.h:8 */

#ifndef GLOBALS_H
/* character array with length, so that \0 can be a member (src/include/globals.h:42-46) */
struct string_len {
    unsigned char *str;
    size_t len;
};
#endif

typedef struct deflate_compr deflate_compr_t;      /* src/include/deflate_ext.h:10 */
deflate_compr_t *spawn_deflate_compr_t(void);       /* :11 (SPAWNABLE, globals.h:20-29) */

void deflate_compr_init(deflate_compr_t *com, int fd_in, int fd_out, int fd_stats, swi sw); /* :13 */
void deflate_compr_deinit(deflate_compr_t *com);                                             /* :14 */

/*
 * :16  Decompress the zlib stream compr_dat into a malloc'ed buffer owned by the
 * caller (decompr_dat->str).  ops & DEFLATE_NULLTERM appends a '\0' when the
 * data does not already end in one.  0 on success, negative error code
 * (include/deflate_errors.h) otherwise, with str = NULL, len = 0.
 */
int deflate_decompress(struct string_len *decompr_dat, struct string_len *compr_dat, int ops);

/*
 * :17  Compress everything readable from fd_in (read(2) until it returns 0,
 * src/deflate_compress.c:139-145) into a zlib (RFC 1950) stream on fd_out.
 * fd_stats >= 0 receives one struct compress_stats per token, -1 disables
 * (src/deflate_compress.c:355-361).  sw is the sliding window (1..32768).
 * 0 on success, negative error code otherwise.  The caller owns the fds.
 */
int deflate_compress(int fd_in, int fd_out, int fd_stats, swi sw, int ops);

/* Per-token record on fd_stats, src/include/deflate_ext.h:19-31.  In this
 * implementation the three cost fields are block-level (the exact cost of the block
 * the token belongs to, every token of a block carries the same three numbers), not
 * the reference's running per-token estimate: tree_bits + ll_bits + d_bits + 3 header
 * bits = the block's size in bits.  bytes is an int as in the reference, so
 * deflate_compress() returns -E_RANGE for inputs of 2 GiB or more when fd_stats >= 0. */
struct compress_stats {
    int bytes;     /* input bytes consumed before this token, + 1 */
    int tree_bits; /* bits of the block's code description (dynamic blocks; else 0) */
    int ll_bits;   /* bits of the block's lit/len symbols + length extra bits (stored block: its bytes + LEN/NLEN + padding) */
    int d_bits;    /* bits of the block's distance symbols + distance extra bits */
    int ll;        /* literal byte, or match length */
    int d;         /* 0 for a literal, else match distance */
};

#ifdef __cplusplus
}
#endif
#endif
