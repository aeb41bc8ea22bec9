/*
 * include/deflate_b200.h -- thin C-ABI over the sm_100a DEFLATE kernels.
 *
 * This is the layer the reference-facing host code (include/deflate_ext.h,
 * csrc/host/deflate_compress.c) calls; it is also what a cgo/JNI/ctypes binding
 * would bind.  Plain pointers and sizes only.  There is NO CPU fallback: every
 * entry point fails with DFL_E_CUDA when no sm_100 device / driver is present.
 *
 * What each entry point replaces in the reference (mparker97/deflate_compression):
 *   dfl_b200_deflate_device / _compress_host
 *       process_loop(), src/deflate_compress.c:219-347: hash-chain longest-match
 *       search (:244-263, check_dup_str :164-180), greedy token decision
 *       (:265-288), per-token code mapping (get_len_code/get_dist_code :182-217)
 *       and the cost accounting it feeds (h_tree_d_lens / h_tree_builder_*,
 *       src/h_tree.c:75-302) -- completed with the encode side the reference
 *       leaves as TODOs (:269,:279): length-limited code lengths, canonical
 *       codes, stored/fixed/dynamic choice (README.md:16,19), bit packing and
 *       the zlib container the decoder demands (src/deflate_decompress.c:347-368,
 *       :402-404).
 *   dfl_b200_tokens
 *       the per-token (ll, d) fields of struct compress_stats written to
 *       fd_stats (src/include/deflate_ext.h:19-31, src/deflate_compress.c:291-309).
 *   dfl_params.window
 *       the `swi sw` argument of deflate_compress (src/include/deflate_ext.h:17).
 */
#ifndef DEFLATE_B200_H
#define DEFLATE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DFL_MIN_MATCH 3       /* src/deflate_compress.c:247 */
#define DFL_MAX_MATCH 258     /* src/include/deflate.h:4 */
#define DFL_MAX_WINDOW 32768  /* README.md:6 */
#define DFL_MAX_SEG 32768     /* one segment = one CTA-resident window pair */
#define DFL_SUB_BYTES 4096    /* block_mode 4: a segment's tokens are cut where they cross multiples of this many input bytes */
#define DFL_SUB_MAX 8         /* DFL_MAX_SEG / DFL_SUB_BYTES: blocks a segment can be split into */
#define DFL_HASH_BUCKETS 15360 /* bucket count (sized to the shared-memory budget); observable only with chain_limit > 0 */
/* bucket of the 3 bytes packed little-endian in v (low 24 bits): multiplicative hash, range-reduced by mul-hi */
#define DFL_HASH3(v) ((uint32_t)(((uint64_t)((((uint32_t)(v)) & 0xFFFFFFu) * 0x9E3779B1u) * DFL_HASH_BUCKETS) >> 32))
/* The index buckets positions by their first FOUR bytes (v = 4 bytes little-endian): the trigram picks a
 * group of 2^DFL_SUB_BITS neighbouring buckets, the fourth byte one bucket of the group.  Every match of
 * >= 4 bytes shares the bucket; a 3-byte match is somewhere in the group. */
#ifndef DFL_SUB_BITS
#define DFL_SUB_BITS 3
#endif
#define DFL_TRI_GROUPS (DFL_HASH_BUCKETS >> DFL_SUB_BITS)
#define DFL_HASH4(v) (((uint32_t)(((uint64_t)((((uint32_t)(v)) & 0xFFFFFFu) * 0x9E3779B1u) * DFL_TRI_GROUPS) >> 32) << DFL_SUB_BITS) | \
                      (((((uint32_t)(v)) >> 24) * 0x9E3779B1u) >> (32 - DFL_SUB_BITS)))

/* error codes: 0 ok, negative on error (reference convention: callers test < 0,
 * src/png/png_decoder.c:249).  Values reuse src/include/global_errors.h:24-35. */
#define DFL_OK 0
#define DFL_E_MALLOC (-2)   /* E_MALLOC */
#define DFL_E_SZ (-6)       /* E_SZ: output buffer too small */
#define DFL_E_RANGE (-10)   /* E_RANGE */
#define DFL_E_INVAL (-11)   /* E_INVAL */
#define DFL_E_CUDA (-100)   /* CUDA runtime / driver / device failure */
#define DFL_E_DATA (-((1 << 24) + 13)) /* -E_ZINV (src/include/deflate_errors.h:21): corrupt stream / wrong index / Adler-32 mismatch */

typedef struct dfl_ctx dfl_ctx;

typedef struct dfl_params {
    int32_t window;      /* 1..32768; reference `sw` */
    int32_t seg_bytes;   /* 1..32768; independent segment (parse restart, one block + sync marker) */
    int32_t chain_limit; /* 0 = walk the whole chain (reference); K = newest K entries of the 4-byte bucket, then the nearest 3-byte match */
    int32_t use_dict;    /* 1 = segments may match into the preceding <= window bytes */
    int32_t block_mode;  /* 0 = cheapest of stored/fixed/dynamic; 1/2/3 force stored/fixed/dynamic; 4 = cheapest, and a
                          * segment may be split into up to DFL_SUB_MAX blocks where that saves bits (adaptive block
                          * splitting, the reference's stated aim: README.md:15-20) */
} dfl_params;

typedef struct dfl_seg_info {
    int32_t btype;       /* 0 stored, 1 fixed, 2 dynamic (block_mode 4: of the segment's first block) */
    int32_t ntokens;
    int64_t block_bits;  /* chosen block(s) incl. 3 header bits each, excl. sync marker (a stored block inside a split
                          * segment counts 8 * (bytes + 5)) */
    int64_t out_bytes;   /* bytes of this segment in the stream incl. sync marker */
    int64_t out_offset;  /* byte offset of the segment in the deflate body */
    /* the block's cost the way struct compress_stats splits it (src/include/deflate_ext.h:19-31);
     * tree_bits + ll_bits + d_bits + 3 * nblocks == block_bits */
    int64_t tree_bits;   /* code description: HLIT/HDIST/HCLEN, code-length code, run-length coded lengths (0 unless dynamic) */
    int64_t ll_bits;     /* lit/len symbols + length extra bits (stored: everything but the 3 header bits) */
    int64_t d_bits;      /* distance symbols + distance extra bits */
    int64_t nblocks;     /* blocks the segment was emitted as: 1 unless block_mode 4 split it */
} dfl_seg_info;

void dfl_b200_default_params(dfl_params *p);
const char *dfl_b200_strerror(int err);

/* Context bound to one CUDA device.  Workspace is allocated lazily and reused. */
int dfl_b200_create(dfl_ctx **ctx, int device, const dfl_params *prm);
void dfl_b200_destroy(dfl_ctx *ctx);
int dfl_b200_set_params(dfl_ctx *ctx, const dfl_params *prm);

/* Upper bound of the raw deflate body for n input bytes. */
size_t dfl_b200_bound(const dfl_params *prm, size_t n);

/*
 * Device-resident compress: d_in[0,n) -> raw DEFLATE body at d_out.
 * d_in[-hist_len, 0) must be readable device memory holding the bytes that
 * precede this range in the stream (dictionary priming for a shard; 0 for a
 * whole stream).  The last segment's sync marker carries BFINAL iff is_final.
 * Asynchronous on `stream` (a cudaStream_t, NULL = default) up to the final
 * size read-back; *out_len and *adler (Adler-32 of d_in[0,n) from start value
 * 1) are valid on return.
 * Stored blocks are copied in aligned 4-byte words of d_in: up to 3 bytes before a segment's first byte and past
 * d_in + n are read when they share an aligned word with the data -- never another word than the data's own, so
 * any allocation that holds the data holds what is read (the match kernel stages exactly [d_in - hist_len, d_in + n)).
 */
int dfl_b200_deflate_device(dfl_ctx *ctx, const uint8_t *d_in, size_t n, size_t hist_len,
                            int is_final, uint8_t *d_out, size_t out_cap, size_t *out_len,
                            uint32_t *adler, void *stream);

/* Same, but enqueue only: sizes stay on the device until dfl_b200_deflate_finish(). */
int dfl_b200_deflate_enqueue(dfl_ctx *ctx, const uint8_t *d_in, size_t n, size_t hist_len,
                             int is_final, uint8_t *d_out, size_t out_cap, void *stream);
int dfl_b200_deflate_finish(dfl_ctx *ctx, size_t *out_len, uint32_t *adler, void *stream);

/*
 * Host-buffer compress (the call behind deflate_compress()): copies in[0,n) to
 * the device, compresses, and copies a complete zlib (RFC 1950) stream back
 * into out.  H2D and D2H are inside the call.
 */
int dfl_b200_compress_host(dfl_ctx *ctx, const uint8_t *in, size_t n, uint8_t *out, size_t cap,
                           size_t *out_len);

/*
 * fd -> fd, the body of deflate_compress() (src/deflate_compress.c:362-376): everything readable from fd_in
 * (read(2) until it returns 0, :139-145; a regular file is read with pread(2) by several threads) becomes a zlib
 * stream on fd_out (-1: discarded).  A reader thread fills pinned pieces of 64 MiB, the copy in, the kernels and
 * the copy out of neighbouring pieces overlap on three streams, a writer thread empties the pinned output
 * pieces; the stream is byte-identical to dfl_b200_compress_host's on the same bytes.
 */
int dfl_b200_compress_fd(dfl_ctx *ctx, int fd_in, int fd_out, unsigned long long *in_bytes, unsigned long long *out_bytes);
int dfl_b200_set_fd_piece(dfl_ctx *ctx, size_t bytes); /* piece size of the fd pipeline (0 = 64 MiB); test support */

/*
 * N GPUs behind one call: ctxs[r] is a context on GPU r, all with the same parameters.
 * in[0, n) shards contiguously into segment-aligned ranges, one per context, each primed with the 32 KiB before
 * it; the bodies land in rank order in out, after the zlib header and before the combined Adler-32.  One host
 * thread drives all devices; the stream equals dfl_b200_compress_host's on one GPU.  Pinned in / out buffers let
 * the copies of different devices overlap.
 */
int dfl_b200_compress_host_multi(dfl_ctx **ctxs, int ndev, const uint8_t *in, size_t n, uint8_t *out, size_t cap,
                                 size_t *out_len);
/* The same from fd to fd (what deflate_compress() runs when DFL_B200_DEVICES names several GPUs): fd_in is read to its
 * end into a pinned buffer that ctxs[0] keeps between calls, the stream is written to fd_out (-1: discarded). */
int dfl_b200_compress_fd_multi(dfl_ctx **ctxs, int ndev, int fd_in, int fd_out, unsigned long long *in_bytes,
                               unsigned long long *out_bytes);

/* cudaMemcpyAsync(dst, src, n, cudaMemcpyDefault, stream): a copy between any two addresses of the unified address
 * space -- e.g. from another GPU's buffer opened through CUDA IPC -- done by the copy engine of the stream's device, no
 * SM involved.  bench.py gathers the ranks' compressed shards on rank 0 with it while the next step's kernels run. */
int dfl_b200_copy_async(void *dst, const void *src, size_t n, void *stream);
/* Adler-32 of a concatenation A | B from adler32(A), adler32(B) (both from start value 1) and len(B). */
uint32_t dfl_b200_adler32_combine(uint32_t a1, uint32_t a2, uint64_t len2);

/* Parity / fd_stats support: keep the token stream of the next compress calls
 * on the host (4 bytes per input byte).  Token = literal byte, or dist<<16|len. */
int dfl_b200_keep_tokens(dfl_ctx *ctx, int enable);
long dfl_b200_num_segments(dfl_ctx *ctx);
int dfl_b200_seg_info(dfl_ctx *ctx, long seg, dfl_seg_info *info);
long dfl_b200_tokens(dfl_ctx *ctx, long seg, uint32_t *tokens, long cap);

/*
 * Segment index + segment-parallel inflate (the GPU half of the reference's
 * deflate_decompress, src/deflate_decompress.c:371-409, for streams this library wrote).
 * Every segment of the body starts byte aligned; the index is the byte offset of each segment in the
 * DEFLATE body plus, as its last entry, the body length (nseg + 1 entries).  It belongs to the last
 * compress call of the context.  With it the Huffman decoding runs one warp per segment; the LZ77 copies,
 * which chain from segment to segment, are resolved by pointer jumping (see k_inflate).
 */
long dfl_b200_seg_index(dfl_ctx *ctx, uint64_t *offsets, long cap); /* -> entries (nseg + 1); offsets may be NULL */
const uint64_t *dfl_b200_seg_index_device(dfl_ctx *ctx);            /* the same array in device memory */
/* d_body/d_out/d_seg_off are device pointers; d_out[-hist_len, 0) must hold the shard's history when the
 * shard was compressed with one.  n = uncompressed length (nseg = ceil(n / seg_bytes)).  Returns the
 * Adler-32 of the n bytes (of this shard alone, as dfl_b200_deflate_device does). */
int dfl_b200_inflate_device(dfl_ctx *ctx, const uint8_t *d_body, size_t body_len, const uint64_t *d_seg_off,
                            size_t nseg, uint32_t seg_bytes, size_t hist_len, uint8_t *d_out, size_t n,
                            uint32_t *adler, void *stream);
int dfl_b200_inflate_rounds(dfl_ctx *ctx); /* pointer-jumping rounds of the last inflate call */
/* Host buffers: a complete zlib stream (header + body + Adler-32) and its index -> n bytes; checks the
 * header fields src/deflate_decompress.c:347-368 checks and the Adler-32 (:402-404). */
int dfl_b200_decompress_host(dfl_ctx *ctx, const uint8_t *in, size_t in_len, const uint64_t *seg_off, size_t nseg,
                             uint32_t seg_bytes, uint8_t *out, size_t n);

/*
 * PNG scanline filtering ahead of the compressor: the forward direction of the
 * reference's defilter() (src/png/png_decoder.c:819-857; filter ids src/png/include/png_decoder.h:57-61).
 * d_pix: rows * row_bytes unfiltered bytes; d_out: rows * (row_bytes + 1) bytes, [type | filtered row] per row.
 * bpp = bytes per pixel.  mode 0 = per row the type with the smallest sum of |residual| (None, Sub, Up,
 * Average, Paeth; first minimum), 1..5 = that type - 1 for every row.
 */
int dfl_b200_png_filter_device(dfl_ctx *ctx, const uint8_t *d_pix, size_t row_bytes, size_t rows, int bpp, int mode,
                               uint8_t *d_out, void *stream);
/* 8-bit grey / grey+alpha / RGB / RGBA pixels (channels 1..4, host memory) -> a PNG file (signature, IHDR, one
 * IDAT, IEND; chunk CRCs as src/include/crc.h:22-30): filter + compress on the GPU. */
size_t dfl_b200_png_bound(const dfl_params *prm, uint32_t width, uint32_t height, uint32_t channels);
int dfl_b200_png_encode_host(dfl_ctx *ctx, const uint8_t *pix, uint32_t width, uint32_t height, uint32_t channels,
                             int filter_mode, uint8_t *out, size_t cap, size_t *out_len);

/* Which kernel resolves the greedy parse of a chunk: 0 = chosen by the chunk's segment count (default),
 * 1 = one warp per segment (k_parse_warp), 2 = one CTA per segment (k_parse).  Same tokens either way. */
int dfl_b200_set_parse_mode(dfl_ctx *ctx, int mode);

/* Which kernel builds the index and searches: 0 = k_match_run (runs of consecutive segments per CTA, the window's
 * sorted index inherited from the segment before) wherever the geometry is the default one (segments and window
 * of 32768 bytes, dictionary priming, no chain limit) and k_match (one CTA per segment) elsewhere; 1 = always
 * k_match.  Same tokens either way. */
int dfl_b200_set_match_mode(dfl_ctx *ctx, int mode);

/* Number of kernels this context has launched (bench.py's gpu_launches). */
long dfl_b200_kernel_launches(dfl_ctx *ctx);
/* Device milliseconds of each kernel class over the last deflate call
 * (CUDA events on the launch stream); names via dfl_b200_kernel_name(). */
int dfl_b200_enable_timing(dfl_ctx *ctx, int enable);
int dfl_b200_num_kernels(void);
const char *dfl_b200_kernel_name(int k);
double dfl_b200_kernel_ms(dfl_ctx *ctx, int k);

/* k_match SM cycles of the last call summed over its CTAs: 0 stage-in, 1 chain build,
 * 2 search, 3 total (profiling aid). */
unsigned long long dfl_b200_phase_cycles(dfl_ctx *ctx, int k);
/* Segments of the last call whose index had to be rebuilt with the slow, ballot-ranked
 * scatter because the fast atomic-ticket scatter failed its order check (expected 0). */
unsigned int dfl_b200_sort_fallbacks(dfl_ctx *ctx);

/* Synthetic corpora generated on the device (bench/test support). */
int dfl_b200_synth_random(uint8_t *d_out, size_t n, uint64_t seed, void *stream);

#ifdef __cplusplus
}
#endif
#endif
