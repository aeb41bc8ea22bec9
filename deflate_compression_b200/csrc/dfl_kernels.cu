/*
 * dfl_kernels.cu -- sm_100a DEFLATE compress kernels + the thin C-ABI
 * (include/deflate_b200.h).  B200-native replacement for the hot path of
 * mparker97/deflate_compression: process_loop() in src/deflate_compress.c:219-347
 * (hash-chain match search, greedy parse) and the Huffman accounting of
 * src/h_tree.c:75-302, completed with the encode side the reference leaves as
 * TODOs (src/deflate_compress.c:269,:279).
 *
 * One translation unit; the kernels live in the headers included below:
 *   k_match.cuh   k_match  : one CTA per segment.  [window | segment] staged into shared memory with TMA bulk
 *                            copies, positions sorted by (bucket, position) with an in-CTA radix sort, then
 *                            speculative chunk parsers find the longest / nearest match of every parse head.
 *   k_parse.cuh   k_parse / k_parse_warp : the greedy decision over the evaluated heads, token compaction,
 *                            lit/len + dist histograms.
 *   k_huff.cuh    k_huff   : length-limited code lengths (15/15/7), canonical codes, RFC 1951 3.2.7 header,
 *                            exact stored / fixed / dynamic cost and choice.
 *   k_emit.cuh    k_scan, k_emit, k_final : byte offsets, prefix-sum bit packer, Adler-32.
 *   k_inflate.cuh k_inflate, k_inf_resolve, k_inf_finish : segment-parallel inflate.
 *   k_png.cuh     k_png_filter, k_png_filter4 : PNG scanline filtering.
 * This file: the context and the C-ABI (include/deflate_b200.h).
 * No tensor cores: nothing here is a dense contraction.
 */
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "deflate_b200.h"

#include "dfl_common.cuh"
#include "k_match.cuh"
#include "k_match_run.cuh"
#include "k_parse.cuh"
#include "k_huff.cuh"
#include "k_emit.cuh"
#include "k_inflate.cuh"
#include "k_png.cuh"

/* ------------------------------------------------------------------------- */
/* host side: context + C-ABI                                                */
/* ------------------------------------------------------------------------- */

struct dfl_ctx {
    int device;
    dfl_params prm;
    /* chunk workspace */
    size_t ws_segs, ws_seg_bytes;
    uint32_t *tokcand; /* k_match: candidate token at every evaluated position */
    uint32_t *tokens;  /* k_parse: the segment's tokens, compacted */
    uint16_t *scratch; /* k_match radix-sort ping buffer, one slot per SM */
    SegArrays A;
    void *seg_blob;
    SplitArrays X;     /* block_mode 4: per-block arrays (X.nblk == NULL otherwise) */
    void *split_blob;
    size_t split_segs;
    DevState *state;
    DevState *h_state; /* pinned */
    /* host-buffer path */
    uint8_t *d_in, *d_out;
    size_t d_in_cap, d_out_cap;
    cudaStream_t s_in, s_run, s_out; /* host path: copy-in, kernels, copy-out */
    cudaEvent_t *e_in, *e_done;
    unsigned long long *h_totals;     /* pinned: end offset of every piece */
    int pipe_cap;
    /* fd pipeline (dfl_b200_compress_fd): pinned pieces, device slots */
    uint8_t *fd_h_in[3], *fd_h_out[2], *fd_d_in[2], *fd_d_out[2];
    cudaEvent_t fd_e_in[2], fd_e_out[2];
    size_t fd_piece, fd_piece_req;
    uint8_t *mh_in, *mh_out; /* dfl_b200_compress_fd_multi (kept by the first context): pinned whole-input / whole-output buffers */
    size_t mh_in_cap, mh_out_cap;
    int parse_mode; /* 0 = by segment count, 1 = k_parse_warp, 2 = k_parse */
    int match_mode; /* 0 = k_match_run where the geometry allows, 1 = always k_match */
    int num_sms;
    size_t scratch_persist; /* bytes of the scratch asked to persist in L2 (0: not supported) */
    /* segment index of the last compress call (byte offset of every segment in the body, + the end) and inflate state */
    unsigned long long *seg_index;
    size_t seg_index_cap, seg_index_n;
    uint8_t *d_png; /* PNG encode: the filtered scanlines */
    size_t d_png_cap;
    uint32_t *inf_w; /* inflate: one word per output byte of a chunk */
    size_t inf_cap;
    int inf_rounds;
    /* bookkeeping */
    long launches;
    long long last_n;
    int pending;
    /* token keeping */
    int keep;
    uint32_t *h_tokens;
    size_t h_tokens_cap;
    long nseg_kept;
    uint32_t *h_ntok, *h_btype, *h_out_bytes, *h_cost;
    unsigned long long *h_block_bits, *h_out_off;
    size_t h_seg_cap;
    /* timing */
    int timing;
    cudaEvent_t *ev;
    int *ev_class;
    int ev_n, ev_cap;
    double k_ms[K_NUM];
    char errbuf[256];
};

#define CK(call)                                                                          \
    do {                                                                                  \
        cudaError_t e_ = (call);                                                          \
        if (e_ != cudaSuccess) {                                                          \
            if (ctx) snprintf(ctx->errbuf, sizeof(ctx->errbuf), "%s: %s", #call, cudaGetErrorString(e_)); \
            return DFL_E_CUDA;                                                            \
        }                                                                                 \
    } while (0)

extern "C" void dfl_b200_default_params(dfl_params *p)
{
    p->window = DFL_MAX_WINDOW;
    p->seg_bytes = DFL_MAX_SEG;
    p->chain_limit = 0;
    p->use_dict = 1;
    p->block_mode = 0;
}

extern "C" const char *dfl_b200_strerror(int err)
{
    switch (err) {
    case DFL_OK: return "ok";
    case DFL_E_MALLOC: return "E_MALLOC: allocation failed";
    case DFL_E_SZ: return "E_SZ: output buffer too small";
    case DFL_E_RANGE: return "E_RANGE: parameter out of range";
    case DFL_E_INVAL: return "E_INVAL: invalid argument";
    case DFL_E_CUDA: return "CUDA failure (no sm_100 device, driver error or kernel fault)";
    case DFL_E_DATA: return "E_ZINV: corrupt stream, wrong segment index or Adler-32 mismatch";
    default: return "unknown error";
    }
}

static int check_params(const dfl_params *p)
{
    if (p->window < 1 || p->window > DFL_MAX_WINDOW) return DFL_E_RANGE;
    if (p->seg_bytes < 1 || p->seg_bytes > DFL_MAX_SEG) return DFL_E_RANGE;
    if (p->chain_limit < 0) return DFL_E_RANGE;
    if (p->block_mode < 0 || p->block_mode > 4) return DFL_E_RANGE;
    return DFL_OK;
}

extern "C" size_t dfl_b200_bound(const dfl_params *prm, size_t n)
{
    size_t nseg = n ? (n + prm->seg_bytes - 1) / prm->seg_bytes : 0;
    size_t b = n + nseg * (5 * ((prm->seg_bytes + 65534) / 65535) + 5 + 8) + 64;
    /* a forced fixed/dynamic block may expand: <= 15 bits per literal + header */
    if (prm->block_mode == 2 || prm->block_mode == 3) b = 2 * n + nseg * 640 + 64;
    /* a split segment never takes more bits than the same segment as one block; a stored block inside it may pad */
    if (prm->block_mode == 4) b += nseg * 6 * DFL_SUB_MAX;
    return b;
}

static void free_ws(dfl_ctx *ctx)
{
    if (ctx->tokcand) cudaFree(ctx->tokcand);
    if (ctx->tokens) cudaFree(ctx->tokens);
    ctx->tokens = NULL;
    if (ctx->seg_blob) cudaFree(ctx->seg_blob);
    ctx->tokcand = NULL;
    ctx->seg_blob = NULL;
    ctx->ws_segs = 0;
    if (ctx->split_blob) cudaFree(ctx->split_blob);
    ctx->split_blob = NULL;
    ctx->split_segs = 0;
    memset(&ctx->X, 0, sizeof(ctx->X));
}

/* grow-only device buffer of the context */
static int grow_dev(dfl_ctx *ctx, uint8_t **buf, size_t *cap, size_t need)
{
    if (*cap >= need) return DFL_OK;
    if (*buf) cudaFree(*buf);
    *buf = NULL;
    *cap = 0;
    CK(cudaMalloc(buf, need));
    *cap = need;
    return DFL_OK;
}

static int ensure_index(dfl_ctx *ctx, size_t nseg)
{
    if (ctx->seg_index_cap < nseg + 1) {
        if (ctx->seg_index) cudaFree(ctx->seg_index);
        ctx->seg_index = NULL;
        ctx->seg_index_cap = 0;
        CK(cudaMalloc(&ctx->seg_index, (nseg + 1) * sizeof(unsigned long long)));
        ctx->seg_index_cap = nseg + 1;
    }
    ctx->seg_index_n = nseg;
    return DFL_OK;
}

/* per-block arrays of block_mode 4, made the first time a context splits */
static int ensure_split(dfl_ctx *ctx, size_t ns)
{
    if (ctx->split_segs >= ns) return DFL_OK;
    if (ctx->split_blob) cudaFree(ctx->split_blob);
    ctx->split_blob = NULL;
    ctx->split_segs = 0;
    const size_t nb = ns * DFL_SUB_MAX;
    CK(cudaMalloc(&ctx->split_blob, ns * 4 + nb * (6 + LLS + DS + HDR_WORDS) * sizeof(uint32_t) + 256));
    uint32_t *p = (uint32_t *)ctx->split_blob;
    SplitArrays &X = ctx->X;
    X.nblk = p; p += ns;
    X.btype = p; p += nb;
    X.hdr_bits = p; p += nb;
    X.tok0 = p; p += nb;
    X.ntok = p; p += nb;
    X.pos0 = p; p += nb;
    X.nbytes = p; p += nb;
    X.lltab = p; p += nb * LLS;
    X.dtab = p; p += nb * DS;
    X.hdr = p; p += nb * HDR_WORDS;
    ctx->split_segs = ns;
    return DFL_OK;
}

static int ensure_ws(dfl_ctx *ctx, size_t nseg_chunk)
{
    if (ctx->ws_segs >= nseg_chunk && ctx->ws_seg_bytes == (size_t)ctx->prm.seg_bytes)
        return ctx->prm.block_mode == 4 ? ensure_split(ctx, ctx->ws_segs) : DFL_OK;
    free_ws(ctx);
    size_t ns = nseg_chunk;
    CK(cudaMalloc(&ctx->tokcand, ns * ctx->prm.seg_bytes * sizeof(uint32_t)));
    CK(cudaMalloc(&ctx->tokens, ns * ctx->prm.seg_bytes * sizeof(uint32_t)));
    size_t per = (LLS + DS + LLS + DS + HDR_WORDS + 8) * sizeof(uint32_t) + 2 * sizeof(unsigned long long);
    CK(cudaMalloc(&ctx->seg_blob, ns * per + 256));
    unsigned char *p = (unsigned char *)ctx->seg_blob;
    SegArrays &A = ctx->A;
    A.block_bits = (unsigned long long *)p; p += ns * 8;
    A.out_off = (unsigned long long *)p; p += ns * 8;
    A.ll_freq = (uint32_t *)p; p += ns * LLS * 4;
    A.d_freq = (uint32_t *)p; p += ns * DS * 4;
    A.lltab = (uint32_t *)p; p += ns * LLS * 4;
    A.dtab = (uint32_t *)p; p += ns * DS * 4;
    A.hdr = (uint32_t *)p; p += ns * HDR_WORDS * 4;
    A.ntok = (uint32_t *)p; p += ns * 4;
    A.extra = (uint32_t *)p; p += ns * 4;
    A.hdr_bits = (uint32_t *)p; p += ns * 4;
    A.btype = (uint32_t *)p; p += ns * 4;
    A.out_bytes = (uint32_t *)p; p += ns * 4;
    A.cost = (uint32_t *)p; p += ns * 8;
    ctx->ws_segs = ns;
    ctx->ws_seg_bytes = ctx->prm.seg_bytes;
    return ctx->prm.block_mode == 4 ? ensure_split(ctx, ns) : DFL_OK;
}

extern "C" int dfl_b200_create(dfl_ctx **out, int device, const dfl_params *prm)
{
    dfl_ctx *ctx = NULL;
    int ndev = 0;
    if (!out) return DFL_E_INVAL;
    *out = NULL;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return DFL_E_CUDA; /* no CPU fallback */
    if (device < 0 || device >= ndev) return DFL_E_RANGE;
    ctx = (dfl_ctx *)calloc(1, sizeof(dfl_ctx));
    if (!ctx) return DFL_E_MALLOC;
    ctx->device = device;
    if (prm) ctx->prm = *prm; else dfl_b200_default_params(&ctx->prm);
    int rc = check_params(&ctx->prm);
    if (rc) { free(ctx); return rc; }
    cudaDeviceProp prop;
    if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&prop, device) != cudaSuccess ||
        prop.major < 10) {
        free(ctx);
        return DFL_E_CUDA;
    }
    if (cudaFuncSetAttribute(k_match<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(MatchSmem)) != cudaSuccess ||
        cudaFuncSetAttribute(k_match<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(MatchSmem)) != cudaSuccess ||
        cudaFuncSetAttribute(k_match_run, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(MatchSmem)) != cudaSuccess ||
        cudaFuncSetAttribute(k_parse, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ParseSmem)) != cudaSuccess ||
        cudaMalloc(&ctx->state, sizeof(DevState)) != cudaSuccess ||
        cudaMalloc(&ctx->scratch, (size_t)SCRATCH_SLOTS * SCRATCH_STRIDE * sizeof(uint16_t)) != cudaSuccess ||
        cudaMallocHost(&ctx->h_state, sizeof(DevState)) != cudaSuccess) {
        free(ctx);
        return DFL_E_CUDA;
    }
    ctx->num_sms = prop.multiProcessorCount < SCRATCH_SLOTS ? prop.multiProcessorCount : SCRATCH_SLOTS;
    /* k_match's scratch (radix park, the window block's offsets: 128 KB per resident CTA, reused segment after segment)
     * is asked to stay in L2: without it the input streaming through evicts it and every segment writes it back to HBM */
    ctx->scratch_persist = (size_t)ctx->num_sms * SCRATCH_STRIDE * sizeof(uint16_t);
    if ((size_t)prop.persistingL2CacheMaxSize < ctx->scratch_persist) ctx->scratch_persist = (size_t)prop.persistingL2CacheMaxSize;
    if ((size_t)prop.accessPolicyMaxWindowSize < ctx->scratch_persist) ctx->scratch_persist = (size_t)prop.accessPolicyMaxWindowSize;
    if (ctx->scratch_persist && cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, ctx->scratch_persist) != cudaSuccess) {
        cudaGetLastError();
        ctx->scratch_persist = 0;
    }
    *out = ctx;
    return DFL_OK;
}

static void fd_free(dfl_ctx *ctx);

extern "C" void dfl_b200_destroy(dfl_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    free_ws(ctx);
    if (ctx->state) cudaFree(ctx->state);
    if (ctx->scratch) cudaFree(ctx->scratch);
    if (ctx->h_state) cudaFreeHost(ctx->h_state);
    if (ctx->d_in) cudaFree(ctx->d_in);
    if (ctx->d_out) cudaFree(ctx->d_out);
    if (ctx->h_tokens) cudaFreeHost(ctx->h_tokens);
    if (ctx->seg_index) cudaFree(ctx->seg_index);
    if (ctx->inf_w) cudaFree(ctx->inf_w);
    if (ctx->d_png) cudaFree(ctx->d_png);
    fd_free(ctx);
    if (ctx->mh_in) cudaFreeHost(ctx->mh_in);
    if (ctx->mh_out) cudaFreeHost(ctx->mh_out);
    for (int i = 0; i < ctx->pipe_cap; i++) { cudaEventDestroy(ctx->e_in[i]); cudaEventDestroy(ctx->e_done[i]); }
    free(ctx->e_in); free(ctx->e_done);
    if (ctx->h_totals) cudaFreeHost(ctx->h_totals);
    if (ctx->s_in) { cudaStreamDestroy(ctx->s_in); cudaStreamDestroy(ctx->s_run); cudaStreamDestroy(ctx->s_out); }
    free(ctx->h_ntok); free(ctx->h_btype); free(ctx->h_out_bytes); free(ctx->h_cost); free(ctx->h_block_bits); free(ctx->h_out_off);
    for (int i = 0; i < ctx->ev_cap; i++) cudaEventDestroy(ctx->ev[i]);
    free(ctx->ev); free(ctx->ev_class);
    free(ctx);
}

extern "C" int dfl_b200_set_params(dfl_ctx *ctx, const dfl_params *prm)
{
    if (!ctx || !prm) return DFL_E_INVAL;
    int rc = check_params(prm);
    if (rc) return rc;
    ctx->prm = *prm;
    return DFL_OK;
}

extern "C" int dfl_b200_set_parse_mode(dfl_ctx *ctx, int mode)
{
    if (!ctx || mode < 0 || mode > 2) return DFL_E_INVAL;
    ctx->parse_mode = mode;
    return DFL_OK;
}

extern "C" int dfl_b200_set_match_mode(dfl_ctx *ctx, int mode)
{
    if (!ctx || mode < 0 || mode > 1) return DFL_E_INVAL;
    ctx->match_mode = mode;
    return DFL_OK;
}

extern "C" int dfl_b200_keep_tokens(dfl_ctx *ctx, int enable)
{
    if (!ctx) return DFL_E_INVAL;
    ctx->keep = enable != 0;
    return DFL_OK;
}

extern "C" int dfl_b200_enable_timing(dfl_ctx *ctx, int enable)
{
    if (!ctx) return DFL_E_INVAL;
    ctx->timing = enable != 0;
    return DFL_OK;
}

extern "C" int dfl_b200_num_kernels(void) { return K_NUM; }
extern "C" const char *dfl_b200_kernel_name(int k) { return (k >= 0 && k < K_NUM) ? k_names[k] : ""; }
extern "C" double dfl_b200_kernel_ms(dfl_ctx *ctx, int k) { return (ctx && k >= 0 && k < K_NUM) ? ctx->k_ms[k] : 0.0; }
extern "C" long dfl_b200_kernel_launches(dfl_ctx *ctx) { return ctx ? ctx->launches : 0; }

static int ev_mark(dfl_ctx *ctx, int cls, cudaStream_t st)
{
    /* records one event; events come in (start, stop) pairs per launch */
    if (!ctx->timing) return DFL_OK;
    if (ctx->ev_n == ctx->ev_cap) {
        int ncap = ctx->ev_cap ? ctx->ev_cap * 2 : 256;
        cudaEvent_t *ne = (cudaEvent_t *)realloc(ctx->ev, ncap * sizeof(cudaEvent_t));
        int *nc = (int *)realloc(ctx->ev_class, ncap * sizeof(int));
        if (!ne || !nc) return DFL_E_MALLOC;
        ctx->ev = ne; ctx->ev_class = nc;
        for (int i = ctx->ev_cap; i < ncap; i++) CK(cudaEventCreate(&ctx->ev[i]));
        ctx->ev_cap = ncap;
    }
    ctx->ev_class[ctx->ev_n] = cls;
    CK(cudaEventRecord(ctx->ev[ctx->ev_n], st));
    ctx->ev_n++;
    return DFL_OK;
}

#define LAUNCH(cls, kern, grid, block, smem, st, ...)                 \
    do {                                                              \
        int r_ = ev_mark(ctx, cls, st);                               \
        if (r_) return r_;                                            \
        kern<<<grid, block, smem, st>>>(__VA_ARGS__);                 \
        ctx->launches++;                                              \
        r_ = ev_mark(ctx, cls, st);                                   \
        if (r_) return r_;                                            \
    } while (0)

static int keep_reserve(dfl_ctx *ctx, size_t n, size_t nseg)
{
    if (ctx->h_tokens_cap < n) {
        if (ctx->h_tokens) cudaFreeHost(ctx->h_tokens);
        ctx->h_tokens = NULL;
        ctx->h_tokens_cap = 0;
        CK(cudaMallocHost(&ctx->h_tokens, (n ? n : 1) * sizeof(uint32_t)));
        ctx->h_tokens_cap = n;
    }
    if (ctx->h_seg_cap < nseg) {
        free(ctx->h_ntok); free(ctx->h_btype); free(ctx->h_out_bytes); free(ctx->h_cost); free(ctx->h_block_bits); free(ctx->h_out_off);
        ctx->h_ntok = (uint32_t *)malloc(nseg * 4);
        ctx->h_cost = (uint32_t *)malloc(nseg * 8);
        ctx->h_btype = (uint32_t *)malloc(nseg * 4);
        ctx->h_out_bytes = (uint32_t *)malloc(nseg * 4);
        ctx->h_block_bits = (unsigned long long *)malloc(nseg * 8);
        ctx->h_out_off = (unsigned long long *)malloc(nseg * 8);
        if (!ctx->h_ntok || !ctx->h_btype || !ctx->h_out_bytes || !ctx->h_cost || !ctx->h_block_bits || !ctx->h_out_off) return DFL_E_MALLOC;
        ctx->h_seg_cap = nseg;
    }
    return DFL_OK;
}

/* The five kernels of one chunk (segments [s0, s0 + ns) of a stream of nseg segments), on stream st. */
static int launch_chunk(dfl_ctx *ctx, const uint8_t *d_in, size_t n, size_t hist_len, int is_final, size_t s0, int ns,
                        size_t nseg, uint8_t *d_out, size_t out_cap, cudaStream_t st)
{
    const dfl_params prm = ctx->prm;
    if (ctx->scratch_persist) {
        cudaStreamAttrValue av;
        memset(&av, 0, sizeof(av));
        av.accessPolicyWindow.base_ptr = ctx->scratch;
        av.accessPolicyWindow.num_bytes = ctx->scratch_persist;
        av.accessPolicyWindow.hitRatio = 1.0f;
        av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        if (cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &av) != cudaSuccess) cudaGetLastError();
    }
    /* the default geometry (the window of a segment is the segment before it): runs of consecutive segments per
     * CTA, the window's index inherited (k_match_run); anything else: one CTA per segment (k_match) */
    if (!ctx->match_mode && prm.seg_bytes == DFL_MAX_SEG && prm.window == DFL_MAX_WINDOW && prm.use_dict && !prm.chain_limit) {
        /* Segments per run.  A run costs r segments plus one more index build (its first segment sorts its window
         * itself: ~0.22 of a segment's time), and the runs go round the SMs in waves: take the r whose waves x run cost
         * is least, and among the ones within 1.5 % of it the shortest (more runs: the counter evens out SMs that
         * were placed late or drew slower data). */
        int run_segs = 1;
        {
            double best = 1e30, cost[RUN_SEGS_MAX + 1];
            for (int r = 1; r <= RUN_SEGS_MAX; r++) {
                const int nr = (ns + r - 1) / r;
                cost[r] = (double)((nr + ctx->num_sms - 1) / ctx->num_sms) * ((r < ns ? r : ns) + 0.22);
                if (cost[r] < best) best = cost[r];
            }
            for (int r = RUN_SEGS_MAX; r >= 1; r--)
                if (cost[r] <= 1.015 * best) run_segs = r;
        }
        const int nruns = (ns + run_segs - 1) / run_segs;
        LAUNCH(K_MATCH, k_match_run, nruns < ctx->num_sms ? nruns : ctx->num_sms, MATCH_THREADS, sizeof(MatchSmem), st, d_in, (long long)n,
               (long long)hist_len, (long long)s0, ns, run_segs, ctx->tokcand, ctx->scratch, ctx->state);
    } else if (prm.chain_limit)
        LAUNCH(K_MATCH, k_match<true>, ns, MATCH_THREADS, sizeof(MatchSmem), st, d_in, (long long)n, (long long)hist_len,
               (long long)s0, prm, ctx->tokcand, ctx->scratch, ctx->state);
    else
        LAUNCH(K_MATCH, k_match<false>, ns, MATCH_THREADS, sizeof(MatchSmem), st, d_in, (long long)n, (long long)hist_len,
               (long long)s0, prm, ctx->tokcand, ctx->scratch, ctx->state);
    /* few segments: a CTA per segment (short dependent chain); enough segments to fill every warp slot
     * of the chip: a warp per segment (fewer instructions per token) */
    const int per_warp = ctx->parse_mode ? ctx->parse_mode == 1 : ns >= PARSE_WARP_MIN_SEGS;
    if (per_warp)
        LAUNCH(K_PARSE, k_parse_warp, (ns + PARSE_WARPS - 1) / PARSE_WARPS, PARSE_WARPS * 32, 0, st, (long long)n,
               (long long)s0, ns, prm, ctx->tokcand, ctx->tokens, ctx->A);
    else
        LAUNCH(K_PARSE, k_parse, ns, PARSE_THREADS, sizeof(ParseSmem), st, (long long)n, (long long)s0, prm, ctx->tokcand,
               ctx->tokens, ctx->A);
    SplitArrays X;
    memset(&X, 0, sizeof(X));
    if (prm.block_mode == 4) {
        X = ctx->X;
        LAUNCH(K_HUFF, k_huff_split, (ns + SPLIT_WARPS - 1) / SPLIT_WARPS, SPLIT_WARPS * 32, 0, st, (long long)n,
               (long long)s0, ns, prm, ctx->tokens, ctx->A, X);
    } else
        LAUNCH(K_HUFF, k_huff, (ns + HUFF_WARPS - 1) / HUFF_WARPS, HUFF_WARPS * 32, 0, st, (long long)n,
               (long long)s0, ns, prm, ctx->A);
    LAUNCH(K_SCAN, k_scan, 1, SCAN_THREADS, 0, st, ns, ctx->A, ctx->state);
    /* segment index: where each segment starts in the body, and (entry nseg) where the body ends so far */
    CK(cudaMemcpyAsync(ctx->seg_index + s0, ctx->A.out_off, (size_t)ns * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, st));
    CK(cudaMemcpyAsync(ctx->seg_index + s0 + ns, &ctx->state->total_out, sizeof(unsigned long long), cudaMemcpyDeviceToDevice, st));
    LAUNCH(K_EMIT, k_emit, ns, EMIT_THREADS, 0, st, d_in, (long long)n, (long long)s0, (long long)(nseg - 1),
           is_final, prm, ctx->tokens, ctx->A, X, d_out, (unsigned long long)out_cap, ctx->state);
    return DFL_OK;
}

extern "C" int dfl_b200_deflate_enqueue(dfl_ctx *ctx, const uint8_t *d_in, size_t n, size_t hist_len,
                                        int is_final, uint8_t *d_out, size_t out_cap, void *stream)
{
    if (!ctx || (!d_in && n) || !d_out) return DFL_E_INVAL;
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaSetDevice(ctx->device));
    const dfl_params prm = ctx->prm;
    const size_t nseg = n ? (n + prm.seg_bytes - 1) / prm.seg_bytes : 0;
    size_t chunk_segs = DFL_CHUNK_BYTES / (size_t)prm.seg_bytes;
    if (chunk_segs > nseg) chunk_segs = nseg;
    if (chunk_segs < 1) chunk_segs = 1;
    int rc = ensure_ws(ctx, chunk_segs);
    if (rc) return rc;
    rc = ensure_index(ctx, nseg);
    if (rc) return rc;
    ctx->ev_n = 0;
    ctx->last_n = (long long)n;
    ctx->nseg_kept = 0;
    if (ctx->keep) {
        rc = keep_reserve(ctx, n, nseg);
        if (rc) return rc;
    }
    CK(cudaMemsetAsync(ctx->state, 0, sizeof(DevState), st));
    if (n == 0) {
        /* empty input: a lone empty stored block */
        static const uint8_t empty_final[5] = {1, 0, 0, 0xFF, 0xFF}, empty_more[5] = {0, 0, 0, 0xFF, 0xFF};
        if (out_cap < 5) return DFL_E_SZ;
        CK(cudaMemcpyAsync(d_out, is_final ? empty_final : empty_more, 5, cudaMemcpyHostToDevice, st));
        unsigned long long five = 5;
        CK(cudaMemcpyAsync(&ctx->state->total_out, &five, 8, cudaMemcpyHostToDevice, st));
    }
    for (size_t s0 = 0; s0 < nseg; s0 += chunk_segs) {
        int ns = (int)(nseg - s0 < chunk_segs ? nseg - s0 : chunk_segs);
        rc = launch_chunk(ctx, d_in, n, hist_len, is_final, s0, ns, nseg, d_out, out_cap, st);
        if (rc) return rc;
        if (ctx->keep) {
            /* debug / fd_stats path: pull this chunk's tokens and segment records */
            size_t pos0 = s0 * (size_t)prm.seg_bytes;
            size_t cnt = (size_t)ns * prm.seg_bytes;
            if (pos0 + cnt > n) cnt = n - pos0;
            CK(cudaMemcpyAsync(ctx->h_tokens + pos0, ctx->tokens, cnt * 4, cudaMemcpyDeviceToHost, st));
            CK(cudaMemcpyAsync(ctx->h_ntok + s0, ctx->A.ntok, ns * 4, cudaMemcpyDeviceToHost, st));
            CK(cudaMemcpyAsync(ctx->h_btype + s0, ctx->A.btype, ns * 4, cudaMemcpyDeviceToHost, st));
            CK(cudaMemcpyAsync(ctx->h_out_bytes + s0, ctx->A.out_bytes, ns * 4, cudaMemcpyDeviceToHost, st));
            CK(cudaMemcpyAsync(ctx->h_cost + 2 * s0, ctx->A.cost, ns * 8, cudaMemcpyDeviceToHost, st));
            CK(cudaMemcpyAsync(ctx->h_block_bits + s0, ctx->A.block_bits, ns * 8, cudaMemcpyDeviceToHost, st));
            CK(cudaMemcpyAsync(ctx->h_out_off + s0, ctx->A.out_off, ns * 8, cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            ctx->nseg_kept = (long)(s0 + ns);
        }
    }
    LAUNCH(K_FINAL, k_final, 1, 32, 0, st, (long long)n, ctx->state);
    CK(cudaGetLastError());
    ctx->pending = 1;
    return DFL_OK;
}

extern "C" int dfl_b200_deflate_finish(dfl_ctx *ctx, size_t *out_len, uint32_t *adler, void *stream)
{
    if (!ctx || !ctx->pending) return DFL_E_INVAL;
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaMemcpyAsync(ctx->h_state, ctx->state, sizeof(DevState), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    ctx->pending = 0;
    if (ctx->timing) {
        for (int k = 0; k < K_NUM; k++) ctx->k_ms[k] = 0;
        for (int i = 0; i + 1 < ctx->ev_n; i += 2) {
            float ms = 0;
            CK(cudaEventElapsedTime(&ms, ctx->ev[i], ctx->ev[i + 1]));
            ctx->k_ms[ctx->ev_class[i]] += ms;
        }
    }
    if (ctx->h_state->error) return ctx->h_state->error == 1 ? DFL_E_SZ : DFL_E_CUDA; /* 1: the output did not fit */
    if (out_len) *out_len = (size_t)ctx->h_state->total_out;
    if (adler) *adler = ctx->h_state->adler;
    return DFL_OK;
}

extern "C" int dfl_b200_deflate_device(dfl_ctx *ctx, const uint8_t *d_in, size_t n, size_t hist_len,
                                       int is_final, uint8_t *d_out, size_t out_cap, size_t *out_len,
                                       uint32_t *adler, void *stream)
{
    int rc = dfl_b200_deflate_enqueue(ctx, d_in, n, hist_len, is_final, d_out, out_cap, stream);
    if (rc) return rc;
    return dfl_b200_deflate_finish(ctx, out_len, adler, stream);
}

#ifndef HOST_PIPE_BYTES
#define HOST_PIPE_BYTES (64u << 20) /* piece size of the host path: H2D, compute and D2H of neighbouring pieces overlap */
#endif

static int host_pipe_init(dfl_ctx *ctx, int npieces)
{
    if (!ctx->s_in) {
        CK(cudaStreamCreateWithFlags(&ctx->s_in, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&ctx->s_run, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&ctx->s_out, cudaStreamNonBlocking));
    }
    if (ctx->pipe_cap < npieces) {
        for (int i = 0; i < ctx->pipe_cap; i++) { cudaEventDestroy(ctx->e_in[i]); cudaEventDestroy(ctx->e_done[i]); }
        free(ctx->e_in); free(ctx->e_done);
        if (ctx->h_totals) cudaFreeHost(ctx->h_totals);
        ctx->pipe_cap = 0;
        ctx->e_in = (cudaEvent_t *)calloc(npieces, sizeof(cudaEvent_t));
        ctx->e_done = (cudaEvent_t *)calloc(npieces, sizeof(cudaEvent_t));
        if (!ctx->e_in || !ctx->e_done) return DFL_E_MALLOC;
        CK(cudaMallocHost(&ctx->h_totals, (size_t)npieces * sizeof(unsigned long long)));
        for (int i = 0; i < npieces; i++) {
            CK(cudaEventCreateWithFlags(&ctx->e_in[i], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&ctx->e_done[i], cudaEventDisableTiming));
        }
        ctx->pipe_cap = npieces;
    }
    return DFL_OK;
}

extern "C" int dfl_b200_compress_host(dfl_ctx *ctx, const uint8_t *in, size_t n, uint8_t *out, size_t cap,
                                      size_t *out_len)
{
    if (!ctx || (!in && n) || !out || !out_len) return DFL_E_INVAL;
    CK(cudaSetDevice(ctx->device));
    const dfl_params prm = ctx->prm;
    size_t bound = dfl_b200_bound(&prm, n);
    { const int rc_ = grow_dev(ctx, &ctx->d_in, &ctx->d_in_cap, n + 16); if (rc_) return rc_; }
    { const int rc_ = grow_dev(ctx, &ctx->d_out, &ctx->d_out_cap, bound + 16); if (rc_) return rc_; }
    uint8_t *d_body = ctx->d_out + 2; /* body at +2 so the zlib header can be prepended without a second copy */
    size_t body = 0;
    uint32_t adler = 1;
    const size_t nseg = n ? (n + prm.seg_bytes - 1) / prm.seg_bytes : 0;
    size_t piece_segs = HOST_PIPE_BYTES / (size_t)prm.seg_bytes;
    if (piece_segs < 1) piece_segs = 1;
    /* piece boundaries (in segments): a quarter piece first so that the kernels start early, then full pieces
     * (measured on 256 MiB: 64 MiB pieces 16.0 GB/s, 32 MiB 15.6, 16 MiB 13.4; an extra small last piece 15.8) */
    const size_t first_segs = piece_segs / 4 ? piece_segs / 4 : 1;
    std::vector<size_t> bnd(1, 0);
    if (nseg) {
        size_t b = first_segs < nseg ? first_segs : nseg;
        bnd.push_back(b);
        while (nseg - b > piece_segs) { b += piece_segs; bnd.push_back(b); }
        if (b < nseg) bnd.push_back(nseg);
    }
    const int npieces = (int)bnd.size() - 1;
#define PIECE_S0(i) (bnd[(i)])
#define PIECE_S1(i) (bnd[(i) + 1])
    if (ctx->keep || npieces < 2) {
        /* small input, or the token-keeping debug path: one copy in, one call, one copy out */
        cudaStream_t st = 0;
        if (n) CK(cudaMemcpyAsync(ctx->d_in, in, n, cudaMemcpyHostToDevice, st));
        int rc = dfl_b200_deflate_device(ctx, ctx->d_in, n, 0, 1, d_body, bound, &body, &adler, st);
        if (rc) return rc;
        if (cap < body + 6) return DFL_E_SZ;
        CK(cudaMemcpyAsync(out + 2, d_body, body, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
    } else {
        /* (every error exit drains the three streams before it returns: copies from `in` and into `out` may be in
         * flight, and the caller is free to release its buffers once the call has failed) */
        const int prc = [&]() -> int {
        /* pipeline over pieces: copy-in of piece i+1, kernels of piece i and copy-out of piece i-1 overlap
         * on three streams; output offsets stay on the device, each piece's end offset comes back with it */
        int rc = ensure_ws(ctx, piece_segs);
        if (rc) return rc;
        rc = ensure_index(ctx, nseg);
        if (rc) return rc;
        ctx->last_n = (long long)n;
        rc = host_pipe_init(ctx, npieces);
        if (rc) return rc;
        ctx->ev_n = 0;
        ctx->nseg_kept = 0;
        CK(cudaMemsetAsync(ctx->state, 0, sizeof(DevState), ctx->s_run));
        for (int i = 0; i < npieces; i++) {
            const size_t lo = PIECE_S0(i) * prm.seg_bytes;
            const size_t hi = (PIECE_S1(i) * prm.seg_bytes < n) ? PIECE_S1(i) * prm.seg_bytes : n;
            CK(cudaMemcpyAsync(ctx->d_in + lo, in + lo, hi - lo, cudaMemcpyHostToDevice, ctx->s_in));
            CK(cudaEventRecord(ctx->e_in[i], ctx->s_in));
        }
        size_t copied = 0;
        for (int i = 0; i <= npieces; i++) {
            if (i < npieces) {
                const size_t s0 = PIECE_S0(i);
                const int ns = (int)(PIECE_S1(i) - s0);
                CK(cudaStreamWaitEvent(ctx->s_run, ctx->e_in[i], 0));
                rc = launch_chunk(ctx, ctx->d_in, n, 0, 1, s0, ns, nseg, d_body, bound, ctx->s_run);
                if (rc) return rc;
                if (i == npieces - 1) {
                    LAUNCH(K_FINAL, k_final, 1, 32, 0, ctx->s_run, (long long)n, ctx->state);
                    CK(cudaMemcpyAsync(ctx->h_state, ctx->state, sizeof(DevState), cudaMemcpyDeviceToHost, ctx->s_run));
                }
                CK(cudaMemcpyAsync(&ctx->h_totals[i], &ctx->state->total_out, sizeof(unsigned long long),
                                   cudaMemcpyDeviceToHost, ctx->s_run));
                CK(cudaEventRecord(ctx->e_done[i], ctx->s_run));
            }
            if (i >= 1) {
                /* piece i-1 is complete: its bytes [copied, h_totals[i-1]) can leave while piece i runs */
                CK(cudaEventSynchronize(ctx->e_done[i - 1]));
                const size_t end = (size_t)ctx->h_totals[i - 1];
                if (cap < end + 6) { /* let everything in flight land before the caller's buffers go away */
                    cudaStreamSynchronize(ctx->s_run);
                    cudaStreamSynchronize(ctx->s_out);
                    return DFL_E_SZ;
                }
                CK(cudaMemcpyAsync(out + 2 + copied, d_body + copied, end - copied, cudaMemcpyDeviceToHost, ctx->s_out));
                copied = end;
            }
        }
        CK(cudaStreamSynchronize(ctx->s_out));
        CK(cudaGetLastError());
        if (ctx->h_state->error) return ctx->h_state->error == 1 ? DFL_E_SZ : DFL_E_CUDA;
        body = (size_t)ctx->h_state->total_out;
        adler = ctx->h_state->adler;
            return DFL_OK;
        }();
        if (prc) {
            cudaStreamSynchronize(ctx->s_in);
            cudaStreamSynchronize(ctx->s_run);
            cudaStreamSynchronize(ctx->s_out);
            cudaGetLastError();
            return prc;
        }
    }
#undef PIECE_S0
#undef PIECE_S1
    out[0] = 0x78; /* CM = 8, CINFO = 7: header the decoder checks at src/deflate_decompress.c:352-360 */
    out[1] = 0x9C;
    out[2 + body + 0] = (uint8_t)(adler >> 24); /* RFC 1950: Adler-32 MSB first */
    out[2 + body + 1] = (uint8_t)(adler >> 16);
    out[2 + body + 2] = (uint8_t)(adler >> 8);
    out[2 + body + 3] = (uint8_t)adler;
    *out_len = body + 6;
    return DFL_OK;
}

extern "C" long dfl_b200_num_segments(dfl_ctx *ctx) { return ctx ? ctx->nseg_kept : 0; }

/* segments of the last call whose index was rebuilt with the ballot-ranked scatter (expected 0) */
extern "C" unsigned int dfl_b200_sort_fallbacks(dfl_ctx *ctx) { return ctx ? ctx->h_state->sort_fallbacks : 0; }

/* k_match SM cycles of the last call, summed over CTAs: 0 stage-in, 1 build, 2 search, 3 total */
extern "C" unsigned long long dfl_b200_phase_cycles(dfl_ctx *ctx, int k)
{
    return (ctx && k >= 0 && k < 24) ? ctx->h_state->t_phase[k] : 0;
}

extern "C" int dfl_b200_seg_info(dfl_ctx *ctx, long seg, dfl_seg_info *info)
{
    if (!ctx || !info || seg < 0 || seg >= ctx->nseg_kept) return DFL_E_RANGE;
    info->btype = (int32_t)(ctx->h_btype[seg] & 0xFF);
    info->nblocks = (int64_t)(ctx->h_btype[seg] >> 8 ? ctx->h_btype[seg] >> 8 : 1);
    info->ntokens = (int32_t)ctx->h_ntok[seg];
    info->block_bits = (int64_t)ctx->h_block_bits[seg];
    info->out_bytes = (int64_t)ctx->h_out_bytes[seg];
    info->out_offset = (int64_t)ctx->h_out_off[seg];
    info->tree_bits = (int64_t)ctx->h_cost[2 * seg];
    info->d_bits = (int64_t)ctx->h_cost[2 * seg + 1];
    info->ll_bits = info->block_bits - 3 * info->nblocks - info->tree_bits - info->d_bits;
    return DFL_OK;
}

extern "C" long dfl_b200_tokens(dfl_ctx *ctx, long seg, uint32_t *tokens, long cap)
{
    if (!ctx || seg < 0 || seg >= ctx->nseg_kept) return DFL_E_RANGE;
    long nt = (long)ctx->h_ntok[seg];
    if (tokens) {
        if (cap < nt) return DFL_E_SZ;
        memcpy(tokens, ctx->h_tokens + (size_t)seg * ctx->prm.seg_bytes, (size_t)nt * 4);
    }
    return nt;
}

extern "C" int dfl_b200_synth_random(uint8_t *d_out, size_t n, uint64_t seed, void *stream)
{
    dfl_ctx *ctx = NULL;
    k_synth_random<<<148 * 8, 256, 0, (cudaStream_t)stream>>>(d_out, (unsigned long long)n, (unsigned long long)seed);
    CK(cudaGetLastError());
    return DFL_OK;
}

/* ------------------------------------------------------------------------- */
/* segment index + inflate                                                   */
/* ------------------------------------------------------------------------- */

extern "C" long dfl_b200_seg_index(dfl_ctx *ctx, uint64_t *offsets, long cap)
{
    if (!ctx) return DFL_E_INVAL;
    const long cnt = (long)ctx->seg_index_n + 1;
    if (!ctx->seg_index) return DFL_E_INVAL;
    if (offsets) {
        if (cap < cnt) return DFL_E_SZ;
        if (cudaSetDevice(ctx->device) != cudaSuccess ||
            cudaMemcpy(offsets, ctx->seg_index, (size_t)cnt * sizeof(uint64_t), cudaMemcpyDeviceToHost) != cudaSuccess)
            return DFL_E_CUDA;
        if (ctx->seg_index_n == 0) offsets[0] = 5; /* empty input: the body is the lone empty stored block */
    }
    return cnt;
}

extern "C" const uint64_t *dfl_b200_seg_index_device(dfl_ctx *ctx) { return ctx ? (const uint64_t *)ctx->seg_index : NULL; }

extern "C" int dfl_b200_inflate_device(dfl_ctx *ctx, const uint8_t *d_body, size_t body_len, const uint64_t *d_seg_off,
                                       size_t nseg, uint32_t seg_bytes, size_t hist_len, uint8_t *d_out, size_t n,
                                       uint32_t *adler, void *stream)
{
    if (!ctx || (!d_body && body_len) || (!d_out && n) || (!d_seg_off && nseg)) return DFL_E_INVAL;
    if (seg_bytes < 1 || seg_bytes > DFL_MAX_SEG) return DFL_E_RANGE;
    if (nseg != (n ? (n + seg_bytes - 1) / seg_bytes : 0)) return DFL_E_INVAL;
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaSetDevice(ctx->device));
    size_t chunk_segs = DFL_CHUNK_BYTES / (size_t)seg_bytes;
    if (chunk_segs > nseg) chunk_segs = nseg;
    if (chunk_segs < 1) chunk_segs = 1;
    const size_t words = chunk_segs * seg_bytes;
    if (ctx->inf_cap < words) {
        if (ctx->inf_w) cudaFree(ctx->inf_w);
        ctx->inf_w = NULL;
        ctx->inf_cap = 0;
        const size_t inf_bytes = words * sizeof(uint32_t) + 16 + words / INF_BLK + 16; /* + one done flag per resolve block */
        CK(cudaMalloc(&ctx->inf_w, inf_bytes));
        CK(cudaMemsetAsync(ctx->inf_w, 0x80, inf_bytes, st)); /* every word starts as a value, never as a pointer */
        ctx->inf_cap = words;
    }
    ctx->ev_n = 0;
    CK(cudaMemsetAsync(ctx->state, 0, sizeof(DevState), st));
    unsigned int *pending = &ctx->state->inf_pending;
    int rounds_total = 0;
    for (size_t s0 = 0; s0 < nseg; s0 += chunk_segs) {
        const int ns = (int)(nseg - s0 < chunk_segs ? nseg - s0 : chunk_segs);
        const long long chunk_lo = (long long)(s0 * seg_bytes);
        const long long cnt = (long long)((s0 + ns) * (size_t)seg_bytes < n ? (size_t)ns * seg_bytes : n - s0 * seg_bytes);
        LAUNCH(K_INFLATE, k_inflate, (ns + INF_WARPS - 1) / INF_WARPS, INF_WARPS * 32, 0, st, d_body, (unsigned long long)body_len,
               (const unsigned long long *)d_seg_off, (long long)s0, ns, (int)seg_bytes, (long long)n, (long long)hist_len,
               ctx->inf_w, ctx->state);
        /* pointer jumping until no word is a pointer any more: every round at least doubles the number of
         * segments a pointer spans, so ceil(log2(ns)) + 2 rounds always suffice */
        int max_rounds = 3;
        for (int v = ns; v > 1; v >>= 1) max_rounds++;
        unsigned char *blk_done = reinterpret_cast<unsigned char *>(ctx->inf_w + ctx->inf_cap + 4);
        CK(cudaMemsetAsync(blk_done, 0, (size_t)((cnt + INF_BLK - 1) / INF_BLK), st));
        for (int r = 0;; r++) {
            CK(cudaMemsetAsync(pending, 0, sizeof(unsigned int), st));
            LAUNCH(K_INF_RESOLVE, k_inf_resolve, 148 * 16, 256, 0, st, ctx->inf_w, cnt, chunk_lo, (long long)hist_len,
                   (const uint8_t *)d_out, blk_done, pending, &ctx->state->error);
            unsigned int left = 0;
            CK(cudaMemcpyAsync(&ctx->h_state->inf_pending, pending, sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            left = ctx->h_state->inf_pending;
            rounds_total++;
            if (!left) break;
            if (r >= max_rounds) return DFL_E_DATA; /* a pointer cycle: not a stream this library wrote */
        }
        LAUNCH(K_INF_FINISH, k_inf_finish, ns, 256, 0, st, (const uint32_t *)ctx->inf_w, (long long)s0, (int)seg_bytes, (long long)n,
               d_out, ctx->state);
    }
    LAUNCH(K_FINAL, k_final, 1, 32, 0, st, (long long)n, ctx->state);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(ctx->h_state, ctx->state, sizeof(DevState), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    ctx->inf_rounds = rounds_total;
    if (ctx->timing) {
        for (int k = 0; k < K_NUM; k++) ctx->k_ms[k] = 0;
        for (int i = 0; i + 1 < ctx->ev_n; i += 2) {
            float ms = 0;
            CK(cudaEventElapsedTime(&ms, ctx->ev[i], ctx->ev[i + 1]));
            ctx->k_ms[ctx->ev_class[i]] += ms;
        }
    }
    if (ctx->h_state->error) return DFL_E_DATA;
    if (adler) *adler = ctx->h_state->adler;
    return DFL_OK;
}

/* pointer-jumping rounds of the last inflate call (profiling aid) */
extern "C" int dfl_b200_inflate_rounds(dfl_ctx *ctx) { return ctx ? ctx->inf_rounds : 0; }

extern "C" int dfl_b200_decompress_host(dfl_ctx *ctx, const uint8_t *in, size_t in_len, const uint64_t *seg_off, size_t nseg,
                                        uint32_t seg_bytes, uint8_t *out, size_t n)
{
    if (!ctx || !in || in_len < 6 || (!out && n) || (!seg_off && nseg)) return DFL_E_INVAL;
    /* zlib wrapper: the checks of src/deflate_decompress.c:347-368 (CM = 8, window <= 32 KiB, FCHECK, no preset dictionary) */
    if ((in[0] & 0x0F) != 8 || (in[0] >> 4) > 7 || ((in[0] << 8) | in[1]) % 31 != 0 || (in[1] & 0x20)) return DFL_E_DATA;
    CK(cudaSetDevice(ctx->device));
    const size_t body_len = in_len - 6;
    { const int rc_ = grow_dev(ctx, &ctx->d_in, &ctx->d_in_cap, in_len + 16); if (rc_) return rc_; }
    { const int rc_ = grow_dev(ctx, &ctx->d_out, &ctx->d_out_cap, n + 16); if (rc_) return rc_; }
    uint64_t *d_off = NULL;
    CK(cudaMalloc(&d_off, (nseg + 1) * sizeof(uint64_t)));
    cudaStream_t st = 0;
    uint32_t adler = 1;
    int rc = DFL_OK;
    if (cudaMemcpyAsync(ctx->d_in, in, in_len, cudaMemcpyHostToDevice, st) != cudaSuccess ||
        cudaMemcpyAsync(d_off, seg_off, (nseg + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, st) != cudaSuccess)
        rc = DFL_E_CUDA;
    if (!rc) rc = dfl_b200_inflate_device(ctx, ctx->d_in + 2, body_len, d_off, nseg, seg_bytes, 0, ctx->d_out, n, &adler, st);
    if (!rc && n && cudaMemcpy(out, ctx->d_out, n, cudaMemcpyDeviceToHost) != cudaSuccess) rc = DFL_E_CUDA;
    cudaFree(d_off);
    if (rc) return rc;
    const uint32_t want = ((uint32_t)in[in_len - 4] << 24) | ((uint32_t)in[in_len - 3] << 16) | ((uint32_t)in[in_len - 2] << 8) | in[in_len - 1];
    return want == adler ? DFL_OK : DFL_E_DATA; /* RFC 1950: Adler-32 of the uncompressed data, MSB first */
}

/* ------------------------------------------------------------------------- */
/* PNG: scanline filter + container                                          */
/* ------------------------------------------------------------------------- */

extern "C" int dfl_b200_png_filter_device(dfl_ctx *ctx, const uint8_t *d_pix, size_t row_bytes, size_t rows, int bpp, int mode,
                                          uint8_t *d_out, void *stream)
{
    if (!ctx || ((!d_pix || !d_out) && rows && row_bytes)) return DFL_E_INVAL;
    if (bpp < 1 || bpp > 8 || mode < 0 || mode > 5) return DFL_E_RANGE;
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaSetDevice(ctx->device));
    ctx->ev_n = 0;
    if (rows && row_bytes) {
        const unsigned int grid = (unsigned int)((rows + PNGF_WARPS - 1) / PNGF_WARPS);
        if (bpp <= 4 && row_bytes % 4 == 0 && ((uintptr_t)d_pix & 3) == 0)
            LAUNCH(K_PNG_FILTER, k_png_filter4, grid, PNGF_WARPS * 32, 0, st, d_pix, (long long)row_bytes, (long long)rows, bpp, mode, d_out);
        else
            LAUNCH(K_PNG_FILTER, k_png_filter, grid, PNGF_WARPS * 32, 0, st, d_pix, (long long)row_bytes, (long long)rows, bpp, mode, d_out);
    }
    CK(cudaGetLastError());
    if (ctx->timing) {
        CK(cudaStreamSynchronize(st));
        for (int k = 0; k < K_NUM; k++) ctx->k_ms[k] = 0;
        for (int i = 0; i + 1 < ctx->ev_n; i += 2) {
            float ms = 0;
            CK(cudaEventElapsedTime(&ms, ctx->ev[i], ctx->ev[i + 1]));
            ctx->k_ms[ctx->ev_class[i]] += ms;
        }
    }
    return DFL_OK;
}

/* CRC-32 of a PNG chunk (calc_crc, src/include/crc.h:22-30) */
static uint32_t png_crc(const uint8_t *p, size_t n)
{
    static uint32_t table[256];
    static int ready = 0;
    if (!ready) {
        for (uint32_t i = 0; i < 256; i++) {
            uint32_t c = i;
            for (int k = 0; k < 8; k++) c = (c >> 1) ^ (0xEDB88320u & (0u - (c & 1u)));
            table[i] = c;
        }
        ready = 1;
    }
    uint32_t c = 0xFFFFFFFFu;
    for (size_t i = 0; i < n; i++) c = table[(c ^ p[i]) & 0xFF] ^ (c >> 8);
    return ~c;
}

static void put_be32(uint8_t *p, uint32_t v) { p[0] = (uint8_t)(v >> 24); p[1] = (uint8_t)(v >> 16); p[2] = (uint8_t)(v >> 8); p[3] = (uint8_t)v; }

extern "C" size_t dfl_b200_png_bound(const dfl_params *prm, uint32_t width, uint32_t height, uint32_t channels)
{
    const size_t filtered = (size_t)height * ((size_t)width * channels + 1);
    return dfl_b200_bound(prm, filtered) + 6 + 8 + 25 + 12 + 12;
}

/* 8-bit grey (1), grey + alpha (2), RGB (3) or RGBA (4) pixels -> a PNG file: signature, IHDR, one IDAT
 * (filter on the GPU, compress on the GPU, zlib wrapper), IEND -- what src/png/png_decoder.c reads. */
extern "C" int dfl_b200_png_encode_host(dfl_ctx *ctx, const uint8_t *pix, uint32_t width, uint32_t height, uint32_t channels,
                                        int filter_mode, uint8_t *out, size_t cap, size_t *out_len)
{
    static const uint8_t sig[8] = {0x89, 0x50, 0x4E, 0x47, 0x0D, 0x0A, 0x1A, 0x0A}; /* PNG_HEAD, src/png/include/png_decoder.h:63 */
    static const uint8_t ctype[5] = {0, 0, 4, 2, 6};
    if (!ctx || !pix || !out || !out_len) return DFL_E_INVAL;
    if (!width || !height || channels < 1 || channels > 4) return DFL_E_RANGE;
    const size_t row_bytes = (size_t)width * channels, npix = row_bytes * height, nf = (size_t)height * (row_bytes + 1);
    if (nf + 64 > 0x7FFFFFFFull) return DFL_E_RANGE; /* one IDAT chunk */
    CK(cudaSetDevice(ctx->device));
    const dfl_params prm = ctx->prm;
    const size_t bound = dfl_b200_bound(&prm, nf);
    if (cap < 8 + 25 + 12 + 12) return DFL_E_SZ;
    { const int rc_ = grow_dev(ctx, &ctx->d_in, &ctx->d_in_cap, npix + 16); if (rc_) return rc_; }
    { const int rc_ = grow_dev(ctx, &ctx->d_png, &ctx->d_png_cap, nf + 16); if (rc_) return rc_; }
    { const int rc_ = grow_dev(ctx, &ctx->d_out, &ctx->d_out_cap, bound + 16); if (rc_) return rc_; }
    cudaStream_t st = 0;
    CK(cudaMemcpyAsync(ctx->d_in, pix, npix, cudaMemcpyHostToDevice, st));
    int rc = dfl_b200_png_filter_device(ctx, ctx->d_in, row_bytes, height, (int)channels, filter_mode, ctx->d_png, st);
    if (rc) return rc;
    size_t body = 0;
    uint32_t adler = 1;
    rc = dfl_b200_deflate_device(ctx, ctx->d_png, nf, 0, 1, ctx->d_out, bound, &body, &adler, st);
    if (rc) return rc;
    const size_t idat = 2 + body + 4, total = 8 + 25 + 12 + idat + 12;
    if (cap < total) return DFL_E_SZ;
    uint8_t *p = out;
    memcpy(p, sig, 8); p += 8;
    put_be32(p, 13); memcpy(p + 4, "IHDR", 4);
    put_be32(p + 8, width); put_be32(p + 12, height);
    p[16] = 8; p[17] = ctype[channels]; p[18] = 0; p[19] = 0; p[20] = 0; /* depth, colour type, deflate, adaptive filtering, no interlace */
    put_be32(p + 21, png_crc(p + 4, 17)); p += 25;
    put_be32(p, (uint32_t)idat); memcpy(p + 4, "IDAT", 4);
    p[8] = 0x78; p[9] = 0x9C;
    CK(cudaMemcpy(p + 10, ctx->d_out, body, cudaMemcpyDeviceToHost));
    put_be32(p + 10 + body, adler);
    put_be32(p + 8 + idat, png_crc(p + 4, 4 + idat)); p += 12 + idat;
    put_be32(p, 0); memcpy(p + 4, "IEND", 4); put_be32(p + 8, png_crc(p + 4, 4));
    *out_len = total;
    return DFL_OK;
}

#include "dfl_pipeline.cuh"
