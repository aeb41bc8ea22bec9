/*
 * host/deflate_compress.c -- the reference's compress entry points, in C, over
 * the CUDA C-ABI (include/deflate_b200.h).  Replaces src/deflate_compress.c:
 * 83-112 (ctor/dtor) and :362-376 (deflate_compress) of
 * mparker97/deflate_compression; process_loop (:219-347) itself runs on the GPU.
 * No CPU fallback: without a B200 the calls fail with -E_GPU.
 */
#define _GNU_SOURCE
#include <errno.h>
#include <limits.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "deflate_b200.h"
#include "deflate_errors.h"
#include "deflate_ext.h"

struct deflate_compr { /* opaque to callers: src/include/deflate_ext.h:10 */
    dfl_ctx *ctx;
    int fd_in, fd_out, fd_stats;
    swi sliding_window;
    int err;
};

static int map_err(int rc)
{
    switch (rc) {
    case DFL_OK: return 0;
    case DFL_E_MALLOC: return -E_MALLOC;
    case DFL_E_SZ: return -E_SZ;
    case DFL_E_RANGE: return -E_RANGE;
    case DFL_E_INVAL: return -E_INVAL;
    default: return -(int)E_GPU;
    }
}

/* A context is expensive to make (device workspace, pinned pipeline pieces), deflate_compress() makes one per call
 * (src/deflate_compress.c:364-366 does): the last one is parked here and taken again by the next call with the same
 * window instead of being destroyed; it is released when the process exits. */
static pthread_mutex_t parked_mu = PTHREAD_MUTEX_INITIALIZER;
static dfl_ctx *parked_ctx;
static int parked_window, parked_device, parked_atexit;

/* The reference's entry points have no device argument: DFL_B200_DEVICE (default 0) picks the GPU. */
static int env_device(void)
{
    const char *e = getenv("DFL_B200_DEVICE"), *m = getenv("DFL_B200_DEVICES");
    if (m && *m) return atoi(m); /* a device list: its first entry */
    return e && *e ? atoi(e) : 0;
}

/* Nor have they a knob for the blocks: DFL_B200_SPLIT=1 turns adaptive block splitting on (block_mode 4: a 32 KiB
 * segment becomes up to 8 blocks where that saves bits -- the reference's stated aim, README.md:15-20). */
static int env_block_mode(void)
{
    const char *e = getenv("DFL_B200_SPLIT");
    return e && *e && atoi(e) ? 4 : 0;
}

/* DFL_B200_DEVICES=0,1,2,...: more than one device puts every call without a record stream on all of them
 * (contiguous shards, one zlib stream: dfl_b200_compress_fd_multi).  The extra contexts stay parked like the first. */
#define MULTI_MAX 16
static dfl_ctx *multi_ctx[MULTI_MAX]; /* contexts on devices 1.. of the list ([0] is the call's own) */
static int multi_dev[MULTI_MAX], multi_n;

static int env_devices(int *devs)
{
    const char *e = getenv("DFL_B200_DEVICES");
    int k = 0;
    while (e && *e && k < MULTI_MAX) {
        char *end;
        const long v = strtol(e, &end, 10);
        if (end == e) break;
        devs[k++] = (int)v;
        e = *end == ',' ? end + 1 : end;
        if (*end != ',') break;
    }
    return k;
}

static void parked_release(void)
{
    int i;
    pthread_mutex_lock(&parked_mu);
    dfl_b200_destroy(parked_ctx);
    parked_ctx = NULL;
    for (i = 0; i < MULTI_MAX; i++) { dfl_b200_destroy(multi_ctx[i]); multi_ctx[i] = NULL; }
    multi_n = 0;
    pthread_mutex_unlock(&parked_mu);
}

deflate_compr_t *spawn_deflate_compr_t(void)
{
    return (deflate_compr_t *)calloc(1, sizeof(deflate_compr_t));
}

/* src/deflate_compress.c:85-104.  Errors are latched in the object (the
 * reference longjmps out of a void function) and reported by deflate_compress. */
void deflate_compr_init(deflate_compr_t *com, int fd_in, int fd_out, int fd_stats, swi sw)
{
    dfl_params prm;
    if (!com) return;
    com->fd_in = fd_in;
    com->fd_out = fd_out;
    com->fd_stats = fd_stats;
    com->sliding_window = sw;
    com->ctx = NULL;
    com->err = 0;
    if (sw < 1 || sw > DFL_MAX_WINDOW) { com->err = -E_RANGE; return; }
    pthread_mutex_lock(&parked_mu);
    if (parked_ctx && parked_window == (int)sw && parked_device == env_device()) {
        com->ctx = parked_ctx;
        parked_ctx = NULL;
    }
    pthread_mutex_unlock(&parked_mu);
    dfl_b200_default_params(&prm);
    prm.window = sw;
    prm.block_mode = env_block_mode();
    if (com->ctx) { /* (the parked context may have been used with the other block mode) */
        com->err = map_err(dfl_b200_set_params(com->ctx, &prm));
        return;
    }
    com->err = map_err(dfl_b200_create(&com->ctx, env_device(), &prm));
}

void deflate_compr_deinit(deflate_compr_t *com)
{
    dfl_ctx *old = NULL;
    if (!com || !com->ctx) return;
    dfl_b200_keep_tokens(com->ctx, 0);
    pthread_mutex_lock(&parked_mu);
    old = parked_ctx;
    parked_ctx = com->ctx;
    parked_window = (int)com->sliding_window;
    parked_device = env_device();
    if (!parked_atexit) { parked_atexit = 1; atexit(parked_release); }
    pthread_mutex_unlock(&parked_mu);
    dfl_b200_destroy(old);
    com->ctx = NULL;
}

static int write_all(int fd, const void *buf, size_t n)
{
    const unsigned char *p = (const unsigned char *)buf;
    while (n) {
        ssize_t w = write(fd, p, n);
        if (w < 0) {
            if (errno == EINTR) continue;
            return -E_INVAL;
        }
        p += w;
        n -= (size_t)w;
    }
    return 0;
}

/* fetch(), src/deflate_compress.c:139-145: read(2) until it returns 0. */
static int read_all(int fd, unsigned char **pbuf, size_t *pn)
{
    size_t cap = 1u << 20, n = 0;
    unsigned char *buf = (unsigned char *)malloc(cap);
    if (!buf) return -E_MALLOC;
    for (;;) {
        ssize_t r;
        if (n == cap) {
            unsigned char *nb = (unsigned char *)realloc(buf, cap * 2);
            if (!nb) { free(buf); return -E_MALLOC; }
            buf = nb;
            cap *= 2;
        }
        r = read(fd, buf + n, cap - n);
        if (r < 0) {
            if (errno == EINTR) continue;
            free(buf);
            return -E_INVAL;
        }
        if (r == 0) break;
        n += (size_t)r;
    }
    *pbuf = buf;
    *pn = n;
    return 0;
}

/* One struct compress_stats per token (src/deflate_compress.c:291-309).  ll / d are the token; bytes counts the
 * input consumed so far (+ 1, :235); the three cost fields are those of the block the token belongs to, so that
 * the reference's rate = (tree_bits + ll_bits + d_bits) / bytes (src/include/deflate_ext.h:29-30) read at the last
 * record of a block is that block's real cost (less its 3 header bits). */
static int write_stats(deflate_compr_t *com)
{
    long nseg = dfl_b200_num_segments(com->ctx), s;
    uint32_t *tok = (uint32_t *)malloc(sizeof(uint32_t) * DFL_MAX_SEG);
    struct compress_stats *recs = (struct compress_stats *)malloc(sizeof(struct compress_stats) * DFL_MAX_SEG);
    int bytes = 1, rc = 0; /* cs.bytes = 1, src/deflate_compress.c:235 */
    if (!tok || !recs) { free(tok); free(recs); return -E_MALLOC; }
    for (s = 0; s < nseg && !rc; s++) {
        dfl_seg_info info;
        long nt = dfl_b200_tokens(com->ctx, s, tok, DFL_MAX_SEG), k;
        if (nt < 0 || dfl_b200_seg_info(com->ctx, s, &info)) { rc = -E_INVAL; break; }
        for (k = 0; k < nt; k++) {
            uint32_t t = tok[k];
            recs[k].bytes = bytes;
            recs[k].tree_bits = (int)info.tree_bits;
            recs[k].ll_bits = (int)info.ll_bits;
            recs[k].d_bits = (int)info.d_bits;
            if ((t >> 16) == 0) { recs[k].ll = (int)(t & 0xFF); recs[k].d = 0; bytes += 1; }
            else { recs[k].ll = (int)(t & 0xFFFF); recs[k].d = (int)(t >> 16); bytes += (int)(t & 0xFFFF); }
        }
        rc = write_all(com->fd_stats, recs, sizeof(struct compress_stats) * (size_t)nt);
    }
    free(tok);
    free(recs);
    return rc;
}

/* The body of deflate_compress once the object exists (process_loop's role). */
static int deflate_compr_run(deflate_compr_t *com)
{
    unsigned char *in = NULL, *out = NULL;
    size_t n = 0, cap, outn = 0;
    int rc;
    dfl_params prm;
    if (com->err) return com->err;
    if (com->fd_stats < 0) {
        /* no record stream wanted.  Several devices named: the whole input, sharded over them */
        int devs[MULTI_MAX];
        const int nd = env_devices(devs);
        if (nd > 1) {
            dfl_ctx *ctxs[MULTI_MAX];
            int i;
            dfl_b200_default_params(&prm);
            prm.window = com->sliding_window;
            prm.block_mode = env_block_mode();
            ctxs[0] = com->ctx; /* made on the first device of the list (env_device) */
            rc = 0;
            pthread_mutex_lock(&parked_mu);
            for (i = 1; i < nd && !rc; i++) {
                if (multi_ctx[i] && multi_dev[i] != devs[i]) { dfl_b200_destroy(multi_ctx[i]); multi_ctx[i] = NULL; }
                if (!multi_ctx[i]) rc = map_err(dfl_b200_create(&multi_ctx[i], devs[i], &prm));
                else rc = map_err(dfl_b200_set_params(multi_ctx[i], &prm));
                multi_dev[i] = devs[i];
                ctxs[i] = multi_ctx[i];
            }
            if (!parked_atexit) { parked_atexit = 1; atexit(parked_release); }
            if (!rc) rc = map_err(dfl_b200_compress_fd_multi(ctxs, nd, com->fd_in, com->fd_out, NULL, NULL));
            pthread_mutex_unlock(&parked_mu);
            return rc;
        }
        /* one device: the pipelined fd -> fd path (reads, copies, kernels and writes overlap) */
        return map_err(dfl_b200_compress_fd(com->ctx, com->fd_in, com->fd_out, NULL, NULL));
    }
    rc = read_all(com->fd_in, &in, &n);
    if (rc) return rc;
    /* compress_stats.bytes is an int (src/include/deflate_ext.h:20): the record stream cannot describe more */
    if (com->fd_stats >= 0 && n >= (size_t)INT_MAX) { free(in); return -E_RANGE; }
    dfl_b200_default_params(&prm);
    prm.window = com->sliding_window;
    prm.block_mode = env_block_mode();
    cap = dfl_b200_bound(&prm, n) + 16;
    out = (unsigned char *)malloc(cap);
    if (!out) { free(in); return -E_MALLOC; }
    if (com->fd_stats >= 0) dfl_b200_keep_tokens(com->ctx, 1);
    rc = map_err(dfl_b200_compress_host(com->ctx, in, n, out, cap, &outn));
    if (!rc && com->fd_out >= 0) rc = write_all(com->fd_out, out, outn);
    if (!rc && com->fd_stats >= 0) rc = write_stats(com);
    free(in);
    free(out);
    return rc;
}

/* src/deflate_compress.c:362-376 */
int deflate_compress(int fd_in, int fd_out, int fd_stats, swi sw, int ops)
{
    deflate_compr_t *com;
    int rc;
    (void)ops; /* bit 0 (DEFLATE_NULLTERM) is a decompress option, :360 */
    com = spawn_deflate_compr_t();
    if (!com) return -E_MALLOC;
    deflate_compr_init(com, fd_in, fd_out, fd_stats, sw);
    rc = deflate_compr_run(com);
    deflate_compr_deinit(com);
    free(com);
    return rc;
}
