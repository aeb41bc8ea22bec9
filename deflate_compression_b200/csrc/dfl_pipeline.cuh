/* dfl_pipeline.cuh -- host-side pipelines over the device compress call (included at the end of dfl_kernels.cu):
 *   dfl_b200_compress_fd          fd -> fd, the body of deflate_compress() (src/deflate_compress.c:139-145, :362-376):
 *                                 reader thread -> pinned pieces -> H2D / kernels / D2H on three streams -> writer thread
 *   dfl_b200_compress_host_multi  one host thread, N contexts (one per GPU), contiguous shards
 *   dfl_b200_adler32_combine      Adler-32 of a concatenation from the parts' checksums
 */
#pragma once

#include <condition_variable>
#include <mutex>
#include <thread>

#include <errno.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

extern "C" uint32_t dfl_b200_adler32_combine(uint32_t a1, uint32_t a2, uint64_t len2)
{
    /* both parts started from 1; RFC 1950 section 8.2 arithmetic mod 65521 */
    const uint64_t M = ADLER_MOD;
    const uint64_t s1a = a1 & 0xFFFF, s2a = a1 >> 16, s1b = a2 & 0xFFFF, s2b = a2 >> 16;
    const uint64_t s1 = (s1a + s1b + M - 1) % M;
    const uint64_t s2 = (s2a + s2b + (len2 % M) * ((s1a + M - 1) % M)) % M;
    return (uint32_t)((s2 << 16) | s1);
}

/* ------------------------------------------------------------------------- */
/* fd -> fd                                                                  */
/* ------------------------------------------------------------------------- */

#ifndef FD_PIECE_BYTES
#define FD_PIECE_BYTES (32u << 20) /* one pipeline piece; the first ones are smaller (8, 16 MiB) so that the pipeline fills fast */
#endif
#define FD_IN_SLOTS 3  /* pinned input pieces: one being read, one being copied, one being compressed */
#define FD_OUT_SLOTS 2 /* pinned output pieces: one being filled by the GPU, one being written */
#define FD_IO_THREADS 8

/* read(2) until n bytes or end of file (fetch(), src/deflate_compress.c:139-145).  A regular file is read by several
 * threads with pread(2): one thread's copy out of the page cache runs at a third of what the GPU compresses. */
static long long fd_read_piece(int fd, bool regular, unsigned long long file_off, uint8_t *buf, size_t n)
{
    if (regular) {
        struct stat st;
        if (fstat(fd, &st) != 0) return -1;
        const unsigned long long left = (unsigned long long)st.st_size > file_off ? (unsigned long long)st.st_size - file_off : 0;
        const size_t want = left < n ? (size_t)left : n;
        if (!want) return 0;
        const int nt = want >= (8u << 20) ? FD_IO_THREADS : 1;
        const size_t per = ((want + nt - 1) / nt + 4095) & ~(size_t)4095;
        bool bad = false;
        std::vector<std::thread> th;
        for (int t = 0; t < nt; t++) {
            const size_t lo = (size_t)t * per, hi = lo + per < want ? lo + per : want;
            if (lo >= hi) break;
            th.emplace_back([=, &bad] {
                size_t p = lo;
                while (p < hi) {
                    const ssize_t r = pread(fd, buf + p, hi - p, (off_t)(file_off + p));
                    if (r < 0 && errno == EINTR) continue;
                    if (r <= 0) { bad = true; return; } /* the file shrank under us, or an I/O error */
                    p += (size_t)r;
                }
            });
        }
        for (auto &t : th) t.join();
        return bad ? -1 : (long long)want;
    }
    size_t got = 0;
    while (got < n) {
        const ssize_t r = read(fd, buf + got, n - got);
        if (r < 0) {
            if (errno == EINTR) continue;
            return -1;
        }
        if (r == 0) break;
        got += (size_t)r;
    }
    return (long long)got;
}

static int fd_write_all(int fd, const uint8_t *p, size_t n)
{
    while (n) {
        const ssize_t w = write(fd, p, n);
        if (w < 0) {
            if (errno == EINTR) continue;
            return -1;
        }
        p += w;
        n -= (size_t)w;
    }
    return 0;
}

static void fd_free(dfl_ctx *ctx)
{
    if (!ctx->fd_h_in[0]) return;
    for (int i = 0; i < FD_IN_SLOTS; i++) { if (ctx->fd_h_in[i]) cudaFreeHost(ctx->fd_h_in[i]); ctx->fd_h_in[i] = NULL; }
    for (int i = 0; i < 2; i++) {
        if (ctx->fd_h_out[i]) cudaFreeHost(ctx->fd_h_out[i]);
        if (ctx->fd_d_in[i]) cudaFree(ctx->fd_d_in[i]);
        if (ctx->fd_d_out[i]) cudaFree(ctx->fd_d_out[i]);
        if (ctx->fd_e_in[i]) cudaEventDestroy(ctx->fd_e_in[i]);
        if (ctx->fd_e_out[i]) cudaEventDestroy(ctx->fd_e_out[i]);
        ctx->fd_h_out[i] = ctx->fd_d_in[i] = ctx->fd_d_out[i] = NULL;
        ctx->fd_e_in[i] = ctx->fd_e_out[i] = NULL;
    }
}

/* pinned host pieces and device slots of the fd pipelines live with the context */
static int fd_slots_init(dfl_ctx *ctx, size_t piece, size_t bound)
{
    const size_t W = DFL_MAX_WINDOW;
    if (ctx->fd_h_in[0] && ctx->fd_piece != piece) fd_free(ctx); /* (piece or segment size changed between calls) */
    if (!ctx->fd_h_in[0]) {
        for (int s = 0; s < FD_IN_SLOTS; s++) CK(cudaMallocHost(&ctx->fd_h_in[s], piece));
        for (int s = 0; s < FD_OUT_SLOTS; s++) CK(cudaMallocHost(&ctx->fd_h_out[s], bound + 16));
        for (int s = 0; s < 2; s++) {
            CK(cudaMalloc(&ctx->fd_d_in[s], W + piece + 32));
            CK(cudaMalloc(&ctx->fd_d_out[s], bound));
        }
        for (int s = 0; s < 2; s++) {
            CK(cudaEventCreateWithFlags(&ctx->fd_e_in[s], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&ctx->fd_e_out[s], cudaEventDisableTiming));
        }
        ctx->fd_piece = piece;
    }
    return DFL_OK;
}

extern "C" int dfl_b200_set_fd_piece(dfl_ctx *ctx, size_t bytes)
{
    if (!ctx) return DFL_E_INVAL;
    ctx->fd_piece_req = bytes;
    return DFL_OK;
}

/* bytes of piece i: a ramp (quarter, half) up to the full piece, in whole segments */
static size_t fd_piece_len(size_t piece, size_t seg, long i)
{
    size_t p = i == 0 ? piece / 4 : i == 1 ? piece / 2 : piece;
    p = p / seg * seg;
    if (p < DFL_MAX_WINDOW) p = (DFL_MAX_WINDOW + seg - 1) / seg * seg; /* a piece holds the next one's window */
    return p < piece ? p : piece;
}

/* One output piece to fd_out.  A regular file that can be mapped (opened read-write: a memfd, a file opened O_RDWR) is
 * grown with fallocate(2) and filled through a shared mapping by several threads -- plain copies at memory speed;
 * write(2) / pwrite(2) into the page cache take a lock per page and stop at ~5 GB/s however many threads call them
 * (measured: the writer was what bounded the fd -> fd call).  Anything else: pwrite(2) from several threads, or write(2). */
static int fd_write_piece(int fd, bool regular, unsigned long long file_off, const uint8_t *p, size_t n)
{
    if (!regular || n < (1u << 20)) {
        if (regular && lseek(fd, (off_t)file_off, SEEK_SET) < 0) return -1;
        return fd_write_all(fd, p, n);
    }
    const int nt = FD_IO_THREADS;
    const size_t per = ((n + nt - 1) / nt + 4095) & ~(size_t)4095;
    {
        struct stat st;
        const unsigned long long end = file_off + n, map_lo = file_off & ~4095ull;
        /* fallocate(2) grows the file AND reserves its pages: a store through the mapping can then not run out of
         * space (which would be a SIGBUS, not an error return) */
        if (fstat(fd, &st) == 0 && fallocate(fd, 0, (off_t)file_off, (off_t)n) == 0) {
            void *m = mmap(NULL, (size_t)(end - map_lo), PROT_READ | PROT_WRITE, MAP_SHARED | MAP_POPULATE, fd, (off_t)map_lo); /* (populated in one go: a fault per page from the copying threads costs more) */
            if (m != MAP_FAILED) {
                uint8_t *dst = (uint8_t *)m + (file_off - map_lo);
                std::vector<std::thread> th;
                for (int t = 0; t < nt; t++) {
                    const size_t lo = (size_t)t * per, hi = lo + per < n ? lo + per : n;
                    if (lo >= hi) break;
                    th.emplace_back([=] { memcpy(dst + lo, p + lo, hi - lo); });
                }
                for (auto &t : th) t.join();
                munmap(m, (size_t)(end - map_lo));
                return lseek(fd, (off_t)end, SEEK_SET) < 0 ? -1 : 0; /* as write(2) would have left it */
            }
        }
    }
    bool bad = false;
    std::vector<std::thread> th;
    for (int t = 0; t < nt; t++) {
        const size_t lo = (size_t)t * per, hi = lo + per < n ? lo + per : n;
        if (lo >= hi) break;
        th.emplace_back([=, &bad] {
            size_t q = lo;
            while (q < hi) {
                const ssize_t w = pwrite(fd, p + q, hi - q, (off_t)(file_off + q));
                if (w < 0 && errno == EINTR) continue;
                if (w <= 0) { bad = true; return; }
                q += (size_t)w;
            }
        });
    }
    for (auto &t : th) t.join();
    if (bad) return -1;
    return lseek(fd, (off_t)(file_off + n), SEEK_SET) < 0 ? -1 : 0; /* as write(2) would have left it */
}

namespace {
/* DFL_B200_FD_TRACE=1: a line per pipeline event on stderr (ms since the call began) -- where an fd -> fd call spends its time */
struct FdTrace {
    bool on;
    double t0;
    std::mutex mu;
    static double now() { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; }
    FdTrace() : on(getenv("DFL_B200_FD_TRACE") && atoi(getenv("DFL_B200_FD_TRACE"))), t0(now()) {}
    void ev(const char *what, long i, size_t n = 0)
    {
        if (!on) return;
        std::lock_guard<std::mutex> lk(mu);
        fprintf(stderr, "[fd %8.3f ms] %-14s piece %ld (%zu B)\n", now() - t0, what, i, n);
    }
};
struct FdPipe {
    std::mutex mu;
    std::condition_variable cv;
    /* reader -> main: piece i is in h_in[i % FD_IN_SLOTS] with in_len[slot] bytes once in_ready > i; the main thread
     * frees a slot by raising in_free (pieces whose kernels are done) */
    long in_ready = 0, in_free = 0;
    size_t in_len[FD_IN_SLOTS] = {0, 0, 0};
    bool in_eof = false, in_err = false;
    /* main -> writer: output piece j is in h_out[j % FD_OUT_SLOTS] with out_len bytes once out_ready > j */
    long out_ready = 0, out_done = 0;
    size_t out_len[FD_OUT_SLOTS] = {0, 0};
    bool out_err = false, out_close = false;
};
} // namespace

extern "C" int dfl_b200_compress_fd(dfl_ctx *ctx, int fd_in, int fd_out, unsigned long long *in_bytes, unsigned long long *out_bytes)
{
    if (!ctx) return DFL_E_INVAL;
    CK(cudaSetDevice(ctx->device));
    const dfl_params prm = ctx->prm;
    size_t piece = (ctx->fd_piece_req ? ctx->fd_piece_req : (size_t)FD_PIECE_BYTES) / (size_t)prm.seg_bytes * (size_t)prm.seg_bytes; /* whole segments */
    if (piece < (size_t)prm.seg_bytes) piece = (size_t)prm.seg_bytes;
    if (piece < DFL_MAX_WINDOW) piece = (DFL_MAX_WINDOW + (size_t)prm.seg_bytes - 1) / (size_t)prm.seg_bytes * (size_t)prm.seg_bytes; /* a piece holds the next one's window */
    const size_t W = DFL_MAX_WINDOW, bound = dfl_b200_bound(&prm, piece) + 16;
    int rc = host_pipe_init(ctx, 1); /* the three streams */
    if (rc) return rc;
    rc = fd_slots_init(ctx, piece, bound);
    if (rc) return rc;
    struct stat st_in;
    const bool regular = fstat(fd_in, &st_in) == 0 && S_ISREG(st_in.st_mode);
    unsigned long long start_off = 0;
    if (regular) {
        const off_t cur = lseek(fd_in, 0, SEEK_CUR);
        start_off = cur > 0 ? (unsigned long long)cur : 0;
    }

    FdPipe P;
    FdTrace T;
    std::thread reader([&] {
        unsigned long long off = start_off;
        for (long i = 0;; i++) {
            {
                std::unique_lock<std::mutex> lk(P.mu);
                P.cv.wait(lk, [&] { return i - P.in_free < FD_IN_SLOTS || P.out_err; });
                if (P.out_err) return;
            }
            const int s = (int)(i % FD_IN_SLOTS);
            const size_t want = fd_piece_len(piece, (size_t)prm.seg_bytes, i);
            T.ev("read begin", i, want);
            const long long got = fd_read_piece(fd_in, regular, off, ctx->fd_h_in[s], want);
            T.ev("read end", i, got > 0 ? (size_t)got : 0);
            std::lock_guard<std::mutex> lk(P.mu);
            if (got < 0) { P.in_err = true; P.in_eof = true; P.cv.notify_all(); return; }
            off += (unsigned long long)got;
            P.in_len[s] = (size_t)got;
            P.in_ready = i + 1;
            if ((size_t)got < want) P.in_eof = true;
            P.cv.notify_all();
            if ((size_t)got < want) return;
        }
    });
    struct stat st_out;
    const bool regular_out = fd_out >= 0 && fstat(fd_out, &st_out) == 0 && S_ISREG(st_out.st_mode);
    unsigned long long out_off = 0;
    if (regular_out) {
        const off_t cur = lseek(fd_out, 0, SEEK_CUR);
        out_off = cur > 0 ? (unsigned long long)cur : 0;
    }
    std::thread writer([&] {
        static const uint8_t hdr[2] = {0x78, 0x9C}; /* CM = 8, CINFO = 7, FCHECK: what src/deflate_decompress.c:352-360 checks */
        bool bad = fd_out >= 0 && fd_write_all(fd_out, hdr, 2) != 0;
        unsigned long long woff = out_off + 2;
        for (long j = 0;; j++) {
            size_t len;
            {
                std::unique_lock<std::mutex> lk(P.mu);
                if (bad) { P.out_err = true; P.cv.notify_all(); }
                P.cv.wait(lk, [&] { return P.out_ready > j || P.out_close; });
                if (P.out_ready <= j) return;
                len = P.out_len[j % FD_OUT_SLOTS];
            }
            T.ev("write begin", j, len);
            if (!bad && fd_out >= 0 && fd_write_piece(fd_out, regular_out, woff, ctx->fd_h_out[j % FD_OUT_SLOTS], len) != 0) bad = true;
            T.ev("write end", j, len);
            woff += len;
            std::lock_guard<std::mutex> lk(P.mu);
            P.out_done = j + 1;
            if (bad) P.out_err = true;
            P.cv.notify_all();
        }
    });

    unsigned long long total_in = 0, total_out = 2;
    uint32_t adler = 1;
    size_t len_cur = 0, len_next = 0, body_prev = 0;
    rc = DFL_OK;
    /* piece i's bytes: the tail of piece i - 1 primes its window (device to device), then the copy in */
    auto copy_in = [&](long i, size_t len, size_t len_before) -> bool {
        const int ds = (int)(i & 1);
        uint8_t *d_piece = ctx->fd_d_in[ds] + W;
        const size_t hist = i == 0 ? 0 : (len_before < W ? len_before : W);
        if (hist && cudaMemcpyAsync(d_piece - hist, ctx->fd_d_in[ds ^ 1] + W + len_before - hist, hist, cudaMemcpyDeviceToDevice, ctx->s_in) != cudaSuccess) return false;
        if (len && cudaMemcpyAsync(d_piece, ctx->fd_h_in[i % FD_IN_SLOTS], len, cudaMemcpyHostToDevice, ctx->s_in) != cudaSuccess) return false;
        return cudaEventRecord(ctx->fd_e_in[ds], ctx->s_in) == cudaSuccess;
    };
    auto wait_piece = [&](long i, size_t *len) -> bool { /* the reader has delivered piece i */
        std::unique_lock<std::mutex> lk(P.mu);
        P.cv.wait(lk, [&] { return P.in_ready > i || P.in_err; });
        if (P.in_err) return false;
        *len = P.in_len[i % FD_IN_SLOTS];
        return true;
    };
    auto publish = [&](long j, size_t olen) { /* output piece j is complete in its pinned slot: over to the writer */
        std::lock_guard<std::mutex> lk(P.mu);
        P.out_len[j % FD_OUT_SLOTS] = olen;
        P.out_ready = j + 1;
        P.cv.notify_all();
    };
    if (!wait_piece(0, &len_cur)) rc = DFL_E_INVAL;
    else if (!copy_in(0, len_cur, 0)) rc = DFL_E_CUDA;
    for (long i = 0; rc == DFL_OK; i++) {
        const int ds = (int)(i & 1);
        const size_t len = len_cur, hist = i == 0 ? 0 : (total_in < W ? (size_t)total_in : W);
        /* a piece that fills its slot may or may not be the last one: the next read tells (it overlapped piece i - 1) */
        bool last = len < fd_piece_len(piece, (size_t)prm.seg_bytes, i);
        if (!last) {
            if (!wait_piece(i + 1, &len_next)) { rc = DFL_E_INVAL; break; }
            last = len_next == 0;
        }
        if (cudaStreamWaitEvent(ctx->s_run, ctx->fd_e_in[ds], 0) != cudaSuccess) { rc = DFL_E_CUDA; break; }
        T.ev("enqueue", i, len);
        rc = dfl_b200_deflate_enqueue(ctx, ctx->fd_d_in[ds] + W, len, hist, last ? 1 : 0, ctx->fd_d_out[ds], bound, ctx->s_run);
        if (rc) break;
        T.ev("enqueued", i, len);
        /* while piece i is compressed: piece i + 1 goes to the device, piece i - 1 goes to the writer */
        if (!last && !copy_in(i + 1, len_next, len)) { rc = DFL_E_CUDA; break; }
        if (i > 0) {
            if (cudaEventSynchronize(ctx->fd_e_out[ds ^ 1]) != cudaSuccess) { rc = DFL_E_CUDA; break; }
            T.ev("d2h done", i - 1, body_prev);
            publish(i - 1, body_prev);
        }
        size_t body = 0;
        uint32_t a = 1;
        rc = dfl_b200_deflate_finish(ctx, &body, &a, ctx->s_run);
        if (rc) break;
        T.ev("kernels done", i, body);
        adler = dfl_b200_adler32_combine(adler, a, len);
        total_in += len;
        {
            /* piece i's input slot can be read into again; its output slot must have been written out (two pieces ago) */
            std::unique_lock<std::mutex> lk(P.mu);
            P.in_free = i + 1;
            P.cv.notify_all();
            P.cv.wait(lk, [&] { return P.out_done >= i - (FD_OUT_SLOTS - 1) || P.out_err; });
            if (P.out_err) { rc = DFL_E_INVAL; break; }
        }
        uint8_t *ho = ctx->fd_h_out[i % FD_OUT_SLOTS];
        if (body && cudaMemcpyAsync(ho, ctx->fd_d_out[ds], body, cudaMemcpyDeviceToHost, ctx->s_out) != cudaSuccess) { rc = DFL_E_CUDA; break; }
        if (cudaEventRecord(ctx->fd_e_out[ds], ctx->s_out) != cudaSuccess) { rc = DFL_E_CUDA; break; }
        total_out += body;
        body_prev = body;
        len_cur = len_next;
        if (last) {
            if (cudaEventSynchronize(ctx->fd_e_out[ds]) != cudaSuccess) { rc = DFL_E_CUDA; break; }
            /* RFC 1950: Adler-32 of the uncompressed data, MSB first */
            ho[body + 0] = (uint8_t)(adler >> 24); ho[body + 1] = (uint8_t)(adler >> 16);
            ho[body + 2] = (uint8_t)(adler >> 8); ho[body + 3] = (uint8_t)adler;
            total_out += 4;
            publish(i, body + 4);
            break;
        }
    }
    {
        std::lock_guard<std::mutex> lk(P.mu);
        P.out_close = true;
        if (rc) P.out_err = true; /* releases the reader */
        P.in_free = 1L << 60;
        P.cv.notify_all();
    }
    reader.join();
    writer.join();
    T.ev("joined", -1);
    if (regular) lseek(fd_in, (off_t)(start_off + total_in), SEEK_SET); /* as read(2) would have left it */
    cudaStreamSynchronize(ctx->s_in);
    cudaStreamSynchronize(ctx->s_run);
    cudaStreamSynchronize(ctx->s_out);
    if (rc) return rc;
    if (P.out_err || P.in_err) return DFL_E_INVAL;
    if (in_bytes) *in_bytes = total_in;
    if (out_bytes) *out_bytes = total_out;
    return DFL_OK;
}

/* ------------------------------------------------------------------------- */
/* N GPUs behind one call                                                    */
/* ------------------------------------------------------------------------- */

/*
 * The input shards contiguously, one segment-aligned range per context (= per GPU), each primed with the 32 KiB
 * before it taken from the caller's buffer (no GPU <-> GPU halo), each ended byte aligned by its sync marker; only
 * the last shard's last block carries BFINAL.  One host thread: every device gets its copy-in and kernels enqueued,
 * then the bodies come back in rank order straight into the caller's buffer (the offset of shard r is known once
 * shards 0..r-1 have reported their sizes) and the Adler-32 values are combined.  The copies overlap across devices
 * when `in` / `out` are pinned (cudaHostRegister / cudaMallocHost); with pageable buffers the result is the same.
 */
extern "C" int dfl_b200_compress_host_multi(dfl_ctx **ctxs, int ndev, const uint8_t *in, size_t n, uint8_t *out, size_t cap,
                                            size_t *out_len)
{
    if (!ctxs || ndev < 1 || (!in && n) || !out || !out_len) return DFL_E_INVAL;
    for (int r = 0; r < ndev; r++)
        if (!ctxs[r] || memcmp(&ctxs[r]->prm, &ctxs[0]->prm, sizeof(dfl_params)) != 0) return DFL_E_INVAL;
    const dfl_params prm = ctxs[0]->prm;
    const size_t nseg = n ? (n + prm.seg_bytes - 1) / prm.seg_bytes : 0;
    if ((size_t)ndev > nseg) ndev = nseg ? (int)nseg : 1; /* every shard holds at least one segment */
    if (ndev == 1) return dfl_b200_compress_host(ctxs[0], in, n, out, cap, out_len);
    const size_t W = (size_t)prm.window;
    std::vector<size_t> lo(ndev), hi(ndev), hist(ndev);
    for (int r = 0; r < ndev; r++) {
        lo[r] = nseg * (size_t)r / (size_t)ndev * (size_t)prm.seg_bytes;
        hi[r] = nseg * (size_t)(r + 1) / (size_t)ndev * (size_t)prm.seg_bytes;
        if (lo[r] > n) lo[r] = n;
        if (hi[r] > n || r == ndev - 1) hi[r] = n;
        hist[r] = prm.use_dict ? (lo[r] < W ? lo[r] : W) : 0;
    }
    /* enqueue everything */
    for (int r = 0; r < ndev; r++) {
        dfl_ctx *ctx = ctxs[r];
        CK(cudaSetDevice(ctx->device));
        int rc = host_pipe_init(ctx, 1);
        if (rc) return rc;
        const size_t len = hi[r] - lo[r];
        const size_t bound = dfl_b200_bound(&prm, len);
        { const int rc_ = grow_dev(ctx, &ctx->d_in, &ctx->d_in_cap, W + len + 32); if (rc_) return rc_; }
        { const int rc_ = grow_dev(ctx, &ctx->d_out, &ctx->d_out_cap, bound + 16); if (rc_) return rc_; }
        /* the shard sits W bytes into the buffer so that its history lies in front of it, whatever its length */
        uint8_t *d_shard = ctx->d_in + W;
        if (hist[r] + len)
            CK(cudaMemcpyAsync(d_shard - hist[r], in + lo[r] - hist[r], hist[r] + len, cudaMemcpyHostToDevice, ctx->s_run));
        rc = dfl_b200_deflate_enqueue(ctx, d_shard, len, hist[r], r == ndev - 1, ctx->d_out, bound, ctx->s_run);
        if (rc) return rc;
    }
    /* collect in rank order */
    size_t pos = 2;
    uint32_t adler = 1;
    int rc = DFL_OK;
    for (int r = 0; r < ndev; r++) {
        dfl_ctx *ctx = ctxs[r];
        size_t body = 0;
        uint32_t a = 1;
        if (cudaSetDevice(ctx->device) != cudaSuccess) { rc = DFL_E_CUDA; break; }
        rc = dfl_b200_deflate_finish(ctx, &body, &a, ctx->s_run);
        if (rc) break;
        if (cap < pos + body + 4) { rc = DFL_E_SZ; break; }
        if (body && cudaMemcpyAsync(out + pos, ctx->d_out, body, cudaMemcpyDeviceToHost, ctx->s_run) != cudaSuccess) { rc = DFL_E_CUDA; break; }
        pos += body;
        adler = dfl_b200_adler32_combine(adler, a, hi[r] - lo[r]);
    }
    for (int r = 0; r < ndev; r++) { /* everything in flight lands before the caller's buffers go away */
        cudaSetDevice(ctxs[r]->device);
        if (ctxs[r]->s_run && cudaStreamSynchronize(ctxs[r]->s_run) != cudaSuccess && !rc) rc = DFL_E_CUDA;
    }
    if (rc) return rc;
    out[0] = 0x78;
    out[1] = 0x9C;
    out[pos + 0] = (uint8_t)(adler >> 24);
    out[pos + 1] = (uint8_t)(adler >> 16);
    out[pos + 2] = (uint8_t)(adler >> 8);
    out[pos + 3] = (uint8_t)adler;
    *out_len = pos + 4;
    return DFL_OK;
}

/*
 * fd -> fd on N GPUs: what deflate_compress() runs when DFL_B200_DEVICES names more than one device.  The pipeline of
 * dfl_b200_compress_fd with the pieces dealt round the contexts: piece i goes to context i mod N, primed with the 32 KiB
 * before it from the pinned slot piece i - 1 was read into (no device-to-device copy); N pieces are in flight, one per
 * context; they are retired in order (size and Adler-32 read back, body copied out, handed to the writer thread).  The
 * reader thread, the writer thread and the pinned slots are the single-device call's (3 input and 2 output slots per
 * context, all used as one ring).  Byte-identical to the one-GPU calls.
 */
extern "C" int dfl_b200_compress_fd_multi(dfl_ctx **ctxs, int ndev, int fd_in, int fd_out, unsigned long long *in_bytes,
                                          unsigned long long *out_bytes)
{
    if (!ctxs || ndev < 1 || !ctxs[0]) return DFL_E_INVAL;
    if (ndev == 1) return dfl_b200_compress_fd(ctxs[0], fd_in, fd_out, in_bytes, out_bytes);
    for (int r = 0; r < ndev; r++)
        if (!ctxs[r] || memcmp(&ctxs[r]->prm, &ctxs[0]->prm, sizeof(dfl_params)) != 0) return DFL_E_INVAL;
    dfl_ctx *ctx = ctxs[0];
    const dfl_params prm = ctx->prm;
    size_t piece = (ctx->fd_piece_req ? ctx->fd_piece_req : (size_t)FD_PIECE_BYTES) / (size_t)prm.seg_bytes * (size_t)prm.seg_bytes;
    if (piece < (size_t)prm.seg_bytes) piece = (size_t)prm.seg_bytes;
    if (piece < DFL_MAX_WINDOW) piece = (DFL_MAX_WINDOW + (size_t)prm.seg_bytes - 1) / (size_t)prm.seg_bytes * (size_t)prm.seg_bytes;
    const size_t W = DFL_MAX_WINDOW, bound = dfl_b200_bound(&prm, piece) + 16;
    const int NIN = FD_IN_SLOTS * ndev, NOUT = FD_OUT_SLOTS * ndev;
    std::vector<uint8_t *> h_in(NIN), h_out(NOUT);
    for (int r = 0; r < ndev; r++) {
        ctx = ctxs[r];
        CK(cudaSetDevice(ctx->device));
        int rc = host_pipe_init(ctx, 1);
        if (rc) return rc;
        rc = fd_slots_init(ctx, piece, bound);
        if (rc) return rc;
        for (int s = 0; s < FD_IN_SLOTS; s++) h_in[r * FD_IN_SLOTS + s] = ctx->fd_h_in[s];
        for (int s = 0; s < FD_OUT_SLOTS; s++) h_out[r * FD_OUT_SLOTS + s] = ctx->fd_h_out[s];
    }
    ctx = ctxs[0];
    struct stat st_in;
    const bool regular = fstat(fd_in, &st_in) == 0 && S_ISREG(st_in.st_mode);
    unsigned long long start_off = 0;
    if (regular) {
        const off_t cur = lseek(fd_in, 0, SEEK_CUR);
        start_off = cur > 0 ? (unsigned long long)cur : 0;
    }
    /* shared with the two threads (the ring sizes differ from FdPipe's) */
    std::mutex mu;
    std::condition_variable cv;
    long in_ready = 0, in_free = 0, out_ready = 0, out_done = 0;
    std::vector<size_t> in_len(NIN, 0), out_len(NOUT, 0);
    bool in_eof = false, in_err = false, out_err = false, out_close = false;
    FdTrace T;
    std::thread reader([&] {
        unsigned long long off = start_off;
        for (long i = 0;; i++) {
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return i - in_free < NIN || out_err; });
                if (out_err) return;
            }
            const int s = (int)(i % NIN);
            const size_t want = fd_piece_len(piece, (size_t)prm.seg_bytes, i);
            T.ev("read begin", i, want);
            const long long got = fd_read_piece(fd_in, regular, off, h_in[s], want);
            T.ev("read end", i, got > 0 ? (size_t)got : 0);
            std::lock_guard<std::mutex> lk(mu);
            if (got < 0) { in_err = true; in_eof = true; cv.notify_all(); return; }
            off += (unsigned long long)got;
            in_len[s] = (size_t)got;
            in_ready = i + 1;
            if ((size_t)got < want) in_eof = true;
            cv.notify_all();
            if ((size_t)got < want) return;
        }
    });
    struct stat st_out;
    const bool regular_out = fd_out >= 0 && fstat(fd_out, &st_out) == 0 && S_ISREG(st_out.st_mode);
    unsigned long long out_off = 0;
    if (regular_out) {
        const off_t cur = lseek(fd_out, 0, SEEK_CUR);
        out_off = cur > 0 ? (unsigned long long)cur : 0;
    }
    std::thread writer([&] {
        static const uint8_t hdr[2] = {0x78, 0x9C};
        bool bad = fd_out >= 0 && fd_write_all(fd_out, hdr, 2) != 0;
        unsigned long long woff = out_off + 2;
        for (long j = 0;; j++) {
            size_t len;
            {
                std::unique_lock<std::mutex> lk(mu);
                if (bad) { out_err = true; cv.notify_all(); }
                cv.wait(lk, [&] { return out_ready > j || out_close; });
                if (out_ready <= j) return;
                len = out_len[j % NOUT];
            }
            T.ev("write begin", j, len);
            if (!bad && fd_out >= 0 && fd_write_piece(fd_out, regular_out, woff, h_out[j % NOUT], len) != 0) bad = true;
            T.ev("write end", j, len);
            woff += len;
            std::lock_guard<std::mutex> lk(mu);
            out_done = j + 1;
            if (bad) out_err = true;
            cv.notify_all();
        }
    });

    unsigned long long total_in = 0, total_out = 2;
    uint32_t adler = 1;
    int rc = DFL_OK;
    std::vector<size_t> plen; /* length of every piece seen so far */
    long n_pieces = -1;       /* known once the last piece has been seen */
    auto wait_piece = [&](long i, size_t *len) -> bool { /* the reader has delivered piece i (or reached the end before it) */
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return in_ready > i || in_err || in_eof; });
        if (in_err) return false;
        *len = in_ready > i ? in_len[i % NIN] : 0;
        return true;
    };
    auto publish = [&](long j, size_t olen) {
        std::lock_guard<std::mutex> lk(mu);
        out_len[j % NOUT] = olen;
        out_ready = j + 1;
        cv.notify_all();
    };
    /* piece i to its context: the window from the slot of piece i - 1, the piece itself, then the kernels */
    auto stage = [&](long i) -> int {
        dfl_ctx *c = ctxs[i % ndev];
        const int ds = (int)((i / ndev) & 1);
        if (cudaSetDevice(c->device) != cudaSuccess) return DFL_E_CUDA;
        uint8_t *d_piece = c->fd_d_in[ds] + W;
        const size_t len = plen[i], before = i ? plen[i - 1] : 0, hist = before < W ? before : W;
        if (hist && cudaMemcpyAsync(d_piece - hist, h_in[(i - 1) % NIN] + before - hist, hist, cudaMemcpyHostToDevice, c->s_in) != cudaSuccess) return DFL_E_CUDA;
        if (len && cudaMemcpyAsync(d_piece, h_in[i % NIN], len, cudaMemcpyHostToDevice, c->s_in) != cudaSuccess) return DFL_E_CUDA;
        return cudaEventRecord(c->fd_e_in[ds], c->s_in) == cudaSuccess ? DFL_OK : DFL_E_CUDA;
    };
    auto enqueue = [&](long i) -> int {
        dfl_ctx *c = ctxs[i % ndev];
        const int ds = (int)((i / ndev) & 1);
        if (cudaSetDevice(c->device) != cudaSuccess) return DFL_E_CUDA;
        const size_t before = i ? plen[i - 1] : 0, hist = before < W ? before : W;
        if (cudaStreamWaitEvent(c->s_run, c->fd_e_in[ds], 0) != cudaSuccess) return DFL_E_CUDA;
        T.ev("enqueue", i, plen[i]);
        return dfl_b200_deflate_enqueue(c, c->fd_d_in[ds] + W, plen[i], hist, i == n_pieces - 1 ? 1 : 0, c->fd_d_out[ds], bound, c->s_run);
    };
    long pending_pub = -1; /* piece whose body is on its way to its pinned slot */
    size_t pending_len = 0;
    auto flush_pub = [&]() -> int {
        if (pending_pub < 0) return DFL_OK;
        dfl_ctx *c = ctxs[pending_pub % ndev];
        if (cudaSetDevice(c->device) != cudaSuccess || cudaEventSynchronize(c->fd_e_out[(pending_pub / ndev) & 1]) != cudaSuccess) return DFL_E_CUDA;
        T.ev("d2h done", pending_pub, pending_len);
        if (pending_pub == n_pieces - 1) { /* RFC 1950: Adler-32 of the uncompressed data, MSB first */
            uint8_t *ho = h_out[pending_pub % NOUT] + pending_len;
            ho[0] = (uint8_t)(adler >> 24); ho[1] = (uint8_t)(adler >> 16); ho[2] = (uint8_t)(adler >> 8); ho[3] = (uint8_t)adler;
            pending_len += 4;
            total_out += 4;
        }
        publish(pending_pub, pending_len);
        pending_pub = -1;
        return DFL_OK;
    };
    auto retire = [&](long j) -> int {
        dfl_ctx *c = ctxs[j % ndev];
        const int ds = (int)((j / ndev) & 1);
        if (cudaSetDevice(c->device) != cudaSuccess) return DFL_E_CUDA;
        size_t body = 0;
        uint32_t a = 1;
        int r_ = dfl_b200_deflate_finish(c, &body, &a, c->s_run);
        if (r_) return r_;
        T.ev("kernels done", j, body);
        adler = dfl_b200_adler32_combine(adler, a, plen[j]);
        total_in += plen[j];
        r_ = flush_pub(); /* (the piece before: its copy out had this piece's kernels to hide behind) */
        if (r_) return r_;
        {
            /* the slot of piece j - 1 can be read into again (piece j's copy in, which read its tail, is done); piece
             * j's output slot must have been written out */
            std::unique_lock<std::mutex> lk(mu);
            in_free = j;
            cv.notify_all();
            cv.wait(lk, [&] { return out_done >= j - (NOUT - 1) || out_err; });
            if (out_err) return DFL_E_INVAL;
        }
        if (body && cudaMemcpyAsync(h_out[j % NOUT], c->fd_d_out[ds], body, cudaMemcpyDeviceToHost, c->s_out) != cudaSuccess) return DFL_E_CUDA;
        if (cudaEventRecord(c->fd_e_out[ds], c->s_out) != cudaSuccess) return DFL_E_CUDA;
        total_out += body;
        pending_pub = j;
        pending_len = body;
        return DFL_OK;
    };
    for (long i = 0; rc == DFL_OK; i++) {
        if (n_pieces < 0) {
            /* piece i, and whether it is the last one (a piece that fills its slot is followed by another read) */
            size_t len = 0, nxt = 0;
            if ((long)plen.size() <= i) {
                if (!wait_piece(i, &len)) { rc = DFL_E_INVAL; break; }
                plen.push_back(len);
            }
            bool last = plen[i] < fd_piece_len(piece, (size_t)prm.seg_bytes, i);
            if (!last) {
                if (!wait_piece(i + 1, &nxt)) { rc = DFL_E_INVAL; break; }
                if (nxt == 0) last = true; else plen.push_back(nxt);
            }
            if (last) n_pieces = i + 1;
            rc = stage(i);
            if (rc) break;
        } else if (i >= n_pieces + ndev) {
            break;
        }
        if (i >= ndev && i - ndev < (n_pieces < 0 ? i : n_pieces)) { rc = retire(i - ndev); if (rc) break; }
        if (n_pieces < 0 || i < n_pieces) { rc = enqueue(i); if (rc) break; }
    }
    if (rc == DFL_OK) rc = flush_pub();
    {
        std::lock_guard<std::mutex> lk(mu);
        out_close = true;
        if (rc) out_err = true; /* releases the reader */
        in_free = 1L << 60;
        cv.notify_all();
    }
    reader.join();
    writer.join();
    T.ev("joined", -1);
    if (regular) lseek(fd_in, (off_t)(start_off + total_in), SEEK_SET); /* as read(2) would have left it */
    for (int r = 0; r < ndev; r++) {
        cudaSetDevice(ctxs[r]->device);
        cudaStreamSynchronize(ctxs[r]->s_in);
        cudaStreamSynchronize(ctxs[r]->s_run);
        cudaStreamSynchronize(ctxs[r]->s_out);
    }
    if (rc) return rc;
    if (out_err || in_err) return DFL_E_INVAL;
    if (in_bytes) *in_bytes = total_in;
    if (out_bytes) *out_bytes = total_out;
    return DFL_OK;
}

extern "C" int dfl_b200_copy_async(void *dst, const void *src, size_t n, void *stream)
{
    dfl_ctx *ctx = NULL;
    if (n) CK(cudaMemcpyAsync(dst, src, n, cudaMemcpyDefault, (cudaStream_t)stream));
    return DFL_OK;
}
