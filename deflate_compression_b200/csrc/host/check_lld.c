/*
 * host/check_lld.c -- the reference's test target (Makefile:26-27), automated.
 *
 * tests/check_lld.c of mparker97/deflate_compression forks, lets the child compress and has the parent re-expand
 * every (lit/len, dist) record through its own 32 KiB ring, printing what a decoder would produce; it is a visual
 * check ("not programmatically verified", :1-6), does not compile (:48, :74, :92) and hands the pipe to fd_out
 * instead of fd_stats (:74, "TODO: wrong").  This program does what that one intends, against
 * include/deflate_ext.h alone: the child calls deflate_compress(fd_in, fd_z, p[1], 32768, 0), the parent expands
 * the record stream and compares it with the input byte for byte, then inflates the zlib stream the child wrote
 * with deflate_decompress and compares that too.
 *
 *   check_lld FILE     exit status 0 iff both decodings reproduce FILE
 */
#include <fcntl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <sys/wait.h>
#include <unistd.h>

#include "deflate_ext.h"

#define SLIDING_WINDOW (1 << 15) /* tests/check_lld.c:18 */

static int read_full(int fd, void *buf, size_t n)
{
    unsigned char *p = (unsigned char *)buf;
    size_t got = 0;
    while (got < n) {
        ssize_t r = read(fd, p + got, n - got);
        if (r < 0) return -1;
        if (r == 0) break;
        got += (size_t)r;
    }
    return (int)got;
}

static unsigned char *slurp(const char *path, size_t *pn)
{
    struct stat st;
    unsigned char *b;
    int fd = open(path, O_RDONLY);
    if (fd < 0 || fstat(fd, &st) != 0) return NULL;
    b = (unsigned char *)malloc(st.st_size ? (size_t)st.st_size : 1);
    if (!b || read_full(fd, b, (size_t)st.st_size) != (int)st.st_size) { free(b); close(fd); return NULL; }
    close(fd);
    *pn = (size_t)st.st_size;
    return b;
}

/* the parent: one record per token; literal = append, match = copy ll bytes from d back (do_write, :20-39) */
static long expand_and_compare(int f, const unsigned char *want, size_t n, long *ntok, long *nmatch)
{
    unsigned char *ring = (unsigned char *)malloc(SLIDING_WINDOW);
    struct compress_stats cs;
    size_t pos = 0;
    long bad = -1;
    if (!ring) return 0;
    while (read_full(f, &cs, sizeof(cs)) == (int)sizeof(cs)) {
        int k, len = cs.d ? cs.ll : 1;
        if (cs.bytes != (int)(pos + 1) && bad < 0) bad = (long)pos; /* bytes consumed so far, + 1 (src/deflate_compress.c:235) */
        (*ntok)++;
        if (cs.d) (*nmatch)++;
        for (k = 0; k < len; k++, pos++) {
            unsigned char c = cs.d ? ring[(pos - (size_t)cs.d) % SLIDING_WINDOW] : (unsigned char)cs.ll;
            ring[pos % SLIDING_WINDOW] = c;
            if ((pos >= n || want[pos] != c) && bad < 0) bad = (long)pos;
        }
    }
    free(ring);
    if (bad < 0 && pos != n) bad = (long)pos;
    return bad;
}

int main(int argc, char *argv[])
{
    int p[2], fd_in, fd_z, status = 0;
    pid_t pid;
    unsigned char *want;
    size_t n = 0;
    char zpath[] = "/tmp/check_lld_XXXXXX";
    long ntok = 0, nmatch = 0, bad;
    if (argc != 2) {
        fprintf(stderr, "USAGE: %s FILE\n", argv[0]);
        return 2;
    }
    want = slurp(argv[1], &n);
    if (!want) {
        fprintf(stderr, "%s: no such file\n", argv[1]);
        return 2;
    }
    fd_z = mkstemp(zpath);
    if (fd_z < 0 || pipe(p) < 0) return 2;
    pid = fork();
    if (pid < 0) return 2;
    if (pid == 0) { /* the child compresses: zlib stream to fd_z, records to the pipe */
        int rc;
        close(p[0]);
        fd_in = open(argv[1], O_RDONLY);
        rc = fd_in < 0 ? -1 : deflate_compress(fd_in, fd_z, p[1], SLIDING_WINDOW, 0);
        close(p[1]);
        _exit(rc == 0 ? 0 : 1);
    }
    close(p[1]);
    bad = expand_and_compare(p[0], want, n, &ntok, &nmatch);
    close(p[0]);
    waitpid(pid, &status, 0);
    if (!WIFEXITED(status) || WEXITSTATUS(status) != 0) {
        fprintf(stderr, "check_lld: deflate_compress failed in the child\n");
        unlink(zpath);
        return 1;
    }
    if (bad >= 0) {
        fprintf(stderr, "check_lld: the expanded records differ from the input at byte %ld\n", bad);
        unlink(zpath);
        return 1;
    }
    { /* and the stream itself, through the other half of the public interface */
        struct string_len z, back;
        size_t zn = 0;
        int rc;
        z.str = slurp(zpath, &zn);
        z.len = zn;
        unlink(zpath);
        if (!z.str) return 2;
        rc = deflate_decompress(&back, &z, 0);
        if (rc < 0 || back.len != n || memcmp(back.str, want, n) != 0) {
            fprintf(stderr, "check_lld: deflate_decompress of the stream does not give the input back (rc %d)\n", rc);
            return 1;
        }
        printf("check_lld ok: %zu bytes, %ld tokens (%ld matches), stream %zu bytes\n", n, ntok, nmatch, zn);
        free(back.str);
        free(z.str);
    }
    free(want);
    return 0;
}
