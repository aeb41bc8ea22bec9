/*
 * host/zencode.c -- the `zencode` CLI the reference's Makefile names
 * (Makefile:13,20-21) but never gives a main().  Intended use per
 * src/deflate_compress.c:362: deflate_compress(STDIN_FILENO, STDOUT_FILENO, ...).
 *
 *   zencode [-d] [-w window] [-s stats_file] [infile [outfile]]
 *     (default)  compress infile|stdin to a zlib stream on outfile|stdout (GPU)
 *     -d         decompress instead (CPU, deflate_decompress)
 *     -w N       sliding window, 1..32768 (default 32768, tests/check_lld.c:18)
 *     -s FILE    write one struct compress_stats per token to FILE (fd_stats)
 */
#include <fcntl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "deflate_ext.h"

static int slurp(int fd, unsigned char **pb, size_t *pn)
{
    size_t cap = 1 << 20, n = 0;
    unsigned char *b = malloc(cap);
    if (!b) return -1;
    for (;;) {
        ssize_t r;
        if (n == cap) {
            unsigned char *nb = realloc(b, cap *= 2);
            if (!nb) { free(b); return -1; }
            b = nb;
        }
        r = read(fd, b + n, cap - n);
        if (r < 0) { free(b); return -1; }
        if (r == 0) break;
        n += (size_t)r;
    }
    *pb = b;
    *pn = n;
    return 0;
}

int main(int argc, char *argv[])
{
    int decompress = 0, window = 32768, fd_in = STDIN_FILENO, fd_out = STDOUT_FILENO, fd_stats = -1, i, rc;
    const char *stats = NULL, *files[2] = {NULL, NULL};
    int nfiles = 0;
    for (i = 1; i < argc; i++) {
        if (!strcmp(argv[i], "-d")) decompress = 1;
        else if (!strcmp(argv[i], "-w") && i + 1 < argc) window = atoi(argv[++i]);
        else if (!strcmp(argv[i], "-s") && i + 1 < argc) stats = argv[++i];
        else if (!strcmp(argv[i], "-b")) setenv("DFL_B200_SPLIT", "1", 1); /* adaptive block splitting */
        else if (argv[i][0] == '-' && argv[i][1]) {
            fprintf(stderr, "USAGE: %s [-d] [-b] [-w window] [-s stats_file] [infile [outfile]]\n", argv[0]);
            return 2;
        } else if (nfiles < 2) files[nfiles++] = argv[i];
    }
    if (files[0] && strcmp(files[0], "-") && (fd_in = open(files[0], O_RDONLY)) < 0) {
        fprintf(stderr, "%s: no such file\n", files[0]);
        return 1;
    }
    /* (read-write: the compressor fills a mappable output file through a shared mapping) */
    if (files[1] && (fd_out = open(files[1], O_RDWR | O_CREAT | O_TRUNC, 0644)) < 0) {
        fprintf(stderr, "%s: cannot create\n", files[1]);
        return 1;
    }
    if (stats && (fd_stats = open(stats, O_WRONLY | O_CREAT | O_TRUNC, 0644)) < 0) {
        fprintf(stderr, "%s: cannot create\n", stats);
        return 1;
    }
    if (!decompress) {
        rc = deflate_compress(fd_in, fd_out, fd_stats, (swi)window, 0);
    } else {
        struct string_len in = {NULL, 0}, out = {NULL, 0};
        size_t off = 0;
        if (slurp(fd_in, &in.str, &in.len)) return 1;
        rc = deflate_decompress(&out, &in, 0);
        while (!rc && off < out.len) {
            ssize_t w = write(fd_out, out.str + off, out.len - off);
            if (w < 0) { rc = -1; break; }
            off += (size_t)w;
        }
        free(in.str);
        free(out.str);
    }
    if (rc < 0) fprintf(stderr, "zencode: error %d\n", rc);
    return rc < 0 ? 1 : 0;
}
