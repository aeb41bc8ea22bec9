/*
 * host/synth.c -- synthetic corpora of BASELINE.json's configs (SURVEY.md
 * section 8(d)): bench / test support, deterministic, no file access.
 *   random : splitmix64 output bytes (config 4; same stream as k_synth_random)
 *   markov : order-2 character Markov chain over a model table derived from the
 *            reference's sample text (configs 1, 2; fixture tests/golden/markov2.npz)
 *   log    : 64 template lines with one 8-hex-digit counter field (config 3)
 *   png    : filtered RGB scanlines, filter types Sub/Up/Paeth cycling (config 5;
 *            filter ids as in src/png/include/png_decoder.h:57-61, PNG-spec Paeth)
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static uint64_t sm64_next(uint64_t *s)
{
    uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

void dfl_synth_random(uint8_t *out, size_t n, uint64_t seed)
{
    uint64_t s = seed;
    size_t i;
    for (i = 0; i + 8 <= n; i += 8) {
        uint64_t z = sm64_next(&s);
        memcpy(out + i, &z, 8);
    }
    if (i < n) {
        uint64_t z = sm64_next(&s);
        memcpy(out + i, &z, n - i);
    }
}

/*
 * ctx_off[65537]: entries of context (c0 << 8 | c1) are [ctx_off[c], ctx_off[c+1])
 * nxt[k]: successor byte; cum[k]: inclusive cumulative count inside the context.
 * A context without successors restarts from start_ctx.
 */
void dfl_synth_markov(uint8_t *out, size_t n, uint64_t seed, const uint32_t *ctx_off, const uint8_t *nxt,
                      const uint32_t *cum, uint32_t start_ctx)
{
    uint64_t s = seed;
    uint32_t ctx = start_ctx & 0xFFFF;
    size_t i = 0;
    if (n > 0) out[i++] = (uint8_t)(ctx >> 8);
    if (n > 1) out[i++] = (uint8_t)ctx;
    while (i < n) {
        uint32_t lo = ctx_off[ctx], hi = ctx_off[ctx + 1], total, r, k;
        if (lo == hi) { ctx = start_ctx & 0xFFFF; lo = ctx_off[ctx]; hi = ctx_off[ctx + 1]; }
        total = cum[hi - 1];
        r = (uint32_t)(sm64_next(&s) % total);
        k = lo;
        if (hi - lo > 8) {
            uint32_t a = lo, b = hi - 1;
            while (a < b) {
                uint32_t m = (a + b) >> 1;
                if (cum[m] > r) b = m; else a = m + 1;
            }
            k = a;
        } else {
            while (cum[k] <= r) k++;
        }
        out[i++] = nxt[k];
        ctx = ((ctx & 0xFF) << 8) | nxt[k];
    }
}

/* 64 templates cut from `text` (>= 64*300 bytes of Markov text); line r is
 * template[r % 64] with an 8-hex-digit field holding r. */
void dfl_synth_log(uint8_t *out, size_t n, uint64_t seed, const uint8_t *text, size_t text_len)
{
    static const char hexd[] = "0123456789abcdef";
    uint8_t tmpl[64][304];
    int tlen[64], foff[64], t, j;
    uint64_t s = seed;
    size_t i = 0;
    uint32_t r = 0;
    for (t = 0; t < 64; t++) {
        size_t base = ((size_t)t * 300) % (text_len > 300 ? text_len - 300 : 1);
        tlen[t] = 120 + (int)(sm64_next(&s) % 181);
        for (j = 0; j < tlen[t] - 1; j++) {
            uint8_t c = text[base + j];
            tmpl[t][j] = (c == '\n' || c == '\r') ? ' ' : c;
        }
        tmpl[t][tlen[t] - 1] = '\n';
        foff[t] = (int)(sm64_next(&s) % (uint64_t)(tlen[t] - 9));
    }
    while (i < n) {
        uint8_t line[304];
        size_t c;
        t = (int)(r & 63);
        memcpy(line, tmpl[t], (size_t)tlen[t]);
        for (j = 0; j < 8; j++) line[foff[t] + j] = (uint8_t)hexd[(r >> (28 - 4 * j)) & 15];
        c = (size_t)tlen[t] < n - i ? (size_t)tlen[t] : n - i;
        memcpy(out + i, line, c);
        i += c;
        r++;
    }
}

static int paeth(int a, int b, int c)
{
    int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

/* Filtered scanlines of a 4096-pixel RGB image: smooth 2-D gradient + uniform
 * noise in [-2, 2]; each 12288-byte row is prefixed by its filter type byte. */
void dfl_synth_png(uint8_t *out, size_t n, uint64_t seed)
{
    enum { WPX = 4096, ROW = WPX * 3 };
    static const uint8_t types[3] = {1, 2, 4};
    uint8_t *cur = (uint8_t *)malloc(ROW), *prev = (uint8_t *)calloc(ROW, 1), *tmp;
    uint64_t s = seed;
    size_t i = 0;
    uint32_t y = 0;
    if (!cur || !prev) { free(cur); free(prev); return; }
    while (i < n) {
        uint8_t ft = types[y % 3];
        int x, c;
        for (x = 0; x < WPX; x++) {
            uint64_t z = sm64_next(&s);
            for (c = 0; c < 3; c++) {
                int base = (int)((x * (c + 1)) / 16 + (y * (3 - c)) / 8 + c * 40);
                int noise = (int)((z >> (8 * c)) % 5) - 2;
                cur[x * 3 + c] = (uint8_t)(base + noise);
            }
        }
        if (i < n) out[i++] = ft;
        for (x = 0; x < ROW && i < n; x++) {
            int a = x >= 3 ? cur[x - 3] : 0, b = prev[x], cc = x >= 3 ? prev[x - 3] : 0, pred;
            pred = ft == 1 ? a : ft == 2 ? b : paeth(a, b, cc);
            out[i++] = (uint8_t)(cur[x] - pred);
        }
        tmp = prev; prev = cur; cur = tmp;
        y++;
    }
    free(cur);
    free(prev);
}
