/* dfl_common.cuh -- constants, per-call device state and the small device helpers the kernels share
 * (included by dfl_kernels.cu, the single translation unit of libdeflate_b200.so). */
#pragma once

#define FULL 0xFFFFFFFFu
#define NLL 286
#define ND 30
#define NCL 19
#define LLS 288          /* stride of per-segment lit/len tables */
#define DS 32            /* stride of per-segment distance tables */
#define HDR_WORDS 160    /* 640 B >= 3+14+57+316*14 bits */
#define NPOS_MAX 65536
#define HASH_SIZE DFL_HASH_BUCKETS
#define MATCH_THREADS 1024
#ifndef REFILL_MIN_IDLE
#define REFILL_MIN_IDLE 8 
#endif
#define DFL_CHUNK_BYTES (256u << 20)
#define ADLER_MOD 65521u

enum { K_MATCH = 0, K_PARSE, K_HUFF, K_SCAN, K_EMIT, K_FINAL, K_INFLATE, K_INF_RESOLVE, K_INF_FINISH, K_PNG_FILTER, K_NUM };
static const char *k_names[K_NUM] = {"k_match", "k_parse", "k_huff", "k_scan", "k_emit", "k_final", "k_inflate", "k_inf_resolve", "k_inf_finish", "k_png_filter"};

struct DevState {
    unsigned long long total_out;  /* bytes of deflate body written so far */
    unsigned long long adl_a;      /* sum of per-segment byte sums (mod 65521 each) */
    unsigned long long adl_b;      /* sum of per-segment weighted sums */
    unsigned int error;            /* 1 = output overflow */
    unsigned int adler;            /* final Adler-32 */
    unsigned long long t_phase[24]; /* k_match SM cycles summed over CTAs: stage-in, build, search, total */
    unsigned int sort_fallbacks;   /* segments whose index had to be rebuilt with the ballot-ranked scatter */
    unsigned int inf_pending;      /* k_inf_resolve: some word is still a pointer */
    unsigned int run_next, run_done; /* k_match_run: next run of segments to hand out, CTAs that have finished (the last one resets both) */
};

struct SegArrays {
    uint32_t *ll_freq;   /* [nseg][LLS] */
    uint32_t *d_freq;    /* [nseg][DS] */
    uint32_t *ntok;      /* [nseg] */
    uint32_t *extra;     /* [nseg] sum of length/distance extra bits */
    uint32_t *lltab;     /* [nseg][LLS]  code_rev | len << 16 of the chosen block type */
    uint32_t *dtab;      /* [nseg][DS] */
    uint32_t *hdr;       /* [nseg][HDR_WORDS] packed block header bits */
    uint32_t *hdr_bits;  /* [nseg] */
    uint32_t *btype;     /* [nseg] */
    uint32_t *out_bytes; /* [nseg] */
    uint32_t *cost;      /* [nseg][2]: bits of the chosen block's code description, bits of its distance symbols + extra bits */
    unsigned long long *block_bits; /* [nseg] */
    unsigned long long *out_off;    /* [nseg] */
};

/* ------------------------------------------------------------------------- */
/* small device helpers                                                      */
/* ------------------------------------------------------------------------- */

__device__ __forceinline__ uint32_t lanemask_lt()
{
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

/* unaligned little-endian 4-byte load from a word-aligned shared array */
__device__ __forceinline__ uint32_t ld4(const uint32_t *w, uint32_t byte_idx)
{
    uint32_t i = byte_idx >> 2;
    uint32_t lo = w[i], hi = w[i + 1];
    return __funnelshift_r(lo, hi, (byte_idx & 3) * 8);
}

/* Explicit shared-window addressing for the hot loops: a 32-bit shared address computed once
 * (ptxas otherwise re-derives the window base with S2R SR_CgaCtaId inside the loops). */
__device__ __forceinline__ uint32_t lds_u32(uint32_t a)
{
    uint32_t v;
    asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t a)
{
    uint32_t v;
    asm("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t lds_u8(uint32_t a)
{
    uint32_t v;
    asm("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts_u16(uint32_t a, uint32_t v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t atoms_add(uint32_t a, uint32_t v)
{
    uint32_t o;
    asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(o) : "r"(a), "r"(v) : "memory");
    return o;
}
__device__ __forceinline__ void reds_add(uint32_t a, uint32_t v) { asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
/* unaligned little-endian 4-byte load at shared address a */
__device__ __forceinline__ uint32_t ld4s(uint32_t a)
{
    const uint32_t w = a & ~3u;
    return __funnelshift_r(lds_u32(w), lds_u32(w + 4), a << 3);
}

/* DFL_HASH4 with the group width chosen at run time: sb = 0 gives plain trigram buckets (DFL_HASH3),
 * sb = DFL_SUB_BITS the 4-byte buckets.  ngrp = DFL_HASH_BUCKETS >> sb. */
__device__ __forceinline__ uint32_t hash4(uint32_t v, uint32_t ngrp, int sb)
{
    const uint32_t grp = __umulhi((v & 0xFFFFFFu) * 0x9E3779B1u, ngrp);
    return (grp << sb) | __funnelshift_l((v >> 24) * 0x9E3779B1u, 0u, sb);
}

/* RFC 1951 3.2.5: length 3..258 -> symbol 257..285 (same numbers as
 * get_len_code, src/deflate_compress.c:182-200; extra bits per the RFC). */
__device__ __forceinline__ int len_sym(int len, int &eb, int &ev)
{
    int t = len - 3;
    if (len == 258) { eb = 0; ev = 0; return 285; }
    if (t < 8) { eb = 0; ev = 0; return 257 + t; }
    int k = 31 - __clz(t);
    eb = k - 2;
    ev = t & ((1 << eb) - 1);
    return 261 + 4 * eb + ((t >> eb) & 3);
}

/* distance 1..32768 -> symbol 0..29 (get_dist_code, src/deflate_compress.c:202-217) */
__device__ __forceinline__ int dist_sym(int dist, int &eb, int &ev)
{
    int t = dist - 1;
    if (t < 4) { eb = 0; ev = 0; return t; }
    int k = 31 - __clz(t);
    eb = k - 1;
    ev = t & ((1 << eb) - 1);
    return 2 * k + ((t >> eb) & 1);
}

__device__ __forceinline__ int fixed_ll_len(int s) { return s < 144 ? 8 : s < 256 ? 9 : s < 280 ? 7 : 8; }

/* mbarrier + TMA bulk copy (global -> shared), sm_90+ PTX */
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    uint32_t a = (uint32_t)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    uint32_t a = (uint32_t)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
    uint32_t d = (uint32_t)__cvta_generic_to_shared(dst_smem);
    uint32_t b = (uint32_t)__cvta_generic_to_shared(bar);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(d), "l"(src_gmem), "r"(bytes), "r"(b)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t a = (uint32_t)__cvta_generic_to_shared(bar);
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(a), "r"(parity)
        : "memory");
}
