/*
 * include/deflate_errors.h -- error numbers of the drop-in API.
 * Same names and values as the reference (src/include/global_errors.h:24-35,
 * src/include/deflate_errors.h:7-22); functions return the NEGATED value
 * (callers test < 0, src/png/png_decoder.c:249,709).  No setjmp/longjmp.
 */
#ifndef DEFLATE_ERRORS_H
#define DEFLATE_ERRORS_H

#define E_LEN 1
#define E_MALLOC 2
#define E_FORK 3
#define E_PIPE 4
#define E_CRC 5
#define E_SZ 6
#define E_EXIST 7
#define E_NEXIST 8
#define E_NONULL 9
#define E_RANGE 10
#define E_INVAL 11
#define E_RESERV 12

#define DEFLATE_ERROR_MASK (1U << 24)
#define E_HUFAMB (DEFLATE_ERROR_MASK + 1)  /* ambiguous Huffman code */
#define E_HUFINV (DEFLATE_ERROR_MASK + 2)  /* invalid Huffman code (input) */
#define E_HUFVAL (DEFLATE_ERROR_MASK + 3)  /* invalid Huffman value (output) */
#define E_HUFDIS (DEFLATE_ERROR_MASK + 4)  /* invalid distance */
#define E_ZADL32 (DEFLATE_ERROR_MASK + 5)  /* invalid adler32 checksum */
#define E_ZHEAD (DEFLATE_ERROR_MASK + 6)   /* invalid header */
#define E_ZFCHCK (DEFLATE_ERROR_MASK + 7)  /* invalid FCHECK */
#define E_ZCMPMT (DEFLATE_ERROR_MASK + 8)  /* invalid compression method */
#define E_ZSLWIN (DEFLATE_ERROR_MASK + 9)  /* invalid sliding window */
#define E_ZPDICT (DEFLATE_ERROR_MASK + 10) /* preset dictionary not supported */
#define E_ZBSZ (DEFLATE_ERROR_MASK + 11)   /* invalid block size */
#define E_ZNLEN (DEFLATE_ERROR_MASK + 12)  /* LEN / NLEN mismatch */
#define E_ZINV (DEFLATE_ERROR_MASK + 13)   /* invalid compression metadata */
#define E_ZBTYPE (DEFLATE_ERROR_MASK + 14) /* invalid block type */
#define E_GPU (DEFLATE_ERROR_MASK + 100)   /* CUDA device / runtime failure (new) */

#endif
