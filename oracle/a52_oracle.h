/* oracle/a52_oracle.h - CPU restatement of the reference AC-3 decode path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is linked, imported or
 * executed by the product path (ac-3-acm-codec_b200/); only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may use it, and only as the checker.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this restatement
 * against the unmodified reference built by `make -C oracle ref`
 * (oracle/_ref/liba52_ref.so) - exponents, baps, coefficients and PCM
 * bit-exact, lfsr_state equal - and against the committed golden vectors in
 * tests/golden/ that were generated from that same reference build.
 */
#ifndef ORACLE_A52_ORACLE_H
#define ORACLE_A52_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ora_dec ora_dec_t;

ora_dec_t * ora_init (void);
void ora_free (ora_dec_t * d);
float * ora_samples (ora_dec_t * d);
int ora_syncinfo (const uint8_t * buf, int * flags, int * sample_rate, int * bit_rate);
int ora_frame (ora_dec_t * d, const uint8_t * buf, int * flags, float * level, float bias);
void ora_dynrng_off (ora_dec_t * d);
int ora_block (ora_dec_t * d);

/* intermediates of the most recent ora_block(); which: 0..4 fbw, 5 lfe, 6 cpl.
 * bap is returned in liba52's private numbering so it compares 1:1 with the
 * reference state. */
void ora_get_expbap (ora_dec_t * d, int which, uint8_t * exp, int8_t * bap);
/* info[] layout identical to ref_get_info() in refbuild/a52_ref_wrap.c */
void ora_get_info (ora_dec_t * d, int * info);
void ora_set_lfsr (ora_dec_t * d, int v);
/* dequantised, gain-applied coefficients of the last block before any
 * mixing/transform: planes 0..4 fbw channels, plane 5 lfe. */
void ora_get_coeffs (ora_dec_t * d, float * coef /* [6][256] */);

void ora_bit_allocate (int fscod, int halfrate, int bai11, int csnroffst, int chbai,
		       int deltbae, const int8_t * deltba, int bndstart, int start,
		       int end, int fastleak, int slowleak, const uint8_t * exp,
		       int8_t * bap /* liba52 numbering */);
void ora_imdct (int kind, float * data, float * delay, float bias);

long ora_decode_stream (ora_dec_t * d, const uint8_t * es, long nbytes, int req_flags,
			float level, float bias, float * out, int nout, int dynrng_off);

#ifdef __cplusplus
}
#endif
#endif
