/* oracle/refbuild/ac3enc_ref_wrap.cpp - TEST INFRASTRUCTURE ONLY.
 *
 * Wrapper translation unit around the UNMODIFIED reference encoder
 * (src/ac3enc/ac3enc.cpp under /root/reference, included at build time; the
 * output lives in oracle/_ref/, git-ignored).  The reference source assumes
 * Win64 (32-bit `long`: put_bits stores through `unsigned long *`,
 * ac3enc.cpp:168, and bswap takes `unsigned long`, :101-108), so the libc
 * headers are included first and `long` is then re-defined to `int` for the
 * duration of the include.  `_M_AMD64` selects the portable bswap.
 * windows.h / crtdbg.h are the shims next to this file.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <assert.h>
#include <stdint.h>

#define _M_AMD64 1
#define long int
#include "ac3enc.cpp"
#undef long

#define REF_API extern "C" __attribute__((visibility("default")))

/* forget every cross-frame carry (last_samples, warm-start csnroffst) */
REF_API int ref_ac3enc_init (int freq, int bitrate, int channels)
{
    memset (&ac3enc_state, 0, sizeof (ac3enc_state));
    return AC3_encode_init (freq, bitrate, channels);
}

/* dst must hold >= 3840 + 8 bytes (put_bits stores whole 32-bit words) */
REF_API int ref_ac3enc_frame (unsigned char * dst, short * samples, unsigned char * chmap)
{
    return AC3_encode_frame (dst, samples, chmap);
}

/* Encode a whole stream: pcm = nframes*1536*nch interleaved int16,
 * out = nframes*frame_bytes.  Returns frame_bytes (0 on bad config). */
REF_API int ref_ac3enc_stream (int freq, int bitrate, int channels,
			       const short * pcm, int nframes,
			       const unsigned char * chmap, unsigned char * out)
{
    unsigned char tmp[3840 + 64];
    unsigned char idmap[6] = {0, 1, 2, 3, 4, 5};
    int fb = ref_ac3enc_init (freq, bitrate, channels);
    if (fb <= 0) return 0;
    for (int f = 0; f < nframes; f++) {
	memset (tmp, 0, sizeof (tmp));
	AC3_encode_frame (tmp, (short *) pcm + (size_t) f * 1536 * channels,
			  (unsigned char *) (chmap ? chmap : idmap));
	memcpy (out + (size_t) f * fb, tmp, fb);
    }
    return fb;
}

/* golden-vector accessors for the file-static intermediates
 * (ac3enc.cpp:80-87), valid after ref_ac3enc_frame */
REF_API void ref_ac3enc_get (int what, void * dst)
{
    switch (what) {
    case 0: memcpy (dst, mdct_coef, sizeof (mdct_coef)); break;       /* int32 [6][6][256] */
    case 1: memcpy (dst, exponent, sizeof (exponent)); break;         /* u8 [6][6][256] */
    case 2: memcpy (dst, exp_strategy, sizeof (exp_strategy)); break; /* u8 [6][6] */
    case 3: memcpy (dst, encoded_exp, sizeof (encoded_exp)); break;   /* u8 [6][6][256] */
    case 4: memcpy (dst, bap, sizeof (bap)); break;                   /* u8 [6][6][256] */
    case 5: memcpy (dst, exp_samples, sizeof (exp_samples)); break;   /* s8 [6][6] */
    case 7: memcpy (dst, ac3_window, sizeof (ac3_window)); break;       /* s16 [256] */
    case 8: { short * q = (short *) dst; memcpy (q, costab, 128); memcpy (q + 64, sintab, 128);
	      memcpy (q + 128, xcos1, 256); memcpy (q + 256, xsin1, 256); } break;
    case 6: { int * p = (int *) dst; p[0] = ac3enc_state.csnroffst;
	      p[1] = ac3enc_state.fsnroffst[0]; p[2] = ac3enc_state.frame_size; } break;
    }
}
