/* oracle/ac3enc_oracle.h - CPU restatement of the reference AC-3 encoder (TEST INFRASTRUCTURE ONLY). */
#ifndef ORACLE_AC3ENC_ORACLE_H
#define ORACLE_AC3ENC_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct ora_enc ora_enc_t;
ora_enc_t * ora_enc_init (int freq, int bitrate, int channels);       /* NULL = rejected (ac3enc.cpp:1019-1072) */
void ora_enc_free (ora_enc_t * s);
int ora_enc_frame_bytes (ora_enc_t * s);
/* samples: 1536 * channels interleaved int16; chmap[coded channel] = source channel (NULL = identity) */
int ora_enc_frame (ora_enc_t * s, unsigned char * dst, const short * samples, const unsigned char * chmap);
/* intermediates of the last frame, numbered like ref_ac3enc_get() in refbuild/ac3enc_ref_wrap.cpp */
void ora_enc_get (ora_enc_t * s, int what, void * dst);
const int16_t * ora_enc_window (ora_enc_t * s);
int ora_enc_stream (int freq, int bitrate, int channels, const short * pcm, int nframes,
		    const unsigned char * chmap, unsigned char * out);
#ifdef __cplusplus
}
#endif
#endif
