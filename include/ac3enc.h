/* include/ac3enc.h - drop-in encoder API of the B200 AC-3 engine.
 *
 * Same two entry points as the reference encoder's header
 * (reference: src/ac3enc/ac3enc.h:6-7; callers AC3ACM.cpp:1762, 1940):
 *   AC3_encode_init  (freq, bitrate in bit/s, channels 1..6) -> frame bytes, 0 = rejected
 *   AC3_encode_frame (dst, 1536 * channels interleaved int16, chmap) -> bytes written
 * chmap[coded channel] = source channel of the interleaved input (ac3enc.cpp:1676;
 * WAV order -> AC-3 order maps at AC3ACM.cpp:1631-1662).  Like the reference, the encoder is a
 * process-wide singleton and not thread-safe.  The encode itself runs on the GPU through the
 * batched path (a batch of one stream, one frame); there is no CPU fallback: AC3_encode_init
 * returns 0 when no CUDA device is usable.
 */
#ifndef AC3ENC_H
#define AC3ENC_H

#ifdef __cplusplus
extern "C" {
#endif

int AC3_encode_init (int freq, int bitrate, int channels);
int AC3_encode_frame (unsigned char * dst, short * samples, unsigned char * chmap);

#ifdef __cplusplus
}
#endif
#endif /* AC3ENC_H */
