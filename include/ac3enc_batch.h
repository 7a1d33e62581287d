/* include/ac3enc_batch.h - batched AC-3 encode entry points of the B200 engine (C ABI).
 *
 * One call encodes nstreams independent PCM streams of nframes frames each; a group of GPU
 * threads walks the frames of one stream in order (the reference encoder carries 256 samples per
 * channel and the warm start of its SNR-offset search from frame to frame, ac3enc.cpp:55, 921,
 * 969), streams in parallel.  Per frame the arithmetic is the reference's integer encoder
 * (ac3enc.cpp:1640-1763): frames are byte-identical to AC3_encode_frame's.
 *
 * Pointers are host pointers unless AC3_BATCH_DEVICE_PTRS is set in mem_flags.
 */
#ifndef AC3ENC_BATCH_H
#define AC3ENC_BATCH_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ac3_batch_s ac3_batch_t;

#define AC3_BATCH_DEVICE_PTRS 1

#define AC3_ST_OK            0
#define AC3_ST_NO_FIT        1   /* no SNR offset fits the frame (the reference prints "Yack" and emits an
				    undecodable frame, ac3enc.cpp:930-933); here: a valid frame of zero mantissas */

/* per-stream carry between calls (optional) */
typedef struct {
    int16_t last_samples[6][256];   /* previous 256 samples of every coded channel (ac3enc.cpp:55) */
    int32_t csnroffst;              /* warm start of the search; 0 in a fresh record means "40" (:1092) */
    int32_t started;                /* 0: fresh stream */
    int32_t reserved[2];
} ac3_stream_carry_t;

/* optional intermediate dumps for parity testing; any pointer may be NULL.  Layouts per frame:
 * coef int32 [6 blocks][6 ch][256], exp_shift int8 [6][6], strategy uint8 [6][6],
 * encoded_exp uint8 [6][6][256], bap uint8 [6][6][256], snr int32 [2] (csnroffst, fsnroffst) */
typedef struct {
    int32_t * coef;
    int8_t *  exp_shift;
    uint8_t * strategy;
    uint8_t * encoded_exp;
    uint8_t * bap;
    int32_t * snr;
} ac3_batch_debug_t;

/* one call at a time per context (it owns device scratch); one context per host thread or stream */
ac3_batch_t * ac3_batch_create (int device);
void ac3_batch_destroy (ac3_batch_t * ctx);
const char * ac3_batch_last_error (ac3_batch_t * ctx);

/* frame size in bytes for a configuration, 0 if the reference encoder would reject it
 * (AC3_encode_init, ac3enc.cpp:1019-1077) */
int ac3_batch_frame_bytes (int freq, int bitrate, int channels);

/* pcm   int16 [nstreams][nframes][1536][channels] (interleaved, as AC3_encode_frame takes it)
 * chmap channels entries or NULL (identity)
 * out   uint8 [nstreams][nframes][frame_bytes]
 * status int32 [nstreams][nframes] or NULL; carry [nstreams] in/out or NULL
 * Returns 0, or negative on a CUDA / argument error. */
int ac3_batch_encode (ac3_batch_t * ctx, const int16_t * pcm, int nstreams, int nframes,
		      int freq, int bitrate, int channels, const uint8_t * chmap,
		      uint8_t * out, int32_t * status, ac3_stream_carry_t * carry,
		      const ac3_batch_debug_t * debug, int mem_flags, void * cuda_stream);

/* ---- encoder front end: what the ACM wrapper does around AC3_encode_frame ---- */

/* WAVE channel order (FL FR FC LFE BL BR) -> coded channel order: chmap[coded channel] = source channel, for
 * 1..6 channels (reference: create_channel_map, AC3ACM.cpp:1631-1662).  Entries past `channels` are written as
 * the reference writes them; returns 0, or -1 for a channel count it does not know. */
int ac3_wav_channel_map (int channels, uint8_t chmap[6]);

/* The bitrate in kb/s that a destination format names through nAvgBytesPerSec, validated as the ACM stream
 * open does (AC3ACM.cpp:128-149, 1913-1936): 125 * one of the 19 AC-3 bitrates, or - 44.1 kHz only - the
 * rounded byte rate of that bitrate's frame; 0 = not a supported format. */
int ac3_acm_bitrate (int freq, uint32_t avg_bytes_per_sec);

/* nBlockAlign of the format (AC3ACM.cpp:128-149, 951-953): frame bytes from the wrapper's table, 44.1 kHz
 * frames unpadded; 0 for an unknown rate / bitrate. */
int ac3_acm_block_align (int freq, int kbps);

/* Streaming PCM -> AC-3 conversion with the buffering of stream_convert_pcm (AC3ACM.cpp:1665-1798): input is
 * gathered into frames of 1536 samples per channel (WAVE channel order, mapped with ac3_wav_channel_map), a
 * frame that does not fit the destination is carried over to the next call, input past a full destination
 * stays unconsumed, a trailing partial frame stays buffered.  All whole frames one call completes are encoded
 * in ONE launch.  ac3_stream_open returns NULL where the wrapper refuses the format (rate below 32 kHz,
 * :1890-1891; byte rate not in the table; configuration AC3_encode_init rejects). */
typedef struct ac3_stream_s ac3_stream_t;
ac3_stream_t * ac3_stream_open (ac3_batch_t * ctx, int freq, uint32_t avg_bytes_per_sec, int channels);
int ac3_stream_convert (ac3_stream_t * s, const void * src, uint32_t src_len, uint32_t * src_used,
			void * dst, uint32_t dst_len, uint32_t * dst_used, int start);
int ac3_stream_frame_bytes (ac3_stream_t * s);
void ac3_stream_close (ac3_stream_t * s);

long ac3_batch_launch_count (ac3_batch_t * ctx);
double ac3_batch_kernel_ms (ac3_batch_t * ctx, int * nlaunches);

#ifdef __cplusplus
}
#endif
#endif /* AC3ENC_BATCH_H */
