/* include/a52_batch.h - batched entry points of the B200 AC-3 engine (C ABI).
 *
 * One call decodes thousands of independent AC-3 streams: every stream is a
 * run of sync frames walked by one pair of GPU warps, frames of one
 * stream in order (overlap-add tail and dither generator carried on chip),
 * streams in parallel.  This is the batched form of the per-frame loop every
 * liba52 caller writes (reference: a52dec.c:240-309 - a52_syncinfo, a52_frame,
 * [a52_dynrng], 6 x a52_block, copy a52_samples()), with the same arithmetic
 * per frame as a52_frame/a52_block (liba52/parse.c:131-205, 558-940).
 *
 * All functions use plain pointers and sizes; no CUDA or torch types appear.
 * Pointers are host pointers unless A52_BATCH_DEVICE_PTRS is set in
 * `mem_flags`, in which case es / frame_off / stream_first / pcm_out /
 * frame_status / frame_flags / state are device pointers (es 16-byte aligned
 * with >= 16 readable bytes after es_bytes) and no copy is made.
 */
#ifndef A52_BATCH_H
#define A52_BATCH_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct a52_batch_s a52_batch_t;

/* PCM layouts (per frame: 6 blocks x 256 samples x nout channels, channel
 * order as liba52: [LFE] then the granted mode's channels) */
#define A52_PCM_F32_PLANAR      0   /* float [6][nout][256]   == 6 x a52_samples() */
#define A52_PCM_F32_INTERLEAVED 1   /* float [1536][nout] */
#define A52_PCM_S16_INTERLEAVED 2   /* int16 [1536][nout], round-to-nearest of x*32768, saturated
				       (== libao convert2s16.c:33-41 applied to bias-384 floats) */
#define A52_PCM_S16_WAV         3   /* int16 [6][256][nout]: block by block in WAV channel order exactly as libao's
				       convert2s16_wav lays it out (convert2s16.c:199-306: L R C LFE SL SR), the
				       2F1R+LFE case with its fall-through into the 3F1R+LFE layout included
				       (:270-285: five-sample groups L S R LFE -32768, cut after 1024 values) */

/* OR-ed into req_flags (with A52_LFE / A52_ADJUST_LEVEL as wanted): every frame is decoded in its own coded
 * mode, i.e. the request is the flags a52_syncinfo reports for that frame - what libao's wav6 driver does by
 * leaving *flags alone (audio_out_wav.c:69-70, 211-214).  Frames keep the stride of six channels. */
#define A52_REQ_AS_CODED        0x100

#define A52_BATCH_DEVICE_PTRS   1

/* dynamic range control */
#define A52_DRC_STREAM          0   /* apply the stream's dynrng words (liba52 default) */
#define A52_DRC_OFF             1   /* == a52_dynrng (state, NULL, NULL) */
#define A52_DRC_TABLE           2   /* every dynrng word stands for the range a52_batch_set_drc_table() gave it:
				       == a52_dynrng (state, call, data) with the callback run by the caller
				       between a52_batch_scan() and a52_batch_decode() */

/* per-frame status written to frame_status[] */
#define A52_ST_OK               0
#define A52_ST_BAD_SYNC         1   /* a52_syncinfo would return 0 (parse.c:98-127) */
#define A52_ST_BAD_FRAME        2   /* a52_frame would return 1 (parse.c:163-164) */
#define A52_ST_BAD_BLOCK        16  /* + block index: a52_block returned 1 in that block */

/* per-stream carry between calls (optional): dither generator position and
 * the overlap-add tail of every output channel */
typedef struct {
    uint32_t dither_index;      /* number of dither_gen() calls so far, mod 65535 */
    uint32_t per_channel;       /* 0: delay planes 0..nout-1 hold the downmixed tail (liba52 downmixed=1,
				   the initial state); 1: planes 0..nfchans-1 hold per-coded-channel tails */
    uint32_t reserved[2];
    float    delay[6][128];     /* overlap-add tails: planes 0..4 main channels, plane 5 LFE */
} a52_stream_carry_t;

/* optional intermediate dumps for parity testing; any pointer may be NULL.
 * Layouts: exp/bap uint8 [nframes][6 blocks][7][256] with index 0..4 fbw,
 * 5 lfe, 6 coupling channel, bap in the A/52 standard's numbering 0..15;
 * coef float [nframes][6][6][256] = dequantised, gain-applied coefficients
 * before mixing (planes 0..4 fbw, 5 lfe); info int32 [nframes][6][16]
 * (endmant[5], cplstrtmant, cplendmant, chincpl, lfsr_state, acmod, lfeon,
 * output, 0, ncplbnd, rematflg, csnroffst). */
typedef struct {
    uint8_t * exp;
    uint8_t * bap;
    float *   coef;
    int32_t * info;
} a52_batch_debug_t;

/* A context owns device scratch (work counters, carry and locate-stage scratch, staging buffers): like an
 * a52_state_t (liba52 is re-entrant per state, one state = one thread at a time, NEWS:3) it serves ONE call at
 * a time.  Asynchronous device-pointer calls of one context must be issued on one CUDA stream (or be ordered
 * by the caller); use one context per host thread or stream for concurrency. */
a52_batch_t * a52_batch_create (int device);
void a52_batch_destroy (a52_batch_t * ctx);
const char * a52_batch_last_error (a52_batch_t * ctx);

/* Host-side frame indexer: scans an elementary stream for sync frames exactly
 * like the resync loop of a52dec.c:240-309 and writes the byte offset of each
 * frame.  Returns the number of frames found (<= max_frames). */
int a52_batch_index (const uint8_t * es, size_t es_bytes, uint64_t * frame_off, int max_frames);

/* The same scan on the GPU for elementary streams that already live in device memory (one thread walks
 * one stream): stream s = bytes [stream_off[s], stream_off[s+1]) of es.  es, frame_off (exactly the
 * frames found are written, at most max_frames) and stream_first (nstreams + 1 entries) are DEVICE
 * pointers, stream_off is a host pointer.  Returns the total number of frames, or a negative value
 * (frame table too small, CUDA error). */
int a52_batch_index_device (a52_batch_t * ctx, const uint8_t * es, const uint64_t * stream_off, int nstreams,
			    uint64_t * frame_off, int max_frames, uint32_t * stream_first, void * cuda_stream);

/* What a frame holds without decoding it to PCM: a pass over the side information, the exponents and the bit
 * allocation only (parse.c:558-804 + the counting half of :336-433). */
typedef struct {
    uint32_t dither_draws;	/* dither_gen() calls the frame makes (parse.c:310-319) */
    int16_t  dynrng[6][2];	/* the 8-bit dynrng word(s) of every block as coded (parse.c:578-598; [1] is the
				   second word of a 1+1 frame), -1 = the block carries none */
    int32_t  status;		/* A52_ST_* the decode would report */
} a52_frame_scan_t;

/* Scan nframes frames (arguments as for a52_batch_decode; scan[nframes] is a device pointer under
 * A52_BATCH_DEVICE_PTRS).  Returns 0 or a negative error. */
int a52_batch_scan (a52_batch_t * ctx, const uint8_t * es, size_t es_bytes,
		    const uint64_t * frame_off, int nframes,
		    const uint32_t * stream_first, int nstreams, int req_flags,
		    a52_frame_scan_t * scan, int mem_flags, void * cuda_stream);

/* The ranges calls with drc_mode A52_DRC_TABLE apply: float [nframes][6][2], entry [f][b][k] for word k of block b
 * of frame f as a52_batch_scan() reported it (entries of absent words are not read).  The range of a word d
 * (signed 8 bit) as liba52 computes it before the callback is (((d & 0x1f) | 0x20) << 13) * 2^-(18 - (d >> 5))
 * (parse.c:586-591).  Host pointer, or device pointer for A52_BATCH_DEVICE_PTRS calls; must stay valid until the
 * decode call has consumed it. */
void a52_batch_set_drc_table (a52_batch_t * ctx, const float * ranges);

/* How a stream longer than a work unit (32 frames) is cut: 0 = choose (default), 1 = slices chained by the carry
 * record (a stream is decoded by one pair of warps at a time: best when streams outnumber the GPU's resident
 * pairs), 2 = frame-independent slices (all slices of a stream run side by side after a scan pass and a prefix
 * sum of the dither draws; every slice decodes one frame of look-back for the overlap-add state: best for few
 * long streams).  Both give the same PCM for conforming streams. */
void a52_batch_set_slice_mode (a52_batch_t * ctx, int mode);

/* bytes one decoded frame occupies in pcm_out for a request (req_flags as for
 * a52_frame): 1536 * nout_requested * sample size */
size_t a52_batch_frame_stride (int req_flags, int out_fmt);

/* Decode nframes frames forming nstreams streams.
 *   frame_off[nframes]      byte offset of each frame in es (frame length is
 *                           derived from its own header as a52_syncinfo does)
 *   stream_first[nstreams+1] index of the first frame of each stream
 *   req_flags, level, bias  as for a52_frame (A52_ADJUST_LEVEL honoured)
 *   pcm_out                 nframes * a52_batch_frame_stride() bytes
 *   frame_status[nframes]   A52_ST_* (may be NULL)
 *   frame_flags[nframes]    granted output flags per frame (may be NULL)
 *   carry[nstreams]         in/out per-stream carry (may be NULL: streams
 *                           start from liba52's initial state)
 *   cuda_stream             a cudaStream_t (NULL = default stream)
 * Returns 0 on success, negative on a CUDA/argument error (see
 * a52_batch_last_error).  Asynchronous when A52_BATCH_DEVICE_PTRS is set. */
int a52_batch_decode (a52_batch_t * ctx,
		      const uint8_t * es, size_t es_bytes,
		      const uint64_t * frame_off, int nframes,
		      const uint32_t * stream_first, int nstreams,
		      int req_flags, float level, float bias, int drc_mode,
		      int out_fmt, void * pcm_out,
		      int32_t * frame_status, int32_t * frame_flags,
		      a52_stream_carry_t * carry,
		      const a52_batch_debug_t * debug,
		      int mem_flags, void * cuda_stream);

/* Optional: upper bound of the frame length (bytes) in the next batches, so that
 * device-pointer calls need no length pre-pass; 0 = derive it (default). */
void a52_batch_set_max_frame_bytes (a52_batch_t * ctx, int nbytes);

/* Optional: frames of the longest stream in the next batches, so that device-pointer calls need no
 * pre-pass to plan the work units (slices of streams); 0 = derive it (default). */
void a52_batch_set_max_stream_frames (a52_batch_t * ctx, int nframes);

/* Bounds-checked builds (-DA52_BOUNDS_CHECK): code of the first violated address / index check of any decode
 * launch of this process, 0 = none.  Regular builds carry no checks and return -2. */
int a52_batch_violations (void);

/* Experiment entry (profiles/r02_imdct_tc_ab.json): the IMDCT-512 of nplanes coefficient planes,
 * global to global, by the production FFT transform (variant 0) or as a 3xTF32 tensor-core GEMM (variant 1; afrag = the
 * transform matrix in mma fragment order, built by tools/dev_imdct_ab.py).  Not part of the decode path. */
int a52_ab_imdct (a52_batch_t * ctx, int variant, const float * x, float * y, int nplanes, const void * afrag,
		  void * cuda_stream);

/* Measured FP32 FMA peak of the context's device in TFLOP/s (a register-only FMA loop; for bench.py's FP32 roofline). */
double a52_ab_fp32_peak (a52_batch_t * ctx);

/* number of kernel launches issued by this context so far (bench bookkeeping) */
long a52_batch_launch_count (a52_batch_t * ctx);
/* average device time (ms) of the decode kernel over the launches since the
 * last call, measured with CUDA events on the launching stream; resets. */
double a52_batch_kernel_ms (a52_batch_t * ctx, int * nlaunches);

#ifdef __cplusplus
}
#endif
#endif /* A52_BATCH_H */
