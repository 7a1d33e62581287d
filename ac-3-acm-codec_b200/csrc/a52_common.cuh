// a52_common.cuh - shared definitions of the B200 AC-3 decode engine.
//
// Data layout, constant tables and small device helpers used by the decode
// kernel (a52_decode.cu).  sm_100a only.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "ac3_tables.h"

namespace a52 {

constexpr int kDitherPeriod = 65535;
constexpr int kDitherWrap = 4096;

// output mode ids == liba52's A52_* flag values (include/a52.h)
enum { M_CHANNEL = 0, M_MONO, M_STEREO, M_3F, M_2F1R, M_3F1R, M_2F2R, M_3F2R,
       M_CHANNEL1, M_CHANNEL2, M_DOLBY, M_MASK = 15, M_LFE = 16, M_ADJUST = 32 };

// ---- read-only tables, one copy per CTA in shared memory -------------------
// (lane-indexed lookups: shared memory, not __constant__, because a constant
// load with 32 different addresses serialises)
struct __align__(16) Tables {
    float    window[256];      // KBD alpha=5 (imdct.c:364-372)
    float2   pre1[128];        // natural-order pre-twiddle of the 512 transform, sign folded
    float2   post1[64];
    float2   pre2[64];
    float2   post2[32];
    float2   wfft[128];        // e^{-2 pi j k / 128}
    // the same twiddles with every factor duplicated, (x, x, y, y): operands of the packed two-channel
    // transform (fma.rn.f32x2 has no scalar-broadcast form, so the broadcast lives in the table)
    float4   pre1d[128];
    float4   wfftd[128];
    float4   post1d[64];
    float4   win2d[128];       // (w[2q], w[2q], w[2q+1], w[2q+1])
    uint4    fftaddr[32];      // per lane: swizzled scratch byte offsets of the paired transform (see imdct512_pair)
    int16_t  q1[3][32];        // grouped 3-level values by (digit, 5-bit code)
    int16_t  q2[3][128];       // grouped 5-level values by (digit, 7-bit code)
    int16_t  q4[2][128];       // grouped 11-level values by (digit, 7-bit code)
    int16_t  q35[24];          // 7-level at [0..7], 15-level at [8..23]
    uint16_t dither_lut[256];  // CRC-16/0xA011 byte step (tables.h:213-246)
    uint16_t exp_lut[128];     // 7-bit exponent group code -> the three deltas + 2 as nibbles, bit 15: not a code (>= 125)
    uint32_t cnt_lut32[20];    // per bap: 5-bit counters n1 | n2 << 5 | n4 << 10 | zero << 15, plain field bits << 20
    uint4    emit_lut[32];     // per bap (+16: bap-0 mantissas of this run are dithered), see build_tables()
    uint4    emit_lut2[32];    // same index: where a mantissa's plan entry goes and what its position word holds
    uint16_t hth[3 * 50];
    uint8_t  masktab[256];
    uint8_t  latab[256];
    uint8_t  baptab[64];
    uint8_t  bndtab[52];
    uint8_t  bap_bits[16];
    uint8_t  pad_[8];
};

// ---- per-(requested output, acmod, mix levels) constants, host-computed ----
struct ModeEntry {            // result of liba52's a52_downmix_init for one BSI combination
    int32_t output;           // granted mode (without LFE bit), -1 = invalid request
    float   level;            // adjusted level (A52_ADJUST_LEVEL folded in), before the x2
};
struct MixEntry {             // which coded channels feed each output channel
    uint8_t nout;
    uint8_t pos[5];
    uint8_t neg[5];
    uint8_t up[5];            // a52_upmix: output plane that moves to coded channel ch (0xff = zeroed)
};

// ---- kernel parameters -----------------------------------------------------
struct StreamCarry {          // == a52_stream_carry_t (include/a52_batch.h)
    uint32_t dither_index;
    uint32_t per_channel;
    uint32_t reserved[2];
    float    delay[6][128];
};

struct FrameScan {            // == a52_frame_scan_t (include/a52_batch.h)
    uint32_t dither_draws;
    int16_t  dynrng[6][2];    // -1: the block carries no word
    int32_t  status;
};

struct DecodeParams {
    const uint8_t*  es;
    uint64_t        es_bytes;
    const uint64_t* frame_off;
    const uint32_t* stream_first;
    int             nstreams;
    int             nframes;         // entries of frame_off
    int             req_flags;
    float           bias;
    int             drc_off;
    int             out_fmt;
    int             nout_req;        // channels of the requested mode (+LFE)
    size_t          frame_stride;    // bytes per decoded frame in pcm
    uint8_t*        pcm;
    int32_t*        status;
    int32_t*        frame_flags;
    StreamCarry*    carry;
    const StreamCarry* carry_in;     // where the initial state is read (frame-independent slices: a copy, since the
                                     // last slice of a stream writes carry[] while its first may not have started)
    const uint16_t* dither_seq;      // state after n dither_gen() calls from seed 1, n = 0..65534, then wrapped
                                     // (kDitherWrap more entries) so that a block never takes a modulo
    int*            work_counter;
    int             fbuf_bytes;      // bytes of the staged-frame buffer (multiple of 16)
    int             warp_bytes;      // shared-memory bytes per warp
    int             nplanes;         // coefficient planes per warp: 5, or 6 when the LFE is requested
    int             group_threads;   // threads walking one stream (64: a pair of warps)
    int             slice_frames;    // pair kernel: frames per work unit
    int             nslices;         // pair kernel: work units per stream (1 = whole streams)
    int             carry_init;      // carry[] holds the caller's initial state (else streams start fresh)
    uint8_t*        plan;            // [resident pairs][kPlanBytes]: plan of the last full locate pass
    int*            slice_done;      // [nstreams] slices completed per stream
    // scan pass (a52_batch_scan): side information, exponents, bit allocation and the locate counts only - what
    // the dither generator drew in every frame and which dynrng words its blocks carry; no PCM
    int             scan_only;
    FrameScan*      scan;            // [nframes]
    // frame-independent slices: every slice of a stream starts on its own, one frame early (that frame rebuilds the
    // overlap-add tails and the tail representation; its PCM goes to a scratch frame), with the dither generator
    // position a scan pass and a prefix sum gave it.  No slice waits for another.
    int             lockstep;        // pairs of a CTA start every frame together (see the frame gate in the kernel)
    int             indep;
    const uint32_t* slice_dither;    // [nstreams][nslices] generator position at the slice's first decoded frame
    uint8_t*        scratch_pcm;     // [resident pairs][frame_stride]
    // dynamic range control by table: the range every dynrng word of the batch stands for (a52_dynrng callbacks)
    const float*    drc_ranges;      // [nframes][6][2] or NULL
    // optional dumps
    uint8_t*        dbg_exp;
    uint8_t*        dbg_bap;
    float*          dbg_coef;
    int32_t*        dbg_info;
    // a52_downmix_init for every BSI combination under this call's request: [acmod_ext 0..8][cmixlev * 4 +
    // surmixlev].  Travels with the launch (constant bank), so contexts and calls share no mutable state.
    ModeEntry       mode[9 * 16];
};

}  // namespace a52
