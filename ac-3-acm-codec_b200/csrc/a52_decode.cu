// a52_decode.cu - batched AC-3 (ATSC A/52) decode for NVIDIA B200 (sm_100a).
//
// One persistent kernel decodes thousands of independent AC-3 streams.  A PAIR
// of warps (64 threads) owns a slice of one stream at a time and walks its
// sync frames in order; everything between the staged bitstream and the PCM
// store lives in that pair's part of shared memory and in registers, so the
// only CTA-wide synchronisation is the one after the table load (pairs meet
// at their own named barrier):
//
//   stage   frame bytes  : TMA bulk copy (cp.async.bulk + mbarrier); the next frame is
//                          requested as soon as the last block's mantissas are out
//   parse   BSI + audio-block side info        (reference: liba52/parse.c:131-205, 558-804)  lane 0
//   exps    exponent groups -> exponents       (parse.c:218-270)   lanes = groups, warp scan
//   alloc   parametric bit allocation          (bit_allocate.c:124-265) lanes = bands / bins
//   locate  lanes own contiguous runs of the block's mantissas in coded order: count pass,
//           warp prefix sums of group counters and bit widths, then an emit pass that writes
//           a descriptor per mantissa and per-class work lists (parse.c:336-433); a block
//           that reallocates nothing reuses the count pass and the scans of the last one
//   unpack  per class, convergent: dither (32 LFSR states kept in the warp, advanced 32 steps
//           at a time), 3/5/11-level groups one lane per group code, plain fields
//   couple  coupling fan-out, rematrix         (parse.c:435-556, 837-865)
//   mix     downmix (coefficient or time domain) (downmix.c:162-619)
//   imdct   512 / 2x256 transform: pre-twiddle, radix-4 FFT in shared memory,
//           post-twiddle                        (imdct.c:258-345)
//   ola     KBD window + overlap-add + PCM store (imdct.c:276-292); the overlap tails
//           stay in shared memory from block to block and frame to frame
// imdct + ola of block b run in the pass of block b + 1, beside its side-info parse.
//
// The dither generator position and the overlap tails enter and leave the call
// (and pass from one slice of a stream to the next) through a52_stream_carry_t.
//
// Integer stages are bit-exact with liba52; the float transform uses a
// different FFT factorisation (tolerance 1e-5 relative RMS, measured ~1e-7).
#include <cuda_runtime.h>
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "a52_common.cuh"

namespace a52 {

// ---------------------------------------------------------------------------
// constant memory (warp-uniform lookups only)
// ---------------------------------------------------------------------------
__constant__ MixEntry  c_mix[8 * 11];      // [acmod][output mode]
__constant__ uint8_t   c_nfchans[8] = {2, 1, 2, 3, 3, 4, 4, 5};
__constant__ uint16_t  c_bitrate[19] = {32, 40, 48, 56, 64, 80, 96, 112, 128, 160, 192, 224, 256, 320,
                                        384, 448, 512, 576, 640};
__constant__ int16_t   c_sgain[4]  = {0x540, 0x4d8, 0x478, 0x410};
__constant__ int16_t   c_dbknee[4] = {0x000, 0x700, 0x900, 0xb00};
__constant__ int16_t   c_floor[8]  = {0x2f0, 0x2b0, 0x270, 0x230, 0x1f0, 0x170, 0x0f0, -0x800};

__device__ Tables g_tables;                // filled by the host once per context

// Bounds-checked build (-DA52_BOUNDS_CHECK, tools/gpu_checked.sh): every computed shared-memory store address,
// plan index, bit-window word and staged-frame size is checked against its region; the first violation's code is
// kept in g_violation and read back through a52_batch_violations().  (compute-sanitizer is not available on the
// B200 pool this repository is measured on; the checked build runs the whole GPU suite instead.)
__device__ int g_violation[4];
#ifdef A52_BOUNDS_CHECK
#define A52_CHECK(cond, code) do { if (!(cond)) atomicCAS(&g_violation[0], 0, (code)); } while (0)
#else
#define A52_CHECK(cond, code) do { } while (0)
#endif

// ---------------------------------------------------------------------------
// per-group shared state
// ---------------------------------------------------------------------------
struct Segment {          // a run of mantissas in coded order
    uint8_t  arr;         // exponent/bap array: 0..4 fbw, 5 lfe, 6 coupling
    uint8_t  plane;       // coefficient plane the descriptors go to
    uint8_t  start;       // first bin
    uint8_t  dith;        // dither flag (fbw) / unused (cpl, lfe)
    uint16_t count;       // number of bins
    uint16_t first;       // flat index of the first bin
};

struct __align__(16) GroupCtl {
    // ---- stream / frame ----
    int      stream;           // current stream index (-1 = done)
    int      frame_ok;         // frame header valid
    int      err;              // error raised inside the current block
    int      tr_ticket;        // transform jobs of the pending block handed out so far
    float    wg2[12];          // wt[o][ch] * gain[ch] of outputs 0 and 1 ([0..4] and [5..9]): the stereo mix fast path
    uint32_t arange[8];        // per array 0..6: start | end << 8 | first band << 16 | end band << 24 of the coded range
    uint32_t cur_blk;          // 6 * frame + block being parsed (index into the per-block tables of the launch)
    int16_t  dynw[2];          // the block's dynrng words as coded (-1: none)
    uint32_t pad1;
    // what the last full locate pass left for blocks that repeat its baps, ranges and flags
    uint32_t loc_ta, loc_tb, loc_tz, loc_mant, loc_bitpos;
    uint8_t  loc_valid, loc_dithflag, loc_chincpl, repeat;
    uint32_t base_bit;         // bit offset of the frame inside the staged buffer
    uint32_t limit_bit;        // end of frame (bits) inside the staged buffer
    uint32_t bitpos;           // cursor (bits), advanced block by block
    uint32_t dither_index;     // dither_gen() calls so far in this stream, mod 65535
    uint32_t per_channel;      // delay planes hold per-coded-channel tails (liba52: downmixed == 0)
    // BSI
    uint8_t  fscod, halfrate, acmod, lfeon, nfchans, nout, out_lfe, pad0;
    int      output;           // granted mode incl. LFE bit
    float    clev, slev, level, dynrng;
    // ---- block side info (parse.c:570-804) ----
    uint8_t  blksw, dithflag, chincpl, phsflginu;
    uint8_t  cplstrtmant, cplendmant, cplstrtbnd, ncplbnd;
    uint32_t cplbndstrc;
    uint8_t  rematflg, csnroffst, cplfleak, cplsleak;
    uint8_t  endmant[5];
    alignas(4) uint8_t expstr[7];   // this block's strategies (0 = reuse)
    uint16_t bai;
    uint8_t  chbai[7];
    uint8_t  deltbae[7];
    uint8_t  do_alloc;         // bit per array 0..6 (liba52's do_bit_alloc: 0..4, 32->5, 64->6)
    uint8_t  zero_alloc;       // zero_snr_offsets() shortcut taken
    uint8_t  uniform_path;     // 1: mix coefficients then nout transforms; 0: per-channel transforms
    uint8_t  nseg;
    uint32_t exp_pos[7];       // bit position of the first 7-bit exponent group
    uint8_t  exp_abs[7];       // starting exponent
    uint8_t  exp_ngrp[7];
    float    gain[6];          // per coded channel (5 = lfe), level*dynrng*mix folded
    float    cplco[5][18];
    int8_t   deltba[7][50];
    Segment  seg[8];
    uint32_t total_bins;
    // ---- lane plan of the locate passes ----
    uint16_t plan_count[8];
    uint8_t  plan_lane0[9];    // first lane of each segment (+ end)
    uint8_t  plan_nseg;
    uint16_t plan_K;           // run length per lane
    float    wt[5][5];         // mix matrix [output][coded channel] in {-1, 0, +1} (downmix.c:480-619)
    int      identity_mix;     // output channel o is exactly coded channel o
    int      nobias_mask;      // main outputs that get NO bias when the channels are transformed one by one:
                               // with slev == 0 a52_downmix leaves 2/1, 2/2 -> stereo and 3/1, 3/2 -> 3F
                               // without calling a mixer, and only the mixers add it (downmix.c:526-530,
                               // 546-552, 563-573; parse.c:893-918)
    uint8_t  gains_dirty;      // dynrng / frame parameters changed since compute_gains()
    uint8_t  segs_dirty;       // coded ranges changed since the segment table was built
};

struct WarpPtrs {
    GroupCtl* ctl;
    uint8_t*  exp;     // [7][256]
    uint8_t*  bap;     // [7][256]  standard numbering 0..15
    float*    plane;   // [nplanes][256]: 5 fbw planes (+ the LFE plane when it is an output)
    uint32_t* fbuf;    // staged frame (native-endian 32-bit words after the swap pass)
    uint64_t* mbar;
};


__host__ __device__ inline int align16(int x) { return (x + 15) & ~15; }

// The planes come first in a pair's shared memory, which starts 128-byte aligned: the paired transform of the
// stereo fast path works in place on planes 0 and 1 and XORs address bits 4..6.
__host__ __device__ inline int align128(int x) { return (x + 127) & ~127; }

// Layout of a pair's shared memory.  Every offset but the size of the staged frame (last) is a compile-time
// constant of the kernel instantiation (NPL = coefficient planes: 5, or 6 when the LFE channel is requested), so
// all of a pair's addresses are one register plus an immediate.
template <int NPL>
struct PairLayout {
    static constexpr int plane = 0;                                    // [NPL][256] float, 128-byte aligned
    static constexpr int dump = plane + NPL * 1024;                    // where the missing members of a short group go
    static constexpr int delay = dump + 16;                            // [NPL][128] float overlap-add tails
    static constexpr int ctl = delay + NPL * 512;
    static constexpr int exp = ctl + (((int)sizeof(GroupCtl) + 15) & ~15);
    static constexpr int bap = exp + 7 * 256;
    static constexpr int xch = bap + 7 * 256;                          // [2][8] scan totals of the two warps
    static constexpr int mbar = xch + 64;
    static constexpr int fbuf = mbar + 16;                             // staged frame, fbuf_bytes
};

// The pair's plan in global memory (see the locate stage): 16-byte group entries, 8-byte plain entries,
// 4-byte zero entries.  A block has at most 1504 coded mantissas and 1287 coefficient slots.
constexpr int kPlanGroups = 768, kPlanPlain = 1536, kPlanZeros = 2560;
constexpr int kPlanPlainOff = kPlanGroups * 16, kPlanZeroOff = kPlanPlainOff + kPlanPlain * 8;
constexpr int kPlanBytes = kPlanZeroOff + kPlanZeros * 4;

constexpr int kTablesBytes = ((int)sizeof(Tables) + 16 + 127) & ~127;       // the tables + the CTA's frame gate word

__host__ __device__ inline int pair_smem_bytes(int fbuf_bytes, int nplanes)
{
    return align128((nplanes == 6 ? PairLayout<6>::fbuf : PairLayout<5>::fbuf) + fbuf_bytes);
}

// ---------------------------------------------------------------------------
// small PTX helpers
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                 :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}

// TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---------------------------------------------------------------------------
// bit reader over the staged frame (native-endian words, MSB first)
// reference: bitstream.h:53-77 / bitstream.c:63-97
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t peek_bits(const uint32_t* w, uint32_t pos, uint32_t n)
{
    // n in 1..32
    uint32_t i = pos >> 5, s = pos & 31;
    A52_CHECK(i + 1 < (3840u + 64u) / 4u && n >= 1 && n <= 32, 401);
    uint32_t hi = w[i], lo = w[i + 1];
    uint32_t v = __funnelshift_l(lo, hi, s);
    return v >> (32 - n);
}

struct BitReader {
    const uint32_t* w;
    uint32_t pos;
    uint32_t limit;
    __device__ __forceinline__ uint32_t get(uint32_t n)
    {
        uint32_t v = 0;
        if (pos + n <= limit) v = peek_bits(w, pos, n);
        pos += n;
        return v;
    }
    __device__ __forceinline__ int32_t get_signed(uint32_t n)
    {
        uint32_t v = get(n);
        return ((int32_t)(v << (32 - n))) >> (32 - n);
    }
    __device__ __forceinline__ void skip(uint32_t n) { pos += n; }
};

// The same reader for the audio blocks' side information, where a hundred fields are read one after the other on one
// thread: a 64-bit window in registers, refilled where the parser says so (need()), so that a field costs a shift and
// two funnel shifts instead of two dependent shared-memory loads.  The staged frame is zero from its last bit to the
// end of the buffer, so fields past the end read as zero without a check (a field that straddles the end keeps its
// leading bits; BitReader returns the whole field as zero: damaged frames only).
struct BitWin {
    const uint32_t* w;
    uint32_t pos;            // bit position of the window's first bit
    uint32_t maxw;           // last word index a refill may start at
    uint32_t hi, lo;
    int      left;           // valid bits in the window
    __device__ __forceinline__ BitWin(const uint32_t* w_, uint32_t pos_, uint32_t fbuf_bytes)
        : w(w_), pos(pos_), maxw(fbuf_bytes / 4 - 3), hi(0), lo(0), left(0) {}
    __device__ __forceinline__ void need(int n)
    {
        if (left < n) {
            const uint32_t i = min(pos >> 5, maxw), s = pos & 31;
            const uint32_t w0 = w[i], w1 = w[i + 1], w2 = w[i + 2];
            hi = __funnelshift_l(w1, w0, s);
            lo = __funnelshift_l(w2, w1, s);
            left = 64;
        }
    }
    __device__ __forceinline__ uint32_t get(uint32_t n)       // n in 1..31, after a need() that covers it
    {
        A52_CHECK(left >= (int)n && n >= 1 && n < 32, 402);
        const uint32_t v = hi >> (32 - n);
        hi = __funnelshift_l(lo, hi, n);
        lo <<= n;
        pos += n;
        left -= (int)n;
        return v;
    }
    __device__ __forceinline__ int32_t get_signed(uint32_t n)
    {
        const uint32_t v = get(n);
        return ((int32_t)(v << (32 - n))) >> (32 - n);
    }
    __device__ __forceinline__ void skip(uint32_t n) { pos += n; left = 0; }
};

// bytes to stage for frame f: from the 16-byte aligned address at or below its
// offset up to the start of the next frame (or the end of the buffer), capped
// end of frame f's bytes: the start of the next frame of the table, or the end of the buffer (for the last frame of
// the table, and for tables that are not ascending: frames of different streams may lie in any order)
__device__ __forceinline__ uint64_t frame_end(const DecodeParams& P, uint32_t f, uint64_t off)
{
    const uint64_t nxt = (f + 1 < (uint32_t)P.nframes) ? P.frame_off[f + 1] : P.es_bytes;
    return (nxt > off) ? nxt : P.es_bytes;
}

__device__ __forceinline__ uint32_t stage_bytes(const DecodeParams& P, uint32_t f)
{
    uint64_t off = P.frame_off[f];
    uint64_t a0 = off & ~(uint64_t)15;
    uint64_t end = frame_end(P, f, off);
    uint64_t n = ((end - a0) + 15) & ~(uint64_t)15;
    uint64_t cap = (uint64_t)P.fbuf_bytes - 16;
    return (uint32_t)(n < cap ? n : cap);
}

__device__ __forceinline__ float pow2neg(int k)   // 2^-k, k in [0, 126]
{
    return __int_as_float((127 - k) << 23);
}

// ---------------------------------------------------------------------------
// frame header: a52_syncinfo + a52_frame (parse.c:86-205)
// run by one thread of the group
// ---------------------------------------------------------------------------
__device__ int parse_frame_header(GroupCtl* c, const uint32_t* w, uint32_t base_bit,
                                  const DecodeParams& P, uint32_t avail_bytes, uint32_t cap_bytes)
{
    BitReader br{w, base_bit, base_bit + avail_bytes * 8};
    uint32_t sync = br.get(16);
    br.skip(16);                               // crc1 (never checked, parse.c has no CRC code)
    uint32_t fscod = br.get(2);
    uint32_t frmsizecod = br.get(6);
    uint32_t bsid = br.get(5);
    br.skip(3);                                // bsmod
    if (sync != 0x0b77 || bsid >= 12 || frmsizecod >= 38 || fscod == 3)
        return 1;                              // A52_ST_BAD_SYNC
    int kbps = c_bitrate[frmsizecod >> 1];
    int bytes = (fscod == 0) ? 4 * kbps : (fscod == 2) ? 6 * kbps
              : 2 * (320 * kbps / 147 + (int)(frmsizecod & 1));
    // a frame longer than the staging buffer (a52_batch_set_max_frame_bytes told us less than the stream holds)
    // cannot be decoded: refuse it instead of decoding a truncated copy
    if ((uint32_t)bytes > cap_bytes) return 2;                 // A52_ST_BAD_FRAME
    if ((uint32_t)bytes > avail_bytes) bytes = avail_bytes;
    c->limit_bit = base_bit + bytes * 8;
    br.limit = c->limit_bit;

    c->fscod = fscod;
    c->halfrate = (bsid > 8) ? bsid - 8 : 0;
    int acmod = br.get(3);
    c->acmod = acmod;
    c->nfchans = c_nfchans[acmod];
    int acmod_ext = acmod;
    if (acmod == 2 && br.get(2) == 2) acmod_ext = 8;          // Dolby surround flagged stereo
    int cmix = 0, smix = 0;
    float clev = 0.f, slev = 0.f;
    if ((acmod & 1) && acmod != 1) {
        cmix = br.get(2);
        clev = (cmix == 0) ? 0.7071067811865476f : (cmix == 2) ? 0.5f : 0.5946035575013605f;
    }
    if (acmod & 4) {
        smix = br.get(2);
        slev = (smix == 0) ? 0.7071067811865476f : (smix == 2) ? 0.f : 0.5f;
    }
    c->clev = clev;
    c->slev = slev;
    c->lfeon = br.get(1);

    ModeEntry me = P.mode[acmod_ext * 16 + cmix * 4 + smix];
    if (me.output < 0) return 2;                               // A52_ST_BAD_FRAME
    c->output = me.output;
    c->out_lfe = (c->lfeon && (P.req_flags & M_LFE)) ? 1 : 0;
    if (c->out_lfe) c->output |= M_LFE;
    c->nout = c_mix[acmod * 11 + me.output].nout;
    c->level = me.level * 2.0f;                                // parse.c:169
    c->dynrng = c->level;
    c->gains_dirty = 1;
    c->segs_dirty = 1;
    c->loc_valid = 0;
    // mix matrix as float weights (so that the mixers are plain multiply-adds); fixed for the frame
    {
        const MixEntry mx = c_mix[acmod * 11 + me.output];
        for (int o = 0; o < 5; o++)
            for (int ch = 0; ch < 5; ch++)
                c->wt[o][ch] = ((mx.pos[o] >> ch) & 1) ? 1.f : ((mx.neg[o] >> ch) & 1) ? -1.f : 0.f;
        int ident = (mx.nout == c->nfchans);
        for (int o = 0; o < mx.nout && ident; o++)
            if (mx.pos[o] != (1u << o) || mx.neg[o]) ident = 0;
        c->identity_mix = ident;
        const int om = me.output & M_MASK;
        c->nobias_mask = (slev != 0.f) ? 0
                       : ((acmod == 4 || acmod == 6) && om == M_STEREO) ? 3
                       : ((acmod == 5 || acmod == 7) && om == M_3F) ? 5 : 0;
    }
    // delta bit allocation is reset per frame for cpl + fbw (parse.c:173-175)
    c->deltbae[6] = 2;
    for (int i = 0; i < 5; i++) c->deltbae[i] = 2;

    int reps = (acmod == 0) ? 2 : 1;
    while (reps--) {
        br.skip(5);
        if (br.get(1)) br.skip(8);
        if (br.get(1)) br.skip(8);
        if (br.get(1)) br.skip(7);
    }
    br.skip(2);
    if (br.get(1)) br.skip(14);
    if (br.get(1)) br.skip(14);
    if (br.get(1)) {
        uint32_t n = br.get(6);
        br.skip(8 * (n + 1));
    }
    c->bitpos = br.pos;
    return 0;
}

// ---------------------------------------------------------------------------
// per-channel gains (downmix.c:162-330), by channel role
// ---------------------------------------------------------------------------
__device__ void compute_gains(GroupCtl* c)
{
    enum { R_L, R_C, R_R, R_S, R_SL, R_SR, R_A, R_B };
    static const uint8_t roles[8][5] = {
        {R_A, R_B, 0, 0, 0}, {R_C, 0, 0, 0, 0}, {R_L, R_R, 0, 0, 0}, {R_L, R_C, R_R, 0, 0},
        {R_L, R_R, R_S, 0, 0}, {R_L, R_C, R_R, R_S, 0}, {R_L, R_R, R_SL, R_SR, 0},
        {R_L, R_C, R_R, R_SL, R_SR}};
    const float level = c->dynrng, clev = c->clev, slev = c->slev;
    const float l3 = (float)((double)level * 0.7071067811865476);
    const int out = c->output & M_MASK;
    const int acmod = c->acmod;
    const bool out_c = (out == M_3F || out == M_3F1R || out == M_3F2R);
    const bool out_s1 = (out == M_2F1R || out == M_3F1R);
    const bool out_s2 = (out == M_2F2R || out == M_3F2R);
    for (int ch = 0; ch < c->nfchans; ch++) {
        float g = level;
        switch (roles[acmod][ch]) {
        case R_A:
            g = (out == M_MONO) ? (float)((double)level * 0.5) : (out == M_CHANNEL2) ? 0.f : level;
            break;
        case R_B:
            g = (out == M_MONO) ? (float)((double)level * 0.5) : (out == M_CHANNEL1) ? 0.f : level;
            break;
        case R_L: case R_R:
            g = (out == M_MONO) ? l3 : level;
            break;
        case R_C:
            if (acmod == 1) g = (out == M_DOLBY) ? l3 : level;
            else if (out_c) g = level;
            else if (out == M_MONO) g = (float)((double)(l3 * clev) * 2.0);
            else if (out == M_DOLBY) g = l3;
            else g = level * clev;
            break;
        case R_S:
            if (out_s1) g = level;
            else if (out_s2 || out == M_DOLBY) g = l3;
            else g = l3 * slev;
            break;
        case R_SL: case R_SR:
            if (out_s2) g = level;
            else if (out_s1 || out == M_DOLBY) g = l3;
            else if (out == M_MONO) g = l3 * slev;
            else g = level * slev;
            break;
        }
        c->gain[ch] = g;
    }
    c->gain[5] = level;     // lfe: state->dynrng (parse.c:869-870)
    for (int o = 0; o < 2; o++)
        for (int ch = 0; ch < 5; ch++) c->wg2[o * 5 + ch] = c->wt[o][ch] * ((ch < c->nfchans) ? c->gain[ch] : 0.f);
}

// ---------------------------------------------------------------------------
// audio block side information (parse.c:570-804), one thread
// ---------------------------------------------------------------------------
__device__ int parse_block(GroupCtl* c, const uint32_t* w, const DecodeParams& P, const Tables& T)
{
    BitWin br(w, c->bitpos, (uint32_t)P.fbuf_bytes);
    const int nfchans = c->nfchans;
    const int acmod = c->acmod;
    const uint32_t chmask = (1u << nfchans) - 1;
    uint32_t blksw, dith, do_alloc = 0;

    // A block that keeps everything - coupling strategy and coordinates, rematrixing, every exponent set, the
    // bit allocation parameters, no dynrng word, no skip field - is, after blksw / dithflag, a run of zero
    // flag bits whose length follows from what is already known (parse.c:578-804 with every `if` not taken).
    // One 64-bit window and one compare recognise it; anything else takes the field-by-field parse below.
    bool fast = false;
    if (c->loc_valid && !c->segs_dirty && c->bitpos + 96 <= c->limit_bit) {
        const uint32_t p0 = c->bitpos, wi = p0 >> 5, sft = p0 & 31;
        const uint32_t w0 = w[wi], w1 = w[wi + 1], w2 = w[wi + 2];
        const uint32_t hi = __funnelshift_l(w1, w0, sft), lo = __funnelshift_l(w2, w1, sft);
        const uint32_t nb = 2 * nfchans;
        const uint32_t cpl = c->chincpl;
        // dynrnge (x2 for 1+1) | cplstre | cplcoe per coupled channel | rematstr (2/0) | cplexpstr | chexpstr |
        // lfeexpstr | baie | snroffste | cplleake | deltbaie | skiple
        const uint32_t nz = (acmod ? 1u : 2u) + 1u + (cpl ? (uint32_t)__popc(cpl) + 3u : 0u) + (acmod == 2 ? 1u : 0u)
                          + nb + c->lfeon + 4u;
        const uint32_t rest = __funnelshift_l(lo, hi, nb);          // the bits after blksw / dithflag
        if ((rest >> (32 - nz)) == 0) {
            fast = true;
            const uint32_t v = hi >> (32 - nb);
            blksw = __brev(v >> nfchans) >> (32 - nfchans);
            dith = __brev(v & chmask) >> (32 - nfchans);
            c->blksw = blksw;
            c->dithflag = dith;
            reinterpret_cast<uint32_t*>(c->expstr)[0] = 0;           // expstr[0..3]
            c->expstr[4] = 0; c->expstr[5] = 0; c->expstr[6] = 0;
            c->zero_alloc = 0;
            br.pos = p0 + nb + nz;
        }
    }
    c->dynw[0] = -1;
    c->dynw[1] = -1;
    if (!fast) {
    br.need(64);                               // up to 44 bits: blksw / dithflag, dynrng, the coupling strategy header
    uint32_t v = br.get(2 * nfchans);          // blksw[nfchans], dithflag[nfchans]
    // the fields are sent channel 0 first (msb): bit-reverse them so that bit i = channel i
    blksw = __brev(v >> nfchans) >> (32 - nfchans);
    dith = __brev(v & chmask) >> (32 - nfchans);
    c->blksw = blksw;
    c->dithflag = dith;

    const int nrep = acmod ? 1 : 2;            // dynrng (parse.c:578-598)
    for (int rep = 0; rep < nrep; rep++) {
        if (br.get(1)) {
            int d = br.get_signed(8);
            c->dynw[rep] = (int16_t)(d & 0xff);
            if (!P.drc_off) {
                float range = (float)(((d & 0x1f) | 0x20) << 13) * pow2neg(15 + 3 - (d >> 5));
                // the caller's own mapping of this word (a52_dynrng with a callback, run on the host between
                // a scan pass and this one)
                if (P.drc_ranges) range = P.drc_ranges[(size_t)c->cur_blk * 2 + rep];
                c->dynrng = c->level * range;
                c->gains_dirty = 1;
            }
        }
    }

    if (br.get(1)) {                           // cplstre (parse.c:600-634)
        c->chincpl = 0;
        c->segs_dirty = 1;
        if (br.get(1)) {
            const uint32_t chincpl = __brev(br.get(nfchans)) >> (32 - nfchans);
            c->chincpl = chincpl;
            if (acmod < 2) return 1;
            if (acmod == 2) c->phsflginu = br.get(1);
            int begf = br.get(4), endf = br.get(4);
            int nsub = endf + 3 - begf;
            if (nsub < 0) return 1;
            int nbnd = nsub;
            c->cplstrtmant = begf * 12 + 37;
            c->cplendmant = endf * 12 + 73;
            uint32_t strc = 0;
            br.need(32);
            for (int i = 0; i < nsub - 1; i++)
                if (br.get(1)) { strc |= 1u << i; nbnd--; }
            c->cplbndstrc = strc;
            c->ncplbnd = nbnd;
        }
    }
    const uint32_t chincpl = c->chincpl;

    if (chincpl) {                             // coupling coordinates (parse.c:636-667)
        int any = 0;
        for (int i = 0; i < nfchans; i++)
            if ((chincpl >> i) & 1) {
                br.need(8);
                if (br.get(1)) {
                    any = 1;
                    int mstr = 3 * br.get(2);
                    for (int j = 0; j < c->ncplbnd; j++) {
                        br.need(8);
                        int e = br.get(4), m = br.get(4);
                        m = (e == 15) ? m << 14 : (m | 0x10) << 13;
                        c->cplco[i][j] = (float)m * pow2neg(15 + e + mstr);
                    }
                }
            }
        if (acmod == 2 && c->phsflginu && any) {
            br.need(32);
            for (int j = 0; j < c->ncplbnd; j++)
                if (br.get(1)) c->cplco[1][j] = -c->cplco[1][j];
        }
    }

    br.need(64);                               // rematrix flags, exponent strategies, channel bandwidth codes: <= 48 bits
    if (acmod == 2 && br.get(1)) {             // rematrix flags (parse.c:669-678)
        int stop = chincpl ? c->cplstrtmant : 253;
        uint32_t f = br.get(1);
        if (25 < stop) f |= br.get(1) << 1;
        if (25 < stop && 37 < stop) f |= br.get(1) << 2;
        if (25 < stop && 37 < stop && 61 < stop) f |= br.get(1) << 3;
        c->rematflg = f;
    }

    // exponent strategies (parse.c:680-701)
    uint8_t expstr[7] = {0, 0, 0, 0, 0, 0, 0};
    if (chincpl) expstr[6] = br.get(2);
    {
        const uint32_t ev = br.get(2 * nfchans);
        for (int i = 0; i < nfchans; i++) expstr[i] = (ev >> (2 * (nfchans - 1 - i))) & 3;
        if (ev) c->segs_dirty = 1;              // a channel's coded range may change with new exponents
    }
    if (c->lfeon) expstr[5] = br.get(1);
    for (int i = 0; i < nfchans; i++)
        if (expstr[i]) {
            if ((chincpl >> i) & 1) c->endmant[i] = c->cplstrtmant;
            else {
                int bw = br.get(6);
                if (bw > 60) return 1;
                c->endmant[i] = bw * 3 + 73;
            }
        }

    // exponent fields: remember where they are, decode later in parallel
    if (expstr[6]) {
        // (x / (3 << k) == (x / 3) >> k for non-negative x, and x / 3 == (x * 43691) >> 17 below 2^16: no divider)
        int ngrp = (int)((((uint32_t)(c->cplendmant - c->cplstrtmant) * 43691u) >> 17) >> (expstr[6] - 1));
        do_alloc |= 64;
        br.need(4);
        c->exp_abs[6] = br.get(4) << 1;
        c->exp_pos[6] = br.pos;
        c->exp_ngrp[6] = ngrp;
        br.skip(7 * ngrp);
    }
    for (int i = 0; i < nfchans; i++)
        if (expstr[i]) {
            int gsz = 3 << (expstr[i] - 1);
            int ngrp = (int)((((uint32_t)(c->endmant[i] + gsz - 4) * 43691u) >> 17) >> (expstr[i] - 1));
            do_alloc |= 1u << i;
            br.need(4);
            c->exp_abs[i] = br.get(4);
            c->exp_pos[i] = br.pos;
            c->exp_ngrp[i] = ngrp;
            br.skip(7 * ngrp + 2);             // + gainrng
        }
    if (expstr[5]) {
        do_alloc |= 32;
        br.need(4);
        c->exp_abs[5] = br.get(4);
        c->exp_pos[5] = br.pos;
        c->exp_ngrp[5] = 2;
        br.skip(14);
    }
    for (int i = 0; i < 7; i++) c->expstr[i] = expstr[i];

    // bit allocation side info (parse.c:738-772)
    br.need(64);                               // baie .. the coupling channel's offsets: <= 26 bits
    if (br.get(1)) { do_alloc = 127; c->bai = br.get(11); }
    if (br.get(1)) {
        do_alloc = 127;
        c->csnroffst = br.get(6);
        if (chincpl) c->chbai[6] = br.get(7);
        br.need(64);                           // five channels and the lfe: 42 bits
        for (int i = 0; i < nfchans; i++) c->chbai[i] = br.get(7);
        if (c->lfeon) c->chbai[5] = br.get(7);
    }
    br.need(32);                               // leak terms, deltbaie, deltbae: <= 20 bits
    if (chincpl && br.get(1)) {
        do_alloc |= 64;
        c->cplfleak = br.get(3);               // kept as coded; standard form uses (x<<8)+768
        c->cplsleak = br.get(3);
    }
    if (br.get(1)) {                           // deltbaie
        do_alloc = 127;
        if (chincpl) c->deltbae[6] = br.get(2);
        for (int i = 0; i < nfchans; i++) c->deltbae[i] = br.get(2);
        for (int k = -1; k < nfchans; k++) {
            int a = (k < 0) ? 6 : k;
            if (k < 0 && !chincpl) continue;
            if (c->deltbae[a] != 1) continue;
            int8_t* dst = c->deltba[a];        // parse_deltba (parse.c:272-294)
            for (int j = 0; j < 50; j++) dst[j] = 0;
            br.need(8);
            int nseg = br.get(3) + 1, band = 0;
            while (nseg--) {
                br.need(16);
                band += br.get(5);
                int len = br.get(4), code = br.get(3);
                int delta = (code >= 4) ? code - 3 : code - 4;
                if (!len) continue;
                if (band + len >= 50) return 1;
                while (len--) dst[band++] = (int8_t)delta;
            }
        }
    }

    // zero_snr_offsets (parse.c:296-308)
    c->zero_alloc = 0;
    if (do_alloc) {
        int allzero = !c->csnroffst && !(chincpl && (c->chbai[6] >> 3)) &&
                      !(c->lfeon && (c->chbai[5] >> 3));
        for (int i = 0; i < nfchans && allzero; i++)
            if (c->chbai[i] >> 3) allzero = 0;
        c->zero_alloc = allzero;
        // liba52 only re-allocates arrays whose bit is set and which exist
        uint32_t m = do_alloc & ((1u << nfchans) - 1);
        if (c->lfeon) m |= do_alloc & 32;
        if (chincpl) m |= do_alloc & 64;
        if (allzero) {
            m = ((1u << nfchans) - 1) | 32 | 64;   // memset of every bap array (parse.c:775-780)
        } else if (m & 32) {
            c->deltbae[5] = 2;                     // parse.c:793
        }
        do_alloc = m;
    }

    br.need(16);
    if (br.get(1)) {                           // skip field (parse.c:800-804)
        uint32_t n = br.get(9);
        br.skip(8 * n);
    }
    }   // !fast
    c->do_alloc = do_alloc;
    c->bitpos = br.pos;
    const uint32_t chincpl = c->chincpl;

    if (c->gains_dirty) {
        compute_gains(c);
        c->gains_dirty = 0;
    }
    // may the locate stage reuse what its last full pass left?  (same baps and exponents: nothing
    // reallocated; same coded ranges; same dither and coupling flags)
    c->repeat = c->loc_valid && !c->segs_dirty && do_alloc == 0 && c->dithflag == c->loc_dithflag
                && c->chincpl == c->loc_chincpl;
    c->loc_dithflag = c->dithflag;
    c->loc_chincpl = c->chincpl;
    if (c->segs_dirty) {
    c->segs_dirty = 0;

    // coded order of the mantissas (parse.c:816-835, 867-879)
    int ns = 0, done_cpl = 0;
    uint32_t flat = 0;
    for (int i = 0; i < nfchans; i++) {
        Segment s;
        s.arr = i; s.plane = i; s.start = 0; s.dith = 0;
        s.count = c->endmant[i]; s.first = flat;
        c->seg[ns++] = s;
        flat += s.count;
        if (((chincpl >> i) & 1) && !done_cpl) {
            done_cpl = 1;
            s.arr = 6; s.plane = i; s.start = c->cplstrtmant; s.dith = 0;
            s.count = c->cplendmant - c->cplstrtmant; s.first = flat;
            c->seg[ns++] = s;
            flat += s.count;
        }
    }
    if (c->lfeon) {
        Segment s;
        s.arr = 5; s.plane = 5; s.start = 0; s.dith = 0; s.count = 7; s.first = flat;
        c->seg[ns++] = s;
        flat += 7;
    }
    c->nseg = ns;
    c->total_bins = flat;
    // the coded range of every exponent / bap array and the bands of the masking curve it covers
    for (int a = 0; a < 7; a++) {
        const int st = (a == 6) ? c->cplstrtmant : 0, en = (a == 5) ? 7 : (a == 6) ? c->cplendmant : c->endmant[a];
        c->arange[a] = (uint32_t)st | ((uint32_t)en << 8) | ((uint32_t)T.masktab[st & 255] << 16) |
                       ((uint32_t)(T.masktab[(en - 1) & 255] + 1) << 24);
    }
    // Lanes of the locate passes: every lane walks a run of at most K mantissas of ONE segment,
    // lanes in coded order (so that one warp scan orders the whole block).  K = smallest run
    // length for which the segments need no more than 32 lanes; recomputed only when the
    // segment sizes change.
    {
        bool same = (ns == c->plan_nseg);
        for (int k = 0; k < ns; k++) same = same && (c->seg[k].count == c->plan_count[k]);
        if (!same) {
            // K odd: consecutive lanes then write their descriptors to different shared-memory banks
            uint32_t K = ((flat + P.group_threads - 1) / P.group_threads) | 1;
            // the 7 LFE bins stay in one lane: when the LFE channel is not requested its lane takes part in
            // the bit count but not in the list cursors, which only works for the last lane of the block
            if (c->lfeon && K < 7) K = 7;
            // ceil(count / K) through one reciprocal per candidate K: (x * r) >> 16 with r = 65536 / K + 1 is exact
            // for count < 300, K < 100 (checked exhaustively; a segment has at most 253 bins, K stays below 50)
            uint32_t r;
            for (;; K += 2) {
                r = 65536u / K + 1u;
                uint32_t lanes = 0;
                for (int k = 0; k < ns; k++) lanes += ((c->seg[k].count + K - 1) * r) >> 16;
                if (lanes <= (uint32_t)P.group_threads) break;
            }
            uint32_t l0 = 0;
            for (int k = 0; k < ns; k++) {
                c->plan_lane0[k] = (uint8_t)l0;
                c->plan_count[k] = c->seg[k].count;
                l0 += ((c->seg[k].count + K - 1) * r) >> 16;
            }
            c->plan_lane0[ns] = (uint8_t)l0;
            c->plan_nseg = ns;
            c->plan_K = K;
        }
    }
    }

    // transform path (parse.c:881-886): mix coefficients first unless block
    // switch flags differ between channels that get mixed
    int uniform = 1;
    if (c->nout < nfchans) {
        uint32_t all = (1u << nfchans) - 1;
        if (blksw != 0 && blksw != all) uniform = 0;
    } else {
        uniform = 0;                           // nothing to mix: per-channel transforms
    }
    c->uniform_path = uniform;

    return 0;
}

// ---------------------------------------------------------------------------
// exponent decode for one array, executed by one warp (parse.c:218-270)
// ---------------------------------------------------------------------------
__device__ int decode_exponents(const Tables& T, const uint32_t* w, uint32_t limit, uint8_t* dst, int strategy,
                                int ngrp, uint32_t pos0, int start, int lane)
{
    const int rep = 1 << (strategy - 1);
    int carry = start;
    int bad = 0;
    for (int g0 = 0; g0 < ngrp; g0 += 32) {
        int g = g0 + lane;
        int d0 = 2, d1 = 2, d2 = 2;            // neutral deltas for idle lanes
        if (g < ngrp) {
            uint32_t p = pos0 + 7 * g;
            uint32_t code = (p + 7 <= limit) ? peek_bits(w, p, 7) : 0;
            const uint32_t L = T.exp_lut[code];
            bad |= (int)(L >> 15);
            d0 = L & 15; d1 = (L >> 4) & 15; d2 = (L >> 8) & 15;
        }
        int s = d0 + d1 + d2 - 6;
        int incl = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        int e0 = carry + incl - s + d0 - 2;
        int e1 = e0 + d1 - 2;
        int e2 = e1 + d2 - 2;
        if (g < ngrp) {
            if ((unsigned)e0 > 24u || (unsigned)e1 > 24u || (unsigned)e2 > 24u) bad = 1;
            uint8_t* q = dst + 3 * rep * g;
            for (int r = 0; r < rep; r++) {
                q[r] = (uint8_t)e0;
                q[rep + r] = (uint8_t)e1;
                q[2 * rep + r] = (uint8_t)e2;
            }
        }
        carry += __shfl_sync(0xffffffffu, incl, 31);
    }
    return __any_sync(0xffffffffu, bad);
}

// ---------------------------------------------------------------------------
// parametric bit allocation for one array, executed by one warp.
// A/52 standard form (psd = 3072 - 128*exp); bit-exact with
// bit_allocate.c:124-265 (tests/test_parity_gpu.py, tests/test_oracle.py).
// ---------------------------------------------------------------------------
__device__ __forceinline__ int lowcomp_step(int a, int b0, int b1, int band)
{
    if (band < 7) {
        if (b0 + 256 == b1) a = 384;
        else if (b0 > b1) a = max(a - 64, 0);
    } else if (band < 20) {
        if (b0 + 256 == b1) a = 320;
        else if (b0 > b1) a = max(a - 64, 0);
    } else {
        a = max(a - 128, 0);
    }
    return a;
}

__device__ __forceinline__ void alloc_range(const GroupCtl* c, int a, int& start, int& end)
{
    const uint32_t r = c->arange[a];          // kept by parse_block whenever a coded range changes
    start = r & 255;
    end = (r >> 8) & 255;
}

// Bit allocation of every array flagged in `todo` (bit a: 0..4 fbw, 5 lfe, 6 coupling), one warp.
//   1. band PSD        lanes = (array, band); bands of equal width together so that the serial
//                      log-add inside a band runs in lock step
//   2. excitation      lanes = arrays; the leak recurrences are serial over bands
//   3. masking curve   lanes = (array, band)
//   4. bap lookup      lanes = bins
// psd / mask scratch: int16 [7][50] each.
struct WarpSync { __device__ __forceinline__ void operator()() const { __syncwarp(); } };
struct PairSync {                        // two warps of one stream: named barrier, 64 threads
    int id;
    __device__ __forceinline__ void operator()() const { asm volatile("bar.sync %0, 64;" :: "r"(id) : "memory"); }
};

template <int NT, class Sync>
__device__ void bit_allocate_block(const Tables& T, const GroupCtl* c, uint32_t todo, const uint8_t* exp_all,
                                   uint8_t* bap_all, int16_t* psd_all, int16_t* mask_all, int lane, Sync sync)
{
    // compact list of the arrays to do, 3 bits each
    uint32_t act = 0;
    int na = 0;
    for (uint32_t m = todo; m; m &= m - 1) act |= (uint32_t)(__ffs(m) - 1) << (3 * na++);
    const int half = c->halfrate;

    // ---- 1a. single-bin bands (0..27) ----
    for (int t = lane; t < na * 28; t += NT) {
        const int j = t / 28, band = t - j * 28;
        const int a = (act >> (3 * j)) & 7;
        int start, end;
        alloc_range(c, a, start, end);
        if (band >= start && band < end) psd_all[a * 50 + band] = (int16_t)(3072 - (exp_all[a * 256 + band] << 7));
    }
    // ---- 1b. integrated bands: width 3 (28..34), 6 (35..40), 12 (41..44), 24 (45..49) ----
#pragma unroll 1
    for (int cls = 0; cls < 4; cls++) {
        const int fb = (cls == 0) ? 28 : (cls == 1) ? 35 : (cls == 2) ? 41 : 45;
        const int nb = (cls == 0) ? 7 : (cls == 1) ? 6 : (cls == 2) ? 4 : 5;
        for (int t0 = 0; t0 < na * nb; t0 += NT) {
            const int t = t0 + lane;
            int b0 = 0, b1 = 0, a = 0, band = 0;
            if (t < na * nb) {
                const int j = t / nb;
                band = fb + (t - j * nb);
                a = (act >> (3 * j)) & 7;
                int start, end;
                alloc_range(c, a, start, end);
                b0 = max((int)T.bndtab[band], start);
                b1 = min((int)T.bndtab[band + 1], end);
            }
            if (b0 < b1) {
                const uint8_t* e = exp_all + a * 256;
                int v = 3072 - (e[b0] << 7);
                for (int bin = b0 + 1; bin < b1; bin++) {
                    const int p = 3072 - (e[bin] << 7);
                    const int adr = min(abs(v - p) >> 1, 255);
                    v = max(v, p) + T.latab[adr];
                }
                psd_all[a * 50 + band] = (int16_t)v;
            }
        }
    }
    sync();

    // ---- 2. excitation (serial recurrences), result left in mask[] ----
    if (lane < na) {
        const int a = (act >> (3 * lane)) & 7;
        const int16_t* bndpsd = psd_all + a * 50;
        int16_t* mask = mask_all + a * 50;
        int start, end;
        alloc_range(c, a, start, end);
        const bool is_lfe = (a == 5);
        int fastleak = 0, slowleak = 0;
        if (a == 6) {
            fastleak = (c->cplfleak << 8) + 768;
            slowleak = (c->cplsleak << 8) + 768;
        }
        const int bai = c->bai, chbai = c->chbai[a];
        const int sdecay = (0x0f + 2 * (bai >> 9)) >> half;
        const int fdecay = (0x3f + 0x14 * ((bai >> 7) & 3)) >> half;
        const int sgain = c_sgain[(bai >> 5) & 3];
        const int fgain = 0x80 * ((chbai & 7) + 1);
        const int bndstrt = (c->arange[a] >> 16) & 255;
        const int bndend = c->arange[a] >> 24;
        int band, begin, lowcomp = 0;
        if (bndstrt == 0) {
            lowcomp = lowcomp_step(lowcomp, bndpsd[0], bndpsd[1], 0);
            mask[0] = bndpsd[0] - fgain - lowcomp;
            lowcomp = lowcomp_step(lowcomp, bndpsd[1], bndpsd[2], 1);
            mask[1] = bndpsd[1] - fgain - lowcomp;
            begin = 7;
            for (band = 2; band < 7; band++) {
                bool last_lfe = is_lfe && band == 6;
                if (!last_lfe) lowcomp = lowcomp_step(lowcomp, bndpsd[band], bndpsd[band + 1], band);
                fastleak = bndpsd[band] - fgain;
                slowleak = bndpsd[band] - sgain;
                mask[band] = fastleak - lowcomp;
                if (!last_lfe && bndpsd[band] <= bndpsd[band + 1]) { begin = band + 1; break; }
            }
            int stop = min(bndend, 22);
            for (band = begin; band < stop; band++) {
                if (!(is_lfe && band == 6))
                    lowcomp = lowcomp_step(lowcomp, bndpsd[band], bndpsd[band + 1], band);
                fastleak = max(fastleak - fdecay, bndpsd[band] - fgain);
                slowleak = max(slowleak - sdecay, bndpsd[band] - sgain);
                mask[band] = max(fastleak - lowcomp, slowleak);
            }
            begin = 22;
        } else {
            begin = bndstrt;
        }
        for (band = begin; band < bndend; band++) {
            fastleak = max(fastleak - fdecay, bndpsd[band] - fgain);
            slowleak = max(slowleak - sdecay, bndpsd[band] - sgain);
            mask[band] = max(fastleak, slowleak);
        }
    }
    sync();

    // ---- 3. masking curve, delta, snr offset (per band) ----
    {
        const int bai = c->bai;
        const int dbknee = c_dbknee[(bai >> 3) & 3];
        const int floorv = c_floor[bai & 7];
        const int csnr = c->csnroffst;
        const uint16_t* hth = T.hth + c->fscod * 50;
        for (int t = lane; t < na * 50; t += NT) {
            const int j = t / 50, band = t - j * 50;
            const int a = (act >> (3 * j)) & 7;
            int start, end;
            alloc_range(c, a, start, end);
            const int bndstrt = (c->arange[a] >> 16) & 255;
            const int bndend = c->arange[a] >> 24;
            if (band < bndstrt || band >= bndend) continue;
            const int snroffset = (((csnr - 15) << 4) + (c->chbai[a] >> 3)) << 2;
            const int deltbae = c->deltbae[a];
            int v = mask_all[a * 50 + band];
            const int p = psd_all[a * 50 + band];
            if (p < dbknee) v += (dbknee - p) >> 2;
            v = max(v, (int)hth[band >> half]);
            if (deltbae == 0 || deltbae == 1) v += c->deltba[a][band] * 128;
            v -= snroffset + floorv;
            v = max(v, 0) & 0x1fe0;
            mask_all[a * 50 + band] = (int16_t)(v + floorv);
        }
    }
    sync();

    // ---- 4. pointer lookup per bin ----
    for (int j = 0; j < na; j++) {
        const int a = (act >> (3 * j)) & 7;
        int start, end;
        alloc_range(c, a, start, end);
        const uint8_t* e = exp_all + a * 256;
        const int16_t* mask = mask_all + a * 50;
        uint8_t* bap = bap_all + a * 256;
        for (int bin = start + lane; bin < end; bin += NT) {
            const int p = 3072 - (e[bin] << 7);
            int q = (p - mask[T.masktab[bin]]) >> 5;
            q = min(max(q, 0), 63);
            bap[bin] = T.baptab[q];
        }
    }
}

// ---------------------------------------------------------------------------
// warp helpers
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane)
{
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// field of n (0..16) bits at bit position pos, zero past the end of the frame
__device__ __forceinline__ uint32_t field_at(const uint32_t* w, uint32_t pos, uint32_t n, uint32_t limit)
{
    return (pos + n <= limit) ? peek_bits(w, pos, n) : 0u;
}

// ---------------------------------------------------------------------------
// mantissa descriptors.  A 32-bit word per coded mantissa, written into the
// coefficient-plane slot the coefficient itself will occupy:
//   [4:0] exponent  [8:5] bap  [24:10] bit position of the field / of the group code
// Planes receive q * 2^-(15+exp) (exact); the channel gain is applied later
// (mix weights, or a gain pass when every channel is transformed on its own).
// ---------------------------------------------------------------------------
// byte mask of the bins k0 .. k0+3 that still belong to a run of n bins
__device__ __forceinline__ uint32_t run_mask(uint32_t n, uint32_t k0)
{
    const int rem = max(0, min((int)n - (int)k0, 4));
    return __funnelshift_rc(0xffffffffu, 0u, 32 - 8 * rem);
}

__device__ __forceinline__ uint32_t make_desc(uint32_t exp, uint32_t bap, uint32_t pos)
{
    return exp | (bap << 5) | (pos << 10);
}

// byte permute without the 0x7777 selector masking __byte_perm adds (selector nibbles here are 0..7)
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel)
{
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
    return r;
}

// read-only table loads through 32-bit shared addresses (keeps the shared-window base out of the loops)
__device__ __forceinline__ uint4 lds_v4(uint32_t a)
{
    uint4 v;
    asm("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ uint2 lds_v2(uint32_t a)
{
    uint2 v;
    asm("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
    return v;
}

// predicated shared-memory stores through 32-bit shared addresses (no branch, no address rebuild)
__device__ __forceinline__ void sts_u16_if(uint32_t addr, uint32_t v, uint32_t pred)
{
    asm volatile("{\n.reg .pred p;\nsetp.ne.u32 p, %2, 0;\n@p st.shared.u16 [%0], %1;\n}\n"
                 :: "r"(addr), "h"((unsigned short)v), "r"(pred) : "memory");
}
__device__ __forceinline__ void sts_u32_if(uint32_t addr, uint32_t v, uint32_t pred)
{
    asm volatile("{\n.reg .pred p;\nsetp.ne.u32 p, %2, 0;\n@p st.shared.u32 [%0], %1;\n}\n"
                 :: "r"(addr), "r"(v), "r"(pred) : "memory");
}

__device__ __forceinline__ void sts_f32_if(uint32_t addr, float v, uint32_t pred)
{
    asm volatile("{\n.reg .pred p;\nsetp.ne.u32 p, %2, 0;\n@p st.shared.f32 [%0], %1;\n}\n"
                 :: "r"(addr), "f"(v), "r"(pred) : "memory");
}
__device__ __forceinline__ void stg_u32_if(void* addr, uint32_t v, uint32_t pred)
{
    asm volatile("{\n.reg .pred p;\nsetp.ne.u32 p, %2, 0;\n@p st.global.cg.u32 [%0], %1;\n}\n"
                 :: "l"(addr), "r"(v), "r"(pred) : "memory");
}
// read-only shared-memory loads (tables, the staged frame) through 32-bit shared addresses
__device__ __forceinline__ uint32_t lds_u32(uint32_t a)
{
    uint32_t v;
    asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ int lds_s16(uint32_t a)
{
    int v;
    asm("ld.shared.s16 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}

// peek without a limit check: descriptor positions are clamped to the frame end at emit time
// and the words past the frame end are zero
__device__ __forceinline__ uint32_t peek_nz(const uint32_t* w, uint32_t pos, uint32_t n)
{
    uint32_t i = pos >> 5, s = pos & 31;
    return __funnelshift_l(w[i + 1], w[i], s) >> (32 - n);
}

// coefficient-domain downmix: out[o] = sum over coded channels of wg[o][ch] * in[ch]
template <int NM, int NT = 32>
__device__ __forceinline__ void mix_planes(float* plane, const GroupCtl* c, int lane)
{
    float wg[NM][5];
#pragma unroll
    for (int o = 0; o < NM; o++)
#pragma unroll
        for (int ch = 0; ch < 5; ch++) wg[o][ch] = c->wt[o][ch] * c->gain[ch];
#pragma unroll 2
    for (int bin = lane; bin < 256; bin += NT) {
        float in[5], acc[NM];
#pragma unroll
        for (int ch = 0; ch < 5; ch++) in[ch] = plane[ch * 256 + bin];       // planes past nfchans are zero
#pragma unroll
        for (int o = 0; o < NM; o++) {
            acc[o] = wg[o][0] * in[0];
#pragma unroll
            for (int ch = 1; ch < 5; ch++) acc[o] = fmaf(wg[o][ch], in[ch], acc[o]);
        }
#pragma unroll
        for (int o = 0; o < NM; o++) plane[o * 256 + bin] = acc[o];
    }
}

// ---------------------------------------------------------------------------
// IMDCT building blocks: one warp transforms one 256-coefficient plane in
// place and leaves U[0..127] (new first-half values) in plane[0..127] and
// V[0..127] (new overlap tail) in plane[128..255].
// Natural-order formulation of imdct.c:258-345:
//   512: z_m = rot(pre1[m]; X[2m], X[255-2m]), B = DFT128(z),
//        (a, b) from B[i], B[127-i] with post1[i]
//   256: two 64-point transforms on even/odd quartets
// ---------------------------------------------------------------------------
__device__ __forceinline__ float2 cmul(float2 a, float2 b)
{
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 mul_mj(float2 a) { return make_float2(a.y, -a.x); }   // a * (-j)

// radix-4 DIF butterfly (forward DFT kernel e^{-2 pi j/4})
__device__ __forceinline__ void bfly4(float2& x0, float2& x1, float2& x2, float2& x3)
{
    float2 t0 = cadd(x0, x2), t1 = csub(x0, x2), t2 = cadd(x1, x3), t3 = mul_mj(csub(x1, x3));
    x0 = cadd(t0, t2);
    x2 = csub(t0, t2);
    x1 = cadd(t1, t3);
    x3 = csub(t1, t3);
}

// FFT scratch index -> shared-memory index.  The butterflies touch z at strides 32, 8, 2 and 1;
// XOR-ing two higher index bits into bits 1..3 spreads every one of those access patterns over the
// banks (64-bit accesses are served half a warp at a time).
__device__ __forceinline__ int zsw(int i)
{
    return i ^ (((i >> 4) & 3) << 1) ^ (((i >> 5) & 1) << 3);
}

__device__ void imdct512_warp(const Tables& T, float* plane, int lane)
{
    float2* z = reinterpret_cast<float2*>(plane);
    float2 x[4];
    // pre-twiddle: lane owns m = lane + 32 q
#pragma unroll
    for (int q = 0; q < 4; q++) {
        int m = lane + 32 * q;
        float a = plane[2 * m], b = plane[255 - 2 * m];
        float2 t = T.pre1[m];
        x[q] = make_float2(t.x * a + t.y * b, t.x * b - t.y * a);
    }
    __syncwarp();
    // stage 1: radix 4, stride 32, twiddle W128^(lane p)
    bfly4(x[0], x[1], x[2], x[3]);
    z[zsw(lane)] = x[0];
    z[zsw(lane + 32)] = cmul(x[1], T.wfft[lane]);
    z[zsw(lane + 64)] = cmul(x[2], T.wfft[2 * lane]);
    z[zsw(lane + 96)] = cmul(x[3], T.wfft[(3 * lane) & 127]);
    __syncwarp();
    // stage 2: 4 blocks of 32, stride 8, twiddle W32^(j p) = W128^(4 j p)
    {
        int b = lane >> 3, j = lane & 7, base = b * 32 + j;
        const int i0 = zsw(base), i1 = zsw(base + 8), i2 = zsw(base + 16), i3 = zsw(base + 24);
        float2 y0 = z[i0], y1 = z[i1], y2 = z[i2], y3 = z[i3];
        bfly4(y0, y1, y2, y3);
        __syncwarp();
        z[i0] = y0;
        z[i1] = cmul(y1, T.wfft[4 * j]);
        z[i2] = cmul(y2, T.wfft[8 * j]);
        z[i3] = cmul(y3, T.wfft[12 * j]);
    }
    __syncwarp();
    // stage 3: 16 blocks of 8, stride 2, twiddle W8^(j p) = W128^(16 j p)
    {
        int b = lane >> 1, j = lane & 1, base = b * 8 + j;
        const int i0 = zsw(base), i1 = zsw(base + 2), i2 = zsw(base + 4), i3 = zsw(base + 6);
        float2 y0 = z[i0], y1 = z[i1], y2 = z[i2], y3 = z[i3];
        bfly4(y0, y1, y2, y3);
        __syncwarp();
        z[i0] = y0;
        z[i1] = cmul(y1, T.wfft[16 * j]);
        z[i2] = cmul(y2, T.wfft[32 * j]);
        z[i3] = cmul(y3, T.wfft[48 * j]);
    }
    __syncwarp();
    // stage 4: radix 2 on adjacent pairs
#pragma unroll
    for (int r = 0; r < 2; r++) {
        int u = lane + 32 * r;
        const int ia = zsw(2 * u), ib = zsw(2 * u + 1);
        float2 a = z[ia], b = z[ib];
        z[ia] = cadd(a, b);
        z[ib] = csub(a, b);
    }
    __syncwarp();
    // post-twiddle: B[k] sits at pos(k) = 32 (k&3) + 8 ((k>>2)&3) + 2 ((k>>4)&3) + (k>>6)
    float u_[4], v_[4];
#pragma unroll
    for (int r = 0; r < 2; r++) {
        int i = lane + 32 * r;
        int k2 = 127 - i;
        float2 p = T.post1[i];
        float2 B1 = z[zsw(32 * (i & 3) + 8 * ((i >> 2) & 3) + 2 * ((i >> 4) & 3) + (i >> 6))];
        float2 B2 = z[zsw(32 * (k2 & 3) + 8 * ((k2 >> 2) & 3) + 2 * ((k2 >> 4) & 3) + (k2 >> 6))];
        u_[2 * r] = p.x * B1.x + p.y * B1.y;         // a_r -> U[2i]
        v_[2 * r] = p.y * B1.x - p.x * B1.y;         // a_i -> V[2i]
        u_[2 * r + 1] = -(p.y * B2.x + p.x * B2.y);  // -b_r -> U[2i+1] (sign folded so that the
                                                     // overlap-add below is the same for every p)
        v_[2 * r + 1] = p.x * B2.x - p.y * B2.y;     // b_i -> V[2i+1]
    }
    __syncwarp();
#pragma unroll
    for (int r = 0; r < 2; r++) {
        int i = lane + 32 * r;
        reinterpret_cast<float2*>(plane)[i] = make_float2(u_[2 * r], u_[2 * r + 1]);
        reinterpret_cast<float2*>(plane)[64 + i] = make_float2(v_[2 * r], v_[2 * r + 1]);
    }
    __syncwarp();
}

__device__ __noinline__ void imdct256_warp(const Tables& T, float* plane, int lane)
{
    float2* z = reinterpret_cast<float2*>(plane);
    // two interleaved 64-point transforms: f = lane >> 4 selects the transform
    // for the butterflies; the pre-twiddle covers m = lane, lane + 32 for both
    float2 x1[2], x2[2];
#pragma unroll
    for (int r = 0; r < 2; r++) {
        int m = lane + 32 * r;
        float2 t = T.pre2[m];
        float a = plane[4 * m], b = plane[254 - 4 * m];
        float c = plane[4 * m + 1], d = plane[255 - 4 * m];
        x1[r] = make_float2(t.x * a + t.y * b, t.x * b - t.y * a);
        x2[r] = make_float2(t.x * c + t.y * d, t.x * d - t.y * c);
    }
    __syncwarp();
#pragma unroll
    for (int r = 0; r < 2; r++) {
        z[lane + 32 * r] = x1[r];
        z[64 + lane + 32 * r] = x2[r];
    }
    __syncwarp();
    const int f = lane >> 4, l = lane & 15;
    float2* zf = z + 64 * f;
    // stage 1: stride 16, twiddle W64^(l p) = W128^(2 l p)
    {
        float2 y0 = zf[l], y1 = zf[l + 16], y2 = zf[l + 32], y3 = zf[l + 48];
        bfly4(y0, y1, y2, y3);
        __syncwarp();
        zf[l] = y0;
        zf[l + 16] = cmul(y1, T.wfft[2 * l]);
        zf[l + 32] = cmul(y2, T.wfft[4 * l]);
        zf[l + 48] = cmul(y3, T.wfft[6 * l]);
    }
    __syncwarp();
    // stage 2: 4 blocks of 16, stride 4, twiddle W16^(j p) = W128^(8 j p)
    {
        int b = l >> 2, j = l & 3, base = b * 16 + j;
        float2 y0 = zf[base], y1 = zf[base + 4], y2 = zf[base + 8], y3 = zf[base + 12];
        bfly4(y0, y1, y2, y3);
        __syncwarp();
        zf[base] = y0;
        zf[base + 4] = cmul(y1, T.wfft[8 * j]);
        zf[base + 8] = cmul(y2, T.wfft[16 * j]);
        zf[base + 12] = cmul(y3, T.wfft[24 * j]);
    }
    __syncwarp();
    // stage 3: 16 blocks of 4, stride 1, no twiddle
    {
        int base = l * 4;
        float2 y0 = zf[base], y1 = zf[base + 1], y2 = zf[base + 2], y3 = zf[base + 3];
        bfly4(y0, y1, y2, y3);
        __syncwarp();
        zf[base] = y0; zf[base + 1] = y1; zf[base + 2] = y2; zf[base + 3] = y3;
    }
    __syncwarp();
    // post: B[k] at pos(k) = 16 (k&3) + 4 ((k>>2)&3) + (k>>4); i = lane (0..31)
    const int i = lane, k2 = 63 - i;
    const int p1 = 16 * (i & 3) + 4 * ((i >> 2) & 3) + (i >> 4);
    const int p2 = 16 * (k2 & 3) + 4 * ((k2 >> 2) & 3) + (k2 >> 4);
    float2 p = T.post2[i];
    float2 A1 = z[p1], A2 = z[p2], C1 = z[64 + p1], C2 = z[64 + p2];
    float ar = p.x * A1.x + p.y * A1.y, ai = p.y * A1.x - p.x * A1.y;
    float br = p.y * A2.x + p.x * A2.y, bi = p.x * A2.x - p.y * A2.y;
    float cr = p.x * C1.x + p.y * C1.y, ci = p.y * C1.x - p.x * C1.y;
    float dr = p.y * C2.x + p.x * C2.y, di = p.x * C2.x - p.y * C2.y;
    __syncwarp();
    plane[2 * i] = ar;        plane[127 - 2 * i] = ai;
    plane[2 * i + 1] = bi;    plane[126 - 2 * i] = br;
    plane[128 + 2 * i] = ci;  plane[128 + 127 - 2 * i] = cr;
    plane[128 + 2 * i + 1] = dr; plane[128 + 126 - 2 * i] = di;
    __syncwarp();
}

// ---------------------------------------------------------------------------
// Stereo fast path: the two mixed planes travel as ONE plane of (L, R) pairs and every arithmetic
// instruction of the transform and of the overlap-add is a packed fma/add/mul.f32x2 (sm_100a), so one warp
// transforms both channels in the instruction count of one.  Same factorisation as imdct512_warp
// (imdct.c:258-345); the radix-2 pass is folded into the post-twiddle.
// ---------------------------------------------------------------------------
struct C2 { float2 re, im; };            // one complex point of both channels: re = (re L, re R)

__device__ __forceinline__ float2 f2neg(float2 a) { return make_float2(-a.x, -a.y); }
__device__ __forceinline__ float2 f2add(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 f2sub(float2 a, float2 b) { return __fadd2_rn(a, f2neg(b)); }
__device__ __forceinline__ float2 f2mul(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 f2fma(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }

__device__ __forceinline__ float4 lds_f4(uint32_t a)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ float2 lds_f2(uint32_t a)
{
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts_f4(uint32_t a, float4 v)
{
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" :: "r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ C2 lds_c2(uint32_t a)
{
    const float4 v = lds_f4(a);
    return C2{make_float2(v.x, v.y), make_float2(v.z, v.w)};
}
__device__ __forceinline__ void sts_c2(uint32_t a, const C2& v) { sts_f4(a, make_float4(v.re.x, v.re.y, v.im.x, v.im.y)); }

__device__ __forceinline__ C2 c2add(const C2& a, const C2& b) { return C2{f2add(a.re, b.re), f2add(a.im, b.im)}; }
__device__ __forceinline__ C2 c2sub(const C2& a, const C2& b) { return C2{f2sub(a.re, b.re), f2sub(a.im, b.im)}; }
// a * w, w = (wr, wr, wi, wi)
__device__ __forceinline__ C2 c2mul_tw(const C2& a, const float4 w)
{
    const float2 wr = make_float2(w.x, w.y), wi = make_float2(w.z, w.w);
    return C2{f2fma(f2neg(a.im), wi, f2mul(a.re, wr)), f2fma(a.re, wi, f2mul(a.im, wr))};
}
__device__ __forceinline__ void bfly4_c2(C2& x0, C2& x1, C2& x2, C2& x3)
{
    const C2 t0 = c2add(x0, x2), t1 = c2sub(x0, x2), t2 = c2add(x1, x3), d = c2sub(x1, x3);
    // t3 = d * (-j) = (d.im, -d.re)
    x0 = c2add(t0, t2);
    x2 = c2sub(t0, t2);
    x1 = C2{f2add(t1.re, d.im), f2sub(t1.im, d.re)};
    x3 = C2{f2sub(t1.re, d.im), f2add(t1.im, d.re)};
}

// x_sa: shared address (128-byte aligned) of 256 (L, R) coefficient pairs in mix2_pairs' order.  On exit the same 2 KB hold, as
// float4 units, [i] = (U_L[2i], U_R[2i], U_L[2i+1], U_R[2i+1]) and [64 + i] = the same of V (i < 64), U / V
// as imdct512_warp leaves them per plane.
__device__ __forceinline__ void imdct512_pair(uint32_t tab_base, uint32_t x_sa, int lane)
{
    const uint32_t pre_t = tab_base + (uint32_t)offsetof(Tables, pre1d) + 16u * lane;
    const uint32_t wf_t = tab_base + (uint32_t)offsetof(Tables, wfftd);
    const uint4 A = lds_v4(tab_base + (uint32_t)offsetof(Tables, fftaddr) + 16u * lane);
    const uint32_t e1 = x_sa + (A.x & 0xffffu), e2 = x_sa + (A.x >> 16), e3 = x_sa + (A.y & 0xffffu);
    const uint32_t ep1 = x_sa + (A.y >> 16), ep2 = x_sa + (A.z & 0xffffu);
    C2 x[4];
    // pre-twiddle: lane owns m = lane + 32 q; a = X[2m], b = X[255 - 2m]
#pragma unroll
    for (int q = 0; q < 4; q++) {
        // X[k] sits at ((k >> 6) & 1) * 1024 + (k >> 7) * 512 + (k & 63) * 8 (see mix2_pairs)
        const uint32_t qa = (q & 1) * 1024u + (q >> 1) * 512u, qb = ((3 - q) & 1) * 1024u + ((3 - q) >> 1) * 512u;
        const float2 a = lds_f2(x_sa + 16u * lane + qa);
        const float2 b = lds_f2(x_sa + 504u - 16u * lane + qb);
        const float4 t = lds_f4(pre_t + 512u * q);
        const float2 tx = make_float2(t.x, t.y), ty = make_float2(t.z, t.w);
        x[q].re = f2fma(ty, b, f2mul(tx, a));
        x[q].im = f2fma(f2neg(ty), a, f2mul(tx, b));
    }
    __syncwarp();
    // pass 1: radix 4, stride 32, twiddle W128^(lane p)
    bfly4_c2(x[0], x[1], x[2], x[3]);
    sts_c2(e1, x[0]);
    sts_c2((e1 ^ 0x20u) + 512u, c2mul_tw(x[1], lds_f4(wf_t + 16u * lane)));
    sts_c2((e1 ^ 0x40u) + 1024u, c2mul_tw(x[2], lds_f4(wf_t + 32u * lane)));
    sts_c2((e1 ^ 0x60u) + 1536u, c2mul_tw(x[3], lds_f4(wf_t + 16u * ((3u * lane) & 127u))));
    __syncwarp();
    // pass 2: 4 blocks of 32, stride 8, twiddle W128^(4 j p)
    {
        const uint32_t j = lane & 7;
        const uint32_t i0 = e2, i1 = (e2 ^ 0x30u) + 128u, i2 = (e2 ^ 0x40u) + 256u, i3 = (e2 ^ 0x70u) + 384u;
        C2 y0 = lds_c2(i0), y1 = lds_c2(i1), y2 = lds_c2(i2), y3 = lds_c2(i3);
        bfly4_c2(y0, y1, y2, y3);
        __syncwarp();
        sts_c2(i0, y0);
        sts_c2(i1, c2mul_tw(y1, lds_f4(wf_t + 64u * j)));
        sts_c2(i2, c2mul_tw(y2, lds_f4(wf_t + 128u * j)));
        sts_c2(i3, c2mul_tw(y3, lds_f4(wf_t + 192u * j)));
    }
    __syncwarp();
    // pass 3: 16 blocks of 8, stride 2, twiddle W128^(16 j p)
    {
        const uint32_t j = lane & 1;
        const uint32_t i0 = e3, i1 = e3 ^ 0x20u, i2 = e3 ^ 0x40u, i3 = e3 ^ 0x60u;
        C2 y0 = lds_c2(i0), y1 = lds_c2(i1), y2 = lds_c2(i2), y3 = lds_c2(i3);
        bfly4_c2(y0, y1, y2, y3);
        __syncwarp();
        sts_c2(i0, y0);
        sts_c2(i1, c2mul_tw(y1, lds_f4(wf_t + 256u * j)));
        sts_c2(i2, c2mul_tw(y2, lds_f4(wf_t + 512u * j)));
        sts_c2(i3, c2mul_tw(y3, lds_f4(wf_t + 768u * j)));
    }
    __syncwarp();
    // radix-2 pass + post-twiddle: B[i] = z[pos] + z[pos + 1] (i < 64), B[127 - i] = z[pos' - 1] - z[pos']
    float4 uo[2], vo[2];
#pragma unroll
    for (int r = 0; r < 2; r++) {
        const uint32_t a1 = ep1 ^ (0x40u * r), a2 = ep2 ^ (0x40u * (1 - r));
        const C2 B1 = c2add(lds_c2(a1), lds_c2(a1 ^ 0x10u));
        const C2 B2 = c2sub(lds_c2(a2), lds_c2(a2 ^ 0x10u));
        const float4 p = lds_f4(tab_base + (uint32_t)offsetof(Tables, post1d) + 16u * lane + 512u * r);
        const float2 px = make_float2(p.x, p.y), py = make_float2(p.z, p.w);
        const float2 u0 = f2fma(py, B1.im, f2mul(px, B1.re));              // U[2i]   =  p.x B1.re + p.y B1.im
        const float2 v0 = f2fma(f2neg(px), B1.im, f2mul(py, B1.re));       // V[2i]   =  p.y B1.re - p.x B1.im
        const float2 u1 = f2neg(f2fma(px, B2.im, f2mul(py, B2.re)));       // U[2i+1] = -(p.y B2.re + p.x B2.im)
        const float2 v1 = f2fma(f2neg(py), B2.im, f2mul(px, B2.re));       // V[2i+1] =  p.x B2.re - p.y B2.im
        uo[r] = make_float4(u0.x, u0.y, u1.x, u1.y);
        vo[r] = make_float4(v0.x, v0.y, v1.x, v1.y);
    }
    __syncwarp();
#pragma unroll
    for (int r = 0; r < 2; r++) {
        sts_f4(x_sa + 16u * lane + 512u * r, uo[r]);
        sts_f4(x_sa + 1024u + 16u * lane + 512u * r, vo[r]);
    }
    __syncwarp();
}

// Coefficient-domain downmix to two outputs, written as (L, R) pairs over planes 0 and 1: thread t owns bins
// 4t .. 4t+3.  Warp w reads the w-th halves of the five planes and writes only into the w-th halves of planes 0
// and 1 (pair k at ((k >> 6) & 1) * 1024 + (k >> 7) * 512 + (k & 63) * 8), so a __syncwarp between its loads and
// its stores is all the ordering the in-place mix needs.
__device__ __forceinline__ void mix2_pairs(const float* plane, uint32_t x_sa, const GroupCtl* c, int t)
{
    const float4* p4 = reinterpret_cast<const float4*>(plane);
    float4 in[5];
#pragma unroll
    for (int ch = 0; ch < 5; ch++) in[ch] = p4[ch * 64 + t];          // planes past nfchans are zero
    const float4* wq = reinterpret_cast<const float4*>(c->wg2);
    const float4 w0 = wq[0], w1 = wq[1], w2 = wq[2];
    const float wl[5] = {w0.x, w0.y, w0.z, w0.w, w1.x}, wr[5] = {w1.y, w1.z, w1.w, w2.x, w2.y};
    float L[4], R[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const float v0 = j == 0 ? in[0].x : j == 1 ? in[0].y : j == 2 ? in[0].z : in[0].w;
        L[j] = wl[0] * v0;
        R[j] = wr[0] * v0;
#pragma unroll
        for (int ch = 1; ch < 5; ch++) {
            const float v = j == 0 ? in[ch].x : j == 1 ? in[ch].y : j == 2 ? in[ch].z : in[ch].w;
            L[j] = fmaf(wl[ch], v, L[j]);
            R[j] = fmaf(wr[ch], v, R[j]);
        }
    }
    __syncwarp();
    const uint32_t o = x_sa + ((t >> 4) & 1) * 1024u + (t >> 5) * 512u + (t & 15) * 32u;
    sts_f4(o, make_float4(L[0], R[0], L[1], R[1]));
    sts_f4(o + 16u, make_float4(L[2], R[2], L[3], R[3]));
}

// ---------------------------------------------------------------------------
// the kernel: one warp per stream
// ---------------------------------------------------------------------------
__device__ __forceinline__ void issue_frame_load(const DecodeParams& P, const WarpPtrs& G, uint32_t f)
{
    uint64_t a0 = P.frame_off[f] & ~(uint64_t)15;
    uint32_t nb = stage_bytes(P, f);
    fence_proxy_async();
    mbar_expect_tx(G.mbar, nb);
    A52_CHECK(nb <= (uint32_t)P.fbuf_bytes && (nb & 15) == 0 && a0 + nb <= ((P.es_bytes + 64 + 15) & ~(uint64_t)15), 301);
    tma_load_1d(G.fbuf, P.es + a0, nb, G.mbar);
}

static_assert(sizeof(GroupCtl) <= 1296, "GroupCtl grew: check the shared-memory budget per stream");

// ===========================================================================
// The decode kernel: TWO warps (64 threads, a "pair") walk one stream.  The two warps split every
// lane loop, meet at a 64-thread named barrier, keep the overlap tails in shared memory and
// exchange scan totals through a few words of shared memory.
// ===========================================================================
struct PairPtrs : WarpPtrs {
    float*    delay;   // [nplanes][128] overlap-add tails
    uint32_t* xch;     // [2][8] scan totals of the two warps
};

// The pair's pointers are built from ONE pinned 32-bit shared address (an opaque register): the compiler then addresses
// every field as that register plus an immediate.  Built from the generic `smem + offset` it re-derives the shared
// window base (S2UR CgaCtaId, UMOV, ULEA) in front of most accesses - 500 warp instructions per frame.
template <class T>
__device__ __forceinline__ T* shared_ptr_at(uint32_t sa) { return reinterpret_cast<T*>(__cvta_shared_to_generic((size_t)sa)); }

template <int NPL>
__device__ __forceinline__ PairPtrs carve_pair(uint32_t base_sa)
{
    using Lay = PairLayout<NPL>;
    PairPtrs g;
    g.plane = shared_ptr_at<float>(base_sa + Lay::plane);
    g.delay = shared_ptr_at<float>(base_sa + Lay::delay);
    g.ctl = shared_ptr_at<GroupCtl>(base_sa + Lay::ctl);
    g.exp = shared_ptr_at<uint8_t>(base_sa + Lay::exp);
    g.bap = shared_ptr_at<uint8_t>(base_sa + Lay::bap);
    g.xch = shared_ptr_at<uint32_t>(base_sa + Lay::xch);
    g.mbar = shared_ptr_at<uint64_t>(base_sa + Lay::mbar);
    g.fbuf = shared_ptr_at<uint32_t>(base_sa + Lay::fbuf);
    return g;
}

// Window + overlap-add (+ time-domain mix) + store of one block, every output format and both tail
// representations; thread q owns positions p = 2q, 2q+1 (and their mirrors) of every plane.  Kept out of
// line: the stereo float fast path in the kernel covers the common request, this covers the rest.
// float -> int16 as libao does it (convert2s16.c:33-41): the sample carries the bias 384, so its bit pattern
// minus that of 384.0f is round-to-nearest-even of x * 32768; the subtraction wraps like the reference's int32
// does, and an output channel the reference left unbiased goes through the same arithmetic
__device__ __forceinline__ int16_t s16_of(float v, bool biased)
{
    const float x = biased ? v + 384.f : v;
    const int i = (int)((uint32_t)__float_as_int(x) - 0x43c00000u);
    return (int16_t)min(max(i, -32768), 32767);
}

// libao's WAV sample order (convert2s16.c:199-306) per granted mode [+11 with LFE]: nibble ch = position
// of liba52 plane ch inside a sample group, bits 24-27 = group stride, bits 28-31 = position filled with
// convert (0) or 15.  Row 2F1R+LFE is the reference's fall-through (:270-285): groups of five.
__constant__ uint32_t c_wavmap[22] = {
    0xf2000010, 0xf1000000, 0xf2000010, 0xf3000120, 0xf3000210, 0xf4003120, 0xf4003210, 0xf5043120,
    0xf1000000, 0xf1000000, 0xf2000010,
    0xf3000102, 0xf2000001, 0xf3000102, 0xf4001203, 0x45001203, 0xf5041203, 0xf5043102, 0xf6541203,
    0xf2000001, 0xf2000001, 0xf3000102};

__device__ __noinline__ void ola_store_generic(const Tables& T, const DecodeParams& P, const PairPtrs& G, const GroupCtl* c,
                                               uint8_t* out_frame, int blk, int gt, int nfchans, int nmain,
                                               bool uniform, int lfe_on)
{
        const int nout = nmain + lfe_on;
        const int nobias = uniform ? 0 : c->nobias_mask;
        const bool identity = c->identity_mix;
        const float2* plane2 = reinterpret_cast<const float2*>(G.plane);
        const float2* win2 = reinterpret_cast<const float2*>(T.window);
        float2* delay2 = reinterpret_cast<float2*>(G.delay);
        const int q = gt;
        if (uniform && c->per_channel) {
            // a52_downmix on the per-channel tails; zero-gain channels are left out
            float2 d[5], m[5];
#pragma unroll
            for (int ch = 0; ch < 5; ch++)
                d[ch] = (ch < nfchans && c->gain[ch] != 0.f) ? delay2[ch * 64 + q] : make_float2(0.f, 0.f);
#pragma unroll
            for (int o = 0; o < 5; o++) {
                float ax = 0.f, ay = 0.f;
#pragma unroll
                for (int ch = 0; ch < 5; ch++) { ax = fmaf(c->wt[o][ch], d[ch].x, ax); ay = fmaf(c->wt[o][ch], d[ch].y, ay); }
                m[o] = make_float2(ax, ay);
            }
#pragma unroll
            for (int o = 0; o < 5; o++)
                if (o < nmain) delay2[o * 64 + q] = m[o];
        } else if (!uniform && !c->per_channel) {
            // a52_upmix: downmixed tails go back to the coded channels they belong to
            const MixEntry mx = c_mix[c->acmod * 11 + (c->output & M_MASK)];
            float2 m[5];
#pragma unroll
            for (int o = 0; o < 5; o++) m[o] = delay2[o * 64 + q];
#pragma unroll
            for (int ch = 0; ch < 5; ch++) {
                if (ch < nfchans) {
                    float2 v = make_float2(0.f, 0.f);
#pragma unroll
                    for (int o = 0; o < 5; o++)
                        if (mx.up[ch] == o) v = m[o];
                    delay2[ch * 64 + q] = v;
                }
            }
        }
        const float2 wl = win2[q], wh = win2[127 - q];      // (w[p], w[p+1]), (w[254-p], w[255-p])
        float y[6][4];                                       // [output][p, p+1, 254-p, 255-p]
#pragma unroll
        for (int o = 0; o < 6; o++)
#pragma unroll
            for (int r = 0; r < 4; r++) y[o][r] = 0.f;
#pragma unroll
        for (int pl = 0; pl < 6; pl++) {
            const bool is_lfe = (pl == 5);
            bool live;
            if (is_lfe) live = lfe_on;
            else if (uniform) live = pl < nmain;
            else live = pl < nfchans && c->gain[pl] != 0.f;
            if (!live) continue;
            const float2 U = plane2[pl * 128 + q], V = plane2[pl * 128 + 64 + q];
            const float2 D = delay2[pl * 64 + q];
            const float a0 = D.x * wh.y - U.x * wl.x;        // sample p
            const float a1 = D.y * wh.x - U.y * wl.y;        // sample p + 1
            const float b0 = D.x * wl.x + U.x * wh.y;        // sample 255 - p
            const float b1 = D.y * wl.y + U.y * wh.x;        // sample 254 - p
            delay2[pl * 64 + q] = V;
            if (is_lfe) {
                y[0][0] = a0; y[0][1] = a1; y[0][2] = b1; y[0][3] = b0;
            } else if (uniform || identity) {
#pragma unroll
                for (int oo = 0; oo < 6; oo++)
                    if (oo == pl + lfe_on) { y[oo][0] = a0; y[oo][1] = a1; y[oo][2] = b1; y[oo][3] = b0; }
            } else {
#pragma unroll
                for (int o = 0; o < 5; o++) {
                    const float wgt = c->wt[o][pl];
#pragma unroll
                    for (int oo = 0; oo < 6; oo++)
                        if (o < nmain && oo == o + lfe_on) {
                            y[oo][0] = fmaf(wgt, a0, y[oo][0]);
                            y[oo][1] = fmaf(wgt, a1, y[oo][1]);
                            y[oo][2] = fmaf(wgt, b1, y[oo][2]);
                            y[oo][3] = fmaf(wgt, b0, y[oo][3]);
                        }
                }
            }
        }
        const int p = 2 * q;
        const uint32_t wav_map = c_wavmap[(c->output & M_MASK) + (lfe_on ? 11 : 0)];
        const int wav_stride = (wav_map >> 24) & 15, wav_fill = wav_map >> 28;
        if (P.out_fmt == 1 && nout == 2 && !nobias) {
            const float bias = P.bias;
            float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(out_frame) + (size_t)blk * 512);
            dst[q] = make_float4(y[0][0] + bias, y[1][0] + bias, y[0][1] + bias, y[1][1] + bias);
            dst[127 - q] = make_float4(y[0][2] + bias, y[1][2] + bias, y[0][3] + bias, y[1][3] + bias);
        } else if (P.out_fmt == 1 && nout == 6 && !nobias) {
            // 5.1 interleaved float: samples p, p + 1 of the six channels are 48 contiguous, 16-byte aligned bytes
            const float bias = P.bias;
            float4* lo = reinterpret_cast<float4*>(reinterpret_cast<float*>(out_frame) + ((size_t)blk * 256 + p) * 6);
            float4* hi = reinterpret_cast<float4*>(reinterpret_cast<float*>(out_frame) + ((size_t)blk * 256 + 254 - p) * 6);
            lo[0] = make_float4(y[0][0] + bias, y[1][0] + bias, y[2][0] + bias, y[3][0] + bias);
            lo[1] = make_float4(y[4][0] + bias, y[5][0] + bias, y[0][1] + bias, y[1][1] + bias);
            lo[2] = make_float4(y[2][1] + bias, y[3][1] + bias, y[4][1] + bias, y[5][1] + bias);
            hi[0] = make_float4(y[0][2] + bias, y[1][2] + bias, y[2][2] + bias, y[3][2] + bias);
            hi[1] = make_float4(y[4][2] + bias, y[5][2] + bias, y[0][3] + bias, y[1][3] + bias);
            hi[2] = make_float4(y[2][3] + bias, y[3][3] + bias, y[4][3] + bias, y[5][3] + bias);
        } else {
#pragma unroll
            for (int oc = 0; oc < 6; oc++) {
                if (oc >= nout) continue;
                const float v0 = y[oc][0], v1 = y[oc][1], v2 = y[oc][2], v3 = y[oc][3];
                const bool biased = !(oc >= lfe_on && ((nobias >> (oc - lfe_on)) & 1));
                const float bias = biased ? P.bias : 0.f;
                if (P.out_fmt == 0) {
                    float* dst = reinterpret_cast<float*>(out_frame) + ((size_t)blk * nout + oc) * 256;
                    *reinterpret_cast<float2*>(dst + p) = make_float2(v0 + bias, v1 + bias);
                    *reinterpret_cast<float2*>(dst + 254 - p) = make_float2(v2 + bias, v3 + bias);
                } else if (P.out_fmt == 1) {
                    float* dst = reinterpret_cast<float*>(out_frame) + (size_t)blk * 256 * nout;
                    dst[p * nout + oc] = v0 + bias;
                    dst[(p + 1) * nout + oc] = v1 + bias;
                    dst[(254 - p) * nout + oc] = v2 + bias;
                    dst[(255 - p) * nout + oc] = v3 + bias;
                } else if (P.out_fmt == 3) {
                    // WAV channel order (convert2s16.c:199-306); a value lands where the reference puts it
                    // and is dropped when that lies past the 256 * nout values wav_play writes
                    int16_t* dst = reinterpret_cast<int16_t*>(out_frame) + (size_t)blk * 256 * nout;
                    const int pos = (wav_map >> (4 * oc)) & 15, lim = 256 * nout;
                    const float vv[4] = {v0, v1, v2, v3};
                    const int ss[4] = {p, p + 1, 254 - p, 255 - p};
#pragma unroll
                    for (int r = 0; r < 4; r++) {
                        const int idx = ss[r] * wav_stride + pos;
                        if (idx < lim) dst[idx] = s16_of(vv[r], biased);
                        // the slot the fall-through case reads from a plane nobody wrote: convert (+0.0f)
                        const int fidx = ss[r] * wav_stride + wav_fill;
                        if (oc == 0 && wav_fill != 15 && fidx < lim) dst[fidx] = (int16_t)-32768;
                    }
                } else {
                    int16_t* dst = reinterpret_cast<int16_t*>(out_frame) + (size_t)blk * 256 * nout;
                    dst[p * nout + oc] = s16_of(v0, biased);
                    dst[(p + 1) * nout + oc] = s16_of(v1, biased);
                    dst[(254 - p) * nout + oc] = s16_of(v2, biased);
                    dst[(255 - p) * nout + oc] = s16_of(v3, biased);
                }
            }
        }
}

#ifndef A52_GATE_SLEEP_NS
#define A52_GATE_SLEEP_NS 400          // (100 .. 400 ns measure the same; 1.5 us and more lose)
#endif
// The gates of a CTA (see the kernel): one shared-memory word each, [7:0] arrived, [15:8] members, [31:16] generation.
__device__ __forceinline__ void gate_arrive_fn(unsigned int* g, bool wait, int gt, int bar_id)
{
    if (gt == 0) {
        unsigned int old = *(volatile unsigned int*)g, seen;
        bool released;
        do {
            seen = old;
            const unsigned int cnt = (seen & 255u) + 1u, act = (seen >> 8) & 255u;
            released = cnt >= act;
            const unsigned int nw = released ? ((((seen >> 16) + 1u) & 0xffffu) << 16) | (act << 8) : seen + 1u;
            old = atomicCAS(g, seen, nw);
        } while (old != seen);
        if (!released && wait)
            while ((*(volatile unsigned int*)g >> 16) == (seen >> 16)) __nanosleep(A52_GATE_SLEEP_NS);
    }
    if (wait) asm volatile("bar.sync %0, 64;" :: "r"(bar_id) : "memory");
}

__device__ __forceinline__ void gate_leave_fn(unsigned int* g)
{
    // if everyone who stays has already arrived, let them go
    unsigned int old = *(volatile unsigned int*)g, seen;
    do {
        seen = old;
        const unsigned int cnt = seen & 255u, act = ((seen >> 8) & 255u) - 1u;
        const unsigned int nw = (cnt && cnt >= act) ? ((((seen >> 16) + 1u) & 0xffffu) << 16) | (act << 8)
                                                    : (seen & ~0xff00u) | (act << 8);
        old = atomicCAS(g, seen, nw);
    } while (old != seen);
}

#ifndef A52_PAIRS_PER_CTA
#define A52_PAIRS_PER_CTA 14
#endif
constexpr int kMaxPairsPerCta = A52_PAIRS_PER_CTA;      // named barriers 1..15 serve the pairs: at most 15

template <int NPL>
__global__ void __launch_bounds__(kMaxPairsPerCta * 64, 1)
a52_decode_kernel(const DecodeParams P)
{
    extern __shared__ __align__(128) uint8_t smem[];
    Tables& T = *reinterpret_cast<Tables*>(smem);
    // Frame gate: the pairs of a CTA that are inside a frame loop start every frame together.  The kernel is bound by
    // instruction fetch - its pairs are all at different places of a program several times the size of the
    // instruction cache - and the per-frame work (header, block 0's side information, exponents, bit allocation,
    // locate passes) is the largest part of that program that a pair runs only once per frame: started together,
    // one pair's fetches serve the others.  One word: [7:0] arrived, [15:8] members, [31:16] generation.
    // Streams whose blocks are mostly of the expensive kind (new exponents / allocation in most blocks: what the last
    // frame looked like decides) also start every BLOCK together, through a second gate of their own.
    unsigned int* const gates = reinterpret_cast<unsigned int*>(smem + sizeof(Tables));  // (inside the tables' padding)
    unsigned int& gate = gates[0];
    unsigned int& bgate = gates[1];
    const int tid = threadIdx.x;
    const int wid = tid >> 5, lane = tid & 31, pair = wid >> 1, w = wid & 1;
    const int gt = w * 32 + lane;
    constexpr int NT = 64;
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(&g_tables);
        uint32_t* dst = reinterpret_cast<uint32_t*>(&T);
        for (int i = tid; i < (int)(sizeof(Tables) / 4); i += blockDim.x) dst[i] = src[i];
    }
    if (tid == 0) { gates[0] = 0; gates[1] = 0; }
    // (the pair's offset is pinned in a register: left to itself the compiler re-derives it - a constant-bank load
    // and a multiply - in front of most shared-memory accesses)
    uint32_t pair_sa;
    asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(pair_sa)
                 : "r"((uint32_t)pair), "r"((uint32_t)P.warp_bytes), "r"(smem_u32(smem) + (uint32_t)kTablesBytes));
    const PairPtrs G = carve_pair<NPL>(pair_sa);
    GroupCtl* c = G.ctl;
    uint32_t* const W = G.fbuf;
    const PairSync sync{pair + 1};
    if (gt == 0) {
        mbar_init(G.mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = gt; i < (int)(sizeof(GroupCtl) / 4); i += NT) reinterpret_cast<uint32_t*>(c)[i] = 0;
    for (int i = gt; i < 7 * 256 * 2 / 4; i += NT) reinterpret_cast<uint32_t*>(G.exp)[i] = 0;
    __syncthreads();
    uint32_t tab_base, plane_sa;
    asm volatile("mov.u32 %0, %1;" : "=r"(tab_base) : "r"(smem_u32(&T)));
    plane_sa = pair_sa + (uint32_t)PairLayout<NPL>::plane;
    constexpr uint32_t kDumpWord = (uint32_t)(PairLayout<NPL>::dump - PairLayout<NPL>::plane);   // scale bits 0: stores 0.0f
    uint32_t phase = 0;
    constexpr int ndelay = NPL;              // tails: planes 0..4 main, 5 LFE
    // this pair's scratch in global memory: the plan of the last full locate pass
    uint8_t* const plan = P.plan + (size_t)(blockIdx.x * (blockDim.x >> 6) + pair) * kPlanBytes;

    // Work units are SLICES of streams (P.slice_frames frames), handed out slice-major from one ticket
    // counter: all first slices, then all second slices, ...  A slice starts from the carry record its
    // predecessor left in global memory; the predecessor holds a smaller ticket, so it is already running
    // on some resident CTA (the grid never exceeds the SM count) and the wait below cannot deadlock.
    // Slicing keeps the last wave of a batch short: the tail of a launch is one slice, not one stream.
    const uint32_t nunits = (uint32_t)P.nstreams * (uint32_t)P.nslices;
    for (;;) {
        if (gt == 0) c->stream = atomicAdd(P.work_counter, 1);
        sync();
        const uint32_t ticket = (uint32_t)c->stream;
        if (ticket >= nunits) break;
        const uint32_t slice = ticket / (uint32_t)P.nstreams;
        const int s = (int)(ticket - slice * (uint32_t)P.nstreams);
        const uint32_t fs0 = P.stream_first[s], fs1 = P.stream_first[s + 1];
        uint32_t f0 = fs0 + slice * (uint32_t)P.slice_frames;
        if (f0 > fs1) f0 = fs1;
        const uint32_t f1 = (slice + 1 < (uint32_t)P.nslices && fs1 - f0 > (uint32_t)P.slice_frames) ? f0 + P.slice_frames : fs1;
        const bool chained = !P.indep && !P.scan_only;
        if (slice > 0 && chained) {
            if (gt == 0) {
                volatile int* done = P.slice_done + s;
                while (*done < (int)slice) __nanosleep(200);
                __threadfence();
            }
            sync();
        }
        // a slice on its own decodes one frame of look-back first (not in a scan: nothing is carried there)
        const bool lookback = P.indep && slice > 0 && f0 < f1;
        const uint32_t fl = lookback ? f0 - 1 : f0;
        const bool have_state = P.carry && !P.scan_only && (chained ? (slice > 0 || P.carry_init) : (slice == 0 && P.carry_init));
        uint32_t dither_index = 0;
        if (have_state) {
            const StreamCarry* cin = (P.carry_in ? P.carry_in : P.carry) + s;
            dither_index = __ldcg(&cin->dither_index) % kDitherPeriod;
            if (gt == 0) c->per_channel = (__ldcg(&cin->per_channel) != 0);
            for (int i = gt; i < ndelay * 128; i += NT) G.delay[i] = __ldcg(&cin->delay[i >> 7][i & 127]);
        } else {
            // a fresh stream: liba52 starts from a52_init's state; so do the parse state (exponents, baps, coupling
            // and allocation parameters a damaged first frame might "reuse") and the tails
            // (not the first four words: the ticket other threads of the pair may still be reading)
            for (int i = 4 + gt; i < (int)(sizeof(GroupCtl) / 4); i += NT) reinterpret_cast<uint32_t*>(c)[i] = 0;
            for (int i = gt; i < 7 * 256 * 2 / 4; i += NT) reinterpret_cast<uint32_t*>(G.exp)[i] = 0;
            for (int i = gt; i < ndelay * 128; i += NT) G.delay[i] = 0.f;
        }
        if (P.indep && f0 < f1) {
            // generator position at frame fl (carry of the call included by the host's prefix sum)
            dither_index = __ldg(P.slice_dither + (size_t)s * P.nslices + slice) % kDitherPeriod;
        }
        if (gt == 0 && fl < f1) issue_frame_load(P, G, fl);
        const bool gated = P.lockstep && fl < f1;
        if (gated && gt == 0) atomicAdd(&gate, 1u << 8);         // a member from here to the end of the unit's frames
        sync();

        // arrive at a gate (thread 0 of the pair), optionally wait for the round to complete; leave a gate
        auto gate_arrive = [&](unsigned int& g, bool wait) { gate_arrive_fn(&g, wait, gt, pair + 1); };
        auto gate_leave = [&](unsigned int& g) { if (gt == 0) gate_leave_fn(&g); };
        int cold_prev = 0;                        // expensive blocks among blocks 1..5 of the unit's previous frame
        for (uint32_t f = fl; f < f1; f++) {
            if (gated) gate_arrive(gate, true);
            const bool bgated = gated && cold_prev >= 3;      // this frame's blocks go through the block gate
            if (bgated && gt == 0) atomicAdd(&bgate, 1u << 8);
            // (P.lockstep >= 2: such streams also meet in front of every stage of a block - their blocks are long, and
            // the instruction cache holds about one stage's code of it)
            const bool sgated = bgated && P.lockstep >= 2;
            const int gates_due = sgated ? 5 + 6 * 5 : 5;
            int cold_now = 0, gates_passed = 0;
            const bool shadow = lookback && f == fl;          // decoded for its state only
            uint32_t frame_draws = 0;
            const uint64_t off = P.frame_off[f];
            bool next_issued = false;
            mbar_wait(G.mbar, phase);
            phase ^= 1;
            for (int i = gt; i < P.fbuf_bytes / 4; i += NT) W[i] = __byte_perm(W[i], 0, 0x0123);
            sync();
            if (gt == 0) {
                uint32_t base_bit = (uint32_t)(off & 15) * 8;
                const uint32_t cap = P.fbuf_bytes - 16 - (uint32_t)(off & 15);     // what the staging buffer holds
                uint32_t avail = cap;
                {
                    const uint64_t end = frame_end(P, f, off);
                    if (end - off < avail) avail = (uint32_t)(end - off);
                }
                c->base_bit = base_bit;
                int st = parse_frame_header(c, W, base_bit, P, avail, cap);
                c->frame_ok = (st == 0);
                c->err = st;
                if (P.frame_flags && !shadow) P.frame_flags[f] = st ? 0 : c->output;
            }
            sync();
            int frame_status = c->err;
            const bool frame_ok = c->frame_ok;
            uint8_t* out_frame = shadow ? P.scratch_pcm + (size_t)(blockIdx.x * (blockDim.x >> 6) + pair) * P.frame_stride
                                        : P.pcm + (size_t)f * P.frame_stride;
            if (frame_ok) {
                uint32_t lim = c->limit_bit;
                for (uint32_t i = (lim >> 5) + gt; i < (uint32_t)P.fbuf_bytes / 4; i += NT) {
                    if (i == (lim >> 5)) {
                        uint32_t keep = lim & 31;
                        W[i] = keep ? (W[i] & (0xffffffffu << (32 - keep))) : 0;
                    } else W[i] = 0;
                }
                sync();
            }

            // The transforms of a block run while thread 0 reads the side information of the NEXT block: the
            // other warp starts on them at once, the parsing warp joins when it is through (jobs by ticket).
            // What they need of the finished block is kept in registers (p_*); a seventh pass flushes block 5.
            int blk = 0;
            bool pend = false;
            int p_blk = 0;
            uint32_t p_live = 0, p_blksw = 0;
            bool p_uniform = false, p_fast = false;
            if (frame_ok)
            for (;; blk++) {
                const bool more = blk < 6;
                if (bgated && blk > 0 && more) { gate_arrive(bgate, true); gates_passed++; }
                // ================= P (block blk) | T (block blk - 1) =================
                if (more && gt == 0) {
                    c->cur_blk = f * 6u + (uint32_t)blk;
                    c->err = parse_block(c, W, P, T);
                    if (P.scan_only && !c->err) {
                        P.scan[f].dynrng[blk][0] = c->dynw[0];
                        P.scan[f].dynrng[blk][1] = c->dynw[1];
                    }
                }
                __syncwarp();
                if (pend && p_fast) {
                    // both mixed planes in one packed transform; whichever warp gets here first takes it
                    int j = 0;
                    if (lane == 0) j = atomicAdd(&c->tr_ticket, 1);
                    j = __shfl_sync(0xffffffffu, j, 0);
                    if (j == 0) imdct512_pair(tab_base, plane_sa, lane);
                } else if (pend) {
                    const int njobs = __popc(p_live);
                    for (;;) {
                        int j = 0;
                        if (lane == 0) j = atomicAdd(&c->tr_ticket, 1);
                        j = __shfl_sync(0xffffffffu, j, 0);
                        if (j >= njobs) break;
                        uint32_t m = p_live;
                        for (int k = 0; k < j; k++) m &= m - 1;
                        const int pl = __ffs(m) - 1;
                        const bool shortblk = (pl < 5) && ((p_blksw >> (p_uniform ? 0 : pl)) & 1);
                        if (shortblk) imdct256_warp(T, G.plane + pl * 256, lane);
                        else imdct512_warp(T, G.plane + pl * 256, lane);
                    }
                }
                sync();
                // ================= O (block blk - 1) =================
                if (pend) {
                    const int nmain = c->nout, lfe_on = c->out_lfe, nfch = c->nfchans;
                    // thread q = gt owns positions p = 2q, 2q+1 (and their mirrors 254-p, 255-p) of every plane
                    if (p_fast) {
                        // stereo fast path: (L, R) pairs all the way, packed arithmetic
                        const int q = gt;
                        const float4 Uq = lds_f4(plane_sa + 16u * q), Vq = lds_f4(plane_sa + 1024u + 16u * q);
                        float2* delay2 = reinterpret_cast<float2*>(G.delay);
                        const float2 DL = delay2[q], DR = delay2[64 + q];
                        const float4 WL = lds_f4(tab_base + (uint32_t)offsetof(Tables, win2d) + 16u * q);
                        const float4 WH = lds_f4(tab_base + (uint32_t)offsetof(Tables, win2d) + 16u * (127 - q));
                        const float2 U0 = make_float2(Uq.x, Uq.y), U1 = make_float2(Uq.z, Uq.w);
                        const float2 D0 = make_float2(DL.x, DR.x), D1 = make_float2(DL.y, DR.y);
                        const float2 wlx = make_float2(WL.x, WL.y), wly = make_float2(WL.z, WL.w);
                        const float2 whx = make_float2(WH.x, WH.y), why = make_float2(WH.z, WH.w);
                        const float2 bias2 = make_float2(P.bias, P.bias);
                        const float2 y0 = f2fma(f2neg(U0), wlx, f2mul(D0, why));      // sample p
                        const float2 y1 = f2fma(f2neg(U1), wly, f2mul(D1, whx));      // sample p + 1
                        const float2 y3 = f2fma(U0, why, f2mul(D0, wlx));             // sample 255 - p
                        const float2 y2 = f2fma(U1, whx, f2mul(D1, wly));             // sample 254 - p
                        delay2[q] = make_float2(Vq.x, Vq.z);
                        delay2[64 + q] = make_float2(Vq.y, Vq.w);
                        if (P.out_fmt == 1) {
                            const float2 a0 = f2add(y0, bias2), a1 = f2add(y1, bias2), a2 = f2add(y2, bias2), a3 = f2add(y3, bias2);
                            float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(out_frame) + (size_t)p_blk * 512);
                            dst[q] = make_float4(a0.x, a0.y, a1.x, a1.y);
                            dst[127 - q] = make_float4(a2.x, a2.y, a3.x, a3.y);
                        } else {
                            auto two = [](float2 v) {
                                return (uint32_t)(uint16_t)s16_of(v.x, true) | ((uint32_t)(uint16_t)s16_of(v.y, true) << 16);
                            };
                            uint2* dst = reinterpret_cast<uint2*>(reinterpret_cast<int16_t*>(out_frame) + (size_t)p_blk * 512);
                            dst[q] = make_uint2(two(y0), two(y1));
                            dst[127 - q] = make_uint2(two(y2), two(y3));
                        }
                    } else if (p_uniform && !c->per_channel && nmain == 2 && !lfe_on && P.out_fmt != 0) {
                        // the common request: two mixed planes, tails already downmixed, interleaved float or
                        // int16 out (two channels: libao's WAV order is liba52's)
                        const float bias = P.bias;
                        const float2* plane2 = reinterpret_cast<const float2*>(G.plane);
                        const float2* win2 = reinterpret_cast<const float2*>(T.window);
                        float2* delay2 = reinterpret_cast<float2*>(G.delay);
                        const int q = gt;
                        const float2 wl = win2[q], wh = win2[127 - q];
                        float y[2][4];
#pragma unroll
                        for (int pl = 0; pl < 2; pl++) {
                            const float2 U = plane2[pl * 128 + q], V = plane2[pl * 128 + 64 + q];
                            const float2 D = delay2[pl * 64 + q];
                            y[pl][0] = D.x * wh.y - U.x * wl.x;        // sample p
                            y[pl][1] = D.y * wh.x - U.y * wl.y;        // sample p + 1
                            y[pl][3] = D.x * wl.x + U.x * wh.y;        // sample 255 - p
                            y[pl][2] = D.y * wl.y + U.y * wh.x;        // sample 254 - p
                            delay2[pl * 64 + q] = V;
                        }
                        if (P.out_fmt == 1) {
                            float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(out_frame) + (size_t)p_blk * 512);
                            dst[q] = make_float4(y[0][0] + bias, y[1][0] + bias, y[0][1] + bias, y[1][1] + bias);
                            dst[127 - q] = make_float4(y[0][2] + bias, y[1][2] + bias, y[0][3] + bias, y[1][3] + bias);
                        } else {
                            // int16 (convert2s16.c:33-41): L R L R of samples p, p + 1 in one 8-byte store
                            auto two = [](float l, float r) {
                                return (uint32_t)(uint16_t)s16_of(l, true) | ((uint32_t)(uint16_t)s16_of(r, true) << 16);
                            };
                            uint2* dst = reinterpret_cast<uint2*>(reinterpret_cast<int16_t*>(out_frame) + (size_t)p_blk * 512);
                            dst[q] = make_uint2(two(y[0][0], y[1][0]), two(y[0][1], y[1][1]));
                            dst[127 - q] = make_uint2(two(y[0][2], y[1][2]), two(y[0][3], y[1][3]));
                        }
                    } else {
                        ola_store_generic(T, P, G, c, out_frame, p_blk, gt, nfch, nmain, p_uniform, lfe_on);
                    }
                    sync();
                    if (gt == 0) {
                        c->per_channel = p_uniform ? 0 : 1;
                        c->tr_ticket = 0;
                    }
                    pend = false;
                }
                if (!more || c->err) break;
                const int nfchans = c->nfchans;
                const uint32_t chincpl = c->chincpl;
                const uint32_t limit = c->limit_bit;

#ifndef A52_NO_STAGE_GATES
                if (sgated) { gate_arrive(bgate, true); gates_passed++; }
#endif
                // ================= E =================
                // (a block that repeats the last allocation has no exponents to decode and nothing to allocate)
                if (!c->repeat) {
                    int bad = 0, k = 0;
                    for (int a = 0; a < 7; a++) {
                        if (!c->expstr[a]) continue;
                        if ((k++ & 1) != w) continue;
                        uint8_t* e = G.exp + a * 256;
                        int dst = (a == 6) ? c->cplstrtmant : 1;
                        if (a != 6 && lane == 0) e[0] = c->exp_abs[a];
                        bad |= decode_exponents(T, W, limit, e + dst, c->expstr[a], c->exp_ngrp[a],
                                                c->exp_pos[a], c->exp_abs[a], lane);
                    }
                    if (bad && lane == 0) c->err = 1;
                    sync();
                    if (c->err) break;
                }

#ifndef A52_NO_STAGE_GATES
                if (sgated) { gate_arrive(bgate, true); gates_passed++; }
#endif
                // ================= B =================
                if (c->do_alloc) {
                    if (c->zero_alloc) {
                        for (int a = 0; a < 7; a++)
                            if ((c->do_alloc >> a) & 1) reinterpret_cast<uint32_t*>(G.bap + a * 256)[gt] = 0;
                    } else {
                        int16_t* scratch = reinterpret_cast<int16_t*>(G.plane);     // (the planes are rebuilt by the locate stage)
                        bit_allocate_block<NT>(T, c, c->do_alloc, G.exp, G.bap, scratch, scratch + 7 * 50, gt, sync);
                    }
                    sync();
                }

#ifndef A52_NO_STAGE_GATES
                if (sgated) { gate_arrive(bgate, true); gates_passed++; }
#endif
                // ================= L =================
                // What the locate passes produce is a PLAN of the block's mantissas: for every class (3-, 5- and
                // 11-level groups, plain fields, dithered zeros) the list of its members in coded order, each with
                // the shared-memory slot of its coefficient, its scale 2^-(15 + exponent) and the bit position of
                // its field.  The plan is a function of the baps, the coded ranges and the dither / coupling flags,
                // up to one common shift of the bit positions, and it lives in the pair's scratch in global memory
                // (L2-resident, read back coalesced).  A block that changes none of its inputs (every exponent set
                // and allocation reused: the rule, not the exception) skips the passes altogether and lets the
                // unpack stage add the shift.
                const bool rep = c->repeat != 0;
                if (!rep && blk > 0) cold_now++;
                const uint32_t bitpos = c->bitpos;
                uint32_t ta, tb, tz, mant_bits, pos_delta = 0;
                for (int i = gt; i < NPL * 256 / 4; i += NT)
                    reinterpret_cast<uint4*>(G.plane)[i] = make_uint4(0, 0, 0, 0);
                if (rep) {
                    ta = c->loc_ta; tb = c->loc_tb; tz = c->loc_tz; mant_bits = c->loc_mant;
                    pos_delta = bitpos - c->loc_bitpos;
                } else {
                    const uint32_t K = c->plan_K;
                    const uint32_t cpl_dith = chincpl & c->dithflag;
                    const uint32_t ncpl_dith = __popc(cpl_dith);
                    const int nseg = c->nseg;
                    uint32_t run_idx = 0, run_slot = 0, run_n = 0, zmode = 0;
                    bool mute = false;
                    {
                        int sgi = -1;
                        for (int k = 0; k < nseg; k++)
                            if (gt >= c->plan_lane0[k] && gt < c->plan_lane0[k + 1]) sgi = k;
                        if (sgi >= 0) {
                            const Segment sg = c->seg[sgi];
                            const uint32_t o = (gt - c->plan_lane0[sgi]) * K;
                            run_idx = sg.arr * 256 + sg.start + o;
                            run_slot = sg.plane * 256 + sg.start + o;
                            run_n = min(K, (uint32_t)sg.count - o);
                            zmode = (sg.arr == 6) ? (ncpl_dith ? 2u : 0u) : ((c->dithflag >> sg.arr) & 1u);
                            mute = (sg.arr == 5) && !c->out_lfe;
                        }
                    }
                    // pass 1: class counts of my run.  The bap bytes are fetched a 32-bit word at a time (runs
                    // start at any byte: each word is funnelled together with its predecessor); one 32-bit LUT
                    // word per bap carries 5-bit counters (3-, 5-, 11-level, zero) and the plain field bits.
                    const uint32_t* bapw = reinterpret_cast<const uint32_t*>(G.bap);
                    const uint32_t* expw = reinterpret_cast<const uint32_t*>(G.exp);
                    const uint32_t wi0 = run_idx >> 2, bsh = (run_idx & 3) * 8;
                    uint32_t cnt = 0;
                    {
                        const uint32_t lut32 = tab_base + (uint32_t)offsetof(Tables, cnt_lut32);
                        uint32_t wlo = bapw[wi0];
                        for (uint32_t k = 0, j = 1; k < K; k += 4, j++) {
                            const uint32_t whi = bapw[wi0 + j];
                            // bytes past my run select the all-zero row 16
                            const uint32_t vm = run_mask(run_n, k);
                            const uint32_t v = (__funnelshift_r(wlo, whi, bsh) & vm) | (0x10101010u & ~vm);
                            wlo = whi;
    #pragma unroll
                            for (int t = 0; t < 4; t++) {
                                uint32_t l;
                                asm("ld.shared.u32 %0, [%1];" : "=r"(l) : "r"(lut32 + prmt(v, 0, 0x4440 + t) * 4));
                                cnt += l;
                            }
                        }
                    }
                    const uint32_t n1 = cnt & 31, n2 = (cnt >> 5) & 31, n4 = (cnt >> 10) & 31, n0 = (cnt >> 15) & 31;
                    const uint32_t fixed = cnt >> 20;
                    const uint32_t np = run_n - n1 - n2 - n4 - n0;
                    const uint32_t nz = n0 * (zmode == 2 ? ncpl_dith : zmode);
                    const uint32_t pa = mute ? 0u : (n1 | (n2 << 16)), pb = mute ? 0u : (n4 | (np << 16));
                    uint32_t ia = warp_incl_scan(pa, lane), ib = warp_incl_scan(pb, lane);
                    uint32_t iz = warp_incl_scan(nz, lane);
                    if (lane == 31) { G.xch[w * 8 + 0] = ia; G.xch[w * 8 + 1] = ib; G.xch[w * 8 + 2] = iz; }
                    sync();
                    ta = G.xch[0] + G.xch[8]; tb = G.xch[1] + G.xch[9]; tz = G.xch[2] + G.xch[10];
                    if (w) { ia += G.xch[0]; ib += G.xch[1]; iz += G.xch[2]; }
                    const uint32_t e1 = (ia - pa) & 0xffff, e2 = (ia - pa) >> 16;
                    const uint32_t e4 = (ib - pb) & 0xffff, ep = (ib - pb) >> 16, ez = iz - nz;
                    const uint32_t t1 = ta & 0xffff, t2 = ta >> 16, t4 = tb & 0xffff;
                    const uint32_t p1 = e1 % 3, p2 = e2 % 3, p4 = e4 & 1;
                    const uint32_t s1 = (p1 + n1 + 2) / 3 - (p1 != 0);
                    const uint32_t s2 = (p2 + n2 + 2) / 3 - (p2 != 0);
                    const uint32_t s4 = (p4 + n4 + 1) / 2 - (p4 != 0);
                    const uint32_t mybits = fixed + 5 * s1 + 7 * (s2 + s4);
                    uint32_t ibits = warp_incl_scan(mybits, lane);
                    if (lane == 31) G.xch[w * 8 + 3] = ibits;
                    sync();
                    mant_bits = G.xch[3] + G.xch[11];
                    if (w) ibits += G.xch[3];
                    const uint32_t ng1 = (t1 + 2) / 3, ng2 = (t2 + 2) / 3;
                    if (!P.scan_only) {
                        uint32_t pos = min(bitpos + ibits - mybits, limit);
                        // exclusive occurrence counts of my run, per class, as 16-bit fields
                        const uint32_t base_lo = e1 | (e2 << 16), base_hi = e4 | (ep << 16);
                        const uint32_t base_z = ez;
                        // first group of each group class inside the one group section: 0, ng1, ng1 + ng2
                        const uint32_t gb_lo = ng1 << 16, gb_hi = ng1 + ng2;
                        uint32_t run_a = 0, run_z = 0;
                        const uint32_t lut_addr = tab_base + (uint32_t)offsetof(Tables, emit_lut);
                        const uint32_t lut2_addr = tab_base + (uint32_t)offsetof(Tables, emit_lut2);
                        const uint32_t zrow4 = (zmode == 1) ? 0x10101010u : 0u;
                        const uint32_t emit_bit = mute ? 0u : 0x1000000u;
                        // four mantissas per trip: one funnelled bap word and one exponent word, their LUT rows
                        // fetched together, then the four updates in coded order
                        // emit_lut row: x: counter increment (classes 1, 2, 4, plain); y: count selector A |
                        // width << 16 | emit << 24; z: count selector B; w: run-counter selector | period << 16 |
                        // zero-class increment << 24
                        // emit_lut2 row: x: 2^17 / period; y: byte offset of the class's plan section; z: entry
                        // stride | offset of the member word << 8 | group-base selector << 16; w: constant bits of
                        // the entry's position word (class, 32 - width, value table), bit 31 = has a position word
                        auto emit_one = [&](uint32_t e, uint32_t slot, const uint4& L, const uint4& M) {
                            const uint32_t cls_cnt = prmt(run_a, run_z, L.w);
                            const uint32_t occ = prmt(prmt(base_lo, base_hi, L.y), base_z, L.z) + cls_cnt;
                            const uint32_t per = prmt(L.w, 0, 0x4442);
                            const uint32_t g = (occ * M.x) >> 17;               // occ / period (occ < 2^15)
                            const uint32_t r = occ - g * per;
                            const uint32_t u = prmt(gb_lo, gb_hi, M.z >> 16) + g;
                            const uint32_t eb = M.y + u * (M.z & 0xffu);        // the entry
                            const uint32_t ea = eb + ((M.z >> 8) & 0xffu) + 4u * r;   // this member's word
                            run_a += L.x;
                            run_z += L.w >> 24;
                            const uint32_t emit = L.y & emit_bit;
                            A52_CHECK(!emit || (ea < (uint32_t)kPlanBytes && eb < (uint32_t)kPlanBytes && slot < NPL * 256u && e <= 24u &&
                                               (eb < (uint32_t)kPlanPlainOff ? u < (uint32_t)kPlanGroups
                                                : eb < (uint32_t)kPlanZeroOff ? occ < (uint32_t)kPlanPlain : occ < (uint32_t)kPlanZeros)), 101);
                            stg_u32_if(plan + ea, (slot << 2) | ((112u - e) << 23), emit);
                            stg_u32_if(plan + eb, pos | M.w, (r == 0 && (int32_t)M.w < 0) ? emit : 0u);
                            pos = min(pos + (r == 0 ? prmt(L.y, 0, 0x4442) : 0u), limit);
                        };
                        uint32_t bw_lo = bapw[wi0], ew_lo = expw[wi0];
                        if (zmode != 2) {
                            for (uint32_t k0 = 0, j = 1; k0 < K; k0 += 4, j++) {
                                const uint32_t bw_hi = bapw[wi0 + j], ew_hi = expw[wi0 + j];
                                const uint32_t bv = __funnelshift_r(bw_lo, bw_hi, bsh), ev = __funnelshift_r(ew_lo, ew_hi, bsh);
                                bw_lo = bw_hi;
                                ew_lo = ew_hi;
                                uint4 Lr[4], Mr[4];
                                // mantissas past my run take the row of an undithered zero: nothing moves
                                const uint32_t rows = (bv + zrow4) & run_mask(run_n, k0);
    #pragma unroll
                                for (int t = 0; t < 4; t++) {
                                    Lr[t] = lds_v4(lut_addr + prmt(rows, 0, 0x4440 + t) * 16);
                                    Mr[t] = lds_v4(lut2_addr + prmt(rows, 0, 0x4440 + t) * 16);
                                }
    #pragma unroll
                                for (int t = 0; t < 4; t++)
                                    emit_one((ev >> (8 * t)) & 0xff, run_slot + k0 + t, Lr[t], Mr[t]);
                            }
                        } else {
                            // coupling channel with dither: a bap-0 bin takes one dither value per coupled
                            // channel, in channel order (parse.c:466-481)
                            for (uint32_t k = 0; k < run_n; k++) {
                                const uint32_t b = G.bap[run_idx + k], e = G.exp[run_idx + k], slot = run_slot + k;
                                if (b == 0) {
                                    uint32_t m = cpl_dith;
                                    while (m) {
                                        const uint32_t ch = __ffs(m) - 1;
                                        m &= m - 1;
                                        const uint32_t s2 = ch * 256 + (slot & 255);
                                        A52_CHECK(base_z + run_z < (uint32_t)kPlanZeros && s2 < NPL * 256u, 102);
                                        stg_u32_if(plan + kPlanZeroOff + 4u * (base_z + run_z), (s2 << 2) | ((112u - e) << 23), 1u);
                                        run_z++;
                                    }
                                } else {
                                    emit_one(e, slot, lds_v4(lut_addr + b * 16), lds_v4(lut2_addr + b * 16));
                                }
                            }
                        }
                    }
                    // the last group of a class may be short: its missing members unpack into a dump word
                    if (gt < 3 && !P.scan_only) {
                        const uint32_t tcl = gt == 0 ? t1 : gt == 1 ? t2 : t4, per = gt == 2 ? 2u : 3u;
                        const uint32_t gb = gt == 0 ? 0u : gt == 1 ? ng1 : ng1 + ng2;
                        const uint32_t full = tcl / per, rem = tcl - full * per;
                        if (rem)
                            for (uint32_t rr = rem; rr < per; rr++)
                                {
                                    A52_CHECK(gb + full < (uint32_t)kPlanGroups, 103);
                                    stg_u32_if(plan + 16u * (gb + full) + 4u + 4u * rr, kDumpWord, 1u);
                                }
                    }
                    if (gt == 0) {
                        c->loc_ta = ta; c->loc_tb = tb; c->loc_tz = tz; c->loc_mant = mant_bits;
                        c->loc_bitpos = bitpos;
                        c->loc_valid = 1;
                    }
                }
                sync();
                if (gt == 0) c->bitpos = bitpos + mant_bits;
                frame_draws += tz;
                if (P.scan_only) continue;        // a scan stops here: the counts are all it wants of the block
                const uint32_t t1 = ta & 0xffff, t2 = ta >> 16, t4 = tb & 0xffff, tp = tb >> 16;

#ifndef A52_NO_STAGE_GATES
                if (sgated) { gate_arrive(bgate, true); gates_passed++; }
#endif
                // ================= U =================
                // Every class reads its part of the plan back in coalesced 16-byte (groups, pairs of plain fields) or
                // 4-byte (zeros: one row of 32 per load, eight rows in flight) loads; a member word is
                // slot << 2 | scale bits, so a coefficient costs a table or shift dequantisation, one multiply and
                // one store.
                {
                    // every first load of the three classes is issued before anything is used: one exposure of the
                    // L2 latency per block instead of three
                    const uint32_t ngroups = (t1 + 2) / 3 + (t2 + 2) / 3 + (t4 + 1) / 2;
                    const uint4* gp = reinterpret_cast<const uint4*>(plan) + gt;
                    const uint4* pp = reinterpret_cast<const uint4*>(plan + kPlanPlainOff) + gt;
                    const uint32_t* zp = reinterpret_cast<const uint32_t*>(plan + kPlanZeroOff) + 32 * w + lane;
                    const uint16_t* ds = P.dither_seq + dither_index + 1 + 32 * w + lane;
                    uint4 Gn = make_uint4(0, 0, 0, 0), Pn = make_uint4(0, 0, 0, 0);
                    if (gt < ngroups) Gn = __ldcg(gp);
                    if (2 * gt < tp) Pn = __ldcg(pp);

                    // zeros: rows of 32 alternate between the warps; zero k takes value dither_index + 1 + k of the
                    // generator's sequence (parse.c:310-319), read from the (wrapped) table
                    const uint32_t nrows = (tz + 31) >> 5;
                    for (uint32_t r0 = w; r0 < nrows; r0 += 16, zp += 512, ds += 512) {
                        uint32_t E[8], D[8];
    #pragma unroll
                        for (int i = 0; i < 8; i++) {
                            E[i] = __ldcg(zp + 64 * i);
                            D[i] = __ldg(ds + 64 * i);
                        }
    #pragma unroll
                        for (int i = 0; i < 8; i++) {
                            const uint32_t k = 32 * (r0 + 2 * i) + lane;
                            const int dv = (3 * (int)(int16_t)D[i]) >> 2;
                            const float v = (float)dv * __uint_as_float(E[i] & 0x7f800000u);
                            A52_CHECK(k >= tz || (E[i] & 0x1fffu) < NPL * 1024u, 201);
                            sts_f32_if(plane_sa + (E[i] & 0x1fffu), v, k < tz ? 1u : 0u);
                        }
                    }
                    dither_index = (dither_index + tz) % kDitherPeriod;

                    const uint32_t w_sa = pair_sa + (uint32_t)PairLayout<NPL>::fbuf;
                    auto window = [&](uint32_t pw) {          // the 32 bits at the entry's (shifted, clamped) position
                        const uint32_t pos = min((pw & 0x7fffu) + pos_delta, limit);
                        const uint32_t a = w_sa + ((pos >> 5) << 2);
                        A52_CHECK(a + 8 <= w_sa + (uint32_t)P.fbuf_bytes, 202);
                        return __funnelshift_l(lds_u32(a + 4), lds_u32(a), pos);
                    };
                    // groups: position word (pos | class << 15 | (32 - width) << 19 | table << 24) + three member words
                    const uint32_t q_sa = tab_base + (uint32_t)offsetof(Tables, q1);
                    for (uint32_t g = gt; g < ngroups; g += NT) {
                        const uint4 E = Gn;
                        if (g + NT < ngroups) Gn = __ldcg(gp + (g + NT - gt));
                        const uint32_t code = window(E.x) >> ((E.x >> 19) & 31u);
                        const uint32_t ta0 = q_sa + ((E.x >> 24) & 0x7fu) * 64u + code * 2u;
                        const uint32_t st = (E.x & 0x18000u) ? 256u : 64u;       // digit stride of the value table
                        const float v0 = (float)lds_s16(ta0) * __uint_as_float(E.y & 0x7f800000u);
                        const float v1 = (float)lds_s16(ta0 + st) * __uint_as_float(E.z & 0x7f800000u);
#ifdef A52_BOUNDS_CHECK
                        {
                            const uint32_t nmem = ((E.x & 0x18000u) != 0x10000u) ? 3u : 2u;      // members of the group
                            A52_CHECK((E.y & 0x1fffu) <= NPL * 1024u && (E.z & 0x1fffu) <= NPL * 1024u &&
                                      (nmem < 3 || (E.w & 0x1fffu) <= NPL * 1024u), 203);
                            A52_CHECK(((E.x >> 24) & 0x7fu) * 64u + code * 2u + (nmem - 1) * st <
                                      (uint32_t)(sizeof(T.q1) + sizeof(T.q2) + sizeof(T.q4)), 205);
                        }
#endif
                        sts_f32_if(plane_sa + (E.y & 0x1fffu), v0, 1u);
                        sts_f32_if(plane_sa + (E.z & 0x1fffu), v1, 1u);
                        if ((E.x & 0x18000u) != 0x10000u) {                        // 11-level codes hold two values
                            const float v2 = (float)lds_s16(ta0 + 2 * st) * __uint_as_float(E.w & 0x7f800000u);
                            sts_f32_if(plane_sa + (E.w & 0x1fffu), v2, 1u);
                        }
                    }
                    // plain fields, two entries per load: position word (.. | bit 30: value table) + member word
                    const uint32_t q35_sa = tab_base + (uint32_t)offsetof(Tables, q35);
                    auto plain_one = [&](uint32_t pw, uint32_t mw, uint32_t pred) {
                        const uint32_t v = window(pw), sh = (pw >> 19) & 31u;
                        int q = ((int)(v & (0xffffffffu << sh))) >> 16;
                        if (pw & 0x40000000u) q = lds_s16(q35_sa + ((pw >> 24) & 0x3fu) * 2u + (v >> sh) * 2u);
                        A52_CHECK(!pred || (mw & 0x1fffu) < NPL * 1024u, 204);
                        sts_f32_if(plane_sa + (mw & 0x1fffu), (float)q * __uint_as_float(mw & 0x7f800000u), pred);
                    };
                    for (uint32_t k = 2 * gt; k < tp; k += 2 * NT) {
                        const uint4 E = Pn;
                        if (k + 2 * NT < tp) Pn = __ldcg(pp + ((k >> 1) + NT - gt));
                        plain_one(E.x, E.y, 1u);
                        plain_one(E.z, E.w, k + 1 < tp ? 1u : 0u);
                    }
                }
                sync();
                if (blk == 5 && f + 1 < f1) {
                    if (gt == 0) issue_frame_load(P, G, f + 1);
                    next_issued = true;
                }

                const bool unif = c->uniform_path;
                if (!unif) {
                    for (int ch = 0; ch < nfchans; ch++) {
                        const float g1 = c->gain[ch];
                        const int end = c->endmant[ch];
                        for (int bin = gt; bin < end; bin += NT) G.plane[ch * 256 + bin] *= g1;
                    }
                }
                if (c->out_lfe && gt < 7) G.plane[5 * 256 + gt] *= c->gain[5];
                if (!unif || c->out_lfe) sync();
                if (chincpl) {
                    const int first = __ffs(chincpl) - 1;
                    for (int bin = c->cplstrtmant + gt; bin < c->cplendmant; bin += NT) {
                        int sub = (bin - c->cplstrtmant) / 12;
                        int bnd = sub - __popc(c->cplbndstrc & ((1u << sub) - 1));
                        bool zero_bap = (G.bap[6 * 256 + bin] == 0);
                        float cv = G.plane[first * 256 + bin];
                        for (int ch = nfchans - 1; ch >= 0; ch--) {
                            if (!((chincpl >> ch) & 1)) continue;
                            float co = unif ? c->cplco[ch][bnd] : c->cplco[ch][bnd] * c->gain[ch];
                            float src = zero_bap ? G.plane[ch * 256 + bin] : cv;
                            G.plane[ch * 256 + bin] = src * co;
                        }
                    }
                    sync();
                }
                if (c->acmod == 2 && c->rematflg) {
                    int end = min(c->endmant[0], c->endmant[1]);
                    for (int bin = 13 + gt; bin < end; bin += NT) {
                        int band = (bin >= 61) ? 3 : (bin >= 37) ? 2 : (bin >= 25) ? 1 : 0;
                        if ((c->rematflg >> band) & 1) {
                            float a = G.plane[bin], b = G.plane[256 + bin];
                            G.plane[bin] = a + b;
                            G.plane[256 + bin] = a - b;
                        }
                    }
                    sync();
                }
                if (P.dbg_exp) {
                    size_t o = ((size_t)f * 6 + blk) * 7 * 256;
                    for (int i = gt; i < 7 * 256; i += NT) {
                        P.dbg_exp[o + i] = G.exp[i];
                        P.dbg_bap[o + i] = G.bap[i];
                    }
                }
                if (P.dbg_coef) {
                    size_t o = ((size_t)f * 6 + blk) * 6 * 256;
                    for (int i = gt; i < 6 * 256; i += NT) {
                        int pl = i >> 8;
                        bool live = (pl < nfchans) || (pl == 5 && c->out_lfe);
                        float v = live ? G.plane[i] : 0.f;
                        if (unif && pl < 5) v *= c->gain[pl];
                        P.dbg_coef[o + i] = v;
                    }
                }
                if (P.dbg_info && gt == 0) {
                    int32_t* o = P.dbg_info + ((size_t)f * 6 + blk) * 16;
                    for (int i = 0; i < 5; i++) o[i] = c->endmant[i];
                    o[5] = c->cplstrtmant; o[6] = c->cplendmant; o[7] = c->chincpl;
                    o[8] = P.dither_seq[dither_index]; o[9] = c->acmod; o[10] = c->lfeon;
                    o[11] = c->output;
                    o[12] = c->blksw | (c->uniform_path << 8) | ((c->clev == 0.f) << 9) | ((c->slev == 0.f) << 10);
                    o[13] = c->ncplbnd; o[14] = c->rematflg;
                    o[15] = c->csnroffst;
                }

#ifndef A52_NO_STAGE_GATES
                if (sgated) { gate_arrive(bgate, true); gates_passed++; }
#endif
                // ================= M =================
                const int nmain = c->nout;
                const bool uniform = c->uniform_path;
                const int lfe_on = c->out_lfe;
                const bool fast2 = uniform && nmain == 2 && !lfe_on && c->blksw == 0 && !c->per_channel &&
                                   P.out_fmt != 0;      // (two channels: libao's WAV order is liba52's)
                if (fast2) {
                    mix2_pairs(G.plane, plane_sa, c, gt);
                } else if (uniform) {
                    switch (nmain) {
                    case 1: mix_planes<1, NT>(G.plane, c, gt); break;
                    case 2: mix_planes<2, NT>(G.plane, c, gt); break;
                    case 3: mix_planes<3, NT>(G.plane, c, gt); break;
                    default: mix_planes<4, NT>(G.plane, c, gt); break;
                    }
                }

                // the transforms wait for the next pass: what they need of this block
                {
                    const int ntr = uniform ? nmain : nfchans;
                    uint32_t live = 0;
                    for (int pl = 0; pl < 6; pl++) {
                        if (pl < 5 ? (pl >= ntr) : !lfe_on) continue;
                        if (!uniform && pl < 5 && c->gain[pl] == 0.f) continue;
                        live |= 1u << pl;
                    }
                    p_live = live;
                    p_blksw = c->blksw;
                    p_uniform = uniform;
                    p_fast = fast2;
                    p_blk = blk;
                    pend = true;
                }
                sync();
            }   // blocks

            if (bgated) {
                for (; gates_passed < gates_due; gates_passed++) gate_arrive(bgate, false);    // a frame cut short still counts
                gate_leave(bgate);
            }
            cold_prev = cold_now;
            if (frame_ok && blk < 6) frame_status = 16 + blk;
            if (P.scan_only) {
                if (gt == 0) {
                    P.scan[f].dither_draws = frame_draws;
                    P.scan[f].status = frame_status;
                }
            } else if (frame_status) {
                int nout = frame_ok ? (c->nout + c->out_lfe) : P.nout_req;
                int ssz = (P.out_fmt >= 2) ? 2 : 4;
                size_t from = (size_t)blk * 256 * nout * ssz;
                size_t to = (size_t)6 * 256 * nout * ssz;
                for (size_t i = from + gt * 4; i < to; i += NT * 4)
                    *reinterpret_cast<uint32_t*>(out_frame + i) = 0;
            }
            if (frame_ok && !(P.req_flags & 0x100) && !P.scan_only) {
                // a frame granted fewer channels than the request's stride holds: the rest of its slot is silence
                const int ssz = (P.out_fmt >= 2) ? 2 : 4;
                const size_t used = (size_t)1536 * (c->nout + c->out_lfe) * ssz;
                for (size_t i = used + gt * 4; i < P.frame_stride; i += NT * 4)
                    *reinterpret_cast<uint32_t*>(out_frame + i) = 0;
            }
            if (gt == 0 && P.status && !shadow && !P.scan_only) P.status[f] = frame_status;
            sync();
            if (!next_issued && f + 1 < f1 && gt == 0) issue_frame_load(P, G, f + 1);
            sync();
        }   // frames
        if (gated) gate_leave(gate);

        // the state after the unit's last frame: for the next slice of the chain, or (frame-independent slices: from
        // the slice that holds the stream's last frame) for the caller
        if (P.carry && !P.scan_only && (chained || (f1 == fs1 && (f0 < f1 || slice == 0)))) {
            for (int i = gt; i < ndelay * 128; i += NT) P.carry[s].delay[i >> 7][i & 127] = G.delay[i];
            if (gt == 0) {
                P.carry[s].dither_index = dither_index;
                P.carry[s].per_channel = c->per_channel;
            }
        }
        if (P.nslices > 1 && chained) {
            __threadfence();
            sync();
            if (gt == 0) atomicExch(P.slice_done + s, (int)slice + 1);
        }
        sync();
    }
}

}  // namespace a52

#include "a52_host.inl"
#include "a52_imdct_ab.inl"
