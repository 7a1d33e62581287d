// a52_decode.cu - batched AC-3 (ATSC A/52) decode for NVIDIA B200 (sm_100a).
//
// One persistent kernel decodes thousands of independent AC-3 streams.  A
// "group" of 128 threads (4 warps) owns one stream at a time and walks its
// sync frames in order; everything between the staged bitstream and the PCM
// store lives in shared memory / registers:
//
//   stage   frame bytes  : TMA bulk copy (cp.async.bulk + mbarrier), double buffered
//   parse   BSI + audio-block side info        (reference: liba52/parse.c:131-205, 558-804)
//   exps    exponent groups -> exponents       (parse.c:218-270)   lanes = groups, warp scan
//   alloc   parametric bit allocation          (bit_allocate.c:124-265) one warp per channel
//   locate  per-bin field positions            (parse.c:336-433)   prefix sums of bit widths
//   unpack  mantissa extract + dequantise + dither (parse.c:310-433)
//   couple  coupling fan-out, rematrix         (parse.c:435-556, 837-865)
//   mix     downmix (coefficient or time domain) (downmix.c:162-619)
//   imdct   512 / 2x256 transform: pre-twiddle, radix-4 FFT in shared memory,
//           post-twiddle                        (imdct.c:258-345)
//   ola     KBD window + overlap-add + PCM store (imdct.c:276-292)
//
// The overlap-add tail and the dither generator position are carried on chip
// from frame to frame of a stream (and in/out of the call through
// a52_stream_carry_t), so frames never wait on another group.
//
// Integer stages are bit-exact with liba52; the float transform uses a
// different FFT factorisation (tolerance 1e-5 relative RMS, measured ~1e-7).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "a52_common.cuh"

namespace a52 {

// ---------------------------------------------------------------------------
// constant memory (warp-uniform lookups only)
// ---------------------------------------------------------------------------
__constant__ ModeEntry c_mode[9 * 16];     // [acmod_ext 0..8][cmixlev*4 + surmixlev]
__constant__ MixEntry  c_mix[8 * 11];      // [acmod][output mode]
__constant__ uint8_t   c_nfchans[8] = {2, 1, 2, 3, 3, 4, 4, 5};
__constant__ uint16_t  c_bitrate[19] = {32, 40, 48, 56, 64, 80, 96, 112, 128, 160, 192, 224, 256, 320,
                                        384, 448, 512, 576, 640};
__constant__ int16_t   c_sgain[4]  = {0x540, 0x4d8, 0x478, 0x410};
__constant__ int16_t   c_dbknee[4] = {0x000, 0x700, 0x900, 0xb00};
__constant__ int16_t   c_floor[8]  = {0x2f0, 0x2b0, 0x270, 0x230, 0x1f0, 0x170, 0x0f0, -0x800};

__device__ Tables g_tables;                // filled by the host once per context

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
    uint8_t  expstr[7];        // this block's strategies (0 = reuse)
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
    // ---- scan scratch ----
    uint32_t scan_a[kGroupWarps], scan_b[kGroupWarps], scan_c[kGroupWarps];
    uint32_t blk_dither;       // dither calls of the block
    uint32_t mant_bits;        // total mantissa bits of the block
};

struct GroupPtrs {
    GroupCtl* ctl;
    uint8_t*  exp;     // [7][256]
    uint8_t*  bap;     // [7][256]  standard numbering 0..15
    int16_t*  band;    // [4 warps][2][50] bit-allocation scratch
    uint8_t*  grp;     // group codes: [0,512) bap1, [512,1024) bap2, [1024,1792) bap4
    float*    plane;   // [6][256]
    float*    delay;   // [ndelay][128]
    uint32_t* fbuf[2]; // staged frames (as native-endian 32-bit words after the swap pass)
    uint64_t* mbar;    // [2]
};

constexpr int kGrpOff2 = 512, kGrpOff4 = 1024, kGrpBytes = 1792;

__host__ __device__ inline int align16(int x) { return (x + 15) & ~15; }

__host__ __device__ inline int group_smem_bytes(int fbuf_bytes, int ndelay)
{
    int n = 0;
    n += align16((int)sizeof(GroupCtl));
    n += 7 * 256 * 2;
    n += align16(kGroupWarps * 2 * 50 * 2);
    n += kGrpBytes;
    n += 6 * 256 * 4;
    n += ndelay * 128 * 4;
    n += 2 * fbuf_bytes;
    n += 16;
    return n;
}

__device__ inline GroupPtrs carve(uint8_t* base, int fbuf_bytes, int ndelay)
{
    GroupPtrs g;
    g.ctl = reinterpret_cast<GroupCtl*>(base);  base += align16((int)sizeof(GroupCtl));
    g.exp = base;                               base += 7 * 256;
    g.bap = base;                               base += 7 * 256;
    g.band = reinterpret_cast<int16_t*>(base);  base += align16(kGroupWarps * 2 * 50 * 2);
    g.grp = base;                               base += kGrpBytes;
    g.plane = reinterpret_cast<float*>(base);   base += 6 * 256 * 4;
    g.delay = reinterpret_cast<float*>(base);   base += ndelay * 128 * 4;
    g.fbuf[0] = reinterpret_cast<uint32_t*>(base); base += fbuf_bytes;
    g.fbuf[1] = reinterpret_cast<uint32_t*>(base); base += fbuf_bytes;
    g.mbar = reinterpret_cast<uint64_t*>(base);
    return g;
}

// ---------------------------------------------------------------------------
// small PTX helpers
// ---------------------------------------------------------------------------
__device__ __forceinline__ void group_sync(int gid)
{
    asm volatile("bar.sync %0, %1;" :: "r"(gid + 1), "n"(kGroupThreads) : "memory");
}

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

// bytes to stage for frame f: from the 16-byte aligned address at or below its
// offset up to the start of the next frame (or the end of the buffer), capped
__device__ __forceinline__ uint32_t stage_bytes(const DecodeParams& P, uint32_t f)
{
    uint64_t off = P.frame_off[f], nxt = P.frame_off[f + 1];
    uint64_t a0 = off & ~(uint64_t)15;
    uint64_t end = (nxt > off) ? nxt : P.es_bytes;
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
                                  const DecodeParams& P, uint32_t avail_bytes)
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

    ModeEntry me = c_mode[acmod_ext * 16 + cmix * 4 + smix];
    if (me.output < 0) return 2;                               // A52_ST_BAD_FRAME
    c->output = me.output;
    c->out_lfe = (c->lfeon && (P.req_flags & M_LFE)) ? 1 : 0;
    if (c->out_lfe) c->output |= M_LFE;
    c->nout = c_mix[acmod * 11 + me.output].nout;
    c->level = me.level * 2.0f;                                // parse.c:169
    c->dynrng = c->level;
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
}

// ---------------------------------------------------------------------------
// audio block side information (parse.c:570-804), one thread
// ---------------------------------------------------------------------------
__device__ int parse_block(GroupCtl* c, const uint32_t* w, const DecodeParams& P)
{
    BitReader br{w, c->bitpos, c->limit_bit};
    const int nfchans = c->nfchans;
    const int acmod = c->acmod;

    uint32_t v = br.get(2 * nfchans);          // blksw[nfchans], dithflag[nfchans]
    uint32_t blksw = 0, dith = 0;
    for (int i = 0; i < nfchans; i++) {
        blksw |= ((v >> (2 * nfchans - 1 - i)) & 1) << i;
        dith  |= ((v >> (nfchans - 1 - i)) & 1) << i;
    }
    c->blksw = blksw;
    c->dithflag = dith;

    int reps = acmod ? 1 : 2;                  // dynrng (parse.c:578-598)
    while (reps--) {
        if (br.get(1)) {
            int d = br.get_signed(8);
            if (!P.drc_off) {
                float range = (float)(((d & 0x1f) | 0x20) << 13) * pow2neg(15 + 3 - (d >> 5));
                c->dynrng = c->level * range;
            }
        }
    }

    if (br.get(1)) {                           // cplstre (parse.c:600-634)
        c->chincpl = 0;
        if (br.get(1)) {
            uint32_t m = br.get(nfchans), chincpl = 0;
            for (int i = 0; i < nfchans; i++) chincpl |= ((m >> (nfchans - 1 - i)) & 1) << i;
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
            if ((chincpl >> i) & 1)
                if (br.get(1)) {
                    any = 1;
                    int mstr = 3 * br.get(2);
                    for (int j = 0; j < c->ncplbnd; j++) {
                        int e = br.get(4), m = br.get(4);
                        m = (e == 15) ? m << 14 : (m | 0x10) << 13;
                        c->cplco[i][j] = (float)m * pow2neg(15 + e + mstr);
                    }
                }
        if (acmod == 2 && c->phsflginu && any)
            for (int j = 0; j < c->ncplbnd; j++)
                if (br.get(1)) c->cplco[1][j] = -c->cplco[1][j];
    }

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
    for (int i = 0; i < nfchans; i++) expstr[i] = br.get(2);
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
    uint32_t do_alloc = 0;
    if (expstr[6]) {
        int ngrp = (c->cplendmant - c->cplstrtmant) / (3 << (expstr[6] - 1));
        do_alloc |= 64;
        c->exp_abs[6] = br.get(4) << 1;
        c->exp_pos[6] = br.pos;
        c->exp_ngrp[6] = ngrp;
        br.skip(7 * ngrp);
    }
    for (int i = 0; i < nfchans; i++)
        if (expstr[i]) {
            int gsz = 3 << (expstr[i] - 1);
            int ngrp = (c->endmant[i] + gsz - 4) / gsz;
            do_alloc |= 1u << i;
            c->exp_abs[i] = br.get(4);
            c->exp_pos[i] = br.pos;
            c->exp_ngrp[i] = ngrp;
            br.skip(7 * ngrp + 2);             // + gainrng
        }
    if (expstr[5]) {
        do_alloc |= 32;
        c->exp_abs[5] = br.get(4);
        c->exp_pos[5] = br.pos;
        c->exp_ngrp[5] = 2;
        br.skip(14);
    }
    for (int i = 0; i < 7; i++) c->expstr[i] = expstr[i];

    // bit allocation side info (parse.c:738-772)
    if (br.get(1)) { do_alloc = 127; c->bai = br.get(11); }
    if (br.get(1)) {
        do_alloc = 127;
        c->csnroffst = br.get(6);
        if (chincpl) c->chbai[6] = br.get(7);
        for (int i = 0; i < nfchans; i++) c->chbai[i] = br.get(7);
        if (c->lfeon) c->chbai[5] = br.get(7);
    }
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
            int nseg = br.get(3) + 1, band = 0;
            while (nseg--) {
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
    c->do_alloc = do_alloc;

    if (br.get(1)) {                           // skip field (parse.c:800-804)
        uint32_t n = br.get(9);
        br.skip(8 * n);
    }
    c->bitpos = br.pos;

    compute_gains(c);

    // coded order of the mantissas (parse.c:816-835, 867-879)
    int ns = 0, done_cpl = 0;
    uint32_t flat = 0;
    for (int i = 0; i < nfchans; i++) {
        Segment s;
        s.arr = i; s.plane = i; s.start = 0; s.dith = (dith >> i) & 1;
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
__device__ int decode_exponents(const uint32_t* w, uint32_t limit, uint8_t* dst, int strategy,
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
            if (code >= 125) bad = 1;
            d0 = code / 25; d1 = (code / 5) % 5; d2 = code % 5;
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

__device__ void bit_allocate_warp(const Tables& T, const GroupCtl* c, int arr, const uint8_t* exp,
                                  uint8_t* bap, int16_t* bndpsd, int16_t* mask, int lane)
{
    int start = 0, end;
    int fastleak = 0, slowleak = 0;
    const bool is_lfe = (arr == 5);
    if (arr == 6) {
        start = c->cplstrtmant; end = c->cplendmant;
        fastleak = (c->cplfleak << 8) + 768;
        slowleak = (c->cplsleak << 8) + 768;
    } else if (arr == 5) {
        end = 7;
    } else {
        end = c->endmant[arr];
    }
    const int bai = c->bai, chbai = c->chbai[arr], half = c->halfrate;
    const int sdecay = (0x0f + 2 * (bai >> 9)) >> half;
    const int fdecay = (0x3f + 0x14 * ((bai >> 7) & 3)) >> half;
    const int sgain = c_sgain[(bai >> 5) & 3];
    const int dbknee = c_dbknee[(bai >> 3) & 3];
    const int floorv = c_floor[bai & 7];
    const int fgain = 0x80 * ((chbai & 7) + 1);
    const int snroffset = (((c->csnroffst - 15) << 4) + (chbai >> 3)) << 2;
    const int deltbae = c->deltbae[arr];
    const int bndstrt = T.masktab[start];
    const int bndend = T.masktab[end - 1] + 1;

    // 1. band PSD integration (log-addition is not associative: serial inside a band)
    for (int band = bndstrt + lane; band < bndend; band += 32) {
        int b0 = max((int)T.bndtab[band], start);
        int b1 = min((int)T.bndtab[band + 1], end);
        int v = 3072 - (exp[b0] << 7);
        for (int bin = b0 + 1; bin < b1; bin++) {
            int p = 3072 - (exp[bin] << 7);
            int d = v - p;
            int adr = min(abs(d) >> 1, 255);
            v = max(v, p) + T.latab[adr];
        }
        bndpsd[band] = (int16_t)v;
    }
    __syncwarp();

    // 2. excitation (serial recurrences; one lane), result left in mask[]
    if (lane == 0) {
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
    __syncwarp();

    // 3. masking curve, delta, snr offset (per band)
    for (int band = bndstrt + lane; band < bndend; band += 32) {
        int v = mask[band];
        int p = bndpsd[band];
        if (p < dbknee) v += (dbknee - p) >> 2;
        v = max(v, (int)T.hth[c->fscod * 50 + (band >> half)]);
        if (deltbae == 0 || deltbae == 1) v += c->deltba[arr][band] * 128;
        v -= snroffset + floorv;
        v = max(v, 0) & 0x1fe0;
        mask[band] = (int16_t)(v + floorv);
    }
    __syncwarp();

    // 4. pointer lookup per bin
    for (int bin = start + lane; bin < end; bin += 32) {
        int p = 3072 - (exp[bin] << 7);
        int a = (p - mask[T.masktab[bin]]) >> 5;
        a = min(max(a, 0), 63);
        bap[bin] = T.baptab[a];
    }
}

// ---------------------------------------------------------------------------
// group-wide exclusive scans (128 threads)
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

// ---------------------------------------------------------------------------
// mantissa descriptors.  A 32-bit word per bin, written into the coefficient
// plane slot the coefficient itself will later occupy:
//   [4:0] exponent  [8:5] bap  [10:9] digit (grouped)  [11] dither enable
//   bap 0        : [31:16] dither value (int16)
//   bap 1,2,4    : [22:12] group index (code sits in the group-code array)
//   other        : [26:12] bit position of the field
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t make_desc(uint32_t exp, uint32_t bap, uint32_t digit, uint32_t payload)
{
    return exp | (bap << 5) | (digit << 9) | (payload << 12);
}

struct Cursor {           // walks the coded-order sequence
    int seg;
    int bin;              // bin within the segment's array
    int left;             // bins left in the segment
};

__device__ __forceinline__ void cursor_seek(Cursor& cu, const GroupCtl* c, uint32_t flat)
{
    int s = 0;
    for (int k = 1; k < c->nseg; k++)
        if (flat >= c->seg[k].first) s = k;
    cu.seg = s;
    uint32_t off = flat - c->seg[s].first;
    cu.bin = c->seg[s].start + off;
    cu.left = c->seg[s].count - off;
}

__device__ __forceinline__ void cursor_next(Cursor& cu, const GroupCtl* c)
{
    cu.bin++;
    if (--cu.left == 0 && cu.seg + 1 < c->nseg) {
        cu.seg++;
        cu.bin = c->seg[cu.seg].start;
        cu.left = c->seg[cu.seg].count;
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
    z[lane] = x[0];
    z[lane + 32] = cmul(x[1], T.wfft[lane]);
    z[lane + 64] = cmul(x[2], T.wfft[2 * lane]);
    z[lane + 96] = cmul(x[3], T.wfft[(3 * lane) & 127]);
    __syncwarp();
    // stage 2: 4 blocks of 32, stride 8, twiddle W32^(j p) = W128^(4 j p)
    {
        int b = lane >> 3, j = lane & 7, base = b * 32 + j;
        float2 y0 = z[base], y1 = z[base + 8], y2 = z[base + 16], y3 = z[base + 24];
        bfly4(y0, y1, y2, y3);
        __syncwarp();
        z[base] = y0;
        z[base + 8] = cmul(y1, T.wfft[4 * j]);
        z[base + 16] = cmul(y2, T.wfft[8 * j]);
        z[base + 24] = cmul(y3, T.wfft[12 * j]);
    }
    __syncwarp();
    // stage 3: 16 blocks of 8, stride 2, twiddle W8^(j p) = W128^(16 j p)
    {
        int b = lane >> 1, j = lane & 1, base = b * 8 + j;
        float2 y0 = z[base], y1 = z[base + 2], y2 = z[base + 4], y3 = z[base + 6];
        bfly4(y0, y1, y2, y3);
        __syncwarp();
        z[base] = y0;
        z[base + 2] = cmul(y1, T.wfft[16 * j]);
        z[base + 4] = cmul(y2, T.wfft[32 * j]);
        z[base + 6] = cmul(y3, T.wfft[48 * j]);
    }
    __syncwarp();
    // stage 4: radix 2 on adjacent pairs
#pragma unroll
    for (int r = 0; r < 2; r++) {
        int u = lane + 32 * r;
        float2 a = z[2 * u], b = z[2 * u + 1];
        z[2 * u] = cadd(a, b);
        z[2 * u + 1] = csub(a, b);
    }
    __syncwarp();
    // post-twiddle: B[k] sits at pos(k) = 32 (k&3) + 8 ((k>>2)&3) + 2 ((k>>4)&3) + (k>>6)
    float u_[4], v_[4];
#pragma unroll
    for (int r = 0; r < 2; r++) {
        int i = lane + 32 * r;
        int k2 = 127 - i;
        float2 p = T.post1[i];
        float2 B1 = z[32 * (i & 3) + 8 * ((i >> 2) & 3) + 2 * ((i >> 4) & 3) + (i >> 6)];
        float2 B2 = z[32 * (k2 & 3) + 8 * ((k2 >> 2) & 3) + 2 * ((k2 >> 4) & 3) + (k2 >> 6)];
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
        plane[2 * i] = u_[2 * r];
        plane[2 * i + 1] = u_[2 * r + 1];
        plane[128 + 2 * i] = v_[2 * r];
        plane[128 + 2 * i + 1] = v_[2 * r + 1];
    }
    __syncwarp();
}

__device__ void imdct256_warp(const Tables& T, float* plane, int lane)
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
// the kernel
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kGroupThreads * kMaxGroupsPerCta, 1)
a52_decode_kernel(const DecodeParams P)
{
    extern __shared__ __align__(128) uint8_t smem[];
    Tables& T = *reinterpret_cast<Tables*>(smem);
    const int tid = threadIdx.x;
    const int gid = tid / kGroupThreads;          // group within the CTA
    const int gt = tid % kGroupThreads;           // thread within the group
    const int warp = gt >> 5, lane = gt & 31;

    // tables: global -> shared, whole CTA
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(&g_tables);
        uint32_t* dst = reinterpret_cast<uint32_t*>(&T);
        for (int i = tid; i < (int)(sizeof(Tables) / 4); i += blockDim.x) dst[i] = src[i];
    }
    GroupPtrs G = carve(smem + align16((int)sizeof(Tables)) + gid * P.group_bytes, P.fbuf_bytes, P.ndelay);
    GroupCtl* c = G.ctl;
    if (gt == 0) {
        mbar_init(&G.mbar[0], 1);
        mbar_init(&G.mbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // zero persistent decoder state once (liba52 leaves it uninitialised; valid streams never read it)
    for (int i = gt; i < (int)(sizeof(GroupCtl) / 4); i += kGroupThreads)
        reinterpret_cast<uint32_t*>(c)[i] = 0;
    for (int i = gt; i < 7 * 256 * 2 / 4; i += kGroupThreads)
        reinterpret_cast<uint32_t*>(G.exp)[i] = 0;
    __syncthreads();

    uint32_t phase[2] = {0, 0};

    for (;;) {
        // ---- claim a stream ----
        if (gt == 0) c->stream = atomicAdd(P.work_counter, 1);
        group_sync(gid);
        const int s = c->stream;
        if (s >= P.nstreams) break;
        const uint32_t f0 = P.stream_first[s], f1 = P.stream_first[s + 1];

        // carry in
        for (int i = gt; i < P.ndelay * 128; i += kGroupThreads)
            G.delay[i] = P.carry ? P.carry[s].delay[i >> 7][i & 127] : 0.f;
        if (gt == 0) {
            c->dither_index = P.carry ? P.carry[s].dither_index % kDitherPeriod : 0;
            c->per_channel = P.carry ? (P.carry[s].per_channel != 0) : 0;
        }

        // prefetch the first frame
        int cur = 0;
        if (gt == 0 && f0 < f1) {
            uint64_t off = P.frame_off[f0];
            uint64_t a0 = off & ~(uint64_t)15;
            uint32_t nb = stage_bytes(P, f0);
            fence_proxy_async();
            mbar_expect_tx(&G.mbar[0], nb);
            tma_load_1d(G.fbuf[0], P.es + a0, nb, &G.mbar[0]);
        }
        group_sync(gid);

        for (uint32_t f = f0; f < f1; f++, cur ^= 1) {
            const uint64_t off = P.frame_off[f];
            uint32_t* W = G.fbuf[cur];
            // wait for this frame, start the next one
            mbar_wait(&G.mbar[cur], phase[cur]);
            phase[cur] ^= 1;
            if (gt == 0 && f + 1 < f1) {
                uint64_t a1 = P.frame_off[f + 1] & ~(uint64_t)15;
                uint32_t nb = stage_bytes(P, f + 1);
                fence_proxy_async();
                mbar_expect_tx(&G.mbar[cur ^ 1], nb);
                tma_load_1d(G.fbuf[cur ^ 1], P.es + a1, nb, &G.mbar[cur ^ 1]);
            }
            // big-endian bytes -> native words
            for (int i = gt; i < P.fbuf_bytes / 4; i += kGroupThreads)
                W[i] = __byte_perm(W[i], 0, 0x0123);
            group_sync(gid);

            if (gt == 0) {
                uint32_t base_bit = (uint32_t)(off & 15) * 8;
                uint32_t avail = P.fbuf_bytes - 16 - (uint32_t)(off & 15);
                // a frame never extends past the start of the next one / the end of the buffer
                {
                    uint64_t nxt = P.frame_off[f + 1];
                    uint64_t end = (nxt > off) ? nxt : P.es_bytes;
                    if (end - off < avail) avail = (uint32_t)(end - off);
                }
                c->base_bit = base_bit;
                int st = parse_frame_header(c, W, base_bit, P, avail);
                c->frame_ok = (st == 0);
                c->err = st;
                if (P.frame_flags) P.frame_flags[f] = st ? 0 : c->output;
            }
            group_sync(gid);
            int frame_status = c->err;            // 0, 1 (sync) or 2 (frame)
            uint8_t* out_frame = P.pcm + (size_t)f * P.frame_stride;

            if (c->frame_ok) {
                // zero the bits past the frame end so that overruns read zeros
                {
                    uint32_t lim = c->limit_bit;
                    for (uint32_t i = (lim >> 5) + gt; i < (uint32_t)P.fbuf_bytes / 4; i += kGroupThreads) {
                        if (i == (lim >> 5)) {
                            uint32_t keep = lim & 31;
                            W[i] = keep ? (W[i] & (0xffffffffu << (32 - keep))) : 0;
                        } else W[i] = 0;
                    }
                }
                group_sync(gid);
            }

            int blk = 0;
            for (; blk < 6 && c->frame_ok; blk++) {
                // ================= P: side info =================
                if (gt == 0) c->err = parse_block(c, W, P);
                group_sync(gid);
                if (c->err) break;
                const int nfchans = c->nfchans;
                const uint32_t chincpl = c->chincpl;

                // ================= E: exponents =================
                {
                    int bad = 0, k = 0;
                    for (int a = 0; a < 7; a++) {
                        if (!c->expstr[a]) continue;
                        if ((k++ & (kGroupWarps - 1)) != warp) continue;
                        uint8_t* e = G.exp + a * 256;
                        int dst = (a == 6) ? c->cplstrtmant : 1;
                        if (a != 6 && lane == 0) e[0] = c->exp_abs[a];
                        bad |= decode_exponents(W, c->limit_bit, e + dst, c->expstr[a], c->exp_ngrp[a],
                                                c->exp_pos[a], c->exp_abs[a], lane);
                    }
                    if (bad && lane == 0) c->err = 1;
                }
                group_sync(gid);
                if (c->err) break;

                // ================= B: bit allocation =================
                if (c->do_alloc) {
                    int k = 0;
                    for (int a = 0; a < 7; a++) {
                        if (!((c->do_alloc >> a) & 1)) continue;
                        if ((k++ & (kGroupWarps - 1)) != warp) continue;
                        uint8_t* bp = G.bap + a * 256;
                        if (c->zero_alloc) {
                            for (int i = lane; i < 64; i += 32) reinterpret_cast<uint32_t*>(bp)[i] = 0;
                        } else {
                            int16_t* scratch = G.band + warp * 100;
                            bit_allocate_warp(T, c, a, G.exp + a * 256, bp, scratch, scratch + 50, lane);
                        }
                    }
                    group_sync(gid);
                }

                // ================= L: locate =================
                const uint32_t total = c->total_bins;
                const uint32_t K = (total + kGroupThreads - 1) / kGroupThreads;
                const uint32_t my0 = min(gt * K, total), my1 = min(my0 + K, total);
                const uint32_t ncpl_dith = __popc(chincpl & c->dithflag);
                uint32_t n1 = 0, n2 = 0, n4 = 0, nd = 0, fixed = 0;
                {
                    Cursor cu;
                    if (my0 < my1) cursor_seek(cu, c, my0);
                    for (uint32_t i = my0; i < my1; i++) {
                        const Segment sg = c->seg[cu.seg];
                        uint32_t b = G.bap[sg.arr * 256 + cu.bin];
                        n1 += (b == 1); n2 += (b == 2); n4 += (b == 4);
                        if (b == 0) nd += (sg.arr == 6) ? ncpl_dith : sg.dith;
                        else if (b != 1 && b != 2 && b != 4) fixed += T.bap_bits[b];
                        cursor_next(cu, c);
                    }
                }
                // scan 1: counts
                uint32_t pa = n1 | (n2 << 16), pb = n4 | (nd << 16);
                uint32_t ia = warp_incl_scan(pa, lane), ib = warp_incl_scan(pb, lane);
                if (lane == 31) { c->scan_a[warp] = ia; c->scan_b[warp] = ib; }
                group_sync(gid);
                uint32_t ea = ia - pa, eb = ib - pb;
                for (int wv = 0; wv < warp; wv++) { ea += c->scan_a[wv]; eb += c->scan_b[wv]; }
                const uint32_t c1 = ea & 0xffff, c2 = ea >> 16, c4 = eb & 0xffff, cd = eb >> 16;
                // group codes started inside my chunk
                uint32_t s1 = (c1 + n1 + 2) / 3 - (c1 + 2) / 3;
                uint32_t s2 = (c2 + n2 + 2) / 3 - (c2 + 2) / 3;
                uint32_t s4 = (c4 + n4 + 1) / 2 - (c4 + 1) / 2;
                uint32_t mybits = fixed + 5 * s1 + 7 * s2 + 7 * s4;
                uint32_t ic = warp_incl_scan(mybits, lane);
                if (lane == 31) c->scan_c[warp] = ic;
                if (gt == kGroupThreads - 1) c->blk_dither = cd + nd;
                group_sync(gid);
                uint32_t pos = c->bitpos + ic - mybits;
                for (int wv = 0; wv < warp; wv++) pos += c->scan_c[wv];
                if (gt == kGroupThreads - 1) c->mant_bits = pos + mybits - c->bitpos;

                // walk 2: write descriptors, group codes, dither values
                {
                    Cursor cu;
                    uint32_t k1 = c1, k2 = c2, k4 = c4;
                    uint32_t lfsr = 0;
                    if (nd) lfsr = P.dither_seq[(c->dither_index + cd) % kDitherPeriod];
                    const uint32_t limit = c->limit_bit;
                    if (my0 < my1) cursor_seek(cu, c, my0);
                    for (uint32_t i = my0; i < my1; i++) {
                        const Segment sg = c->seg[cu.seg];
                        const uint32_t b = G.bap[sg.arr * 256 + cu.bin];
                        const uint32_t e = G.exp[sg.arr * 256 + cu.bin];
                        uint32_t* slot = reinterpret_cast<uint32_t*>(G.plane) + sg.plane * 256 + cu.bin;
                        if (b == 0) {
                            if (sg.arr == 6) {
                                // one dither value per coupled channel, channel order (parse.c:466-481)
                                for (int ch = 0; ch < nfchans; ch++) {
                                    if (!((chincpl >> ch) & 1)) continue;
                                    uint32_t d = make_desc(e, 0, 0, 0);
                                    if ((c->dithflag >> ch) & 1) {
                                        lfsr = (T.dither_lut[lfsr >> 8] ^ (lfsr << 8)) & 0xffff;
                                        int dv = (3 * (int)(int16_t)lfsr) >> 2;
                                        d = make_desc(e, 0, 0, 0) | (1u << 11) | ((uint32_t)dv << 16);
                                    }
                                    reinterpret_cast<uint32_t*>(G.plane)[ch * 256 + cu.bin] = d;
                                }
                            } else {
                                uint32_t d = make_desc(e, 0, 0, 0);
                                if (sg.dith) {
                                    lfsr = (T.dither_lut[lfsr >> 8] ^ (lfsr << 8)) & 0xffff;
                                    int dv = (3 * (int)(int16_t)lfsr) >> 2;
                                    d |= (1u << 11) | ((uint32_t)dv << 16);
                                }
                                *slot = d;
                            }
                        } else if (b == 1 || b == 2 || b == 4) {
                            uint32_t k, per, wbits, goff;
                            if (b == 1) { k = k1++; per = 3; wbits = 5; goff = 0; }
                            else if (b == 2) { k = k2++; per = 3; wbits = 7; goff = kGrpOff2; }
                            else { k = k4++; per = 2; wbits = 7; goff = kGrpOff4; }
                            uint32_t gi = k / per, dg = k - gi * per;
                            if (dg == 0) {
                                uint32_t code = (pos + wbits <= limit) ? peek_bits(W, pos, wbits) : 0;
                                G.grp[goff + gi] = (uint8_t)code;
                                pos += wbits;
                            }
                            *slot = make_desc(e, b, dg, gi);
                        } else {
                            uint32_t wbits = T.bap_bits[b];
                            *slot = make_desc(e, b, 0, pos);
                            pos += wbits;
                        }
                        cursor_next(cu, c);
                    }
                }
                group_sync(gid);
                if (gt == 0) {
                    c->bitpos += c->mant_bits;
                    c->dither_index = (c->dither_index + c->blk_dither) % kDitherPeriod;
                }

                // ================= U: unpack + dequantise =================
                {
                    const uint32_t limit = c->limit_bit;
                    const int nplanes = c->out_lfe ? 6 : nfchans;
                    const int first_cpl = chincpl ? __ffs(chincpl) - 1 : -1;
                    for (int pl = 0; pl < nplanes; pl++) {
                        if (pl >= nfchans && pl < 5) continue;
                        int own_end, slot_end;
                        float gain;
                        if (pl == 5) { own_end = slot_end = 7; gain = c->gain[5]; }
                        else {
                            own_end = c->endmant[pl];
                            slot_end = ((chincpl >> pl) & 1) ? c->cplendmant : own_end;
                            gain = c->gain[pl];
                        }
                        for (int bin = gt; bin < 256; bin += kGroupThreads) {
                            float* slotf = G.plane + pl * 256 + bin;
                            float val = 0.f;
                            // coupling range of a coupled channel: only the first coupled channel holds
                            // the coupling channel's descriptors; the others hold dither descriptors
                            // (coupling bap 0) or nothing yet (filled by the fan-out below)
                            if (bin >= own_end && bin < slot_end && pl != first_cpl && G.bap[6 * 256 + bin] != 0)
                                continue;
                            if (bin < slot_end) {
                                uint32_t d = __float_as_uint(*slotf);
                                uint32_t e = d & 31, b = (d >> 5) & 15, dg = (d >> 9) & 3;
                                // coupling-range slots hold the coupling channel's value without channel gain
                                float g = (bin < own_end) ? gain : 1.0f;
                                int q = 0;
                                if (b == 0) {
                                    q = (int)d >> 16;              // dither value or 0
                                } else if (b == 1) {
                                    q = T.q1[dg][G.grp[(d >> 12) & 0x7ff] & 31];
                                } else if (b == 2) {
                                    q = T.q2[dg][G.grp[kGrpOff2 + ((d >> 12) & 0x7ff)] & 127];
                                } else if (b == 4) {
                                    q = T.q4[dg][G.grp[kGrpOff4 + ((d >> 12) & 0x7ff)] & 127];
                                } else {
                                    uint32_t wbits = T.bap_bits[b];
                                    uint32_t p = (d >> 12) & 0x7fff;
                                    uint32_t raw = (p + wbits <= limit) ? peek_bits(W, p, wbits) : 0;
                                    if (b == 3) q = T.q3[raw];
                                    else if (b == 5) q = T.q5[raw];
                                    else q = ((int)(raw << (32 - wbits))) >> 16;
                                }
                                val = (float)q * (g * pow2neg(15 + e));
                            }
                            *slotf = val;
                        }
                    }
                }
                group_sync(gid);

                // ================= C: coupling fan-out (parse.c:435-556) =================
                if (chincpl) {
                    int first = __ffs(chincpl) - 1;
                    for (int bin = c->cplstrtmant + gt; bin < c->cplendmant; bin += kGroupThreads) {
                        // coupling band of this bin
                        int sub = (bin - c->cplstrtmant) / 12;
                        int bnd = sub - __popc(c->cplbndstrc & ((1u << sub) - 1));
                        bool zero_bap = (G.bap[6 * 256 + bin] == 0);
                        float cv = G.plane[first * 256 + bin];
                        for (int ch = nfchans - 1; ch >= 0; ch--) {
                            if (!((chincpl >> ch) & 1)) continue;
                            float co = c->cplco[ch][bnd] * c->gain[ch];
                            float src = zero_bap ? G.plane[ch * 256 + bin] : cv;
                            G.plane[ch * 256 + bin] = src * co;
                        }
                    }
                    group_sync(gid);
                }

                // ================= rematrix (parse.c:837-865) =================
                if (c->acmod == 2 && c->rematflg) {
                    int end = min(c->endmant[0], c->endmant[1]);
                    for (int bin = 13 + gt; bin < end; bin += kGroupThreads) {
                        int band = (bin >= 61) ? 3 : (bin >= 37) ? 2 : (bin >= 25) ? 1 : 0;
                        if ((c->rematflg >> band) & 1) {
                            float a = G.plane[bin], b = G.plane[256 + bin];
                            G.plane[bin] = a + b;
                            G.plane[256 + bin] = a - b;
                        }
                    }
                    group_sync(gid);
                }

                // ---- optional dumps ----
                if (P.dbg_exp) {
                    size_t o = ((size_t)f * 6 + blk) * 7 * 256;
                    for (int i = gt; i < 7 * 256; i += kGroupThreads) {
                        P.dbg_exp[o + i] = G.exp[i];
                        P.dbg_bap[o + i] = G.bap[i];
                    }
                }
                if (P.dbg_coef) {
                    size_t o = ((size_t)f * 6 + blk) * 6 * 256;
                    for (int i = gt; i < 6 * 256; i += kGroupThreads) {
                        int pl = i >> 8;
                        bool live = (pl < nfchans) || (pl == 5 && c->out_lfe);
                        P.dbg_coef[o + i] = live ? G.plane[i] : 0.f;
                    }
                }
                if (P.dbg_info && gt == 0) {
                    int32_t* o = P.dbg_info + ((size_t)f * 6 + blk) * 16;
                    for (int i = 0; i < 5; i++) o[i] = c->endmant[i];
                    o[5] = c->cplstrtmant; o[6] = c->cplendmant; o[7] = c->chincpl;
                    o[8] = P.dither_seq[c->dither_index]; o[9] = c->acmod; o[10] = c->lfeon;
                    o[11] = c->output; o[12] = c->blksw | (c->uniform_path << 8) | ((c->clev == 0.f) << 9) | ((c->slev == 0.f) << 10); o[13] = c->ncplbnd; o[14] = c->rematflg;
                    o[15] = c->csnroffst;
                }

                // ================= M: coefficient-domain mix =================
                const MixEntry mx = c_mix[c->acmod * 11 + (c->output & M_MASK)];
                const int nmain = mx.nout;
                const bool uniform = c->uniform_path;
                if (uniform) {
                    for (int bin = gt; bin < 256; bin += kGroupThreads) {
                        float in[5], outv[5];
#pragma unroll
                        for (int ch = 0; ch < 5; ch++) in[ch] = (ch < nfchans) ? G.plane[ch * 256 + bin] : 0.f;
#pragma unroll
                        for (int o = 0; o < 5; o++) {
                            float acc = 0.f;
                            if (o < nmain) {
#pragma unroll
                                for (int ch = 0; ch < 5; ch++) {
                                    if ((mx.pos[o] >> ch) & 1) acc += in[ch];
                                    else if ((mx.neg[o] >> ch) & 1) acc -= in[ch];
                                }
                            }
                            outv[o] = acc;
                        }
#pragma unroll
                        for (int o = 0; o < 5; o++)
                            if (o < nmain) G.plane[o * 256 + bin] = outv[o];
                    }
                    group_sync(gid);
                }

                // ================= T: transforms =================
                {
                    const int ntr = uniform ? nmain : nfchans;
                    const int njobs = ntr + (c->out_lfe ? 1 : 0);
                    for (int j = warp; j < njobs; j += kGroupWarps) {
                        int pl = (j < ntr) ? j : 5;
                        bool shortblk = (pl < 5) && ((c->blksw >> (uniform ? 0 : pl)) & 1);
                        if (shortblk) imdct256_warp(T, G.plane + pl * 256, lane);
                        else imdct512_warp(T, G.plane + pl * 256, lane);
                    }
                }
                group_sync(gid);

                // ================= O: window + overlap-add (+ time-domain mix) + store ========
                // The delay planes follow liba52's state machine (parse.c:881-937): after a block that
                // mixed coefficients they hold the downmixed tail (planes 0..nout-1); after a block that
                // transformed every coded channel they hold per-channel tails (planes 0..nfchans-1).
                // Switching representation = a52_downmix / a52_upmix on the delay (downmix.c:480-685);
                // a channel whose gain is zero keeps its tail untouched and unheard (parse.c:897-909).
                {
                    const int lfe_on = c->out_lfe;
                    const int nout = nmain + lfe_on;
                    const float bias = P.bias;
                    float* D = G.delay;
                    const int p = gt;                              // kGroupThreads == 128 positions
                    const float w0 = T.window[p], w1 = T.window[255 - p];
                    float y0[6], y1[6];                            // by output channel (0 = LFE when present)
#pragma unroll
                    for (int o = 0; o < 6; o++) { y0[o] = 0.f; y1[o] = 0.f; }
                    if (lfe_on) {
                        float U = G.plane[5 * 256 + p], V = G.plane[5 * 256 + 128 + p], Dv = D[5 * 128 + p];
                        y0[0] = Dv * w1 - U * w0;
                        y1[0] = Dv * w0 + U * w1;
                        D[5 * 128 + p] = V;
                    }
                    if (uniform) {
                        if (c->per_channel) {
                            float d[5], m[5];
#pragma unroll
                            for (int ch = 0; ch < 5; ch++)
                                d[ch] = (ch < nfchans && c->gain[ch] != 0.f) ? D[ch * 128 + p] : 0.f;
#pragma unroll
                            for (int o = 0; o < 5; o++) {
                                float acc = 0.f;
#pragma unroll
                                for (int ch = 0; ch < 5; ch++) {
                                    if ((mx.pos[o] >> ch) & 1) acc += d[ch];
                                    else if ((mx.neg[o] >> ch) & 1) acc -= d[ch];
                                }
                                m[o] = acc;
                            }
#pragma unroll
                            for (int o = 0; o < 5; o++)
                                if (o < nmain) D[o * 128 + p] = m[o];
                        }
#pragma unroll
                        for (int o = 0; o < 5; o++) {
                            if (o < nmain) {
                                float U = G.plane[o * 256 + p], V = G.plane[o * 256 + 128 + p], Dv = D[o * 128 + p];
                                float a = Dv * w1 - U * w0, b = Dv * w0 + U * w1;
                                D[o * 128 + p] = V;
                                if (lfe_on) { y0[o + 1 > 5 ? 5 : o + 1] = a; y1[o + 1 > 5 ? 5 : o + 1] = b; }
                                else { y0[o] = a; y1[o] = b; }
                            }
                        }
                    } else {
                        if (!c->per_channel) {
                            float m[5];
#pragma unroll
                            for (int o = 0; o < 5; o++) m[o] = D[o * 128 + p];
#pragma unroll
                            for (int ch = 0; ch < 5; ch++) {
                                if (ch < nfchans) {
                                    float v = 0.f;
#pragma unroll
                                    for (int o = 0; o < 5; o++)
                                        if (mx.up[ch] == o) v = m[o];
                                    D[ch * 128 + p] = v;
                                }
                            }
                        }
#pragma unroll
                        for (int ch = 0; ch < 5; ch++) {
                            if (ch < nfchans && c->gain[ch] != 0.f) {
                                float U = G.plane[ch * 256 + p], V = G.plane[ch * 256 + 128 + p], Dv = D[ch * 128 + p];
                                float a = Dv * w1 - U * w0, b = Dv * w0 + U * w1;
                                D[ch * 128 + p] = V;
#pragma unroll
                                for (int o = 0; o < 5; o++) {
                                    const int oo = lfe_on ? (o + 1 > 5 ? 5 : o + 1) : o;
                                    if ((mx.pos[o] >> ch) & 1) { y0[oo] += a; y1[oo] += b; }
                                    else if ((mx.neg[o] >> ch) & 1) { y0[oo] -= a; y1[oo] -= b; }
                                }
                            }
                        }
                    }
#pragma unroll
                    for (int oc = 0; oc < 6; oc++) {
                        if (oc < nout) {
                            float a = y0[oc] + bias, b = y1[oc] + bias;
                            if (P.out_fmt == 0) {
                                float* dst = reinterpret_cast<float*>(out_frame) + ((size_t)blk * nout + oc) * 256;
                                dst[p] = a;
                                dst[255 - p] = b;
                            } else if (P.out_fmt == 1) {
                                float* dst = reinterpret_cast<float*>(out_frame) + (size_t)blk * 256 * nout;
                                dst[p * nout + oc] = a;
                                dst[(255 - p) * nout + oc] = b;
                            } else {
                                int16_t* dst = reinterpret_cast<int16_t*>(out_frame) + (size_t)blk * 256 * nout;
                                int ia = __float2int_rn(y0[oc] * 32768.f), ib = __float2int_rn(y1[oc] * 32768.f);
                                ia = min(max(ia, -32768), 32767);
                                ib = min(max(ib, -32768), 32767);
                                dst[p * nout + oc] = (int16_t)ia;
                                dst[(255 - p) * nout + oc] = (int16_t)ib;
                            }
                        }
                    }
                }
                group_sync(gid);
                if (gt == 0) c->per_channel = uniform ? 0 : 1;
            }   // blocks

            if (c->frame_ok && blk < 6) frame_status = 16 + blk;     // A52_ST_BAD_BLOCK + block
            if (frame_status) {
                // silence for everything not produced
                int nout = c->frame_ok ? (c->nout + c->out_lfe) : P.nout_req;
                int ssz = (P.out_fmt == 2) ? 2 : 4;
                size_t from = (size_t)blk * 256 * nout * ssz;
                size_t to = (size_t)6 * 256 * nout * ssz;
                for (size_t i = from + gt * 4; i < to; i += kGroupThreads * 4)
                    *reinterpret_cast<uint32_t*>(out_frame + i) = 0;
            }
            if (gt == 0 && P.status) P.status[f] = frame_status;
            group_sync(gid);
        }   // frames

        // carry out
        if (P.carry) {
            for (int i = gt; i < P.ndelay * 128; i += kGroupThreads)
                P.carry[s].delay[i >> 7][i & 127] = G.delay[i];
            if (gt == 0) {
                P.carry[s].dither_index = c->dither_index;
                P.carry[s].per_channel = c->per_channel;
            }
        }
        group_sync(gid);
    }
}

}  // namespace a52

#include "a52_host.inl"
