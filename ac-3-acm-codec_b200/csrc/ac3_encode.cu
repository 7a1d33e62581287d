// ac3_encode.cu - batched AC-3 encode for NVIDIA B200 (sm_100a).
//
// Bit-exact re-implementation of the reference's integer encoder
// (reference: src/ac3enc/ac3enc.cpp:1640-1763) as one persistent kernel.  A CTA of
// six warps owns one PCM stream at a time and walks its frames in order (the
// reference carries the previous 256 samples per channel and the warm start of the
// SNR-offset search across frames, ac3enc.cpp:55, 921, 969); warp w is coded channel w
// for everything that is per channel.  Everything between the PCM load and the frame
// store lives in shared memory:
//
//   E1  window, block-floating normalisation, 512-point fixed-point MDCT (128-point
//       radix-2 FFT, 16-bit storage, halving butterflies), exponents   (:1673-1722, 485-603)
//   E2  exponent strategy, run minima, group minima, +-2 delta constraint as two
//       min-plus warp scans                                            (:617-761, 1725-1749)
//   E3  masking curve once per exponent set (warp scans: no serial lane), then the SNR-offset
//       search on per-set class counts: a pass evaluates the next three likely probes by bap
//       address (packed counters, no re-derivation of the curve) and every warp replays the
//       reference's state machine on the results                      (:220-421, 764-975)
//   E4  side information as 64-bit runs; grouped exponents by channel warps; mantissas by
//       BLOCK warps (a lane owns eight consecutive bins, positions from one packed scan,
//       group members in per-block rings, ungrouped fields in 64-bit runs); both CRCs over
//       all six warps with per-thread multipliers                     (:1113-1638)
//
// Frames are byte-identical to the reference's (tests/test_encoder_gpu.py).
#include <cuda_runtime.h>
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <new>
#include <vector>

#include "ac3_tables.h"
#include "../../include/ac3enc.h"
#include "../../include/ac3enc_batch.h"

namespace ac3e {

constexpr int kWarps = 6;
constexpr int kThreads = kWarps * 32;
constexpr int kFrameWords = 968;         // staged output frame, 32-bit words (3840 bytes max + slack)

struct EncTables {
    uint32_t taba[16];                   // per bap: plain field bits | 3-level << 8 | 5-level << 16 | 11-level << 24 (counters)
    uint2    tabq[16];                   // per bap: x = levels (0 = asymmetric) | quantiser bits << 8 | field width << 16 | class << 24,
                                         //          y = what the bap adds to the packed class counters (1 << 10 (class - 1), or 0)
    uint32_t tabc[64];                   // taba[baptab[address]]: a search probe counts without looking at the bap
    uint8_t  masktab[256];               // (8-byte aligned: read eight bins at a time)
    int16_t  window[256];
    int16_t  costab[64], sintab[64], xcos1[128], xsin1[128];
    uint16_t crc_table[256];
    uint16_t hth[150];
    uint8_t  rev[128];
    uint8_t  latab[256];
    uint8_t  baptab[64];
    uint8_t  bndtab[52];
    uint8_t  plain_bits[16];             // field width of the ungrouped baps, 0 for bap 0, 1, 2, 4
    uint8_t  width[16];                  // field / group-code width per bap
    uint8_t  pad_[8];
};

__device__ EncTables g_enc_tables;

struct EncCarry {                        // == ac3_stream_carry_t
    int16_t last_samples[6][256];
    int32_t csnroffst;
    int32_t started;
    int32_t reserved[2];
};

struct EncParams {
    const int16_t* pcm;
    uint8_t*  out;
    int32_t*  status;
    EncCarry* carry;
    int*      work_counter;
    int*      slice_done;                // [nstreams] slices finished per stream (work units are slices of streams)
    int       slice_frames, nslices;
    int nstreams, nframes;
    int nch_all, nch, lfe, acmod, fscod, halfrate, bsid, frmsizecod, frame_words;
    uint32_t crc_inv;                    // x^-(16 fs58 - 16) mod poly (ac3enc.cpp:1627)
    uint8_t chmap[8];
    // optional dumps
    int32_t* dbg_coef;
    int8_t*  dbg_shift;
    uint8_t* dbg_strategy;
    uint8_t* dbg_enc;
    uint8_t* dbg_bap;
    int32_t* dbg_snr;
};

static_assert(offsetof(EncTables, masktab) % 8 == 0, "masktab is read as uint2");

struct EncShared {
    int32_t  coef[6][6][256];            // MDCT coefficients; a (block, channel) slot first holds its 512 input samples
    uint8_t  expo[6][6][256];            // raw exponents, later the baps
    uint8_t  enc[6][6][256];             // exponents as the decoder will see them (run heads)
    int16_t  last[6][256];               // previous 256 samples per coded channel
    union {
        // E1..E3: FFT scratch, masking curves before the snr offset (run heads)
        // (pcnt: the class counts of a search pass - three snr offsets at a time -, two buffers used in turn)
        struct { uint32_t z[6][128]; int16_t mask[6][6][50]; int pcnt[2][3][6][6][4]; } e1;
        // E4: the frame being packed; per block (= warp) the quantised members of the 3- / 5- / 11-level groups by
        // occurrence number and the bit positions of the group codes, as rings (a channel adds at most 223 members
        // and 112 groups to a class, and at most one group of a class stays open across a channel boundary)
        struct { uint32_t frame[kFrameWords]; uint8_t ring_v[6][3][256]; uint16_t ring_p[6][3][128]; } e4;
    } u;
    int      cnt[6][6][4];               // per exponent set: 3-, 5-, 11-level mantissas, bits of plain fields
    uint8_t  strategy[6][6];
    uint8_t  head[6][6];                 // block holding the exponent set a (block, channel) uses
    int8_t   exp_shift[6][6];
    uint32_t exp_pos[6][6];              // bit position of a channel's exponent section in a block
    uint32_t mant_pos[6];                // bit position of a block's first mantissa
    int      exp_bits[6];                // per channel: bits of its exponent sections (ac3enc.cpp:760)
    int      frame_bits;                 // everything but mantissas
    int      cs, fs, failed;             // result of the search (and the warm start of the next frame's)
    uint32_t crcp[6];                    // the warps' parts of the two CRCs
};

// exponent groups of a coded range for strategy 1 / 2 / 3 (one, two or four bins per exponent; :684-700):
// (ncoef + 3 gs - 4) / (3 gs) for the two ranges there are - 223 bins (74, 37, 19) and the LFE's 7 (2, 1, 1)
__device__ __forceinline__ int exp_groups(int strategy, bool is_lfe)
{
    return is_lfe ? (strategy == 1 ? 2 : 1) : (strategy == 1 ? 74 : strategy == 2 ? 37 : 19);
}

__device__ __forceinline__ int ilog2(uint32_t v) { return v ? 31 - __clz(v) : 0; }

// two 16-bit halves -> one word (low halves of re and im), one PRMT
__device__ __forceinline__ uint32_t pack16(int re, int im) { return __byte_perm((uint32_t)re, (uint32_t)im, 0x5410); }

__device__ __forceinline__ void put_bits_atomic(uint32_t* frame, uint32_t pos, uint32_t n, uint32_t v)
{
    // n in 1..16, v < 2^n; big-endian bit order: bit 0 of the frame = msb of word 0
    if (pos + n > kFrameWords * 32) return;
    uint32_t wi = pos >> 5, sh = pos & 31;
    uint64_t x = (uint64_t)v << (64 - n - sh);
    uint32_t hi = (uint32_t)(x >> 32), lo = (uint32_t)x;
    if (hi) atomicOr(&frame[wi], hi);
    if (lo) atomicOr(&frame[wi + 1], lo);
}

// A run of at most 64 bits of side information collected in a register and written in one go.
struct BitRun {
    uint64_t acc;
    uint32_t n;
    __device__ __forceinline__ void put(uint32_t k, uint32_t v) { acc = (acc << k) | v; n += k; }
    // writes the run at bit `pos` of the frame (big-endian bit order) and returns the position after it
    __device__ __forceinline__ uint32_t flush(uint32_t* frame, uint32_t pos) const
    {
        if (n == 0 || pos + n > kFrameWords * 32) return pos + n;
        const uint64_t x = acc << (64 - n);
        const uint32_t hi = (uint32_t)(x >> 32), lo = (uint32_t)x, wi = pos >> 5, sh = pos & 31;
        const uint32_t w0 = hi >> sh, w1 = __funnelshift_r(lo, hi, sh), w2 = __funnelshift_r(0u, lo, sh);
        if (w0) atomicOr(&frame[wi], w0);
        if (w1) atomicOr(&frame[wi + 1], w1);
        if (w2) atomicOr(&frame[wi + 2], w2);
        return pos + n;
    }
};

__device__ __forceinline__ int sym_quant(int c, int e, int levels)      // ac3enc.cpp:1150-1166
{
    int v;
    if (c >= 0) { v = (levels * (c << e)) >> 24; v = (v + 1) >> 1; v = (levels >> 1) + v; }
    else { v = (levels * ((-c) << e)) >> 24; v = (v + 1) >> 1; v = (levels >> 1) - v; }
    return v;
}

__device__ __forceinline__ int asym_quant(int c, int e, int qbits)       // ac3enc.cpp:1169-1190
{
    int lshift = e + qbits - 24, v;
    v = lshift >= 0 ? c << lshift : c >> (-lshift);
    v = (v + 1) >> 1;
    int m = 1 << (qbits - 1);
    if (v >= m) v = m - 1;
    return v & ((1 << qbits) - 1);
}

__device__ __noinline__ uint32_t mul_poly(uint32_t a, uint32_t b)      // ac3enc.cpp:1513-1524, poly 0x18005
{
    // (a loop on purpose, and out of line: it runs a dozen times per frame and must not cost instruction-cache space)
    uint32_t c = 0;
#pragma unroll 1
    for (int i = 0; i < 16; i++) {
        if (a & 1) c ^= b;
        a >>= 1;
        b <<= 1;
        if (b & 0x10000) b ^= 0x18005;
    }
    return c;
}

// CRC-16 (poly 0x8005, init 0) over a byte range of the frame by 96 threads: every thread runs the table CRC over
// an equal chunk (chunks are aligned to the END of the range: leading zero bytes do not change a zero-initialised
// CRC) and multiplies it by x^(8 * bytes after its chunk) - a constant of the thread, worked out once per launch -
// so that the range's CRC is the XOR of the threads' values (the CRC is linear over GF(2)).
__device__ __forceinline__ uint32_t crc_chunk(const EncTables& T, const uint32_t* frame, int b0, int start, int L)
{
    uint32_t crc = 0;
#pragma unroll 1
    for (int k = 0; k < L; k++) {
        const int idx = start + k;
        const uint32_t byte = (idx >= b0) ? ((frame[idx >> 2] >> (24 - 8 * (idx & 3))) & 0xff) : 0u;
        crc = (T.crc_table[byte ^ (crc >> 8)] ^ (crc << 8)) & 0xffff;
    }
    return crc;
}

// ---------------------------------------------------------------------------
// E1: one (block, channel): window, normalise, MDCT-512, exponents.  One warp.
// ---------------------------------------------------------------------------
// `smp[r]` = the block's new sample lane + 32 r of this channel (the caller fetches them a block ahead)
__device__ __forceinline__ void e1_transform(EncShared& S, const EncTables& T, int blk, int ch, int lane,
                                             const int (&smp_in)[8])
{
    int16_t* in = reinterpret_cast<int16_t*>(S.coef[blk][ch]);       // 512 samples live in the coefficient slot
    uint32_t* z = S.u.e1.z[ch];
    // previous 256 samples | new 256 samples, windowed (ac3enc.cpp:1673-1693)
    uint32_t amax = 0;
#pragma unroll
    for (int r = 0; r < 8; r++) {
        const int j = lane + 32 * r;
        const int old = S.last[ch][j];
        const int cur = smp_in[r];
        S.last[ch][j] = (int16_t)cur;
        const int a = (int16_t)((old * T.window[j]) >> 15);
        const int b = (int16_t)((cur * T.window[255 - j]) >> 15);
        in[j] = (int16_t)a;
        in[256 + j] = (int16_t)b;
        amax |= (uint32_t)abs(a) | (uint32_t)abs(b);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) amax |= __shfl_xor_sync(0xffffffffu, amax, o);
    int sh = 14 - ilog2(amax);                                          // :1697-1700
    if (sh < 0) sh = 0;
    if (lane == 0) S.exp_shift[blk][ch] = (int8_t)(sh - 9);
    __syncwarp();
    // pre-rotation (:576-589) on the shifted samples; rot[k] = k < 128 ? -in[k + 384] : in[k - 128]
#pragma unroll
    for (int r = 0; r < 4; r++) {
        const int i = lane + 32 * r;
        auto smp = [&](int k) -> int { return (int)(int16_t)(in[k] << sh); };
        auto rot = [&](int k) -> int { return k < 128 ? (int)(int16_t)(-smp(k + 384)) : smp(k - 128); };
        const int re = (rot(2 * i) - rot(511 - 2 * i)) >> 1;
        const int im = -(rot(256 + 2 * i) - rot(255 - 2 * i)) >> 1;
        const int bre = -T.xcos1[i], bim = T.xsin1[i];
        const int xr = (re * bre - im * bim) >> 15, xi = (re * bim + bre * im) >> 15;
        z[T.rev[i]] = pack16(xr, xi);
    }
    __syncwarp();
    // 128-point FFT (:485-568): the reference's seven halving radix-2 passes, butterfly for butterfly (same products,
    // same shifts, results cut to 16 bits as its IComplex stores cut them), but two passes at a time in registers:
    // a lane holds the four points of a 4-point group, so the scratch is crossed four times instead of seven.
    struct CI { int re, im; };
    auto ld = [&](int i) { const uint32_t p = z[i]; return CI{(int)(int16_t)(p & 0xffff), (int)(int16_t)(p >> 16)}; };
    auto stz = [&](int i, const CI& v) { z[i] = pack16(v.re, v.im); };
    // one butterfly; t == 0: no product (the reference's first butterfly of a group), quarter: multiply by -j
    auto bf = [&](CI& pp, CI& q, int l, bool plain, bool quarter) {
        int ax, ay;
        if (quarter) { ax = q.im; ay = -q.re; }
        else {
            const int c = T.costab[l], sn = -T.sintab[l];
            ax = (c * q.re - sn * q.im) >> 15;
            ay = (c * q.im + q.re * sn) >> 15;
            if (plain) { ax = q.re; ay = q.im; }
        }
        const int r0 = (pp.re + ax) >> 1, i0 = (pp.im + ay) >> 1, r1 = (pp.re - ax) >> 1, i1 = (pp.im - ay) >> 1;
        pp.re = (int16_t)r0; pp.im = (int16_t)i0; q.re = (int16_t)r1; q.im = (int16_t)i1;
    };
    auto bf_plain = [&](CI& pp, CI& q) {
        const int r0 = (pp.re + q.re) >> 1, i0 = (pp.im + q.im) >> 1, r1 = (pp.re - q.re) >> 1, i1 = (pp.im - q.im) >> 1;
        pp.re = (int16_t)r0; pp.im = (int16_t)i0; q.re = (int16_t)r1; q.im = (int16_t)i1;
    };
    CI y0, y1, y2, y3;
    {   // passes 0, 1: points 4 lane .. 4 lane + 3
        const uint4 v = reinterpret_cast<const uint4*>(z)[lane];
        y0 = CI{(int)(int16_t)(v.x & 0xffff), (int)(int16_t)(v.x >> 16)};
        y1 = CI{(int)(int16_t)(v.y & 0xffff), (int)(int16_t)(v.y >> 16)};
        y2 = CI{(int)(int16_t)(v.z & 0xffff), (int)(int16_t)(v.z >> 16)};
        y3 = CI{(int)(int16_t)(v.w & 0xffff), (int)(int16_t)(v.w >> 16)};
        bf_plain(y0, y1); bf_plain(y2, y3);
        bf_plain(y0, y2); bf(y1, y3, 0, false, true);
        reinterpret_cast<uint4*>(z)[lane] = make_uint4(pack16(y0.re, y0.im), pack16(y1.re, y1.im), pack16(y2.re, y2.im), pack16(y3.re, y3.im));
    }
    __syncwarp();
    {   // passes 2, 3: points base + 4 j
        const int t = lane & 3, base = (lane >> 2) * 16 + t;
        y0 = ld(base); y1 = ld(base + 4); y2 = ld(base + 8); y3 = ld(base + 12);
        bf(y0, y1, 16 * t, t == 0, false); bf(y2, y3, 16 * t, t == 0, false);
        bf(y0, y2, 8 * t, t == 0, false);  bf(y1, y3, 8 * (t + 4), false, false);
        stz(base, y0); stz(base + 4, y1); stz(base + 8, y2); stz(base + 12, y3);
    }
    __syncwarp();
    {   // passes 4, 5: points base + 16 j
        const int t = lane & 15, base = (lane >> 4) * 64 + t;
        y0 = ld(base); y1 = ld(base + 16); y2 = ld(base + 32); y3 = ld(base + 48);
        bf(y0, y1, 4 * t, t == 0, false); bf(y2, y3, 4 * t, t == 0, false);
        bf(y0, y2, 2 * t, t == 0, false); bf(y1, y3, 2 * (t + 16), false, false);
        stz(base, y0); stz(base + 16, y1); stz(base + 32, y2); stz(base + 48, y3);
    }
    __syncwarp();
    {   // pass 6: (lane, lane + 64) and (lane + 32, lane + 96); the results stay in registers for the post-rotation
        y0 = ld(lane); y2 = ld(lane + 64); y1 = ld(lane + 32); y3 = ld(lane + 96);
        bf(y0, y2, lane, lane == 0, false);
        bf(y1, y3, lane + 32, false, false);
    }
    // post-rotation (:594-602) into the coefficient slot (the samples are no longer needed)
    int32_t* out = S.coef[blk][ch];
    int o0[4], o1[4];
    {
        const CI* yy[4] = {&y0, &y1, &y2, &y3};
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const int i = lane + 32 * r;
            const int re = yy[r]->re, im = yy[r]->im;
            o1[r] = (re * T.xsin1[i] - im * T.xcos1[i]) >> 15;               // re1 -> out[255 - 2i]
            o0[r] = (re * T.xcos1[i] + T.xsin1[i] * im) >> 15;               // im1 -> out[2i]
        }
    }
    __syncwarp();
#pragma unroll
    for (int r = 0; r < 4; r++) {
        const int i = lane + 32 * r;
        out[2 * i] = o0[r];
        out[255 - 2 * i] = o1[r];
    }
    __syncwarp();
    // exponents (:1707-1722)
    const int es = sh - 9;
#pragma unroll
    for (int r = 0; r < 8; r++) {
        const int j = lane + 32 * r;
        const int a = abs(out[j]);
        // 23 - ilog2(a) + es = clz(a) - 8 + es; zero and everything at or past 24 read 24, and the latter is zeroed
        int e = a ? __clz(a) - 8 + es : 24;
        if (a && e >= 24) out[j] = 0;
        S.expo[blk][ch][j] = (uint8_t)min(e, 24);
    }
}

// ---------------------------------------------------------------------------
// E2: strategies + exponent sets of one channel.  One warp.  Returns the exponent bits.
// ---------------------------------------------------------------------------
__device__ int e2_exponents(EncShared& S, const EncParams& P, int ch, int lane, uint32_t& sets)
{
    const bool is_lfe = P.lfe && ch == 5;
    const int ncoef = is_lfe ? 7 : 223;
    // new exponents when the L1 distance to the previous block exceeds 1000 over all 256 bins (:617-640)
    uint32_t newmask = 1;
    {
        // (a lane takes eight consecutive bins: sums of absolute byte differences, four bytes per instruction)
        uint2 prev = *reinterpret_cast<const uint2*>(S.expo[0][ch] + 8 * lane);
        for (int blk = 1; blk < 6; blk++) {
            const uint2 cur = *reinterpret_cast<const uint2*>(S.expo[blk][ch] + 8 * lane);
            const uint32_t d = __reduce_add_sync(0xffffffffu, __vsadu4(cur.x, prev.x) + __vsadu4(cur.y, prev.y));
            if (d > 1000) newmask |= 1u << blk;
            prev = cur;
        }
    }
    sets = newmask;                                                      // the blocks that start an exponent set
    int bits = 0;
    for (int i = 0; i < 6;) {
        int j = i + 1;
        while (j < 6 && !((newmask >> j) & 1)) j++;
        // run [i, j): strategy by run length (:645-668); the LFE only knows new / reuse
        const int strat = is_lfe ? 1 : (j - i == 1) ? 3 : (j - i <= 3) ? 2 : 1;
        if (lane == 0) {
            S.strategy[i][ch] = (uint8_t)strat;
            for (int k = i; k < j; k++) {
                if (k > i) S.strategy[k][ch] = 0;
                S.head[k][ch] = (uint8_t)i;
            }
        }
        // minimum over the run (:1736-1739), coded bins only
        for (int k = lane; k < ncoef; k += 32) {
            int m = S.expo[i][ch][k];
            for (int b = i + 1; b < j; b++) m = min(m, (int)S.expo[b][ch][k]);
            S.expo[i][ch][k] = (uint8_t)m;
        }
        __syncwarp();
        // group minima, dc <= 15 (:684-726); lane owns values 7 lane .. 7 lane + 6 of e1[0 .. ng]
        const int gs = strat == 1 ? 1 : strat == 2 ? 2 : 4;
        const int ng = 3 * exp_groups(strat, is_lfe);
        const uint8_t* ex = S.expo[i][ch];
        int e1[7];
#pragma unroll
        for (int r = 0; r < 7; r++) {
            const int g = 7 * lane + r;
            int m = 64;
            if (g == 0) m = min((int)ex[0], 15);
            else if (g <= ng) {
                const int k = 1 + (g - 1) * gs;
                m = ex[k];
                for (int q = 1; q < gs; q++) m = min(m, (int)ex[k + q]);
            }
            e1[r] = m;
        }
        // |delta| <= 2 (:732-748): largest sequence below e1 = forward then backward min-plus sweep
        {
            int run = 1 << 20;                                           // min over my values of e - 2 g
#pragma unroll
            for (int r = 0; r < 7; r++) run = min(run, e1[r] - 2 * (7 * lane + r));
            int incl = run;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl = min(incl, t);
            }
            int carry = __shfl_up_sync(0xffffffffu, incl, 1);
            if (lane == 0) carry = 1 << 20;
#pragma unroll
            for (int r = 0; r < 7; r++) {
                const int g = 7 * lane + r;
                carry = min(carry, e1[r] - 2 * g);
                e1[r] = carry + 2 * g;
            }
            run = 1 << 20;                                               // min over my values of e + 2 g
#pragma unroll
            for (int r = 0; r < 7; r++)
                if (7 * lane + r <= ng) run = min(run, e1[r] + 2 * (7 * lane + r));
            incl = run;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int t = __shfl_down_sync(0xffffffffu, incl, o);
                if (lane + o < 32) incl = min(incl, t);
            }
            carry = __shfl_down_sync(0xffffffffu, incl, 1);
            if (lane == 31) carry = 1 << 20;
#pragma unroll
            for (int r = 6; r >= 0; r--) {
                const int g = 7 * lane + r;
                if (g <= ng) {
                    carry = min(carry, e1[r] + 2 * g);
                    e1[r] = carry - 2 * g;
                }
            }
        }
        // what the decoder will see (:750-758)
        uint8_t* en = S.enc[i][ch];
#pragma unroll
        for (int r = 0; r < 7; r++) {
            const int g = 7 * lane + r;
            if (g == 0) en[0] = (uint8_t)e1[r];
            else if (g <= ng) {
                const int k = 1 + (g - 1) * gs;
                for (int q = 0; q < gs; q++) en[k + q] = (uint8_t)e1[r];
            }
        }
        bits += 4 + (ng / 3) * 7;
        __syncwarp();
        i = j;
    }
    return bits;
}

// ---------------------------------------------------------------------------
// E3a: masking curve of one exponent set (everything of ac3enc.cpp:220-383 that does not
// depend on the snr offset).  Encoder parameters are fixed (:861-869).  One warp.
// ---------------------------------------------------------------------------
__device__ void e3_mask(EncShared& S, const EncTables& T, const EncParams& P, int blk, int ch, int lane, int16_t* psd)
{
    const bool is_lfe = P.lfe && ch == 5;
    const int end = is_lfe ? 7 : 223;
    const uint8_t* ex = S.enc[blk][ch];
    const int sdecay = 0x13 >> P.halfrate, fdecay = 0x53 >> P.halfrate, sgain = 0x4d8, dbknee = 0x900, fgain = 0x280;
    const int bndend = T.masktab[end - 1] + 1;
    for (int band = lane; band < bndend; band += 32) {
        const int b0 = T.bndtab[band], b1 = min((int)T.bndtab[band + 1], end);
        int v = 3072 - (ex[b0] << 7);
        for (int bin = b0 + 1; bin < b1; bin++) {
            const int p = 3072 - (ex[bin] << 7);
            const int adr = min(abs(v - p) >> 1, 255);
            v = max(v, p) + T.latab[adr];
        }
        psd[band] = (int16_t)v;
    }
    __syncwarp();
    int16_t* mk = S.u.e1.mask[blk][ch];
    int fast21 = 0, slow21 = 0;                                          // the recurrences' state after band 21
    if (bndend > 22) {
        // Bands 0..21 of a full-bandwidth channel, lane = band (:262-330 of the reference, which walks them one by one).
        // The low-frequency compensation is a chain of steps x -> max(x - a, c) - "reset to 384 / 320" (a = inf), "down by
        // 64 / 128, not below 0", "hold" (a = 0, c = -inf) - chosen by the band's psd and the next; such steps compose
        // ((a1, c1) then (a2, c2) = (a1 + a2, max(c1 - a2, c2))), so the chain is a warp scan.  fast and slow restart at
        // band begin - 1 (the first band in 2..6 whose psd does not fall, else 6) and decay from there: prefix maxima of
        // psd + band * decay, as for the bands above 21.
        constexpr int kInf = 1 << 20;
        const int b = lane;
        const bool on = b < 22;
        const int p0 = psd[on ? b : 0], p1 = psd[on ? b + 1 : 0];
        int a, c;
        if (b < 20) {
            if (p0 + 256 == p1) { a = kInf; c = b < 7 ? 384 : 320; }
            else if (p0 > p1) { a = 64; c = 0; }
            else { a = 0; c = -kInf; }
        } else { a = 128; c = 0; }
        const uint32_t rise = __ballot_sync(0xffffffffu, b >= 2 && b <= 6 && p0 <= p1);
        const int begin = rise ? __ffs(rise) : 7;                        // (lowest such band) + 1
        const bool live = on && b >= begin - 1;
        int pf = live ? p0 - fgain + b * fdecay : -kInf, ps = live ? p0 - sgain + b * sdecay : -kInf;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int a1 = __shfl_up_sync(0xffffffffu, a, o), c1 = __shfl_up_sync(0xffffffffu, c, o);
            const int tf = __shfl_up_sync(0xffffffffu, pf, o), ts = __shfl_up_sync(0xffffffffu, ps, o);
            if (lane >= o) {
                c = max(c1 - a, c);
                a = min(a1 + a, kInf);
                pf = max(pf, tf);
                ps = max(ps, ts);
            }
        }
        const int lowcomp = max(-a, c);
        const int f = pf - b * fdecay, sl = ps - b * sdecay;
        if (on) mk[b] = (int16_t)(b < begin ? p0 - fgain - lowcomp : max(f - lowcomp, sl));
        fast21 = __shfl_sync(0xffffffffu, f, 21);
        slow21 = __shfl_sync(0xffffffffu, sl, 21);
    } else if (lane == 0) {
        // the LFE channel's seven bands, one by one
        auto lc1 = [](int a, int b0, int b1) { return (b0 + 256 == b1) ? 384 : (b0 > b1) ? max(a - 64, 0) : a; };
        int lowcomp = 0, fast = 0, slow = 0, begin = 7, bin;
        lowcomp = lc1(lowcomp, psd[0], psd[1]);
        mk[0] = (int16_t)(psd[0] - fgain - lowcomp);
        lowcomp = lc1(lowcomp, psd[1], psd[2]);
        mk[1] = (int16_t)(psd[1] - fgain - lowcomp);
        for (bin = 2; bin < 7; bin++) {
            const bool last_lfe = is_lfe && bin == 6;
            if (!last_lfe) lowcomp = lc1(lowcomp, psd[bin], psd[bin + 1]);
            fast = psd[bin] - fgain;
            slow = psd[bin] - sgain;
            mk[bin] = (int16_t)(fast - lowcomp);
            if (!last_lfe && psd[bin] <= psd[bin + 1]) { begin = bin + 1; break; }
        }
        for (bin = begin; bin < bndend; bin++) {
            if (!(is_lfe && bin == 6)) {
                const int b0 = psd[bin], b1 = psd[bin + 1];
                if (bin < 7) lowcomp = lc1(lowcomp, b0, b1);
                else if (bin < 20) lowcomp = (b0 + 256 == b1) ? 320 : (b0 > b1) ? max(lowcomp - 64, 0) : lowcomp;
                else lowcomp = max(lowcomp - 128, 0);
            }
            fast = max(fast - fdecay, psd[bin] - fgain);
            slow = max(slow - sdecay, psd[bin] - sgain);
            mk[bin] = (int16_t)max(fast - lowcomp, slow);
        }
    }
    // bands 22 and up: fast = max(fast - fdecay, psd - fgain) and the same for slow are max-plus recurrences without
    // side conditions, i.e. prefix maxima of psd + j * decay: lane j takes band 22 + j
    if (bndend > 22) {
        const int bin = 22 + lane;
        const bool on = bin < bndend;
        const int p = on ? psd[bin] : -(1 << 20);
        int pf = p - fgain + lane * fdecay, ps = p - sgain + lane * sdecay;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int tf = __shfl_up_sync(0xffffffffu, pf, o), ts = __shfl_up_sync(0xffffffffu, ps, o);
            if (lane >= o) { pf = max(pf, tf); ps = max(ps, ts); }
        }
        const int f = max(fast21 - (lane + 1) * fdecay, pf - lane * fdecay);
        const int sl = max(slow21 - (lane + 1) * sdecay, ps - lane * sdecay);
        if (on) mk[bin] = (int16_t)max(f, sl);
    }
    __syncwarp();
    for (int band = lane; band < bndend; band += 32) {
        int v1 = mk[band];
        const int tmp = dbknee - psd[band];
        if (tmp > 0) v1 += tmp >> 2;
        mk[band] = (int16_t)max(v1, (int)T.hth[P.fscod * 50 + (band >> P.halfrate)]);
    }
    __syncwarp();
}

// E3b: baps of one exponent set for an snr offset (:393-420) + class counts.  One warp, in two steps.
// With mask' = (max(mask - snroffset - 0x1f0, 0) & 0x1fe0) + 0x1f0 = 32 k + 496 and psd = 3072 - 128 exp, the bap
// address (psd - mask') >> 5 is 80 - 4 exp - k: step 1 (lanes = bands) leaves k per band in `kb`, step 2 (a lane owns
// eight consecutive bins) looks the class counters up by address and sums them in one packed word.
__device__ __forceinline__ void e3_bands(const EncShared& S, const EncTables& T, const EncParams& P, int blk, int ch,
                                         int lane, int snroffset, uint8_t* kb)
{
    const int end = (P.lfe && ch == 5) ? 7 : 223;
    const int bndend = T.masktab[end - 1] + 1;
    const int16_t* mk = S.u.e1.mask[blk][ch];
    for (int band = lane; band < bndend; band += 32)
        kb[band] = (uint8_t)((max(mk[band] - snroffset - 0x1f0, 0) >> 5) & 0xff);
}

__device__ __forceinline__ void e3_probe(EncShared& S, const EncTables& T, const EncParams& P, int blk, int ch, int lane,
                                         const uint8_t* kb, bool store, int (*cnt)[6][4])
{
    const int end = (P.lfe && ch == 5) ? 7 : 223;
    const int i0 = 8 * lane;
    const int nvalid = min(max(end - i0, 0), 8);
    const uint2 e8 = *reinterpret_cast<const uint2*>(S.enc[blk][ch] + i0);
    const uint2 m8 = *reinterpret_cast<const uint2*>(T.masktab + i0);
    uint32_t acc = 0;
    uint2 b8 = make_uint2(0u, 0u);
#pragma unroll
    for (int k = 0; k < 8; k++) {
        if (k < nvalid) {
            const int sh = 8 * (k & 3);
            const int ex = (int)(((k < 4 ? e8.x : e8.y) >> sh) & 0xff);
            const int band = (int)(((k < 4 ? m8.x : m8.y) >> sh) & 0xff);
            const int a = min(max(80 - 4 * ex - (int)kb[band], 0), 63);
            acc += T.tabc[a];
            if (store) {
                const uint32_t b = T.baptab[a];
                if (k < 4) b8.x |= b << sh; else b8.y |= b << sh;
            }
        }
    }
    if (store && nvalid) *reinterpret_cast<uint2*>(S.expo[blk][ch] + i0) = b8;   // raw exponents are dead: baps live there
    const uint32_t n = __reduce_add_sync(0xffffffffu, acc >> 8);
    const uint32_t fixed = __reduce_add_sync(0xffffffffu, acc & 0xff);
    if (lane == 0) {
        cnt[blk][ch][0] = (int)(n & 0xff);
        cnt[blk][ch][1] = (int)((n >> 8) & 0xff);
        cnt[blk][ch][2] = (int)(n >> 16);
        cnt[blk][ch][3] = (int)fixed;
    }
}

// The same for THREE snr offsets in one pass over the set (the search evaluates its next three likely candidates at a
// time, see the kernel): the exponent, band and mask of a bin are fetched once, k is worked out per candidate on the spot
// (no band table), and three packed counters run side by side.  cnt[candidate][blk][ch][4].
__device__ __forceinline__ void e3_probe3(EncShared& S, const EncTables& T, const EncParams& P, int blk, int ch, int lane,
                                          int snro0, int snro1, int snro2, int (*cnt)[6][6][4])
{
    const int end = (P.lfe && ch == 5) ? 7 : 223;
    const int i0 = 8 * lane;
    const int nvalid = min(max(end - i0, 0), 8);
    uint2 e8 = *reinterpret_cast<const uint2*>(S.enc[blk][ch] + i0);
    uint2 m8 = *reinterpret_cast<const uint2*>(T.masktab + i0);
    const int16_t* mk = S.u.e1.mask[blk][ch];
    uint32_t acc0 = 0, acc1 = 0, acc2 = 0;
#pragma unroll 1
    for (int it = 0; it < 4; it++) {
#pragma unroll
        for (int j = 0; j < 2; j++) {
            if (2 * it + j < nvalid) {
                const int ex = (int)((e8.x >> (8 * j)) & 0xff);
                const int band = (int)((m8.x >> (8 * j)) & 0xff);
                const int q = 80 - 4 * ex, t = mk[band] - 0x1f0;
                acc0 += T.tabc[min(max(q - ((max(t - snro0, 0) >> 5) & 0xff), 0), 63)];
                acc1 += T.tabc[min(max(q - ((max(t - snro1, 0) >> 5) & 0xff), 0), 63)];
                acc2 += T.tabc[min(max(q - ((max(t - snro2, 0) >> 5) & 0xff), 0), 63)];
            }
        }
        e8.x = __funnelshift_r(e8.x, e8.y, 16); e8.y >>= 16;
        m8.x = __funnelshift_r(m8.x, m8.y, 16); m8.y >>= 16;
    }
    const uint32_t n0 = __reduce_add_sync(0xffffffffu, acc0 >> 8), f0 = __reduce_add_sync(0xffffffffu, acc0 & 0xff);
    const uint32_t n1 = __reduce_add_sync(0xffffffffu, acc1 >> 8), f1 = __reduce_add_sync(0xffffffffu, acc1 & 0xff);
    const uint32_t n2 = __reduce_add_sync(0xffffffffu, acc2 >> 8), f2 = __reduce_add_sync(0xffffffffu, acc2 & 0xff);
    if (lane < 3) {
        const uint32_t n = lane == 0 ? n0 : lane == 1 ? n1 : n2, f = lane == 0 ? f0 : lane == 1 ? f1 : f2;
        *reinterpret_cast<int4*>(cnt[lane][blk][ch]) = make_int4((int)(n & 0xff), (int)((n >> 8) & 0xff), (int)(n >> 16), (int)f);
    }
}

// bits left in the frame for each of the three candidates of a pass: lanes 8 c + block
__device__ __forceinline__ void bits_left3(const EncShared& S, const EncParams& P, int lane, const int (*cnt)[6][6][4],
                                           int& left0, int& left1, int& left2)
{
    int used = 0;
    const int c = lane >> 3, b = lane & 7;
    if (c < 3 && b < 6) {
        int n1 = 0, n2 = 0, n4 = 0;
        for (int ch = 0; ch < P.nch_all; ch++) {
            const int4 q = *reinterpret_cast<const int4*>(cnt[c][S.head[b][ch]][ch]);
            n1 += q.x; n2 += q.y; n4 += q.z; used += q.w;
        }
        used += 5 * ((n1 + 2) / 3) + 7 * ((n2 + 2) / 3) + 7 * ((n4 + 1) / 2);
    }
    used += __shfl_xor_sync(0xffffffffu, used, 1);
    used += __shfl_xor_sync(0xffffffffu, used, 2);
    used += __shfl_xor_sync(0xffffffffu, used, 4);
    const int left = 16 * P.frame_words - S.frame_bits - used;
    left0 = __shfl_sync(0xffffffffu, left, 0);
    left1 = __shfl_sync(0xffffffffu, left, 8);
    left2 = __shfl_sync(0xffffffffu, left, 16);
}

// The search of compute_bit_allocation (:921-967) as a state machine fed with one probe result
// at a time.  phase 0: csnr down by 4 until it fits; 1: csnr up by 4; 2: csnr up by 1;
// 3: fsnr up by 4; 4: fsnr up by 1; 5: done.  Every warp runs its own copy on the same inputs, so a
// probe costs one CTA barrier, not two.
struct Search {
    int phase, cs, fs, probe_cs, probe_fs, done, failed;
};

__device__ __forceinline__ void search_step(Search& q, int left)
{
    int ph = q.phase;
    if (ph == 0) {
        if (left >= 0) { q.cs = q.probe_cs; ph = 1; }
        else {
            q.probe_cs -= 4;
            if (q.probe_cs < 0) { q.failed = 1; q.cs = 0; q.fs = 0; q.phase = 5; q.done = 1; }
            return;
        }
    } else if (left >= 0) {
        q.cs = q.probe_cs;
        q.fs = q.probe_fs;
    } else {
        ph++;
    }
    // next candidate of the current phase, falling through exhausted phases
    for (;;) {
        if (ph == 1) { if (q.cs + 4 <= 63) { q.probe_cs = q.cs + 4; q.probe_fs = 0; break; } ph = 2; }
        else if (ph == 2) { if (q.cs + 1 <= 63) { q.probe_cs = q.cs + 1; q.probe_fs = 0; break; } ph = 3; }
        else if (ph == 3) { if (q.fs + 4 <= 15) { q.probe_cs = q.cs; q.probe_fs = q.fs + 4; break; } ph = 4; }
        else if (ph == 4) { if (q.fs + 1 <= 15) { q.probe_cs = q.cs; q.probe_fs = q.fs + 1; break; } ph = 5; }
        else { q.done = 1; break; }
    }
    q.phase = ph;
}

// ---------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 3)
ac3_encode_kernel(const EncParams P)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    EncTables& T = *reinterpret_cast<EncTables*>(smem_raw);
    EncShared& S = *reinterpret_cast<EncShared*>(smem_raw + ((sizeof(EncTables) + 15) & ~(size_t)15));
    __shared__ int s_stream;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(&g_enc_tables);
        uint32_t* dst = reinterpret_cast<uint32_t*>(&T);
        for (int i = tid; i < (int)(sizeof(EncTables) / 4); i += kThreads) dst[i] = src[i];
    }
    __syncthreads();
    const int nbytes = P.frame_words * 2;
    const bool active = warp < P.nch_all;                              // warp = coded channel
    // CRC constants of this thread (frame end, :1599-1638): warps 0..2 share crc1's range, bytes [4, 5/8 of the frame),
    // warps 3..5 crc2's, [5/8, end - 2); crc1's multipliers also carry the inverse-polynomial factor (:1627)
    const int fs58 = (P.frame_words >> 1) + (P.frame_words >> 3);
    const int crc_b0 = warp < 3 ? 4 : 2 * fs58, crc_b1 = warp < 3 ? 2 * fs58 : nbytes - 2;
    const int crc_L = (crc_b1 - crc_b0 + 95) / 96;
    const int crc_start = crc_b1 - (96 - (tid - (warp < 3 ? 0 : 96))) * crc_L;
    uint32_t crc_mult = 1;
    {
        const int after = crc_b1 - (crc_start + crc_L);
#pragma unroll 1
        for (int k = 0; k < after; k++) crc_mult = (T.crc_table[crc_mult >> 8] ^ (crc_mult << 8)) & 0xffff;
        if (warp < 3) crc_mult = mul_poly(crc_mult, P.crc_inv);
    }

    // Work units are slices of streams (P.slice_frames frames), handed out slice-major from one ticket counter;
    // a slice starts from the carry record its predecessor left in global memory.  The predecessor holds a
    // smaller ticket and every CTA of the grid is resident, so the wait below cannot deadlock.  Slices keep
    // the last wave of a launch short (4096 streams on 444 resident CTAs are 9.2 waves of whole streams).
    const int nunits = P.nstreams * P.nslices;
    for (;;) {
        if (tid == 0) s_stream = atomicAdd(P.work_counter, 1);
        __syncthreads();
        const int ticket = s_stream;
        if (ticket >= nunits) break;
        const int slice = ticket / P.nstreams, s = ticket - slice * P.nstreams;
        const int f_begin = slice * P.slice_frames;
        const int f_end = (P.nslices > 1 && f_begin + P.slice_frames < P.nframes) ? f_begin + P.slice_frames : P.nframes;
        if (slice > 0) {
            if (tid == 0) {
                volatile int* done = P.slice_done + s;
                while (*done < slice) __nanosleep(200);
                __threadfence();
            }
            __syncthreads();
        }
        // carry in
        const bool started = P.carry && __ldcg(&P.carry[s].started);
        for (int i = tid; i < 6 * 256; i += kThreads)
            S.last[i >> 8][i & 255] = started ? __ldcg(&P.carry[s].last_samples[i >> 8][i & 255]) : (int16_t)0;
        if (tid == 0) S.cs = started ? __ldcg(&P.carry[s].csnroffst) : 40;      // :1092
        __syncthreads();

        for (int f = f_begin; f < f_end; f++) {
            const size_t fidx = (size_t)s * P.nframes + f;
            const int16_t* pcm = P.pcm + fidx * 1536 * P.nch_all;
            // ================= E1 =================
            // a warp works through its channel's six blocks on its own: samples straight from global memory
            // (the six warps of the CTA share the lines), one block ahead of the transform that uses them
            if (active) {
                const int16_t* mine = pcm + P.chmap[warp];
                int nxt[8];
#pragma unroll
                for (int r = 0; r < 8; r++) nxt[r] = __ldg(mine + (size_t)(lane + 32 * r) * P.nch_all);
#pragma unroll 1
                for (int blk = 0; blk < 6; blk++) {
                    int cur[8];
#pragma unroll
                    for (int r = 0; r < 8; r++) cur[r] = nxt[r];
                    if (blk < 5) {
#pragma unroll
                        for (int r = 0; r < 8; r++)
                            nxt[r] = __ldg(mine + (size_t)((blk + 1) * 256 + lane + 32 * r) * P.nch_all);
                    }
                    e1_transform(S, T, blk, warp, lane, cur);
                }
            }
            if (P.dbg_coef) __syncthreads();
            if (P.dbg_coef) {
                for (int i = tid; i < 6 * 6 * 256; i += kThreads) {
                    const int ch = (i >> 8) % 6;
                    P.dbg_coef[fidx * 9216 + i] = ch < P.nch_all ? (&S.coef[0][0][0])[i] : 0;
                }
                if (tid < 36) P.dbg_shift[fidx * 36 + tid] = (tid % 6) < P.nch_all ? (&S.exp_shift[0][0])[tid] : 0;
            }
            // ================= E2 =================
            uint32_t sets = 0;                                           // this channel's exponent sets, one bit per block
            if (active) {
                __syncwarp();
                const int bits = e2_exponents(S, P, warp, lane, sets);
                // this channel's share of everything but mantissas (:880-916): its exponent sections and, per set of a
                // full-bandwidth channel, chbwcod and the gain range field
                if (lane == 0) S.exp_bits[warp] = bits + (warp < P.nch ? 8 * __popc(sets) : 0);
            }
            // ================= E3 =================
            if (active) {
                int16_t* psd = reinterpret_cast<int16_t*>(S.u.e1.z[warp]);      // scratch: 50 band values
                for (uint32_t m = sets; m; m &= m - 1) e3_mask(S, T, P, __ffs(m) - 1, warp, lane, psd);
            }
            __syncthreads();
            {
                // the rest of it is the same for every frame of the call
                static const int inc[8] = {0, 0, 2, 2, 2, 4, 2, 4};
                int fb = 65 + inc[P.acmod] + 6 * (P.nch * 2 + 2 + (P.acmod == 2 ? 1 : 0) + 2 * P.nch + (P.lfe ? 1 : 0) + 4)
                       + 1 + 2 * 4 + 3 + 6 + P.nch_all * (4 + 3) + 2 + 16;
                for (int ch = 0; ch < P.nch_all; ch++) fb += S.exp_bits[ch];
                if (tid == 0) S.frame_bits = fb;                         // (read after the first pass's barrier)
            }
            Search q;
            q.cs = S.cs;
            q.probe_cs = q.cs;                                           // warm start (:921)
            q.probe_fs = q.fs = q.phase = q.done = q.failed = 0;
            // The reference probes one (csnr, fsnr) at a time (:921-967).  Here a pass evaluates the NEXT THREE probes the
            // search will most likely ask for - following "fits" in phases 0, 3, 4 and "does not fit" in phases 1, 2 - and
            // the state machine is then replayed on the results of this and the previous pass for as long as it finds
            // what it asks for: the same decisions on the same numbers, 7.7 probes in 3.0 passes (and barriers) on average.
            int key_a0 = -1, key_a1 = -1, key_a2 = -1, key_b0 = -1, key_b1 = -1, key_b2 = -1;      // previous, current pass
            int left_a0 = 0, left_a1 = 0, left_a2 = 0, left_b0 = 0, left_b1 = 0, left_b2 = 0;
            for (int par = 0;; par ^= 1) {
                key_a0 = key_b0; key_a1 = key_b1; key_a2 = key_b2;
                left_a0 = left_b0; left_a1 = left_b1; left_a2 = left_b2;
                {
                    Search t = q;
                    key_b0 = t.probe_cs * 16 + t.probe_fs;
                    search_step(t, (t.phase == 0 || t.phase >= 3) ? 0 : -1);
                    key_b1 = t.done ? key_b0 : t.probe_cs * 16 + t.probe_fs;
                    if (!t.done) search_step(t, (t.phase == 0 || t.phase >= 3) ? 0 : -1);
                    key_b2 = t.done ? key_b1 : t.probe_cs * 16 + t.probe_fs;
                }
                if (active) {
                    const int s0 = ((((key_b0 >> 4) - 15) << 4) + (key_b0 & 15)) << 2;
                    const int s1 = ((((key_b1 >> 4) - 15) << 4) + (key_b1 & 15)) << 2;
                    const int s2 = ((((key_b2 >> 4) - 15) << 4) + (key_b2 & 15)) << 2;
                    for (uint32_t m = sets; m; m &= m - 1) e3_probe3(S, T, P, __ffs(m) - 1, warp, lane, s0, s1, s2, S.u.e1.pcnt[par]);
                }
                __syncthreads();
                bits_left3(S, P, lane, S.u.e1.pcnt[par], left_b0, left_b1, left_b2);
                while (!q.done) {
                    const int key = q.probe_cs * 16 + q.probe_fs;
                    int left;
                    if (key == key_b0) left = left_b0;
                    else if (key == key_b1) left = left_b1;
                    else if (key == key_b2) left = left_b2;
                    else if (key == key_a0) left = left_a0;
                    else if (key == key_a1) left = left_a1;
                    else if (key == key_a2) left = left_a2;
                    else break;
                    search_step(q, left);
                }
                if (q.done) break;
            }
            {
                // the accepted allocation (or, after a failed search, all-zero baps)
                const int snro = (((q.cs - 15) << 4) + q.fs) << 2;
                if (tid == 0) { S.cs = q.cs; S.fs = q.fs; S.failed = q.failed; }
                if (active) {
                    uint8_t* kb = reinterpret_cast<uint8_t*>(S.u.e1.z[warp]);
                    if (!q.failed) {
                        for (uint32_t m = sets; m; m &= m - 1) {
                            const int blk = __ffs(m) - 1;
                            e3_bands(S, T, P, blk, warp, lane, snro, kb + 64 * blk);
                        }
                    }
                    __syncwarp();
                    for (uint32_t m = sets; m; m &= m - 1) {
                        const int blk = __ffs(m) - 1;
                        if (q.failed) {
                            for (int i = lane; i < 256; i += 32) S.expo[blk][warp][i] = 0;
                            if (lane < 4) S.cnt[blk][warp][lane] = 0;
                        } else {
                            e3_probe(S, T, P, blk, warp, lane, kb + 64 * blk, true, S.cnt);
                        }
                    }
                }
            }
            __syncthreads();
            if (P.dbg_bap) {
                for (int i = tid; i < 6 * 6 * 256; i += kThreads) {
                    const int blk = i / 1536, ch = (i >> 8) % 6, k = i & 255;
                    const bool ok = ch < P.nch_all && k < ((P.lfe && ch == 5) ? 7 : 223);
                    const int h = ok ? S.head[blk][ch] : 0;
                    P.dbg_bap[fidx * 9216 + i] = ok ? S.expo[h][ch][k] : 0;
                    P.dbg_enc[fidx * 9216 + i] = ok ? S.enc[h][ch][k] : 0;
                }
                if (tid < 36) P.dbg_strategy[fidx * 36 + tid] = (tid % 6) < P.nch_all ? (&S.strategy[0][0])[tid] : 0;
                if (tid == 0) { P.dbg_snr[fidx * 2] = S.cs; P.dbg_snr[fidx * 2 + 1] = S.fs; }
            }
            // ================= E4 =================
            uint32_t* frame = S.u.e4.frame;
            for (int i = tid; i < kFrameWords; i += kThreads) frame[i] = 0;
            __syncthreads();
            if (warp == 0) {
                // side information (:1113-1147, 1210-1259, 1316-1337): lane 0 writes the BSI, lanes 0..5 then size one
                // block each and a prefix sum places the blocks.  A block's fields are two runs of at most 64 bits -
                // up to the exponents, and from the bit allocation flag to the skip flag - which a lane collects in
                // a register and writes with three atomic ORs each; the exponent and mantissa sections between and
                // after them are written later by the block's warp, their positions recorded here.
                if (lane == 0) {
                    BitRun w{0, 0};
                    w.put(2, P.fscod);
                    w.put(6, P.frmsizecod);
                    w.put(5, P.bsid);
                    w.put(3, 0);
                    w.put(3, P.acmod);
                    if ((P.acmod & 1) && P.acmod != 1) w.put(2, 1);
                    if (P.acmod & 4) w.put(2, 1);
                    if (P.acmod == 2) w.put(2, 0);
                    w.put(1, P.lfe);
                    w.put(5, 31);
                    w.put(4, 0);
                    w.put(1, 1);
                    w.put(3, 0);
                    atomicOr(&frame[0], 0x0b770000u);                    // sync word; crc1 comes at the frame end
                    w.flush(frame, 32);
                }
                const uint32_t bsi_bits = 16 + 16 + 2 + 6 + 5 + 3 + 3 + (((P.acmod & 1) && P.acmod != 1) ? 2 : 0)
                                        + ((P.acmod & 4) ? 2 : 0) + (P.acmod == 2 ? 2 : 0) + 1 + 5 + 4 + 1 + 3;
                const int blk = lane < 6 ? lane : 5;
                // (this is the one serial stretch the other five warps wait for: strategies are read once, runs of
                // equal-width fields are put together in 32-bit arithmetic before they enter the 64-bit run)
                uint32_t exp_len[6], len = 0, mant = 0;
                uint32_t expstr = 0, bwcod = 0;                          // chexpstr of the full-bandwidth channels; chbwcod fields
                int nnew = 0, lfe_st = 0;
                {
                    int n1 = 0, n2 = 0, n4 = 0;
#pragma unroll
                    for (int ch = 0; ch < 6; ch++) {
                        exp_len[ch] = 0;
                        if (ch >= P.nch_all) continue;
                        const int st = S.strategy[blk][ch];
                        const bool ch_lfe = P.lfe && ch == 5;
                        if (st) exp_len[ch] = 4 + 7 * exp_groups(st, ch_lfe) + (ch_lfe ? 0 : 2);
                        if (ch_lfe) lfe_st = st;
                        else {
                            expstr = (expstr << 2) | (uint32_t)st;
                            if (st) { bwcod = (bwcod << 6) | 50u; nnew++; }
                        }
                        const int4 q = *reinterpret_cast<const int4*>(S.cnt[S.head[blk][ch]][ch]);
                        n1 += q.x; n2 += q.y; n4 += q.z; mant += q.w;
                        len += exp_len[ch];
                    }
                    mant += 5 * ((n1 + 2) / 3) + 7 * ((n2 + 2) / 3) + 7 * ((n4 + 1) / 2);
                    len += 2 * P.nch + 1 + (blk == 0 ? 2 : 1) + (P.acmod == 2 ? (blk == 0 ? 5 : 1) : 0)
                         + 2 * P.nch + (P.lfe ? 1 : 0) + 6 * nnew
                         + 1 + (blk == 0 ? 11 : 0) + 1 + (blk == 0 ? 6 + 7 * P.nch_all : 0) + 2 + mant;
                    if (lane >= 6) len = 0;
                }
                uint32_t incl = len;
#pragma unroll
                for (int o = 1; o < 8; o <<= 1) {
                    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += t;
                }
                if (lane < 6) {
                    uint32_t pos = bsi_bits + incl - len;
                    BitRun w{0, 0};
                    w.put(2 * P.nch + 1, ((1u << P.nch) - 1) << 1);      // blksw (zeros), dithflag (ones), dynrnge (0)
                    if (blk == 0) w.put(2, 2); else w.put(1, 0);         // cplstre [cplinu]
                    if (P.acmod == 2) { if (blk == 0) w.put(5, 16); else w.put(1, 0); }
                    w.put(2 * P.nch, expstr);                            // chexpstr
                    if (P.lfe) w.put(1, lfe_st);                         // lfeexpstr
                    w.put(6 * nnew, bwcod);                              // chbwcod of the channels with new exponents
                    pos = w.flush(frame, pos);                           // at most 11 + 2 + 5 + 10 + 1 + 30 = 59 bits
#pragma unroll
                    for (int ch = 0; ch < 6; ch++) {
                        if (ch >= P.nch_all || !exp_len[ch]) continue;
                        S.exp_pos[blk][ch] = pos;
                        pos += exp_len[ch];
                    }
                    w = BitRun{0, 0};
                    if (blk == 0) {
                        // baie + the five allocation parameters, snroffste + csnroffst, then fsnroffst / fgaincod per channel
                        w.put(12, (1u << 11) | (2u << 9) | (1u << 7) | (1u << 5) | (2u << 3) | 4u);
                        w.put(7, (1u << 6) | (uint32_t)S.cs);
                        const uint32_t f7 = ((uint32_t)S.fs << 3) | 4u, f21 = f7 | f7 << 7 | f7 << 14;
                        int ch = 0;
                        for (; ch + 3 <= P.nch_all; ch += 3) w.put(21, f21);
                        for (; ch < P.nch_all; ch++) w.put(7, f7);
                        w.put(2, 0);
                    } else {
                        w.put(4, 0);                                     // baie, snroffste, deltbaie, skiple: all off
                    }
                    S.mant_pos[blk] = w.flush(frame, pos);               // at most 1 + 11 + 1 + 6 + 42 + 2 = 63 bits
                }
            }
            __syncthreads();
            if (active) {
                // grouped exponents (:1261-1314): warp = channel (the sets are spread evenly over the channels, not
                // over the blocks), lanes = groups
                const int ch = warp;
                for (int blk = 0; blk < 6; blk++) {
                    const int st = S.strategy[blk][ch];
                    if (!st) continue;
                    const uint8_t* en = S.enc[blk][ch];
                    const int gs = st == 1 ? 1 : st == 2 ? 2 : 4;
                    const int ng = exp_groups(st, P.lfe && ch == 5);
                    const uint32_t p0 = S.exp_pos[blk][ch];
                    if (lane == 0) put_bits_atomic(frame, p0, 4, en[0]);
                    for (int g = lane; g < ng; g += 32) {
                        const int k = 1 + 3 * g * gs;
                        const int ea = g ? en[k - gs] : en[0];
                        const int d0 = en[k] - ea + 2, d1 = en[k + gs] - en[k] + 2, d2 = en[k + 2 * gs] - en[k + gs] + 2;
                        put_bits_atomic(frame, p0 + 4 + 7 * g, 7, (uint32_t)((d0 * 5 + d1) * 5 + d2));
                    }
                }
            }
            {
                // mantissas (:1346-1501): warp = audio block, walking the block's channels in
                // coded order (the occurrence numbers of the grouped classes run across channels); a lane owns eight
                // consecutive bins.  No barrier and no accumulator shared between warps: group members go to the
                // block's rings, the lanes then emit the codes of the groups the channel closed.
                const int blk = warp;
                uint8_t* rv = &S.u.e4.ring_v[warp][0][0];
                uint16_t* rp = &S.u.e4.ring_p[warp][0][0];
                int N1 = 0, N2 = 0, N4 = 0;
                uint32_t pos0 = S.mant_pos[blk];
                auto g3 = [](int x) { return (int)(((uint32_t)(x + 2) * 43691u) >> 17); };   // members x' < x with x' % 3 == 0
                auto d3 = [](int x) { return (int)(((uint32_t)x * 43691u) >> 17); };         // x / 3 below 2^15
                for (int ch = 0; ch < P.nch_all; ch++) {
                    const bool is_lfe = P.lfe && ch == 5;
                    const int ncoef = is_lfe ? 7 : 223;
                    const int h = S.head[blk][ch];
                    const uint8_t* en = S.enc[h][ch];
                    // mantissas
                    const int i0 = 8 * lane;
                    const int nvalid = min(max(ncoef - i0, 0), 8);
                    uint2 b8 = *reinterpret_cast<const uint2*>(S.expo[h][ch] + i0);
                    const uint2 e8 = *reinterpret_cast<const uint2*>(en + i0);
                    {
                        // bins past the coded range hold no bap
                        const uint64_t keep = nvalid >= 8 ? ~0ull : ((1ull << (8 * nvalid)) - 1);
                        b8.x &= (uint32_t)keep;
                        b8.y &= (uint32_t)(keep >> 32);
                    }
                    uint32_t acc = 0;
#pragma unroll
                    for (int k = 0; k < 8; k++) acc += T.taba[((k < 4 ? b8.x : b8.y) >> (8 * (k & 3))) & 15];
                    const uint32_t cnt = acc >> 8, pl = acc & 0xff;
                    uint32_t icnt = cnt, ipl = pl;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const uint32_t t = __shfl_up_sync(0xffffffffu, icnt, o), u = __shfl_up_sync(0xffffffffu, ipl, o);
                        if (lane >= o) { icnt += t; ipl += u; }
                    }
                    const uint32_t tcnt = __shfl_sync(0xffffffffu, icnt, 31), tpl = __shfl_sync(0xffffffffu, ipl, 31);
                    const uint32_t ec = icnt - cnt;
                    int X1 = N1 + (int)(ec & 0xff), X2 = N2 + (int)((ec >> 8) & 0xff), X4 = N4 + (int)(ec >> 16);
                    uint32_t pos = pos0 + (ipl - pl) + 5 * (g3(X1) - g3(N1)) + 7 * (g3(X2) - g3(N2)) + 7 * (((X4 + 1) >> 1) - ((N4 + 1) >> 1));
                    const int gexp = S.exp_shift[blk][ch];
                    int cf[8];
                    {
                        const int4 c0 = *reinterpret_cast<const int4*>(&S.coef[blk][ch][i0]);
                        const int4 c1 = *reinterpret_cast<const int4*>(&S.coef[blk][ch][i0 + 4]);
                        cf[0] = c0.x; cf[1] = c0.y; cf[2] = c0.z; cf[3] = c0.w;
                        cf[4] = c1.x; cf[5] = c1.y; cf[6] = c1.z; cf[7] = c1.w;
                    }
                    // (a real loop, two mantissas per trip: unrolled eight times the quantisers miss the instruction
                    // cache; no branch on the bap - lanes hold different classes, so every path is predicated work)
                    uint2 e8r = e8;
                    uint32_t Xm;                                         // the classes' occurrence counters at this lane
                    {
                        const uint32_t m1 = N1 >= 768 ? N1 - 768 : N1, m2 = N2 >= 768 ? N2 - 768 : N2, m4 = N4 >= 768 ? N4 - 768 : N4;
                        Xm = (m1 + (ec & 0xff)) | (m2 + ((ec >> 8) & 0xff)) << 10 | (m4 + (ec >> 16)) << 20;
                    }
                    BitRun run{0, 0};
                    uint32_t run_pos = pos;
#pragma unroll 1
                    for (int it = 0; it < 4; it++) {
#pragma unroll
                        for (int j = 0; j < 2; j++) {
                            const uint32_t b = (b8.x >> (8 * j)) & 15;
                            const int c = cf[j];
                            const int e = (int)((e8r.x >> (8 * j)) & 0xff) - gexp;
                            const uint2 tq = T.tabq[b];
                            const int lv = (int)(tq.x & 0xff), qb = (int)((tq.x >> 8) & 0xff), cl = (int)(tq.x >> 24);
                            const uint32_t wd = (tq.x >> 16) & 0xff;
                            int v;
                            {
                                // symmetric (:1150-1166) and asymmetric (:1169-1190) quantiser side by side
                                const int a = abs(c);
                                int vs = (lv * (a << e)) >> 24;
                                vs = (vs + 1) >> 1;
                                vs = (lv >> 1) + (c >= 0 ? vs : -vs);
                                const int lshift = e + qb - 24;
                                int va = lshift >= 0 ? c << lshift : c >> (-lshift);
                                va = (va + 1) >> 1;
                                const int m = 1 << (qb - 1);
                                va = min(va, m - 1);
                                va &= 2 * m - 1;
                                v = lv ? vs : va;
                            }
                            // grouped classes: occurrence number x; x % 3 == 0 <=> x * 0xAAAAAAAB <= 0x55555555 (mod 2^32),
                            // x even <=> x * 2^31 == 0: the member that opens a group reserves the code's place
                            // (counters modulo 768 = 3 * 256 in ten-bit fields: x is only used modulo 256, 3 and 2)
                            const uint32_t xsh = (uint32_t)(10 * cl + 22) & 31;
                            const uint32_t x = (Xm >> xsh) & 0x3ff;
                            Xm += tq.y;                                  // 1 << xsh for a grouped bap, else 0
                            const bool pairs = cl == 3;
                            const bool opens = x * (pairs ? 0x80000000u : 0xAAAAAAABu) <= (pairs ? 0u : 0x55555555u);
                            if (cl) {
                                rv[(cl - 1) * 256 + (x & 255)] = (uint8_t)v;
                                if (opens) rp[(cl - 1) * 128 + ((x >> 1) & 127)] = (uint16_t)min(pos, 65535u);
                            }
                            // the lane's mantissas are consecutive in the stream: ungrouped fields (and zeros where a
                            // group code will go) collect in a run that is written every four mantissas (<= 64 bits)
                            const uint32_t adv = (!cl || opens) ? wd : 0u;
                            run.acc = (run.acc << adv) | ((b && !cl) ? (uint32_t)v : 0u);
                            run.n += adv;
                            pos += adv;
                        }
                        if (it & 1) {
                            run.flush(frame, run_pos);
                            run_pos = pos;
                            run = BitRun{0, 0};
                        }
                        b8.x = __funnelshift_r(b8.x, b8.y, 16); b8.y >>= 16;
                        e8r.x = __funnelshift_r(e8r.x, e8r.y, 16); e8r.y >>= 16;
#pragma unroll
                        for (int j = 0; j < 6; j++) cf[j] = cf[j + 2];
                    }
                    __syncwarp();
                    // the groups this channel closed: lanes = groups
                    const int t1 = (int)(tcnt & 0xff), t2 = (int)((tcnt >> 8) & 0xff), t4 = (int)(tcnt >> 16);
                    {
                        const int d1a = d3(N1), d2a = d3(N2), d4a = N4 >> 1;
                        const int n1g = d3(N1 + t1) - d1a, n2g = d3(N2 + t2) - d2a, n4g = ((N4 + t4) >> 1) - d4a;
                        for (int i = lane; i < n1g + n2g + n4g; i += 32) {
                            int g, cls, m0, m1;
                            if (i < n1g) { g = d1a + i; cls = 0; m0 = 9; m1 = 3; }
                            else if (i < n1g + n2g) { g = d2a + i - n1g; cls = 1; m0 = 25; m1 = 5; }
                            else { g = d4a + i - n1g - n2g; cls = 2; m0 = 11; m1 = 1; }
                            const uint8_t* r = rv + cls * 256;
                            const int x0 = cls == 2 ? 2 * g : 3 * g;
                            int code = m0 * r[x0 & 255] + m1 * r[(x0 + 1) & 255];
                            if (cls != 2) code += r[(x0 + 2) & 255];
                            put_bits_atomic(frame, rp[cls * 128 + ((x0 >> 1) & 127)], cls == 0 ? 5 : 7, (uint32_t)code);
                        }
                    }
                    pos0 += tpl + 5 * (g3(N1 + t1) - g3(N1)) + 7 * (g3(N2 + t2) - g3(N2)) + 7 * (((N4 + t4 + 1) >> 1) - ((N4 + 1) >> 1));
                    N1 += t1; N2 += t2; N4 += t4;
                    __syncwarp();
                }
                // groups still open at the end of the block carry what they have (:1375-1421)
                if (lane < 3) {
                    const int n = lane == 0 ? N1 : lane == 1 ? N2 : N4;
                    const int g = lane == 2 ? n >> 1 : d3(n);
                    const int rem = n - (lane == 2 ? 2 : 3) * g;
                    if (rem) {
                        const uint8_t* r = rv + lane * 256;
                        const int x0 = (lane == 2 ? 2 : 3) * g;
                        const int m0 = lane == 0 ? 9 : lane == 1 ? 25 : 11, m1 = lane == 0 ? 3 : 5;
                        int code = m0 * r[x0 & 255];
                        if (rem == 2) code += m1 * r[(x0 + 1) & 255];
                        put_bits_atomic(frame, rp[lane * 128 + ((x0 >> 1) & 127)], lane == 0 ? 5 : 7, (uint32_t)code);
                    }
                }
            }
            __syncthreads();
            // frame end (:1599-1638): crc1 over the first 5/8 through the inverse polynomial trick, crc2 over the
            // rest, stored over the last two bytes whatever spilled into them (the stereo bit-accounting slip, :889)
            {
                uint32_t crc = mul_poly(crc_chunk(T, frame, crc_b0, crc_start, crc_L), crc_mult);
                crc = __reduce_xor_sync(0xffffffffu, crc);
                if (lane == 0) S.crcp[warp] = crc;
                __syncthreads();
            }
            // store (big-endian bytes): frames start at even offsets, so 16-bit stores; the threads that hold the
            // CRC fields (bytes 2, 3 and the last two) put the sums of the warps' parts there
            {
                const uint32_t c1 = (S.crcp[0] ^ S.crcp[1] ^ S.crcp[2]) & 0xffff, c2 = (S.crcp[3] ^ S.crcp[4] ^ S.crcp[5]) & 0xffff;
                uint16_t* dst = reinterpret_cast<uint16_t*>(P.out + fidx * nbytes);
                for (int i = tid; i < P.frame_words; i += kThreads) {
                    const uint32_t wv = frame[i >> 1];
                    uint32_t h = (i & 1) ? (wv & 0xffff) : (wv >> 16);
                    if (i == 1) h = c1;
                    if (i == P.frame_words - 1) h = c2;
                    dst[i] = (uint16_t)(((h & 0xff) << 8) | (h >> 8));
                }
                if (tid == 0 && P.status) P.status[fidx] = S.failed ? AC3_ST_NO_FIT : AC3_ST_OK;
            }
            __syncthreads();
        }   // frames

        if (P.carry) {
            for (int i = tid; i < 6 * 256; i += kThreads) P.carry[s].last_samples[i >> 8][i & 255] = S.last[i >> 8][i & 255];
            if (tid == 0) { P.carry[s].csnroffst = S.cs; P.carry[s].started = 1; }
        }
        if (P.nslices > 1) {
            __threadfence();
            __syncthreads();
            if (tid == 0) atomicExch(P.slice_done + s, slice + 1);
        }
        __syncthreads();
    }
}

}  // namespace ac3e

#include "ac3_encode_host.inl"
