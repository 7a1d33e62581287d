// a52_host.inl - host side of the engine: context, constant tables, the C ABI
// declared in include/a52_batch.h and the drop-in liba52 API of include/a52.h.
// Included at the end of a52_decode.cu (same translation unit as the kernel).

#include <new>
#include <vector>

#include "../../include/a52.h"
#include "../../include/a52_batch.h"

namespace a52 {

// ---------------------------------------------------------------------------
// host tables
// ---------------------------------------------------------------------------
static const uint8_t h_nfchans[11] = {2, 1, 2, 3, 3, 4, 4, 5, 1, 1, 2};

static int host_syncinfo(const uint8_t* b, int* flags, int* sample_rate, int* bit_rate)
{
    // reference: liba52/parse.c:86-129
    static const uint8_t lfe_bit[8] = {0x10, 0x10, 0x04, 0x04, 0x04, 0x01, 0x04, 0x01};
    if (b[0] != 0x0b || b[1] != 0x77) return 0;
    int bsid = b[5] >> 3;
    if (bsid >= 12) return 0;
    int half = bsid > 8 ? bsid - 8 : 0;
    int acmod = b[6] >> 5;
    *flags = (((b[6] & 0xf8) == 0x50) ? M_DOLBY : acmod) | ((b[6] & lfe_bit[acmod]) ? M_LFE : 0);
    int cod = b[4] & 63;
    if (cod >= 38) return 0;
    int kbps = ac3_bitrate_kbps[cod >> 1];
    *bit_rate = (kbps * 1000) >> half;
    switch (b[4] >> 6) {
    case 0: *sample_rate = 48000 >> half; return 4 * kbps;
    case 1: *sample_rate = 44100 >> half; return 2 * (320 * kbps / 147 + (cod & 1));
    case 2: *sample_rate = 32000 >> half; return 6 * kbps;
    }
    return 0;
}

// output-mode negotiation + level adjustment, reference: liba52/downmix.c:34-160.
// `input` = acmod, or M_DOLBY for a 2/0 stream flagged Dolby surround.
static int host_downmix_init(int input, int flags, float* level, float clev, float slev)
{
    static const char* grant[11] = {"0A222222", "11111111", "0A222222", "0A232323", "0A224444",
                                    "0A224545", "0A236666", "0A236767", "81111111", "91111111",
                                    "0A2AAAAA"};
    const double k3 = 0.7071067811865476, kp3 = 1.4142135623730951;
    int req = flags & M_MASK;
    if (req > M_DOLBY) return -1;
    char ch = grant[req][input & 7];
    int out = ch >= 'A' ? ch - 'A' + 10 : ch - '0';
    // the reference compares the float clev with a double constant here
    // (downmix.c:69-71), which never matches in the float build
    if (out == M_STEREO && (input == M_DOLBY || (input == M_3F && (double)clev == k3))) out = M_DOLBY;
    if (!(flags & M_ADJUST)) return out;
    double adj;
    const int key = (out << 3) + (input & 7);
#define K(i, o) (((o) << 3) + (i))
    switch (key) {
    case K(M_3F, M_MONO): adj = k3 / (1 + clev); break;
    case K(M_STEREO, M_MONO): case K(M_2F2R, M_2F1R): case K(M_3F2R, M_3F1R): adj = k3; break;
    case K(M_3F2R, M_2F1R):
        if (clev < kp3 - 1) { adj = k3; break; }
        /* fall through */
    case K(M_3F, M_STEREO): case K(M_3F1R, M_2F1R): case K(M_3F1R, M_2F2R): case K(M_3F2R, M_2F2R):
        adj = 1 / (1 + clev); break;
    case K(M_2F1R, M_MONO): adj = kp3 / (2 + slev); break;
    case K(M_2F1R, M_STEREO): case K(M_3F1R, M_3F): adj = 1 / (1 + slev * k3); break;
    case K(M_3F1R, M_MONO): adj = k3 / (1 + clev + slev * 0.5); break;
    case K(M_3F1R, M_STEREO): adj = 1 / (1 + clev + slev * k3); break;
    case K(M_2F2R, M_MONO): adj = k3 / (1 + slev); break;
    case K(M_2F2R, M_STEREO): case K(M_3F2R, M_3F): adj = 1 / (1 + slev); break;
    case K(M_3F2R, M_MONO): adj = k3 / (1 + clev + slev); break;
    case K(M_3F2R, M_STEREO): adj = 1 / (1 + clev + slev); break;
    case K(M_MONO, M_DOLBY): adj = kp3; break;
    case K(M_3F, M_DOLBY): case K(M_2F1R, M_DOLBY): adj = 1 / (1 + k3); break;
    case K(M_3F1R, M_DOLBY): case K(M_2F2R, M_DOLBY): adj = 1 / (1 + 2 * k3); break;
    case K(M_3F2R, M_DOLBY): adj = 1 / (1 + 3 * k3); break;
    default: return out;
    }
#undef K
    float a = (float)adj;
    *level = *level * a;
    return out;
}

static void host_mix_levels(int acmod, int cmix, int smix, float* clev, float* slev)
{
    // parse.c:134-137, 152-158
    static const double cm[4] = {0.7071067811865476, 0.5946035575013605, 0.5, 0.5946035575013605};
    static const double sm[4] = {0.7071067811865476, 0.5, 0, 0.5};
    *clev = *slev = 0;
    if ((acmod & 1) && acmod != 1) *clev = (float)cm[cmix];
    if (acmod & 4) *slev = (float)sm[smix];
}

static void build_mode_table(ModeEntry* tab, int req_flags, float level)
{
    for (int ae = 0; ae < 9; ae++)
        for (int cm = 0; cm < 4; cm++)
            for (int sm = 0; sm < 4; sm++) {
                int acmod = ae == 8 ? 2 : ae;
                float clev, slev, lv = level;
                host_mix_levels(acmod, cm, sm, &clev, &slev);
                // as-coded request: the frame's own mode is the request (what a52_syncinfo reports)
                const int in_mode = ae == 8 ? M_DOLBY : acmod;
                const int req = (req_flags & A52_REQ_AS_CODED) ? (in_mode | (req_flags & (M_LFE | M_ADJUST))) : req_flags;
                int out = host_downmix_init(in_mode, req, &lv, clev, slev);
                tab[ae * 16 + cm * 4 + sm].output = out;
                tab[ae * 16 + cm * 4 + sm].level = lv;
            }
}

// which coded channels feed which output channel (downmix.c:480-619 as a sign matrix)
static void build_mix_table(MixEntry* tab)
{
    enum { R_L, R_C, R_R, R_S, R_SL, R_SR, R_A, R_B, R_NONE };
    static const uint8_t in_roles[8][5] = {
        {R_A, R_B, R_NONE, R_NONE, R_NONE}, {R_C, R_NONE, R_NONE, R_NONE, R_NONE},
        {R_L, R_R, R_NONE, R_NONE, R_NONE}, {R_L, R_C, R_R, R_NONE, R_NONE},
        {R_L, R_R, R_S, R_NONE, R_NONE},    {R_L, R_C, R_R, R_S, R_NONE},
        {R_L, R_R, R_SL, R_SR, R_NONE},     {R_L, R_C, R_R, R_SL, R_SR}};
    static const uint8_t out_roles[11][5] = {
        {R_A, R_B, R_NONE, R_NONE, R_NONE},   // CHANNEL
        {R_C, R_NONE, R_NONE, R_NONE, R_NONE},// MONO (everything sums here)
        {R_L, R_R, R_NONE, R_NONE, R_NONE},   // STEREO
        {R_L, R_C, R_R, R_NONE, R_NONE},      // 3F
        {R_L, R_R, R_S, R_NONE, R_NONE},      // 2F1R
        {R_L, R_C, R_R, R_S, R_NONE},         // 3F1R
        {R_L, R_R, R_SL, R_SR, R_NONE},       // 2F2R
        {R_L, R_C, R_R, R_SL, R_SR},          // 3F2R
        {R_A, R_NONE, R_NONE, R_NONE, R_NONE},// CHANNEL1
        {R_B, R_NONE, R_NONE, R_NONE, R_NONE},// CHANNEL2
        {R_L, R_R, R_NONE, R_NONE, R_NONE}};  // DOLBY
    memset(tab, 0, sizeof(MixEntry) * 8 * 11);
    for (int acmod = 0; acmod < 8; acmod++)
        for (int out = 0; out < 11; out++) {
            MixEntry& m = tab[acmod * 11 + out];
            int nout = h_nfchans[out];
            m.nout = nout;
            bool has[9] = {false};
            for (int o = 0; o < nout; o++) has[out_roles[out][o]] = true;
            for (int ch = 0; ch < h_nfchans[acmod]; ch++) {
                int r = in_roles[acmod][ch];
                for (int o = 0; o < nout; o++) {
                    int t = out_roles[out][o];
                    int sign = 0;
                    if (out == M_MONO) sign = 1;
                    else if (out == M_CHANNEL1 || out == M_CHANNEL2 || out == M_CHANNEL) sign = (t == r) ? 1 : 0;
                    else if (t == r) sign = 1;                                 // straight through
                    else if (r == R_C && !has[R_C] && (t == R_L || t == R_R)) sign = 1;
                    else if (r == R_S && !has[R_S]) {
                        if (has[R_SL]) sign = (t == R_SL || t == R_SR) ? 1 : 0;    // 1 surround -> 2
                        else if (out == M_DOLBY) sign = (t == R_L) ? -1 : (t == R_R) ? 1 : 0;
                        else sign = (t == R_L || t == R_R) ? 1 : 0;
                    } else if ((r == R_SL || r == R_SR) && !has[R_SL]) {
                        if (has[R_S]) sign = (t == R_S) ? 1 : 0;                   // 2 surrounds -> 1
                        else if (out == M_DOLBY) sign = (t == R_L) ? -1 : (t == R_R) ? 1 : 0;
                        else sign = ((r == R_SL && t == R_L) || (r == R_SR && t == R_R)) ? 1 : 0;
                    }
                    if (sign > 0) m.pos[o] |= 1u << ch;
                    if (sign < 0) m.neg[o] |= 1u << ch;
                }
            }
            // a52_upmix (downmix.c:621-685): where each downmixed delay plane goes when the decoder
            // falls back to per-channel transforms; the remaining channels' planes are zeroed
            for (int ch = 0; ch < 5; ch++) m.up[ch] = 0xff;
            for (int o = 0; o < nout; o++) {
                int t = out_roles[out][o], dst = -1;
                for (int ch = 0; ch < h_nfchans[acmod] && dst < 0; ch++)
                    if (in_roles[acmod][ch] == t) dst = ch;
                if (dst < 0 && out == M_MONO) dst = 0;
                if (dst < 0 && t == R_S)
                    for (int ch = 0; ch < h_nfchans[acmod] && dst < 0; ch++)
                        if (in_roles[acmod][ch] == R_SL) dst = ch;
                if (dst >= 0 && m.up[dst] == 0xff) m.up[dst] = (uint8_t)o;
            }
        }
}

static void build_tables(Tables* T)
{
    memset(T, 0, sizeof(*T));
    // KBD window, alpha = 5 (imdct.c:347-372): same double recurrence, rounded to float
    {
        double sum = 0, cum[256];
        for (int i = 0; i < 256; i++) {
            double x = i * (256 - i) * (5 * M_PI / 256) * (5 * M_PI / 256), b = 1;
            for (int k = 100; k > 0; k--) b = b * x / (k * k) + 1;
            sum += b;
            cum[i] = sum;
        }
        sum++;
        for (int i = 0; i < 256; i++) T->window[i] = (float)sqrt(cum[i] / sum);
    }
    for (int m = 0; m < 128; m++) {
        double th = (M_PI / 256) * (m + 64 - 0.25), sg = (m & 1) ? -1.0 : 1.0;   // imdct.c:386-396
        T->pre1[m] = make_float2((float)(sg * cos(th)), (float)(sg * sin(th)));
        T->wfft[m] = make_float2((float)cos(2 * M_PI * m / 128), (float)-sin(2 * M_PI * m / 128));
    }
    for (int i = 0; i < 64; i++) {
        T->post1[i] = make_float2((float)cos((M_PI / 256) * (i + 0.5)), (float)sin((M_PI / 256) * (i + 0.5)));
        T->pre2[i] = make_float2((float)cos((M_PI / 128) * (i - 0.25)), (float)sin((M_PI / 128) * (i - 0.25)));
    }
    for (int i = 0; i < 32; i++)
        T->post2[i] = make_float2((float)cos((M_PI / 128) * (i + 0.5)), (float)sin((M_PI / 128) * (i + 0.5)));
    // operands of the packed two-channel transform: every factor twice
    for (int m = 0; m < 128; m++) {
        T->pre1d[m] = make_float4(T->pre1[m].x, T->pre1[m].x, T->pre1[m].y, T->pre1[m].y);
        T->wfftd[m] = make_float4(T->wfft[m].x, T->wfft[m].x, T->wfft[m].y, T->wfft[m].y);
        T->win2d[m] = make_float4(T->window[2 * m], T->window[2 * m], T->window[2 * m + 1], T->window[2 * m + 1]);
    }
    for (int i = 0; i < 64; i++) T->post1d[i] = make_float4(T->post1[i].x, T->post1[i].x, T->post1[i].y, T->post1[i].y);
    // Scratch addressing of the paired transform.  A point is 16 bytes (re L, re R, im L, im R); point i lives at
    // sw(i) = i ^ 3 b3 ^ 4 b4 ^ 2 b5 ^ 4 b6 (b_k = bit k of i), which keeps every access pattern of the
    // radix-4 passes and of the post-twiddle on eight different 16-byte bank groups per quarter warp.  All
    // addresses of a lane are one of five lane constants XOR / plus compile-time constants.
    {
        auto sw = [](int i) { return i ^ (((i >> 3) & 1) * 3) ^ (((i >> 4) & 1) * 4) ^ (((i >> 5) & 1) * 2) ^ (((i >> 6) & 1) * 4); };
        for (int l = 0; l < 32; l++) {
            const int e1 = sw(l);                                         // pass 1 stores: (e1 ^ {0,2,4,6}) + 32 p
            const int b2 = l >> 3, j2 = l & 7;
            const int e2 = (b2 * 32 + j2) ^ ((b2 & 1) * 2) ^ ((b2 >> 1) * 4);   // pass 2: (e2 ^ {0,3,4,7}) + 8 k
            const int b3 = l >> 1, j3 = l & 1;
            const int s3 = ((b3 & 1) * 3) ^ (((b3 >> 1) & 1) * 4) ^ (((b3 >> 2) & 1) * 2) ^ (((b3 >> 3) & 1) * 4);
            const int e3 = 8 * b3 + (j3 ^ s3);                            // pass 3: e3 ^ 2 k
            const int p0 = 32 * (l & 3) + 8 * ((l >> 2) & 3) + 2 * (l >> 4);
            const int nl = (~l) & 31;
            const int p0b = 32 * (nl & 3) + 8 * ((nl >> 2) & 3) + 2 * ((nl >> 4) & 1);
            T->fftaddr[l] = make_uint4((uint32_t)(e1 * 16) | ((uint32_t)(e2 * 16) << 16),
                                       (uint32_t)(e3 * 16) | ((uint32_t)(sw(p0) * 16) << 16),
                                       (uint32_t)(sw(p0b) * 16), 0u);
        }
    }
    for (int c = 0; c < 32; c++)
        for (int d = 0; d < 3; d++) {
            int dig[3] = {c / 9, (c / 3) % 3, c % 3};
            T->q1[d][c] = c < 27 ? ac3_q3[dig[d]] : 0;
        }
    for (int c = 0; c < 128; c++) {
        int d5[3] = {c / 25, (c / 5) % 5, c % 5};
        int d11[2] = {c / 11, c % 11};
        for (int d = 0; d < 3; d++) T->q2[d][c] = c < 125 ? ac3_q5[d5[d]] : 0;
        for (int d = 0; d < 2; d++) T->q4[d][c] = c < 121 ? ac3_q11[d11[d]] : 0;
    }
    for (int c = 0; c < 128; c++)
        T->exp_lut[c] = (uint16_t)(c < 125 ? (c / 25) | (((c / 5) % 5) << 4) | ((c % 5) << 8) : 0x8000 | 2 | (2 << 4) | (2 << 8));
    for (int i = 0; i < 8; i++) T->q35[i] = ac3_q7[i];
    for (int i = 0; i < 16; i++) T->q35[8 + i] = ac3_q15[i];
    for (int i = 0; i < 256; i++) {
        T->dither_lut[i] = ac3_dither_lut[i];
        T->masktab[i] = ac3_masktab[i];
        T->latab[i] = ac3_latab[i];
    }
    for (int i = 0; i < 150; i++) T->hth[i] = ac3_hth[i];
    for (int i = 0; i < 64; i++) T->baptab[i] = ac3_baptab[i];
    for (int i = 0; i < 51; i++) T->bndtab[i] = ac3_bndtab[i];
    T->bndtab[51] = 253;
    for (int i = 0; i < 16; i++) T->bap_bits[i] = ac3_bap_bits[i];
    for (int b = 0; b < 16; b++) {
        T->cnt_lut32[b] = (b == 1 ? 1u : 0u) | (b == 2 ? 1u << 5 : 0u) | (b == 4 ? 1u << 10 : 0u) | (b == 0 ? 1u << 15 : 0u) |
                          ((b != 0 && b != 1 && b != 2 && b != 4) ? (uint32_t)ac3_bap_bits[b] << 20 : 0u);
    }
    // emit_lut: how pass 2 of the locate stage treats a mantissa of a given bap.
    //   classes: 0 = 3-level groups, 1 = 5-level, 2 = 11-level, 3 = plain fields, 4 = dithered zero
    //   x: increment of the packed per-class counters (bytes 0..3 = classes 0..3)
    //   y: [15:0] byte selector picking the class's 16-bit list base out of (base_lo, base_hi),
    //      [23:16] bits taken when the mantissa starts a field / group, [24] emit list entry + descriptor
    //   z: [15:0] byte selector keeping that base or replacing it by base_z, [31:16] 256 / period
    //   w: [15:0] byte selector picking the class's counter out of (run_a, run_z) (also used on the
    //      packed phases, where the zero class reads 0), [23:16] period, [31:24] increment of run_z
    for (int z = 0; z < 2; z++)
        for (int b = 0; b < 16; b++) {
            int cls = b == 0 ? 4 : b == 1 ? 0 : b == 2 ? 1 : b == 4 ? 2 : 3;
            uint32_t emit = (b != 0 || z) ? 1 : 0;
            uint32_t per = cls < 2 ? 3 : cls == 2 ? 2 : 1;
            uint32_t recip = cls < 2 ? 86 : cls == 2 ? 128 : 256;     // (x * recip) >> 8 == x / period, x < 128
            uint32_t width = cls == 0 ? 5 : cls <= 2 ? 7 : cls == 3 ? ac3_bap_bits[b] : 0;
            static const uint32_t selA[5] = {0x4410, 0x4432, 0x4454, 0x4476, 0x4410};
            uint32_t selB = cls == 4 ? 0x7654 : 0x7610;               // bytes 6, 7 of base_z are zero
            uint32_t selC = cls == 4 ? 0x7774 : (0x7770u | (uint32_t)cls);   // bytes 5..7 of run_z are zero
            uint32_t x = cls < 4 ? (1u << (8 * cls)) : 0;
            uint32_t incz = (cls == 4 && z) ? 1 : 0;
            T->emit_lut[z * 16 + b] = make_uint4(x, selA[cls] | (width << 16) | (emit << 24),
                                                 selB | (recip << 16), selC | (per << 16) | (incz << 24));
            // emit_lut2: the mantissa's place in the pair's plan (a52_decode.cu, locate stage)
            //   x: 2^17 / period (rounded up): (n * x) >> 17 == n / period for n < 2^15
            //   y: byte offset of the class's section; z: [7:0] entry stride, [15:8] offset of the first member
            //      word, [31:16] byte selector of the class's first group out of (ng1 << 16, ng1 + ng2)
            //   w: constant bits of the position word: class << 15, (32 - width) << 19, value table << 24
            //      (groups: offset from q1 in 64-byte units; plain: index into q35, bit 30 = use it), bit 31 = the
            //      entry has a position word (not for zeros)
            const uint32_t recip17 = per == 3 ? 43691u : per == 2 ? 65536u : 131072u;
            const uint32_t sec = cls < 3 ? 0u : cls == 3 ? (uint32_t)kPlanPlainOff : (uint32_t)kPlanZeroOff;
            const uint32_t stride = cls < 3 ? 16u : cls == 3 ? 8u : 4u, offa = cls < 4 ? 4u : 0u;
            static const uint32_t selG[5] = {0x7610, 0x7632, 0x7654, 0x7676, 0x7676};
            uint32_t pw = 0;
            if (cls < 3) {
                const uint32_t tbl = cls == 0 ? 0u : cls == 1 ? (uint32_t)sizeof(T->q1) / 64u
                                                              : (uint32_t)(sizeof(T->q1) + sizeof(T->q2)) / 64u;
                pw = 0x80000000u | ((uint32_t)cls << 15) | ((32u - width) << 19) | (tbl << 24);
            } else if (cls == 3) {
                pw = 0x80000000u | (3u << 15) | ((32u - width) << 19);
                if (b == 3 || b == 5) pw |= 0x40000000u | ((uint32_t)((b & 4) * 2) << 24);
            }
            T->emit_lut2[z * 16 + b] = make_uint4(recip17, sec, stride | (offa << 8) | (selG[cls] << 16), pw);
        }
    static_assert(offsetof(Tables, q2) == offsetof(Tables, q1) + sizeof(T->q1) &&
                  offsetof(Tables, q4) == offsetof(Tables, q2) + sizeof(T->q2), "value tables must be contiguous");
}

static int nout_of_flags(int flags)
{
    if (flags & A52_REQ_AS_CODED) return 6;
    int m = flags & M_MASK;
    if (m > M_DOLBY) m = M_STEREO;
    return h_nfchans[m] + ((flags & M_LFE) ? 1 : 0);
}

// frame length pre-pass for device-resident input: max over frames of the
// length a52_syncinfo derives from each header
__global__ void a52_maxlen_kernel(const uint8_t* es, const uint64_t* off, int nframes, int* out)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int len = 0;
    if (i < nframes) {
        const uint8_t* b = es + off[i];
        int cod = b[4] & 63, fscod = b[4] >> 6;
        if (b[0] == 0x0b && b[1] == 0x77 && cod < 38 && fscod < 3) {
            int kbps = c_bitrate[cod >> 1];
            len = fscod == 0 ? 4 * kbps : fscod == 2 ? 6 * kbps : 2 * (320 * kbps / 147 + (cod & 1));
        }
    }
    for (int o = 16; o; o >>= 1) len = max(len, __shfl_xor_sync(0xffffffffu, len, o));
    if ((threadIdx.x & 31) == 0 && len) atomicMax(out, len);
}

// Frame indexer on the device: one WARP walks one elementary stream with the resync discipline of
// a52dec.c:240-309 (slide one byte until a52_syncinfo accepts a header, then hop by the frame length).
// The walk is a chain p -> p + len(p); the warp follows it 32 links at a time by speculation: lane j tests the
// header at p + j * L (L = the length of the last accepted frame).  Every lane up to the first that does not find
// a valid header of the same length is exactly where the byte-serial walk would have landed, so those frames are
// accepted at once; at the first miss the warp falls back to what the reference does there - a different length
// is taken as it is, a bad header starts the sliding search, 32 byte positions per step (ballot, first hit wins).
// Pass 0 counts the frames of every stream, pass 1 writes their offsets behind stream_first[s].
__device__ __forceinline__ int dev_syncinfo_len(const uint8_t* b)
{
    if (b[0] != 0x0b || b[1] != 0x77) return 0;          // parse.c:98-127
    if ((b[5] >> 3) >= 12) return 0;
    const int cod = b[4] & 63, fscod = b[4] >> 6;
    if (cod >= 38 || fscod == 3) return 0;
    const int kbps = c_bitrate[cod >> 1];
    return fscod == 0 ? 4 * kbps : fscod == 2 ? 6 * kbps : 2 * (320 * kbps / 147 + (cod & 1));
}

__global__ void a52_index_kernel(const uint8_t* es, const uint64_t* stream_off, int nstreams, int* count,
                                 const uint32_t* stream_first, uint64_t* frame_off, int max_frames)
{
    const int s = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (s >= nstreams) return;
    uint64_t pos = stream_off[s];
    const uint64_t end = stream_off[s + 1];
    int n = 0;
    const uint32_t slot = stream_first ? stream_first[s] : 0;
    uint32_t L = 0;                                      // length of the last accepted frame (0: none yet)
    while (pos + 7 <= end) {
        if (L) {
            // speculative hop: frames of the same length back to back
            const uint64_t p = pos + (uint64_t)lane * L;
            const bool ok = p + 7 <= end && (uint32_t)dev_syncinfo_len(es + p) == L && p + L <= end;
            const uint32_t miss = ~__ballot_sync(0xffffffffu, ok);
            const int run = miss ? __ffs(miss) - 1 : 32;
            if (lane < run && frame_off && (int)(slot + n + lane) < max_frames) frame_off[slot + n + lane] = p;
            n += run;
            pos += (uint64_t)run * L;
            if (run == 32) continue;
            if (pos + 7 > end) break;
        }
        // the walk's next step, as the reference takes it: the header at pos decides
        const int len = dev_syncinfo_len(es + pos);      // (every lane reads the same 7 bytes)
        if (len) {
            if (pos + len > end) break;
            if (lane == 0 && frame_off && (int)(slot + n) < max_frames) frame_off[slot + n] = pos;
            n++;
            pos += len;
            L = len;
            continue;
        }
        // resync: first byte position after pos that holds a valid header
        L = 0;
        for (;;) {
            const uint64_t p = pos + 1 + lane;
            const bool hit = p + 7 <= end && dev_syncinfo_len(es + p) != 0;
            const uint32_t m = __ballot_sync(0xffffffffu, hit);
            if (m) { pos += (uint64_t)__ffs(m); break; }
            pos += 32;
            if (pos + 7 > end) break;
        }
    }
    if (count && lane == 0) count[s] = n;
}

// Frame-independent slices: position of the dither generator at the first frame every slice decodes (its
// look-back frame), from the per-frame draw counts of a scan pass.  One warp per stream.
__global__ void a52_slice_dither_kernel(const FrameScan* scan, const uint32_t* stream_first, int nstreams, int nslices,
                                        int slice_frames, const StreamCarry* carry_in, uint32_t* slice_dither)
{
    const int s = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (s >= nstreams) return;
    const uint32_t fs0 = stream_first[s], fs1 = stream_first[s + 1];
    uint32_t run = carry_in ? carry_in[s].dither_index % kDitherPeriod : 0u;       // position at the stream's first frame
    for (int j = 0; j < nslices; j++) {
        uint32_t f0 = fs0 + (uint32_t)j * slice_frames;
        if (f0 > fs1) f0 = fs1;
        const uint32_t f1 = (j + 1 < nslices && fs1 - f0 > (uint32_t)slice_frames) ? f0 + slice_frames : fs1;
        if (lane == 0)
            slice_dither[(size_t)s * nslices + j] =
                (j > 0 && f0 < fs1) ? (run + kDitherPeriod - scan[f0 - 1].dither_draws % kDitherPeriod) % kDitherPeriod : run;
        uint32_t sum = 0;
        for (uint32_t f = f0 + lane; f < f1; f += 32) sum += scan[f].dither_draws;
        for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        run = (run + sum) % kDitherPeriod;
    }
}

// longest stream (frames) of a device-resident batch
__global__ void a52_maxstream_kernel(const uint32_t* first, int nstreams, int* out)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int n = (i < nstreams) ? (int)(first[i + 1] - first[i]) : 0;
    for (int o = 16; o; o >>= 1) n = max(n, __shfl_xor_sync(0xffffffffu, n, o));
    if ((threadIdx.x & 31) == 0 && n) atomicMax(out, n);
}

}  // namespace a52

// ---------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------
struct a52_batch_s {
    int device = 0;
    int num_sms = 0;
    int warps_per_cta = 0;         // 0 = as many as fit
    int pair_kernel = 1;           // (kept for the launch arithmetic: two warps per stream)
    int max_frame_hint = 0;
    int max_stream_hint = 0;       // frames of the longest stream of the next batches (0 = derive)
    int slice_frames = 32;         // pair kernel: frames per work unit
    int slice_mode = 0;            // 0 = choose, 1 = slices of a stream chained by the carry record, 2 = frame-independent
    int lockstep = 2;              // frame gate; block and stage gates for streams of expensive blocks (A52_B200_LOCKSTEP=0 / 1: none / no stage gates)
    int zero_copy = 1;             // host-pointer calls: the kernel stores PCM straight into a pinned, mapped caller
                                   // buffer (A52_B200_ZERO_COPY=0: always stage in device memory and copy)
    const float* drc_table = nullptr;   // a52_batch_set_drc_table: ranges for the next A52_DRC_TABLE call
    uint16_t* d_dither = nullptr;
    int host_chunk_streams = 128;  // streams per pipelined chunk of a host-pointer call (A52_B200_HOST_CHUNK_STREAMS)
    int host_chunk_env = 0;
    int host_concurrency = 8;      // chunk kernels in flight at once (a chunk alone is latency-bound: its CTAs
                                   // are small, several launches share the SMs)
    cudaStream_t s_runs[8] = {};
    int* d_counter = nullptr;      // [0..31] work counters (one per pipelined chunk), [63] max frame length
    cudaStream_t s_in = nullptr, s_out = nullptr, s_run = nullptr;   // host-mode pipeline: H2D, D2H, kernels
    cudaEvent_t ev_in[32] = {nullptr}, ev_run[32] = {nullptr};
    char err[256] = {0};
    long launches = 0;
    // timing
    std::vector<cudaEvent_t> ev_pool;
    size_t ev_used = 0;
    // cached mode table key
    int mode_flags = -1;
    float mode_level = 0;
    a52::ModeEntry mode_tab[9 * 16];   // the request's grant / level table, copied into every launch's parameters
    // host-mode scratch
    struct Buf { void* p = nullptr; size_t cap = 0; } b_es, b_off, b_first, b_pcm, b_status, b_flags,
        b_carry, b_dexp, b_dbap, b_dcoef, b_dinfo, b_slice, b_done, b_snap, b_scan, b_sdith, b_spcm, b_cin, b_ranges;
};

#define A52_CUDA(call)                                                                       \
    do {                                                                                     \
        cudaError_t e_ = (call);                                                             \
        if (e_ != cudaSuccess) {                                                             \
            snprintf(ctx->err, sizeof(ctx->err), "%s failed: %s", #call, cudaGetErrorString(e_)); \
            return -1;                                                                       \
        }                                                                                    \
    } while (0)

static int ensure(a52_batch_t* ctx, a52_batch_s::Buf& b, size_t bytes)
{
    if (bytes <= b.cap) return 0;
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.cap = 0;
    size_t cap = bytes + bytes / 8 + 256;
    A52_CUDA(cudaMalloc(&b.p, cap));
    b.cap = cap;
    return 0;
}

#pragma GCC visibility push(default)
extern "C" {

a52_batch_t* a52_batch_create(int device)
{
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return nullptr;
    if (cudaSetDevice(device) != cudaSuccess) return nullptr;
    a52_batch_t* ctx = new (std::nothrow) a52_batch_s();
    if (!ctx) return nullptr;
    ctx->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete ctx; return nullptr; }
    ctx->num_sms = prop.multiProcessorCount;
    const char* g = getenv("A52_B200_WARPS_PER_CTA");
    if (g) ctx->warps_per_cta = atoi(g);
    const char* sf = getenv("A52_B200_SLICE_FRAMES");
    if (sf && atoi(sf) > 0) ctx->slice_frames = atoi(sf);
    const char* lk = getenv("A52_B200_LOCKSTEP");
    if (lk) ctx->lockstep = atoi(lk);
    const char* zc = getenv("A52_B200_ZERO_COPY");
    if (zc) ctx->zero_copy = atoi(zc);
    const char* sm = getenv("A52_B200_SLICE_MODE");
    if (sm && atoi(sm) >= 0 && atoi(sm) <= 2) ctx->slice_mode = atoi(sm);
    const char* hc = getenv("A52_B200_HOST_CHUNK_STREAMS");
    if (hc && atoi(hc) > 0) { ctx->host_chunk_streams = atoi(hc); ctx->host_chunk_env = 1; }
    const char* hq = getenv("A52_B200_HOST_CONCURRENCY");
    if (hq && atoi(hq) > 0 && atoi(hq) <= 8) ctx->host_concurrency = atoi(hq);

    // constant tables
    a52::Tables* T = new a52::Tables;
    a52::build_tables(T);
    bool ok = cudaMemcpyToSymbol(a52::g_tables, T, sizeof(*T)) == cudaSuccess;
    delete T;
    a52::MixEntry mix[8 * 11];
    a52::build_mix_table(mix);
    ok = ok && cudaMemcpyToSymbol(a52::c_mix, mix, sizeof(mix)) == cudaSuccess;
    // dither sequence: state after n calls from seed 1 (parse.c:310-319)
    std::vector<uint16_t> seq(a52::kDitherPeriod + a52::kDitherWrap);
    uint16_t s = 1;
    for (int n = 0; n < a52::kDitherPeriod; n++) {
        seq[n] = s;
        s = (uint16_t)(ac3_dither_lut[s >> 8] ^ (uint16_t)(s << 8));
    }
    for (int n = 0; n < a52::kDitherWrap; n++) seq[a52::kDitherPeriod + n] = seq[n];
    ok = ok && cudaMalloc(&ctx->d_dither, seq.size() * 2) == cudaSuccess;
    ok = ok && cudaMemcpy(ctx->d_dither, seq.data(), seq.size() * 2, cudaMemcpyHostToDevice) == cudaSuccess;
    ok = ok && cudaMalloc(&ctx->d_counter, 64 * sizeof(int)) == cudaSuccess;
    ok = ok && cudaFuncSetAttribute(a52::a52_decode_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    227 * 1024) == cudaSuccess;
    ok = ok && cudaFuncSetAttribute(a52::a52_decode_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    227 * 1024) == cudaSuccess;
    if (!ok) {
        a52_batch_destroy(ctx);
        return nullptr;
    }
    return ctx;
}

void a52_batch_destroy(a52_batch_t* ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    a52_batch_s::Buf* bufs[] = {&ctx->b_es, &ctx->b_off, &ctx->b_first, &ctx->b_pcm, &ctx->b_status,
                                &ctx->b_flags, &ctx->b_carry, &ctx->b_dexp, &ctx->b_dbap, &ctx->b_dcoef,
                                &ctx->b_dinfo, &ctx->b_slice, &ctx->b_done, &ctx->b_snap, &ctx->b_scan, &ctx->b_sdith,
                                &ctx->b_spcm, &ctx->b_cin, &ctx->b_ranges};
    for (auto* b : bufs)
        if (b->p) cudaFree(b->p);
    if (ctx->d_dither) cudaFree(ctx->d_dither);
    if (ctx->d_counter) cudaFree(ctx->d_counter);
    if (ctx->s_in) cudaStreamDestroy(ctx->s_in);
    if (ctx->s_out) cudaStreamDestroy(ctx->s_out);
    if (ctx->s_run) cudaStreamDestroy(ctx->s_run);
    for (int i = 0; i < 8; i++)
        if (ctx->s_runs[i]) cudaStreamDestroy(ctx->s_runs[i]);
    for (int i = 0; i < 32; i++) {
        if (ctx->ev_in[i]) cudaEventDestroy(ctx->ev_in[i]);
        if (ctx->ev_run[i]) cudaEventDestroy(ctx->ev_run[i]);
    }
    for (auto e : ctx->ev_pool) cudaEventDestroy(e);
    delete ctx;
}

const char* a52_batch_last_error(a52_batch_t* ctx) { return ctx ? ctx->err : "no context"; }

int a52_batch_index(const uint8_t* es, size_t es_bytes, uint64_t* frame_off, int max_frames)
{
    // same resync discipline as the reference driver (a52dec.c:240-309): slide one
    // byte at a time until a52_syncinfo accepts a header, then hop by the frame length
    size_t pos = 0;
    int n = 0;
    while (pos + 7 <= es_bytes && n < max_frames) {
        int fl, sr, br;
        int len = a52::host_syncinfo(es + pos, &fl, &sr, &br);
        if (!len) { pos++; continue; }
        if (pos + len > es_bytes) break;
        frame_off[n++] = pos;
        pos += len;
    }
    return n;
}

int a52_batch_index_device(a52_batch_t* ctx, const uint8_t* es, const uint64_t* stream_off, int nstreams,
                           uint64_t* frame_off, int max_frames, uint32_t* stream_first, void* cuda_stream)
{
    using namespace a52;
    if (!ctx || nstreams < 0) return -3;
    if (nstreams == 0) return 0;
    A52_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    if (ensure(ctx, ctx->b_off, (size_t)(nstreams + 1) * 8)) return -1;
    if (ensure(ctx, ctx->b_done, (size_t)nstreams * sizeof(int))) return -1;
    A52_CUDA(cudaMemcpyAsync(ctx->b_off.p, stream_off, (size_t)(nstreams + 1) * 8, cudaMemcpyHostToDevice, st));
    const int tpb = 128, grid = (nstreams + 3) / 4;      // one warp per stream
    a52_index_kernel<<<grid, tpb, 0, st>>>(es, (const uint64_t*)ctx->b_off.p, nstreams, (int*)ctx->b_done.p, nullptr,
                                          nullptr, 0);
    std::vector<int> cnt(nstreams);
    A52_CUDA(cudaMemcpyAsync(cnt.data(), ctx->b_done.p, (size_t)nstreams * sizeof(int), cudaMemcpyDeviceToHost, st));
    A52_CUDA(cudaStreamSynchronize(st));
    std::vector<uint32_t> first(nstreams + 1);
    uint64_t total = 0;
    for (int s = 0; s < nstreams; s++) { first[s] = (uint32_t)total; total += (uint64_t)cnt[s]; }
    first[nstreams] = (uint32_t)total;
    if (total > (uint64_t)max_frames) {
        snprintf(ctx->err, sizeof(ctx->err), "frame table too small: %llu frames found", (unsigned long long)total);
        return -4;
    }
    A52_CUDA(cudaMemcpyAsync(stream_first, first.data(), (size_t)(nstreams + 1) * 4, cudaMemcpyHostToDevice, st));
    a52_index_kernel<<<grid, tpb, 0, st>>>(es, (const uint64_t*)ctx->b_off.p, nstreams, nullptr, stream_first, frame_off,
                                          max_frames);
    A52_CUDA(cudaGetLastError());
    A52_CUDA(cudaStreamSynchronize(st));       // `first` lives on this stack frame
    ctx->launches += 2;
    return (int)total;
}

size_t a52_batch_frame_stride(int req_flags, int out_fmt)
{
    return (size_t)1536 * a52::nout_of_flags(req_flags) * (out_fmt >= A52_PCM_S16_INTERLEAVED ? 2 : 4);
}

void a52_batch_set_max_frame_bytes(a52_batch_t* ctx, int nbytes) { ctx->max_frame_hint = nbytes; }

void a52_batch_set_max_stream_frames(a52_batch_t* ctx, int nframes) { ctx->max_stream_hint = nframes; }

long a52_batch_launch_count(a52_batch_t* ctx) { return ctx->launches; }

double a52_batch_kernel_ms(a52_batch_t* ctx, int* nlaunches)
{
    double total = 0;
    int n = 0;
    cudaSetDevice(ctx->device);
    for (size_t i = 0; i + 1 < ctx->ev_used; i += 2) {
        float ms = 0;
        cudaEventSynchronize(ctx->ev_pool[i + 1]);
        if (cudaEventElapsedTime(&ms, ctx->ev_pool[i], ctx->ev_pool[i + 1]) == cudaSuccess) {
            total += ms;
            n++;
        }
    }
    ctx->ev_used = 0;
    if (nlaunches) *nlaunches = n;
    return n ? total / n : 0.0;
}

// scratch_base / scratch_total: position of this launch's streams inside the per-stream scratch arrays when
// several launches of one call are in flight at once (host pipeline); 0 / 0 = a launch on its own
enum { RUN_CHAINED = 0, RUN_SCAN = 1, RUN_INDEP = 2 };

static int launch_decode(a52_batch_t* ctx, a52::DecodeParams& P, int nframes, int max_frame_bytes,
                         float level, cudaStream_t st, int counter_slot = 0, int max_stream_frames = 0,
                         int scratch_base = 0, int scratch_total = 0, int snap_streams = 0, int run_mode = RUN_CHAINED)
{
    using namespace a52;
    // work units of the pair kernel: slices of streams (see a52_decode_kernel)
    P.slice_frames = ctx->slice_frames;
    P.nslices = 1;
    P.carry_init = P.carry != nullptr;
    P.slice_done = nullptr;
    P.lockstep = ctx->lockstep;
    P.scan_only = run_mode == RUN_SCAN;
    P.indep = run_mode == RUN_INDEP;
    if (run_mode != RUN_CHAINED) {
        if (max_stream_frames > ctx->slice_frames) P.nslices = (max_stream_frames + ctx->slice_frames - 1) / ctx->slice_frames;
    } else if (ctx->pair_kernel && max_stream_frames > ctx->slice_frames) {
        P.nslices = (max_stream_frames + ctx->slice_frames - 1) / ctx->slice_frames;
        const size_t total = scratch_total ? (size_t)scratch_total : (size_t)P.nstreams;
        if (ensure(ctx, ctx->b_done, total * sizeof(int))) return -1;
        A52_CUDA(cudaMemsetAsync((int*)ctx->b_done.p + scratch_base, 0, (size_t)P.nstreams * sizeof(int), st));
        P.slice_done = (int*)ctx->b_done.p + scratch_base;
        if (!P.carry) {
            if (ensure(ctx, ctx->b_slice, total * sizeof(StreamCarry))) return -1;
            P.carry = (StreamCarry*)ctx->b_slice.p + scratch_base;
        }
    }
    // per-request constants
    if (ctx->mode_flags != P.req_flags || ctx->mode_level != level) {
        build_mode_table(ctx->mode_tab, P.req_flags, level);
        ctx->mode_flags = P.req_flags;
        ctx->mode_level = level;
    }
    memcpy(P.mode, ctx->mode_tab, sizeof(P.mode));
    if (max_frame_bytes < 128) max_frame_bytes = 128;
    if (max_frame_bytes > 3840) max_frame_bytes = 3840;
    P.fbuf_bytes = align16(max_frame_bytes + 15) + 16 + 16;
    P.nplanes = (P.req_flags & M_LFE) ? 6 : 5;
    const bool pair = ctx->pair_kernel != 0;
    P.group_threads = pair ? 64 : 32;
    P.warp_bytes = pair_smem_bytes(P.fbuf_bytes, P.nplanes);
    P.dither_seq = ctx->d_dither;
    P.work_counter = ctx->d_counter + counter_slot;
    const int tables = kTablesBytes;
    int fit = (227 * 1024 - tables) / P.warp_bytes;
    const int fit_max = kMaxPairsPerCta;
    if (fit > fit_max) fit = fit_max;
    if (fit < 1) {
        snprintf(ctx->err, sizeof(ctx->err), "decode kernel does not fit: %d bytes of shared memory per stream", P.warp_bytes);
        return -2;
    }
    int G = fit;
    if (ctx->warps_per_cta > 0 && ctx->warps_per_cta < G) G = ctx->warps_per_cta;
    // small batches: spread the work over all SMs.  What runs side by side: whole streams when their slices are
    // chained, every slice when they are independent of each other
    const long long par = run_mode == RUN_CHAINED ? (long long)P.nstreams : (long long)P.nstreams * P.nslices;
    long long per_sm = (par + ctx->num_sms - 1) / ctx->num_sms;
    if (per_sm < G) G = per_sm < 1 ? 1 : (int)per_sm;
    const int threads = G * P.group_threads;
    const size_t smem = (size_t)tables + (size_t)G * P.warp_bytes;
    long long grid_ll = (par + G - 1) / G;
    int grid = grid_ll > ctx->num_sms ? ctx->num_sms : (int)grid_ll;
    if (grid > ctx->num_sms) grid = ctx->num_sms;
    if (grid < 1) grid = 1;
    A52_CUDA(cudaMemsetAsync(ctx->d_counter + counter_slot, 0, sizeof(int), st));
    // scratch of the locate stage: one plan per resident pair.  Launches of one host-pointer call run side by
    // side: each takes the region of its counter slot, all regions sized for the largest chunk.
    {
        const long long nmax = scratch_total ? snap_streams : par;
        size_t pairs = (size_t)nmax + ctx->num_sms;
        const size_t cap = (size_t)ctx->num_sms * fit;
        if (pairs > cap) pairs = cap;
        const size_t region = pairs * (size_t)kPlanBytes;
        const int regions = scratch_total ? 32 : 1;
        if (ensure(ctx, ctx->b_snap, region * regions)) return -1;
        P.plan = (uint8_t*)ctx->b_snap.p + region * (scratch_total ? counter_slot : 0);
        if (run_mode == RUN_INDEP) {
            // where the look-back frame of every slice puts its PCM
            if (ensure(ctx, ctx->b_spcm, pairs * P.frame_stride)) return -1;
            P.scratch_pcm = (uint8_t*)ctx->b_spcm.p;
        }
    }
    // timing events
    if (ctx->ev_used + 2 > ctx->ev_pool.size()) {
        if (ctx->ev_pool.size() < 8192) {
            cudaEvent_t a, b;
            A52_CUDA(cudaEventCreate(&a));
            A52_CUDA(cudaEventCreate(&b));
            ctx->ev_pool.push_back(a);
            ctx->ev_pool.push_back(b);
        } else {
            ctx->ev_used = 0;
        }
    }
    cudaEvent_t e0 = ctx->ev_pool[ctx->ev_used], e1 = ctx->ev_pool[ctx->ev_used + 1];
    ctx->ev_used += 2;
    A52_CUDA(cudaEventRecord(e0, st));
    if (P.nplanes == 6) a52_decode_kernel<6><<<grid, threads, smem, st>>>(P);
    else a52_decode_kernel<5><<<grid, threads, smem, st>>>(P);
    A52_CUDA(cudaEventRecord(e1, st));
    A52_CUDA(cudaGetLastError());
    ctx->launches++;
    (void)nframes;
    return 0;
}

// Frame-independent slices of device-resident streams: scan pass, prefix sum of the dither draws, decode.
static int launch_indep(a52_batch_t* ctx, a52::DecodeParams& P, int nframes, int maxlen, int maxstream, float level,
                        cudaStream_t st)
{
    using namespace a52;
    const int nstreams = P.nstreams;
    if (ensure(ctx, ctx->b_scan, (size_t)nframes * sizeof(FrameScan))) return -1;
    const int nslices = (maxstream + ctx->slice_frames - 1) / ctx->slice_frames;
    if (ensure(ctx, ctx->b_sdith, (size_t)nstreams * nslices * sizeof(uint32_t))) return -1;
    DecodeParams Ps = P;
    Ps.pcm = nullptr; Ps.status = nullptr; Ps.frame_flags = nullptr; Ps.carry = nullptr;
    Ps.dbg_exp = nullptr; Ps.dbg_bap = nullptr; Ps.dbg_coef = nullptr; Ps.dbg_info = nullptr;
    Ps.scan = (FrameScan*)ctx->b_scan.p;
    int rc = launch_decode(ctx, Ps, nframes, maxlen, level, st, 0, maxstream, 0, 0, 0, RUN_SCAN);
    if (rc) return rc;
    const StreamCarry* cin = nullptr;
    if (P.carry) {
        if (ensure(ctx, ctx->b_cin, (size_t)nstreams * sizeof(StreamCarry))) return -1;
        A52_CUDA(cudaMemcpyAsync(ctx->b_cin.p, P.carry, (size_t)nstreams * sizeof(StreamCarry), cudaMemcpyDeviceToDevice, st));
        cin = (const StreamCarry*)ctx->b_cin.p;
    }
    a52_slice_dither_kernel<<<(nstreams + 3) / 4, 128, 0, st>>>((const FrameScan*)ctx->b_scan.p, P.stream_first, nstreams,
                                                               nslices, ctx->slice_frames, cin, (uint32_t*)ctx->b_sdith.p);
    A52_CUDA(cudaGetLastError());
    ctx->launches++;
    P.carry_in = cin;
    P.slice_dither = (const uint32_t*)ctx->b_sdith.p;
    return launch_decode(ctx, P, nframes, maxlen, level, st, 0, maxstream, 0, 0, 0, RUN_INDEP);
}

int a52_batch_violations(void)
{
    int v[4] = {0, 0, 0, 0};
    if (cudaMemcpyFromSymbol(v, a52::g_violation, sizeof(v)) != cudaSuccess) return -1;
#ifdef A52_BOUNDS_CHECK
    return v[0];
#else
    return v[0] ? v[0] : -2;          // -2: this build carries no checks
#endif
}

void a52_batch_set_slice_mode(a52_batch_t* ctx, int mode)
{
    if (mode >= 0 && mode <= 2) ctx->slice_mode = mode;
}

void a52_batch_set_drc_table(a52_batch_t* ctx, const float* ranges) { ctx->drc_table = ranges; }

int a52_batch_scan(a52_batch_t* ctx, const uint8_t* es, size_t es_bytes, const uint64_t* frame_off, int nframes,
                   const uint32_t* stream_first, int nstreams, int req_flags, a52_frame_scan_t* scan, int mem_flags,
                   void* cuda_stream)
{
    using namespace a52;
    static_assert(sizeof(a52_frame_scan_t) == sizeof(FrameScan), "scan record layout");
    if (!ctx) return -1;
    ctx->err[0] = 0;
    if (nframes < 0 || nstreams < 0 || (req_flags & M_MASK) > M_DOLBY || !scan) {
        snprintf(ctx->err, sizeof(ctx->err), "bad argument");
        return -3;
    }
    if (nframes == 0 || nstreams == 0) return 0;
    A52_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    DecodeParams P;
    memset(&P, 0, sizeof(P));
    P.es_bytes = es_bytes;
    P.nstreams = nstreams;
    P.nframes = nframes;
    P.req_flags = req_flags;
    P.out_fmt = A52_PCM_F32_PLANAR;
    P.nout_req = nout_of_flags(req_flags);
    P.frame_stride = a52_batch_frame_stride(req_flags, A52_PCM_F32_PLANAR);
    int maxlen = 0, maxstream = 0;
    if (mem_flags & A52_BATCH_DEVICE_PTRS) {
        if (((uintptr_t)es & 15) != 0) {
            snprintf(ctx->err, sizeof(ctx->err), "device bitstream pointer must be 16-byte aligned");
            return -3;
        }
        P.es = es;
        P.frame_off = frame_off;
        P.stream_first = stream_first;
        P.scan = (FrameScan*)scan;
        maxstream = ctx->max_stream_hint;
        if (maxstream <= 0) {
            A52_CUDA(cudaMemsetAsync(ctx->d_counter + 62, 0, sizeof(int), st));
            a52_maxstream_kernel<<<(nstreams + 255) / 256, 256, 0, st>>>(stream_first, nstreams, ctx->d_counter + 62);
            A52_CUDA(cudaMemcpyAsync(&maxstream, ctx->d_counter + 62, sizeof(int), cudaMemcpyDeviceToHost, st));
            A52_CUDA(cudaStreamSynchronize(st));
            ctx->launches++;
        }
        maxlen = ctx->max_frame_hint;
        if (maxlen <= 0) {
            A52_CUDA(cudaMemsetAsync(ctx->d_counter + 63, 0, sizeof(int), st));
            a52_maxlen_kernel<<<(nframes + 255) / 256, 256, 0, st>>>(es, frame_off, nframes, ctx->d_counter + 63);
            A52_CUDA(cudaMemcpyAsync(&maxlen, ctx->d_counter + 63, sizeof(int), cudaMemcpyDeviceToHost, st));
            A52_CUDA(cudaStreamSynchronize(st));
            ctx->launches++;
        }
        A52_CUDA(cudaMemsetAsync(scan, 0xff, (size_t)nframes * sizeof(FrameScan), st));
        return launch_decode(ctx, P, nframes, maxlen, 1.0f, st, 0, maxstream, 0, 0, 0, RUN_SCAN);
    }
    // host pointers: stage in, scan, stage out (synchronous)
    for (int i = 0; i < nframes; i++) {
        int fl, sr, br;
        if (frame_off[i] + 7 <= es_bytes) {
            int len = host_syncinfo(es + frame_off[i], &fl, &sr, &br);
            if (len > maxlen) maxlen = len;
        }
    }
    for (int q = 0; q < nstreams; q++) {
        int n = (int)(stream_first[q + 1] - stream_first[q]);
        if (n > maxstream) maxstream = n;
    }
    if (ensure(ctx, ctx->b_es, es_bytes + 64)) return -1;
    if (ensure(ctx, ctx->b_off, (size_t)(nframes + 1) * 8)) return -1;
    if (ensure(ctx, ctx->b_first, (size_t)(nstreams + 1) * 4)) return -1;
    if (ensure(ctx, ctx->b_scan, (size_t)nframes * sizeof(FrameScan))) return -1;
    A52_CUDA(cudaMemsetAsync((uint8_t*)ctx->b_es.p + es_bytes, 0, 64, st));
    A52_CUDA(cudaMemcpyAsync(ctx->b_es.p, es, es_bytes, cudaMemcpyHostToDevice, st));
    A52_CUDA(cudaMemcpyAsync(ctx->b_off.p, frame_off, (size_t)nframes * 8, cudaMemcpyHostToDevice, st));
    A52_CUDA(cudaMemcpyAsync(ctx->b_first.p, stream_first, (size_t)(nstreams + 1) * 4, cudaMemcpyHostToDevice, st));
    A52_CUDA(cudaMemsetAsync(ctx->b_scan.p, 0xff, (size_t)nframes * sizeof(FrameScan), st));
    P.es = (const uint8_t*)ctx->b_es.p;
    P.frame_off = (const uint64_t*)ctx->b_off.p;
    P.stream_first = (const uint32_t*)ctx->b_first.p;
    P.scan = (FrameScan*)ctx->b_scan.p;
    int rc = launch_decode(ctx, P, nframes, maxlen, 1.0f, st, 0, maxstream, 0, 0, 0, RUN_SCAN);
    if (rc) { cudaStreamSynchronize(st); return rc; }
    A52_CUDA(cudaMemcpyAsync(scan, ctx->b_scan.p, (size_t)nframes * sizeof(FrameScan), cudaMemcpyDeviceToHost, st));
    A52_CUDA(cudaStreamSynchronize(st));
    return 0;
}

static int batch_decode_impl(a52_batch_t* ctx, const uint8_t* es, size_t es_bytes, const uint64_t* frame_off,
                             int nframes, const uint32_t* stream_first, int nstreams, int req_flags, float level,
                             float bias, int drc_mode, int out_fmt, void* pcm_out, int32_t* frame_status,
                             int32_t* frame_flags, a52_stream_carry_t* carry, const a52_batch_debug_t* debug,
                             int mem_flags, void* cuda_stream);

int a52_batch_decode(a52_batch_t* ctx, const uint8_t* es, size_t es_bytes, const uint64_t* frame_off,
                     int nframes, const uint32_t* stream_first, int nstreams, int req_flags, float level,
                     float bias, int drc_mode, int out_fmt, void* pcm_out, int32_t* frame_status,
                     int32_t* frame_flags, a52_stream_carry_t* carry, const a52_batch_debug_t* debug,
                     int mem_flags, void* cuda_stream)
{
    if (!ctx) return -1;
    const int rc = batch_decode_impl(ctx, es, es_bytes, frame_off, nframes, stream_first, nstreams, req_flags, level, bias,
                                     drc_mode, out_fmt, pcm_out, frame_status, frame_flags, carry, debug, mem_flags,
                                     cuda_stream);
    if (rc && !(mem_flags & A52_BATCH_DEVICE_PTRS)) {
        // a host-pointer call that fails half way may have copies in flight from / into the caller's buffers:
        // let them settle before the caller gets its buffers back
        if (ctx->s_in) cudaStreamSynchronize(ctx->s_in);
        if (ctx->s_out) cudaStreamSynchronize(ctx->s_out);
        if (ctx->s_run) cudaStreamSynchronize(ctx->s_run);
        for (int i = 0; i < 8; i++)
            if (ctx->s_runs[i]) cudaStreamSynchronize(ctx->s_runs[i]);
        if (cuda_stream) cudaStreamSynchronize((cudaStream_t)cuda_stream);
        cudaGetLastError();
    }
    return rc;
}

static int batch_decode_impl(a52_batch_t* ctx, const uint8_t* es, size_t es_bytes, const uint64_t* frame_off,
                             int nframes, const uint32_t* stream_first, int nstreams, int req_flags, float level,
                             float bias, int drc_mode, int out_fmt, void* pcm_out, int32_t* frame_status,
                             int32_t* frame_flags, a52_stream_carry_t* carry, const a52_batch_debug_t* debug,
                             int mem_flags, void* cuda_stream)
{
    using namespace a52;
    ctx->err[0] = 0;
    if (nframes < 0 || nstreams < 0 || (req_flags & M_MASK) > M_DOLBY || out_fmt < 0 || out_fmt > A52_PCM_S16_WAV) {
        snprintf(ctx->err, sizeof(ctx->err), "bad argument");
        return -3;
    }
    if (nframes == 0 || nstreams == 0) return 0;
    A52_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const size_t stride = a52_batch_frame_stride(req_flags, out_fmt);

    DecodeParams P;
    memset(&P, 0, sizeof(P));
    P.es_bytes = es_bytes;
    P.nstreams = nstreams;
    P.nframes = nframes;
    P.req_flags = req_flags;
    P.bias = bias;
    P.drc_off = (drc_mode == A52_DRC_OFF);
    P.out_fmt = out_fmt;
    P.nout_req = nout_of_flags(req_flags);
    P.frame_stride = stride;

    if (mem_flags & A52_BATCH_DEVICE_PTRS) {
        if (((uintptr_t)es & 15) != 0) {
            snprintf(ctx->err, sizeof(ctx->err), "device bitstream pointer must be 16-byte aligned");
            return -3;
        }
        P.es = es;
        P.frame_off = frame_off;
        P.stream_first = stream_first;
        P.pcm = (uint8_t*)pcm_out;
        P.status = frame_status;
        P.frame_flags = frame_flags;
        P.carry = (StreamCarry*)carry;
        if (debug) { P.dbg_exp = debug->exp; P.dbg_bap = debug->bap; P.dbg_coef = debug->coef; P.dbg_info = debug->info; }
        int maxstream = ctx->max_stream_hint;
        if (maxstream <= 0 && ctx->pair_kernel) {
            A52_CUDA(cudaMemsetAsync(ctx->d_counter + 62, 0, sizeof(int), st));
            a52_maxstream_kernel<<<(nstreams + 255) / 256, 256, 0, st>>>(stream_first, nstreams, ctx->d_counter + 62);
            A52_CUDA(cudaMemcpyAsync(&maxstream, ctx->d_counter + 62, sizeof(int), cudaMemcpyDeviceToHost, st));
            A52_CUDA(cudaStreamSynchronize(st));
            ctx->launches++;
        }
        int maxlen = ctx->max_frame_hint;
        if (maxlen <= 0) {
            A52_CUDA(cudaMemsetAsync(ctx->d_counter + 63, 0, sizeof(int), st));
            a52_maxlen_kernel<<<(nframes + 255) / 256, 256, 0, st>>>(es, frame_off, nframes, ctx->d_counter + 63);
            A52_CUDA(cudaMemcpyAsync(&maxlen, ctx->d_counter + 63, sizeof(int), cudaMemcpyDeviceToHost, st));
            A52_CUDA(cudaStreamSynchronize(st));
            ctx->launches++;
        }
        if (drc_mode == A52_DRC_TABLE) P.drc_ranges = ctx->drc_table;        // (device pointer in this mode)
        // few long streams: chained slices would leave most of the GPU idle (a stream is one pair at a time), so
        // the slices are made independent of each other: a scan pass counts what the dither generator draws in
        // every frame, a prefix sum turns that into every slice's starting position, and every slice rebuilds the
        // overlap-add state by decoding one frame of look-back
        const long long resident = (long long)ctx->num_sms * kMaxPairsPerCta;
        const bool indep = ctx->slice_mode == 2 ||
                           (ctx->slice_mode == 0 && 2LL * nstreams <= resident && maxstream >= 4 * ctx->slice_frames);
        if (!indep || maxstream <= ctx->slice_frames) return launch_decode(ctx, P, nframes, maxlen, level, st, 0, maxstream);
        return launch_indep(ctx, P, nframes, maxlen, maxstream, level, st);
    }

    // ---- host pointers: stage in, decode, stage out (synchronous for the caller) ----
    // Large batches are cut into chunks of streams that flow through three CUDA streams:
    // bitstream H2D | decode kernel | PCM D2H, so that with pinned host buffers the PCIe copies
    // overlap each other and the kernels.  (Pageable buffers work too, without the overlap.)
    if (!ctx->s_in) {
        A52_CUDA(cudaStreamCreateWithFlags(&ctx->s_in, cudaStreamNonBlocking));
        A52_CUDA(cudaStreamCreateWithFlags(&ctx->s_out, cudaStreamNonBlocking));
        A52_CUDA(cudaStreamCreateWithFlags(&ctx->s_run, cudaStreamNonBlocking));
        for (int i = 0; i < 8; i++) A52_CUDA(cudaStreamCreateWithFlags(&ctx->s_runs[i], cudaStreamNonBlocking));
        for (int i = 0; i < 32; i++) {
            A52_CUDA(cudaEventCreateWithFlags(&ctx->ev_in[i], cudaEventDisableTiming));
            A52_CUDA(cudaEventCreateWithFlags(&ctx->ev_run[i], cudaEventDisableTiming));
        }
    }
    cudaStream_t s_in = ctx->s_in, s_out = ctx->s_out, s_run = st ? st : ctx->s_run;
    // longest frame of the batch: the caller's bound when there is one (a52_batch_set_max_frame_bytes; a frame
    // longer than that is refused by the kernel, not truncated), else one pass over the headers - a cache miss per
    // frame on a large batch, all of it before the first byte moves
    int maxlen = ctx->max_frame_hint;
    if (maxlen <= 0)
        for (int i = 0; i < nframes; i++) {
            int fl, sr, br;
            if (frame_off[i] + 7 <= es_bytes) {
                int len = host_syncinfo(es + frame_off[i], &fl, &sr, &br);
                if (len > maxlen) maxlen = len;
            }
        }
    if (ensure(ctx, ctx->b_es, es_bytes + 64)) return -1;
    if (ensure(ctx, ctx->b_off, (size_t)(nframes + 1) * 8)) return -1;
    if (ensure(ctx, ctx->b_first, (size_t)(nstreams + 1) * 4)) return -1;
    // A pinned (page-locked, hence mapped) caller buffer takes the PCM straight from the kernel's stores: no staging
    // buffer, no device-to-host copy pass, and the transfer overlaps the decode at the granularity of a block.
    // (Small batches only - the drop-in a52_block path, short files: there it saves a copy and a synchronisation.
    // Large ones go through the copy engine, which moves 4 % more per second than the SMs' stores do over the same
    // link: 301.6 against 311.5 ms for the 18 GB of bench.py's step; A52_B200_ZERO_COPY=2 forces it regardless.)
    uint8_t* pcm_alias = nullptr;
    if (ctx->zero_copy == 2 || (ctx->zero_copy && stride * (size_t)nframes <= ((size_t)64 << 20))) {
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, pcm_out) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer)
            pcm_alias = (uint8_t*)at.devicePointer;
        else
            cudaGetLastError();
    }
    if (!pcm_alias && ensure(ctx, ctx->b_pcm, stride * nframes)) return -1;
    if (ensure(ctx, ctx->b_status, (size_t)nframes * 4)) return -1;
    if (ensure(ctx, ctx->b_flags, (size_t)nframes * 4)) return -1;
    A52_CUDA(cudaMemsetAsync((uint8_t*)ctx->b_es.p + es_bytes, 0, 64, s_run));
    std::vector<uint64_t> off(frame_off, frame_off + nframes);
    off.push_back(es_bytes);
    A52_CUDA(cudaMemcpyAsync(ctx->b_off.p, off.data(), off.size() * 8, cudaMemcpyHostToDevice, s_run));
    A52_CUDA(cudaMemcpyAsync(ctx->b_first.p, stream_first, (size_t)(nstreams + 1) * 4, cudaMemcpyHostToDevice, s_run));
    P.es = (const uint8_t*)ctx->b_es.p;
    P.frame_off = (const uint64_t*)ctx->b_off.p;
    P.pcm = pcm_alias ? pcm_alias : (uint8_t*)ctx->b_pcm.p;
    P.status = (int32_t*)ctx->b_status.p;
    P.frame_flags = (int32_t*)ctx->b_flags.p;
    if (carry) {
        if (ensure(ctx, ctx->b_carry, sizeof(StreamCarry) * (size_t)nstreams)) return -1;
        A52_CUDA(cudaMemcpyAsync(ctx->b_carry.p, carry, sizeof(StreamCarry) * (size_t)nstreams, cudaMemcpyHostToDevice, s_run));
    }
    const size_t nblk = (size_t)nframes * 6;
    if (drc_mode == A52_DRC_TABLE && ctx->drc_table) {
        if (ensure(ctx, ctx->b_ranges, nblk * 2 * sizeof(float))) return -1;
        A52_CUDA(cudaMemcpyAsync(ctx->b_ranges.p, ctx->drc_table, nblk * 2 * sizeof(float), cudaMemcpyHostToDevice, s_run));
        P.drc_ranges = (const float*)ctx->b_ranges.p;
    }
    if (debug) {
        if (debug->exp && debug->bap) {
            if (ensure(ctx, ctx->b_dexp, nblk * 7 * 256) || ensure(ctx, ctx->b_dbap, nblk * 7 * 256)) return -1;
            P.dbg_exp = (uint8_t*)ctx->b_dexp.p;
            P.dbg_bap = (uint8_t*)ctx->b_dbap.p;
            A52_CUDA(cudaMemsetAsync(P.dbg_exp, 0, nblk * 7 * 256, s_run));
            A52_CUDA(cudaMemsetAsync(P.dbg_bap, 0, nblk * 7 * 256, s_run));
        }
        if (debug->coef) {
            if (ensure(ctx, ctx->b_dcoef, nblk * 6 * 256 * 4)) return -1;
            P.dbg_coef = (float*)ctx->b_dcoef.p;
            A52_CUDA(cudaMemsetAsync(P.dbg_coef, 0, nblk * 6 * 256 * 4, s_run));
        }
        if (debug->info) {
            if (ensure(ctx, ctx->b_dinfo, nblk * 16 * 4)) return -1;
            P.dbg_info = (int32_t*)ctx->b_dinfo.p;
            A52_CUDA(cudaMemsetAsync(P.dbg_info, 0, nblk * 16 * 4, s_run));
        }
    }
    // the small staging copies above read pageable host vectors: finish them before they go away
    A52_CUDA(cudaStreamSynchronize(s_run));

    // chunk plan: contiguous stream ranges; each chunk's bitstream is the byte span of its frames
    // chunk size: small chunks start the PCM copy early (what counts when the copy is the bottleneck: float
    // output), larger ones keep more streams in flight (what counts when the kernels are: int16 output)
    const int chunk_streams = ctx->host_chunk_env ? ctx->host_chunk_streams : (stride >= 8192 ? 128 : 256);
    int nchunks = nstreams / chunk_streams;
    if (nchunks > 32) nchunks = 32;
    if (nchunks < 1) nchunks = 1;
    int longest = 0;
    for (int q = 0; q < nstreams; q++) {
        int n = (int)(stream_first[q + 1] - stream_first[q]);
        if (n > longest) longest = n;
    }
    // few long streams: frame-independent slices (see the device-pointer path), the whole batch in one go
    const bool indep = longest > ctx->slice_frames &&
                       (ctx->slice_mode == 2 || (ctx->slice_mode == 0 && 2LL * nstreams <= (long long)ctx->num_sms * kMaxPairsPerCta &&
                                                 longest >= 4 * ctx->slice_frames));
    if (indep) nchunks = 1;
    std::vector<size_t> cb0(nchunks), cb1(nchunks);
    size_t span_total = 0;
    for (int cidx = 0; cidx < nchunks; cidx++) {
        int s0 = (int)((long long)cidx * nstreams / nchunks), s1 = (int)((long long)(cidx + 1) * nstreams / nchunks);
        size_t lo = es_bytes, hi = 0;
        for (uint32_t f = stream_first[s0]; f < stream_first[s1]; f++) {
            size_t a = (size_t)frame_off[f];
            size_t b = a + (size_t)maxlen;
            if (a < lo) lo = a;
            if (b > hi) hi = b;
        }
        if (hi > es_bytes) hi = es_bytes;
        if (lo > hi) lo = hi;
        cb0[cidx] = lo & ~(size_t)15;
        cb1[cidx] = hi;
        span_total += cb1[cidx] - cb0[cidx];
    }
    const bool slice_input = span_total <= es_bytes + es_bytes / 2 + 4096;
    if (!slice_input) {
        // frames of different chunks interleave in memory: one copy of everything
        A52_CUDA(cudaMemcpyAsync(ctx->b_es.p, es, es_bytes, cudaMemcpyHostToDevice, s_in));
        A52_CUDA(cudaEventRecord(ctx->ev_in[0], s_in));
    }
    for (int cidx = 0; cidx < nchunks; cidx++) {
        int s0 = (int)((long long)cidx * nstreams / nchunks), s1 = (int)((long long)(cidx + 1) * nstreams / nchunks);
        uint32_t fa = stream_first[s0], fb = stream_first[s1];
        if (s1 == s0) continue;
        // chunk kernels go round the run streams: a caller-given stream keeps them in order on itself
        cudaStream_t s_k = st ? st : ctx->s_runs[cidx % ctx->host_concurrency];
        if (slice_input) {
            if (cb1[cidx] > cb0[cidx])
                A52_CUDA(cudaMemcpyAsync((uint8_t*)ctx->b_es.p + cb0[cidx], es + cb0[cidx], cb1[cidx] - cb0[cidx],
                                         cudaMemcpyHostToDevice, s_in));
            A52_CUDA(cudaEventRecord(ctx->ev_in[cidx], s_in));
            A52_CUDA(cudaStreamWaitEvent(s_k, ctx->ev_in[cidx], 0));
        } else {
            A52_CUDA(cudaStreamWaitEvent(s_k, ctx->ev_in[0], 0));
        }
        DecodeParams Pc = P;
        Pc.stream_first = (const uint32_t*)ctx->b_first.p + s0;
        Pc.nstreams = s1 - s0;
        Pc.carry = carry ? (StreamCarry*)ctx->b_carry.p + s0 : nullptr;
        int chunk_max = 0;
        for (int q = s0; q < s1; q++) {
            int n = (int)(stream_first[q + 1] - stream_first[q]);
            if (n > chunk_max) chunk_max = n;
        }
        int rc = indep ? launch_indep(ctx, Pc, nframes, maxlen, chunk_max, level, s_k)
                       : launch_decode(ctx, Pc, nframes, maxlen, level, s_k, cidx, chunk_max, s0, nstreams,
                                       (nstreams + nchunks - 1) / nchunks + 1);
        if (rc) return rc;
        A52_CUDA(cudaEventRecord(ctx->ev_run[cidx], s_k));
        A52_CUDA(cudaStreamWaitEvent(s_out, ctx->ev_run[cidx], 0));
        if (fb > fa && !pcm_alias)
            A52_CUDA(cudaMemcpyAsync((uint8_t*)pcm_out + (size_t)fa * stride, Pc.pcm + (size_t)fa * stride,
                                     (size_t)(fb - fa) * stride, cudaMemcpyDeviceToHost, s_out));
    }
    if (!st)
        for (int i = 0; i < ctx->host_concurrency; i++) A52_CUDA(cudaStreamSynchronize(ctx->s_runs[i]));
    A52_CUDA(cudaStreamSynchronize(s_run));
    if (frame_status) A52_CUDA(cudaMemcpyAsync(frame_status, P.status, (size_t)nframes * 4, cudaMemcpyDeviceToHost, s_run));
    if (frame_flags) A52_CUDA(cudaMemcpyAsync(frame_flags, P.frame_flags, (size_t)nframes * 4, cudaMemcpyDeviceToHost, s_run));
    if (carry) A52_CUDA(cudaMemcpyAsync(carry, ctx->b_carry.p, sizeof(StreamCarry) * (size_t)nstreams, cudaMemcpyDeviceToHost, s_run));
    if (debug) {
        if (P.dbg_exp) {
            A52_CUDA(cudaMemcpyAsync(debug->exp, P.dbg_exp, nblk * 7 * 256, cudaMemcpyDeviceToHost, s_run));
            A52_CUDA(cudaMemcpyAsync(debug->bap, P.dbg_bap, nblk * 7 * 256, cudaMemcpyDeviceToHost, s_run));
        }
        if (P.dbg_coef) A52_CUDA(cudaMemcpyAsync(debug->coef, P.dbg_coef, nblk * 6 * 256 * 4, cudaMemcpyDeviceToHost, s_run));
        if (P.dbg_info) A52_CUDA(cudaMemcpyAsync(debug->info, P.dbg_info, nblk * 16 * 4, cudaMemcpyDeviceToHost, s_run));
    }
    A52_CUDA(cudaStreamSynchronize(s_run));
    A52_CUDA(cudaStreamSynchronize(s_out));
    A52_CUDA(cudaStreamSynchronize(s_in));
    return 0;
}

// ===========================================================================
// drop-in liba52 API (include/a52.h) on top of the batched path
// ===========================================================================
struct a52_state_s {
    a52_batch_t* ctx;
    sample_t* samples;          // 256 * 12 floats, pinned host memory
    float* frame_pcm;           // [6][6][256] decoded frame (planar)
    const uint8_t* frame;       // caller's buffer (valid through the 6 a52_block calls)
    int frame_len;
    int req_flags;              // request incl. A52_ADJUST_LEVEL as passed to a52_frame
    int out_flags;
    float level_in, bias;
    int drc_off;
    level_t (*drc_call)(level_t, void*);   // a52_dynrng's callback (NULL: compression as coded)
    void* drc_data;
    int blk;                    // next block to hand out
    int decoded;                // frame_pcm valid for the staged frame
    int status;
    a52_stream_carry_t carry;
};

a52_state_t* a52_init(uint32_t mm_accel)
{
    (void)mm_accel;
    int dev = 0;
    const char* e = getenv("A52_B200_DEVICE");
    if (e) dev = atoi(e);
    a52_batch_t* ctx = a52_batch_create(dev);
    if (!ctx) return nullptr;       // no CPU fallback by design
    a52_state_t* s = (a52_state_t*)calloc(1, sizeof(*s));
    if (!s) { a52_batch_destroy(ctx); return nullptr; }
    s->ctx = ctx;
    if (cudaMallocHost((void**)&s->samples, 256 * 12 * sizeof(sample_t)) != cudaSuccess ||
        cudaMallocHost((void**)&s->frame_pcm, 6 * 6 * 256 * sizeof(float)) != cudaSuccess) {
        a52_free(s);
        return nullptr;
    }
    memset(s->samples, 0, 256 * 12 * sizeof(sample_t));
    return s;
}

sample_t* a52_samples(a52_state_t* s) { return s->samples; }

int a52_syncinfo(uint8_t* buf, int* flags, int* sample_rate, int* bit_rate)
{
    return a52::host_syncinfo(buf, flags, sample_rate, bit_rate);
}

int a52_frame(a52_state_t* s, uint8_t* buf, int* flags, level_t* level, sample_t bias)
{
    using namespace a52;
    // BSI fields needed for the mode negotiation (parse.c:139-167); everything
    // else of the frame is parsed on the GPU
    int acmod = buf[6] >> 5;
    uint32_t bits = ((uint32_t)buf[6] << 24) | ((uint32_t)buf[7] << 16) | ((uint32_t)buf[8] << 8);
    int p = 3, input = acmod, cmix = 0, smix = 0;
    auto take = [&](int n) { int v = (bits << p) >> (32 - n); p += n; return v; };
    if (acmod == 2 && take(2) == 2) input = M_DOLBY;
    if ((acmod & 1) && acmod != 1) cmix = take(2);
    if (acmod & 4) smix = take(2);
    int lfeon = take(1);
    float clev, slev;
    host_mix_levels(acmod, cmix, smix, &clev, &slev);
    float lv = *level;
    int out = host_downmix_init(input, *flags, &lv, clev, slev);
    if (out < 0) return 1;
    s->req_flags = *flags;
    s->level_in = *level;
    if (lfeon && (*flags & M_LFE)) out |= M_LFE;
    *flags = out;
    *level = lv;
    s->out_flags = out;
    s->bias = bias;
    s->drc_off = 0;              // a52_frame re-arms compression as coded, without a callback (parse.c:170-171)
    s->drc_call = nullptr;
    s->drc_data = nullptr;
    s->frame = buf;
    int fl, sr, br;
    s->frame_len = host_syncinfo(buf, &fl, &sr, &br);
    if (s->frame_len <= 0) s->frame_len = 3840;      // a52_frame does not validate the header itself
    s->blk = 0;
    s->decoded = 0;
    s->status = 0;
    return 0;
}

void a52_dynrng(a52_state_t* s, level_t (*call)(level_t, void*), void* data)
{
    // NULL callback = dynamic range compression off; otherwise every dynrng word of the frame goes through the
    // callback before it is applied (parse.c:207-216, 586-595).  The words are coded inside the audio blocks, so the
    // first a52_block of the frame scans the frame for them, runs the callback on the host in block order and
    // decodes with the ranges it returned (a52_batch_scan + A52_DRC_TABLE).
    s->drc_off = (call == nullptr);
    s->drc_call = call;
    s->drc_data = data;
}

int a52_block(a52_state_t* s)
{
    if (s->blk >= 6) return 1;
    if (!s->decoded) {
        uint64_t off[2] = {0, (uint64_t)s->frame_len};
        uint32_t first[2] = {0, 1};
        int32_t status = 0, fflags = 0;
        int drc_mode = s->drc_off ? A52_DRC_OFF : A52_DRC_STREAM;
        float ranges[6][2];
        if (s->drc_call) {
            a52_frame_scan_t sc;
            if (a52_batch_scan(s->ctx, s->frame, (size_t)s->frame_len, off, 1, first, 1, s->req_flags, &sc, 0, nullptr)) return 1;
            for (int b = 0; b < 6; b++)
                for (int k = 0; k < 2; k++) {
                    ranges[b][k] = 1.0f;
                    if (sc.dynrng[b][k] < 0) continue;
                    const int d = (int)(int8_t)sc.dynrng[b][k];
                    // parse.c:586-591: (((dynrng & 0x1f) | 0x20) << 13) * scale_factor[3 - (dynrng >> 5)]
                    const float range = (float)(((d & 0x1f) | 0x20) << 13) * ldexpf(1.0f, -(15 + 3 - (d >> 5)));
                    ranges[b][k] = s->drc_call(range, s->drc_data);
                }
            a52_batch_set_drc_table(s->ctx, &ranges[0][0]);
            drc_mode = A52_DRC_TABLE;
        }
        int rc = a52_batch_decode(s->ctx, s->frame, (size_t)s->frame_len, off, 1, first, 1, s->req_flags,
                                  s->level_in, s->bias, drc_mode, A52_PCM_F32_PLANAR, s->frame_pcm, &status, &fflags,
                                  &s->carry, nullptr, 0, nullptr);
        a52_batch_set_drc_table(s->ctx, nullptr);
        if (rc) return 1;
        s->status = status;
        s->decoded = 1;
    }
    int b = s->blk++;
    if (s->status && (s->status < A52_ST_BAD_BLOCK || b >= s->status - A52_ST_BAD_BLOCK)) return 1;
    int nout = a52::nout_of_flags(s->out_flags);
    memcpy(s->samples, s->frame_pcm + (size_t)b * nout * 256, (size_t)nout * 256 * sizeof(float));
    return 0;
}

void a52_free(a52_state_t* s)
{
    if (!s) return;
    if (s->samples) cudaFreeHost(s->samples);
    if (s->frame_pcm) cudaFreeHost(s->frame_pcm);
    if (s->ctx) a52_batch_destroy(s->ctx);
    free(s);
}

}  // extern "C"
#pragma GCC visibility pop
