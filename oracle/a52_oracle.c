/* oracle/a52_oracle.c - CPU restatement of the reference AC-3 decoder.
 *
 * TEST INFRASTRUCTURE ONLY (see a52_oracle.h).  This is an independent
 * re-implementation, written from the behaviour of the reference
 * (a52dec-0.7.5-cvs/liba52), of exactly the arithmetic the B200 kernels must
 * reproduce.  Every function names the reference lines it follows.  It keeps
 * the reference's *order of floating-point operations* so that PCM is
 * bit-identical to liba52 built with the same compiler flags; the integer
 * stages (exponents, bit allocation, mantissas) are written in the A/52
 * standard's own form (positive PSD, bap 0..15) - the same form the CUDA
 * kernels use - and are pinned bit-exact to liba52's inverted-sign variant by
 * tests/test_oracle.py.
 *
 * Parity status: PINNED against oracle/_ref/liba52_ref.so (the unmodified
 * reference) and the golden vectors under tests/golden/.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "a52_oracle.h"
#include "a52_tables.h"

enum {
    M_CHANNEL = 0, M_MONO, M_STEREO, M_3F, M_2F1R, M_3F1R, M_2F2R, M_3F2R,
    M_CHANNEL1, M_CHANNEL2, M_DOLBY, M_MASK = 15, M_LFE = 16, M_ADJUST = 32
};

#define K_3DB   0.7071067811865476
#define K_P3DB  1.4142135623730951
#define K_45DB  0.5946035575013605

static const uint8_t nfchans_of[11] = {2, 1, 2, 3, 3, 4, 4, 5, 1, 1, 2};

struct ora_dec {
    /* BSI (parse.c:131-205) */
    int fscod, halfrate, acmod, lfeon;
    float clev, slev;
    int output;
    float level, bias;
    int dynrnge;
    float dynrng;
    /* coupling (parse.c:600-667) */
    int chincpl, phsflginu, cplstrtmant, cplendmant, cplstrtbnd, ncplbnd;
    uint32_t cplbndstrc;
    float cplco[5][18];
    int rematflg;
    int endmant[5];
    /* bit allocation side info; index 0..4 fbw, 5 lfe, 6 cpl */
    int bai, csnroffst;
    int chbai[7];
    int deltbae[7];
    int8_t deltba[7][50];
    int cplfleak, cplsleak;
    uint8_t exp[7][256];
    uint8_t bap[7][256];	/* standard numbering 0..15 */
    /* bit cursor (bitstream.c) */
    const uint8_t * buf;
    uint32_t bitpos;
    uint16_t lfsr;
    float * samples_raw;
    float * samples;		/* 12 planes of 256: 6 output + 6 delay */
    int downmixed;
    float coef_dump[6][256];
};

/* ------------------------------------------------------------------------
 * bit reader: MSB-first fields out of a byte string (bitstream.h:53-77,
 * bitstream.c:63-97 - the reference reads big-endian 32-bit words; the
 * result of every read is the same n bits).
 */
static uint32_t getbits (ora_dec_t * d, int n)
{
    uint32_t v = 0;
    int i;
    for (i = 0; i < n; i++) {
	uint32_t p = d->bitpos + i;
	v = (v << 1) | ((d->buf[p >> 3] >> (7 - (p & 7))) & 1);
    }
    d->bitpos += n;
    return v;
}

static int32_t getbits_signed (ora_dec_t * d, int n)
{
    uint32_t v = getbits (d, n);
    if (n && (v >> (n - 1)))
	return (int32_t) v - (int32_t) (1u << n);
    return (int32_t) v;
}

/* ------------------------------------------------------------------------
 * tables computed at init (imdct.c:358-412), double then rounded to float
 */
static float w_kbd[256];
static float tw_pre1[128][2], tw_post1[64][2], tw_pre2[64][2], tw_post2[32][2];
static float root_tab[5][32];	/* root_tab[k] for transform size 16<<k */
static uint8_t bitrev_order[128];
static int tables_ready;

static void build_tables (void)
{
    int i, k;
    double acc = 0, cum[256];
    if (tables_ready)
	return;
    /* the order in which imdct.c:49-58 feeds the 128-point transform: even
     * indices are the 7-bit reversal doubled; the split-radix recursion used
     * here wants exactly that permutation.  It is generated, not tabulated:
     * entry i = 2*rev7(i) with the odd quarter-blocks mirrored (255-k). */
    {
	/* derive the permutation by running the index recursion of a
	 * conjugate-pair split-radix decimation in time */
	int n, pos = 0;
	int stack_base[256], stack_step[256], stack_n[256], sp = 0;
	stack_base[sp] = 0; stack_step[sp] = 1; stack_n[sp] = 128; sp++;
	/* iterative DFS: children pushed in reverse so they pop in order */
	while (sp) {
	    int base, step;
	    sp--;
	    base = stack_base[sp]; step = stack_step[sp]; n = stack_n[sp];
	    if (n == 1) {
		bitrev_order[pos++] = (uint8_t) (((base % 128) + 128) % 128 * 2);
		continue;
	    }
	    if (n == 2) {
		stack_base[sp] = base + step; stack_step[sp] = step * 2; stack_n[sp] = 1; sp++;
		stack_base[sp] = base; stack_step[sp] = step * 2; stack_n[sp] = 1; sp++;
		continue;
	    }
	    /* x[4k-1] (i.e. base - step), x[4k+1], x[2k] */
	    stack_base[sp] = base - step; stack_step[sp] = step * 4; stack_n[sp] = n / 4; sp++;
	    stack_base[sp] = base + step; stack_step[sp] = step * 4; stack_n[sp] = n / 4; sp++;
	    stack_base[sp] = base; stack_step[sp] = step * 2; stack_n[sp] = n / 2; sp++;
	}
    }
    /* Kaiser-Bessel derived window, alpha = 5 (imdct.c:347-372) */
    for (i = 0; i < 256; i++) {
	double x = i * (256 - i) * (5 * M_PI / 256) * (5 * M_PI / 256);
	double b = 1;
	for (k = 100; k > 0; k--)
	    b = b * x / (k * k) + 1;
	acc += b;
	cum[i] = acc;
    }
    acc++;
    for (i = 0; i < 256; i++)
	w_kbd[i] = (float) sqrt (cum[i] / acc);
    /* cosine roots for the split-radix passes (imdct.c:374-384) */
    for (k = 0; k < 4; k++) {
	int n = 16 << k;
	for (i = 0; i < n / 4 - 1; i++)
	    root_tab[k][i] = (float) cos ((M_PI / (n / 2)) * (i + 1));
    }
    /* pre/post twiddles (imdct.c:386-412) */
    for (i = 0; i < 128; i++) {
	double sgn = (i < 64) ? 1.0 : -1.0;
	k = bitrev_order[i] / 2 + 64;
	tw_pre1[i][0] = (float) (sgn * cos ((M_PI / 256) * (k - 0.25)));
	tw_pre1[i][1] = (float) (sgn * sin ((M_PI / 256) * (k - 0.25)));
    }
    for (i = 0; i < 64; i++) {
	tw_post1[i][0] = (float) cos ((M_PI / 256) * (i + 0.5));
	tw_post1[i][1] = (float) sin ((M_PI / 256) * (i + 0.5));
	k = bitrev_order[i] / 4;
	tw_pre2[i][0] = (float) cos ((M_PI / 128) * (k - 0.25));
	tw_pre2[i][1] = (float) sin ((M_PI / 128) * (k - 0.25));
    }
    for (i = 0; i < 32; i++) {
	tw_post2[i][0] = (float) cos ((M_PI / 128) * (i + 0.5));
	tw_post2[i][1] = (float) sin ((M_PI / 128) * (i + 0.5));
    }
    tables_ready = 1;
}

/* ------------------------------------------------------------------------
 * split-radix inverse FFT, same operation order as imdct.c:75-256
 */
typedef struct { float re, im; } cpx;

/* one conjugate-pair butterfly on (a0,a1,a2,a3) given the rotated inputs
 * (p = rotated a2, q = rotated a3 already combined into t1..t4) */
#define SR_COMBINE(a0, a1, a2, a3, t1, t2, t3, t4) do {	\
    a2.re = a0.re - t1; a2.im = a0.im - t2;		\
    a3.re = a1.re - t3; a3.im = a1.im - t4;		\
    a0.re += t1; a0.im += t2;				\
    a1.re += t3; a1.im += t4;				\
} while (0)

static void sr_pass (cpx * z, const float * roots, int n)
{
    /* imdct.c:194-220; n = quarter length */
    cpx * z0 = z, * z1 = z + n, * z2 = z + 2 * n, * z3 = z + 3 * n;
    float t1, t2, t3, t4, t5, t6, t7, t8;
    int k;
    /* k = 0: unit twiddle (imdct.c:145-158) */
    t1 = z2[0].re + z3[0].re;
    t2 = z2[0].im + z3[0].im;
    t3 = z2[0].im - z3[0].im;
    t4 = z3[0].re - z2[0].re;
    SR_COMBINE (z0[0], z1[0], z2[0], z3[0], t1, t2, t3, t4);
    for (k = 1; k < n; k++) {
	float wr = roots[k - 1], wi = roots[n - k - 1];
	/* imdct.c:111-141 */
	t5 = wi * z2[k].im + wr * z2[k].re;
	t6 = wr * z2[k].im - wi * z2[k].re;
	t8 = wi * z3[k].re + wr * z3[k].im;
	t7 = wr * z3[k].re - wi * z3[k].im;
	t1 = t5 + t7;
	t2 = t6 + t8;
	t3 = t6 - t8;
	t4 = t7 - t5;
	SR_COMBINE (z0[k], z1[k], z2[k], z3[k], t1, t2, t3, t4);
    }
}

static void sr_fft (cpx * z, int n)
{
    float t1, t2, t3, t4, t5, t6, t7, t8;
    if (n == 2) {		/* imdct.c:75-85 */
	float r = z[0].re, i = z[0].im;
	z[0].re += z[1].re;
	z[0].im += z[1].im;
	z[1].re = r - z[1].re;
	z[1].im = i - z[1].im;
	return;
    }
    if (n == 4) {		/* imdct.c:87-109 */
	t1 = z[0].re + z[1].re;
	t2 = z[3].re + z[2].re;
	t3 = z[0].im + z[1].im;
	t4 = z[2].im + z[3].im;
	t5 = z[0].re - z[1].re;
	t6 = z[0].im - z[1].im;
	t7 = z[2].im - z[3].im;
	t8 = z[3].re - z[2].re;
	z[0].re = t1 + t2; z[0].im = t3 + t4;
	z[2].re = t1 - t2; z[2].im = t3 - t4;
	z[1].re = t5 + t7; z[1].im = t6 + t8;
	z[3].re = t5 - t7; z[3].im = t6 - t8;
	return;
    }
    if (n == 8) {		/* imdct.c:183-192 */
	float w = root_tab[0][1];
	sr_fft (z, 4);
	sr_fft (z + 4, 2);
	sr_fft (z + 6, 2);
	t1 = z[4].re + z[6].re;
	t2 = z[4].im + z[6].im;
	t3 = z[4].im - z[6].im;
	t4 = z[6].re - z[4].re;
	SR_COMBINE (z[0], z[2], z[4], z[6], t1, t2, t3, t4);
	/* wr == wi case, imdct.c:162-181 */
	t5 = (z[5].re + z[5].im) * w;
	t6 = (z[5].im - z[5].re) * w;
	t7 = (z[7].re - z[7].im) * w;
	t8 = (z[7].im + z[7].re) * w;
	t1 = t5 + t7;
	t2 = t6 + t8;
	t3 = t6 - t8;
	t4 = t7 - t5;
	SR_COMBINE (z[1], z[3], z[5], z[7], t1, t2, t3, t4);
	return;
    }
    /* imdct.c:222-256 */
    sr_fft (z, n / 2);
    sr_fft (z + n / 2, n / 4);
    sr_fft (z + 3 * n / 4, n / 4);
    {
	int k = 0, m = n;
	while (m > 16) { m >>= 1; k++; }
	sr_pass (z, root_tab[k], n / 4);
    }
}

/* rotation helper of imdct.c:99-104: (W0,W1,d0,d1) -> (W1*d1+W0*d0, W0*d1-W1*d0) */
#define ROT(o0, o1, W0, W1, d0, d1) do {	\
    o0 = (W1) * (d1) + (W0) * (d0);		\
    o1 = (W0) * (d1) - (W1) * (d0);		\
} while (0)
/* windowed overlap-add, imdct.c:106-111 */
#define WOLA(o0, o1, W0, W1, d0, d1) do {		\
    o0 = ((d1) * (W1) + (d0) * (W0)) + bias;		\
    o1 = ((d1) * (W0) - (d0) * (W1)) + bias;		\
} while (0)

static void imdct_long (float * data, float * delay, float bias)
{
    /* imdct.c:258-293 */
    cpx z[128];
    int i;
    for (i = 0; i < 128; i++) {
	int k = bitrev_order[i];
	ROT (z[i].re, z[i].im, tw_pre1[i][0], tw_pre1[i][1], data[k], data[255 - k]);
    }
    sr_fft (z, 128);
    for (i = 0; i < 64; i++) {
	float ar, ai, br, bi;
	float pr = tw_post1[i][0], pi = tw_post1[i][1];
	ROT (ar, ai, pi, pr, z[i].im, z[i].re);
	ROT (br, bi, pr, pi, z[127 - i].im, z[127 - i].re);
	WOLA (data[255 - 2 * i], data[2 * i], w_kbd[255 - 2 * i], w_kbd[2 * i], ar, delay[2 * i]);
	delay[2 * i] = ai;
	WOLA (data[2 * i + 1], data[254 - 2 * i], w_kbd[2 * i + 1], w_kbd[254 - 2 * i], br, delay[2 * i + 1]);
	delay[2 * i + 1] = bi;
    }
}

static void imdct_short (float * data, float * delay, float bias)
{
    /* imdct.c:295-345 */
    cpx z1[64], z2[64];
    int i;
    for (i = 0; i < 64; i++) {
	int k = bitrev_order[i];
	ROT (z1[i].re, z1[i].im, tw_pre2[i][0], tw_pre2[i][1], data[k], data[254 - k]);
	ROT (z2[i].re, z2[i].im, tw_pre2[i][0], tw_pre2[i][1], data[k + 1], data[255 - k]);
    }
    sr_fft (z1, 64);
    sr_fft (z2, 64);
    for (i = 0; i < 32; i++) {
	float ar, ai, br, bi, cr, ci, dr, di;
	float pr = tw_post2[i][0], pi = tw_post2[i][1];
	ROT (ar, ai, pi, pr, z1[i].im, z1[i].re);
	ROT (br, bi, pr, pi, z1[63 - i].im, z1[63 - i].re);
	ROT (cr, ci, pi, pr, z2[i].im, z2[i].re);
	ROT (dr, di, pr, pi, z2[63 - i].im, z2[63 - i].re);
	WOLA (data[255 - 2 * i], data[2 * i], w_kbd[255 - 2 * i], w_kbd[2 * i], ar, delay[2 * i]);
	delay[2 * i] = ci;
	WOLA (data[128 + 2 * i], data[127 - 2 * i], w_kbd[128 + 2 * i], w_kbd[127 - 2 * i], ai, delay[127 - 2 * i]);
	delay[127 - 2 * i] = cr;
	WOLA (data[254 - 2 * i], data[2 * i + 1], w_kbd[254 - 2 * i], w_kbd[2 * i + 1], bi, delay[2 * i + 1]);
	delay[2 * i + 1] = dr;
	WOLA (data[129 + 2 * i], data[126 - 2 * i], w_kbd[129 + 2 * i], w_kbd[126 - 2 * i], br, delay[126 - 2 * i]);
	delay[126 - 2 * i] = di;
    }
}

void ora_imdct (int kind, float * data, float * delay, float bias)
{
    build_tables ();
    if (kind == 256)
	imdct_short (data, delay, bias);
    else
	imdct_long (data, delay, bias);
}

/* ------------------------------------------------------------------------
 * state
 */
ora_dec_t * ora_init (void)
{
    ora_dec_t * d = (ora_dec_t *) calloc (1, sizeof (*d));
    if (!d)
	return NULL;
    d->samples_raw = (float *) calloc (256 * 12 + 16, sizeof (float));
    if (!d->samples_raw) {
	free (d);
	return NULL;
    }
    d->samples = (float *) (((uintptr_t) d->samples_raw + 63) & ~(uintptr_t) 63);
    d->downmixed = 1;		/* parse.c:72 */
    d->lfsr = 1;		/* parse.c:74 */
    build_tables ();
    return d;
}

void ora_free (ora_dec_t * d)
{
    if (d) {
	free (d->samples_raw);
	free (d);
    }
}

float * ora_samples (ora_dec_t * d) { return d->samples; }
void ora_set_lfsr (ora_dec_t * d, int v) { d->lfsr = (uint16_t) v; }

/* ------------------------------------------------------------------------
 * sync info (parse.c:86-129)
 */
int ora_syncinfo (const uint8_t * b, int * flags, int * sample_rate, int * bit_rate)
{
    static const uint8_t lfe_bit[8] = {0x10, 0x10, 0x04, 0x04, 0x04, 0x01, 0x04, 0x01};
    int bsid, half, acmod, cod, kbps;
    if (b[0] != 0x0b || b[1] != 0x77)
	return 0;
    bsid = b[5] >> 3;
    if (bsid >= 12)
	return 0;
    half = (bsid > 8) ? bsid - 8 : 0;
    acmod = b[6] >> 5;
    *flags = (((b[6] & 0xf8) == 0x50) ? M_DOLBY : acmod) | ((b[6] & lfe_bit[acmod]) ? M_LFE : 0);
    cod = b[4] & 63;
    if (cod >= 38)
	return 0;
    kbps = ac3_bitrate_kbps[cod >> 1];
    *bit_rate = (kbps * 1000) >> half;
    switch (b[4] >> 6) {
    case 0: *sample_rate = 48000 >> half; return 4 * kbps;
    case 1: *sample_rate = 44100 >> half; return 2 * (320 * kbps / 147 + (cod & 1));
    case 2: *sample_rate = 32000 >> half; return 6 * kbps;
    }
    return 0;
}

/* ------------------------------------------------------------------------
 * output-mode negotiation and level adjustment (downmix.c:34-160)
 */
static int downmix_setup (int input, int flags, float * level, float clev, float slev)
{
    /* granted mode by [requested][acmod], one hex digit each (downmix.c:37-60) */
    static const char * grant[11] = {
	"0A222222", "11111111", "0A222222", "0A232323", "0A224444", "0A224545",
	"0A236666", "0A236767", "81111111", "91111111", "0A2AAAAA"
    };
    int req = flags & M_MASK, out, c;
    double adj;
    if (req > M_DOLBY)
	return -1;
    c = grant[req][input & 7];
    out = (c >= 'A') ? c - 'A' + 10 : c - '0';
    /* NB: the reference compares the float clev with the DOUBLE constant
     * (downmix.c:69-71), which is never equal in the float build - so the
     * "3F with -3 dB centre -> Dolby" rule is dead code.  Mirrored. */
    if (out == M_STEREO && (input == M_DOLBY || (input == M_3F && (double) clev == K_3DB)))
	out = M_DOLBY;
    if (!(flags & M_ADJUST))
	return out;
    /* downmix.c:73-157.  The reference evaluates these in C's usual
     * arithmetic conversions: float+float sums, then double division. */
#define PAIR(i, o) (((o) << 3) + (i))
    switch (PAIR (input & 7, out)) {
    case PAIR (M_3F, M_MONO):
	adj = K_3DB / (1 + clev); break;
    case PAIR (M_STEREO, M_MONO): case PAIR (M_2F2R, M_2F1R): case PAIR (M_3F2R, M_3F1R):
	adj = K_3DB; break;
    case PAIR (M_3F2R, M_2F1R):
	if (clev < K_P3DB - 1) { adj = K_3DB; break; }
	/* fall through */
    case PAIR (M_3F, M_STEREO): case PAIR (M_3F1R, M_2F1R):
    case PAIR (M_3F1R, M_2F2R): case PAIR (M_3F2R, M_2F2R):
	adj = 1 / (1 + clev); break;
    case PAIR (M_2F1R, M_MONO):
	adj = K_P3DB / (2 + slev); break;
    case PAIR (M_2F1R, M_STEREO): case PAIR (M_3F1R, M_3F):
	adj = 1 / (1 + slev * K_3DB); break;
    case PAIR (M_3F1R, M_MONO):
	adj = K_3DB / (1 + clev + slev * 0.5); break;
    case PAIR (M_3F1R, M_STEREO):
	adj = 1 / (1 + clev + slev * K_3DB); break;
    case PAIR (M_2F2R, M_MONO):
	adj = K_3DB / (1 + slev); break;
    case PAIR (M_2F2R, M_STEREO): case PAIR (M_3F2R, M_3F):
	adj = 1 / (1 + slev); break;
    case PAIR (M_3F2R, M_MONO):
	adj = K_3DB / (1 + clev + slev); break;
    case PAIR (M_3F2R, M_STEREO):
	adj = 1 / (1 + clev + slev); break;
    case PAIR (M_MONO, M_DOLBY):
	adj = K_P3DB; break;
    case PAIR (M_3F, M_DOLBY): case PAIR (M_2F1R, M_DOLBY):
	adj = 1 / (1 + K_3DB); break;
    case PAIR (M_3F1R, M_DOLBY): case PAIR (M_2F2R, M_DOLBY):
	adj = 1 / (1 + 2 * K_3DB); break;
    case PAIR (M_3F2R, M_DOLBY):
	adj = 1 / (1 + 3 * K_3DB); break;
    default:
	return out;
    }
    {
	float a = (float) adj;
	*level = *level * a;
    }
    return out;
}

/* per-channel gains folded into dequantisation (downmix.c:162-330).
 * Returns the mask of channels that get summed into another one. */
static int downmix_gains (float * g, int acmod, int output, float level, float clev, float slev)
{
    float l3 = (float) (level * K_3DB);
    float lc = level * clev, ls = level * slev, l3s = l3 * slev;
    float c_mono = (float) ((l3 * clev) * 2.0);
    int i;
    switch (PAIR (acmod, output & M_MASK)) {
    case PAIR (M_CHANNEL, M_CHANNEL): case PAIR (M_MONO, M_MONO):
    case PAIR (M_STEREO, M_STEREO): case PAIR (M_3F, M_3F):
    case PAIR (M_2F1R, M_2F1R): case PAIR (M_3F1R, M_3F1R):
    case PAIR (M_2F2R, M_2F2R): case PAIR (M_3F2R, M_3F2R):
    case PAIR (M_STEREO, M_DOLBY):
	for (i = 0; i < 5; i++) g[i] = level;
	return 0;
    case PAIR (M_CHANNEL, M_MONO):
	g[0] = g[1] = (float) (level * 0.5); return 3;
    case PAIR (M_STEREO, M_MONO):
	g[0] = g[1] = l3; return 3;
    case PAIR (M_3F, M_MONO):
	g[0] = g[2] = l3; g[1] = c_mono; return 7;
    case PAIR (M_2F1R, M_MONO):
	g[0] = g[1] = l3; g[2] = l3s; return 7;
    case PAIR (M_2F2R, M_MONO):
	g[0] = g[1] = l3; g[2] = g[3] = l3s; return 15;
    case PAIR (M_3F1R, M_MONO):
	g[0] = g[2] = l3; g[1] = c_mono; g[3] = l3s; return 15;
    case PAIR (M_3F2R, M_MONO):
	g[0] = g[2] = l3; g[1] = c_mono; g[3] = g[4] = l3s; return 31;
    case PAIR (M_MONO, M_DOLBY):
	g[0] = l3; return 0;
    case PAIR (M_3F, M_DOLBY):
	g[0] = g[2] = g[3] = g[4] = level; g[1] = l3; return 7;
    case PAIR (M_3F, M_STEREO): case PAIR (M_3F1R, M_2F1R): case PAIR (M_3F2R, M_2F2R):
	g[0] = g[2] = g[3] = g[4] = level; g[1] = lc; return 7;
    case PAIR (M_2F1R, M_DOLBY):
	g[0] = g[1] = level; g[2] = l3; return 7;
    case PAIR (M_2F1R, M_STEREO):
	g[0] = g[1] = level; g[2] = l3s; return 7;
    case PAIR (M_3F1R, M_DOLBY):
	g[0] = g[2] = level; g[1] = g[3] = l3; return 15;
    case PAIR (M_3F1R, M_STEREO):
	g[0] = g[2] = level; g[1] = lc; g[3] = l3s; return 15;
    case PAIR (M_2F2R, M_DOLBY):
	g[0] = g[1] = level; g[2] = g[3] = l3; return 15;
    case PAIR (M_2F2R, M_STEREO):
	g[0] = g[1] = level; g[2] = g[3] = ls; return 15;
    case PAIR (M_3F2R, M_DOLBY):
	g[0] = g[2] = level; g[1] = g[3] = g[4] = l3; return 31;
    case PAIR (M_3F2R, M_2F1R):
	g[0] = g[2] = level; g[1] = lc; g[3] = g[4] = l3; return 31;
    case PAIR (M_3F2R, M_STEREO):
	g[0] = g[2] = level; g[1] = lc; g[3] = g[4] = ls; return 31;
    case PAIR (M_3F1R, M_3F):
	g[0] = g[1] = g[2] = level; g[3] = l3s; return 13;
    case PAIR (M_3F2R, M_3F):
	g[0] = g[1] = g[2] = level; g[3] = g[4] = ls; return 29;
    case PAIR (M_2F2R, M_2F1R):
	g[0] = g[1] = level; g[2] = g[3] = l3; return 12;
    case PAIR (M_3F2R, M_3F1R):
	g[0] = g[1] = g[2] = level; g[3] = g[4] = l3; return 24;
    case PAIR (M_2F1R, M_2F2R):
	g[0] = g[1] = level; g[2] = l3; return 0;
    case PAIR (M_3F1R, M_2F2R):
	g[0] = g[2] = level; g[1] = lc; g[3] = l3; return 7;
    case PAIR (M_3F1R, M_3F2R):
	g[0] = g[1] = g[2] = level; g[3] = l3; return 0;
    case PAIR (M_CHANNEL, M_CHANNEL1):
	g[0] = level; g[1] = 0; return 0;
    case PAIR (M_CHANNEL, M_CHANNEL2):
	g[0] = 0; g[1] = level; return 0;
    }
    return -1;
}

/* ------------------------------------------------------------------------
 * plane mixing (downmix.c:332-619).  P(n) = plane n of 256 floats.
 */
#define P(n) (s + 256 * (n))
#define LOOP for (i = 0; i < 256; i++)

static void add_plane (float * dst, const float * src, float bias)
{
    int i;
    LOOP dst[i] += src[i] + bias;	/* downmix.c:332-338 */
}

static void fold_centre (float * s, float bias)
{
    int i;				/* downmix.c:368-379 */
    LOOP {
	float c = P (1)[i] + bias;
	P (0)[i] += c;
	P (1)[i] = P (2)[i] + c;
    }
}

static void mix_planes (float * s, int acmod, int output, float bias, float slev)
{
    int i;
    int o = output & M_MASK;
    switch (PAIR (acmod, o)) {
    case PAIR (M_CHANNEL, M_CHANNEL2):
	memcpy (P (0), P (1), 1024);
	break;
    case PAIR (M_CHANNEL, M_MONO): case PAIR (M_STEREO, M_MONO):
    sum2:
	add_plane (P (0), P (1), bias);
	break;
    case PAIR (M_2F1R, M_MONO):
	if (slev == 0) goto sum2;
	/* fall through */
    case PAIR (M_3F, M_MONO):
    sum3:
	LOOP P (0)[i] += (P (1)[i] + P (2)[i]) + bias;
	break;
    case PAIR (M_3F1R, M_MONO):
	if (slev == 0) goto sum3;
	/* fall through */
    case PAIR (M_2F2R, M_MONO):
	if (slev == 0) goto sum2;
	LOOP P (0)[i] += (P (1)[i] + P (2)[i] + P (3)[i]) + bias;
	break;
    case PAIR (M_3F2R, M_MONO):
	if (slev == 0) goto sum3;
	LOOP P (0)[i] += (P (1)[i] + P (2)[i] + P (3)[i] + P (4)[i]) + bias;
	break;
    case PAIR (M_MONO, M_DOLBY):
	memcpy (P (1), P (0), 1024);
	break;
    case PAIR (M_3F, M_STEREO): case PAIR (M_3F, M_DOLBY):
    centre:
	fold_centre (s, bias);
	break;
    case PAIR (M_2F1R, M_STEREO):
	if (slev == 0) break;
	LOOP {
	    float c = P (2)[i] + bias;
	    P (0)[i] += c;
	    P (1)[i] += c;
	}
	break;
    case PAIR (M_2F1R, M_DOLBY):
	LOOP {
	    float sr = P (2)[i];
	    P (0)[i] += -sr + bias;
	    P (1)[i] += sr + bias;
	}
	break;
    case PAIR (M_3F1R, M_STEREO):
	if (slev == 0) goto centre;
	LOOP {
	    float c = (P (1)[i] + P (3)[i]) + bias;
	    P (0)[i] += c;
	    P (1)[i] = P (2)[i] + c;
	}
	break;
    case PAIR (M_3F1R, M_DOLBY):
	LOOP {
	    float c = P (1)[i] + bias, sr = P (3)[i];
	    P (0)[i] += c - sr;
	    P (1)[i] = P (2)[i] + c + sr;
	}
	break;
    case PAIR (M_2F2R, M_STEREO):
	if (slev == 0) break;
	add_plane (P (0), P (2), bias);
	add_plane (P (1), P (3), bias);
	break;
    case PAIR (M_2F2R, M_DOLBY):
	LOOP {
	    float sr = P (2)[i] + P (3)[i];
	    P (0)[i] += -sr + bias;
	    P (1)[i] += sr + bias;
	}
	break;
    case PAIR (M_3F2R, M_STEREO):
	if (slev == 0) goto centre;
	LOOP {
	    float c = P (1)[i] + bias;
	    P (0)[i] += c + P (3)[i];
	    P (1)[i] = c + P (2)[i] + P (4)[i];
	}
	break;
    case PAIR (M_3F2R, M_DOLBY):
	LOOP {
	    float c = P (1)[i] + bias, sr = P (3)[i] + P (4)[i];
	    P (0)[i] += c - sr;
	    P (1)[i] = P (2)[i] + c + sr;
	}
	break;
    case PAIR (M_3F1R, M_3F):
	if (slev == 0) break;
	LOOP {
	    float c = P (3)[i] + bias;
	    P (0)[i] += c;
	    P (2)[i] += c;
	}
	break;
    case PAIR (M_3F2R, M_3F):
	if (slev == 0) break;
	add_plane (P (0), P (3), bias);
	add_plane (P (2), P (4), bias);
	break;
    case PAIR (M_3F1R, M_2F1R):
	fold_centre (s, bias);
	memcpy (P (2), P (3), 1024);
	break;
    case PAIR (M_2F2R, M_2F1R):
	add_plane (P (2), P (3), bias);
	break;
    case PAIR (M_3F2R, M_2F1R):
	fold_centre (s, bias);
	LOOP P (2)[i] = (P (3)[i] + P (4)[i]) + bias;
	break;
    case PAIR (M_3F2R, M_3F1R):
	add_plane (P (3), P (4), bias);
	break;
    case PAIR (M_2F1R, M_2F2R):
	memcpy (P (3), P (2), 1024);
	break;
    case PAIR (M_3F1R, M_2F2R):
	fold_centre (s, bias);
	memcpy (P (2), P (3), 1024);
	break;
    case PAIR (M_3F2R, M_2F2R):
	fold_centre (s, bias);
	memcpy (P (2), P (3), 1024);
	memcpy (P (3), P (4), 1024);
	break;
    case PAIR (M_3F1R, M_3F2R):
	memcpy (P (4), P (3), 1024);
	break;
    }
}

/* undo the plane packing of a downmixed delay buffer (downmix.c:621-685) */
static void unmix_planes (float * s, int acmod, int output)
{
    int o = output & M_MASK;
    int zero_from = -1;		/* planes [zero_from, zero_to) cleared afterwards */
    int zero_to = 0;
    switch (PAIR (acmod, o)) {
    case PAIR (M_CHANNEL, M_CHANNEL2):
	memcpy (P (1), P (0), 1024);
	return;
    case PAIR (M_3F2R, M_MONO): zero_from = 1; zero_to = 5; break;
    case PAIR (M_3F1R, M_MONO): case PAIR (M_2F2R, M_MONO): zero_from = 1; zero_to = 4; break;
    case PAIR (M_3F, M_MONO): case PAIR (M_2F1R, M_MONO): zero_from = 1; zero_to = 3; break;
    case PAIR (M_CHANNEL, M_MONO): case PAIR (M_STEREO, M_MONO): zero_from = 1; zero_to = 2; break;
    case PAIR (M_3F2R, M_STEREO): case PAIR (M_3F2R, M_DOLBY):
	memset (P (4), 0, 1024);
	/* fall through */
    case PAIR (M_3F1R, M_STEREO): case PAIR (M_3F1R, M_DOLBY):
	memset (P (3), 0, 1024);
	/* fall through */
    case PAIR (M_3F, M_STEREO): case PAIR (M_3F, M_DOLBY):
	memcpy (P (2), P (1), 1024);
	memset (P (1), 0, 1024);
	return;
    case PAIR (M_2F2R, M_STEREO): case PAIR (M_2F2R, M_DOLBY): zero_from = 2; zero_to = 4; break;
    case PAIR (M_2F1R, M_STEREO): case PAIR (M_2F1R, M_DOLBY): zero_from = 2; zero_to = 3; break;
    case PAIR (M_3F2R, M_3F): zero_from = 3; zero_to = 5; break;
    case PAIR (M_3F1R, M_3F): case PAIR (M_2F2R, M_2F1R): zero_from = 3; zero_to = 4; break;
    case PAIR (M_3F2R, M_3F1R): zero_from = 4; zero_to = 5; break;
    case PAIR (M_3F2R, M_2F1R):
	memset (P (4), 0, 1024);
	memcpy (P (3), P (2), 1024);
	memcpy (P (2), P (1), 1024);
	memset (P (1), 0, 1024);
	return;
    case PAIR (M_3F1R, M_2F1R):
	memcpy (P (3), P (2), 1024);
	memcpy (P (2), P (1), 1024);
	memset (P (1), 0, 1024);
	return;
    case PAIR (M_3F2R, M_2F2R):
	memcpy (P (4), P (3), 1024);
	memcpy (P (3), P (2), 1024);
	memcpy (P (2), P (1), 1024);
	memset (P (1), 0, 1024);
	return;
    default:
	return;
    }
    if (zero_from >= 0)
	memset (P (zero_from), 0, (size_t) (zero_to - zero_from) * 1024);
}

/* ------------------------------------------------------------------------
 * frame header / BSI (parse.c:131-205)
 */
int ora_frame (ora_dec_t * d, const uint8_t * buf, int * flags, float * level, float bias)
{
    static const double cmix[4] = {K_3DB, K_45DB, 0.5, K_45DB};
    static const double smix[4] = {K_3DB, 0.5, 0, 0.5};
    int acmod, bsid, i, reps;

    d->fscod = buf[4] >> 6;
    bsid = buf[5] >> 3;
    d->halfrate = (bsid > 8 && bsid < 12) ? bsid - 8 : 0;
    d->acmod = acmod = buf[6] >> 5;
    d->buf = buf;
    d->bitpos = 6 * 8 + 3;
    if (acmod == 2 && getbits (d, 2) == 2)
	acmod = M_DOLBY;
    d->clev = d->slev = 0;
    if ((acmod & 1) && acmod != 1)
	d->clev = (float) cmix[getbits (d, 2)];
    if (acmod & 4)
	d->slev = (float) smix[getbits (d, 2)];
    d->lfeon = getbits (d, 1);

    d->output = downmix_setup (acmod, *flags, level, d->clev, d->slev);
    if (d->output < 0)
	return 1;
    if (d->lfeon && (*flags & M_LFE))
	d->output |= M_LFE;
    *flags = d->output;
    d->dynrng = d->level = (float) (*level * 2.0);	/* parse.c:169 */
    d->bias = bias;
    d->dynrnge = 1;
    for (i = 0; i < 7; i++)
	if (i != 5)
	    d->deltbae[i] = 2;				/* DELTA_BIT_NONE, parse.c:173-175 */

    reps = acmod ? 1 : 2;
    while (reps--) {
	getbits (d, 5);
	if (getbits (d, 1)) getbits (d, 8);
	if (getbits (d, 1)) getbits (d, 8);
	if (getbits (d, 1)) getbits (d, 7);
    }
    getbits (d, 2);
    if (getbits (d, 1)) getbits (d, 14);
    if (getbits (d, 1)) getbits (d, 14);
    if (getbits (d, 1)) {
	int n = getbits (d, 6);
	d->bitpos += 8 * (n + 1);
    }
    return 0;
}

void ora_dynrng_off (ora_dec_t * d) { d->dynrnge = 0; }	/* parse.c:207-216 with call == NULL */

/* ------------------------------------------------------------------------
 * exponents (parse.c:218-270, tables.h:24-47)
 */
static int unpack_exponents (ora_dec_t * d, int strategy, int ngrps, int start, uint8_t * dst)
{
    int rep = 1 << (strategy - 1);	/* D15 1, D25 2, D45 4 */
    uint8_t e = (uint8_t) start;	/* the reference keeps the running sum in a uint8_t */
    while (ngrps--) {
	int code = getbits (d, 7);
	int dig[3], j, r;
	if (code >= 125) {
	    /* tables.h pads 125..127 with +25: the first delta already overflows */
	    return 1;
	}
	dig[0] = code / 25; dig[1] = (code / 5) % 5; dig[2] = code % 5;
	for (j = 0; j < 3; j++) {
	    e = (uint8_t) (e + dig[j] - 2);
	    if (e > 24)
		return 1;
	    for (r = 0; r < rep; r++)
		*dst++ = e;
	}
    }
    return 0;
}

/* delta bit allocation segments (parse.c:272-294) */
static int unpack_deltba (ora_dec_t * d, int8_t * dst)
{
    int nseg, band = 0;
    memset (dst, 0, 50);
    nseg = getbits (d, 3) + 1;
    while (nseg--) {
	int len, code, delta;
	band += getbits (d, 5);
	len = getbits (d, 4);
	code = getbits (d, 3);
	delta = (code >= 4) ? code - 3 : code - 4;
	if (!len)
	    continue;
	if (band + len >= 50)
	    return 1;
	while (len--)
	    dst[band++] = (int8_t) delta;
    }
    return 0;
}

/* ------------------------------------------------------------------------
 * parametric bit allocation in the A/52 standard's form; bit-exact with the
 * inverted-sign variant at bit_allocate.c:124-265 (pinned by fuzzing in
 * tests/test_oracle.py).
 *   psd = 3072 - 128*exp; band integration by log-addition; excitation with
 *   low-frequency compensation and fast/slow leak; hearing threshold; delta;
 *   snr offset; bap = baptab[(psd - mask) >> 5].
 */
static int lowcomp_step (int a, int b0, int b1, int band)
{
    if (band < 7) {
	if (b0 + 256 == b1) a = 384;
	else if (b0 > b1) { a -= 64; if (a < 0) a = 0; }
    } else if (band < 20) {
	if (b0 + 256 == b1) a = 320;
	else if (b0 > b1) { a -= 64; if (a < 0) a = 0; }
    } else {
	a -= 128; if (a < 0) a = 0;
    }
    return a;
}

static void bit_allocate (int fscod, int halfrate, int bai, int csnroffst, int chbai,
			  int deltbae, const int8_t * deltba, int start, int end,
			  int fastleak, int slowleak, int is_lfe,
			  const uint8_t * exp, uint8_t * bap)
{
    int sdecay = (0x0f + 2 * (bai >> 9)) >> halfrate;
    int fdecay = (0x3f + 0x14 * ((bai >> 7) & 3)) >> halfrate;
    int sgain = ac3_sgain[(bai >> 5) & 3];
    int dbknee = ac3_dbknee[(bai >> 3) & 3];
    int floorv = ac3_floor[bai & 7];
    int fgain = 0x80 * ((chbai & 7) + 1);
    int snroffset = (((csnroffst - 15) << 4) + (chbai >> 3)) << 2;
    int psd[256], bndpsd[50], excite[50], mask[50];
    int bin, band, bndstrt, bndend, lowcomp = 0, begin;

    for (bin = start; bin < end; bin++)
	psd[bin] = 3072 - (exp[bin] << 7);

    /* band integration */
    bndstrt = ac3_masktab[start];
    bndend = ac3_masktab[end - 1] + 1;
    bin = start;
    for (band = bndstrt; band < bndend; band++) {
	int last = ac3_bndtab[band + 1] < end ? ac3_bndtab[band + 1] : end;
	int v = psd[bin++];
	for (; bin < last; bin++) {
	    int c = v - psd[bin];
	    int adr = (c >= 0 ? c : -c) >> 1;
	    if (adr > 255) adr = 255;
	    v = (c >= 0 ? v : psd[bin]) + ac3_latab[adr];
	}
	bndpsd[band] = v;
    }

    /* excitation */
    if (bndstrt == 0) {
	lowcomp = lowcomp_step (lowcomp, bndpsd[0], bndpsd[1], 0);
	excite[0] = bndpsd[0] - fgain - lowcomp;
	lowcomp = lowcomp_step (lowcomp, bndpsd[1], bndpsd[2], 1);
	excite[1] = bndpsd[1] - fgain - lowcomp;
	begin = 7;
	for (band = 2; band < 7; band++) {
	    int last_lfe = is_lfe && band == 6;
	    if (!last_lfe)
		lowcomp = lowcomp_step (lowcomp, bndpsd[band], bndpsd[band + 1], band);
	    fastleak = bndpsd[band] - fgain;
	    slowleak = bndpsd[band] - sgain;
	    excite[band] = fastleak - lowcomp;
	    if (!last_lfe && bndpsd[band] <= bndpsd[band + 1]) {
		begin = band + 1;
		break;
	    }
	}
	for (band = begin; band < (bndend < 22 ? bndend : 22); band++) {
	    int v;
	    if (!(is_lfe && band == 6))
		lowcomp = lowcomp_step (lowcomp, bndpsd[band], bndpsd[band + 1], band);
	    fastleak -= fdecay;
	    if (fastleak < bndpsd[band] - fgain) fastleak = bndpsd[band] - fgain;
	    slowleak -= sdecay;
	    if (slowleak < bndpsd[band] - sgain) slowleak = bndpsd[band] - sgain;
	    v = fastleak - lowcomp;
	    excite[band] = v > slowleak ? v : slowleak;
	}
	begin = 22;
    } else {
	begin = bndstrt;	/* coupling channel: leaks supplied by the caller */
    }
    for (band = begin; band < bndend; band++) {
	fastleak -= fdecay;
	if (fastleak < bndpsd[band] - fgain) fastleak = bndpsd[band] - fgain;
	slowleak -= sdecay;
	if (slowleak < bndpsd[band] - sgain) slowleak = bndpsd[band] - sgain;
	excite[band] = fastleak > slowleak ? fastleak : slowleak;
    }

    /* masking curve + delta + snr offset, then the pointer lookup */
    for (band = bndstrt; band < bndend; band++) {
	int v = excite[band], h;
	if (bndpsd[band] < dbknee)
	    v += (dbknee - bndpsd[band]) >> 2;
	h = ac3_hth[fscod * 50 + (band >> halfrate)];
	if (h > v) v = h;
	if (deltbae == 0 || deltbae == 1)
	    v += deltba[band] * 128;
	v -= snroffset + floorv;
	if (v < 0) v = 0;
	v &= 0x1fe0;
	mask[band] = v + floorv;
    }
    for (bin = start; bin < end; bin++) {
	int a = (psd[bin] - mask[ac3_masktab[bin]]) >> 5;
	if (a < 0) a = 0;
	if (a > 63) a = 63;
	bap[bin] = ac3_baptab[a];
    }
}

void ora_bit_allocate (int fscod, int halfrate, int bai11, int csnroffst, int chbai,
		       int deltbae, const int8_t * deltba, int bndstart, int start,
		       int end, int fastleak, int slowleak, const uint8_t * exp,
		       int8_t * bap)
{
    uint8_t b[256];
    int8_t zero[50];
    int i;
    (void) bndstart;
    memset (b, 0, sizeof (b));
    memset (zero, 0, sizeof (zero));
    /* liba52 passes the coupling leaks as (9-x)<<8 in its inverted domain;
     * standard domain: 3072 - that (= (x<<8) + 768) */
    bit_allocate (fscod, halfrate, bai11, csnroffst, chbai, deltbae,
		  deltba ? deltba : zero, start, end,
		  start ? 3072 - fastleak : 0, start ? 3072 - slowleak : 0,
		  (start == 0 && end == 7), exp, b);
    memset (bap, 0, 256);
    for (i = start; i < end; i++)
	bap[i] = ac3_bap_liba52[b[i]];
}

/* ------------------------------------------------------------------------
 * mantissas (parse.c:310-433, 435-556)
 */
typedef struct {
    int n1, n2, n4;		/* values still pending from the last group code */
    int v1[2], v2[2], v4;
} grp_state;

static int dither_next (ora_dec_t * d)
{
    /* parse.c:310-319; the LUT is the CRC-16 step for x^16+x^15+x^13+x^4+1... (0xA011) */
    uint16_t s = d->lfsr;
    int16_t n = (int16_t) (ac3_dither_lut[s >> 8] ^ (uint16_t) (s << 8));
    d->lfsr = (uint16_t) n;
    return (3 * n) >> 2;
}

/* Returns the integer quantiser value (Q15 scale) of the next mantissa with
 * pointer `b` (b != 0), consuming bits as required. */
static int next_mantissa (ora_dec_t * d, grp_state * g, int b)
{
    int code;
    switch (b) {
    case 1:
	if (g->n1) return g->v1[--g->n1];
	code = getbits (d, 5);
	if (code >= 27) { g->n1 = 2; g->v1[0] = g->v1[1] = 0; return 0; }
	g->n1 = 2;
	g->v1[1] = ac3_q3[(code / 3) % 3];
	g->v1[0] = ac3_q3[code % 3];
	return ac3_q3[code / 9];
    case 2:
	if (g->n2) return g->v2[--g->n2];
	code = getbits (d, 7);
	if (code >= 125) { g->n2 = 2; g->v2[0] = g->v2[1] = 0; return 0; }
	g->n2 = 2;
	g->v2[1] = ac3_q5[(code / 5) % 5];
	g->v2[0] = ac3_q5[code % 5];
	return ac3_q5[code / 25];
    case 3:
	return ac3_q7[getbits (d, 3)];
    case 4:
	if (g->n4) { g->n4 = 0; return g->v4; }
	code = getbits (d, 7);
	g->n4 = 1;
	if (code >= 121) { g->v4 = 0; return 0; }
	g->v4 = ac3_q11[code % 11];
	return ac3_q11[code / 11];
    case 5:
	return ac3_q15[getbits (d, 4)];
    default:
	{
	    int w = ac3_bap_bits[b];
	    return getbits_signed (d, w) * (1 << (16 - w));
	}
    }
}

static void unpack_channel (ora_dec_t * d, float * coef, const uint8_t * exp, const uint8_t * bap,
			    grp_state * g, float gain, int dither, int end)
{
    /* parse.c:336-433 */
    float factor[25];
    int i;
    for (i = 0; i <= 24; i++)
	factor[i] = ldexpf (1.0f, -(15 + i)) * gain;
    for (i = 0; i < end; i++) {
	if (bap[i] == 0)
	    coef[i] = dither ? (float) dither_next (d) * factor[exp[i]] : 0.0f;
	else
	    coef[i] = (float) next_mantissa (d, g, bap[i]) * factor[exp[i]];
    }
}

static void unpack_coupling (ora_dec_t * d, int nfchans, const float * gain, float * planes,
			     grp_state * g, const uint8_t * dithflag)
{
    /* parse.c:435-556 */
    const uint8_t * exp = d->exp[6], * bap = d->bap[6];
    uint32_t strc = d->cplbndstrc;
    int bnd = 0, i = d->cplstrtmant, ch;
    while (i < d->cplendmant) {
	float co[5];
	int stop = i + 12;
	while (strc & 1) { strc >>= 1; stop += 12; }
	strc >>= 1;
	for (ch = 0; ch < nfchans; ch++)
	    co[ch] = d->cplco[ch][bnd] * gain[ch];
	bnd++;
	for (; i < stop; i++) {
	    float sf = ldexpf (1.0f, -(15 + exp[i]));
	    if (bap[i] == 0) {
		for (ch = 0; ch < nfchans; ch++)
		    if ((d->chincpl >> ch) & 1)
			planes[256 * ch + i] = dithflag[ch] ? (sf * co[ch]) * (float) dither_next (d) : 0.0f;
	    } else {
		float m = (float) next_mantissa (d, g, bap[i]) * sf;
		for (ch = 0; ch < nfchans; ch++)
		    if ((d->chincpl >> ch) & 1)
			planes[256 * ch + i] = m * co[ch];
	    }
	}
    }
}

/* ------------------------------------------------------------------------
 * one audio block (parse.c:558-940)
 */
int ora_block (ora_dec_t * d)
{
    static const int remat_edge[4] = {25, 37, 61, 253};
    static const uint8_t cpl_band_of[16] = {31, 35, 37, 39, 41, 42, 43, 44, 45, 45, 46, 46, 47, 47, 48, 48};
    int nfchans = nfchans_of[d->acmod];
    int blksw[5], chexpstr[5], cplexpstr = 0, lfeexpstr = 0, do_alloc = 0;
    uint8_t dithflag[5];
    float gain[5];
    int i, j, reps, chanbias;
    float * s;
    grp_state g;

    for (i = 0; i < nfchans; i++) blksw[i] = getbits (d, 1);
    for (i = 0; i < nfchans; i++) dithflag[i] = (uint8_t) getbits (d, 1);

    /* dynamic range (parse.c:578-598) */
    reps = d->acmod ? 1 : 2;
    while (reps--) {
	if (getbits (d, 1)) {
	    int w = getbits_signed (d, 8);
	    if (d->dynrnge) {
		float range = (float) (((w & 0x1f) | 0x20) << 13) * ldexpf (1.0f, -(15 + 3 - (w >> 5)));
		d->dynrng = d->level * range;
	    }
	}
    }

    /* coupling strategy (parse.c:600-634) */
    if (getbits (d, 1)) {
	d->chincpl = 0;
	if (getbits (d, 1)) {
	    int begf, endf, nsub;
	    for (i = 0; i < nfchans; i++)
		d->chincpl |= getbits (d, 1) << i;
	    if (d->acmod < 2)
		return 1;
	    if (d->acmod == 2)
		d->phsflginu = getbits (d, 1);
	    begf = getbits (d, 4);
	    endf = getbits (d, 4);
	    nsub = endf + 3 - begf;
	    if (nsub < 0)
		return 1;
	    d->ncplbnd = nsub;
	    d->cplstrtbnd = cpl_band_of[begf];
	    d->cplstrtmant = begf * 12 + 37;
	    d->cplendmant = endf * 12 + 73;
	    d->cplbndstrc = 0;
	    for (i = 0; i < nsub - 1; i++)
		if (getbits (d, 1)) {
		    d->cplbndstrc |= 1u << i;
		    d->ncplbnd--;
		}
	}
    }

    /* coupling coordinates (parse.c:636-667) */
    if (d->chincpl) {
	int any = 0;
	for (i = 0; i < nfchans; i++)
	    if ((d->chincpl >> i) & 1)
		if (getbits (d, 1)) {
		    int mstr = 3 * getbits (d, 2);
		    any = 1;
		    for (j = 0; j < d->ncplbnd; j++) {
			int e = getbits (d, 4), m = getbits (d, 4);
			m = (e == 15) ? m << 14 : (m | 0x10) << 13;
			d->cplco[i][j] = (float) m * ldexpf (1.0f, -(15 + e + mstr));
		    }
		}
	if (d->acmod == 2 && d->phsflginu && any)
	    for (j = 0; j < d->ncplbnd; j++)
		if (getbits (d, 1))
		    d->cplco[1][j] = -d->cplco[1][j];
    }

    /* rematrixing flags (parse.c:669-678) */
    if (d->acmod == 2 && getbits (d, 1)) {
	int stop = d->chincpl ? d->cplstrtmant : 253;
	d->rematflg = 0;
	i = 0;
	do
	    d->rematflg |= getbits (d, 1) << i;
	while (remat_edge[i++] < stop);
    }

    /* exponent strategies and bandwidth (parse.c:680-701) */
    if (d->chincpl) cplexpstr = getbits (d, 2);
    for (i = 0; i < nfchans; i++) chexpstr[i] = getbits (d, 2);
    if (d->lfeon) lfeexpstr = getbits (d, 1);
    for (i = 0; i < nfchans; i++)
	if (chexpstr[i]) {
	    if ((d->chincpl >> i) & 1)
		d->endmant[i] = d->cplstrtmant;
	    else {
		int bw = getbits (d, 6);
		if (bw > 60)
		    return 1;
		d->endmant[i] = bw * 3 + 73;
	    }
	}

    /* exponents (parse.c:703-736) */
    if (cplexpstr) {
	int ngrps = (d->cplendmant - d->cplstrtmant) / (3 << (cplexpstr - 1));
	int absexp = getbits (d, 4) << 1;
	do_alloc = 64;
	if (unpack_exponents (d, cplexpstr, ngrps, absexp, d->exp[6] + d->cplstrtmant))
	    return 1;
    }
    for (i = 0; i < nfchans; i++)
	if (chexpstr[i]) {
	    int gsz = 3 << (chexpstr[i] - 1);
	    int ngrps = (d->endmant[i] + gsz - 4) / gsz;
	    do_alloc |= 1 << i;
	    d->exp[i][0] = (uint8_t) getbits (d, 4);
	    if (unpack_exponents (d, chexpstr[i], ngrps, d->exp[i][0], d->exp[i] + 1))
		return 1;
	    getbits (d, 2);	/* gainrng */
	}
    if (lfeexpstr) {
	do_alloc |= 32;
	d->exp[5][0] = (uint8_t) getbits (d, 4);
	if (unpack_exponents (d, lfeexpstr, 2, d->exp[5][0], d->exp[5] + 1))
	    return 1;
    }

    /* bit-allocation side info (parse.c:738-772) */
    if (getbits (d, 1)) {
	do_alloc = 127;
	d->bai = getbits (d, 11);
    }
    if (getbits (d, 1)) {
	do_alloc = 127;
	d->csnroffst = getbits (d, 6);
	if (d->chincpl) d->chbai[6] = getbits (d, 7);
	for (i = 0; i < nfchans; i++) d->chbai[i] = getbits (d, 7);
	if (d->lfeon) d->chbai[5] = getbits (d, 7);
    }
    if (d->chincpl && getbits (d, 1)) {
	do_alloc |= 64;
	d->cplfleak = 9 - getbits (d, 3);
	d->cplsleak = 9 - getbits (d, 3);
    }
    if (getbits (d, 1)) {
	do_alloc = 127;
	if (d->chincpl) d->deltbae[6] = getbits (d, 2);
	for (i = 0; i < nfchans; i++) d->deltbae[i] = getbits (d, 2);
	if (d->chincpl && d->deltbae[6] == 1 && unpack_deltba (d, d->deltba[6]))
	    return 1;
	for (i = 0; i < nfchans; i++)
	    if (d->deltbae[i] == 1 && unpack_deltba (d, d->deltba[i]))
		return 1;
    }

    /* bit allocation (parse.c:774-798) */
    if (do_alloc) {
	int allzero = !d->csnroffst && !(d->chincpl && (d->chbai[6] >> 3)) &&
		      !(d->lfeon && (d->chbai[5] >> 3));
	for (i = 0; i < nfchans && allzero; i++)
	    if (d->chbai[i] >> 3)
		allzero = 0;
	if (allzero) {
	    memset (d->bap[6], 0, 256);
	    for (i = 0; i < nfchans; i++)
		memset (d->bap[i], 0, 256);
	    memset (d->bap[5], 0, 256);
	} else {
	    if (d->chincpl && (do_alloc & 64))
		bit_allocate (d->fscod, d->halfrate, d->bai, d->csnroffst, d->chbai[6],
			      d->deltbae[6], d->deltba[6], d->cplstrtmant, d->cplendmant,
			      3072 - ((d->cplfleak) << 8), 3072 - ((d->cplsleak) << 8), 0,
			      d->exp[6], d->bap[6]);
	    for (i = 0; i < nfchans; i++)
		if (do_alloc & (1 << i))
		    bit_allocate (d->fscod, d->halfrate, d->bai, d->csnroffst, d->chbai[i],
				  d->deltbae[i], d->deltba[i], 0, d->endmant[i], 0, 0, 0,
				  d->exp[i], d->bap[i]);
	    if (d->lfeon && (do_alloc & 32)) {
		d->deltbae[5] = 2;
		bit_allocate (d->fscod, d->halfrate, d->bai, d->csnroffst, d->chbai[5],
			      2, d->deltba[5], 0, 7, 0, 0, 1, d->exp[5], d->bap[5]);
	    }
	}
    }

    /* skip field (parse.c:800-804) */
    if (getbits (d, 1)) {
	int n = getbits (d, 9);
	d->bitpos += 8 * n;
    }

    s = d->samples;
    if (d->output & M_LFE)
	s += 256;
    chanbias = downmix_gains (gain, d->acmod, d->output, d->dynrng, d->clev, d->slev);

    /* mantissas (parse.c:813-835) */
    memset (&g, 0, sizeof (g));
    {
	int done_cpl = 0;
	for (i = 0; i < nfchans; i++) {
	    unpack_channel (d, s + 256 * i, d->exp[i], d->bap[i], &g, gain[i], dithflag[i], d->endmant[i]);
	    if ((d->chincpl >> i) & 1) {
		if (!done_cpl) {
		    done_cpl = 1;
		    unpack_coupling (d, nfchans, gain, s, &g, dithflag);
		}
		j = d->cplendmant;
	    } else
		j = d->endmant[i];
	    for (; j < 256; j++)
		s[256 * i + j] = 0;
	}
    }

    /* rematrixing (parse.c:837-865) */
    if (d->acmod == 2) {
	int end = d->endmant[0] < d->endmant[1] ? d->endmant[0] : d->endmant[1];
	int flags = d->rematflg, k = 0;
	j = 13;
	do {
	    int edge = remat_edge[k++];
	    if (flags & 1) {
		if (edge > end) edge = end;
		do {
		    float a = s[j], b = s[256 + j];
		    s[j] = a + b;
		    s[256 + j] = a - b;
		} while (++j < edge);
	    } else
		j = edge;
	    flags >>= 1;
	} while (j < end);
    }

    for (i = 0; i < 5; i++)
	if (i < nfchans) memcpy (d->coef_dump[i], s + 256 * i, 1024);
	else memset (d->coef_dump[i], 0, 1024);
    memset (d->coef_dump[5], 0, 1024);

    /* LFE (parse.c:867-879) */
    if (d->lfeon) {
	if (d->output & M_LFE) {
	    unpack_channel (d, s - 256, d->exp[5], d->bap[5], &g, d->dynrng, 0, 7);
	    for (i = 7; i < 256; i++) (s - 256)[i] = 0;
	    memcpy (d->coef_dump[5], s - 256, 1024);
	    imdct_long (s - 256, s + 1536 - 256, d->bias);
	} else {
	    float scratch[8];
	    unpack_channel (d, scratch, d->exp[5], d->bap[5], &g, 0, 0, 7);
	}
    }

    /* transform + mix (parse.c:881-937) */
    i = 0;
    if (nfchans_of[d->output & M_MASK] < nfchans)
	for (i = 1; i < nfchans; i++)
	    if (blksw[i] != blksw[0])
		break;
    if (i < nfchans) {
	/* per-channel transforms, then mix in the time domain */
	if (d->downmixed) {
	    d->downmixed = 0;
	    unmix_planes (s + 1536, d->acmod, d->output);
	}
	for (i = 0; i < nfchans; i++) {
	    float b = (chanbias & (1 << i)) ? 0 : d->bias;
	    if (gain[i]) {
		if (blksw[i]) imdct_short (s + 256 * i, s + 1536 + 256 * i, b);
		else imdct_long (s + 256 * i, s + 1536 + 256 * i, b);
	    } else
		for (j = 0; j < 256; j++)
		    s[256 * i + j] = b;
	}
	mix_planes (s, d->acmod, d->output, d->bias, d->slev);
    } else {
	/* mix coefficients, then only the output channels are transformed */
	int nout = nfchans_of[d->output & M_MASK];
	mix_planes (s, d->acmod, d->output, 0, d->slev);
	if (!d->downmixed) {
	    d->downmixed = 1;
	    mix_planes (s + 1536, d->acmod, d->output, 0, d->slev);
	}
	for (i = 0; i < nout; i++) {
	    if (blksw[0]) imdct_short (s + 256 * i, s + 1536 + 256 * i, d->bias);
	    else imdct_long (s + 256 * i, s + 1536 + 256 * i, d->bias);
	}
    }
    return 0;
}

/* ------------------------------------------------------------------------
 * accessors + stream loop
 */
void ora_get_expbap (ora_dec_t * d, int which, uint8_t * exp, int8_t * bap)
{
    int i;
    memcpy (exp, d->exp[which], 256);
    for (i = 0; i < 256; i++)
	bap[i] = ac3_bap_liba52[d->bap[which][i]];
}

void ora_get_info (ora_dec_t * d, int * info)
{
    int i;
    for (i = 0; i < 5; i++) info[i] = d->endmant[i];
    info[5] = d->cplstrtmant; info[6] = d->cplendmant; info[7] = d->chincpl;
    info[8] = d->lfsr;        info[9] = d->acmod;      info[10] = d->lfeon;
    info[11] = d->output;     info[12] = d->downmixed; info[13] = d->ncplbnd;
    info[14] = d->rematflg;   info[15] = d->csnroffst;
}

void ora_get_coeffs (ora_dec_t * d, float * coef)
{
    memcpy (coef, d->coef_dump, sizeof (d->coef_dump));
}

long ora_decode_stream (ora_dec_t * d, const uint8_t * es, long nbytes, int req_flags,
			float level_in, float bias, float * out, int nout, int dynrng_off)
{
    long pos = 0, frames = 0;
    int own = 0;
    if (!d) {
	d = ora_init ();
	own = 1;
	if (!d) return -1;
    }
    while (pos + 7 <= nbytes) {
	int flags, sr, br, len, b;
	float level = level_in;
	len = ora_syncinfo (es + pos, &flags, &sr, &br);
	if (!len || pos + len > nbytes)
	    break;
	flags = req_flags;
	if (ora_frame (d, es + pos, &flags, &level, bias)) {
	    frames = -(frames + 1);
	    break;
	}
	if (dynrng_off)
	    ora_dynrng_off (d);
	for (b = 0; b < 6; b++) {
	    if (ora_block (d)) {
		frames = -(frames + 1);
		goto done;
	    }
	    if (out) {
		memcpy (out, d->samples, (size_t) nout * 256 * sizeof (float));
		out += (size_t) nout * 256;
	    }
	}
	pos += len;
	frames++;
    }
done:
    if (own)
	ora_free (d);
    return frames;
}
