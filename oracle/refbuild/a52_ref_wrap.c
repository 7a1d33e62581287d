/* oracle/refbuild/a52_ref_wrap.c - TEST INFRASTRUCTURE ONLY.
 *
 * Wrapper translation unit around the UNMODIFIED reference decoder
 * (a52dec-0.7.5-cvs/liba52/parse.c, compiled from where it lies under
 * /root/reference; see oracle/Makefile).  Nothing from the reference is
 * copied into this repository: parse.c is pulled in by #include at build
 * time and the result goes to oracle/_ref/ (git-ignored).
 *
 * What the wrapper adds:
 *   - the two IMDCT entry points are interposed by macro so that the
 *     dequantised, gain-applied coefficients handed to the transform
 *     (parse.c:873,902-905,931-935) can be recorded as golden vectors;
 *   - ref_* accessors for the private decoder state (a52_internal.h:35-88):
 *     exponents, baps, endmant, coupling range, lfsr_state;
 *   - ref_* forwarding entry points, because the library is built with
 *     -fvisibility=hidden so that its a52_* symbols never collide with the
 *     product library's drop-in a52_* symbols inside one test process;
 *   - an in-memory decode loop used for the CPU baseline timing
 *     (mirrors a52dec.c:240-309 without file I/O).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define REF_API __attribute__((visibility("default")))

/* liba52 mallocs its state without clearing it (parse.c:59-67), so a damaged frame that falls back
 * on fields no earlier frame has sent reads heap garbage.  The checker makes that deterministic:
 * inside this translation unit malloc() is calloc(), i.e. "never sent" reads as zero. */
#define malloc(n) calloc (1, (n))
#define a52_imdct_512 ref_hook_imdct_512
#define a52_imdct_256 ref_hook_imdct_256
#include "parse.c"
#undef a52_imdct_512
#undef a52_imdct_256
#undef malloc

void a52_imdct_512 (sample_t * data, sample_t * delay, sample_t bias);
void a52_imdct_256 (sample_t * data, sample_t * delay, sample_t bias);

/* ---- coefficient capture --------------------------------------------- */
#define REF_MAX_CAPTURE 16
static __thread int    cap_on;
static __thread int    cap_n;
static __thread float  cap_coef[REF_MAX_CAPTURE][256];
static __thread int    cap_kind[REF_MAX_CAPTURE];   /* 512 or 256 */
static __thread long   cap_plane[REF_MAX_CAPTURE];  /* (data - samples)/256 */
static __thread sample_t * cap_base;

static void cap_record (sample_t * data, int kind)
{
    if (cap_on && cap_n < REF_MAX_CAPTURE) {
	memcpy (cap_coef[cap_n], data, 256 * sizeof (float));
	cap_kind[cap_n] = kind;
	cap_plane[cap_n] = cap_base ? (long)(data - cap_base) / 256 : -1;
	cap_n++;
    }
}

void ref_hook_imdct_512 (sample_t * data, sample_t * delay, sample_t bias)
{
    cap_record (data, 512);
    a52_imdct_512 (data, delay, bias);
}

void ref_hook_imdct_256 (sample_t * data, sample_t * delay, sample_t bias)
{
    cap_record (data, 256);
    a52_imdct_256 (data, delay, bias);
}

REF_API void ref_capture_begin (a52_state_t * st)
{
    cap_on = 1; cap_n = 0; cap_base = st->samples;
}
REF_API int ref_capture_count (void) { return cap_n; }
REF_API int ref_capture_get (int i, float * coef, int * kind, int * plane)
{
    if (i < 0 || i >= cap_n) return -1;
    memcpy (coef, cap_coef[i], 256 * sizeof (float));
    *kind = cap_kind[i];
    *plane = (int) cap_plane[i];
    return 0;
}
REF_API void ref_capture_end (void) { cap_on = 0; }

/* ---- stock API, forwarded -------------------------------------------- */
REF_API a52_state_t * ref_a52_init (uint32_t mm_accel) { return a52_init (mm_accel); }
REF_API sample_t * ref_a52_samples (a52_state_t * s) { return a52_samples (s); }
REF_API int ref_a52_syncinfo (uint8_t * buf, int * flags, int * sample_rate, int * bit_rate)
{ return a52_syncinfo (buf, flags, sample_rate, bit_rate); }
REF_API int ref_a52_frame (a52_state_t * s, uint8_t * buf, int * flags, level_t * level, sample_t bias)
{ return a52_frame (s, buf, flags, level, bias); }
REF_API void ref_a52_dynrng (a52_state_t * s, level_t (* call) (level_t, void *), void * data)
{ a52_dynrng (s, call, data); }
REF_API int ref_a52_block (a52_state_t * s) { return a52_block (s); }
REF_API void ref_a52_free (a52_state_t * s) { a52_free (s); }

/* ---- private-state accessors ------------------------------------------ */
/* which: 0..4 = fbw channel, 5 = lfe, 6 = coupling channel */
REF_API void ref_get_expbap (a52_state_t * s, int which, uint8_t * exp, int8_t * bap)
{
    expbap_t * e = (which == 6) ? &s->cpl_expbap :
		   (which == 5) ? &s->lfe_expbap : &s->fbw_expbap[which];
    memcpy (exp, e->exp, 256);
    memcpy (bap, e->bap, 256);
}

/* info[0..4]=endmant, 5=cplstrtmant, 6=cplendmant, 7=chincpl, 8=lfsr_state,
 * 9=acmod, 10=lfeon, 11=output, 12=downmixed, 13=ncplbnd, 14=rematflg,
 * 15=csnroffst */
REF_API void ref_get_info (a52_state_t * s, int * info)
{
    int i;
    for (i = 0; i < 5; i++) info[i] = s->endmant[i];
    info[5] = s->cplstrtmant; info[6] = s->cplendmant; info[7] = s->chincpl;
    info[8] = s->lfsr_state;  info[9] = s->acmod;      info[10] = s->lfeon;
    info[11] = s->output;     info[12] = s->downmixed; info[13] = s->ncplbnd;
    info[14] = s->rematflg;   info[15] = s->csnroffst;
}

REF_API void ref_set_lfsr (a52_state_t * s, int v) { s->lfsr_state = (uint16_t) v; }

/* direct access to the bit allocator for fuzzing (bit_allocate.c:124).
 * bai11 = state->bai, chbai = ba->bai (fsnroffst<<3|fgaincod), deltbae:
 * DELTA_BIT_NONE(2) or NEW(1) with deltba[50]. */
REF_API void ref_bit_allocate (int fscod, int halfrate, int bai11, int csnroffst,
			       int chbai, int deltbae, const int8_t * deltba,
			       int bndstart, int start, int end,
			       int fastleak, int slowleak,
			       const uint8_t * exp, int8_t * bap)
{
    static __thread a52_state_t st;
    static __thread ba_t ba;
    static __thread expbap_t eb;
    st.fscod = fscod; st.halfrate = halfrate; st.bai = bai11;
    st.csnroffst = csnroffst;
    ba.bai = chbai; ba.deltbae = deltbae;
    if (deltba) memcpy (ba.deltba, deltba, 50); else memset (ba.deltba, 0, 50);
    memcpy (eb.exp, exp, 256);
    memset (eb.bap, 0, 256);
    a52_bit_allocate (&st, &ba, bndstart, start, end, fastleak, slowleak, &eb);
    memcpy (bap, eb.bap, 256);
}

REF_API void ref_imdct (int kind, float * data, float * delay, float bias)
{
    if (kind == 256) a52_imdct_256 (data, delay, bias);
    else a52_imdct_512 (data, delay, bias);
}

/* ---- in-memory decode loop (CPU baseline; a52dec.c:240-309 semantics) -- */
/* Decodes consecutive frames from es[0..nbytes); writes nout*256 planar
 * floats per block to out (if non-NULL) and returns the number of frames
 * decoded, or -(frames+1) on the first error.  `st` may be NULL (a fresh
 * state is created and freed). */
REF_API long ref_decode_stream (a52_state_t * st, const uint8_t * es, long nbytes,
				int req_flags, float level_in, float bias,
				float * out, int nout, int dynrng_off)
{
    long pos = 0, frames = 0;
    int own = 0;
    if (!st) { st = a52_init (0); own = 1; if (!st) return -1; }
    while (pos + 7 <= nbytes) {
	int flags, sr, br, len, b;
	level_t level = level_in;
	len = a52_syncinfo ((uint8_t *) es + pos, &flags, &sr, &br);
	if (!len || pos + len > nbytes) break;
	flags = req_flags;
	if (a52_frame (st, (uint8_t *) es + pos, &flags, &level, bias)) {
	    frames = -(frames + 1); break;
	}
	if (dynrng_off) a52_dynrng (st, NULL, NULL);
	for (b = 0; b < 6; b++) {
	    if (a52_block (st)) { frames = -(frames + 1); goto done; }
	    if (out) {
		memcpy (out, st->samples, (size_t) nout * 256 * sizeof (float));
		out += (size_t) nout * 256;
	    }
	}
	pos += len;
	frames++;
    }
done:
    if (own) a52_free (st);
    return frames;
}
