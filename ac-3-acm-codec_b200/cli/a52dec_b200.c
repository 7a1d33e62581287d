/* a52dec_b200 - the a52dec command line on the batched B200 engine.
 *
 * Same options, output drivers and byte-for-byte output files as the reference's CLI
 * (a52dec-0.7.5-cvs/src/a52dec.c:155-238 options, :240-309 framing, :311-588 demultiplexers, :600-645 main;
 * libao/audio_out_wav.c, audio_out_aif.c, audio_out_float.c, audio_out_peak.c, audio_out_null.c), but the
 * per-frame a52_frame / 6 x a52_block loop is replaced by a52_batch_decode(): the whole input is framed on
 * the host (a52_syncinfo, the 7-byte sliding resync of a52dec.c), handed to the GPU in chunks of frames with
 * the per-stream carry record between chunks, and the PCM comes back in the layout the driver writes
 * (A52_PCM_S16_WAV does libao's convert2s16_wav on the device).
 *
 * New surface: several input files form one batch (one stream each) with `-O <dir>`; outputs are named
 * <dir>/<input basename>.<wav|aif|raw|txt>.
 *
 * Host language: C over the C ABI of liba52_b200.so (include/a52.h, include/a52_batch.h).  No CPU decode path
 * exists here: without a CUDA device a52_batch_create() fails and so does this program.
 */
#define _GNU_SOURCE
#include <errno.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>
#include <unistd.h>

#include "a52.h"
#include "a52_batch.h"

/* ---------------------------------------------------------------------------------------------------- */
/* output drivers (libao/audio_out.c:55-92 order and names)                                              */
/* ---------------------------------------------------------------------------------------------------- */
enum { K_WAV, K_AIF, K_PEAK, K_NULL, K_FLOAT };

typedef struct {
    const char * name;
    int kind;
    int req;      /* request passed as a52_frame's flags (before A52_ADJUST_LEVEL) */
    int fmt;      /* PCM layout asked of the engine */
    float bias;
    const char * ext;
} driver_t;

static const driver_t drivers[] = {
    {"wav", K_WAV, A52_STEREO, A52_PCM_S16_WAV, 384.f, "wav"},
    {"wavdolby", K_WAV, A52_DOLBY, A52_PCM_S16_WAV, 384.f, "wav"},
    {"wav6", K_WAV, A52_REQ_AS_CODED | A52_LFE, A52_PCM_S16_WAV, 384.f, "wav"},
    {"aif", K_AIF, A52_STEREO, A52_PCM_S16_INTERLEAVED, 384.f, "aif"},
    {"aifdolby", K_AIF, A52_DOLBY, A52_PCM_S16_INTERLEAVED, 384.f, "aif"},
    {"peak", K_PEAK, A52_STEREO, A52_PCM_F32_PLANAR, 0.f, "txt"},
    {"peakdolby", K_PEAK, A52_DOLBY, A52_PCM_F32_PLANAR, 0.f, "txt"},
    {"null", K_NULL, A52_STEREO, A52_PCM_F32_PLANAR, 384.f, "raw"},
    {"null4", K_NULL, A52_2F2R, A52_PCM_F32_PLANAR, 384.f, "raw"},
    {"null6", K_NULL, A52_3F2R | A52_LFE, A52_PCM_F32_PLANAR, 384.f, "raw"},
    {"float", K_FLOAT, A52_STEREO, A52_PCM_F32_PLANAR, 0.f, "raw"},
    {NULL, 0, 0, 0, 0.f, NULL}
};

/* per-output-file state: what libao keeps in its instance structs */
typedef struct {
    FILE * fp;
    int set_params;        /* header not written yet (audio_out_wav.c:124, audio_out_aif.c:95) */
    int sample_rate;
    uint32_t speaker_flags;
    int size;              /* payload bytes so far */
    float peak;
} sink_t;

static void le32 (uint8_t * p, uint32_t v) { p[0] = v; p[1] = v >> 8; p[2] = v >> 16; p[3] = v >> 24; }
static void le16 (uint8_t * p, uint32_t v) { p[0] = v; p[1] = v >> 8; }
static void be32 (uint8_t * p, uint32_t v) { p[0] = v >> 24; p[1] = v >> 16; p[2] = v >> 8; p[3] = v; }
static void be16 (uint8_t * p, uint32_t v) { p[0] = v >> 8; p[1] = v; }

/* channels and WAVE speaker mask of a granted mode (audio_out_wav.c:97-116) */
static int wav_channels (int flags, uint32_t * speakers)
{
    static const uint16_t spk[11] = {3, 4, 3, 7, 0x103, 0x107, 0x33, 0x37, 4, 4, 3};
    static const uint8_t nch[11] = {2, 1, 2, 3, 3, 4, 4, 5, 1, 1, 2};
    int m = flags & A52_CHANNEL_MASK, n = nch[m];
    *speakers = spk[m];
    if (flags & A52_LFE) {
	*speakers |= 8;
	n++;
    }
    return n;
}

/* RIFF header: plain PCM for mono / stereo, WAVE_FORMAT_EXTENSIBLE otherwise (audio_out_wav.c:44-58,
 * 124-142, 158-174).  `size` < 0 writes the open-ended sizes of a stream still being written. */
static size_t wav_header (uint8_t * h, uint32_t speakers, int chans, int rate, int size)
{
    const int plain = (speakers == 3 || speakers == 4);
    const size_t n = plain ? 44 : 68;
    memset (h, 0, 68);
    memcpy (h, "RIFF", 4);
    memcpy (h + 8, "WAVEfmt ", 8);
    le32 (h + 16, plain ? 16 : 40);
    le16 (h + 20, plain ? 1 : 0xfffe);
    le16 (h + 22, chans);
    le32 (h + 24, rate);
    le32 (h + 28, rate * 2 * chans);
    le16 (h + 32, 2 * chans);
    le16 (h + 34, 16);
    if (!plain) {
	static const uint8_t guid_tail[14] = {0, 0, 0, 0, 0x10, 0x00, 0x80, 0, 0, 0xaa, 0, 0x38, 0x9b, 0x71};
	le16 (h + 36, 22);
	le16 (h + 38, 16);
	le32 (h + 40, speakers);
	le16 (h + 44, 1);
	memcpy (h + 46, guid_tail, 14);
    }
    memcpy (h + n - 8, "data", 4);
    if (size < 0) {
	le32 (h + 4, plain ? 0xfffffffc : 0xfffffff0);
	le32 (h + n - 4, plain ? 0xffffffd8 : 0xffffffb4);
    } else {
	le32 (h + 4, size + (plain ? 36 : 60));
	le32 (h + n - 4, size);
    }
    return n;
}

/* AIFF header (audio_out_aif.c:42-47, 95-99, 111-121) */
static size_t aif_header (uint8_t * h, int rate, int size)
{
    static const uint8_t proto[54] = {
	'F', 'O', 'R', 'M', 0xff, 0xff, 0xff, 0xfe, 'A', 'I', 'F', 'F', 'C', 'O', 'M', 'M', 0, 0, 0, 18,
	0, 2, 0x3f, 0xff, 0xff, 0xf4, 0, 16, 0x40, 0x0e, 0, 0, 0, 0, 0, 0, 0, 0,
	'S', 'S', 'N', 'D', 0xff, 0xff, 0xff, 0xd8, 0, 0, 0, 0, 0, 0, 0, 0};
    memcpy (h, proto, sizeof (proto));
    be16 (h + 30, rate);
    if (size >= 0) {
	be32 (h + 4, size + 46);
	be32 (h + 22, size / 4);
	be32 (h + 42, size + 8);
    }
    return sizeof (proto);
}

/* ---------------------------------------------------------------------------------------------------- */
/* input: elementary stream, or one of the three demultiplexers                                          */
/* ---------------------------------------------------------------------------------------------------- */
typedef struct {
    uint8_t * p;
    size_t n, cap;
} bytes_t;

static void bytes_add (bytes_t * b, const uint8_t * src, size_t n)
{
    if (b->n + n + 32 > b->cap) {
	b->cap = (b->n + n + 32) * 2;
	b->p = (uint8_t *) realloc (b->p, b->cap);
	if (!b->p) {
	    fprintf (stderr, "out of memory\n");
	    exit (1);
	}
    }
    memcpy (b->p + b->n, src, n);
    b->n += n;
}

static int read_all (FILE * f, bytes_t * b)
{
    uint8_t tmp[65536];
    size_t n;
    while ((n = fread (tmp, 1, sizeof (tmp), f)) > 0)
	bytes_add (b, tmp, n);
    return ferror (f) ? -1 : 0;
}

/* PES header sizes as a52dec.c:478-507 derives them; 0 = the header is cut off by the end of the input.
 * Returns the offset of the byte that carries the substream id. */
static size_t ps_private1_header (const uint8_t * h, size_t avail)
{
    static const int mpeg1_skip[16] = {0, 0, 4, 9, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    size_t len;
    if (avail < 7)
	return 0;
    if ((h[6] & 0xc0) == 0x80) {	/* mpeg2 */
	if (avail < 9)
	    return 0;
	len = 10 + h[8];
    } else {				/* mpeg1: stuffing, STD buffer, time stamps */
	len = 7;
	while (h[len - 1] == 0xff) {
	    len++;
	    if (avail < len)
		return 0;
	    if (len == 23) {
		fprintf (stderr, "too much stuffing\n");
		break;
	    }
	}
	if ((h[len - 1] & 0xc0) == 0x40) {
	    len += 2;
	    if (avail < len)
		return 0;
	}
	len += mpeg1_skip[h[len - 1] >> 4] + 1;
    }
    return avail < len ? 0 : len;
}

/* Program stream (-s) and bare PES (-T) over the whole input (a52dec.c:311-529, ps_loop :531-540).
 * Returns 1 when the reference would have called exit (1) at this point. */
static int demux_ps (const uint8_t * buf, size_t n, int track, int pes_only, bytes_t * es)
{
    size_t i = 0;
    while (n - i >= 4) {
	const uint8_t * h = buf + i;
	const size_t avail = n - i;
	size_t len;
	long payload;
	if (h[0] || h[1] || h[2] != 1) {
	    i++;
	    continue;
	}
	if (pes_only) {
	    if (h[3] != 0xbd) {
		fprintf (stderr, "bad stream id %x\n", h[3]);
		return 1;
	    }
	    if (avail < 9)
		break;
	    if ((h[6] & 0xc0) != 0x80) {
		fprintf (stderr, "bad multiplex - not mpeg2\n");
		return 1;
	    }
	    len = 9 + h[8];
	    if (avail < len)
		break;
	    payload = 6 + (h[4] << 8) + h[5] - (long) len;
	    i += len;
	    if (payload > 0) {
		size_t take = (size_t) payload > n - i ? n - i : (size_t) payload;
		bytes_add (es, buf + i, take);
		i += take;
	    }
	    continue;
	}
	switch (h[3]) {
	case 0xb9:	/* program end code */
	    return 0;
	case 0xba:	/* pack header */
	    if (avail < 5)
		return 0;
	    if ((h[4] & 0xc0) == 0x40) {
		if (avail < 14)
		    return 0;
		len = 14 + (h[13] & 7);
	    } else if ((h[4] & 0xf0) == 0x20) {
		len = 12;
	    } else {
		fprintf (stderr, "weird pack header\n");
		len = 5;
	    }
	    if (avail < len)
		return 0;
	    i += len;
	    break;
	case 0xbd:	/* private stream 1 */
	    len = ps_private1_header (h, avail);
	    if (!len)
		return 0;
	    if (h[len - 1] != track) {
		payload = 6 + (h[4] << 8) + h[5] - (long) len;
		i += len;
		if (payload > 0)
		    i += (size_t) payload > n - i ? n - i : (size_t) payload;
		break;
	    }
	    len += 3;	/* frame count and first-access-unit pointer */
	    if (avail < len)
		return 0;
	    payload = 6 + (h[4] << 8) + h[5] - (long) len;
	    i += len;
	    if (payload > 0) {
		size_t take = (size_t) payload > n - i ? n - i : (size_t) payload;
		bytes_add (es, buf + i, take);
		i += take;
	    }
	    break;
	default:
	    if (h[3] < 0xb9) {
		fprintf (stderr, "looks like a video stream, not system stream\n");
		return 1;
	    }
	    if (avail < 6)
		return 0;
	    payload = (h[4] << 8) + h[5];
	    i += 6;
	    i += (size_t) payload > n - i ? n - i : (size_t) payload;
	}
    }
    return 0;
}

/* Transport stream (-t pid): a52dec.c:542-581 picks the packets, :311-447 walks the PES inside them.  The
 * PES header may straddle packets; once it is through, every payload byte up to the next payload-start
 * packet is audio unless the PES packet ended inside the packet that carried its header. */
static int demux_ts (const uint8_t * buf, size_t n, int pid, bytes_t * es)
{
    enum { IN_HEADER, IN_DATA, IN_SKIP } st = IN_SKIP;
    uint8_t head[268];
    size_t hlen = 0, i = 0;
    while (i + 188 <= n) {
	const uint8_t * pk = buf + i, * next = pk + 188, * d;
	if (pk[0] != 0x47) {
	    fprintf (stderr, "bad sync byte\n");
	    i++;
	    continue;
	}
	i += 188;
	if ((((pk[1] << 8) + pk[2]) & 0x1fff) != pid)
	    continue;
	d = pk + 4;
	if (pk[3] & 0x20) {
	    d = pk + 5 + pk[4];
	    if (d > next)
		continue;
	}
	if (!(pk[3] & 0x10))
	    continue;
	if (pk[1] & 0x40) {		/* payload unit start */
	    st = IN_HEADER;
	    hlen = 0;
	} else if (st == IN_HEADER && hlen == 0) {
	    st = IN_SKIP;
	}
	if (st == IN_DATA) {
	    bytes_add (es, d, (size_t) (next - d));
	    continue;
	}
	if (st == IN_SKIP)
	    continue;
	/* header bytes: 4 for the start code, 9 for the flags, then 9 + header_data_length */
	for (;;) {
	    size_t need = hlen < 4 ? 4 : hlen < 9 ? 9 : (size_t) 9 + head[8];
	    size_t take;
	    if (hlen >= 9 && hlen == need)
		break;
	    take = need - hlen;
	    if (take > (size_t) (next - d))
		take = (size_t) (next - d);
	    memcpy (head + hlen, d, take);
	    hlen += take;
	    d += take;
	    if (hlen < need)
		break;			/* continues in the next packet of this pid */
	    if (hlen == 4) {
		if (head[0] || head[1] || head[2] != 1) {
		    st = IN_SKIP;
		    break;
		}
		if (head[3] != 0xbd) {
		    fprintf (stderr, "bad stream id %x\n", head[3]);
		    return 1;
		}
	    } else if (hlen == 9 && (head[6] & 0xc0) != 0x80) {
		fprintf (stderr, "bad multiplex - not mpeg2\n");
		return 1;
	    }
	}
	if (st == IN_HEADER && hlen >= 9 && hlen == (size_t) 9 + head[8]) {
	    const long payload = 6 + (head[4] << 8) + head[5] - (long) hlen;
	    const long here = next - d;
	    if (payload > here) {
		bytes_add (es, d, (size_t) here);
		st = IN_DATA;
	    } else {
		if (payload > 0)
		    bytes_add (es, d, (size_t) payload);
		st = IN_SKIP;
	    }
	}
    }
    return 0;
}

/* ---------------------------------------------------------------------------------------------------- */
/* framing: the sliding 7-byte resync of a52_decode_data (a52dec.c:240-309)                              */
/* ---------------------------------------------------------------------------------------------------- */
typedef struct {
    uint64_t * off;
    int * rate;
    int * flags;
    int n, cap;
    long skipped;
} frames_t;

static void index_frames (uint8_t * es, size_t n, frames_t * fr)
{
    size_t p = 0;
    while (p + 7 <= n) {
	int flags, rate, bit_rate;
	int len = a52_syncinfo (es + p, &flags, &rate, &bit_rate);
	if (!len) {
	    fr->skipped++;
	    p++;
	    continue;
	}
	if (p + len > n)
	    break;		/* the reference waits for more input here and never gets it */
	if (fr->n == fr->cap) {
	    fr->cap = fr->cap ? fr->cap * 2 : 1024;
	    fr->off = (uint64_t *) realloc (fr->off, fr->cap * sizeof (uint64_t));
	    fr->rate = (int *) realloc (fr->rate, fr->cap * sizeof (int));
	    fr->flags = (int *) realloc (fr->flags, fr->cap * sizeof (int));
	    if (!fr->off || !fr->rate || !fr->flags) {
		fprintf (stderr, "out of memory\n");
		exit (1);
	    }
	}
	fr->off[fr->n] = p;
	fr->rate[fr->n] = rate;
	fr->flags[fr->n] = flags;
	fr->n++;
	p += len;
    }
}

/* ---------------------------------------------------------------------------------------------------- */
/* one input = one stream of the batch                                                                   */
/* ---------------------------------------------------------------------------------------------------- */
typedef struct {
    const char * path;
    bytes_t es;
    frames_t fr;
    int next;		/* next frame to hand to the engine */
    int fatal;		/* a demultiplexer hit one of the reference's exit (1) cases */
    sink_t sink;
    a52_stream_carry_t carry;
    long errors;
} input_t;

static const driver_t * drv;
static int disable_dynrng, disable_adjust;
static double gain = 1;
static int demux_track, demux_pid, demux_pes;
static const char * out_dir;
static int extract_only;	/* -x: write the demultiplexed elementary stream (what extract_a52 does) */

static void usage (const char * argv0)
{
    int i;
    fprintf (stderr,
	     "usage: %s [-h] [-o <mode>] [-s [<track>]] [-t <pid>] [-T] [-c] [-r] [-a] \\\n"
	     "\t\t[-g <gain>] [-O <dir>] [-C <frames>] [-x] <file> [<file> ...]\n"
	     "\t-h\tdisplay help and available audio output modes\n"
	     "\t-s\tuse program stream demultiplexer, track 0-7 or 0x80-0x87\n"
	     "\t-t\tuse transport stream demultiplexer, pid 0x10-0x1ffe\n"
	     "\t-T\tuse transport stream PES demultiplexer\n"
	     "\t-c\taccepted for compatibility (there is one implementation: the GPU's)\n"
	     "\t-r\tdisable dynamic range compression\n"
	     "\t-a\tdisable level adjustment based on output mode\n"
	     "\t-g\tadd specified gain in decibels, -96.0 to +96.0\n"
	     "\t-O\tdecode all files as one batch, one output per file in <dir>\n"
	     "\t-C\tframes of a stream per engine call (default 4096)\n"
	     "\t-x\tonly demultiplex: write the AC-3 elementary stream (extract_a52; default -s 0x80)\n"
	     "\t-o\taudio output mode\n", argv0);
    for (i = 0; drivers[i].name; i++)
	fprintf (stderr, "\t\t\t%s\n", drivers[i].name);
    exit (1);
}

/* writes the blocks of one decoded frame the way the driver's play() would; returns 1 when play() would
 * have failed (frame abandoned) */
static int play_frame (input_t * in, const uint8_t * pcm, int granted, int nblocks, int rate)
{
    sink_t * k = &in->sink;
    uint32_t speakers;
    const int chans = wav_channels (granted, &speakers);
    uint8_t hdr[68];
    int b;
    if (nblocks <= 0)
	return 0;
    switch (drv->kind) {
    case K_WAV:
	if (k->set_params) {
	    k->set_params = 0;
	    k->speaker_flags = speakers;
	    k->sample_rate = rate;
	    fwrite (hdr, wav_header (hdr, speakers, chans, rate, -1), 1, k->fp);
	} else if (speakers != k->speaker_flags)
	    return 1;
	fwrite (pcm, (size_t) 512 * chans, nblocks, k->fp);
	k->size += 512 * chans * nblocks;
	break;
    case K_AIF:
	if (k->set_params) {
	    k->set_params = 0;
	    k->sample_rate = rate;
	    fwrite (hdr, aif_header (hdr, rate, -1), 1, k->fp);
	}
	for (b = 0; b < nblocks; b++) {
	    /* convert2s16_2 + s16_BE on planes 0 and 1; a one-channel grant leaves plane 1 at +0.0f,
	     * which convert() turns into -32768 (convert2s16.c:33-41) */
	    const int16_t * s = (const int16_t *) pcm + (size_t) b * 256 * chans;
	    uint8_t o[1024];
	    int i;
	    for (i = 0; i < 256; i++) {
		be16 (o + 4 * i, (uint16_t) s[i * chans]);
		be16 (o + 4 * i + 2, chans > 1 ? (uint16_t) s[i * chans + 1] : 0x8000);
	    }
	    fwrite (o, 1, sizeof (o), k->fp);
	}
	k->size += 1024 * nblocks;
	break;
    case K_FLOAT:
    case K_PEAK:
	for (b = 0; b < nblocks; b++) {
	    /* the driver takes 512 floats whatever the grant: planes 0 and 1 of a52_samples () */
	    const float * s = (const float *) pcm + (size_t) b * 256 * chans;
	    float o[512];
	    int i;
	    memcpy (o, s, 256 * sizeof (float) * (chans > 1 ? 2 : 1));
	    if (chans == 1)
		memset (o + 256, 0, 256 * sizeof (float));
	    if (drv->kind == K_FLOAT)
		fwrite (o, sizeof (float), 512, k->fp);
	    else
		for (i = 0; i < 512; i++)
		    if (k->peak < fabsf (o[i]))
			k->peak = fabsf (o[i]);
	}
	break;
    default:
	break;
    }
    return 0;
}

static void close_sink (input_t * in)
{
    sink_t * k = &in->sink;
    uint8_t hdr[68];
    if (drv->kind == K_PEAK)
	fprintf (k->fp, "peak level = %.4f (%+.2f dB)\n", k->peak, 6 * log (k->peak) / log (2));
    if ((drv->kind == K_WAV || drv->kind == K_AIF) && fseek (k->fp, 0, SEEK_SET) >= 0) {
	if (drv->kind == K_WAV) {
	    uint32_t sp = k->speaker_flags;
	    int chans = 0, m;
	    /* channel count as written at open time: recover it from the speaker mask */
	    for (m = sp; m; m &= m - 1)
		chans++;
	    fwrite (hdr, wav_header (hdr, sp, chans, k->sample_rate, k->size), 1, k->fp);
	} else
	    fwrite (hdr, aif_header (hdr, k->sample_rate, k->size), 1, k->fp);
    }
    if (k->fp != stdout)
	fclose (k->fp);
    else
	fflush (stdout);
}

int main (int argc, char ** argv)
{
    int c, i, ninputs, fatal = 0;
    char * s;
    input_t * in;
    a52_batch_t * ctx;
    struct timeval t0, t1;
    long total_frames = 0;
    int chunk = 4096;		/* frames of one stream per engine call (-C) */

    fprintf (stderr, "a52dec_b200 - a52dec's command line on the batched B200 AC-3 engine\n");
    while ((c = getopt (argc, argv, "hs::t:Tcrag:o:O:C:x")) != -1)
	switch (c) {
	case 'o':
	    for (i = 0; drivers[i].name; i++)
		if (strcmp (drivers[i].name, optarg) == 0)
		    drv = drivers + i;
	    if (!drv) {
		fprintf (stderr, "Invalid video driver: %s\n", optarg);
		usage (argv[0]);
	    }
	    break;
	case 's':
	    demux_track = 0x80;
	    if (optarg) {
		demux_track = strtol (optarg, &s, 0);
		if (demux_track < 0x80)
		    demux_track += 0x80;
		if (demux_track < 0x80 || demux_track > 0x87 || *s) {
		    fprintf (stderr, "Invalid track number: %s\n", optarg);
		    usage (argv[0]);
		}
	    }
	    break;
	case 't':
	    demux_pid = strtol (optarg, &s, 0);
	    if (demux_pid < 0x10 || demux_pid > 0x1ffe || *s) {
		fprintf (stderr, "Invalid pid: %s\n", optarg);
		usage (argv[0]);
	    }
	    break;
	case 'T':
	    demux_pes = 1;
	    break;
	case 'c':
	    break;
	case 'r':
	    disable_dynrng = 1;
	    break;
	case 'a':
	    disable_adjust = 1;
	    break;
	case 'g':
	    gain = strtod (optarg, &s);
	    if (gain < -96 || gain > 96 || *s) {
		fprintf (stderr, "Invalid gain: %s\n", optarg);
		usage (argv[0]);
	    }
	    gain = pow (2, gain / 6);
	    break;
	case 'O':
	    out_dir = optarg;
	    break;
	case 'x':
	    extract_only = 1;
	    break;
	case 'C':
	    chunk = strtol (optarg, &s, 0);
	    if (chunk < 1 || *s) {
		fprintf (stderr, "Invalid chunk: %s\n", optarg);
		usage (argv[0]);
	    }
	    break;
	default:
	    usage (argv[0]);
	}
    if (!drv)
	drv = drivers;
    if (extract_only && !demux_pid && !demux_pes && !demux_track)
	demux_track = 0x80;	/* extract_a52.c:40: program stream, first AC-3 track */
    ninputs = argc - optind;
    if (ninputs > 1 && !out_dir) {
	fprintf (stderr, "several inputs need -O <dir>\n");
	usage (argv[0]);
    }
    if (ninputs == 0)
	ninputs = 1;		/* stdin */
    in = (input_t *) calloc (ninputs, sizeof (input_t));

    for (i = 0; i < ninputs; i++) {
	FILE * f = stdin;
	bytes_t raw = {NULL, 0, 0};
	in[i].path = optind + i < argc ? argv[optind + i] : "stdin";
	if (optind + i < argc && !(f = fopen (argv[optind + i], "rb"))) {
	    fprintf (stderr, "%s - could not open file %s\n", strerror (errno), argv[optind + i]);
	    return 1;
	}
	read_all (f, &raw);
	if (f != stdin)
	    fclose (f);
	if (demux_pid) {
	    in[i].fatal = demux_ts (raw.p, raw.n, demux_pid, &in[i].es);
	    free (raw.p);
	} else if (demux_track || demux_pes) {
	    in[i].fatal = demux_ps (raw.p, raw.n, demux_track, demux_pes && !demux_track, &in[i].es);
	    free (raw.p);
	} else
	    in[i].es = raw;
	if (!in[i].es.p)
	    bytes_add (&in[i].es, (const uint8_t *) "", 0);
	if (extract_only) {
	    /* extract_a52 (src/extract_a52.c): the demultiplexers above with fwrite in the place of the
	     * decoder; no GPU is touched */
	    FILE * fp = stdout;
	    if (out_dir) {
		char path[4096];
		const char * base = strrchr (in[i].path, '/');
		base = base ? base + 1 : in[i].path;
		snprintf (path, sizeof (path), "%s/%s.ac3", out_dir, base);
		if (!(fp = fopen (path, "wb"))) {
		    fprintf (stderr, "%s - could not open file %s\n", strerror (errno), path);
		    return 1;
		}
	    }
	    fwrite (in[i].es.p, 1, in[i].es.n, fp);
	    if (fp != stdout)
		fclose (fp);
	    else
		fflush (stdout);
	    fatal |= in[i].fatal;
	    continue;
	}
	memset (in[i].es.p + in[i].es.n, 0, 16);
	index_frames (in[i].es.p, in[i].es.n, &in[i].fr);
	for (long k = 0; k < in[i].fr.skipped; k++)
	    fprintf (stderr, "skip\n");
	fatal |= in[i].fatal;
	in[i].sink.set_params = 1;
	if (out_dir) {
	    char path[4096];
	    const char * base = strrchr (in[i].path, '/');
	    base = base ? base + 1 : in[i].path;
	    snprintf (path, sizeof (path), "%s/%s.%s", out_dir, base, drv->ext);
	    if (!(in[i].sink.fp = fopen (path, "wb"))) {
		fprintf (stderr, "%s - could not open file %s\n", strerror (errno), path);
		return 1;
	    }
	} else
	    in[i].sink.fp = stdout;
    }

    if (extract_only)
	return fatal ? 1 : 0;
    ctx = a52_batch_create (0);
    if (!ctx) {
	fprintf (stderr, "A52 init failed\n");
	return 1;
    }
    gettimeofday (&t0, NULL);
    {
	const int req = drv->req | (disable_adjust ? 0 : A52_ADJUST_LEVEL);
	const size_t stride = a52_batch_frame_stride (req, drv->fmt);
	int left = 1;
	while (left) {
	    /* this round's frames: up to `chunk` of every stream, packed into one bitstream buffer so that
	     * offsets stay small; frames the driver's setup() refuses (sample rate changed after the header
	     * went out: audio_out_wav.c:63-66) never reach the decoder, as in the reference */
	    bytes_t pack = {NULL, 0, 0};
	    uint64_t * off;
	    uint32_t * first = (uint32_t *) malloc ((ninputs + 1) * sizeof (uint32_t));
	    int * src;
	    int nf = 0, cap = 0, f;
	    uint8_t * pcm;
	    int32_t * status, * granted;
	    a52_stream_carry_t * carry = (a52_stream_carry_t *) malloc (ninputs * sizeof (a52_stream_carry_t));
	    for (i = 0; i < ninputs; i++)
		cap += in[i].fr.n - in[i].next < chunk ? in[i].fr.n - in[i].next : chunk;
	    off = (uint64_t *) malloc ((cap + 1) * sizeof (uint64_t));
	    src = (int *) malloc ((cap + 1) * sizeof (int));
	    left = 0;
	    for (i = 0; i < ninputs; i++) {
		int taken = 0;
		first[i] = nf;
		while (in[i].next < in[i].fr.n && taken < chunk) {
		    const int k = in[i].next++;
		    int len, fl, sr, br;
		    taken++;
		    if ((drv->kind == K_WAV || drv->kind == K_AIF) && !in[i].sink.set_params
			&& in[i].fr.rate[k] != in[i].sink.sample_rate) {
			fprintf (stderr, "error\n");
			in[i].errors++;
			continue;
		    }
		    len = a52_syncinfo (in[i].es.p + in[i].fr.off[k], &fl, &sr, &br);
		    /* frames start on 16-byte boundaries of the packed buffer */
		    while (pack.n & 15)
			bytes_add (&pack, (const uint8_t *) "", 1);
		    off[nf] = pack.n;
		    src[nf] = k;
		    bytes_add (&pack, in[i].es.p + in[i].fr.off[k], len);
		    nf++;
		    /* until the header is out the rate may still change: decode at most up to the first
		     * frame of a new rate, then look again */
		    if ((drv->kind == K_WAV || drv->kind == K_AIF) && in[i].sink.set_params && in[i].next < in[i].fr.n
			&& in[i].fr.rate[in[i].next] != in[i].fr.rate[k])
			break;
		}
		if (in[i].next < in[i].fr.n)
		    left = 1;
		carry[i] = in[i].carry;
	    }
	    first[ninputs] = nf;
	    if (nf) {
		off[nf] = pack.n;
		bytes_add (&pack, (const uint8_t *) "\0\0\0\0\0\0\0\0\0\0\0\0\0\0\0\0", 16);
		pcm = (uint8_t *) malloc (stride * nf);
		status = (int32_t *) malloc (nf * sizeof (int32_t));
		granted = (int32_t *) malloc (nf * sizeof (int32_t));
		if (a52_batch_decode (ctx, pack.p, pack.n - 16, off, nf, first, ninputs, req, (float) gain, drv->bias,
				      disable_dynrng ? A52_DRC_OFF : A52_DRC_STREAM, drv->fmt, pcm, status, granted,
				      carry, NULL, 0, NULL)) {
		    fprintf (stderr, "decode failed: %s\n", a52_batch_last_error (ctx));
		    return 1;
		}
		for (i = 0; i < ninputs; i++) {
		    in[i].carry = carry[i];
		    for (f = first[i]; f < (int) first[i + 1]; f++) {
			const int st = status[f];
			const int nblocks = st == A52_ST_OK ? 6 : st >= A52_ST_BAD_BLOCK ? st - A52_ST_BAD_BLOCK : 0;
			int bad = st != A52_ST_OK;
			bad |= play_frame (&in[i], pcm + stride * f, granted[f], nblocks, in[i].fr.rate[src[f]]);
			if (bad) {
			    fprintf (stderr, "error\n");
			    in[i].errors++;
			} else
			    total_frames++;
		    }
		}
		free (pcm);
		free (status);
		free (granted);
	    }
	    free (pack.p);
	    free (off);
	    free (src);
	    free (first);
	    free (carry);
	}
    }
    gettimeofday (&t1, NULL);
    a52_batch_destroy (ctx);
    {
	const double el = (t1.tv_sec - t0.tv_sec) + (t1.tv_usec - t0.tv_usec) * 1e-6;
	fprintf (stderr, "\n%ld frames decoded in %.2f seconds (%.2f fps)\n", total_frames, el,
		 el > 0 ? total_frames / el : 0.0);
    }
    if (fatal)
	return 1;		/* the reference exits inside the demultiplexer and never finalises the header */
    for (i = 0; i < ninputs; i++)
	close_sink (&in[i]);
    return 0;
}
