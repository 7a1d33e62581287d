/* ac3enc_b200 - WAV to AC-3 on the batched B200 encoder.
 *
 * The reference has no encoder command line: its encoder is reached through the Windows ACM wrapper, which
 * validates the format pair, maps WAVE channel order to coded order and cuts the PCM stream into frames of
 * 1536 samples (src/AC3ACM.cpp:1631-1662 create_channel_map, :1665-1798 stream_convert_pcm, :1890-1936 stream
 * open).  This tool is that front end for files: every input WAV is one stream of the batch, streams of equal
 * (sample rate, bitrate, channels) are encoded together through ac3_batch_encode (), frames are byte-identical
 * to what AC3_encode_frame produces for the same samples.  As in the wrapper, a trailing partial frame is not
 * encoded (it would wait in the gather buffer for input that never comes).
 *
 * usage: ac3enc_b200 [-b <kbps>] [-o <file> | -O <dir>] [-C <frames>] <in.wav> [<in.wav> ...]
 *
 * Host language: C over the C ABI (include/ac3enc_batch.h).  No CPU encoder exists here.
 */
#define _GNU_SOURCE
#include <errno.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>
#include <unistd.h>

#include "ac3enc_batch.h"

typedef struct {
    const char * path;
    uint8_t * raw;
    size_t raw_bytes;
    const int16_t * pcm;	/* interleaved, WAVE channel order */
    long nframes, next;
    int freq, channels, kbps, frame_bytes;
    uint8_t chmap[8];
    ac3_stream_carry_t carry;
    FILE * out;
    long no_fit;
} input_t;

static uint32_t rd32 (const uint8_t * p) { return p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t) p[3] << 24); }
static uint32_t rd16 (const uint8_t * p) { return p[0] | (p[1] << 8); }

static uint8_t * read_all (FILE * f, size_t * n)
{
    size_t cap = 1 << 20, len = 0, got;
    uint8_t * p = (uint8_t *) malloc (cap);
    while (p && (got = fread (p + len, 1, cap - len, f)) > 0) {
	len += got;
	if (len == cap)
	    p = (uint8_t *) realloc (p, cap *= 2);
    }
    *n = len;
    return p;
}

/* RIFF/WAVE, 16-bit PCM, plain or WAVE_FORMAT_EXTENSIBLE; an open-ended data chunk (as a52dec writes to a
 * pipe, libao/audio_out_wav.c:44-58) runs to the end of the file */
static int parse_wav (input_t * in)
{
    const uint8_t * p = in->raw, * end = in->raw + in->raw_bytes;
    int have_fmt = 0;
    if (in->raw_bytes < 12 || memcmp (p, "RIFF", 4) || memcmp (p + 8, "WAVE", 4))
	return -1;
    p += 12;
    while (p + 8 <= end) {
	uint32_t len = rd32 (p + 4);
	const uint8_t * body = p + 8;
	if (!memcmp (p, "fmt ", 4)) {
	    uint32_t tag;
	    if (len < 16 || body + 16 > end)
		return -1;
	    tag = rd16 (body);
	    if (tag == 0xfffe && len >= 40 && body + 40 <= end)
		tag = rd16 (body + 24);		/* SubFormat: KSDATAFORMAT_SUBTYPE_PCM starts with 1 */
	    if (tag != 1 || rd16 (body + 14) != 16)
		return -2;
	    in->channels = rd16 (body + 2);
	    in->freq = rd32 (body + 4);
	    have_fmt = 1;
	} else if (!memcmp (p, "data", 4)) {
	    size_t avail = end - body;
	    if (!have_fmt)
		return -1;
	    if (len > avail)
		len = avail;
	    in->pcm = (const int16_t *) body;
	    in->nframes = len / (1536 * 2 * in->channels);
	    return 0;
	}
	if (len > (size_t) (end - body))
	    break;
	p = body + len + (len & 1);
    }
    return -1;
}

int main (int argc, char ** argv)
{
    int c, i, ninputs, kbps = 0, chunk = 1024;
    const char * out_path = NULL, * out_dir = NULL;
    char * s;
    input_t * in;
    ac3_batch_t * ctx;
    long total = 0;
    struct timeval t0, t1;

    while ((c = getopt (argc, argv, "b:o:O:C:h")) != -1)
	switch (c) {
	case 'b':
	    kbps = strtol (optarg, &s, 0);
	    if (*s)
		kbps = -1;
	    break;
	case 'o':
	    out_path = optarg;
	    break;
	case 'O':
	    out_dir = optarg;
	    break;
	case 'C':
	    chunk = strtol (optarg, &s, 0);
	    if (chunk < 1 || *s)
		chunk = 0;
	    break;
	default:
	    chunk = 0;
	}
    ninputs = argc - optind;
    if (!chunk || ninputs < 1 || (ninputs > 1 && !out_dir) || kbps < 0) {
	fprintf (stderr, "usage: %s [-b <kbps>] [-o <file> | -O <dir>] [-C <frames>] <in.wav> [<in.wav> ...]\n"
		 "\t-b\tbitrate: 32 40 48 56 64 80 96 112 128 160 192 224 256 320 384 448 512 576 640\n"
		 "\t\t(default 192 up to two channels, 448 above)\n"
		 "\t-o\toutput file (one input; default stdout)\n"
		 "\t-O\toutput directory: all inputs are encoded as one batch, <dir>/<name>.ac3 each\n"
		 "\t-C\tframes of a stream per engine call (default 1024)\n", argv[0]);
	return 1;
    }
    in = (input_t *) calloc (ninputs, sizeof (input_t));
    for (i = 0; i < ninputs; i++) {
	FILE * f = fopen (argv[optind + i], "rb");
	int rc;
	in[i].path = argv[optind + i];
	if (!f) {
	    fprintf (stderr, "%s - could not open file %s\n", strerror (errno), in[i].path);
	    return 1;
	}
	in[i].raw = read_all (f, &in[i].raw_bytes);
	fclose (f);
	rc = in[i].raw ? parse_wav (&in[i]) : -1;
	if (rc) {
	    fprintf (stderr, "%s: %s\n", in[i].path, rc == -2 ? "only 16-bit PCM is encoded" : "not a RIFF/WAVE file");
	    return 1;
	}
	/* what the ACM stream open checks (AC3ACM.cpp:1890-1936, 1940-1943) */
	in[i].kbps = kbps ? kbps : (in[i].channels <= 2 ? 192 : 448);
	if (in[i].freq < 32000) {
	    fprintf (stderr, "%s: sample rates below 32000 Hz are not encoded\n", in[i].path);
	    return 1;
	}
	if (ac3_acm_bitrate (in[i].freq, 125u * in[i].kbps) != in[i].kbps) {
	    fprintf (stderr, "%s: %d kb/s is not an AC-3 bitrate\n", in[i].path, in[i].kbps);
	    return 1;
	}
	if (ac3_wav_channel_map (in[i].channels, in[i].chmap)
	    || !(in[i].frame_bytes = ac3_batch_frame_bytes (in[i].freq, in[i].kbps * 1000, in[i].channels))) {
	    fprintf (stderr, "%s: %d Hz, %d channels, %d kb/s: the encoder refuses this format\n", in[i].path,
		     in[i].freq, in[i].channels, in[i].kbps);
	    return 1;
	}
	if (out_dir) {
	    char path[4096];
	    const char * base = strrchr (in[i].path, '/');
	    base = base ? base + 1 : in[i].path;
	    snprintf (path, sizeof (path), "%s/%s.ac3", out_dir, base);
	    in[i].out = fopen (path, "wb");
	} else
	    in[i].out = out_path ? fopen (out_path, "wb") : stdout;
	if (!in[i].out) {
	    fprintf (stderr, "%s - could not open the output for %s\n", strerror (errno), in[i].path);
	    return 1;
	}
    }
    ctx = ac3_batch_create (0);
    if (!ctx) {
	fprintf (stderr, "AC-3 encoder init failed (no CUDA device)\n");
	return 1;
    }
    gettimeofday (&t0, NULL);
    for (;;) {
	/* one round: the streams of one format that still have frames, the same number of frames each */
	int lead = -1, n = 0, k;
	long nfr = chunk;
	int * who = (int *) malloc (ninputs * sizeof (int));
	for (i = 0; i < ninputs; i++) {
	    if (in[i].next >= in[i].nframes)
		continue;
	    if (lead < 0)
		lead = i;
	    if (in[i].freq != in[lead].freq || in[i].kbps != in[lead].kbps || in[i].channels != in[lead].channels)
		continue;
	    who[n++] = i;
	    if (in[i].nframes - in[i].next < nfr)
		nfr = in[i].nframes - in[i].next;
	}
	if (!n) {
	    free (who);
	    break;
	}
	{
	    const int ch = in[lead].channels, fb = in[lead].frame_bytes;
	    const size_t per = (size_t) nfr * 1536 * ch;
	    int16_t * pcm = (int16_t *) malloc (per * n * sizeof (int16_t));
	    uint8_t * out = (uint8_t *) malloc ((size_t) n * nfr * fb + 8);
	    int32_t * status = (int32_t *) malloc ((size_t) n * nfr * sizeof (int32_t));
	    ac3_stream_carry_t * carry = (ac3_stream_carry_t *) malloc (n * sizeof (ac3_stream_carry_t));
	    for (k = 0; k < n; k++) {
		memcpy (pcm + per * k, in[who[k]].pcm + (size_t) in[who[k]].next * 1536 * ch, per * sizeof (int16_t));
		carry[k] = in[who[k]].carry;
	    }
	    if (ac3_batch_encode (ctx, pcm, n, (int) nfr, in[lead].freq, in[lead].kbps * 1000, ch, in[lead].chmap, out,
				  status, carry, NULL, 0, NULL)) {
		fprintf (stderr, "encode failed: %s\n", ac3_batch_last_error (ctx));
		return 1;
	    }
	    for (k = 0; k < n; k++) {
		long f;
		input_t * s1 = &in[who[k]];
		s1->carry = carry[k];
		s1->next += nfr;
		fwrite (out + (size_t) k * nfr * fb, fb, nfr, s1->out);
		for (f = 0; f < nfr; f++)
		    s1->no_fit += status[k * nfr + f] == AC3_ST_NO_FIT;
		total += nfr;
	    }
	    free (pcm);
	    free (out);
	    free (status);
	    free (carry);
	}
	free (who);
    }
    gettimeofday (&t1, NULL);
    ac3_batch_destroy (ctx);
    for (i = 0; i < ninputs; i++) {
	if (in[i].no_fit)
	    fprintf (stderr, "%s: %ld frames had no fitting SNR offset at %d kb/s\n", in[i].path, in[i].no_fit, in[i].kbps);
	if (in[i].out != stdout)
	    fclose (in[i].out);
    }
    {
	const double el = (t1.tv_sec - t0.tv_sec) + (t1.tv_usec - t0.tv_usec) * 1e-6;
	fprintf (stderr, "%ld frames encoded in %.2f seconds (%.2f fps)\n", total, el, el > 0 ? total / el : 0.0);
    }
    return 0;
}
