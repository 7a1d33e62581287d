/* include/a52.h - drop-in public decoder API of the B200 AC-3 engine.
 *
 * Same seven entry points, flag values and sample/level types as liba52's
 * public header (reference: a52dec-0.7.5-cvs/include/a52.h:27-65), so a
 * program written against liba52 (a52dec.c:270-297, AC3ACM.cpp:1478-1578,
 * 2042-2120) links against liba52_b200.so unchanged.  The decode itself runs
 * on the GPU: a52_frame() stages the frame and the first a52_block() of a
 * frame launches the batched kernel with a batch of one frame.  There is no
 * CPU fallback: a52_init() returns NULL when no CUDA device is usable.
 *
 * Differences from the reference header that do not affect callers:
 * <stdint.h> is included here (the reference relies on the caller doing so),
 * and only the float build exists (sample_t = level_t = float, the default
 * and the configuration the ACM wrapper ships, vc++/config.h:75-79).
 */
#ifndef A52_H
#define A52_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef float sample_t;
typedef float level_t;

typedef struct a52_state_s a52_state_t;

/* output channel configurations (request in *flags, granted on return) */
#define A52_CHANNEL       0
#define A52_MONO          1
#define A52_STEREO        2
#define A52_3F            3
#define A52_2F1R          4
#define A52_3F1R          5
#define A52_2F2R          6
#define A52_3F2R          7
#define A52_CHANNEL1      8
#define A52_CHANNEL2      9
#define A52_DOLBY        10
#define A52_CHANNEL_MASK 15

#define A52_LFE          16
#define A52_ADJUST_LEVEL 32

a52_state_t * a52_init (uint32_t mm_accel);
sample_t * a52_samples (a52_state_t * state);
int a52_syncinfo (uint8_t * buf, int * flags, int * sample_rate, int * bit_rate);
int a52_frame (a52_state_t * state, uint8_t * buf, int * flags, level_t * level, sample_t bias);
void a52_dynrng (a52_state_t * state, level_t (* call) (level_t, void *), void * data);
int a52_block (a52_state_t * state);
void a52_free (a52_state_t * state);

#ifdef __cplusplus
}
#endif
#endif /* A52_H */
