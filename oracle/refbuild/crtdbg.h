/* shim for MSVC <crtdbg.h>: the shipped release build defines NDEBUG, which
 * makes _ASSERT a no-op; mirror that (a live assert aborts on stereo input,
 * SURVEY.md section 7 hard part 11). */
#ifndef _ASSERT
#define _ASSERT(x) ((void)0)
#endif
