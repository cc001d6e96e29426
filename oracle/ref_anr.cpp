/* TEST INFRASTRUCTURE — the REFERENCE's LMS automatic notch / noise reduction, Minimal-SDR.ino:702-770, extracted by line range
 * at build time into _ref/anr_extract.inc (oracle/Makefile) and compiled verbatim inside a function that supplies the names the
 * fragment refers to (p_dac, ANR_on, AUDIO_BLOCK_SAMPLES, float32_t).  Its state lives in block-scope statics: one stream per
 * process; tests run every trajectory in a fresh process. */
#include <stdint.h>
#include <math.h>
typedef float float32_t;
#define AUDIO_BLOCK_SAMPLES 128
static int ANR_on = 1;

extern "C" void ref_anr_block(int mode, int16_t *p_dac)
{
  ANR_on = mode;
#include "anr_extract.inc"
}
