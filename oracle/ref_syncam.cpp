/* TEST INFRASTRUCTURE — the REFERENCE's synchronous-AM PLL demodulator: the `case SYNCAM: { ... }` arm of the demodulation switch,
 * Minimal-SDR.ino:631-688, extracted by line range at build time into _ref/syncam_extract.inc (oracle/Makefile) and compiled
 * verbatim inside a switch that supplies the names it refers to.  Its state lives in block-scope statics: one stream per
 * process; tests run every trajectory in a fresh process. */
#include <stdint.h>
#include <math.h>
typedef float float32_t;
#define AUDIO_BLOCK_SAMPLES 128
#define PI 3.1415926535897932384626433832795 /* Arduino.h */
#define _IF 6000                              /* Minimal-SDR.ino:84-85 */
#define SAMPLE_RATE (_IF * 4)
enum { SYNCAM = 0 };                          /* stations.h:4 */

extern "C" void ref_syncam_block(const int16_t *I_buffer, const int16_t *Q_buffer, int16_t *p_dac)
{
  const int mode = SYNCAM;
  switch (mode) {
#include "syncam_extract.inc"
  }
}
