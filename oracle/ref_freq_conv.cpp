/* TEST INFRASTRUCTURE — extern "C" driver around the REFERENCE's AudioEffectFreqConv (freq_conv.{h,cpp}, compiled where it lies
 * under /root/reference against oracle/ref_shim/stubs_fc).  The class is dead code in the sketch and its oscillator tables are
 * defined nowhere (freq_conv.h:33-34): they are defined here and filled per call. */
#include "freq_conv.h"
#include <string.h>

q15_t Osc_Q_buffer_i[AUDIO_BLOCK_SAMPLES];
q15_t Osc_I_buffer_i[AUDIO_BLOCK_SAMPLES];
int AudioStream::fail_alloc = 0;

extern "C" {
/* same shape as orc_freq_conv: n samples (a multiple of 128) in place on I and Q, oscI/oscQ of n entries.
 * have_I / have_Q = 0 models a missing input block; fail = 1 an allocation failure: the object then transmits nothing
 * (freq_conv.cpp:40-47,64,111) and the function returns 0 for that block; otherwise 1. */
int ref_freq_conv_ex(int dir, int pass, int16_t *I, int16_t *Q, const int16_t *oscI, const int16_t *oscQ, uint32_t n, int have_I, int have_Q, int fail)
{
  AudioEffectFreqConv fc;
  fc.direction(dir != 0);
  fc.passthrough(pass != 0);
  int all = 1;
  for (uint32_t b = 0; b + AUDIO_BLOCK_SAMPLES <= n; b += AUDIO_BLOCK_SAMPLES) {
    audio_block_t bi, bq;
    memcpy(bi.data, I + b, sizeof(bi.data));
    memcpy(bq.data, Q + b, sizeof(bq.data));
    memcpy(Osc_I_buffer_i, oscI + b, sizeof(Osc_I_buffer_i));
    memcpy(Osc_Q_buffer_i, oscQ + b, sizeof(Osc_Q_buffer_i));
    fc.in_slot[0] = have_I ? &bi : NULL;
    fc.in_slot[1] = have_Q ? &bq : NULL;
    fc.out_slot[0] = fc.out_slot[1] = NULL;
    AudioStream::fail_alloc = fail;
    fc.update();
    AudioStream::fail_alloc = 0;
    if (fc.out_slot[0] && fc.out_slot[1]) {
      memcpy(I + b, fc.out_slot[0]->data, sizeof(bi.data));
      memcpy(Q + b, fc.out_slot[1]->data, sizeof(bq.data));
    } else all = 0;
  }
  return all;
}
void ref_freq_conv(int dir, int pass, int16_t *I, int16_t *Q, const int16_t *oscI, const int16_t *oscQ, uint32_t n)
{
  ref_freq_conv_ex(dir, pass, I, Q, oscI, oscQ, n, 1, 1, 0);
}
}
