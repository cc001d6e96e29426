/* TEST INFRASTRUCTURE — extern "C" handle around the REFERENCE's AudioFilterBiquad
 * (src/Audio/filter_biquad.{h,cpp}, compiled from /root/reference by oracle/Makefile against the
 * stubs in oracle/ref_shim/).  Also exposes the double-precision coefficient designers of
 * filter_biquad.h:44-149 so the host-side designers can be checked against them. */
#include "filter_biquad.h"
#include <string.h>

namespace {
struct RefBiquad : public AudioFilterBiquad {
  audio_block_t blk;
};
}

extern "C" {
void *ref_biquad_new(void) { return new RefBiquad(); }
void ref_biquad_free(void *p) { delete static_cast<RefBiquad *>(p); }
void ref_biquad_set_coefficients(void *p, uint32_t stage, const int32_t *coef)
{
  static_cast<RefBiquad *>(p)->setCoefficients(stage, reinterpret_cast<const int *>(coef));
}
void ref_biquad_set_coefficients_double(void *p, uint32_t stage, const double *coef)
{
  static_cast<RefBiquad *>(p)->setCoefficients(stage, coef);
}
/* kind: 0 lowpass 1 highpass 2 bandpass 3 notch 4 lowshelf 5 highshelf; p2 = q (0..3) or gain (4,5); p3 = slope */
void ref_biquad_design(void *p, int kind, uint32_t stage, float frequency, float p2, float p3)
{
  RefBiquad *b = static_cast<RefBiquad *>(p);
  switch (kind) {
  case 0: b->setLowpass(stage, frequency, p2); break;
  case 1: b->setHighpass(stage, frequency, p2); break;
  case 2: b->setBandpass(stage, frequency, p2); break;
  case 3: b->setNotch(stage, frequency, p2); break;
  case 4: b->setLowShelf(stage, frequency, p2, p3); break;
  case 5: b->setHighShelf(stage, frequency, p2, p3); break;
  default: break;
  }
}
/* one AudioStream tick: 128 samples in place (filter_biquad.cpp:33-82) */
void ref_biquad_update(void *p, int16_t *block128)
{
  RefBiquad *b = static_cast<RefBiquad *>(p);
  memcpy(b->blk.data, block128, sizeof(b->blk.data));
  b->in_slot[0] = &b->blk;
  b->out_slot[0] = NULL;
  b->update();
  if (b->out_slot[0]) memcpy(block128, b->out_slot[0]->data, sizeof(b->blk.data));
}
/* the private `int32_t definition[32]` (filter_biquad.h:152) sits right after the AudioStream base;
 * we read it through a layout-compatible view for state round-trip tests. */
void ref_biquad_get_definition(void *p, int32_t *out32)
{
  struct View : public AudioStream { View() : AudioStream(1, NULL) {} void update() {} int32_t definition[32]; };
  RefBiquad *b = static_cast<RefBiquad *>(p);
  const View *v = reinterpret_cast<const View *>(static_cast<AudioFilterBiquad *>(b));
  memcpy(out32, v->definition, sizeof(int32_t) * 32);
}
void ref_biquad_set_definition(void *p, const int32_t *in32)
{
  struct View : public AudioStream { View() : AudioStream(1, NULL) {} void update() {} int32_t definition[32]; };
  RefBiquad *b = static_cast<RefBiquad *>(p);
  View *v = reinterpret_cast<View *>(static_cast<AudioFilterBiquad *>(b));
  memcpy(v->definition, in32, sizeof(int32_t) * 32);
}
double ref_audio_sample_rate_exact(void) { return AUDIO_SAMPLE_RATE_EXACT; }
}
