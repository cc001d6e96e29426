/* TEST INFRASTRUCTURE — driver around the REFERENCE's own hot-path sources.
 *
 * Built into oracle/_ref/libmsdr_ref.so by oracle/Makefile, which compiles — from where they lie
 * under /root/reference, nothing is copied into this repo —
 *   src/CMSIS_5/arm_fir_fast_q15.c, arm_fir_init_q15.c, arm_copy_q15.c, arm_sqrt_q31.c   (as C)
 *   src/Audio/filter_biquad.cpp                                                          (as C++)
 * with the shims in oracle/ref_shim/.  This file calls those reference functions and restates, line
 * by line, the two pieces of Minimal-SDR.ino:518-775 `demodulation()` that cannot be compiled
 * because they live among Teensy globals: the fs/4 mix (.ino:546-558) and the demodulation
 * switch (.ino:589-628).  It is the parity anchor for oracle/msdr_oracle.c and, through it, for
 * the CUDA path.  Nothing under minimal-sdr_b200/ may link or call this.
 */
#include "arm_math.h"
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define AUDIO_BLOCK_SAMPLES 128

/* stations.h:4  enum { SYNCAM, AM, LSB, USB, CW } */
enum { MODE_SYNCAM = 0, MODE_AM = 1, MODE_LSB = 2, MODE_USB = 3, MODE_CW = 4 };

/* C++ side (ref_driver_biquad.cpp): the reference's AudioFilterBiquad */
void *ref_biquad_new(void);
void ref_biquad_free(void *bq);
void ref_biquad_set_coefficients(void *bq, uint32_t stage, const int32_t *coef);
void ref_biquad_update(void *bq, int16_t *block128);

/* ---------------------------------------------------------------- stage-level entry points */

/* Minimal-SDR.ino:546-558, literally (n must be a multiple of 4) */
void ref_mix_fs4(const int16_t *p_adc, int16_t *I_buffer, int16_t *Q_buffer, uint32_t n)
{
  for (uint32_t i = 0; i < n; i += 4) {
    I_buffer[i]     = p_adc[i + 0];
    I_buffer[i + 1] = 0;
    I_buffer[i + 2] = -p_adc[i + 2];
    I_buffer[i + 3] = 0;

    Q_buffer[i + 0] = 0;
    Q_buffer[i + 1] = p_adc[i + 1];
    Q_buffer[i + 3] = -p_adc[i + 3];
    Q_buffer[i + 2] = 0;
  }
}

typedef struct {
  arm_fir_instance_q15 S;
  q15_t *state;   /* numTaps + blockSize, as .ino:113-114 */
  q15_t *coeffs;  /* the reference borrows the pointer (arm_fir_init_q15.c:103); we own the storage */
  uint32_t blockSize;
} ref_fir;

ref_fir *ref_fir_new(uint16_t numTaps, const int16_t *coeffs, uint32_t blockSize, int *status)
{
  ref_fir *f = (ref_fir *)calloc(1, sizeof(ref_fir));
  f->state = (q15_t *)calloc((size_t)numTaps + blockSize + 8, sizeof(q15_t));
  f->coeffs = (q15_t *)calloc((size_t)numTaps + 8, sizeof(q15_t));
  memcpy(f->coeffs, coeffs, (size_t)numTaps * sizeof(q15_t));
  f->blockSize = blockSize;
  arm_status st = arm_fir_init_q15(&f->S, numTaps, f->coeffs, f->state, blockSize);
  if (status) *status = (int)st;
  return f;
}
void ref_fir_set_coefficients(ref_fir *f, const int16_t *coeffs)
{ /* models the in-place rewrite of the borrowed table, .ino:222 / UI.cpp:337-345 */
  memcpy(f->coeffs, coeffs, (size_t)f->S.numTaps * sizeof(q15_t));
}
void ref_fir_free(ref_fir *f) { if (f) { free(f->state); free(f->coeffs); free(f); } }

/* processes n samples as ceil(n/blockSize) calls of arm_fir_fast_q15 (last call may be short) */
void ref_fir_run(ref_fir *f, const int16_t *src, int16_t *dst, uint32_t n)
{
  uint32_t done = 0;
  while (done < n) {
    uint32_t m = n - done < f->blockSize ? n - done : f->blockSize;
    arm_fir_fast_q15(&f->S, (q15_t *)(src + done), dst + done, m);
    done += m;
  }
}
/* state buffer view: first numTaps-1 entries are the carried history */
const int16_t *ref_fir_state(ref_fir *f) { return f->state; }

int ref_sqrt_q31(int32_t in, int32_t *out) { return (int)arm_sqrt_q31(in, out); }

/* kind: 0 LSB (.ino:591-596), 1 USB (:598-604), 2 AM/CW f32 (:606-616), 3 AM/CW/SYNCAM q31 (:617-627) */
void ref_demod(int kind, const int16_t *I_buffer, const int16_t *Q_buffer, int16_t *p_dac, uint32_t n)
{
  switch (kind) {
  case 0:
    for (uint32_t i = 0; i < n; i++) p_dac[i] = I_buffer[i] - Q_buffer[i];
    break;
  case 1:
    for (uint32_t i = 0; i < n; i++) p_dac[i] = I_buffer[i] + Q_buffer[i];
    break;
  case 2: {
    float32_t audio;
    for (uint32_t i = 0; i < n; i++) {
      arm_sqrt_f32(I_buffer[i] * I_buffer[i] + Q_buffer[i] * Q_buffer[i], &audio);
      p_dac[i] = audio;
    }
    break;
  }
  case 3: {
    q31_t audio;
    for (uint32_t i = 0; i < n; i++) {
      arm_sqrt_q31(I_buffer[i] * I_buffer[i] + Q_buffer[i] * Q_buffer[i], &audio);
      p_dac[i] = audio >> 16;
    }
    break;
  }
  default: break;
  }
}

/* ---------------------------------------------------------------- whole chain, batched */

typedef struct {
  int mode;
  ref_fir *fir_i, *fir_q;
  void *biquad1, *biquad2;   /* biquad1_dac, biquad2_dac  (.ino:71-72) */
} ref_channel;

typedef struct {
  uint32_t n_channels;
  int am_q31;                /* 0: Teensy 3.5/3.6 f32 envelope, 1: Teensy 3.2 q31 envelope */
  ref_channel *ch;
} ref_chain;

ref_chain *ref_chain_new(uint32_t n_channels, int am_q31)
{
  ref_chain *c = (ref_chain *)calloc(1, sizeof(ref_chain));
  c->n_channels = n_channels;
  c->am_q31 = am_q31;
  c->ch = (ref_channel *)calloc(n_channels, sizeof(ref_channel));
  for (uint32_t i = 0; i < n_channels; i++) {
    c->ch[i].mode = MODE_AM;
    c->ch[i].biquad1 = ref_biquad_new();
    c->ch[i].biquad2 = ref_biquad_new();
  }
  return c;
}
void ref_chain_free(ref_chain *c)
{
  if (!c) return;
  for (uint32_t i = 0; i < c->n_channels; i++) {
    ref_fir_free(c->ch[i].fir_i); ref_fir_free(c->ch[i].fir_q);
    ref_biquad_free(c->ch[i].biquad1); ref_biquad_free(c->ch[i].biquad2);
  }
  free(c->ch); free(c);
}
int ref_chain_set_mode(ref_chain *c, uint32_t ch0, uint32_t nch, int mode)
{
  if (ch0 + nch > c->n_channels || mode < 0 || mode > 4) return -1;
  for (uint32_t i = ch0; i < ch0 + nch; i++) c->ch[i].mode = mode;
  return 0;
}
/* init_FIR(): zero state, bind taps (.ino:901-930) */
int ref_chain_fir_init(ref_chain *c, uint32_t ch0, uint32_t nch, uint16_t numTaps, const int16_t *cI, const int16_t *cQ)
{
  if (ch0 + nch > c->n_channels) return -1;
  int st = 0, s1, s2;
  for (uint32_t i = ch0; i < ch0 + nch; i++) {
    ref_fir_free(c->ch[i].fir_i); ref_fir_free(c->ch[i].fir_q);
    c->ch[i].fir_i = ref_fir_new(numTaps, cI, AUDIO_BLOCK_SAMPLES, &s1);
    c->ch[i].fir_q = ref_fir_new(numTaps, cQ, AUDIO_BLOCK_SAMPLES, &s2);
    if (s1) st = s1;
    if (s2) st = s2;
  }
  return st;
}
int ref_chain_fir_set_coefficients(ref_chain *c, uint32_t ch0, uint32_t nch, const int16_t *cI, const int16_t *cQ)
{
  if (ch0 + nch > c->n_channels) return -1;
  for (uint32_t i = ch0; i < ch0 + nch; i++) {
    if (!c->ch[i].fir_i) return -1;
    ref_fir_set_coefficients(c->ch[i].fir_i, cI);
    ref_fir_set_coefficients(c->ch[i].fir_q, cQ);
  }
  return 0;
}
int ref_chain_biquad_set_coefficients(ref_chain *c, int obj, uint32_t ch0, uint32_t nch, uint32_t stage, const int32_t *coef)
{
  if (ch0 + nch > c->n_channels || obj < 0 || obj > 1) return -1;
  for (uint32_t i = ch0; i < ch0 + nch; i++)
    ref_biquad_set_coefficients(obj ? c->ch[i].biquad2 : c->ch[i].biquad1, stage, coef);
  return 0;
}

/* one pass of demodulation() (.ino:518-775) + the two biquad updates of the IRQ graph (CS-2) */
static void ref_channel_block(ref_chain *c, ref_channel *k, const int16_t *p_adc, int16_t *p_dac)
{
  int16_t I_buffer[AUDIO_BLOCK_SAMPLES];
  int16_t Q_buffer[AUDIO_BLOCK_SAMPLES];
  ref_mix_fs4(p_adc, I_buffer, Q_buffer, AUDIO_BLOCK_SAMPLES);
  {
    q15_t I_FIR_out[AUDIO_BLOCK_SAMPLES];
    q15_t Q_FIR_out[AUDIO_BLOCK_SAMPLES];
    arm_fir_fast_q15(&k->fir_i->S, I_buffer, I_FIR_out, AUDIO_BLOCK_SAMPLES);
    arm_fir_fast_q15(&k->fir_q->S, Q_buffer, Q_FIR_out, AUDIO_BLOCK_SAMPLES);
    arm_copy_q15(I_FIR_out, I_buffer, AUDIO_BLOCK_SAMPLES);
    arm_copy_q15(Q_FIR_out, Q_buffer, AUDIO_BLOCK_SAMPLES);
  }
  switch (k->mode) {
  case MODE_LSB: ref_demod(0, I_buffer, Q_buffer, p_dac, AUDIO_BLOCK_SAMPLES); break;
  case MODE_USB: ref_demod(1, I_buffer, Q_buffer, p_dac, AUDIO_BLOCK_SAMPLES); break;
  case MODE_CW:
  case MODE_AM:  ref_demod(c->am_q31 ? 3 : 2, I_buffer, Q_buffer, p_dac, AUDIO_BLOCK_SAMPLES); break;
  case MODE_SYNCAM:
  default:       ref_demod(3, I_buffer, Q_buffer, p_dac, AUDIO_BLOCK_SAMPLES); break; /* only reachable with am_q31 */
  }
  ref_biquad_update(k->biquad1, p_dac);
  ref_biquad_update(k->biquad2, p_dac);
}

/* in/out: [n_channels][stride] int16, n_blocks*128 samples used per row.  Returns threads used. */
int ref_chain_run(ref_chain *c, const int16_t *in, int16_t *out, uint32_t n_blocks, size_t stride, int n_threads)
{
  int used = 1;
#ifdef _OPENMP
  if (n_threads > 0) omp_set_num_threads(n_threads);
  used = omp_get_max_threads();
#pragma omp parallel for schedule(static)
#endif
  for (long i = 0; i < (long)c->n_channels; i++) {
    ref_channel *k = &c->ch[i];
    if (!k->fir_i || !k->fir_q) continue;
    for (uint32_t b = 0; b < n_blocks; b++)
      ref_channel_block(c, k, in + (size_t)i * stride + (size_t)b * AUDIO_BLOCK_SAMPLES,
                        out + (size_t)i * stride + (size_t)b * AUDIO_BLOCK_SAMPLES);
  }
  return used;
}
