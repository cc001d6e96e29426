/* TEST INFRASTRUCTURE — CPU oracle for the Minimal-SDR receive DSP chain.
 *
 * A plain-C RESTATEMENT (written from the algorithm, not copied) of the reference's hot path:
 *   fs/4 mix -> Q15 FIR pair -> SSB / AM demodulation -> fixed-point biquad cascade.
 * Each function cites the reference file:line it follows.  Pinned against the reference itself:
 * oracle/_ref/libmsdr_ref.so compiles the reference's own arm_fir_fast_q15.c / arm_fir_init_q15.c /
 * arm_copy_q15.c / arm_sqrt_q31.c / filter_biquad.cpp (see oracle/Makefile); tests/test_oracle_vs_ref.py
 * checks this file against it bit-for-bit, and tests/golden/ holds vectors generated from it
 * (tests/golden/make_golden.py).  The reference repository ships no tests or golden vectors of its own.
 * orc_freq_conv: the reference's freq_conv.cpp calls arm_mult_q15/arm_add_q15/arm_sub_q15, which are NOT vendored in
 * the reference (CMSIS-DSP V1.5.1, prebuilt on Teensy).  It is pinned against the reference's own freq_conv.cpp compiled
 * with those three routines built from the saturating primitives the reference DOES vendor as portable C
 * (clip_q31_to_q15, __QADD16, __QSUB16: oracle/ref_q15_prims.c, oracle/ref_freq_conv.cpp); what stays unpinned is the
 * three one-line element loops around those primitives, restated from the CMSIS documentation.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this
 * library.  The product (minimal-sdr_b200/) never links or calls it and has no CPU fallback.
 *
 * Build: gcc -O2 -fwrapv -fno-strict-aliasing -fopenmp -shared -fPIC  (oracle/Makefile).
 * All int32 arithmetic below that may overflow is done in uint32_t so it wraps like the Cortex-M4.
 */
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_BLOCK 128 /* AUDIO_BLOCK_SAMPLES (Teensy core; implied by Minimal-SDR.ino:113-114,525-526) */

/* stations.h:4 */
enum { ORC_SYNCAM = 0, ORC_AM = 1, ORC_LSB = 2, ORC_USB = 3, ORC_CW = 4 };

static inline int32_t orc_ssat16(int32_t v) { return v > 32767 ? 32767 : (v < -32768 ? -32768 : v); }

/* ------------------------------------------------------------------------------------------------
 * A1  fs/4 mix.  Minimal-SDR.ino:546-558:  I = {x,0,-x,0}, Q = {0,x,0,-x} by n mod 4; the negation
 * is computed in int and narrowed back to int16, so -(-32768) stays -32768.
 */
static inline int16_t orc_neg16(int16_t x) { return (int16_t)(uint16_t)(0u - (uint32_t)(int32_t)x); }

void orc_mix_fs4(const int16_t *x, int16_t *I, int16_t *Q, uint32_t n)
{
  for (uint32_t i = 0; i < n; i++) {
    switch (i & 3u) {
    case 0: I[i] = x[i];            Q[i] = 0;                break;
    case 1: I[i] = 0;               Q[i] = x[i];             break;
    case 2: I[i] = orc_neg16(x[i]); Q[i] = 0;                break;
    default: I[i] = 0;              Q[i] = orc_neg16(x[i]);  break;
    }
  }
}

/* ------------------------------------------------------------------------------------------------
 * A2/A3  Q15 FIR.  arm_fir_init_q15.c:78-138 (even numTaps >= 4 else ARGUMENT_ERROR, state zeroed,
 * coefficient pointer kept) and arm_fir_fast_q15.c:60-329.  With s = hist[T-1] || block:
 *   acc  = sum_{k<T} c[k]*s[n+k]   in a 32-bit register that WRAPS (SMLAD/SMLADX, :132-183)
 *   y[n] = ssat16(acc >> 15)       (:234-238), then hist <- last T-1 of s (:296-327).
 * The 4-way unrolled loop, the numTaps%4==2 tail (:195-227) and the blockSize%4 tail (:260-294)
 * all compute this same sum; mod-2^32 addition is order independent, so direct form is exact.
 */
typedef struct {
  uint16_t numTaps;
  int16_t *coeffs; /* owned copy; orc_fir_set_coefficients models the reference's in-place rewrite */
  int16_t *hist;   /* numTaps-1 carried samples */
} orc_fir;

orc_fir *orc_fir_new(uint16_t numTaps, const int16_t *coeffs, uint32_t blockSize, int *status)
{
  (void)blockSize;
  orc_fir *f = (orc_fir *)calloc(1, sizeof(orc_fir));
  f->numTaps = numTaps;
  f->coeffs = (int16_t *)calloc((size_t)numTaps + 1, sizeof(int16_t));
  f->hist = (int16_t *)calloc((size_t)numTaps + 1, sizeof(int16_t));
  memcpy(f->coeffs, coeffs, (size_t)numTaps * sizeof(int16_t));
  /* arm_fir_init_q15.c:93-96: "numTaps & 1" -> ARM_MATH_ARGUMENT_ERROR (-1).  (The header comment asks
   * for >= 4 taps, the code only rejects odd counts.) */
  if (status) *status = (numTaps & 1u) ? -1 : 0;
  return f;
}
void orc_fir_set_coefficients(orc_fir *f, const int16_t *coeffs) { memcpy(f->coeffs, coeffs, (size_t)f->numTaps * sizeof(int16_t)); }
void orc_fir_free(orc_fir *f) { if (f) { free(f->coeffs); free(f->hist); free(f); } }
const int16_t *orc_fir_state(orc_fir *f) { return f->hist; }

void orc_fir_run(orc_fir *f, const int16_t *src, int16_t *dst, uint32_t n)
{
  enum { CHUNK = 1024 };
  const uint32_t T = f->numTaps, H = T - 1u;
  int16_t stackbuf[CHUNK + 1024];
  int16_t *s = (H + CHUNK <= sizeof(stackbuf) / sizeof(stackbuf[0])) ? stackbuf : (int16_t *)malloc(((size_t)H + CHUNK) * sizeof(int16_t));
  for (uint32_t done = 0; done < n; done += CHUNK) {
    const uint32_t m = (n - done < CHUNK) ? n - done : CHUNK;
    memcpy(s, f->hist, (size_t)H * sizeof(int16_t));
    memcpy(s + H, src + done, (size_t)m * sizeof(int16_t));
    for (uint32_t i = 0; i < m; i++) {
      uint32_t acc = 0;
      for (uint32_t k = 0; k < T; k++) acc += (uint32_t)((int32_t)f->coeffs[k] * (int32_t)s[i + k]);
      dst[done + i] = (int16_t)orc_ssat16((int32_t)acc >> 15);
    }
    memcpy(f->hist, s + m, (size_t)H * sizeof(int16_t));
  }
  if (s != stackbuf) free(s);
}

/* ------------------------------------------------------------------------------------------------
 * A5c  arm_sqrt_q31.c:50-138.  CLZ normalisation, float initial guess via 0x5f3759df, three Newton
 * steps in Q31 with 64-bit products, rescale.  in <= 0 -> 0 and ARM_MATH_ARGUMENT_ERROR.
 */
int orc_sqrt_q31(int32_t in, int32_t *pOut)
{
  if (in <= 0) { *pOut = 0; return -1; }
  int32_t signBits = (int32_t)__builtin_clz((uint32_t)in) - 1;
  int32_t sh = (signBits % 2 == 0) ? signBits : signBits - 1;
  int32_t number = (int32_t)((uint32_t)in << sh);
  int32_t half = number >> 1;
  int32_t temp1 = number;
  union { int32_t i; float f; } cv;
  cv.f = (float)number * 4.6566128731e-010f;
  cv.i = 0x5f3759df - (cv.i >> 1);
  int32_t var1 = (int32_t)(cv.f * 1073741824.0f);
  for (int it = 0; it < 3; it++) {
    int32_t sq = (int32_t)(((int64_t)var1 * var1) >> 31);
    int32_t t = (int32_t)(((int64_t)sq * (int64_t)half) >> 31);
    int32_t d = (int32_t)(0x30000000u - (uint32_t)t);
    var1 = (int32_t)((uint32_t)(int32_t)(((int64_t)var1 * d) >> 31) << 2);
  }
  var1 = (int32_t)((uint32_t)(int32_t)(((int64_t)temp1 * var1) >> 31) << 1);
  var1 = var1 >> (sh / 2);
  *pOut = var1;
  return 0;
}

/* ------------------------------------------------------------------------------------------------
 * A5  demodulation switch, Minimal-SDR.ino:589-628.
 *   kind 0 LSB  (:591-596)  out = (int16)(I - Q)           plain int add, narrowing store wraps
 *   kind 1 USB  (:598-604)  out = (int16)(I + Q)
 *   kind 2 AM/CW f32 (:606-616, arm_math.h:5733-5760)  s = I*I+Q*Q (int32, wraps only for I=Q=-32768);
 *          r = s >= 0 ? sqrtf((float)s) : 0;  out = (int16)(int32)r   (truncate, then keep low 16 bits)
 *   kind 3 AM/CW/SYNCAM q31 (:617-627)  out = (int16)(arm_sqrt_q31(s) >> 16)
 */
void orc_demod(int kind, const int16_t *I, const int16_t *Q, int16_t *out, uint32_t n)
{
  for (uint32_t i = 0; i < n; i++) {
    const int32_t a = I[i], b = Q[i];
    switch (kind) {
    case 0: out[i] = (int16_t)(uint16_t)(uint32_t)(a - b); break;
    case 1: out[i] = (int16_t)(uint16_t)(uint32_t)(a + b); break;
    case 2: {
      int32_t s = (int32_t)((uint32_t)(a * a) + (uint32_t)(b * b));
      float fs = (float)s;
      float r = (fs >= 0.0f) ? sqrtf(fs) : 0.0f;
      out[i] = (int16_t)(uint16_t)(uint32_t)(int32_t)r;
      break;
    }
    case 3: {
      int32_t s = (int32_t)((uint32_t)(a * a) + (uint32_t)(b * b));
      int32_t r;
      orc_sqrt_q31(s, &r);
      out[i] = (int16_t)(uint16_t)(uint32_t)(r >> 16);
      break;
    }
    default: break;
    }
  }
}

/* ------------------------------------------------------------------------------------------------
 * A7/A8  AudioFilterBiquad.  State `definition[32]` (filter_biquad.h:152): 4 stages x
 *   {b0, b1, b2, -a1, -a2, bprev = x[n-1]:x[n-2], aprev = y[n-1]:y[n-2], residual | flag(bit31)}.
 * update() (filter_biquad.cpp:33-82), per stage and sample:
 *   sum = residual + sum_i trunc32((int64)coef_i * v_i >> 16)     five SMLAWB/SMLAWT, adds wrap
 *   y   = ssat16(sum >> 14)                                       (ssat #16, asr #14; dspinst.h:33-51)
 *   residual = sum & 0x3FFF
 * Stages run stage-major over the 128-sample block, the next stage runs iff bit31 of word 7 is set.
 * setCoefficients (filter_biquad.cpp:84-100): stage >= 4 ignored; sets the flag on the PREVIOUS stage;
 * stores b0,b1,b2,-a1,-a2; keeps x/y history; clears the residual but keeps this stage's flag.
 */
typedef struct { int32_t definition[32]; } orc_biquad;

orc_biquad *orc_biquad_new(void) { return (orc_biquad *)calloc(1, sizeof(orc_biquad)); } /* filter_biquad.h:36-39 */
void orc_biquad_free(orc_biquad *b) { free(b); }
void orc_biquad_get_definition(orc_biquad *b, int32_t *out32) { memcpy(out32, b->definition, sizeof(b->definition)); }
void orc_biquad_set_definition(orc_biquad *b, const int32_t *in32) { memcpy(b->definition, in32, sizeof(b->definition)); }

void orc_biquad_set_coefficients(orc_biquad *b, uint32_t stage, const int32_t *coef)
{
  if (stage >= 4) return;
  int32_t *dest = b->definition + (stage << 3);
  if (stage > 0) dest[-1] = (int32_t)((uint32_t)dest[-1] | 0x80000000u);
  dest[0] = coef[0];
  dest[1] = coef[1];
  dest[2] = coef[2];
  dest[3] = (int32_t)(0u - (uint32_t)coef[3]);
  dest[4] = (int32_t)(0u - (uint32_t)coef[4]);
  dest[7] = (int32_t)((uint32_t)dest[7] & 0x80000000u);
}

static inline uint32_t orc_smulw(int32_t c, int32_t v16) { return (uint32_t)(int32_t)(((int64_t)c * (int64_t)v16) >> 16); }

void orc_biquad_update(orc_biquad *b, int16_t *block, uint32_t n /* even; 128 in the reference */)
{
  int32_t *st = b->definition;
  uint32_t flag;
  do {
    const int32_t b0 = st[0], b1 = st[1], b2 = st[2], a1 = st[3], a2 = st[4];
    int32_t x1 = (int16_t)((uint32_t)st[5] >> 16), x2 = (int16_t)((uint32_t)st[5] & 0xFFFF);
    int32_t y1 = (int16_t)((uint32_t)st[6] >> 16), y2 = (int16_t)((uint32_t)st[6] & 0xFFFF);
    uint32_t sum = (uint32_t)st[7] & 0x3FFFu;
    for (uint32_t i = 0; i < n; i++) {
      const int32_t x0 = block[i];
      sum += orc_smulw(b0, x0) + orc_smulw(b1, x1) + orc_smulw(b2, x2) + orc_smulw(a1, y1) + orc_smulw(a2, y2);
      const int32_t y0 = orc_ssat16((int32_t)sum >> 14);
      sum &= 0x3FFFu;
      x2 = x1; x1 = x0; y2 = y1; y1 = y0;
      block[i] = (int16_t)y0;
    }
    flag = (uint32_t)st[7] & 0x80000000u;
    st[7] = (int32_t)(sum | flag);
    st[6] = (int32_t)(((uint32_t)y1 << 16) | ((uint32_t)y2 & 0xFFFFu));
    st[5] = (int32_t)(((uint32_t)x1 << 16) | ((uint32_t)x2 & 0xFFFFu));
    st += 8;
  } while (flag);
}

/* ------------------------------------------------------------------------------------------------
 * A6  AudioEffectFreqConv::update, freq_conv.cpp:30-116 (dead code in the sketch).  arm_mult_q15 = ssat16((a*b)>>15),
 * arm_add_q15 / arm_sub_q15 saturating (CMSIS-DSP V1.5.1 docs); checked against the reference's own class compiled with the
 * vendored primitives (tests/test_oracle_vs_ref.py::test_freq_conv_matches_reference_class; see the file header).
 * pass == 0 forwards the inputs unchanged (freq_conv.cpp:49-56; note the inverted naming).
 */
static inline int16_t orc_mult_q15(int16_t a, int16_t b) { return (int16_t)orc_ssat16(((int32_t)a * b) >> 15); }

void orc_freq_conv(int dir, int pass, int16_t *I, int16_t *Q, const int16_t *oscI, const int16_t *oscQ, uint32_t n)
{
  if (!pass) return;
  for (uint32_t i = 0; i < n; i++) {
    const int16_t vi = I[i], vq = Q[i];
    if (!dir) {
      const int32_t A = orc_mult_q15(vi, oscQ[i]), B = orc_mult_q15(vq, oscI[i]);
      const int32_t C = orc_mult_q15(vq, oscQ[i]), D = orc_mult_q15(vi, oscI[i]);
      I[i] = (int16_t)orc_ssat16(A + B);
      Q[i] = (int16_t)orc_ssat16(C - D);
    } else {
      const int32_t A = orc_mult_q15(vq, oscQ[i]), B = orc_mult_q15(vi, oscI[i]);
      const int32_t C = orc_mult_q15(vi, oscQ[i]), D = orc_mult_q15(vq, oscI[i]);
      Q[i] = (int16_t)orc_ssat16(A + B);
      I[i] = (int16_t)orc_ssat16(C - D);
    }
  }
}

/* ------------------------------------------------------------------------------------------------
 * Whole chain, batched over independent channels: demodulation() (Minimal-SDR.ino:518-775, CS-1) followed
 * by biquad1_dac.update(), biquad2_dac.update() of the IRQ graph (CS-2; .ino:71-72,77,79-80).
 */
typedef struct {
  int mode;
  orc_fir *fir_i, *fir_q;
  orc_biquad bq[2];
  float pll[3]; /* SYNCAM PLL: fil_out, omega2, phzerror (.ino:643-645) */
  int anr_on;   /* ANR_on (.ino:99): 0 off, 1 notch, 2 noise reduction */
  void *anr;    /* LMS state, allocated when first switched on */
} orc_channel;
/* the LMS block of demodulation() (.ino:702-770), between the demodulation switch and the DAC queue; defined below */
void *orc_anr_alloc1(void);
void orc_anr_block_v(void *state, int mode, int16_t *p_dac, uint32_t n);
/* `case SYNCAM` on the Teensy 3.5/3.6 (f32 PLL, .ino:631-688); defined with the other next-row restatements below */
void orc_syncam_block3(float *state3, const int16_t *I_buffer, const int16_t *Q_buffer, int16_t *p_dac, uint32_t n);

typedef struct {
  uint32_t n_channels;
  int am_q31;
  orc_channel *ch;
} orc_chain;

orc_chain *orc_chain_new(uint32_t n_channels, int am_q31)
{
  orc_chain *c = (orc_chain *)calloc(1, sizeof(orc_chain));
  c->n_channels = n_channels;
  c->am_q31 = am_q31;
  c->ch = (orc_channel *)calloc(n_channels, sizeof(orc_channel));
  for (uint32_t i = 0; i < n_channels; i++) c->ch[i].mode = ORC_AM; /* .ino:98 */
  return c;
}
void orc_chain_free(orc_chain *c)
{
  if (!c) return;
  for (uint32_t i = 0; i < c->n_channels; i++) { orc_fir_free(c->ch[i].fir_i); orc_fir_free(c->ch[i].fir_q); free(c->ch[i].anr); }
  free(c->ch); free(c);
}
int orc_chain_set_mode(orc_chain *c, uint32_t ch0, uint32_t nch, int mode)
{
  if (ch0 + nch > c->n_channels || mode < 0 || mode > 4) return -1;
  for (uint32_t i = ch0; i < ch0 + nch; i++) c->ch[i].mode = mode;
  return 0;
}
int orc_chain_fir_init(orc_chain *c, uint32_t ch0, uint32_t nch, uint16_t numTaps, const int16_t *cI, const int16_t *cQ)
{
  if (ch0 + nch > c->n_channels) return -1;
  int st = 0, s1, s2;
  for (uint32_t i = ch0; i < ch0 + nch; i++) {
    orc_fir_free(c->ch[i].fir_i); orc_fir_free(c->ch[i].fir_q);
    c->ch[i].fir_i = orc_fir_new(numTaps, cI, ORC_BLOCK, &s1);
    c->ch[i].fir_q = orc_fir_new(numTaps, cQ, ORC_BLOCK, &s2);
    if (s1) st = s1;
    if (s2) st = s2;
  }
  return st;
}
int orc_chain_fir_set_coefficients(orc_chain *c, uint32_t ch0, uint32_t nch, const int16_t *cI, const int16_t *cQ)
{
  if (ch0 + nch > c->n_channels) return -1;
  for (uint32_t i = ch0; i < ch0 + nch; i++) {
    if (!c->ch[i].fir_i) return -1;
    orc_fir_set_coefficients(c->ch[i].fir_i, cI);
    orc_fir_set_coefficients(c->ch[i].fir_q, cQ);
  }
  return 0;
}
int orc_chain_biquad_set_coefficients(orc_chain *c, int obj, uint32_t ch0, uint32_t nch, uint32_t stage, const int32_t *coef)
{
  if (ch0 + nch > c->n_channels || obj < 0 || obj > 1) return -1;
  for (uint32_t i = ch0; i < ch0 + nch; i++) orc_biquad_set_coefficients(&c->ch[i].bq[obj], stage, coef);
  return 0;
}
/* state access for checkpoint/resume tests: raw FIR history is reconstructed from the I/Q delay lines
 * (even positions live in the I line, odd ones in the Q line; samples at n%4 in {2,3} were negated). */
int orc_chain_set_anr(orc_chain *c, uint32_t ch0, uint32_t nch, int anr_on)
{
  if ((uint64_t)ch0 + nch > c->n_channels || anr_on < 0 || anr_on > 2) return -1;
  for (uint32_t i = ch0; i < ch0 + nch; i++) {
    if (anr_on && !c->ch[i].anr) c->ch[i].anr = orc_anr_alloc1();
    c->ch[i].anr_on = anr_on;
  }
  return 0;
}
void orc_chain_get_biquad_definition(orc_chain *c, uint32_t ch, int obj, int32_t *out32) { memcpy(out32, c->ch[ch].bq[obj].definition, 128); }

static int orc_demod_kind(const orc_chain *c, int mode)
{
  switch (mode) {
  case ORC_LSB: return 0;
  case ORC_USB: return 1;
  case ORC_CW:
  case ORC_AM:  return c->am_q31 ? 3 : 2;
  default:      return 3; /* SYNCAM shares the q31 envelope on Teensy 3.2 (.ino:618-620); f32 PLL is a "next" row */
  }
}

int orc_chain_run(orc_chain *c, const int16_t *in, int16_t *out, uint32_t n_blocks, size_t stride, int n_threads)
{
  int used = 1;
#ifdef _OPENMP
  if (n_threads > 0) omp_set_num_threads(n_threads);
  used = omp_get_max_threads();
#pragma omp parallel for schedule(static)
#endif
  for (long i = 0; i < (long)c->n_channels; i++) {
    orc_channel *k = &c->ch[i];
    if (!k->fir_i || !k->fir_q) continue;
    int16_t I[ORC_BLOCK], Q[ORC_BLOCK], If[ORC_BLOCK], Qf[ORC_BLOCK];
    const int kind = orc_demod_kind(c, k->mode);
    for (uint32_t b = 0; b < n_blocks; b++) {
      const int16_t *p_adc = in + (size_t)i * stride + (size_t)b * ORC_BLOCK;
      int16_t *p_dac = out + (size_t)i * stride + (size_t)b * ORC_BLOCK;
      orc_mix_fs4(p_adc, I, Q, ORC_BLOCK);
      orc_fir_run(k->fir_i, I, If, ORC_BLOCK);
      orc_fir_run(k->fir_q, Q, Qf, ORC_BLOCK);
      if (k->mode == ORC_SYNCAM && !c->am_q31) orc_syncam_block3(k->pll, If, Qf, p_dac, ORC_BLOCK);
      else orc_demod(kind, If, Qf, p_dac, ORC_BLOCK);
      if (k->anr_on > 0 && k->anr) orc_anr_block_v(k->anr, k->anr_on, p_dac, ORC_BLOCK);
      orc_biquad_update(&k->bq[0], p_dac, ORC_BLOCK);
      orc_biquad_update(&k->bq[1], p_dac, ORC_BLOCK);
    }
  }
  return used;
}

/* ------------------------------------------------------------------------------------------------------------------
 * Front-end conditioning: ADC DC-blocking high-pass, AudioAmplifier gain, block-maximum AGC.
 * Signal order in the sketch: adc1 -> amp_adc -> queue_adc (Minimal-SDR.ino:76-78); AGC(p_adc) is called on every
 * block read from the queue, before the mix (.ino:530-534), and moves amp_adc.gain() for the blocks that follow.
 * Batch semantics here: zero queue latency, i.e. block k+1 is amplified with the gain AGC set after block k.
 * ------------------------------------------------------------------------------------------------------------------ */

#define ORC_COEF_HPF_DCBLOCK (1048300 << 10) /* input_adc.cpp:32, S1.30 */
#define ORC_AGCBUF 25                        /* .ino:445 */

/* FRACMUL_SHL(x, y, 1) (dspinst.h:358-368): SMULL, then hi << 2 | lo >> 30 */
static inline int32_t orc_fracmul_shl1(int32_t x, int32_t y)
{
  const int64_t p = (int64_t)x * y;
  return (int32_t)(((uint32_t)(int32_t)(p >> 32) << 2) | ((uint32_t)p >> 30));
}

/* input_adc.cpp:198-211.  in: raw unsigned ADC codes; state x1/y1 (input_adc.cpp:37-38; begin() presets x1 to the first
 * reading << 14 and y1 to 0, :59-63 - the caller does that). */
void orc_adc_hpf(const uint16_t *in, int16_t *out, uint32_t n, int32_t *hpf_x1, int32_t *hpf_y1)
{
  int32_t x1 = *hpf_x1, y1 = *hpf_y1;
  for (uint32_t i = 0; i < n; i++) {
    const int32_t tmp = (int32_t)((uint32_t)in[i] << 14);
    int32_t acc = (int32_t)((uint32_t)y1 - (uint32_t)x1);
    acc = (int32_t)((uint32_t)acc + (uint32_t)tmp);
    y1 = orc_fracmul_shl1(acc, ORC_COEF_HPF_DCBLOCK);
    x1 = tmp;
    out[i] = (int16_t)orc_ssat16(y1 >> 14); /* signed_saturate_rshift(y1, 16, 14) */
  }
  *hpf_x1 = x1; *hpf_y1 = y1;
}

/* AudioAmplifier::gain (mixer.h:75-79) */
int32_t orc_amp_multiplier(float n)
{
  if (n > 32767.0f) n = 32767.0f;
  else if (n < -32767.0f) n = -32767.0f;
  return (int32_t)(n * 65536.0f);
}

/* AudioAmplifier::update (mixer.cpp:134-159) + applyGain (:34-47).  mult == 0: the reference transmits NO block at all;
 * a batch has to put something there - zeros, returned 0 so the caller can tell. */
int orc_amp_apply(int16_t *data, uint32_t n, int32_t mult)
{
  if (mult == 0) { memset(data, 0, n * sizeof(int16_t)); return 0; }
  if (mult == 65536) return 1;
  for (uint32_t i = 0; i < n; i++)
    data[i] = (int16_t)orc_ssat16((int32_t)(((int64_t)mult * data[i]) >> 16)); /* SMULWB/T, SSAT #16 */
  return 1;
}

/* ARMv7E-M SIMD helpers used by AGC(): SSUB16 sets APSR.GE[1:0] / [3:2] when the low / high halfword difference is >= 0,
 * SEL picks bytes of the first operand where GE is set. */
static inline uint32_t orc_ssub16(uint32_t a, uint32_t b, uint32_t *ge)
{
  const int32_t lo = (int16_t)(a & 0xFFFF) - (int16_t)(b & 0xFFFF), hi = (int16_t)(a >> 16) - (int16_t)(b >> 16);
  *ge = (lo >= 0 ? 3u : 0u) | (hi >= 0 ? 12u : 0u);
  return ((uint32_t)hi << 16) | ((uint32_t)lo & 0xFFFFu);
}
static inline uint32_t orc_sel(uint32_t a, uint32_t b, uint32_t ge)
{
  const uint32_t m = ((ge & 1u) ? 0x000000FFu : 0u) | ((ge & 2u) ? 0x0000FF00u : 0u) | ((ge & 4u) ? 0x00FF0000u : 0u) | ((ge & 8u) ? 0xFF000000u : 0u);
  return (a & m) | (b & ~m);
}

typedef struct {
  int16_t buf[ORC_AGCBUF];
  int idx;     /* .ino:451: starts at AGCBUF_SIZE */
  float val;   /* AGC_val, .ino:104 */
  float max;   /* AGC_Max, .ino:95 */
  int on;      /* AGC_on, .ino:100 */
} orc_agc;

void orc_agc_init(orc_agc *a, float start, float max, int on)
{
  memset(a, 0, sizeof(*a));
  a->idx = ORC_AGCBUF; a->val = start; a->max = max; a->on = on;
}

/* .ino:453-479: block maximum of |x| with the halfword SIMD idiom, literally (note abs() of whole 32-bit words whose upper
 * halves hold left-overs of the parallel compare). */
uint16_t orc_agc_absmax(const int16_t *block)
{
  const uint32_t *p = (const uint32_t *)block;
  int minv = 32767, maxv = -minv;
  uint32_t ge;
  for (int i = 0; i < ORC_BLOCK / 2; i++) {
    const uint32_t data = p[i];
    (void)orc_ssub16((uint32_t)maxv, data, &ge); maxv = (int)orc_sel((uint32_t)maxv, data, ge);
    (void)orc_ssub16(data, (uint32_t)minv, &ge); minv = (int)orc_sel((uint32_t)minv, data, ge);
  }
  (void)orc_ssub16((uint32_t)maxv, (uint32_t)(maxv >> 16), &ge); maxv = (int)orc_sel((uint32_t)maxv, (uint32_t)(maxv >> 16), ge);
  (void)orc_ssub16((uint32_t)(minv >> 16), (uint32_t)minv, &ge); minv = (int)orc_sel((uint32_t)minv, (uint32_t)(minv >> 16), ge);
  minv = (int)(minv < 0 ? 0u - (uint32_t)minv : (uint32_t)minv); /* abs(), INT_MIN stays INT_MIN like the ARM code */
  maxv = (int)(maxv < 0 ? 0u - (uint32_t)maxv : (uint32_t)maxv);
  (void)orc_ssub16((uint32_t)maxv, (uint32_t)minv, &ge);
  return (uint16_t)orc_sel((uint32_t)maxv, (uint32_t)minv, ge);
}

/* .ino:481-514.  Returns 1 and the new amplifier multiplier in *mult when amp_adc.gain() was called.
 * Reference defect kept visible: `agc_buffer[--agc_idx] = absmax; if (agc_idx < 0) agc_idx = AGCBUF_SIZE;` (.ino:481-482)
 * stores every 26th value at index -1, outside the array (undefined behaviour).  Here, and in the compiled reference
 * (oracle/Makefile gives the array a guard element in front), that store lands nowhere: the value is dropped. */
int orc_agc_update(orc_agc *a, uint16_t absmax, int32_t *mult)
{
  if (!a->on) return 0;
  --a->idx;
  if (a->idx >= 0) a->buf[a->idx] = (int16_t)absmax;
  if (a->idx < 0) a->idx = ORC_AGCBUF;
  int m = 0;
  for (int i = 0; i < ORC_AGCBUF; i++) m += a->buf[i];
  const int d = m / ORC_AGCBUF;
  const float x = 16000;
  const float f = x / d;
  int changed = 0;
  if (f > 1.3) {
    const float fagc = a->val + (a->val * f / 1500);
    if (fagc < a->max) { a->val = fagc; changed = 1; }
  } else if (a->val > 0.1) {
    if (f < 0.6) { a->val = a->val - (a->val * f / 50); changed = 1; }
    else if (f < 0.7) { a->val = a->val - (a->val * f / 200); changed = 1; }
    else if (f < 0.8) { a->val = a->val - (a->val * f / 2000); changed = 1; }
    else if (f < 0.9) { a->val = a->val - (a->val * f / 4000); changed = 1; }
  }
  if (changed) *mult = orc_amp_multiplier(a->val);
  return changed;
}

typedef struct {
  int32_t hpf_x1, hpf_y1;
  int32_t mult;
  orc_agc agc;
} orc_frontend;

orc_frontend *orc_frontend_new(uint32_t n_channels, float agc_start, float agc_max, int agc_on)
{
  orc_frontend *f = (orc_frontend *)calloc(n_channels ? n_channels : 1, sizeof(orc_frontend));
  for (uint32_t i = 0; f && i < n_channels; i++) {
    orc_agc_init(&f[i].agc, agc_start, agc_max, agc_on);
    f[i].mult = orc_amp_multiplier(agc_start); /* .ino:385 */
  }
  return f;
}
void orc_frontend_free(orc_frontend *f) { free(f); }
void orc_frontend_preset(orc_frontend *f, uint32_t ch, uint16_t first_reading) /* AudioInputAnalog::init, input_adc.cpp:59-63 */
{
  f[ch].hpf_x1 = (int32_t)((uint32_t)first_reading << 14);
  f[ch].hpf_y1 = 0;
}
void orc_frontend_get(const orc_frontend *f, uint32_t ch, int32_t *x1, int32_t *y1, int32_t *mult, float *agc_val, int *agc_idx, int16_t *agc_buf)
{
  *x1 = f[ch].hpf_x1; *y1 = f[ch].hpf_y1; *mult = f[ch].mult; *agc_val = f[ch].agc.val; *agc_idx = f[ch].agc.idx;
  memcpy(agc_buf, f[ch].agc.buf, sizeof(f[ch].agc.buf));
}
/* raw ADC codes [n_channels][stride] -> conditioned int16 IF samples, block by block: HPF, amplifier, AGC */
void orc_frontend_run(orc_frontend *f, uint32_t n_channels, const uint16_t *adc, int16_t *out, uint32_t n_blocks, size_t stride)
{
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
  for (long c = 0; c < (long)n_channels; c++) {
    orc_frontend *k = &f[c];
    for (uint32_t b = 0; b < n_blocks; b++) {
      int16_t *o = out + (size_t)c * stride + (size_t)b * ORC_BLOCK;
      orc_adc_hpf(adc + (size_t)c * stride + (size_t)b * ORC_BLOCK, o, ORC_BLOCK, &k->hpf_x1, &k->hpf_y1);
      orc_amp_apply(o, ORC_BLOCK, k->mult);
      int32_t m;
      if (orc_agc_update(&k->agc, orc_agc_absmax(o), &m)) k->mult = m;
    }
  }
}

/* ------------------------------------------------------------------------------------------------------------------
 * LMS automatic notch / noise reduction: Minimal-SDR.ino:702-770, between the demodulator and the DAC
 * queue (the biquads follow as audio objects).  Variable-leak LMS after Warren Pratt's wdsp.  float32_t state, with the
 * `double` sub-expressions C's usual arithmetic conversions give the literals 1.0 and 1e-10.  Every operation below is a
 * separately rounded IEEE operation in the order the source evaluates it (x86-64 gcc without -ffast-math does the same).
 * ------------------------------------------------------------------------------------------------------------------ */
#define ORC_ANR_DLINE 512 /* .ino:707 */
#define ORC_ANR_TAPS 64   /* .ino:708 */
#define ORC_ANR_DELAY 16  /* .ino:709 */

typedef struct {
  float d[ORC_ANR_DLINE]; /* .ino:724 */
  float w[ORC_ANR_DLINE]; /* .ino:725 (only the first 64 are used) */
  float lidx;             /* .ino:715: 120 */
  float ngamma;           /* .ino:718: 0.001 */
  int in_idx;             /* .ino:723 */
} orc_anr;

void orc_anr_init(orc_anr *a)
{
  memset(a, 0, sizeof(*a));
  a->lidx = 120.0f;
  a->ngamma = 0.001f;
}

/* mode 1: notch filter (output = error), mode 2: noise reduction (output = y)  (.ino:749-750) */
void orc_anr_block(orc_anr *a, int mode, int16_t *p_dac, uint32_t n)
{
  const float two_mu = 0.001f, gamma = 0.1f, lidx_min = 0.0f, lidx_max = 200.0f, den_mult = 6.25e-10f, lincr = 1.0f, ldecr = 3.0f;
  const int mask = ORC_ANR_DLINE - 1;
  for (uint32_t i = 0; i < n; i++) {
    a->d[a->in_idx] = p_dac[i];
    float y = 0, sigma = 0;
    for (int j = 0; j < ORC_ANR_TAPS; j++) {
      const int idx = (a->in_idx + j + ORC_ANR_DELAY) & mask;
      y += a->w[j] * a->d[idx];
      sigma += a->d[idx] * a->d[idx];
    }
    const float inv_sigp = 1.0 / (sigma + 1e-10);
    const float error = a->d[a->in_idx] - y;
    if (mode == 1) p_dac[i] = error; else p_dac[i] = y;
    float nel, nev;
    if ((nel = error * (1.0 - two_mu * sigma * inv_sigp)) < 0.0) nel = -nel;
    if ((nev = a->d[a->in_idx] - (1.0 - two_mu * a->ngamma) * y - two_mu * error * sigma * inv_sigp) < 0.0) nev = -nev;
    if (nev < nel) {
      if ((a->lidx += lincr) > lidx_max) a->lidx = lidx_max;
      else if ((a->lidx -= ldecr) < lidx_min) a->lidx = lidx_min;
    }
    a->ngamma = gamma * (a->lidx * a->lidx) * (a->lidx * a->lidx) * den_mult;
    const float c0 = 1.0 - two_mu * a->ngamma;
    const float c1 = two_mu * error * inv_sigp;
    for (int j = 0; j < ORC_ANR_TAPS; j++) {
      const int idx = (a->in_idx + j + ORC_ANR_DELAY) & mask;
      a->w[j] = c0 * a->w[j] + c1 * a->d[idx];
    }
    a->in_idx = (a->in_idx + mask) & mask;
  }
}

void *orc_anr_alloc1(void)
{
  orc_anr *a = (orc_anr *)malloc(sizeof(orc_anr));
  if (a) orc_anr_init(a);
  return a;
}
void orc_anr_block_v(void *state, int mode, int16_t *p_dac, uint32_t n) { orc_anr_block((orc_anr *)state, mode, p_dac, n); }

orc_anr *orc_anr_new(uint32_t n_channels)
{
  orc_anr *a = (orc_anr *)malloc((n_channels ? n_channels : 1) * sizeof(orc_anr));
  for (uint32_t i = 0; a && i < n_channels; i++) orc_anr_init(&a[i]);
  return a;
}
void orc_anr_free(orc_anr *a) { free(a); }
void orc_anr_run(orc_anr *a, uint32_t n_channels, int mode, int16_t *data, uint32_t n_blocks, size_t stride)
{
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
  for (long c = 0; c < (long)n_channels; c++)
    for (uint32_t b = 0; b < n_blocks; b++) orc_anr_block(&a[c], mode, data + (size_t)c * stride + (size_t)b * ORC_BLOCK, ORC_BLOCK);
}
void orc_anr_get(const orc_anr *a, uint32_t ch, float *lidx, float *ngamma, int *in_idx, float *w64, float *d512)
{
  *lidx = a[ch].lidx; *ngamma = a[ch].ngamma; *in_idx = a[ch].in_idx;
  memcpy(w64, a[ch].w, ORC_ANR_TAPS * sizeof(float));
  memcpy(d512, a[ch].d, ORC_ANR_DLINE * sizeof(float));
}

/* ------------------------------------------------------------------------------------------------------------------
 * Synchronous AM demodulator with PLL: `case SYNCAM` of the demodulation switch, Minimal-SDR.ino:631-688
 * (Teensy 3.5/3.6 branch; after wdsp).  Input: the FIR-filtered I and Q blocks; output: corr[0] narrowed to int16.
 * float32 with libm sinf/cosf/atan2f, `double` where the literals 2.0 * PI force it.  Restated operation by operation; the
 * constants are evaluated with the same expressions as the sketch's static initialisers (SAMPLE_RATE = 24000, .ino:84-85).
 * ------------------------------------------------------------------------------------------------------------------ */
#define ORC_PI 3.1415926535897932384626433832795 /* Arduino.h PI */
#define ORC_SAMPLE_RATE (6000 * 4)

typedef struct {
  float fil_out, omega2, phzerror; /* .ino:643-645 */
} orc_syncam;

void orc_syncam_constants(float *omega_min, float *omega_max, float *g1, float *g2)
{
  const float omegaN = 400.0, zeta = 0.45;
  *omega_min = 2.0 * ORC_PI * -4000.0 / ORC_SAMPLE_RATE;
  *omega_max = 2.0 * ORC_PI * 4000.0 / ORC_SAMPLE_RATE;
  const float g1v = 1.0 - exp(-2.0 * omegaN * zeta / ORC_SAMPLE_RATE);
  *g1 = g1v;
  /* the sketch is C++: exp() of a float argument is the float overload, and `1 - float * float` stays float */
  *g2 = -g1v + 2.0 * (1 - expf(-omegaN * zeta / ORC_SAMPLE_RATE) * cosf(omegaN / ORC_SAMPLE_RATE * sqrtf(1.0 - zeta * zeta)));
}

void orc_syncam_block(orc_syncam *s, const int16_t *I_buffer, const int16_t *Q_buffer, int16_t *p_dac, uint32_t n)
{
  float omega_min, omega_max, g1, g2;
  orc_syncam_constants(&omega_min, &omega_max, &g1, &g2);
  for (uint32_t i = 0; i < n; i++) {
    const float Sin = sinf(s->phzerror), Cos = cosf(s->phzerror);
    const float ai = Cos * I_buffer[i], bi = Sin * I_buffer[i], aq = Cos * Q_buffer[i], bq = Sin * Q_buffer[i];
    float corr[2];
    corr[0] = +ai + bq;
    corr[1] = -bi + aq;
    p_dac[i] = corr[0];
    const float det = atan2f(corr[1], corr[0]);
    const float del_out = s->fil_out;
    s->omega2 = s->omega2 + g2 * det;
    if (s->omega2 < omega_min) s->omega2 = omega_min;
    else if (s->omega2 > omega_max) s->omega2 = omega_max;
    s->fil_out = g1 * det + s->omega2;
    s->phzerror = s->phzerror + del_out;
    while (s->phzerror >= 2 * ORC_PI) s->phzerror -= 2.0 * ORC_PI;
    while (s->phzerror < 0.0) s->phzerror += 2.0 * ORC_PI;
  }
}

void orc_syncam_block3(float *state3, const int16_t *I_buffer, const int16_t *Q_buffer, int16_t *p_dac, uint32_t n)
{
  orc_syncam s = {state3[0], state3[1], state3[2]};
  orc_syncam_block(&s, I_buffer, Q_buffer, p_dac, n);
  state3[0] = s.fil_out; state3[1] = s.omega2; state3[2] = s.phzerror;
}

orc_syncam *orc_syncam_new(uint32_t n_channels) { return (orc_syncam *)calloc(n_channels ? n_channels : 1, sizeof(orc_syncam)); }
void orc_syncam_free(orc_syncam *s) { free(s); }
void orc_syncam_run(orc_syncam *s, uint32_t n_channels, const int16_t *I, const int16_t *Q, int16_t *out, uint32_t n_blocks, size_t stride)
{
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
  for (long c = 0; c < (long)n_channels; c++)
    for (uint32_t b = 0; b < n_blocks; b++) {
      const size_t off = (size_t)c * stride + (size_t)b * ORC_BLOCK;
      orc_syncam_block(&s[c], I + off, Q + off, out + off, ORC_BLOCK);
    }
}
void orc_syncam_get(const orc_syncam *s, uint32_t ch, float *fil_out, float *omega2, float *phzerror)
{
  *fil_out = s[ch].fil_out; *omega2 = s[ch].omega2; *phzerror = s[ch].phzerror;
}
