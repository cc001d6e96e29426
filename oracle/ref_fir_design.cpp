/* TEST INFRASTRUCTURE — compiles the reference's Kaiser FIR designer (calc_FIR_coeffs, m_sinc, Izero,
 * Minimal-SDR.ino:779-899) which oracle/Makefile extracts BY LINE RANGE at build time into
 * oracle/_ref/fir_design_extract.inc (git-ignored; no reference source is committed).
 * Used only to pin minimal-sdr_b200's own host-side designer and to generate the AM table fixture. */
#include <stdint.h>
#include <math.h>
typedef float float32_t;
#ifndef PI
#define PI 3.1415926535897932384626433832795 /* Teensy core wiring.h value; arm_math.h:365 only defines PI if absent */
#endif
float m_sinc(int m, float fc);
float32_t Izero(float32_t x);
#include "fir_design_extract.inc"
extern "C" void ref_calc_FIR_coeffs(int16_t *coeffs, int numCoeffs, float fc, float Astop, int type, float dfc, float Fsamprate)
{
  calc_FIR_coeffs(coeffs, numCoeffs, fc, Astop, type, dfc, Fsamprate);
}
