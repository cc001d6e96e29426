/* TEST INFRASTRUCTURE — minimal Arduino.h stand-in so the reference's Teensy Audio sources
 * (src/Audio/filter_biquad.cpp, mixer.cpp) compile on the host.  Selects the Cortex-M4
 * (KINETISK) code paths, which is what the Teensy 3.x targets of the sketch use. */
#ifndef MSDR_STUB_ARDUINO_H
#define MSDR_STUB_ARDUINO_H
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <math.h>
#define KINETISK 1
static inline void __disable_irq(void) {}
static inline void __enable_irq(void) {}
#endif
