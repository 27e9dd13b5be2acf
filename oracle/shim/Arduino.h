/* Host shim for the reference's `#include <Arduino.h>` (test infrastructure only).
 *
 * The reference (schreibfaul1/ESP32-Opus-Player, src/*.cpp) is firmware: it expects the Arduino core
 * for min/max, PROGMEM, ps_malloc and the log_* macros.  This header supplies just enough of that
 * surface for g++ on x86-64 to compile the five codec translation units *unmodified*
 * (SURVEY.md section 8c).  Nothing in the shipped library includes this file.
 */
#pragma once
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <stdarg.h>
#include <limits.h>
#include <assert.h>
#include <math.h>
#include <algorithm>
#include <memory>

using std::max;
using std::min;

#ifndef _min
#define _min(a, b) ((a) < (b) ? (a) : (b))
#endif
#ifndef _max
#define _max(a, b) ((a) > (b) ? (a) : (b))
#endif

#define PROGMEM
typedef bool boolean;

/* PSRAM allocator -> zeroed host heap (the decoded PCM does not depend on the fill). */
static inline void* ps_malloc(size_t n) { return calloc(1, n ? n : 1); }
static inline void* ps_calloc(size_t a, size_t b) { return calloc(a ? a : 1, b ? b : 1); }

static inline void ref_shim_log_sink(const char*, ...) {}
#define log_v(...) ref_shim_log_sink(__VA_ARGS__)
#define log_d(...) ref_shim_log_sink(__VA_ARGS__)
#define log_i(...) ref_shim_log_sink(__VA_ARGS__)
#define log_w(...) ref_shim_log_sink(__VA_ARGS__)
#define log_e(...) ref_shim_log_sink(__VA_ARGS__)
