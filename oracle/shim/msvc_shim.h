// Force-included (-include) when compiling cpp/cordic_sincos.cpp, which uses the MSVC-only
// fopen_s / errno_t (cpp/cordic_sincos.cpp:131).  TEST INFRASTRUCTURE ONLY.
#ifndef BHW_ORACLE_MSVC_SHIM_H_
#define BHW_ORACLE_MSVC_SHIM_H_
#include <stdio.h>
typedef int errno_t;
static inline errno_t fopen_s(FILE** f, const char* name, const char* mode) {
  *f = fopen(name, mode[0] == 'w' ? "w" : mode);
  return *f ? 0 : 1;
}
#endif
