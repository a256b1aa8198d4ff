// Scalar types shared by every translation unit of libcmfrec_b200_{f32,f64}.so.
// Mirrors the reference's compile-time switch (reference src/cmfrec.h:232-305): one library per
// real_t, int_t = 32-bit int, CSR index pointers are size_t.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cfloat>

#ifdef USE_FLOAT
typedef float real_t;
#define CMF_EPS FLT_EPSILON
#define CMF_REAL_NAME "f32"
#else
typedef double real_t;
#define CMF_EPS DBL_EPSILON
#define CMF_REAL_NAME "f64"
#endif
typedef int int_t;

// Device row stride of a factor matrix with `kk` coordinates per row: rows are padded to a whole number of
// 128-byte cache lines and start line-aligned, so that a gathered row is a run of FULL lines (the gathers are
// bound by the L1/TEX tag rate: a row that straddles lines costs extra tag look-ups on every pass).
static inline int cmf_ld_for(int kk)
{
    const int q = (int)(128 / sizeof(real_t));
    return ((kk + q - 1) / q) * q;
}
