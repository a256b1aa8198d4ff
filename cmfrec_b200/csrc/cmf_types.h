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

// Device row stride of a factor matrix holding `kk` solved coordinates plus one bias slot:
// rows start 16-byte aligned so that a whole row is one vector/bulk copy.
static inline int cmf_ld_for(int kk_plus_slot)
{
    const int q = (int)(16 / sizeof(real_t));
    return ((kk_plus_slot + q - 1) / q) * q;
}
