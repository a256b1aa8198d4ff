// Small device-side helpers shared by the sweep kernels.
#pragma once
#include <cuda_runtime.h>
#include "cmf_types.h"

namespace cmfb200 {

#define CMF_FULL_MASK 0xffffffffu

template <typename T> struct VecOf;
template <> struct VecOf<float> { typedef float4 type; static constexpr int N = 4; };
template <> struct VecOf<double> { typedef double2 type; static constexpr int N = 2; };

// read-only (non-coherent) vector load of one 16-byte piece of an opposing-factor row
__device__ __forceinline__ void ldg_vec(const float *p, float *out)
{
    const float4 v = __ldg(reinterpret_cast<const float4 *>(p));
    out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w;
}
__device__ __forceinline__ void ldg_vec(const double *p, double *out)
{
    const double2 v = __ldg(reinterpret_cast<const double2 *>(p));
    out[0] = v.x; out[1] = v.y;
}

// Load C consecutive coordinates starting at column c0 of a factor row whose allocated width is ld
// (ld is a multiple of the 16-byte vector width, rows are 16-byte aligned); columns >= ld read as 0.
template <typename T, int C>
__device__ __forceinline__ void load_row_chunk(const T *row, int c0, int ld, bool valid, T (&v)[C])
{
    constexpr int VN = VecOf<T>::N;
    if constexpr (C % VN == 0) {
#pragma unroll
        for (int q = 0; q < C / VN; q++) {
            if (valid && c0 + q * VN < ld) {
                ldg_vec(row + c0 + q * VN, &v[q * VN]);
            } else {
#pragma unroll
                for (int e = 0; e < VN; e++) v[q * VN + e] = T(0);
            }
        }
    } else {
#pragma unroll
        for (int c = 0; c < C; c++) v[c] = (valid && c0 + c < ld) ? __ldg(row + c0 + c) : T(0);
    }
}

// sum over the L lanes of a sub-warp group (L a power of two, groups are aligned lane ranges)
template <int L, typename T> __device__ __forceinline__ T group_sum(T x)
{
#pragma unroll
    for (int off = L / 2; off >= 1; off >>= 1) x += __shfl_xor_sync(CMF_FULL_MASK, x, off);
    return x;
}

// sum over the 32/L groups of a warp: lanes with equal (lane % L) are combined
template <int L, typename T> __device__ __forceinline__ T across_groups_sum(T x)
{
#pragma unroll
    for (int off = 16; off >= L; off >>= 1) x += __shfl_xor_sync(CMF_FULL_MASK, x, off);
    return x;
}

}  // namespace cmfb200
