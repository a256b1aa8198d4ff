// Micro-benchmark: FP64 tensor-core (DMMA) issue rate of the mma.sync shapes on sm_100a, next to plain DFMA.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_peak dmma_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int SHAPE, int NACC> __global__ void k(double *out, int iters)
{
    double a[4] = {1.0 + threadIdx.x * 1e-9, 1.0, 1.0, 1.0}, b[2] = {1.0 - threadIdx.x * 1e-9, 1.0};
    double c[NACC][4];
    for (int i = 0; i < NACC; i++) for (int j = 0; j < 4; j++) c[i][j] = 0.0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) {
            if (SHAPE == 0) {
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a[0]), "d"(b[0]));
            } else if (SHAPE == 1) {
                asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                             : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3]) : "d"(a[0]), "d"(a[1]), "d"(b[0]));
            } else if (SHAPE == 2) {
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                             : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
            } else {
#pragma unroll
                for (int j = 0; j < 4; j++) c[i][j] = fma(a[j], b[j & 1], c[i][j]);
            }
        }
    }
    double s = 0;
    for (int i = 0; i < NACC; i++) for (int j = 0; j < 4; j++) s += c[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int SHAPE, int NACC> void run(const char *name, double flop_per_inst, int threads)
{
    double *out;
    cudaMalloc(&out, 148 * 8 * 1024 * sizeof(double));
    const int iters = 20000, blocks = 148 * 2;
    k<SHAPE, NACC><<<blocks, threads>>>(out, 10);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<SHAPE, NACC><<<blocks, threads>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double insts = (double)blocks * (threads / 32) * iters * NACC;
    printf("%-12s threads=%4d nacc=%d  %.3f ms  %.2f TFLOP/s  (%.1f clk per warp-inst per SMSP at 1.965 GHz)\n", name, threads, NACC, ms,
           insts * flop_per_inst / ms / 1e9, ms * 1e-3 * 1.965e9 / (insts / (148.0 * 4)));
    cudaFree(out);
}

int main()
{
    run<0, 8>("m8n8k4", 512, 256);
    run<0, 8>("m8n8k4", 512, 512);
    run<0, 16>("m8n8k4", 512, 256);
    run<1, 8>("m16n8k4", 1024, 256);
    run<2, 8>("m16n8k8", 2048, 256);
    run<2, 8>("m16n8k8", 2048, 512);
    run<3, 8>("dfma x4", 4 * 32 * 2, 256);
    run<3, 8>("dfma x4", 4 * 32 * 2, 512);
    return 0;
}
