#include "sweep.h"
namespace cmfb200 {
int launch_explicit_chol_sweep(const CgSweepParams &, cudaStream_t) { return 2; }
int launch_implicit_chol_sweep(const CgSweepParams &, cudaStream_t) { return 2; }
}
