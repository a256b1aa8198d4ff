// Routes the cached CG half-sweep (sweep_cg_resident.cu, compiled once per model) by model.
#include "sweep.h"

namespace cmfb200 {

int resident_sweep_explicit(const CgSweepParams &p, cudaStream_t stream, int *n_launches);
int resident_sweep_implicit(const CgSweepParams &p, cudaStream_t stream, int *n_launches);
int resident_sweep_collective(const CgSweepParams &p, cudaStream_t stream, int *n_launches);

// 0 = launched, 3 = this shape is not covered (nothing was launched: use the direct kernel), other = error
int launch_explicit_cg_sweep_resident(const CgSweepParams &p, cudaStream_t stream, int *n_launches)
{
    return (p.gram || p.qvec || p.solve_all_rows) ? resident_sweep_collective(p, stream, n_launches)
                                                  : resident_sweep_explicit(p, stream, n_launches);
}
int launch_implicit_cg_sweep_resident(const CgSweepParams &p, cudaStream_t stream, int *n_launches)
{
    return resident_sweep_implicit(p, stream, n_launches);
}

}  // namespace cmfb200
