// Explicit-feedback model with dense side information and/or implicit features (see collective.cu).
#pragma once
#include <vector>
#include "als.h"

namespace cmfb200 {

struct CollectiveConfig {
    int p = 0, q = 0;                 // columns of U / I (0 = absent)
    bool implicit_features = false;
    real_t w_user = 1, w_item = 1, w_implicit = 1;     // already divided by w_main
    real_t lam_C = 0, lam_D = 0, lam_Bi = 0, lam_Ai = 0;  // already divided by their weight and scaled by the row count
};

class CollectiveState {
public:
    AlsState *st = nullptr;
    CollectiveConfig cc;
    DevBuf<real_t> Uc, Ic, C, D, Ai, Bi;
    DevBuf<real_t> QA, QB, G1, G2, T1, Ldev, ws, qA, qB;

    int setup(AlsState *state, const CollectiveConfig &c, const real_t *Uc_host, const real_t *Ic_host);
    int iteration(int it, int solver);
    int iterate(int niter, bool use_cg, bool finalize_chol);
    int download(real_t *hC, real_t *hD, real_t *hAi, real_t *hBi);

private:
    int update_side_factor(const real_t *F, int ldF, int_t rows, const real_t *S, int p, real_t lam, real_t *Cout);
    int update_implicit_factor(const DeviceSide &side, const real_t *F, int ldF, int_t rowsF, real_t lam, real_t *Out);
    int build_extras(int which, const DeviceSide &side, int_t rows, const real_t *S, int p, const real_t *Cfac, real_t w_side,
                     const real_t *Fi_opp, int_t rows_opp);
};

int center_side_info(const real_t *S, int_t rows, int p, real_t *colmeans, std::vector<real_t> &centred);

}  // namespace cmfb200
