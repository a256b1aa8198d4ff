#include "topn.h"
#include <cstdio>
namespace cmfb200 {
int top_n(real_t *, int_t, real_t *, int_t, real_t *, real_t, real_t, int_t, int_t, int_t *, int_t, int_t *, int_t, int_t *,
          real_t *, int_t, int_t, int)
{
    std::fprintf(stderr, "cmfrec_b200: topN: not implemented yet\n");
    return 2;
}
}
