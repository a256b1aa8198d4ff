#include "popular.h"
#include <cstdio>
namespace cmfb200 {
int most_popular(real_t *, real_t *, real_t *, real_t, real_t, bool, bool, real_t, int_t, int_t, int_t *, int_t *, real_t *,
                 size_t, real_t *, real_t *, bool, bool, bool, bool, bool, real_t *, int)
{
    std::fprintf(stderr, "cmfrec_b200: fit_most_popular: not implemented yet\n");
    return 2;
}
}
