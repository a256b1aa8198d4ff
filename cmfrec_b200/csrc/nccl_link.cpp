#include "nccl_link.h"
#include <dlfcn.h>
#include <cstdio>
#include <cstring>
#include <vector>

namespace cmfb200 {

namespace {
// the handful of NCCL entry points used, declared locally (ABI per nccl.h 2.27/2.28)
typedef struct { char internal[128]; } id_t;
typedef int (*GetUniqueId_t)(id_t *);
typedef int (*CommInitRank_t)(void **, int, id_t, int);
typedef int (*CommDestroy_t)(void *);
typedef int (*AllGather_t)(const void *, void *, size_t, int, void *, cudaStream_t);
typedef int (*AllReduce_t)(const void *, void *, size_t, int, int, void *, cudaStream_t);
typedef const char *(*GetErrorString_t)(int);
typedef int (*Group_t)(void);
enum { kNcclInt8 = 0, kNcclFloat32 = 7, kNcclFloat64 = 8, kNcclSum = 0 };

struct Api {
    void *handle = nullptr;
    GetUniqueId_t GetUniqueId = nullptr;
    CommInitRank_t CommInitRank = nullptr;
    CommDestroy_t CommDestroy = nullptr;
    AllGather_t AllGather = nullptr;
    AllReduce_t AllReduce = nullptr;
    GetErrorString_t GetErrorString = nullptr;
    Group_t GroupStart = nullptr, GroupEnd = nullptr;
    bool ok = false;
};

Api &api()
{
    static Api a;
    if (a.handle) return a;
    a.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!a.handle) {
        std::fprintf(stderr, "cmfrec_b200: cannot load libnccl.so.2: %s\n", dlerror());
        return a;
    }
    a.GetUniqueId = (GetUniqueId_t)dlsym(a.handle, "ncclGetUniqueId");
    a.CommInitRank = (CommInitRank_t)dlsym(a.handle, "ncclCommInitRank");
    a.CommDestroy = (CommDestroy_t)dlsym(a.handle, "ncclCommDestroy");
    a.AllGather = (AllGather_t)dlsym(a.handle, "ncclAllGather");
    a.AllReduce = (AllReduce_t)dlsym(a.handle, "ncclAllReduce");
    a.GetErrorString = (GetErrorString_t)dlsym(a.handle, "ncclGetErrorString");
    a.GroupStart = (Group_t)dlsym(a.handle, "ncclGroupStart");
    a.GroupEnd = (Group_t)dlsym(a.handle, "ncclGroupEnd");
    a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllGather && a.AllReduce;
    return a;
}

int check(int rc, const char *what)
{
    if (rc != 0) {
        Api &a = api();
        std::fprintf(stderr, "cmfrec_b200: %s failed: %s\n", what, a.GetErrorString ? a.GetErrorString(rc) : "?");
        return 1;
    }
    return 0;
}
}  // namespace

// Communicators are cached per process under the unique id they were created from: a second state created with the
// same id reuses the communicator instead of paying ncclCommInitRank again (seconds), which is what a caller that fits
// many models in one process wants.  They live until the process exits.
namespace {
struct CachedComm { id_t id; int rank, world; void *comm; };
std::vector<CachedComm> &comm_cache()
{
    static std::vector<CachedComm> c;
    return c;
}
}  // namespace

NcclLink::~NcclLink() {}

int NcclLink::unique_id(void *out128)
{
    Api &a = api();
    if (!a.ok) return 1;
    id_t id;
    if (check(a.GetUniqueId(&id), "ncclGetUniqueId")) return 1;
    std::memcpy(out128, &id, sizeof(id));
    return 0;
}

int NcclLink::init(const void *id128, int rank_, int world_)
{
    Api &a = api();
    if (!a.ok) return 1;
    rank = rank_;
    world = world_;
    id_t id;
    std::memcpy(&id, id128, sizeof(id));
    for (const CachedComm &c : comm_cache())
        if (c.rank == rank && c.world == world && std::memcmp(&c.id, &id, sizeof(id)) == 0) {
            comm = c.comm;
            return 0;
        }
    if (check(a.CommInitRank(&comm, world, id, rank), "ncclCommInitRank")) return 1;
    comm_cache().push_back(CachedComm{id, rank, world, comm});
    return 0;
}

int NcclLink::all_gather_inplace(void *base, size_t bytes_per_rank, cudaStream_t stream)
{
    Api &a = api();
    const char *send = (const char *)base + (size_t)rank * bytes_per_rank;
    return check(a.AllGather(send, base, bytes_per_rank, kNcclInt8, comm, stream), "ncclAllGather");
}

// calls between group_begin() and group_end() are fused into one NCCL launch (no-ops when the symbols are missing)
int NcclLink::group_begin()
{
    Api &a = api();
    return a.GroupStart ? check(a.GroupStart(), "ncclGroupStart") : 0;
}
int NcclLink::group_end()
{
    Api &a = api();
    return a.GroupEnd ? check(a.GroupEnd(), "ncclGroupEnd") : 0;
}

int NcclLink::all_reduce_sum(void *buf, size_t count, bool is_double, cudaStream_t stream)
{
    Api &a = api();
    return check(a.AllReduce(buf, buf, count, is_double ? kNcclFloat64 : kNcclFloat32, kNcclSum, comm, stream),
                 "ncclAllReduce");
}

}  // namespace cmfb200
