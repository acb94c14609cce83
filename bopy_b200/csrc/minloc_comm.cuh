// The one exchange step of the sharded sweep: a min-loc reduction of one 16-byte
// (value, global index) record per rank, on the sweep's own stream.
//
// NCCL has no MINLOC operator, so the records are all-gathered (ncclAllGather of 16 bytes per rank over NVLink) and a
// one-warp kernel reduces them with np.argmin's ordering (minloc_better: NaN first unless skipped, then value, then the
// lowest index) -- every rank ends up with the same winner in its own (val, idx) buffers and no host round trip.
//
// NCCL is bound at run time (dlopen of the libnccl.so.2 the process already has -- the one PyTorch loaded -- or the
// system one), so the library keeps loading and serving single-GPU callers where NCCL is absent.
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdint.h>

#include <mutex>
#include <string>

#include "common.cuh"

namespace bopy {

// the five NCCL entry points used, declared here so that no NCCL header is needed at build time (stable C ABI)
struct NcclApi {
    typedef struct ncclComm* comm_t;
    struct unique_id {
        char internal[128];
    };
    int (*GetUniqueId)(unique_id*) = nullptr;
    int (*CommInitRank)(comm_t*, int, unique_id, int) = nullptr;
    int (*CommDestroy)(comm_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int /*ncclDataType_t*/, comm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    int (*GetVersion)(int*) = nullptr;
    void* handle = nullptr;
    std::string error;
    bool ok = false;
};

inline NcclApi& nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* nm : names) {   // prefer the copy already mapped into the process (PyTorch's)
            api.handle = dlopen(nm, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
            if (api.handle) break;
        }
        if (!api.handle)
            for (const char* nm : names) {
                api.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
                if (api.handle) break;
            }
        if (!api.handle) {
            api.error = std::string("libnccl.so.2 not found: ") + (dlerror() ? dlerror() : "");
            return;
        }
        auto sym = [&](const char* s) { return dlsym(api.handle, s); };
        api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
        api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
        api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
        api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
        api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
        api.GetVersion = reinterpret_cast<decltype(api.GetVersion)>(sym("ncclGetVersion"));
        api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllGather && api.GetErrorString;
        if (!api.ok) api.error = "libnccl.so.2 lacks one of ncclGetUniqueId / ncclCommInitRank / ncclCommDestroy / ncclAllGather";
    });
    return api;
}

constexpr int NCCL_INT8 = 0;   // ncclInt8 / ncclChar

// (val, idx) -> this rank's record
__global__ void minloc_pack_kernel(const double* __restrict__ val, const long long* __restrict__ idx, MinLoc* rec) {
    rec->val = *val;
    rec->idx = *idx;
}

// one warp: reduce `world` gathered records; nan_skip: records whose value is NaN lose against every number
__global__ void minloc_gathered_kernel(const MinLoc* __restrict__ recs, int world, int nan_skip, double* val_out, long long* idx_out) {
    MinLoc v;
    v.val = 0.0;
    v.idx = -1;
    for (int r = threadIdx.x; r < world; r += 32) {
        MinLoc c = recs[r];
        if (nan_skip && c.val != c.val) c.idx = -1;
        if (minloc_better(c, v)) v = c;
    }
    v = minloc_warp_reduce(v);
    if (threadIdx.x == 0) {
        *val_out = v.val;
        *idx_out = v.idx;
    }
}

}  // namespace bopy
