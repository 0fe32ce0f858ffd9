// Internal: NCCL bound at run time. libanemoi_b200.so does not link libnccl: the first collective call dlopen()s
// libnccl.so.2, preferring a copy that is ALREADY mapped into the process (RTLD_NOLOAD) -- under PyTorch that is the
// NCCL torch itself uses, so both see one library; a Rust / C host gets the system one. Only the handful of entry
// points the sharded Merkle builder needs are bound (stable since NCCL 2.0); the declarations below restate the
// public nccl.h prototypes so that building the library needs no NCCL headers.
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstddef>
#include <mutex>

namespace anemoi {
namespace nccl {

typedef struct ncclComm* comm_t;
typedef struct { char internal[128]; } unique_id;  // NCCL_UNIQUE_ID_BYTES
enum { kSuccess = 0 };
enum { kUint8 = 1 };  // ncclUint8

struct Api {
    void* handle = nullptr;
    int (*GetUniqueId)(unique_id*) = nullptr;
    int (*CommInitRank)(comm_t*, int, unique_id, int) = nullptr;
    int (*CommInitAll)(comm_t*, int, const int*) = nullptr;
    int (*CommDestroy)(comm_t) = nullptr;
    int (*CommCount)(const comm_t, int*) = nullptr;
    int (*CommUserRank)(const comm_t, int*) = nullptr;
    int (*CommCuDevice)(const comm_t, int*) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, comm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    int (*GetVersion)(int*) = nullptr;
    bool ok = false;
};

inline const Api& api() {
    static Api a;
    static std::once_flag once;
    std::call_once(once, []() {
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* nm : names) {
            a.handle = dlopen(nm, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
            if (a.handle) break;
        }
        for (int i = 0; !a.handle && i < 2; i++) a.handle = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
        if (!a.handle) return;
#define ANEMOI_NCCL_BIND(field, sym) *(void**)(&a.field) = dlsym(a.handle, sym)
        ANEMOI_NCCL_BIND(GetUniqueId, "ncclGetUniqueId");
        ANEMOI_NCCL_BIND(CommInitRank, "ncclCommInitRank");
        ANEMOI_NCCL_BIND(CommInitAll, "ncclCommInitAll");
        ANEMOI_NCCL_BIND(CommDestroy, "ncclCommDestroy");
        ANEMOI_NCCL_BIND(CommCount, "ncclCommCount");
        ANEMOI_NCCL_BIND(CommUserRank, "ncclCommUserRank");
        ANEMOI_NCCL_BIND(CommCuDevice, "ncclCommCuDevice");
        ANEMOI_NCCL_BIND(AllGather, "ncclAllGather");
        ANEMOI_NCCL_BIND(GroupStart, "ncclGroupStart");
        ANEMOI_NCCL_BIND(GroupEnd, "ncclGroupEnd");
        ANEMOI_NCCL_BIND(GetErrorString, "ncclGetErrorString");
        ANEMOI_NCCL_BIND(GetVersion, "ncclGetVersion");
#undef ANEMOI_NCCL_BIND
        a.ok = a.GetUniqueId && a.CommInitRank && a.CommInitAll && a.CommDestroy && a.CommCount && a.CommUserRank &&
               a.AllGather && a.GroupStart && a.GroupEnd && a.GetErrorString;
    });
    return a;
}

}  // namespace nccl
}  // namespace anemoi
