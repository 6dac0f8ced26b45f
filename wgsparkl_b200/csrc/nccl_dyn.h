// NCCL resolved at run time (dlopen), so that libb200mpm.so has no link-time dependency on it and shares the
// copy already loaded by the host process (torch's bundled libnccl.so.2 in the Python host, the system one in a
// Rust / C++ host). Only the handful of entry points the slab exchange needs.
#pragma once

#include <dlfcn.h>
#include <nccl.h>

#include <string>

namespace b2 {

struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
    std::string error;
};

inline const NcclApi& nccl_api() {
    static NcclApi api = [] {
        NcclApi a;
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) {
            a.error = std::string("cannot load libnccl.so.2: ") + dlerror();
            return a;
        }
#define B2_NCCL_SYM(field, name)                                   \
    a.field = (decltype(a.field))dlsym(h, name);                   \
    if (!a.field) {                                                \
        a.error = std::string("libnccl misses symbol ") + name;    \
        return a;                                                  \
    }
        B2_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
        B2_NCCL_SYM(CommInitRank, "ncclCommInitRank")
        B2_NCCL_SYM(CommDestroy, "ncclCommDestroy")
        B2_NCCL_SYM(GroupStart, "ncclGroupStart")
        B2_NCCL_SYM(GroupEnd, "ncclGroupEnd")
        B2_NCCL_SYM(Send, "ncclSend")
        B2_NCCL_SYM(Recv, "ncclRecv")
        B2_NCCL_SYM(AllReduce, "ncclAllReduce")
        B2_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef B2_NCCL_SYM
        a.ok = true;
        return a;
    }();
    return api;
}

} // namespace b2
