// nccl_dl.h -- NCCL resolved at run time with dlopen/dlsym instead of a link-time dependency:
// the host process usually has PyTorch's bundled libnccl.so.2 loaded already (torch.distributed
// does the rendezvous), and a second, older system copy bound at link time would shadow its
// symbols.  dlopen("libnccl.so.2") returns the copy that is already in the process, if any.
#pragma once
#include <dlfcn.h>
#include <nccl.h>

namespace mdbg {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*CommSplit)(ncclComm_t, int, int, ncclComm_t*, void*) = nullptr;   // optional (NCCL >= 2.18)
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    bool ok = false;
};

inline NcclApi& nccl() {
    static NcclApi api;
    if (api.handle) return api;
    api.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
    if (!api.handle) api.handle = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
    if (!api.handle) return api;
#define MDBG_NCCL_SYM(field, name) api.field = (decltype(api.field))dlsym(api.handle, name)
    MDBG_NCCL_SYM(GetUniqueId, "ncclGetUniqueId");
    MDBG_NCCL_SYM(CommInitRank, "ncclCommInitRank");
    MDBG_NCCL_SYM(CommDestroy, "ncclCommDestroy");
    MDBG_NCCL_SYM(CommSplit, "ncclCommSplit");
    MDBG_NCCL_SYM(GetErrorString, "ncclGetErrorString");
    MDBG_NCCL_SYM(GroupStart, "ncclGroupStart");
    MDBG_NCCL_SYM(GroupEnd, "ncclGroupEnd");
    MDBG_NCCL_SYM(Send, "ncclSend");
    MDBG_NCCL_SYM(Recv, "ncclRecv");
    MDBG_NCCL_SYM(AllGather, "ncclAllGather");
    MDBG_NCCL_SYM(AllReduce, "ncclAllReduce");
#undef MDBG_NCCL_SYM
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.GetErrorString && api.GroupStart &&
             api.GroupEnd && api.Send && api.Recv && api.AllGather && api.AllReduce;
    return api;
}

}  // namespace mdbg
