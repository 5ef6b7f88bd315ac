// nccl_dl.h -- NCCL is bound at RUN TIME (dlopen), never at link time: a process that also imports torch must
// end up with exactly one libnccl.so.2 (torch bundles its own, newer than the system one), so we take whichever
// is already loaded, else $ITB_NCCL_LIB, else the default search path.
#pragma once
#include <nccl.h>

namespace infini {
struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId *);
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
    const char *(*GetErrorString)(ncclResult_t);
    ncclResult_t (*GetVersion)(int *);
};
const NcclApi &nccl();  // throws infini::Exception if no NCCL can be loaded
}  // namespace infini
