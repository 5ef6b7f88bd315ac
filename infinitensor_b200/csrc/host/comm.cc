// comm.cc -- NCCL communicator for the in-graph collectives (AllReduce / AllGather).
// Same rendezvous contract as the reference (include/cuda/nccl_communicator.h:22-67): rank 0 creates the
// ncclUniqueId and publishes it in ./<name>_nccl_id.bin, the other ranks poll for the file (<= 10 s);
// plus a variant that takes the id bytes from the caller (torchrun-style launchers broadcast it).
// Errors throw infini::Exception instead of the reference's exit(EXIT_FAILURE) macros.
#include <dlfcn.h>

#include <chrono>
#include <cstdio>
#include <fstream>
#include <thread>
#include <mutex>
#include <cstdlib>

#include "b200_runtime.h"
#include "nccl_dl.h"

namespace infini {

const NcclApi &nccl() {
    static NcclApi api{};
    static bool ok = false;
    static std::once_flag once;
    std::call_once(once, [] {
        void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);  // already in the process (torch)?
        if (!h)
            if (const char *p = std::getenv("ITB_NCCL_LIB")) h = dlopen(p, RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) return;
#define LOAD(field, sym) api.field = reinterpret_cast<decltype(api.field)>(dlsym(h, sym))
        LOAD(GetUniqueId, "ncclGetUniqueId");
        LOAD(CommInitRank, "ncclCommInitRank");
        LOAD(CommDestroy, "ncclCommDestroy");
        LOAD(AllReduce, "ncclAllReduce");
        LOAD(AllGather, "ncclAllGather");
        LOAD(GetErrorString, "ncclGetErrorString");
        LOAD(GetVersion, "ncclGetVersion");
#undef LOAD
        ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce && api.AllGather &&
             api.GetErrorString;
    });
    IT_ASSERT(ok, "NCCL (libnccl.so.2) could not be loaded; set ITB_NCCL_LIB");
    return api;
}

#define checkNcclError(call)                                                                   \
    do {                                                                                       \
        ncclResult_t r__ = (call);                                                             \
        if (r__ != ncclSuccess)                                                                \
            throw ::infini::Exception(string("NCCL error: ") + nccl().GetErrorString(r__) + " at " + __FILE__ + ":" + \
                                      std::to_string(__LINE__));                               \
    } while (0)

class NcclCommunicatorObj final : public CommunicatorObj {
    ncclComm_t comm = nullptr;

  public:
    NcclCommunicatorObj(const ncclUniqueId &id, int worldSize, int rank) : CommunicatorObj(worldSize, rank) {
        checkNcclError(nccl().CommInitRank(&comm, worldSize, id, rank));
    }
    ~NcclCommunicatorObj() override {
        if (comm) nccl().CommDestroy(comm);
    }
    void *getNcclComm() const override { return comm; }
};

Ref<CommunicatorObj> makeNcclCommunicator(const string &name, int worldSize, int rank) {
    const string path = "./" + name + "_nccl_id.bin";
    ncclUniqueId id;
    if (rank == 0) {
        checkNcclError(nccl().GetUniqueId(&id));
        const string tmp = path + ".tmp";
        {
            std::ofstream f(tmp, std::ios::binary);
            f.write(reinterpret_cast<const char *>(&id), sizeof(id));
        }
        std::rename(tmp.c_str(), path.c_str());  // atomic publish
    } else {
        auto t0 = std::chrono::steady_clock::now();
        for (;;) {
            std::ifstream f(path, std::ios::binary);
            if (f && f.read(reinterpret_cast<char *>(&id), sizeof(id))) break;
            IT_ASSERT(std::chrono::steady_clock::now() - t0 < std::chrono::seconds(10),
                      "timed out waiting for " + path);
            std::this_thread::sleep_for(std::chrono::milliseconds(20));
        }
    }
    auto c = make_ref<NcclCommunicatorObj>(id, worldSize, rank);
    if (rank == 0) std::remove(path.c_str());
    return c;
}

Ref<CommunicatorObj> makeNcclCommunicatorWithId(const void *idBytes, int n, int worldSize, int rank) {
    IT_ASSERT(n == (int)sizeof(ncclUniqueId), "bad ncclUniqueId size");
    ncclUniqueId id;
    std::memcpy(&id, idBytes, sizeof(id));
    return make_ref<NcclCommunicatorObj>(id, worldSize, rank);
}

int ncclUniqueIdBytes(void *out, int outBytes) {
    if (outBytes < (int)sizeof(ncclUniqueId)) return -(int)sizeof(ncclUniqueId);
    ncclUniqueId id;
    checkNcclError(nccl().GetUniqueId(&id));
    std::memcpy(out, &id, sizeof(id));
    return (int)sizeof(id);
}

}  // namespace infini
