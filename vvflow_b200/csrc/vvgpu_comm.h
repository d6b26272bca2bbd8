// Multi-GPU transport of libvvgpu: ONE primitive, an all-gather of equal-sized blocks on the
// context's stream, in two forms:
//   * NCCL, for one process per GPU (torchrun): libnccl.so.2 is looked up at run time (dlopen), so a single-GPU
//     user needs no NCCL at all; the host layer only has to carry the 128-byte unique id from rank 0 to the others;
//   * an in-process group, for ONE process that drives several contexts from one thread each (what the vvflow
//     binary itself would do): every rank copies its block straight into its peers' receive buffers with
//     cudaMemcpyPeerAsync (NVLink P2P when the devices differ), ordered by events, two host barriers per collective
//     and no device synchronisation. Ranks may share a device, which is how the sharded step is tested on a
//     one-GPU box.
// Sums over ranks (fric, counters) are an all-gather followed by a local sum in rank order: deterministic and
// bit-identical on every rank, which an all-reduce is not required to be.
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <condition_variable>
#include <mutex>
#include <string>
#include <vector>

namespace vv {

struct NcclApi {
    struct UniqueId { char internal[128]; };   // ncclUniqueId (NCCL_UNIQUE_ID_BYTES = 128)
    void* lib = nullptr;
    int (*GetUniqueId)(UniqueId*) = nullptr;
    int (*CommInitRank)(void**, int, UniqueId, int) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int /*ncclDataType_t*/, void*, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;

    static NcclApi* get(std::string* err) {
        static NcclApi api;
        static std::mutex mu;
        std::lock_guard<std::mutex> lk(mu);
        if (api.lib) return &api;
        // the copy a host framework already loaded (torch bundles its own) wins over the system one
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) { if (err) *err = std::string("libnccl.so.2 not found: ") + dlerror(); return nullptr; }
        api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(h, "ncclGetUniqueId");
        api.CommInitRank = (decltype(api.CommInitRank))dlsym(h, "ncclCommInitRank");
        api.CommDestroy = (decltype(api.CommDestroy))dlsym(h, "ncclCommDestroy");
        api.AllGather = (decltype(api.AllGather))dlsym(h, "ncclAllGather");
        api.GetErrorString = (decltype(api.GetErrorString))dlsym(h, "ncclGetErrorString");
        if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllGather || !api.GetErrorString) {
            if (err) *err = "libnccl.so.2 lacks a required symbol";
            return nullptr;
        }
        api.lib = h;
        return &api;
    }
};

// the threads of an in-process group meet here
struct LocalGroup {
    struct Slot {
        const void* send = nullptr;
        void* recv = nullptr;
        int device = 0;
        cudaStream_t stream = nullptr;
        cudaEvent_t ready = nullptr, done = nullptr;
    };
    int n = 0, refs = 0;
    std::vector<Slot> slot;
    std::mutex mu;
    std::condition_variable cv;
    int arrived = 0;
    long gen = 0;
    bool broken = false;   // a rank failed: release everybody

    void barrier() {
        std::unique_lock<std::mutex> lk(mu);
        const long g = gen;
        if (++arrived == n) { arrived = 0; gen++; cv.notify_all(); }
        else cv.wait(lk, [&] { return gen != g || broken; });
    }
    void abort() {
        std::lock_guard<std::mutex> lk(mu);
        broken = true;
        cv.notify_all();
    }
};

struct Comm {
    enum Kind { NONE = 0, NCCL = 1, LOCAL = 2 } kind = NONE;
    int rank = 0, nranks = 1;
    void* nccl = nullptr;
    LocalGroup* grp = nullptr;
};

// all-gather of `bytes` per rank: recv[r * bytes ...] = rank r's send, on `st`. Returns "" or an error text.
inline std::string comm_allgather(Comm& cm, int device, cudaStream_t st, const void* send, void* recv, size_t bytes) {
    if (cm.kind == Comm::NCCL) {
        NcclApi* api = NcclApi::get(nullptr);
        const int rc = api->AllGather(send, recv, bytes, 0 /* ncclInt8 */, cm.nccl, st);
        if (rc) return std::string("ncclAllGather: ") + api->GetErrorString(rc);
        return "";
    }
    if (cm.kind == Comm::LOCAL) {
        LocalGroup& g = *cm.grp;
        LocalGroup::Slot& me = g.slot[cm.rank];
        me.send = send; me.recv = recv; me.device = device; me.stream = st;
        if (cudaEventRecord(me.ready, st) != cudaSuccess) { g.abort(); return "cudaEventRecord failed"; }
        g.barrier();
        if (g.broken) return "a rank of the group failed";
        for (int p = 0; p < g.n; p++) {
            const LocalGroup::Slot& o = g.slot[p];
            cudaStreamWaitEvent(st, o.ready, 0);   // peer p's receive buffer is free, its own send block final
            if (cudaMemcpyPeerAsync((char*)o.recv + (size_t)cm.rank * bytes, o.device, send, device, bytes, st) != cudaSuccess) {
                g.abort();
                return "cudaMemcpyPeerAsync failed";
            }
        }
        cudaEventRecord(me.done, st);
        g.barrier();
        if (g.broken) return "a rank of the group failed";
        for (int p = 0; p < g.n; p++) cudaStreamWaitEvent(st, g.slot[p].done, 0);
        return "";
    }
    return "no communicator";
}

}  // namespace vv
