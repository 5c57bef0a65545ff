// Multi-GPU all-pairs keyframe matching (config 5): one process per GPU, the query keyframes are sharded by contiguous
// blocks, and the ONE exchange of the path -- every rank needs every other rank's descriptors as db columns -- is a
// chunked ncclAllGather over NVLink that overlaps the matching of the chunks that have already landed.
//   reference semantics: SearchByBoW(KF, KF) inner loop, ORBmatcher.cc:566-618, 634-652 (see hamming.cu::allpairs_kernel)
//
// NCCL is bound at run time (dlopen of libnccl.so.2, the library torch / the system ships): liborbb200.so has no link
// dependency on it, single-GPU users never load it.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "matcher.h"

using namespace orbb;

namespace {

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
};

int nccl_api(NcclApi** out) {
    static NcclApi api;
    if (!api.lib) {
        // an already loaded libnccl.so.2 (e.g. torch's) is reused by the loader; otherwise the system one is opened
        void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) return fail(ORB_ERR_CUDA, "NCCL is not available: %s", dlerror());
#define ORB_NCCL_SYM(name)                                                                \
    *(void**)(&api.name) = dlsym(lib, "nccl" #name);                                      \
    if (!api.name) return fail(ORB_ERR_CUDA, "libnccl has no symbol nccl" #name)
        ORB_NCCL_SYM(GetUniqueId);
        ORB_NCCL_SYM(CommInitRank);
        ORB_NCCL_SYM(CommDestroy);
        ORB_NCCL_SYM(AllGather);
        ORB_NCCL_SYM(GroupStart);
        ORB_NCCL_SYM(GroupEnd);
        ORB_NCCL_SYM(GetErrorString);
        ORB_NCCL_SYM(GetVersion);
#undef ORB_NCCL_SYM
        api.lib = lib;
    }
    *out = &api;
    return ORB_OK;
}

#define ORB_NCCL(api, call)                                                                                      \
    do {                                                                                                         \
        ncclResult_t r__ = (call);                                                                               \
        if (r__ != ncclSuccess) return fail(ORB_ERR_CUDA, "%s:%d: %s -> %s", __FILE__, __LINE__, #call, (api)->GetErrorString(r__)); \
    } while (0)

}  // namespace

struct orbm_comm_s {
    int device = 0, rank = 0, world = 1;
    ncclComm_t comm = nullptr;
    cudaStream_t commStream = nullptr;
    std::vector<cudaEvent_t> landed;      // one per gathered chunk
    cudaEvent_t ready = nullptr, gatherBegin = nullptr, gatherEnd = nullptr;
    DevBuf stageDesc, stageAngles, padDesc, padAngles;
    size_t lastGatherBytes = 0;           // bytes this rank received in the last call
    int lastChunks = 0;
};

extern "C" {

int orbm_comm_unique_id(uint8_t* id128) {
    if (!id128) return fail(ORB_ERR_INVALID, "orbm_comm_unique_id: null out");
    NcclApi* api;
    ORB_CHECK(nccl_api(&api));
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    ORB_NCCL(api, api->GetUniqueId(&id));
    std::memcpy(id128, &id, 128);
    return ORB_OK;
}

int orbm_comm_create(const uint8_t* id128, int rank, int world, int device, orbm_comm* out) {
    if (!out) return fail(ORB_ERR_INVALID, "orbm_comm_create: null out");
    *out = nullptr;
    if (!id128 || world < 1 || world > 17 || rank < 0 || rank >= world)
        return fail(ORB_ERR_INVALID, "orbm_comm_create: need 1 <= world <= 17 and 0 <= rank < world");
    NcclApi* api;
    ORB_CHECK(nccl_api(&api));
    DeviceGuard g(device);
    if (!g.ok) return fail(ORB_ERR_CUDA, "orbm_comm_create: cannot select device %d", device);
    orbm_comm_s* c = new orbm_comm_s;
    c->device = device;
    c->rank = rank;
    c->world = world;
    ncclUniqueId id;
    std::memcpy(&id, id128, 128);
    ncclResult_t r = api->CommInitRank(&c->comm, world, id, rank);
    if (r != ncclSuccess) {
        delete c;
        return fail(ORB_ERR_CUDA, "ncclCommInitRank(rank %d of %d) -> %s", rank, world, api->GetErrorString(r));
    }
    // highest priority: the gather's few CTAs must not queue behind the matcher's thousands of long-running ones
    int prLow = 0, prHigh = 0;
    cudaDeviceGetStreamPriorityRange(&prLow, &prHigh);
    cudaError_t ce = cudaStreamCreateWithPriority(&c->commStream, cudaStreamNonBlocking, prHigh);
    if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&c->ready, cudaEventDisableTiming);
    if (ce == cudaSuccess) ce = cudaEventCreate(&c->gatherBegin);
    if (ce == cudaSuccess) ce = cudaEventCreate(&c->gatherEnd);
    if (ce != cudaSuccess) {
        orbm_comm_destroy(c);
        return fail(ORB_ERR_CUDA, "orbm_comm_create: %s", cudaGetErrorString(ce));
    }
    *out = c;
    return ORB_OK;
}

int orbm_comm_destroy(orbm_comm c) {
    if (!c) return ORB_OK;
    DeviceGuard g(c->device);
    if (c->commStream) cudaStreamSynchronize(c->commStream);
    NcclApi* api;
    if (c->comm && nccl_api(&api) == ORB_OK) api->CommDestroy(c->comm);
    for (cudaEvent_t e : c->landed) cudaEventDestroy(e);
    if (c->ready) cudaEventDestroy(c->ready);
    if (c->gatherBegin) cudaEventDestroy(c->gatherBegin);
    if (c->gatherEnd) cudaEventDestroy(c->gatherEnd);
    if (c->commStream) cudaStreamDestroy(c->commStream);
    c->stageDesc.release();
    c->stageAngles.release();
    c->padDesc.release();
    c->padAngles.release();
    delete c;
    return ORB_OK;
}

int orbm_comm_info(orbm_comm c, int* rank, int* world, int* ncclVersion) {
    if (!c) return fail(ORB_ERR_INVALID, "orbm_comm_info: null communicator");
    if (rank) *rank = c->rank;
    if (world) *world = c->world;
    if (ncclVersion) {
        NcclApi* api;
        ORB_CHECK(nccl_api(&api));
        ORB_NCCL(api, api->GetVersion(ncclVersion));
    }
    return ORB_OK;
}

int orbm_allpairs_sharded(orbm_handle h, orbm_comm c, const uint8_t* dLocalDesc, const float* dLocalAngles, const int* kfPerRank,
                          int nDesc, int qCount, int chunkKf, float ratio, int checkOri, int* dCounts, void* stream) {
    ORBM_ENTER(h);
    if (!c || !kfPerRank || !dCounts || nDesc < 1 || chunkKf < 1) return fail(ORB_ERR_INVALID, "orbm_allpairs_sharded: bad arguments");
    if (c->device != h->device) return fail(ORB_ERR_INVALID, "orbm_allpairs_sharded: matcher and communicator live on different devices");
    NcclApi* api;
    ORB_CHECK(nccl_api(&api));
    const int world = c->world, rank = c->rank;
    std::vector<int> base(world + 1, 0);
    int maxLocal = 0;
    for (int r = 0; r < world; ++r) {
        if (kfPerRank[r] < 0) return fail(ORB_ERR_INVALID, "orbm_allpairs_sharded: negative block size");
        base[r + 1] = base[r] + kfPerRank[r];
        maxLocal = std::max(maxLocal, kfPerRank[r]);
    }
    const int nLocal = kfPerRank[rank], nKf = base[world];
    const int nQ = qCount < 0 ? nLocal : std::min(qCount, nLocal);   // query keyframes = the first nQ of the local block
    if (nLocal > 0 && (!dLocalDesc || (checkOri && !dLocalAngles))) return fail(ORB_ERR_INVALID, "orbm_allpairs_sharded: null table");
    cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
    c->lastGatherBytes = 0;
    c->lastChunks = 0;
    if (nKf == 0) return ORB_OK;
    const size_t kfDescBytes = (size_t)nDesc * 32, kfAngBytes = (size_t)nDesc * 4;
    const int nChunks = world > 1 ? ceil_div(maxLocal, chunkKf) : 0;
    ORB_CUDA(cudaEventRecord(c->ready, st));                       // the local table is complete at this point of `st`
    if (nChunks > 0) {
        // staging: [chunk][rank][chunkKf] keyframes; a rank's last chunk may be partly (or wholly) padding, which is sent from
        // a zeroed pad buffer and never read
        ORB_CHECK(c->stageDesc.reserve((size_t)nChunks * world * chunkKf * kfDescBytes));
        if (checkOri) ORB_CHECK(c->stageAngles.reserve((size_t)nChunks * world * chunkKf * kfAngBytes));
        ORB_CHECK(c->padDesc.reserve((size_t)chunkKf * kfDescBytes));
        if (checkOri) ORB_CHECK(c->padAngles.reserve((size_t)chunkKf * kfAngBytes));
        while ((int)c->landed.size() < nChunks) {
            cudaEvent_t e;
            ORB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            c->landed.push_back(e);
        }
    }
    // 1. all gathers are enqueued first, on the communicator's own high-priority stream: their few CTAs are placed before
    //    the matcher's thousands of long-running ones (enqueued behind a matcher that already fills the SMs, every gather
    //    waited ~20 ms for two ranks' NCCL kernels to be resident at the same time: 590 ms of gathers per 1.2 s step)
    cudaStream_t cs = c->commStream;
    if (nChunks > 0) {
        ORB_CUDA(cudaStreamWaitEvent(cs, c->ready, 0));
        ORB_CUDA(cudaEventRecord(c->gatherBegin, cs));
    }
    for (int k = 0; k < nChunks; ++k) {
        const int first = k * chunkKf, have = std::max(0, std::min(chunkKf, nLocal - first));
        const uint8_t* sendD = dLocalDesc + (size_t)first * kfDescBytes;
        const float* sendA = dLocalAngles ? dLocalAngles + (size_t)first * nDesc : nullptr;
        if (have < chunkKf) {
            ORB_CUDA(cudaMemsetAsync(c->padDesc.p, 0, (size_t)chunkKf * kfDescBytes, cs));
            if (have > 0) ORB_CUDA(cudaMemcpyAsync(c->padDesc.p, sendD, (size_t)have * kfDescBytes, cudaMemcpyDeviceToDevice, cs));
            sendD = c->padDesc.as<uint8_t>();
            if (checkOri) {
                ORB_CUDA(cudaMemsetAsync(c->padAngles.p, 0, (size_t)chunkKf * kfAngBytes, cs));
                if (have > 0) ORB_CUDA(cudaMemcpyAsync(c->padAngles.p, sendA, (size_t)have * kfAngBytes, cudaMemcpyDeviceToDevice, cs));
                sendA = c->padAngles.as<float>();
            }
        }
        uint8_t* recvD = c->stageDesc.as<uint8_t>() + (size_t)k * world * chunkKf * kfDescBytes;
        float* recvA = checkOri ? c->stageAngles.as<float>() + (size_t)k * world * chunkKf * nDesc : nullptr;
        ORB_NCCL(api, api->GroupStart());
        ORB_NCCL(api, api->AllGather(sendD, recvD, (size_t)chunkKf * kfDescBytes, ncclUint8, c->comm, cs));
        if (checkOri) ORB_NCCL(api, api->AllGather(sendA, recvA, (size_t)chunkKf * nDesc, ncclFloat32, c->comm, cs));
        ORB_NCCL(api, api->GroupEnd());
        ORB_CUDA(cudaEventRecord(c->landed[k], cs));
        c->lastGatherBytes += (size_t)(world - 1) * chunkKf * (kfDescBytes + (checkOri ? kfAngBytes : 0));
    }
    if (nChunks > 0) ORB_CUDA(cudaEventRecord(c->gatherEnd, cs));
    c->lastChunks = nChunks;
    // Matching starts when the whole table has landed (1.8 ms for 147 MB on 2 GPUs).  Letting it run under the gathers
    // (ORBB_SHARD_OVERLAP=1) was measured slower: NCCL's CTAs then queue for SM slots that the matcher's CTAs hold for
    // ~8 ms each, the gathers stretch to 420 ms and the spinning CTAs cost the matcher 4.5 % (experiments/README.md).
    static const bool overlap = getenv("ORBB_SHARD_OVERLAP") && atoi(getenv("ORBB_SHARD_OVERLAP")) != 0;
    if (!overlap && nChunks > 0) ORB_CUDA(cudaStreamWaitEvent(st, c->landed[nChunks - 1], 0));
    // 2. matching: the own block needs no communication and goes first; then chunk by chunk as the gathers land, the other
    //    ranks' parts in ring order
    ORB_CHECK(launch_allpairs_ex(dLocalDesc, dLocalAngles, 0, nQ, dLocalDesc, dLocalAngles, 0, nLocal, nDesc, nKf, base[rank], ratio,
                                 checkOri, dCounts, st, &h->launches));
    for (int k = 0; k < nChunks; ++k) {
        const int first = k * chunkKf;
        const uint8_t* recvD = c->stageDesc.as<uint8_t>() + (size_t)k * world * chunkKf * kfDescBytes;
        const float* recvA = checkOri ? c->stageAngles.as<float>() + (size_t)k * world * chunkKf * nDesc : nullptr;
        ORB_CUDA(cudaStreamWaitEvent(st, c->landed[k], 0));
        // one launch per chunk: the other ranks' parts of it, in ring order, form the launch's db range
        ApSegments sg;
        sg.n = 0;
        sg.start[0] = 0;
        for (int s = 1; s < world && sg.n < 16; ++s) {
            const int r = (rank + s) % world;
            const int cnt = std::max(0, std::min(chunkKf, kfPerRank[r] - first));
            if (cnt == 0) continue;
            sg.row[sg.n] = r * chunkKf;
            sg.col[sg.n] = base[r] + first;
            sg.start[sg.n + 1] = sg.start[sg.n] + cnt;
            ++sg.n;
        }
        if (sg.n == 0 || nQ == 0) continue;
        ORB_CHECK(launch_allpairs_ex(dLocalDesc, dLocalAngles, 0, nQ, recvD, recvA, 0, sg.start[sg.n], nDesc, nKf, 0, ratio, checkOri,
                                     dCounts, st, &h->launches, &sg));
    }
    return ORB_OK;
}

int orbm_comm_last_gather(orbm_comm c, double* ms, double* bytesReceived, int* chunks) {
    if (!c) return fail(ORB_ERR_INVALID, "orbm_comm_last_gather: null communicator");
    DeviceGuard g(c->device);
    float t = 0;
    if (c->lastChunks > 0) {
        ORB_CUDA(cudaEventSynchronize(c->gatherEnd));
        ORB_CUDA(cudaEventElapsedTime(&t, c->gatherBegin, c->gatherEnd));
    }
    if (ms) *ms = t;
    if (bytesReceived) *bytesReceived = (double)c->lastGatherBytes;
    if (chunks) *chunks = c->lastChunks;
    return ORB_OK;
}

}  // extern "C"
