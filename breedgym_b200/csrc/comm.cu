// Reward all-gather of the env-sharded vector env (include/breedgym_b200.h, "reward all-gather").
//
// Replaces the host-pipe exchange of DistributedBreedGym (breedgym/vector/vec_env.py:197-219): the only data that
// crosses GPUs is float32[E/G] rewards per rank, gathered with ncclAllGather over NVLink on the step's own stream.
// NCCL is bound at run time (dlopen of the libnccl.so.2 the process already carries -- torch's -- else the system's),
// so the library links against nothing but the CUDA runtime and still loads on a box without NCCL.
#include <dlfcn.h>
#include <string.h>

#include <mutex>

#include "bg_internal.h"

namespace {

struct NcclUniqueId {
    char internal[BG_COMM_ID_BYTES];
};
typedef struct ncclComm *ncclComm_t;
typedef int ncclResult_t;  // ncclSuccess = 0
constexpr int NCCL_FLOAT32 = 7;

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(NcclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, NcclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    std::string why;
};

NcclApi *nccl()
{
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char *names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char *n : names)  // the copy already mapped into the process (torch's), if any
            if (!api.handle) api.handle = dlopen(n, RTLD_NOW | RTLD_NOLOAD);
        for (const char *n : names)
            if (!api.handle) api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (!api.handle) {
            api.why = "libnccl.so.2 not found (dlopen)";
            return;
        }
        api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(api.handle, "ncclGetUniqueId"));
        api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(api.handle, "ncclCommInitRank"));
        api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(api.handle, "ncclCommDestroy"));
        api.AllGather = reinterpret_cast<decltype(api.AllGather)>(dlsym(api.handle, "ncclAllGather"));
        api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(api.handle, "ncclGetErrorString"));
        if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllGather || !api.GetErrorString) {
            api.why = "libnccl is missing a required symbol";
            api.handle = nullptr;
        }
    });
    return &api;
}

int nccl_fail(NcclApi *api, ncclResult_t r, const char *what)
{
    bg_set_error(std::string(what) + ": " + (api->GetErrorString ? api->GetErrorString(r) : "NCCL error"));
    return BG_ECUDA;
}

}  // namespace

struct bg_comm {
    int device = 0;
    int world = 1, rank = 0;
    ncclComm_t comm = nullptr;
    float *warm = nullptr;
};

extern "C" {

int bg_comm_unique_id(uint8_t id_out[BG_COMM_ID_BYTES])
{
    BG_REQUIRE(id_out, BG_EINVAL, "bg_comm_unique_id: null argument");
    NcclApi *api = nccl();
    BG_REQUIRE(api->handle, BG_ESTATE, api->why);
    NcclUniqueId id;
    const ncclResult_t r = api->GetUniqueId(&id);
    if (r) return nccl_fail(api, r, "ncclGetUniqueId");
    memcpy(id_out, id.internal, BG_COMM_ID_BYTES);
    return BG_OK;
}

int bg_comm_create(bg_engine *eng, const uint8_t id[BG_COMM_ID_BYTES], int world, int rank, bg_comm **out)
{
    BG_REQUIRE(eng && id && out, BG_EINVAL, "bg_comm_create: null argument");
    BG_REQUIRE(world >= 1 && rank >= 0 && rank < world, BG_EINVAL, "bg_comm_create: bad world / rank");
    NcclApi *api = nccl();
    BG_REQUIRE(api->handle, BG_ESTATE, api->why);
    int prev = -1;
    cudaGetDevice(&prev);
    BG_CUDA(cudaSetDevice(eng->device));
    bg_comm *c = new (std::nothrow) bg_comm();
    BG_REQUIRE(c, BG_ENOMEM, "out of host memory");
    c->device = eng->device;
    c->world = world;
    c->rank = rank;
    NcclUniqueId uid;
    memcpy(uid.internal, id, BG_COMM_ID_BYTES);
    ncclResult_t r = api->CommInitRank(&c->comm, world, uid, rank);
    if (r) {
        delete c;
        if (prev >= 0) cudaSetDevice(prev);
        return nccl_fail(api, r, "ncclCommInitRank");
    }
    // warm-up: NCCL sets its channels / proxies up lazily on the first collective (milliseconds)
    int rc = BG_OK;
    if (cudaMalloc(&c->warm, sizeof(float) * (size_t)(world + 1)) != cudaSuccess) rc = BG_ENOMEM;
    if (!rc) {
        cudaMemset(c->warm, 0, sizeof(float) * (size_t)(world + 1));
        r = api->AllGather(c->warm + world, c->warm, 1, NCCL_FLOAT32, c->comm, nullptr);
        if (r) rc = nccl_fail(api, r, "ncclAllGather (warm-up)");
        else if (cudaStreamSynchronize(nullptr) != cudaSuccess) rc = BG_ECUDA;
    }
    if (prev >= 0) cudaSetDevice(prev);
    if (rc) {
        bg_comm_destroy(c);
        return rc;
    }
    *out = c;
    return BG_OK;
}

int bg_comm_destroy(bg_comm *c)
{
    if (!c) return BG_OK;
    NcclApi *api = nccl();
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(c->device);
    if (c->comm && api->handle) api->CommDestroy(c->comm);
    cudaFree(c->warm);
    if (prev >= 0) cudaSetDevice(prev);
    delete c;
    return BG_OK;
}

int bg_allgather_f32(bg_comm *c, const float *send_dev, float *recv_dev, int64_t count, void *stream)
{
    BG_REQUIRE(c && c->comm, BG_EINVAL, "bg_allgather_f32: null communicator");
    BG_REQUIRE(count >= 0 && (count == 0 || (send_dev && recv_dev)), BG_EINVAL, "bg_allgather_f32: bad argument");
    if (count == 0) return BG_OK;
    NcclApi *api = nccl();
    int prev = -1;
    cudaGetDevice(&prev);
    if (prev != c->device) BG_CUDA(cudaSetDevice(c->device));
    const ncclResult_t r = api->AllGather(send_dev, recv_dev, (size_t)count, NCCL_FLOAT32, c->comm, (cudaStream_t)stream);
    if (prev >= 0 && prev != c->device) cudaSetDevice(prev);
    if (r) return nccl_fail(api, r, "ncclAllGather");
    return BG_OK;
}

}  // extern "C"
