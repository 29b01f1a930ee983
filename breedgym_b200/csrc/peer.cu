// Reward exchange of the env-sharded vector env over peer memory (include/breedgym_b200.h, "reward exchange over
// peer memory").
//
// Replaces the reward half of DistributedBreedGym.step_wait (breedgym/vector/vec_env.py:197-219: every shard's rewards
// travel through a host pipe).  The exchange is part of the kernel that PRODUCES the rewards: the block that reduces
// env e's GEBVs to max(GEBV) (vec_env.py:95-100) stores the value into the receive window of every rank -- plain
// peer-to-peer stores, which NVSwitch carries at full bandwidth to any GPU of the box -- and the last block to finish
// raises this rank's epoch flag in every window (system-scope release).  There is no collective launch, no second
// kernel and no host work at an episode's end; bg_peer_wait is a one-warp kernel for consumers.
//
// Window (one per rank, cudaMalloc, exported with CUDA IPC; never freed, so a late peer cannot fault):
//   [0, 256)      uint32 flags[r]  = last epoch rank r has published here
//   [256, 512)    private: arrival counter of the publishing grid, timeout counter
//   [512, ...)    float data[2][total]: epoch k lives in half k & 1; rank r owns [offset_r, offset_r + count_r)
// Flow control: epoch k is stored only after the local flags show that every rank has published epoch k - 1 (it has
// then ended the episode in which it could still read epoch k - 2, whose half is about to be overwritten).
#include <string.h>
#include <unistd.h>

#include <atomic>
#include <chrono>

#include "bg_internal.h"

namespace {

constexpr size_t WIN_FLAGS = 0, WIN_PRIV = 256, WIN_DATA = 512;
constexpr unsigned FULL = 0xffffffffu;

struct PeerArgs {
    uint8_t *win[BG_PEER_MAX_WORLD];
    int world, rank;
    uint32_t epoch;
    long long total, offset;
    long long timeout_cycles;
};

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// spin until flags[r] has reached `epoch` (wrap-safe); false: the bound was hit
__device__ bool wait_flag(const uint32_t *flag, uint32_t epoch, long long timeout_cycles)
{
    if ((int32_t)(ld_acquire_sys(flag) - epoch) >= 0) return true;
    const long long t0 = clock64();
    for (;;) {
        __nanosleep(100);
        if ((int32_t)(ld_acquire_sys(flag) - epoch) >= 0) return true;
        if (clock64() - t0 > timeout_cycles) return false;
    }
}

// tail shared by the two publishing kernels: thread 0 of a block has the value of slot `e`
__device__ void publish_value(const PeerArgs &pa, int64_t e, float v, unsigned nblocks)
{
    const size_t slot = WIN_DATA + sizeof(float) * ((size_t)(pa.epoch & 1) * pa.total + (size_t)pa.offset + e);
    for (int p = 0; p < pa.world; ++p) *reinterpret_cast<float *>(pa.win[p] + slot) = v;
    __threadfence_system();
    unsigned *counter = reinterpret_cast<unsigned *>(pa.win[pa.rank] + WIN_PRIV);
    if (atomicAdd(counter, 1u) == nblocks - 1) {  // every block's stores are out: raise the flag everywhere
        *counter = 0;
        __threadfence_system();
        for (int p = 0; p < pa.world; ++p) st_release_sys(reinterpret_cast<uint32_t *>(pa.win[p] + WIN_FLAGS) + pa.rank, pa.epoch);
    }
}

// flow control, thread 0 of every block: the half about to be overwritten holds epoch - 2
__device__ void wait_previous_epoch(const PeerArgs &pa)
{
    if (pa.epoch < 2) return;
    const uint32_t *flags = reinterpret_cast<const uint32_t *>(pa.win[pa.rank] + WIN_FLAGS);
    for (int p = 0; p < pa.world; ++p)
        if (!wait_flag(flags + p, pa.epoch - 1, pa.timeout_cycles))
            atomicAdd(reinterpret_cast<unsigned *>(pa.win[pa.rank] + WIN_PRIV) + 1, 1u);
}

// max reward of env blockIdx.x (the same arithmetic as reduce_env_kernel, gebv.cu) + publication
__global__ void __launch_bounds__(256) reduce_publish_kernel(const float *__restrict__ in, int64_t per_env, float *__restrict__ out,
                                                             const PeerArgs pa)
{
    __shared__ double sh[8];
    const float *x = in + (int64_t)blockIdx.x * per_env;
    double v = -INFINITY;
    for (int64_t i = threadIdx.x; i < per_env; i += blockDim.x) v = fmax(v, (double)x[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(FULL, v, o));
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) v = fmax(v, sh[w]);
        const float r = (float)v;
        out[blockIdx.x] = r;
        wait_previous_epoch(pa);
        publish_value(pa, blockIdx.x, r, gridDim.x);
    }
}

__global__ void publish_kernel(const float *__restrict__ in, const PeerArgs pa)
{
    wait_previous_epoch(pa);
    publish_value(pa, blockIdx.x, in[blockIdx.x], gridDim.x);
}

__global__ void peer_wait_kernel(uint8_t *win, int world, uint32_t epoch, long long timeout_cycles)
{
    const int p = threadIdx.x;
    if (p < world && !wait_flag(reinterpret_cast<const uint32_t *>(win + WIN_FLAGS) + p, epoch, timeout_cycles))
        atomicAdd(reinterpret_cast<unsigned *>(win + WIN_PRIV) + 1, 1u);
}

// identifies this process in a handle: ranks of the same process connect through raw pointers (CUDA IPC cannot open
// a handle in the process that exported it)
uint64_t process_nonce()
{
    static const uint64_t nonce = [] {
        const uint64_t t = (uint64_t)std::chrono::steady_clock::now().time_since_epoch().count();
        return (t * 0x9E3779B97F4A7C15ull) ^ ((uint64_t)getpid() << 32) ^ 0xB200u;
    }();
    return nonce;
}

struct Handle {  // BG_PEER_HANDLE_BYTES
    cudaIpcMemHandle_t ipc;  // 64 bytes
    uint64_t nonce;
    uint64_t ptr;
    int32_t device;
    int32_t world;
    int64_t total;
};
static_assert(sizeof(Handle) == BG_PEER_HANDLE_BYTES, "handle layout");

}  // namespace

struct bg_peer {
    int device = 0, world = 1, rank = 0;
    int64_t total = 0, offset = 0;
    size_t bytes = 0;
    uint8_t *win = nullptr;                       // this rank's window
    uint8_t *peer_win[BG_PEER_MAX_WORLD] = {};   // every rank's window as seen from this device
    bool imported[BG_PEER_MAX_WORLD] = {};
    bool connected = false;
    uint32_t epoch = 0;
    long long timeout_cycles = 0;
    int clock_khz = 1965000;
    bg_engine *attached = nullptr;
};

namespace {

struct DevGuard {
    int prev = -1;
    bool ok = true;
    explicit DevGuard(int dev)
    {
        cudaGetDevice(&prev);
        if (prev != dev) ok = cudaSetDevice(dev) == cudaSuccess;
        else prev = -1;
    }
    ~DevGuard()
    {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

PeerArgs make_args(bg_peer *p)
{
    PeerArgs pa;
    for (int r = 0; r < BG_PEER_MAX_WORLD; ++r) pa.win[r] = r < p->world ? p->peer_win[r] : nullptr;
    pa.world = p->world;
    pa.rank = p->rank;
    pa.epoch = p->epoch;
    pa.total = p->total;
    pa.offset = p->offset;
    pa.timeout_cycles = p->timeout_cycles;
    return pa;
}

}  // namespace

// bg_vec_step's reward reduction when a peer exchange is attached to the engine
int bg_launch_reduce_publish(const float *in, int64_t E, int64_t per_env, float *out, bg_peer *p, cudaStream_t st)
{
    BG_REQUIRE(p && p->connected, BG_ESTATE, "peer exchange is not connected");
    BG_REQUIRE(E > 0 && p->offset + E <= p->total && per_env > 0, BG_EINVAL, "peer exchange: more envs than the window holds");
    ++p->epoch;
    reduce_publish_kernel<<<(unsigned)E, 256, 0, st>>>(in, per_env, out, make_args(p));
    BG_LAUNCHED();
    return BG_OK;
}

void bg_peer_engine_gone(bg_engine *eng)
{
    if (eng && eng->peer) {
        eng->peer->attached = nullptr;
        eng->peer = nullptr;
    }
}

extern "C" {

int bg_peer_create(bg_engine *eng, int world, int rank, int64_t total, int64_t offset, bg_peer **out)
{
    BG_REQUIRE(eng && out, BG_EINVAL, "bg_peer_create: null argument");
    BG_REQUIRE(world >= 1 && world <= BG_PEER_MAX_WORLD && rank >= 0 && rank < world, BG_EINVAL, "bg_peer_create: bad world / rank");
    BG_REQUIRE(total > 0 && total < (int64_t(1) << 26) && offset >= 0 && offset < total, BG_EINVAL, "bg_peer_create: bad window size / offset");
    DevGuard g(eng->device);
    BG_REQUIRE(g.ok, BG_ECUDA, "cudaSetDevice failed");
    bg_peer *p = new (std::nothrow) bg_peer();
    BG_REQUIRE(p, BG_ENOMEM, "out of host memory");
    p->device = eng->device;
    p->world = world;
    p->rank = rank;
    p->total = total;
    p->offset = offset;
    p->bytes = WIN_DATA + sizeof(float) * 2 * (size_t)total;
    if (cudaMalloc(&p->win, p->bytes) != cudaSuccess || cudaMemset(p->win, 0, p->bytes) != cudaSuccess ||
        cudaDeviceSynchronize() != cudaSuccess) {
        cudaGetLastError();
        delete p;
        bg_set_error("bg_peer_create: cannot allocate the receive window");
        return BG_ENOMEM;
    }
    cudaDeviceGetAttribute(&p->clock_khz, cudaDevAttrClockRate, eng->device);
    if (p->clock_khz <= 0) p->clock_khz = 1965000;
    p->timeout_cycles = 20000LL * p->clock_khz;
    *out = p;
    return BG_OK;
}

int bg_peer_handle(bg_peer *p, uint8_t handle_out[BG_PEER_HANDLE_BYTES])
{
    BG_REQUIRE(p && handle_out, BG_EINVAL, "bg_peer_handle: null argument");
    DevGuard g(p->device);
    Handle h;
    memset(&h, 0, sizeof(h));
    // (same-process ranks never open it; a failure here only matters once another process tries to)
    if (cudaIpcGetMemHandle(&h.ipc, p->win) != cudaSuccess) {
        cudaGetLastError();
        memset(&h.ipc, 0, sizeof(h.ipc));
    }
    h.nonce = process_nonce();
    h.ptr = (uint64_t)(uintptr_t)p->win;
    h.device = p->device;
    h.world = p->world;
    h.total = p->total;
    memcpy(handle_out, &h, sizeof(h));
    return BG_OK;
}

int bg_peer_connect(bg_peer *p, const uint8_t *handles)
{
    BG_REQUIRE(p && handles, BG_EINVAL, "bg_peer_connect: null argument");
    BG_REQUIRE(!p->connected, BG_ESTATE, "bg_peer_connect: already connected");
    DevGuard g(p->device);
    BG_REQUIRE(g.ok, BG_ECUDA, "cudaSetDevice failed");
    for (int r = 0; r < p->world; ++r) {
        Handle h;
        memcpy(&h, handles + (size_t)r * BG_PEER_HANDLE_BYTES, sizeof(h));
        BG_REQUIRE(h.world == p->world && h.total == p->total, BG_EINVAL, "bg_peer_connect: ranks disagree on world / window size");
        if (r == p->rank) {
            BG_REQUIRE(h.ptr == (uint64_t)(uintptr_t)p->win, BG_EINVAL, "bg_peer_connect: handle table is not in rank order");
            p->peer_win[r] = p->win;
        } else if (h.nonce == process_nonce()) {  // a rank of this very process: its pointer is valid here
            if (h.device != p->device) {
                int can = 0;
                cudaDeviceCanAccessPeer(&can, p->device, h.device);
                BG_REQUIRE(can, BG_ESTATE, "bg_peer_connect: no peer access between two devices of this process");
                const cudaError_t e = cudaDeviceEnablePeerAccess(h.device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return bg_cuda_fail(e, "cudaDeviceEnablePeerAccess");
                cudaGetLastError();
            }
            p->peer_win[r] = reinterpret_cast<uint8_t *>((uintptr_t)h.ptr);
        } else {
            void *ptr = nullptr;
            const cudaError_t e = cudaIpcOpenMemHandle(&ptr, h.ipc, cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) {
                cudaGetLastError();
                for (int q = 0; q < r; ++q)
                    if (p->imported[q]) {
                        cudaIpcCloseMemHandle(p->peer_win[q]);
                        p->imported[q] = false;
                    }
                return bg_cuda_fail(e, "cudaIpcOpenMemHandle (peer window)");
            }
            p->peer_win[r] = static_cast<uint8_t *>(ptr);
            p->imported[r] = true;
        }
    }
    p->connected = true;
    return BG_OK;
}

int bg_engine_set_peer(bg_engine *eng, bg_peer *peer)
{
    BG_REQUIRE(eng, BG_EINVAL, "null engine");
    BG_REQUIRE(!peer || (peer->connected && peer->device == eng->device), BG_ESTATE,
               "bg_engine_set_peer: the exchange is not connected or lives on another device");
    if (eng->peer) eng->peer->attached = nullptr;
    eng->peer = peer;
    if (peer) peer->attached = eng;
    return BG_OK;
}

int bg_peer_publish_f32(bg_peer *p, const float *send_dev, int64_t count, void *stream)
{
    BG_REQUIRE(p && p->connected, BG_ESTATE, "bg_peer_publish_f32: the exchange is not connected");
    BG_REQUIRE(send_dev && count > 0 && p->offset + count <= p->total, BG_EINVAL, "bg_peer_publish_f32: bad argument");
    DevGuard g(p->device);
    ++p->epoch;
    publish_kernel<<<(unsigned)count, 1, 0, (cudaStream_t)stream>>>(send_dev, make_args(p));
    BG_LAUNCHED();
    return BG_OK;
}

int bg_peer_wait(bg_peer *p, void *stream)
{
    BG_REQUIRE(p && p->connected, BG_ESTATE, "bg_peer_wait: the exchange is not connected");
    if (p->epoch == 0) return BG_OK;
    DevGuard g(p->device);
    peer_wait_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(p->win, p->world, p->epoch, p->timeout_cycles);
    BG_LAUNCHED();
    return BG_OK;
}

int64_t bg_peer_epoch(bg_peer *p) { return p ? (int64_t)p->epoch : -1; }

float *bg_peer_result(bg_peer *p, int parity)
{
    if (!p) return nullptr;
    return reinterpret_cast<float *>(p->win + WIN_DATA) + (size_t)(parity & 1) * p->total;
}

int bg_peer_set_timeout_ms(bg_peer *p, int64_t ms)
{
    BG_REQUIRE(p && ms > 0, BG_EINVAL, "bg_peer_set_timeout_ms: bad argument");
    p->timeout_cycles = (long long)ms * p->clock_khz;
    return BG_OK;
}

int64_t bg_peer_timeouts(bg_peer *p)
{
    if (!p) return -1;
    DevGuard g(p->device);
    unsigned n = 0;
    if (cudaDeviceSynchronize() != cudaSuccess || cudaMemcpy(&n, p->win + WIN_PRIV + 4, 4, cudaMemcpyDeviceToHost) != cudaSuccess) {
        cudaGetLastError();
        return -1;
    }
    return (int64_t)n;
}

int bg_peer_destroy(bg_peer *p)
{
    if (!p) return BG_OK;
    DevGuard g(p->device);
    cudaDeviceSynchronize();  // this rank's own publishing kernels
    if (p->attached && p->attached->peer == p) p->attached->peer = nullptr;
    for (int r = 0; r < p->world; ++r)
        if (p->imported[r]) cudaIpcCloseMemHandle(p->peer_win[r]);
    // p->win is deliberately NOT freed: a slower rank may still store its last epoch into it (a few hundred bytes per
    // exchange, reclaimed when the process exits)
    cudaGetLastError();
    delete p;
    return BG_OK;
}

}  // extern "C"
