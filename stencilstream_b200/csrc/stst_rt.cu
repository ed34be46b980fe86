/*
 * libstst_rt — the C-ABI device runtime underneath the StencilStream-B200 header templates.
 * Interface and reference counterparts: include/stst_rt.h.
 *
 * Everything here is plain CUDA runtime / driver API; there is no kernel in this file (kernels are
 * header templates instantiated with the user's transition function, see
 * stencilstream_b200/include/StencilStream/cuda/internal/TileKernel.hpp).
 */
#include "stst_rt.h"

#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>
#if defined(__linux__)
    #include <sys/mman.h>
#endif

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <mutex>
#include <thread>
#include <string>
#include <unordered_map>
#include <vector>

namespace {

thread_local std::string g_last_error;

int fail(int code, const char *what, const char *detail) {
    g_last_error = std::string(what) + ": " + (detail ? detail : "unknown error");
    return code == 0 ? -1 : code;
}

#define STST_CUDA(call)                                                                            \
    do {                                                                                           \
        cudaError_t err__ = (call);                                                                \
        if (err__ != cudaSuccess) {                                                                \
            (void)cudaGetLastError(); /* do not leave it for an unrelated later check */           \
            return fail(int(err__), #call, cudaGetErrorString(err__));                             \
        }                                                                                          \
    } while (0)

struct DeviceGuard {
    int previous = -1;
    bool active = false;
    cudaError_t enter(int device) {
        cudaError_t err = cudaGetDevice(&previous);
        if (err != cudaSuccess)
            return err;
        if (previous != device) {
            err = cudaSetDevice(device);
            active = (err == cudaSuccess);
        }
        return err;
    }
    ~DeviceGuard() {
        if (active)
            cudaSetDevice(previous);
    }
};

std::mutex g_stream_mutex;
std::vector<cudaStream_t> g_default_streams;
std::vector<bool> g_pool_configured;

cudaStream_t as_stream(stst_stream_t s) { return static_cast<cudaStream_t>(s); }
cudaEvent_t as_event(stst_event_t e) { return static_cast<cudaEvent_t>(e); }

// ---- NCCL, resolved lazily so that the runtime has no link-time dependency on it ---------------
struct NcclApi {
    void *lib = nullptr;
    bool tried = false;
    int (*GetUniqueId)(void *) = nullptr;
    int (*CommInitRank)(void **, int, /*ncclUniqueId by value*/ struct UidByValue, int) = nullptr;
    int (*CommDestroy)(void *) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*Send)(const void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    int (*Recv)(void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};
struct UidByValue {
    char internal[STST_NCCL_UNIQUE_ID_BYTES];
};
NcclApi g_nccl;
std::mutex g_nccl_mutex;

bool load_nccl() {
    std::lock_guard<std::mutex> lock(g_nccl_mutex);
    if (g_nccl.tried)
        return g_nccl.lib != nullptr;
    g_nccl.tried = true;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *name : names) {
        g_nccl.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.lib)
            break;
    }
    if (!g_nccl.lib)
        return false;
    auto sym = [&](const char *n) { return dlsym(g_nccl.lib, n); };
    g_nccl.GetUniqueId = reinterpret_cast<decltype(g_nccl.GetUniqueId)>(sym("ncclGetUniqueId"));
    g_nccl.CommInitRank = reinterpret_cast<decltype(g_nccl.CommInitRank)>(sym("ncclCommInitRank"));
    g_nccl.CommDestroy = reinterpret_cast<decltype(g_nccl.CommDestroy)>(sym("ncclCommDestroy"));
    g_nccl.GroupStart = reinterpret_cast<decltype(g_nccl.GroupStart)>(sym("ncclGroupStart"));
    g_nccl.GroupEnd = reinterpret_cast<decltype(g_nccl.GroupEnd)>(sym("ncclGroupEnd"));
    g_nccl.Send = reinterpret_cast<decltype(g_nccl.Send)>(sym("ncclSend"));
    g_nccl.Recv = reinterpret_cast<decltype(g_nccl.Recv)>(sym("ncclRecv"));
    g_nccl.GetErrorString =
        reinterpret_cast<decltype(g_nccl.GetErrorString)>(sym("ncclGetErrorString"));
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.CommDestroy || !g_nccl.GroupStart ||
        !g_nccl.GroupEnd || !g_nccl.Send || !g_nccl.Recv) {
        dlclose(g_nccl.lib);
        g_nccl.lib = nullptr;
        return false;
    }
    return true;
}

#define STST_NCCL(call)                                                                            \
    do {                                                                                           \
        int res__ = (call);                                                                        \
        if (res__ != 0) {                                                                          \
            return fail(1000 + res__, #call,                                                       \
                        g_nccl.GetErrorString ? g_nccl.GetErrorString(res__) : "nccl error");      \
        }                                                                                          \
    } while (0)

constexpr int kNcclChar = 0; // ncclInt8 / ncclChar

} // namespace

extern "C" {

int stst_rt_abi_version(void) { return STST_RT_ABI_VERSION; }

const char *stst_last_error(void) { return g_last_error.c_str(); }

int stst_device_count(int *count) {
    if (!count)
        return fail(-1, "stst_device_count", "null argument");
    *count = 0;
    STST_CUDA(cudaGetDeviceCount(count));
    return 0;
}

int stst_get_device_info(int device, stst_device_info *info) {
    if (!info)
        return fail(-1, "stst_get_device_info", "null argument");
    std::memset(info, 0, sizeof(*info));
    cudaDeviceProp prop;
    STST_CUDA(cudaGetDeviceProperties(&prop, device));
    info->sm_count = prop.multiProcessorCount;
    info->cc_major = prop.major;
    info->cc_minor = prop.minor;
    info->max_smem_per_block_optin = int(prop.sharedMemPerBlockOptin);
    info->max_smem_per_sm = int(prop.sharedMemPerMultiprocessor);
    info->l2_bytes = prop.l2CacheSize;
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device);
    info->clock_khz = khz;
    std::strncpy(info->name, prop.name, sizeof(info->name) - 1);
    DeviceGuard guard;
    STST_CUDA(guard.enter(device));
    size_t free_b = 0, total_b = 0;
    STST_CUDA(cudaMemGetInfo(&free_b, &total_b));
    info->free_mem = free_b;
    info->total_mem = total_b;
    return 0;
}

int stst_set_device(int device) {
    STST_CUDA(cudaSetDevice(device));
    return 0;
}

int stst_malloc(int device, size_t bytes, stst_stream_t stream, void **ptr) {
    if (!ptr)
        return fail(-1, "stst_malloc", "null argument");
    *ptr = nullptr;
    DeviceGuard guard;
    STST_CUDA(guard.enter(device));
    {
        // Keep freed blocks cached in the pool: StencilUpdate allocates its ping-pong grids on
        // every call, and returning them to the OS each time would serialise on cudaFree.
        std::lock_guard<std::mutex> lock(g_stream_mutex);
        if (g_pool_configured.size() <= size_t(device))
            g_pool_configured.resize(device + 1, false);
        if (!g_pool_configured[device]) {
            cudaMemPool_t pool;
            if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
                uint64_t threshold = UINT64_MAX;
                cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold);
            }
            g_pool_configured[device] = true;
        }
    }
    if (bytes == 0)
        bytes = 256;
    STST_CUDA(cudaMallocAsync(ptr, bytes, as_stream(stream)));
    return 0;
}

int stst_free(int device, void *ptr, stst_stream_t stream) {
    if (!ptr)
        return 0;
    DeviceGuard guard;
    STST_CUDA(guard.enter(device));
    STST_CUDA(cudaFreeAsync(ptr, as_stream(stream)));
    return 0;
}

int stst_malloc_ipc(int device, size_t bytes, void **ptr) {
    if (!ptr)
        return fail(-1, "stst_malloc_ipc", "null argument");
    DeviceGuard guard;
    STST_CUDA(guard.enter(device));
    STST_CUDA(cudaMalloc(ptr, bytes == 0 ? 256 : bytes));
    return 0;
}

int stst_free_ipc(int device, void *ptr) {
    if (!ptr)
        return 0;
    DeviceGuard guard;
    STST_CUDA(guard.enter(device));
    STST_CUDA(cudaFree(ptr));
    return 0;
}

// Pinned host blocks are expensive to create (cudaHostAlloc of 1 GiB takes hundreds of
// milliseconds), and every result grid that host code looks at needs one. Freed blocks are therefore
// kept, keyed by their exact size, up to STST_PINNED_CACHE_MB (default 16384) in total.
namespace {
std::mutex g_pinned_mutex;
std::multimap<size_t, void *> g_pinned_free;
std::unordered_map<void *, size_t> g_pinned_size;
std::unordered_map<void *, size_t> g_pinned_registered; ///< blocks pinned piecewise: rounded size
size_t g_pinned_cached_bytes = 0;

size_t pinned_cache_limit() {
    static const size_t limit = [] {
        const char *env = std::getenv("STST_PINNED_CACHE_MB");
        const size_t mb = (env && *env) ? size_t(std::strtoull(env, nullptr, 10)) : size_t(16384);
        return mb << 20;
    }();
    return limit;
}
} // namespace

int stst_malloc_host(size_t bytes, void **ptr) {
    if (!ptr)
        return fail(-1, "stst_malloc_host", "null argument");
    if (bytes == 0)
        bytes = 64;
    // STST_PIN_LIMIT_MB: refuse larger requests (callers then use pageable memory and the staged
    // transfer pipeline) — for boxes with little pinnable memory, and to test that path.
    if (const char *env = std::getenv("STST_PIN_LIMIT_MB")) {
        if (*env && bytes > (size_t(std::strtoull(env, nullptr, 10)) << 20))
            return fail(-1, "stst_malloc_host", "request exceeds STST_PIN_LIMIT_MB");
    }
    {
        std::lock_guard<std::mutex> lock(g_pinned_mutex);
        auto it = g_pinned_free.find(bytes);
        if (it != g_pinned_free.end()) {
            *ptr = it->second;
            g_pinned_free.erase(it);
            g_pinned_cached_bytes -= bytes;
            return 0;
        }
    }
    cudaError_t err = cudaHostAlloc(ptr, bytes, cudaHostAllocPortable);
    if (err != cudaSuccess) {
        // Out of pinnable memory: drop the cache and retry once.
        (void)cudaGetLastError();
        (void)stst_host_cache_trim();
        err = cudaHostAlloc(ptr, bytes, cudaHostAllocPortable);
    }
    if (err != cudaSuccess) {
        // Some hosts refuse one large cudaHostAlloc but accept the same pages registered piecewise:
        // ordinary memory (2 MiB aligned, transparent huge pages requested), first touched by THIS
        // thread — so that it lands on the NUMA node the caller is bound to — and page-locked in
        // chunks of at most 1 GiB. All or nothing: a refused chunk undoes the others.
        (void)cudaGetLastError();
        constexpr size_t huge = size_t(2) << 20, chunk = size_t(1) << 30;
        const size_t rounded = (bytes + huge - 1) / huge * huge;
        void *block = std::aligned_alloc(huge, rounded);
        if (!block)
            return fail(-1, "stst_malloc_host", "out of host memory");
#if defined(__linux__) && defined(MADV_HUGEPAGE)
        (void)madvise(block, rounded, MADV_HUGEPAGE);
#endif
        std::memset(block, 0, rounded);
        size_t done = 0;
        for (; done < rounded; done += chunk) {
            const size_t n = std::min(chunk, rounded - done);
            if (cudaHostRegister(static_cast<char *>(block) + done, n, cudaHostRegisterPortable) !=
                cudaSuccess)
                break;
        }
        if (done < rounded) {
            const cudaError_t why = cudaGetLastError();
            for (size_t off = 0; off < done; off += chunk)
                (void)cudaHostUnregister(static_cast<char *>(block) + off);
            std::free(block);
            return fail(int(why ? why : cudaErrorMemoryAllocation), "stst_malloc_host",
                        "cudaHostAlloc and piecewise cudaHostRegister both refused the request");
        }
        *ptr = block;
        std::lock_guard<std::mutex> lock(g_pinned_mutex);
        g_pinned_size[*ptr] = bytes;
        g_pinned_registered[*ptr] = rounded;
        return 0;
    }
    std::lock_guard<std::mutex> lock(g_pinned_mutex);
    g_pinned_size[*ptr] = bytes;
    return 0;
}

namespace {
/// Give a block back to the system: cudaFreeHost, or unregister + free for piecewise-registered ones.
cudaError_t release_pinned_block(void *ptr) {
    size_t registered = 0;
    {
        std::lock_guard<std::mutex> lock(g_pinned_mutex);
        auto it = g_pinned_registered.find(ptr);
        if (it != g_pinned_registered.end()) {
            registered = it->second;
            g_pinned_registered.erase(it);
        }
    }
    if (registered == 0)
        return cudaFreeHost(ptr);
    constexpr size_t chunk = size_t(1) << 30;
    for (size_t off = 0; off < registered; off += chunk)
        (void)cudaHostUnregister(static_cast<char *>(ptr) + off);
    std::free(ptr);
    return cudaSuccess;
}
} // namespace

int stst_free_host(void *ptr) {
    if (!ptr)
        return 0;
    {
        std::lock_guard<std::mutex> lock(g_pinned_mutex);
        auto it = g_pinned_size.find(ptr);
        if (it != g_pinned_size.end() && g_pinned_cached_bytes + it->second <= pinned_cache_limit()) {
            g_pinned_free.emplace(it->second, ptr);
            g_pinned_cached_bytes += it->second;
            return 0;
        }
        if (it != g_pinned_size.end())
            g_pinned_size.erase(it);
    }
    STST_CUDA(release_pinned_block(ptr));
    return 0;
}

int stst_host_cache_trim(void) {
    std::multimap<size_t, void *> blocks;
    {
        std::lock_guard<std::mutex> lock(g_pinned_mutex);
        blocks.swap(g_pinned_free);
        g_pinned_cached_bytes = 0;
        for (auto const &b : blocks)
            g_pinned_size.erase(b.second);
    }
    for (auto const &b : blocks)
        (void)release_pinned_block(b.second);
    return 0;
}

int stst_host_register(void *ptr, size_t bytes) {
    STST_CUDA(cudaHostRegister(ptr, bytes, cudaHostRegisterPortable));
    return 0;
}

int stst_host_unregister(void *ptr) {
    STST_CUDA(cudaHostUnregister(ptr));
    return 0;
}

int stst_memset_async(void *ptr, int value, size_t bytes, stst_stream_t stream) {
    STST_CUDA(cudaMemsetAsync(ptr, value, bytes, as_stream(stream)));
    return 0;
}

int stst_memcpy_h2d_async(void *dst, const void *src, size_t bytes, stst_stream_t stream) {
    STST_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, as_stream(stream)));
    return 0;
}

int stst_memcpy_d2h_async(void *dst, const void *src, size_t bytes, stst_stream_t stream) {
    STST_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, as_stream(stream)));
    return 0;
}

int stst_memcpy_d2d_async(void *dst, const void *src, size_t bytes, stst_stream_t stream) {
    STST_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, as_stream(stream)));
    return 0;
}

int stst_memcpy_2d_async(void *dst, size_t dst_pitch, const void *src, size_t src_pitch,
                         size_t width_bytes, size_t height, int kind, stst_stream_t stream) {
    cudaMemcpyKind k = kind == 0   ? cudaMemcpyHostToDevice
                       : kind == 1 ? cudaMemcpyDeviceToHost
                                   : cudaMemcpyDeviceToDevice;
    if (width_bytes == 0 || height == 0)
        return 0;
    STST_CUDA(cudaMemcpy2DAsync(dst, dst_pitch, src, src_pitch, width_bytes, height, k,
                                as_stream(stream)));
    return 0;
}

// ---- staged copies: pageable host memory <-> device through a small pinned ring -------------------
//
// Pinning a whole grid image is the fastest host side for a transfer, but it is not always possible:
// boxes cap pinnable memory (the round-1 B200 pool: cudaHostAlloc starts failing around 4 GiB although
// RLIMIT_MEMLOCK is unlimited), callers hand in ordinary numpy / std::vector memory, and pinning costs
// ~0.4 s per GiB the first time. A plain cudaMemcpy from pageable memory runs at a fraction of the
// link rate because the driver stages it through one small buffer with a single copying thread. The
// pipeline below does that staging itself: a ring of pinned slots, host-side copies spread over a
// few worker threads, the DMA of slot i overlapping the host copy of slot i+1.
namespace {

class CopyWorkers {
  public:
    static CopyWorkers &instance() {
        static CopyWorkers *pool = new CopyWorkers(); // leaked on purpose: no shutdown-order issues
        return *pool;
    }

    /// dst[0, bytes) = src[0, bytes), split over the workers (the caller copies a share as well).
    void copy(void *dst, const void *src, size_t bytes) {
        std::lock_guard<std::mutex> one_job(call_mutex);
        const size_t n = threads.size() + 1;
        if (bytes < (size_t(4) << 20) || n == 1) {
            std::memcpy(dst, src, bytes);
            return;
        }
        const size_t slice = ((bytes + n - 1) / n + 4095) / 4096 * 4096;
        {
            std::lock_guard<std::mutex> lock(mutex);
            job_dst = static_cast<unsigned char *>(dst);
            job_src = static_cast<const unsigned char *>(src);
            job_bytes = bytes;
            job_slice = slice;
            pending = threads.size();
            generation++;
        }
        wake.notify_all();
        run_slice(0);
        std::unique_lock<std::mutex> lock(mutex);
        done.wait(lock, [&] { return pending == 0; });
    }

  private:
    CopyWorkers() {
        unsigned n = std::max(1u, std::min(8u, std::thread::hardware_concurrency() / 2));
        if (const char *env = std::getenv("STST_COPY_THREADS"))
            n = std::max(1, std::atoi(env));
        for (unsigned i = 1; i < n; i++)
            threads.emplace_back([this, i] { loop(i); }).detach();
    }

    void run_slice(size_t index) {
        const size_t lo = index * job_slice;
        if (lo < job_bytes)
            std::memcpy(job_dst + lo, job_src + lo, std::min(job_slice, job_bytes - lo));
    }

    void loop(size_t index) {
        unsigned long long seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lock(mutex);
                wake.wait(lock, [&] { return generation != seen; });
                seen = generation;
            }
            run_slice(index);
            {
                std::lock_guard<std::mutex> lock(mutex);
                pending--;
            }
            done.notify_one();
        }
    }

    std::vector<std::thread> threads;
    std::mutex call_mutex, mutex;
    std::condition_variable wake, done;
    unsigned char *job_dst = nullptr;
    const unsigned char *job_src = nullptr;
    size_t job_bytes = 0, job_slice = 0, pending = 0;
    unsigned long long generation = 0;
};

struct StagingRing {
    static constexpr int slots = 3;
    size_t slot_bytes = 0;
    void *slot[slots] = {};
    cudaEvent_t idle[slots] = {}; // recorded behind the last DMA that touched the slot
    bool used[slots] = {};
};

std::mutex g_staging_mutex;
std::map<int, StagingRing> g_staging; // per device (events belong to a device)

size_t staging_slot_bytes() {
    static const size_t bytes = [] {
        const char *env = std::getenv("STST_STAGING_SLOT_MB");
        const size_t mb = (env && *env) ? size_t(std::strtoull(env, nullptr, 10)) : size_t(16);
        return std::max<size_t>(mb, 1) << 20;
    }();
    return bytes;
}

bool host_pointer_is_pinned(const void *ptr) {
    cudaPointerAttributes attr{};
    if (cudaPointerGetAttributes(&attr, ptr) != cudaSuccess) {
        (void)cudaGetLastError();
        return false;
    }
    return attr.type == cudaMemoryTypeHost || attr.type == cudaMemoryTypeManaged;
}

} // namespace

int stst_host_memcpy(void *dst, const void *src, size_t bytes) {
    if (bytes != 0 && (!dst || !src))
        return fail(-1, "stst_host_memcpy", "null argument");
    CopyWorkers::instance().copy(dst, src, bytes);
    return 0;
}

int stst_host_is_pinned(const void *ptr, int *pinned) {
    if (!pinned)
        return fail(-1, "stst_host_is_pinned", "null argument");
    *pinned = host_pointer_is_pinned(ptr) ? 1 : 0;
    return 0;
}

int stst_memcpy_2d_staged(void *dev, size_t dev_pitch, void *host, size_t host_pitch,
                          size_t row_bytes, size_t rows, int kind, int device,
                          stst_stream_t stream) {
    if (row_bytes == 0 || rows == 0)
        return 0;
    if (kind != 0 && kind != 1)
        return fail(-1, "stst_memcpy_2d_staged", "kind must be 0 (h2d) or 1 (d2h)");
    const bool h2d = kind == 0;
    const size_t slot_bytes = std::max(staging_slot_bytes(), row_bytes);
    std::lock_guard<std::mutex> lock(g_staging_mutex);
    DeviceGuard guard;
    STST_CUDA(guard.enter(device));
    StagingRing &ring = g_staging[device];
    if (ring.slot_bytes < slot_bytes) {
        for (int i = 0; i < StagingRing::slots; i++) {
            if (ring.slot[i]) {
                if (ring.used[i])
                    STST_CUDA(cudaEventSynchronize(ring.idle[i]));
                STST_CUDA(cudaFreeHost(ring.slot[i]));
                ring.slot[i] = nullptr;
            }
            if (!ring.idle[i])
                STST_CUDA(cudaEventCreateWithFlags(&ring.idle[i], cudaEventDisableTiming));
            STST_CUDA(cudaHostAlloc(&ring.slot[i], slot_bytes, cudaHostAllocPortable));
            ring.used[i] = false;
        }
        ring.slot_bytes = slot_bytes;
    }
    const size_t chunk_rows = std::max<size_t>(1, ring.slot_bytes / row_bytes);
    const size_t n_chunks = (rows + chunk_rows - 1) / chunk_rows;
    unsigned char *hp = static_cast<unsigned char *>(host);
    unsigned char *dp = static_cast<unsigned char *>(dev);
    CopyWorkers &workers = CopyWorkers::instance();
    cudaStream_t cs = as_stream(stream);

    // host rows <-> densely packed rows in a slot
    auto host_copy = [&](int s, size_t row0, size_t n, bool into_slot) {
        unsigned char *slot = static_cast<unsigned char *>(ring.slot[s]);
        unsigned char *h = hp + row0 * host_pitch;
        if (host_pitch == row_bytes) {
            if (into_slot)
                workers.copy(slot, h, n * row_bytes);
            else
                workers.copy(h, slot, n * row_bytes);
        } else {
            for (size_t r = 0; r < n; r++) {
                if (into_slot)
                    std::memcpy(slot + r * row_bytes, h + r * host_pitch, row_bytes);
                else
                    std::memcpy(h + r * host_pitch, slot + r * row_bytes, row_bytes);
            }
        }
    };

    if (h2d) {
        for (size_t c = 0; c < n_chunks; c++) {
            const int s = int(c % StagingRing::slots);
            const size_t row0 = c * chunk_rows, n = std::min(chunk_rows, rows - row0);
            if (ring.used[s])
                STST_CUDA(cudaEventSynchronize(ring.idle[s]));
            host_copy(s, row0, n, true);
            STST_CUDA(cudaMemcpy2DAsync(dp + row0 * dev_pitch, dev_pitch, ring.slot[s], row_bytes,
                                        row_bytes, n, cudaMemcpyHostToDevice, cs));
            STST_CUDA(cudaEventRecord(ring.idle[s], cs));
            ring.used[s] = true;
        }
        return 0;
    }
    auto drain = [&](size_t c) -> int {
        const int s = int(c % StagingRing::slots);
        const size_t row0 = c * chunk_rows, n = std::min(chunk_rows, rows - row0);
        STST_CUDA(cudaEventSynchronize(ring.idle[s]));
        host_copy(s, row0, n, false);
        return 0;
    };
    for (size_t c = 0; c < n_chunks; c++) {
        const int s = int(c % StagingRing::slots);
        const size_t row0 = c * chunk_rows, n = std::min(chunk_rows, rows - row0);
        if (c >= size_t(StagingRing::slots)) {
            if (int st = drain(c - StagingRing::slots))
                return st;
        } else if (ring.used[s]) {
            STST_CUDA(cudaEventSynchronize(ring.idle[s]));
        }
        STST_CUDA(cudaMemcpy2DAsync(ring.slot[s], row_bytes, dp + row0 * dev_pitch, dev_pitch,
                                    row_bytes, n, cudaMemcpyDeviceToHost, cs));
        STST_CUDA(cudaEventRecord(ring.idle[s], cs));
        ring.used[s] = true;
    }
    for (size_t c = n_chunks > size_t(StagingRing::slots) ? n_chunks - StagingRing::slots : 0;
         c < n_chunks; c++) {
        if (int st = drain(c))
            return st;
    }
    return 0;
}

int stst_memcpy_2d_auto(void *dev, size_t dev_pitch, void *host, size_t host_pitch,
                        size_t row_bytes, size_t rows, int kind, int device, stst_stream_t stream) {
    if (row_bytes == 0 || rows == 0)
        return 0;
    if (host_pointer_is_pinned(host)) {
        if (kind == 0)
            return stst_memcpy_2d_async(dev, dev_pitch, host, host_pitch, row_bytes, rows, 0, stream);
        return stst_memcpy_2d_async(host, host_pitch, dev, dev_pitch, row_bytes, rows, 1, stream);
    }
    return stst_memcpy_2d_staged(dev, dev_pitch, host, host_pitch, row_bytes, rows, kind, device,
                                 stream);
}

int stst_memcpy_peer_async(void *dst, int dst_device, const void *src, int src_device,
                           size_t bytes, stst_stream_t stream) {
    if (bytes == 0)
        return 0;
    STST_CUDA(cudaMemcpyPeerAsync(dst, dst_device, src, src_device, bytes, as_stream(stream)));
    return 0;
}

int stst_default_stream(int device, stst_stream_t *stream) {
    if (!stream)
        return fail(-1, "stst_default_stream", "null argument");
    std::lock_guard<std::mutex> lock(g_stream_mutex);
    if (g_default_streams.size() <= size_t(device))
        g_default_streams.resize(device + 1, nullptr);
    if (!g_default_streams[device]) {
        DeviceGuard guard;
        STST_CUDA(guard.enter(device));
        cudaStream_t s;
        STST_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
        g_default_streams[device] = s;
    }
    *stream = g_default_streams[device];
    return 0;
}

int stst_stream_create(int device, int high_priority, stst_stream_t *stream) {
    if (!stream)
        return fail(-1, "stst_stream_create", "null argument");
    DeviceGuard guard;
    STST_CUDA(guard.enter(device));
    int least = 0, greatest = 0;
    STST_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
    cudaStream_t s;
    STST_CUDA(cudaStreamCreateWithPriority(&s, cudaStreamNonBlocking,
                                           high_priority ? greatest : least));
    *stream = s;
    return 0;
}

int stst_stream_destroy(stst_stream_t stream) {
    STST_CUDA(cudaStreamDestroy(as_stream(stream)));
    return 0;
}

int stst_stream_synchronize(stst_stream_t stream) {
    STST_CUDA(cudaStreamSynchronize(as_stream(stream)));
    return 0;
}

int stst_stream_wait_event(stst_stream_t stream, stst_event_t event) {
    STST_CUDA(cudaStreamWaitEvent(as_stream(stream), as_event(event), 0));
    return 0;
}

int stst_event_create(int with_timing, stst_event_t *event) {
    if (!event)
        return fail(-1, "stst_event_create", "null argument");
    cudaEvent_t e;
    STST_CUDA(cudaEventCreateWithFlags(&e, with_timing ? cudaEventDefault : cudaEventDisableTiming));
    *event = e;
    return 0;
}

int stst_event_create_ipc(stst_event_t *event) {
    if (!event)
        return fail(-1, "stst_event_create_ipc", "null argument");
    cudaEvent_t e;
    STST_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming | cudaEventInterprocess));
    *event = e;
    return 0;
}

int stst_event_destroy(stst_event_t event) {
    if (!event)
        return 0;
    STST_CUDA(cudaEventDestroy(as_event(event)));
    return 0;
}

int stst_event_record(stst_event_t event, stst_stream_t stream) {
    STST_CUDA(cudaEventRecord(as_event(event), as_stream(stream)));
    return 0;
}

int stst_event_synchronize(stst_event_t event) {
    STST_CUDA(cudaEventSynchronize(as_event(event)));
    return 0;
}

int stst_event_elapsed_ms(stst_event_t start, stst_event_t stop, float *ms) {
    if (!ms)
        return fail(-1, "stst_event_elapsed_ms", "null argument");
    STST_CUDA(cudaEventElapsedTime(ms, as_event(start), as_event(stop)));
    return 0;
}

int stst_device_synchronize(int device) {
    DeviceGuard guard;
    STST_CUDA(guard.enter(device));
    STST_CUDA(cudaDeviceSynchronize());
    return 0;
}

// ---- stream-ordered flags (cuStreamWriteValue32 / cuStreamWaitValue32) ----------------------------
namespace {
using StreamValueFn = CUresult (*)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
StreamValueFn g_write_value32 = nullptr, g_wait_value32 = nullptr;

int load_stream_value_ops() {
    if (g_write_value32 && g_wait_value32)
        return 0;
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    STST_CUDA(cudaGetDriverEntryPoint("cuStreamWriteValue32", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess)
        return fail(-1, "stst_stream_write_value32", "cuStreamWriteValue32 unavailable");
    g_write_value32 = reinterpret_cast<StreamValueFn>(fn);
    fn = nullptr;
    STST_CUDA(cudaGetDriverEntryPoint("cuStreamWaitValue32", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess)
        return fail(-1, "stst_stream_wait_value32_geq", "cuStreamWaitValue32 unavailable");
    g_wait_value32 = reinterpret_cast<StreamValueFn>(fn);
    return 0;
}

int driver_fail(const char *what, CUresult res) {
    return fail(int(res), what, ("CUDA driver error " + std::to_string(int(res))).c_str());
}
} // namespace

int stst_stream_write_value32(stst_stream_t stream, void *device_ptr, uint32_t value) {
    if (int rc = load_stream_value_ops())
        return rc;
    CUresult res = g_write_value32(static_cast<CUstream>(stream),
                                   reinterpret_cast<CUdeviceptr>(device_ptr), value,
                                   CU_STREAM_WRITE_VALUE_DEFAULT);
    return res == CUDA_SUCCESS ? 0 : driver_fail("cuStreamWriteValue32", res);
}

int stst_stream_wait_value32_geq(stst_stream_t stream, void *device_ptr, uint32_t value) {
    if (int rc = load_stream_value_ops())
        return rc;
    CUresult res = g_wait_value32(static_cast<CUstream>(stream),
                                  reinterpret_cast<CUdeviceptr>(device_ptr), value,
                                  CU_STREAM_WAIT_VALUE_GEQ);
    return res == CUDA_SUCCESS ? 0 : driver_fail("cuStreamWaitValue32", res);
}

int stst_tensor_map_encode_2d(void *tensor_map_out, const void *base, int elem_bytes,
                              uint64_t width, uint64_t height, uint64_t pitch_bytes,
                              uint32_t box_w, uint32_t box_h) {
    using EncodeFn = CUresult (*)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                  const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn encode = nullptr;
    if (!encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        STST_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (!fn || qres != cudaDriverEntryPointSuccess)
            return fail(-1, "stst_tensor_map_encode_2d", "cuTensorMapEncodeTiled unavailable");
        encode = reinterpret_cast<EncodeFn>(fn);
    }
    CUtensorMapDataType dtype;
    switch (elem_bytes) {
    case 1: dtype = CU_TENSOR_MAP_DATA_TYPE_UINT8; break;
    case 2: dtype = CU_TENSOR_MAP_DATA_TYPE_UINT16; break;
    case 4: dtype = CU_TENSOR_MAP_DATA_TYPE_UINT32; break;
    case 8: dtype = CU_TENSOR_MAP_DATA_TYPE_UINT64; break;
    default: return fail(-1, "stst_tensor_map_encode_2d", "element size must be 1, 2, 4 or 8");
    }
    if (pitch_bytes % 16 != 0 || (reinterpret_cast<uintptr_t>(base) % 16) != 0)
        return fail(-1, "stst_tensor_map_encode_2d", "base and pitch must be 16-byte aligned");
    if (box_w == 0 || box_h == 0 || box_w > 256 || box_h > 256 ||
        (uint64_t(box_w) * uint64_t(elem_bytes)) % 16 != 0)
        return fail(-1, "stst_tensor_map_encode_2d", "illegal box");
    cuuint64_t dims[2] = {width, height};
    cuuint64_t strides[1] = {pitch_bytes};
    cuuint32_t box[2] = {box_w, box_h};
    cuuint32_t elem_strides[2] = {1, 1};
    CUresult res = encode(static_cast<CUtensorMap *>(tensor_map_out), dtype, 2,
                          const_cast<void *>(base), dims, strides, box, elem_strides,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (res != CUDA_SUCCESS) {
        char buf[64];
        std::snprintf(buf, sizeof(buf), "CUresult %d", int(res));
        return fail(int(res), "cuTensorMapEncodeTiled", buf);
    }
    return 0;
}

int stst_peer_can_access(int device, int peer, int *can) {
    if (!can)
        return fail(-1, "stst_peer_can_access", "null argument");
    STST_CUDA(cudaDeviceCanAccessPeer(can, device, peer));
    return 0;
}

int stst_peer_enable(int device, int peer) {
    DeviceGuard guard;
    STST_CUDA(guard.enter(device));
    cudaError_t err = cudaDeviceEnablePeerAccess(peer, 0);
    if (err == cudaErrorPeerAccessAlreadyEnabled) {
        cudaGetLastError();
        return 0;
    }
    STST_CUDA(err);
    return 0;
}

int stst_ipc_get_mem_handle(void *ptr, unsigned char handle[STST_IPC_HANDLE_BYTES]) {
    static_assert(sizeof(cudaIpcMemHandle_t) == STST_IPC_HANDLE_BYTES);
    cudaIpcMemHandle_t h;
    STST_CUDA(cudaIpcGetMemHandle(&h, ptr));
    std::memcpy(handle, &h, sizeof(h));
    return 0;
}

int stst_ipc_open_mem_handle(const unsigned char handle[STST_IPC_HANDLE_BYTES], void **ptr) {
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, sizeof(h));
    STST_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}

int stst_ipc_close_mem_handle(void *ptr) {
    STST_CUDA(cudaIpcCloseMemHandle(ptr));
    return 0;
}

int stst_ipc_get_event_handle(stst_event_t event, unsigned char handle[STST_IPC_HANDLE_BYTES]) {
    static_assert(sizeof(cudaIpcEventHandle_t) == STST_IPC_HANDLE_BYTES);
    cudaIpcEventHandle_t h;
    STST_CUDA(cudaIpcGetEventHandle(&h, as_event(event)));
    std::memcpy(handle, &h, sizeof(h));
    return 0;
}

int stst_ipc_open_event_handle(const unsigned char handle[STST_IPC_HANDLE_BYTES],
                               stst_event_t *event) {
    cudaIpcEventHandle_t h;
    std::memcpy(&h, handle, sizeof(h));
    cudaEvent_t e;
    STST_CUDA(cudaIpcOpenEventHandle(&e, h));
    *event = e;
    return 0;
}

int stst_nccl_available(void) { return load_nccl() ? 1 : 0; }

int stst_nccl_get_unique_id(unsigned char id[STST_NCCL_UNIQUE_ID_BYTES]) {
    if (!load_nccl())
        return fail(-1, "stst_nccl_get_unique_id", "libnccl.so.2 not found");
    STST_NCCL(g_nccl.GetUniqueId(id));
    return 0;
}

int stst_nccl_comm_init_rank(stst_nccl_comm_t *comm, int n_ranks,
                             const unsigned char id[STST_NCCL_UNIQUE_ID_BYTES], int rank) {
    if (!load_nccl())
        return fail(-1, "stst_nccl_comm_init_rank", "libnccl.so.2 not found");
    UidByValue uid;
    std::memcpy(uid.internal, id, sizeof(uid.internal));
    void *c = nullptr;
    STST_NCCL(g_nccl.CommInitRank(&c, n_ranks, uid, rank));
    *comm = c;
    return 0;
}

int stst_nccl_comm_destroy(stst_nccl_comm_t comm) {
    if (!comm || !g_nccl.lib)
        return 0;
    STST_NCCL(g_nccl.CommDestroy(comm));
    return 0;
}

int stst_nccl_neighbor_exchange(stst_nccl_comm_t comm, int n, const int *peer,
                                const void *const *send_buf, const size_t *send_bytes,
                                void *const *recv_buf, const size_t *recv_bytes,
                                stst_stream_t stream) {
    if (!g_nccl.lib)
        return fail(-1, "stst_nccl_neighbor_exchange", "NCCL not initialised");
    STST_NCCL(g_nccl.GroupStart());
    for (int i = 0; i < n; i++) {
        if (send_bytes[i] > 0)
            STST_NCCL(g_nccl.Send(send_buf[i], send_bytes[i], kNcclChar, peer[i], comm,
                                  as_stream(stream)));
        if (recv_bytes[i] > 0)
            STST_NCCL(g_nccl.Recv(recv_buf[i], recv_bytes[i], kNcclChar, peer[i], comm,
                                  as_stream(stream)));
    }
    STST_NCCL(g_nccl.GroupEnd());
    return 0;
}

} // extern "C"
