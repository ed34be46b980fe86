/*
 * stst_rt.h — C ABI of the StencilStream-B200 device runtime (libstst_rt.so).
 *
 * This is the thin layer the header-only C++ backend (stencilstream_b200/include/StencilStream/cuda)
 * sits on: device memory, pinned host memory, streams, events, copies, peer/IPC mappings, TMA
 * descriptors and (optionally) NCCL point-to-point exchange. It replaces what the reference's cuda
 * backend obtains from the SYCL runtime:
 *
 *   stst_malloc* / stst_free*         <- sycl::buffer<Cell,2> construction/destruction
 *                                        (reference StencilStream/cuda/Grid.hpp:66-74,184-187;
 *                                         cuda/internal/Helpers.hpp:37-45 alloc_field_buffers)
 *   stst_memcpy_*                     <- sycl::host_accessor migration, copy_from/to_buffer
 *                                        (reference StencilStream/cuda/Grid.hpp:109-134,145-153)
 *   stst_stream_* / stst_event_*      <- sycl::queue, queue.wait(), sycl::event profiling
 *                                        (reference StencilStream/cuda/StencilUpdate.hpp:124-135,184-198)
 *   stst_peer_* / stst_ipc_* / stst_nccl_*
 *                                     <- no counterpart: the reference cuda backend is single-GPU
 *                                        (StencilStream/cuda/StencilUpdate.hpp:83); used by the new
 *                                        row-sharding partitioner.
 *   stst_tensor_map_encode_2d         <- no counterpart (TMA descriptors for shared-memory tile loads).
 *
 * Conventions: every function returns 0 on success and a non-zero CUDA/NCCL-derived code on
 * failure; stst_last_error() then returns a thread-local, human-readable message. All handles are
 * opaque pointers. Buffers are caller-owned. The library is thread-compatible, not thread-safe.
 * There is no CPU fallback: without a CUDA device every call that needs one fails.
 */
#ifndef STST_RT_H
#define STST_RT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define STST_RT_ABI_VERSION 1

typedef void *stst_stream_t; /* a CUstream / cudaStream_t */
typedef void *stst_event_t;  /* a CUevent / cudaEvent_t   */

typedef struct stst_device_info {
    int sm_count;
    int cc_major;
    int cc_minor;
    int max_smem_per_block_optin; /* bytes */
    int max_smem_per_sm;          /* bytes */
    int l2_bytes;
    int clock_khz;
    size_t total_mem;
    size_t free_mem;
    char name[128];
} stst_device_info;

/* --- introspection ----------------------------------------------------------------------- */
int stst_rt_abi_version(void);
const char *stst_last_error(void);
int stst_device_count(int *count);
int stst_get_device_info(int device, stst_device_info *info);
int stst_set_device(int device);

/* --- memory ------------------------------------------------------------------------------ */
/* Stream-ordered pool allocation (cudaMallocAsync) on `device`; 256-byte aligned. */
int stst_malloc(int device, size_t bytes, stst_stream_t stream, void **ptr);
int stst_free(int device, void *ptr, stst_stream_t stream);
/* Classic, IPC-exportable allocation (cudaMalloc). */
int stst_malloc_ipc(int device, size_t bytes, void **ptr);
int stst_free_ipc(int device, void *ptr);
/* Pinned, portable host memory. Freed blocks are cached by size (STST_PINNED_CACHE_MB, default
 * 16384) because pinning gigabytes costs hundreds of milliseconds; stst_host_cache_trim releases them.
 * Requests above STST_PIN_LIMIT_MB (if set) fail, like requests the box cannot pin. */
int stst_malloc_host(size_t bytes, void **ptr);
int stst_free_host(void *ptr);
int stst_host_cache_trim(void);
int stst_host_register(void *ptr, size_t bytes); /* pin caller-owned memory */
int stst_host_unregister(void *ptr);
int stst_memset_async(void *ptr, int value, size_t bytes, stst_stream_t stream);

/* --- copies (all asynchronous w.r.t. the host when the host side is pinned) ---------------- */
int stst_memcpy_h2d_async(void *dst, const void *src, size_t bytes, stst_stream_t stream);
int stst_memcpy_d2h_async(void *dst, const void *src, size_t bytes, stst_stream_t stream);
int stst_memcpy_d2d_async(void *dst, const void *src, size_t bytes, stst_stream_t stream);
/* Pitched variants: `height` rows of `width_bytes`; kind: 0 = h2d, 1 = d2h, 2 = d2d. */
int stst_memcpy_2d_async(void *dst, size_t dst_pitch, const void *src, size_t src_pitch,
                         size_t width_bytes, size_t height, int kind, stst_stream_t stream);
int stst_memcpy_peer_async(void *dst, int dst_device, const void *src, int src_device,
                           size_t bytes, stst_stream_t stream);
/*
 * Host memory that is not pinned (ordinary malloc / numpy / std::vector memory, or a grid image the
 * box refused to pin): `rows` rows of `row_bytes` between host rows `host_pitch` apart and device
 * rows `dev_pitch` apart, staged through a ring of three pinned slots (STST_STAGING_SLOT_MB, default
 * 16) with the host-side copies spread over worker threads (STST_COPY_THREADS, default min(8,
 * cores/2)) and overlapped with the DMA. kind: 0 = h2d, 1 = d2h. h2d returns once `host` may be
 * reused (the last DMAs may still be in flight on `stream`); d2h returns when `host` holds the data.
 * stst_memcpy_2d_auto picks the direct asynchronous copy when `host` is pinned, else the staged one.
 */
int stst_memcpy_2d_staged(void *dev, size_t dev_pitch, void *host, size_t host_pitch,
                          size_t row_bytes, size_t rows, int kind, int device, stst_stream_t stream);
int stst_memcpy_2d_auto(void *dev, size_t dev_pitch, void *host, size_t host_pitch, size_t row_bytes,
                        size_t rows, int kind, int device, stst_stream_t stream);
int stst_host_memcpy(void *dst, const void *src, size_t bytes); /* multi-threaded memcpy */
int stst_host_is_pinned(const void *ptr, int *pinned);

/* --- streams and events -------------------------------------------------------------------- */
/* The per-device stream every StencilStream-B200 object uses unless told otherwise. */
int stst_default_stream(int device, stst_stream_t *stream);
int stst_stream_create(int device, int high_priority, stst_stream_t *stream);
int stst_stream_destroy(stst_stream_t stream);
int stst_stream_synchronize(stst_stream_t stream);
int stst_stream_wait_event(stst_stream_t stream, stst_event_t event);
int stst_event_create(int with_timing, stst_event_t *event);
int stst_event_destroy(stst_event_t event);
int stst_event_record(stst_event_t event, stst_stream_t stream);
int stst_event_synchronize(stst_event_t event);
int stst_event_elapsed_ms(stst_event_t start, stst_event_t stop, float *ms);
int stst_device_synchronize(int device);
/*
 * Stream-ordered 32-bit flags in device memory (cuStreamWriteValue32 / cuStreamWaitValue32 with
 * CU_STREAM_WAIT_VALUE_GEQ). `device_ptr` may be a peer- or IPC-mapped address for writes; waits
 * poll memory of the stream's own device. Used by the slab partitioner to order halo pushes between
 * GPUs (and between processes) without host synchronisation.
 */
int stst_stream_write_value32(stst_stream_t stream, void *device_ptr, uint32_t value);
int stst_stream_wait_value32_geq(stst_stream_t stream, void *device_ptr, uint32_t value);

/* --- TMA ------------------------------------------------------------------------------------- */
/*
 * Encode a 2-D tiled tensor map (CUtensorMap, 128 bytes, 64-byte aligned) over a row-major plane
 * of `height` x `width` elements of `elem_bytes` (1, 2, 4 or 8) whose rows are `pitch_bytes` apart
 * (multiple of 16). The box is `box_w` x `box_h` elements (each <= 256, box_w*elem_bytes % 16 == 0);
 * out-of-bounds elements are zero-filled, no swizzle, no interleave.
 */
int stst_tensor_map_encode_2d(void *tensor_map_out, const void *base, int elem_bytes,
                              uint64_t width, uint64_t height, uint64_t pitch_bytes,
                              uint32_t box_w, uint32_t box_h);

/* --- multi-GPU: peers in one process, IPC across processes, NCCL p2p ------------------------ */
int stst_peer_can_access(int device, int peer, int *can);
int stst_peer_enable(int device, int peer);
#define STST_IPC_HANDLE_BYTES 64
int stst_ipc_get_mem_handle(void *ptr, unsigned char handle[STST_IPC_HANDLE_BYTES]);
int stst_ipc_open_mem_handle(const unsigned char handle[STST_IPC_HANDLE_BYTES], void **ptr);
int stst_ipc_close_mem_handle(void *ptr);
int stst_ipc_get_event_handle(stst_event_t event, unsigned char handle[STST_IPC_HANDLE_BYTES]);
int stst_ipc_open_event_handle(const unsigned char handle[STST_IPC_HANDLE_BYTES],
                               stst_event_t *event);
int stst_event_create_ipc(stst_event_t *event); /* interprocess-capable, timing disabled */

#define STST_NCCL_UNIQUE_ID_BYTES 128
typedef void *stst_nccl_comm_t;
int stst_nccl_available(void);
int stst_nccl_get_unique_id(unsigned char id[STST_NCCL_UNIQUE_ID_BYTES]);
int stst_nccl_comm_init_rank(stst_nccl_comm_t *comm, int n_ranks,
                             const unsigned char id[STST_NCCL_UNIQUE_ID_BYTES], int rank);
int stst_nccl_comm_destroy(stst_nccl_comm_t comm);
/*
 * One grouped neighbour exchange: send `send_bytes[i]` from `send_buf[i]` to `peer[i]` and receive
 * `recv_bytes[i]` into `recv_buf[i]` from the same peer, for i in [0, n). Zero-byte legs are skipped.
 */
int stst_nccl_neighbor_exchange(stst_nccl_comm_t comm, int n, const int *peer,
                                const void *const *send_buf, const size_t *send_bytes,
                                void *const *recv_buf, const size_t *recv_bytes,
                                stst_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* STST_RT_H */
