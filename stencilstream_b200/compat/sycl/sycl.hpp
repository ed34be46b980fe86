/*
 * StencilStream-B200 — minimal header-only stand-in for the SYCL 2020 surface that
 * StencilStream user code touches.
 *
 * This is NOT a SYCL implementation. It provides exactly the vocabulary types the reference's
 * user-facing API is written in (reference: StencilStream/Stencil.hpp:62-65 uses sycl::id/range;
 * StencilStream/cuda/Grid.hpp:66,145 uses sycl::buffer / sycl::host_accessor; the example mains use
 * sycl::access::mode, sycl::device, sycl::exception_list, sycl::cos/exp/isinf) so that those sources
 * compile under nvcc (device side: id/range only) and under g++ (host side: everything, including a
 * sequential/OpenMP `queue::submit` + `handler::parallel_for` that is only used to build the
 * reference's cpu backend as the parity oracle).
 */
#pragma once

#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <fstream>
#include <functional>
#include <iostream>
#include <limits>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <tuple>
#include <type_traits>
#include <utility>
#include <variant>
#include <vector>

#if defined(__CUDACC__)
    #define STST_HD __host__ __device__
    #define STST_FORCEINLINE __forceinline__
#else
    #define STST_HD
    #define STST_FORCEINLINE inline __attribute__((always_inline))
#endif

namespace sycl {

// ------------------------------------------------------------------------------------------------
// id / range
// ------------------------------------------------------------------------------------------------

namespace detail {
template <int N> struct index_array {
    std::size_t v[N];

    STST_HD constexpr index_array() : v{} {}

    STST_HD constexpr std::size_t &operator[](int i) { return v[i]; }
    STST_HD constexpr std::size_t const &operator[](int i) const { return v[i]; }

    STST_HD constexpr bool equals(index_array const &o) const {
        bool eq = true;
        for (int i = 0; i < N; i++)
            eq = eq && (v[i] == o.v[i]);
        return eq;
    }
};
} // namespace detail

template <int N = 1> class range : public detail::index_array<N> {
    static_assert(N >= 1 && N <= 3);

  public:
    constexpr range() = default;
    STST_HD constexpr range(std::size_t d0)
        requires(N == 1)
    {
        this->v[0] = d0;
    }
    STST_HD constexpr range(std::size_t d0, std::size_t d1)
        requires(N == 2)
    {
        this->v[0] = d0;
        this->v[1] = d1;
    }
    STST_HD constexpr range(std::size_t d0, std::size_t d1, std::size_t d2)
        requires(N == 3)
    {
        this->v[0] = d0;
        this->v[1] = d1;
        this->v[2] = d2;
    }

    STST_HD constexpr std::size_t size() const {
        std::size_t s = 1;
        for (int i = 0; i < N; i++)
            s *= this->v[i];
        return s;
    }
    STST_HD constexpr std::size_t get(int i) const { return this->v[i]; }

    STST_HD friend constexpr bool operator==(range const &a, range const &b) { return a.equals(b); }
    STST_HD friend constexpr bool operator!=(range const &a, range const &b) {
        return !a.equals(b);
    }
};

template <int N = 1> class id : public detail::index_array<N> {
    static_assert(N >= 1 && N <= 3);

  public:
    constexpr id() = default;
    STST_HD constexpr id(std::size_t d0)
        requires(N == 1)
    {
        this->v[0] = d0;
    }
    STST_HD constexpr id(std::size_t d0, std::size_t d1)
        requires(N == 2)
    {
        this->v[0] = d0;
        this->v[1] = d1;
    }
    STST_HD constexpr id(std::size_t d0, std::size_t d1, std::size_t d2)
        requires(N == 3)
    {
        this->v[0] = d0;
        this->v[1] = d1;
        this->v[2] = d2;
    }
    STST_HD constexpr id(range<N> const &r) {
        for (int i = 0; i < N; i++)
            this->v[i] = r[i];
    }

    STST_HD constexpr std::size_t get(int i) const { return this->v[i]; }
    STST_HD constexpr operator std::size_t() const
        requires(N == 1)
    {
        return this->v[0];
    }

    STST_HD friend constexpr bool operator==(id const &a, id const &b) { return a.equals(b); }
    STST_HD friend constexpr bool operator!=(id const &a, id const &b) { return !a.equals(b); }
};

id(std::size_t) -> id<1>;
id(std::size_t, std::size_t) -> id<2>;
id(std::size_t, std::size_t, std::size_t) -> id<3>;
range(std::size_t) -> range<1>;
range(std::size_t, std::size_t) -> range<2>;
range(std::size_t, std::size_t, std::size_t) -> range<3>;

// ------------------------------------------------------------------------------------------------
// access tags, device, event, exceptions, properties
// ------------------------------------------------------------------------------------------------

namespace access {
enum class mode { read, write, read_write, discard_write, discard_read_write, atomic };
enum class target { device, host_task, global_buffer, constant_buffer, local, host_buffer };
enum class address_space { global_space, local_space, constant_space, private_space };
enum class decorated { no, yes, legacy };
} // namespace access
using access_mode = access::mode;

template <access::mode m> struct mode_tag_t {
    explicit mode_tag_t() = default;
};
inline constexpr mode_tag_t<access::mode::read> read_only{};
inline constexpr mode_tag_t<access::mode::write> write_only{};
inline constexpr mode_tag_t<access::mode::read_write> read_write{};

struct no_init_t {
    explicit no_init_t() = default;
};
inline constexpr no_init_t no_init{};

class device {
  public:
    device() = default;
    bool is_cpu() const { return true; }
    bool is_gpu() const { return false; }
    friend bool operator==(device const &, device const &) { return true; }
};

struct default_selector {};
inline constexpr int default_selector_v = 0;
inline constexpr int cpu_selector_v = 1;
inline constexpr int gpu_selector_v = 2;

class exception : public std::exception {
  public:
    exception() = default;
    explicit exception(std::string msg) : msg(std::move(msg)) {}
    const char *what() const noexcept override { return msg.c_str(); }

  private:
    std::string msg;
};

using exception_list = std::vector<std::exception_ptr>;
using async_handler = std::function<void(exception_list)>;

namespace info {
namespace event_profiling {
struct command_submit {};
struct command_start {};
struct command_end {};
} // namespace event_profiling
} // namespace info

class event {
  public:
    event() = default;
    event(std::uint64_t start_ns, std::uint64_t end_ns) : start_ns(start_ns), end_ns(end_ns) {}
    void wait() {}
    template <typename Param> std::uint64_t get_profiling_info() const {
        if constexpr (std::is_same_v<Param, info::event_profiling::command_end>) {
            return end_ns;
        } else {
            return start_ns;
        }
    }

  private:
    std::uint64_t start_ns = 0, end_ns = 0;
};

namespace property {
namespace queue {
struct enable_profiling {};
struct in_order {};
} // namespace queue
struct no_init {};
} // namespace property

class property_list {
  public:
    property_list() = default;
    template <typename... Props> property_list(Props...) {}
};

// ------------------------------------------------------------------------------------------------
// buffer: shared, reference-counted host storage (SYCL buffers have reference semantics)
// ------------------------------------------------------------------------------------------------

template <typename T, int N = 1> class buffer {
  public:
    using value_type = T;
    static constexpr int dimensions = N;

    buffer(range<N> r) : r(r), storage(allocate(r.size())) {}
    buffer(T *host_data, range<N> r) : r(r), storage(allocate(r.size())) {
        std::memcpy(storage.get(), host_data, r.size() * sizeof(T));
    }

    range<N> get_range() const { return r; }
    std::size_t size() const { return r.size(); }
    std::size_t byte_size() const { return r.size() * sizeof(T); }

    /// Shim-only: raw storage; stable for the lifetime of any copy of this buffer.
    T *stst_data() const { return storage.get(); }

    friend bool operator==(buffer const &a, buffer const &b) { return a.storage == b.storage; }

  private:
    static std::shared_ptr<T[]> allocate(std::size_t n) {
        // Raw, 64-byte aligned, zero-initialised storage; deliberately never std::vector<bool>.
        std::size_t bytes = std::max<std::size_t>(n * sizeof(T), 1);
        bytes = (bytes + 63) / 64 * 64;
        void *p = std::aligned_alloc(64, bytes);
        if (p == nullptr)
            throw std::bad_alloc();
        std::memset(p, 0, bytes);
        return std::shared_ptr<T[]>(static_cast<T *>(p), [](T *q) { std::free(q); });
    }

    range<N> r;
    std::shared_ptr<T[]> storage;
};

// ------------------------------------------------------------------------------------------------
// accessors. Both flavours are plain views onto the buffer's host storage.
// ------------------------------------------------------------------------------------------------

namespace detail {
template <typename Ref> class row_view {
  public:
    using pointer = std::remove_reference_t<Ref> *;
    row_view(pointer row) : row(row) {}
    Ref operator[](std::size_t c) const { return row[c]; }

  private:
    pointer row;
};

template <typename T, int N, access::mode mode> class accessor_base {
  public:
    static constexpr int dimensions = N;
    static constexpr bool is_read_only = (mode == access::mode::read);
    using value_type = std::conditional_t<is_read_only, const T, T>;
    using reference = value_type &;

    accessor_base() : r(), data(nullptr) {}
    accessor_base(buffer<T, N> const &buf) : r(buf.get_range()), data(buf.stst_data()) {}

    range<N> get_range() const { return r; }
    std::size_t size() const { return r.size(); }
    std::size_t byte_size() const { return r.size() * sizeof(T); }
    value_type *get_pointer() const { return data; }

    reference operator[](id<N> i) const {
        std::size_t lin = 0;
        for (int d = 0; d < N; d++)
            lin = lin * r[d] + i[d];
        return data[lin];
    }

    auto operator[](std::size_t i) const {
        if constexpr (N == 1) {
            return static_cast<reference>(data[i]);
        } else {
            static_assert(N == 2, "only 1-D and 2-D accessors are supported by this shim");
            return row_view<reference>(data + i * r[1]);
        }
    }

  protected:
    range<N> r;
    value_type *data;
};
} // namespace detail

template <typename T, int N = 1, access::mode mode = access::mode::read_write>
class host_accessor : public detail::accessor_base<T, N, mode> {
    using base = detail::accessor_base<T, N, mode>;

  public:
    host_accessor() = default;
    host_accessor(buffer<T, N> const &buf) : base(buf) {}
    host_accessor(buffer<T, N> const &buf, mode_tag_t<mode>) : base(buf) {}
    host_accessor(buffer<T, N> const &buf, mode_tag_t<mode>, property_list const &) : base(buf) {}
};

template <typename T, int N> host_accessor(buffer<T, N>) -> host_accessor<T, N>;
template <typename T, int N, access::mode mode>
host_accessor(buffer<T, N>, mode_tag_t<mode>) -> host_accessor<T, N, mode>;

class handler;

template <typename T, int N = 1, access::mode mode = access::mode::read_write,
          access::target tgt = access::target::device>
class accessor : public detail::accessor_base<T, N, mode> {
    using base = detail::accessor_base<T, N, mode>;

  public:
    accessor() = default;
    accessor(buffer<T, N> const &buf) : base(buf) {}
    accessor(buffer<T, N> const &buf, handler &) : base(buf) {}
    accessor(buffer<T, N> const &buf, handler &, mode_tag_t<mode>) : base(buf) {}
    accessor(buffer<T, N> const &buf, handler &, mode_tag_t<mode>, property_list const &)
        : base(buf) {}
};

template <typename T, int N> accessor(buffer<T, N>, handler &) -> accessor<T, N>;
template <typename T, int N, access::mode mode>
accessor(buffer<T, N>, handler &, mode_tag_t<mode>) -> accessor<T, N, mode>;

template <typename T> class global_ptr {
  public:
    global_ptr(T *p = nullptr) : p(p) {}
    template <int N, access::mode mode, access::target tgt>
    global_ptr(accessor<std::remove_const_t<T>, N, mode, tgt> const &ac)
        : p(const_cast<T *>(ac.get_pointer())) {}
    T &operator[](std::size_t i) const { return p[i]; }
    T &operator*() const { return *p; }
    T *get() const { return p; }

  private:
    T *p;
};

// ------------------------------------------------------------------------------------------------
// queue / handler: synchronous execution on the calling host (OpenMP over rows when available).
// Only the oracle build of the reference's cpu backend uses this.
// ------------------------------------------------------------------------------------------------

class handler {
  public:
    template <typename KernelName = void, typename K> void parallel_for(range<1> r, K const &k) {
        for (std::size_t i = 0; i < r[0]; i++)
            k(id<1>(i));
    }

    template <typename KernelName = void, typename K> void parallel_for(range<2> r, K const &k) {
        const long long rows = static_cast<long long>(r[0]);
        const std::size_t cols = r[1];
#if defined(_OPENMP)
    #pragma omp parallel for schedule(static)
#endif
        for (long long row = 0; row < rows; row++) {
            for (std::size_t col = 0; col < cols; col++) {
                k(id<2>(static_cast<std::size_t>(row), col));
            }
        }
    }

    template <typename KernelName = void, typename K> void single_task(K const &k) { k(); }
};

class queue {
  public:
    queue() = default;
    queue(device const &) {}
    queue(device const &, property_list const &) {}
    queue(device const &, async_handler const &) {}
    queue(device const &, async_handler const &, property_list const &) {}
    queue(property_list const &) {}
    template <typename Selector>
        requires(std::is_integral_v<Selector>)
    queue(Selector) {}

    template <typename CGF> event submit(CGF const &cgf) {
        handler cgh;
        cgf(cgh);
        return event();
    }

    void wait() {}
    void wait_and_throw() {}
    device get_device() const { return device(); }
};

// ------------------------------------------------------------------------------------------------
// math used by example host code
// ------------------------------------------------------------------------------------------------

template <typename T> STST_HD inline T cos(T x) { return ::cos(x); }
STST_HD inline float cos(float x) { return ::cosf(x); }
template <typename T> STST_HD inline T sin(T x) { return ::sin(x); }
STST_HD inline float sin(float x) { return ::sinf(x); }
template <typename T> STST_HD inline T exp(T x) { return ::exp(x); }
STST_HD inline float exp(float x) { return ::expf(x); }
template <typename T> STST_HD inline T sqrt(T x) { return ::sqrt(x); }
STST_HD inline float sqrt(float x) { return ::sqrtf(x); }
template <typename T> STST_HD inline T fabs(T x) { return ::fabs(x); }
STST_HD inline float fabs(float x) { return ::fabsf(x); }
template <typename T> STST_HD inline bool isinf(T x) {
#if defined(__CUDA_ARCH__)
    return ::isinf(x);
#else
    return std::isinf(x);
#endif
}

} // namespace sycl
