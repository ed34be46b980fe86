/*
 * StencilStream-B200 stand-in for Intel's `ac_int<W, Signed>` arbitrary-width integer
 * (reference use: examples/fdtd/src/defines.hpp:26,46 — a small unsigned ring index and a loop
 * counter in examples/fdtd/src/material/RenderResolver.hpp:63). Values are held in the smallest
 * standard integer that fits and are wrapped to W bits on every store, which is all those uses need.
 */
#pragma once
#include <cstdint>
#include <type_traits>

#if defined(__CUDACC__)
    #define STST_AC_HD __host__ __device__
#else
    #define STST_AC_HD
#endif

template <int W, bool Signed = true> class ac_int {
    static_assert(W >= 1 && W <= 64);
    using storage_t = std::conditional_t<
        Signed,
        std::conditional_t<(W <= 8), std::int8_t,
                           std::conditional_t<(W <= 16), std::int16_t,
                                              std::conditional_t<(W <= 32), std::int32_t, std::int64_t>>>,
        std::conditional_t<(W <= 8), std::uint8_t,
                           std::conditional_t<(W <= 16), std::uint16_t,
                                              std::conditional_t<(W <= 32), std::uint32_t, std::uint64_t>>>>;

  public:
    STST_AC_HD constexpr ac_int() : value(0) {}
    template <typename I>
        requires std::is_arithmetic_v<I>
    STST_AC_HD constexpr ac_int(I v) : value(wrap(static_cast<std::int64_t>(v))) {}

    STST_AC_HD constexpr operator storage_t() const { return value; }

    STST_AC_HD constexpr ac_int &operator++() {
        value = wrap(static_cast<std::int64_t>(value) + 1);
        return *this;
    }
    STST_AC_HD constexpr ac_int operator++(int) {
        ac_int old = *this;
        ++(*this);
        return old;
    }
    STST_AC_HD constexpr ac_int &operator--() {
        value = wrap(static_cast<std::int64_t>(value) - 1);
        return *this;
    }
    template <typename I> STST_AC_HD constexpr ac_int &operator+=(I v) {
        value = wrap(static_cast<std::int64_t>(value) + static_cast<std::int64_t>(v));
        return *this;
    }
    template <typename I> STST_AC_HD constexpr ac_int &operator-=(I v) {
        value = wrap(static_cast<std::int64_t>(value) - static_cast<std::int64_t>(v));
        return *this;
    }

  private:
    STST_AC_HD static constexpr storage_t wrap(std::int64_t v) {
        if constexpr (W == 64) {
            return static_cast<storage_t>(v);
        } else {
            std::uint64_t mask = (std::uint64_t(1) << W) - 1;
            std::uint64_t u = static_cast<std::uint64_t>(v) & mask;
            if constexpr (Signed) {
                if (u & (std::uint64_t(1) << (W - 1)))
                    u |= ~mask;
            }
            return static_cast<storage_t>(u);
        }
    }

    storage_t value;
};
