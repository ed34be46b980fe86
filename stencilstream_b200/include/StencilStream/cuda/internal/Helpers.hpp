/*
 * StencilStream-B200 — compile-time description of how a `Cell` is laid out in HBM and shared memory.
 *
 * Counterpart of the reference's SoA plumbing (StencilStream/cuda/internal/Helpers.hpp:37-67:
 * `alloc_field_buffers`, `FieldBuffers`, `for_each_in_two_tuples`), which builds one 1-D buffer per
 * entry of the opt-in `Cell::fields` tuple of member pointers. Here the same hook decides the
 * *native* device layout of `cuda::Grid<Cell>`:
 *
 *   - `Cell::fields` present and covering every byte of the cell  -> one row-major plane per field
 *     (struct-of-arrays; coalesced 128-bit accesses per field),
 *   - otherwise (scalars such as `float`/`bool`, or opaque structs)  -> a single plane of whole cells.
 *
 * Nothing in here touches the device; it is index/tuple arithmetic shared by host and device code.
 */
#pragma once
#include <sycl/sycl.hpp>

#include <cstddef>
#include <tuple>
#include <type_traits>
#include <utility>

namespace stencil {
namespace cuda {
namespace internal {

/// Upper bound on planes per cell (sizes the POD kernel-argument structs).
inline constexpr std::size_t max_planes = 16;

template <typename Cell>
concept HasFieldList = requires {
    Cell::fields;
    requires(std::tuple_size_v<std::remove_cvref_t<decltype(Cell::fields)>> >= 1);
};

namespace detail {
template <typename Cell, typename MemberPtr>
using member_type_t =
    std::remove_cvref_t<decltype(std::declval<Cell &>().*std::declval<MemberPtr>())>;

template <typename Cell> constexpr std::size_t listed_field_bytes() {
    if constexpr (HasFieldList<Cell>) {
        return std::apply(
            [](auto... ptrs) { return (std::size_t(0) + ... + sizeof(member_type_t<Cell, decltype(ptrs)>)); },
            Cell::fields);
    } else {
        return 0;
    }
}

template <typename Cell> constexpr std::size_t listed_field_count() {
    if constexpr (HasFieldList<Cell>) {
        return std::tuple_size_v<std::remove_cvref_t<decltype(Cell::fields)>>;
    } else {
        return 0;
    }
}
} // namespace detail

/**
 * Layout traits of a cell type. `n_planes` planes, plane `I` holding elements of `plane_t<I>`;
 * `get<I>(cell)` is the lvalue inside an (AoS, register- or host-resident) cell that plane `I` stores.
 */
template <typename Cell> struct CellLayout {
    static_assert(std::is_trivially_copyable_v<Cell>,
                  "StencilStream-B200 cells must be trivially copyable");

    /// True if the cell is stored as one plane per `Cell::fields` entry.
    static constexpr bool is_split =
        HasFieldList<Cell> && detail::listed_field_bytes<Cell>() == sizeof(Cell) &&
        detail::listed_field_count<Cell>() <= max_planes;

    static constexpr std::size_t n_planes = is_split ? detail::listed_field_count<Cell>() : 1;

    template <std::size_t I> struct plane {
        static_assert(I < n_planes);
        static constexpr auto pick() {
            if constexpr (is_split) {
                using P = std::tuple_element_t<I, std::remove_cvref_t<decltype(Cell::fields)>>;
                return std::type_identity<detail::member_type_t<Cell, P>>{};
            } else {
                return std::type_identity<Cell>{};
            }
        }
        using type = typename decltype(pick())::type;
    };
    template <std::size_t I> using plane_t = typename plane<I>::type;

    template <std::size_t I> STST_HD static constexpr plane_t<I> &get(Cell &cell) {
        if constexpr (is_split) {
            constexpr auto member = std::get<I>(Cell::fields);
            return cell.*member;
        } else {
            return cell;
        }
    }

    template <std::size_t I> STST_HD static constexpr plane_t<I> const &get(Cell const &cell) {
        if constexpr (is_split) {
            constexpr auto member = std::get<I>(Cell::fields);
            return cell.*member;
        } else {
            return cell;
        }
    }

    STST_HD static constexpr std::size_t plane_bytes(std::size_t i) {
        std::size_t sizes[n_planes] = {};
        fill_sizes(sizes, std::make_index_sequence<n_planes>{});
        return sizes[i];
    }

    STST_HD static constexpr std::size_t max_plane_bytes() {
        std::size_t m = 0;
        for (std::size_t i = 0; i < n_planes; i++)
            m = plane_bytes(i) > m ? plane_bytes(i) : m;
        return m;
    }

  private:
    template <std::size_t... Is>
    STST_HD static constexpr void fill_sizes(std::size_t *sizes, std::index_sequence<Is...>) {
        ((sizes[Is] = sizeof(plane_t<Is>)), ...);
    }
};

/**
 * Opt-in declaration of fields a cell's transition functions never change (B200 extension; the
 * reference API has nothing like it — SURVEY.md H4):
 *
 *     struct HotspotCell { float temp, power;
 *         static constexpr auto fields = std::make_tuple(&HotspotCell::temp, &HotspotCell::power);
 *         static constexpr auto constant_fields = std::make_tuple(&HotspotCell::power); };
 *
 * The tile sweeps then leave those planes in place instead of copying them from tile buffer to tile
 * buffer, and keep them in ONE tile buffer (taller tiles) — the same kernels that the run-time
 * detection of such fields uses (speculative plane pass-through, StencilUpdate.hpp), but decided at
 * compile time: no observing launch, no verification read-back, so `blocking = false` calls stay
 * asynchronous. Every entry must also be listed in `Cell::fields`. It is a contract: a transition
 * function that does change a declared field gets the old value back (the kernels still detect it;
 * set STST_VERIFY_CONSTANT_FIELDS=1 to have updates check and throw).
 */
template <typename Cell>
concept HasConstantFieldList = HasFieldList<Cell> && requires {
    Cell::constant_fields;
    requires(std::tuple_size_v<std::remove_cvref_t<decltype(Cell::constant_fields)>> >= 1);
};

namespace detail {
template <typename A, typename B> constexpr bool same_member(A a, B b) {
    if constexpr (std::is_same_v<A, B>)
        return a == b;
    else
        return false;
}

template <typename Cell, typename MemberPtr> constexpr unsigned plane_bit_of(MemberPtr member) {
    unsigned bit = 0, index = 0;
    std::apply([&](auto... listed) { ((bit |= same_member(listed, member) ? (1u << index) : 0u, index++), ...); },
               Cell::fields);
    return bit;
}
} // namespace detail

/// Bit i set: plane i holds a field listed in `Cell::constant_fields` (0 without such a list).
template <typename Cell> constexpr unsigned constant_fields_mask() {
    if constexpr (HasConstantFieldList<Cell> && CellLayout<Cell>::is_split) {
        unsigned mask = 0;
        bool all_listed = true;
        std::apply(
            [&](auto... constant) {
                ((mask |= detail::plane_bit_of<Cell>(constant),
                  all_listed = all_listed && detail::plane_bit_of<Cell>(constant) != 0),
                 ...);
            },
            Cell::constant_fields);
        return all_listed ? mask : ~0u; // ~0u: an entry that is not in Cell::fields (static_assert'ed)
    } else {
        return 0;
    }
}

/// Invoke `f(std::integral_constant<size_t, I>{})` for every plane index of `Cell`.
#if defined(__CUDACC__)
    #pragma nv_exec_check_disable
#endif
template <typename Cell, typename Fn> STST_HD constexpr void for_each_plane(Fn &&f) {
    [&]<std::size_t... Is>(std::index_sequence<Is...>) {
        (f(std::integral_constant<std::size_t, Is>{}), ...);
    }(std::make_index_sequence<CellLayout<Cell>::n_planes>{});
}

/**
 * POD description of the device planes of one grid slab, passed by value into kernels.
 * Plane `i` is row-major with `pitch[i]` *elements* between consecutive rows (the byte pitch is a
 * multiple of 128 so that every row start is 128-bit aligned and TMA-addressable).
 */
struct PlaneSet {
    void *base[max_planes];
    unsigned long long pitch[max_planes];
};

} // namespace internal
} // namespace cuda
} // namespace stencil
