/*
 * StencilStream-B200 — `accessor[i][j]...` chaining helper for accessors that natively only
 * understand `accessor[sycl::id<N>]`.
 *
 * Same class template name and template parameters as the reference
 * (StencilStream/AccessorSubscript.hpp:46-141). The B200 grid accessor does not need it (it hands
 * out row views directly, see cuda/Grid.hpp) but user code written against several backends may
 * name it.
 */
#pragma once
#include <sycl/sycl.hpp>

namespace stencil {

template <typename Cell, typename Accessor, sycl::access::mode access_mode,
          std::size_t current_subdim = 0>
class AccessorSubscript {
  public:
    static constexpr std::size_t dimensions = Accessor::dimensions;
    static constexpr bool is_last = (current_subdim + 2 == dimensions);
    using id_t = sycl::id<int(dimensions)>;
    using cell_ref =
        std::conditional_t<access_mode == sycl::access::mode::read, Cell const &, Cell &>;

    /// Start a subscript chain with the index of the outermost dimension.
    AccessorSubscript(Accessor &ac, std::size_t i)
        requires(current_subdim == 0)
        : ac(ac), prefix() {
        prefix[int(current_subdim)] = i;
    }

    /// Continue a subscript chain.
    AccessorSubscript(Accessor &ac, id_t prefix, std::size_t i) : ac(ac), prefix(prefix) {
        this->prefix[int(current_subdim)] = i;
    }

    AccessorSubscript<Cell, Accessor, access_mode, current_subdim + 1> operator[](std::size_t i)
        requires(!is_last)
    {
        return AccessorSubscript<Cell, Accessor, access_mode, current_subdim + 1>(ac, prefix, i);
    }

    cell_ref operator[](std::size_t i)
        requires(is_last)
    {
        id_t full = prefix;
        full[int(current_subdim) + 1] = i;
        return ac[full];
    }

  private:
    Accessor &ac;
    id_t prefix;
};

} // namespace stencil
