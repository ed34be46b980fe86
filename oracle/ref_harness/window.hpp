// TEST INFRASTRUCTURE — oracle: where a crop of a larger grid lies (see oracle_ref::Shifted, harness.hpp).
#pragma once
#include <cstddef>

namespace oracle_ref {

/// The cells handed to the backend are rows [row0, row0 + rows) x columns [col0, col0 + cols) of a
/// `global_rows` x `global_cols` grid.
struct Window {
    std::size_t row0 = 0, col0 = 0, global_rows = 0, global_cols = 0;
    bool active() const { return global_rows != 0 || global_cols != 0; }
};

} // namespace oracle_ref
