// StencilStream-B200 SYCL stand-in: `#include <sycl.hpp>` spelling (reference: StencilStream/Concepts.hpp:25).
#pragma once
#include "sycl/sycl.hpp"
