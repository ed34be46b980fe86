// StencilStream-B200 SYCL stand-in: legacy `#include <CL/sycl.hpp>` spelling.
#pragma once
#include "../sycl/sycl.hpp"
