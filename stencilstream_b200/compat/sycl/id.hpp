// StencilStream-B200 SYCL stand-in (reference include: StencilStream/Stencil.hpp:23).
#pragma once
#include "sycl.hpp"
