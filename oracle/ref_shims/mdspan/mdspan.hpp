// TEST INFRASTRUCTURE (oracle/_ref build glue) -- not part of the product.
// Stand-in for Kokkos' <mdspan/mdspan.hpp> (pinned by the reference in
// external/macis/cmake/macis-mdspan.cmake, not vendored): forwards to the
// CCCL implementation that ships with CUDA 12.9. Only layout_left views are
// used on the CI hot path (external/macis/include/macis/types.hpp:20,115-150).
#pragma once
#include <cuda/std/mdspan>
namespace Kokkos {
using ::cuda::std::mdspan;
using ::cuda::std::extents;
using ::cuda::std::dextents;
using ::cuda::std::layout_left;
using ::cuda::std::layout_right;
using ::cuda::std::layout_stride;
using ::cuda::std::full_extent;
using ::cuda::std::full_extent_t;
using ::cuda::std::dynamic_extent;
using ::cuda::std::default_accessor;
using ::cuda::std::submdspan;
}  // namespace Kokkos
namespace KokkosEx {
using ::cuda::std::submdspan;
}
