// search_pq.cu -- traversal kernel K1, variant: ADC over u8 codes.
#include "search_launch.cuh"

namespace phnsw {
cudaError_t launch_search_pq(int metric, const SearchArgs &a, int grid, int block, size_t smem,
                             cudaStream_t stream) {
  return launch_metric<1, 0>(metric, a, grid, block, smem, stream);
}
}  // namespace phnsw
