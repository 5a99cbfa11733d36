// search_seq.cu -- traversal kernel K1, variant: sequential-order distances, rows staged by bulk copies.
#include "search_launch.cuh"

namespace phnsw {
cudaError_t launch_search_seq(int metric, const SearchArgs &a, int grid, int block, size_t smem,
                             cudaStream_t stream) {
  return launch_metric<0, 0>(metric, a, grid, block, smem, stream);
}
}  // namespace phnsw
