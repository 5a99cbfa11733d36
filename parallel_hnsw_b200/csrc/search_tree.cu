// search_tree.cu -- traversal kernel K1, variant: tree-order distances (PHNSW_SUM_TREE).
#include "search_launch.cuh"

namespace phnsw {
cudaError_t launch_search_tree(int metric, const SearchArgs &a, int grid, int block, size_t smem,
                             cudaStream_t stream) {
  return launch_metric<0, 1>(metric, a, grid, block, smem, stream);
}
}  // namespace phnsw
