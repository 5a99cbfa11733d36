// search_pq8q.cu -- traversal kernel K1, variant: ADC over u8 codes with quantised (u8)
// per-query tables in shared memory (tables written by adc_lut.cu).
#include "search_launch.cuh"

namespace phnsw {
cudaError_t launch_search_pq8q(int metric, const SearchArgs &a, int grid, int block, size_t smem,
                               cudaStream_t stream) {
  return launch_metric<2, 0>(metric, a, grid, block, smem, stream);
}
}  // namespace phnsw
