// search_launch.cuh -- instantiates one variant (PQ, TREE) of the traversal kernel for the four
// metrics; included by search_seq.cu / search_tree.cu / search_pq.cu.
#pragma once
#include "internal.h"
#include <mutex>
#include <string.h>

namespace phnsw {

template <int METRIC, int PQ, int TREE, int MODE>
static cudaError_t launch_moded(const SearchArgs &a, int grid, int block, size_t smem,
                                cudaStream_t stream) {
  // the attribute belongs to the function (per device), not to the calling thread: keep one
  // high-water mark per device for all host threads, and only ever raise it
  static std::mutex mu;
  static size_t configured[16] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  {
    std::lock_guard<std::mutex> g(mu);
    if (dev >= 16 || configured[dev] < smem) {
      cudaError_t e = cudaFuncSetAttribute(search_kernel<METRIC, PQ, TREE, MODE>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
      if (dev < 16) configured[dev] = smem;
    }
  }
  // several CTAs per SM (batch overlap): the protocol needs every CTA of the launch resident at
  // once, so ask for the largest shared-memory carve-out and check what the device grants; the
  // caller falls back to one CTA per SM on cudaErrorInvalidConfiguration
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (sms > 0 && grid > sms) {
    static std::mutex omu;
    static int ok_block[16] = {0}, ok_per_sm[16] = {0};
    static size_t ok_smem[16] = {0};
    std::lock_guard<std::mutex> g(omu);
    const int per_sm = (grid + sms - 1) / sms;
    if (dev >= 16 || ok_block[dev] != block || ok_smem[dev] != smem || ok_per_sm[dev] < per_sm) {
      cudaFuncSetAttribute(search_kernel<METRIC, PQ, TREE, MODE>,
                           cudaFuncAttributePreferredSharedMemoryCarveout,
                           (int)cudaSharedmemCarveoutMaxShared);
      int nb = 0;
      cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(
          &nb, search_kernel<METRIC, PQ, TREE, MODE>, block, smem);
      if (e != cudaSuccess || nb < per_sm) return cudaErrorInvalidConfiguration;
      if (dev < 16) { ok_block[dev] = block; ok_smem[dev] = smem; ok_per_sm[dev] = nb; }
    }
  }
  if (a.overlap == 2) {  // chained behind another search launch: programmatic dependent launch
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, search_kernel<METRIC, PQ, TREE, MODE>, a);
  }
  search_kernel<METRIC, PQ, TREE, MODE><<<grid, block, smem, stream>>>(a);
  return cudaGetLastError();
}
template <int METRIC, int PQ, int TREE>
static cudaError_t launch_typed(const SearchArgs &a, int grid, int block, size_t smem,
                                cudaStream_t stream) {
  if (a.mode == 0) return launch_moded<METRIC, PQ, TREE, 0>(a, grid, block, smem, stream);
  if (PQ) return cudaErrorInvalidValue;  // the ADC store supports search_layers only
  if (a.mode == 1) return launch_moded<METRIC, PQ ? 0 : PQ, TREE, PQ ? 0 : 1>(a, grid, block, smem, stream);
  return launch_moded<METRIC, PQ ? 0 : PQ, TREE, PQ ? 0 : 2>(a, grid, block, smem, stream);
}
template <int PQ, int TREE>
static cudaError_t launch_metric(int metric, const SearchArgs &a, int grid, int block, size_t smem,
                                 cudaStream_t stream) {
  switch (metric) {
    case kCosHalf: return launch_typed<kCosHalf, PQ, TREE>(a, grid, block, smem, stream);
    case kOneMinusDot: return launch_typed<kOneMinusDot, PQ, TREE>(a, grid, block, smem, stream);
    case kL2Sqrt: return launch_typed<kL2Sqrt, PQ, TREE>(a, grid, block, smem, stream);
    default: return launch_typed<kCosClamp, PQ, TREE>(a, grid, block, smem, stream);
  }
}

}  // namespace phnsw
