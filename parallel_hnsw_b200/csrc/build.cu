// build.cu -- placeholder until the device build lands (next commit)
#include "internal.h"
using namespace phnsw;
extern "C" {
phnsw_status phnsw_generate(phnsw_store *, const uint64_t *, uint64_t, const phnsw_build_params *,
                            uint64_t, phnsw_progress_fn, void *, phnsw_index **) {
  set_error("generate: not implemented yet");
  return PHNSW_ERR_INVALID;
}
phnsw_status phnsw_improve_index(phnsw_index *, const phnsw_build_params *, phnsw_progress_fn,
                                 void *, float *) {
  set_error("improve_index: not implemented yet");
  return PHNSW_ERR_INVALID;
}
phnsw_status phnsw_stochastic_recall(const phnsw_index *, const phnsw_optimization_params *,
                                     float *) {
  set_error("stochastic_recall: not implemented yet");
  return PHNSW_ERR_INVALID;
}
}
