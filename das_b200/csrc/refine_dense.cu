// Dense RecursiveUpdateLayer over a whole level (layers 1..L-1 when num_layers > 1).
// Reference: recursive_update.py:220-235 (layer), :186-197 (projections + gated blend), :34-82 (sampling).
#include "das_common.cuh"

extern "C" int das_refine_dense_layer(const das_levels* d_levels, const das_levels* h_levels, int32_t level,
                                      int32_t layer, const das_decode_cfg* cfg, const float* weights,
                                      const float* uvd_in, float* uvd_out, float* proj, void* stream) {
    (void)d_levels; (void)h_levels; (void)level; (void)layer; (void)cfg; (void)weights; (void)uvd_in; (void)uvd_out; (void)proj; (void)stream;
    das::set_error("das_refine_dense_layer: not built yet (num_layers > 1)");
    return DAS_ERR_UNSUPPORTED;
}
