// tcgen05 (MFT_PREC_TF32) edge-MLP layers.  Placeholder until the tensor-core
// kernels land: reports the path as unavailable, never computes on another path.
#include "common.cuh"
#include "wcompute.cuh"
#include "umma.cuh"

namespace mft {

bool umma_shape_supported(int, int) { return false; }

int wcompute_fwd_layers_tf32(const float*, int, int, int, const mft_wcompute_params*, const WcLayout&,
                             const PairGeom&, cudaStream_t) {
    set_error(MFT_ERR_UNSUPPORTED, "TF32 tcgen05 path not built into this library");
    return MFT_ERR_UNSUPPORTED;
}
int wcompute_bwd_layer_tf32(int, float*, float*, const float*, int, float*, int, int, const mft_wcompute_params*,
                            const mft_wcompute_grads*, const WcLayout&, const PairGeom&, cudaStream_t) {
    set_error(MFT_ERR_UNSUPPORTED, "TF32 tcgen05 path not built into this library");
    return MFT_ERR_UNSUPPORTED;
}

}  // namespace mft
