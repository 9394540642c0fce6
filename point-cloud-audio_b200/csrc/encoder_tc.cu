// tcgen05 / TMEM set-encoder kernels for the audio dims (D=64, H=8, M=64) -- placeholder until the
// kernels land; reports "unsupported" so callers fail loudly instead of silently falling back.
#include "common.cuh"
namespace pca {
size_t st_tc_workspace_bytes(const pca_st_dims*, int, int) { return 0; }
int st_tc_supported(const pca_st_dims*, int) { return 0; }
int st_tc_forward(const float*, int, int, const pca_st_dims*, const float*, float*, void*, size_t, cudaStream_t) {
    return fail(PCA_EUNSUPPORTED, "tcgen05 encoder path not built");
}
}  // namespace pca
