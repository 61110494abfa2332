// l2_knn.cu -- L2 kNN (k = 2) for float32 descriptors.  (placeholder until the kernel lands)
#include "knn.cuh"

namespace mcv {

int launch_l2_knn2(Context* ctx, cudaStream_t stream, const KnnArgs& a) {
    (void)ctx; (void)stream; (void)a;
    set_error("NORM_L2 matcher kernel is not built into this library yet");
    return -1;
}

}  // namespace mcv
