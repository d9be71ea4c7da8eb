// placeholder until the tcgen05 kernel lands (replaced in the next commit)
#include "common.cuh"
extern "C" long occnerf_mlp_packed_bytes(int) { return 0; }
extern "C" int occnerf_mlp_pack_weights(const occnerf_mlp_params *, int, void *, occnerf_stream_t) {
    occnerf_set_error("mlp_tc: not built"); return OCCNERF_EINVAL; }
extern "C" int occnerf_mlp_forward_tc(const float *, int, const void *, int, float *, int, void *, occnerf_stream_t) {
    occnerf_set_error("mlp_tc: not built"); return OCCNERF_EINVAL; }
