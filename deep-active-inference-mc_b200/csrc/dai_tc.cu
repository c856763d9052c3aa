// Tensor-core path (tcgen05 / TMEM / TMA) of the contraction layers.
#include "dai_tc.h"

namespace dai {

int tc_pack_weights(const std::map<std::string, std::vector<float>>& raw, TcWeights* out, std::vector<void*>* allocs,
                    std::string* err) {
    (void)raw; (void)allocs; (void)err;
    out->impl = nullptr;
    return 0;
}

void tc_release(TcWeights* w) { w->impl = nullptr; }

int tc_decoder_chunk(const TcWeights&, const DevWeights&, int, const float*, const uint32_t*, int, void*, void*, void*,
                     void*, const Ct4Args&, cudaStream_t, std::string* err) {
    if (err) *err = "tensor-core decoder not built yet";
    return -1;
}

}  // namespace dai
