// placeholder: replaced by the tcgen05/TMEM decoder
#include "common.cuh"
using namespace gnb;
extern "C" int64_t gnb_decoder_packed_bytes(const GnbDecoderWeights* w) { (void)w; return 0; }
extern "C" int gnb_decoder_pack_bf16(const GnbDecoderWeights* w, void* packed, void* stream) {
    (void)w, (void)packed, (void)stream;
    set_error("gnb_decoder_pack_bf16: not built yet");
    return GNB_E_UNSUPPORTED;
}
extern "C" int gnb_decode_bf16(const GnbDecoderWeights* w, const void* packed, const float* xyz, const float* feat,
                               int64_t n_rows, float* out, float* tsdf, void* stream) {
    (void)w, (void)packed, (void)xyz, (void)feat, (void)n_rows, (void)out, (void)tsdf, (void)stream;
    set_error("gnb_decode_bf16: not built yet");
    return GNB_E_UNSUPPORTED;
}
extern "C" int gnb_query_fused_bf16(const GnbSampleParams* s, const GnbDecoderWeights* w, const void* packed, float* out,
                                    float* tsdf, void* stream) {
    (void)s, (void)w, (void)packed, (void)out, (void)tsdf, (void)stream;
    set_error("gnb_query_fused_bf16: not built yet");
    return GNB_E_UNSUPPORTED;
}
