// synth.cu -- whole-network bf16 engine (placeholder while the tcgen05 path is brought up).
#include "common.cuh"
using namespace sg2;
struct sg2_synth { int size; };
extern "C" int sg2_synth_create(sg2_synth **plan, int, int, int, const sg2_conv_params *, int, const float *, const float *) {
    if (plan) *plan = nullptr;
    set_error("synthesis engine not built yet");
    return SG2_ERR_UNSUPPORTED;
}
extern "C" void sg2_synth_destroy(sg2_synth *p) { delete p; }
extern "C" int64_t sg2_synth_workspace_bytes(const sg2_synth *) { return 0; }
extern "C" int sg2_synth_describe(const sg2_synth *, char *, int) { return 0; }
extern "C" int sg2_synth_pack(sg2_synth *, void *, sg2_stream_t) { set_error("synthesis engine not built yet"); return SG2_ERR_UNSUPPORTED; }
extern "C" int sg2_synth_forward(sg2_synth *, void *, const float *, int64_t, const float *const *, const int64_t *, float *, sg2_stream_t) { set_error("synthesis engine not built yet"); return SG2_ERR_UNSUPPORTED; }
