// Voice-activity gate (placeholder until the conv+LSTM kernel lands in this file).
#include "../../include/wjb200.h"
#include "kernels.h"
extern "C" {
size_t wjb_vad_weights_bytes(void) { return 0; }
int wjb_vad_forward(const float*, int64_t, const int32_t*, int, const void*, float*, int, void*) {
    return wjb::set_error("wjb_vad_forward: not built yet");
}
}
