// placeholder translation unit: replaced by the FMM inpaint implementation
#include "common.cuh"
extern "C" {
int ofxcv_inpaint_u8(ofxcv_ctx*, ofxcv_stream, const uint8_t*, ptrdiff_t, int, const uint8_t*, ptrdiff_t, uint8_t*, ptrdiff_t, int, int, double, int) { return OFXCV_ERR_UNSUPPORTED; }
int ofxcv_inpaint_u8_host(ofxcv_ctx*, const uint8_t*, ptrdiff_t, int, const uint8_t*, ptrdiff_t, uint8_t*, ptrdiff_t, int, int, double, int) { return OFXCV_ERR_UNSUPPORTED; }
size_t ofxcv_inpaint_workspace_bytes(int, int, int) { return 0; }
int ofxcv_inpaint_last_stats(const ofxcv_ctx*, int64_t*) { return OFXCV_ERR_UNSUPPORTED; }
}
