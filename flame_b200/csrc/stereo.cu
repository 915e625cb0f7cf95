// placeholder — replaced by the epipolar search kernels
#include "internal.h"
namespace fb {
int stereo_set_camera(void**, int, int, const float*, const float*) { set_error("stereo: not built yet"); return FLAME_B200_ESTATE; }
int stereo_frame_upload(void**, cudaStream_t, int, const uint8_t*, const float*, const float*, int, const float*) { set_error("stereo: not built yet"); return FLAME_B200_ESTATE; }
int stereo_update(void*, cudaStream_t, const flame_b200_track_params*, int, int, flame_b200_feature*, int32_t, flame_b200_track_diag*, int32_t*) { set_error("stereo: not built yet"); return FLAME_B200_ESTATE; }
void stereo_destroy(void*) {}
}
