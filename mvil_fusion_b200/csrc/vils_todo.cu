// Entry points of include/vils_cabi.h that are not implemented yet return VILS_ERR_BAD_ARG with a message.
#include "common.h"
extern "C" {
#define TODO(name) return vils::fail(VILS_ERR_BAD_ARG, name ": not implemented yet")
int vils_ba_marginalize(vils_ba*, int32_t, int32_t, vils_prior_out*) { TODO("vils_ba_marginalize"); }
int vils_ba_sharded_buffer(vils_ba*, void**, size_t*) { TODO("vils_ba_sharded_buffer"); }
int vils_ba_sharded_linearize(vils_ba*, int32_t) { TODO("vils_ba_sharded_linearize"); }
int vils_ba_sharded_update(vils_ba*, const vils_solve_opts*) { TODO("vils_ba_sharded_update"); }
int vils_klt_create(int32_t, int32_t, int32_t, int32_t, int32_t, int32_t, vils_klt**) { TODO("vils_klt_create"); }
void vils_klt_destroy(vils_klt*) {}
int vils_klt_track(vils_klt*, const uint8_t*, const uint8_t*, int32_t, const float*, int32_t, float*, uint8_t*, float*) { TODO("vils_klt_track"); }
int vils_klt_upload(vils_klt*, const uint8_t*, const uint8_t*, int32_t, const float*, int32_t) { TODO("vils_klt_upload"); }
int vils_klt_track_device(vils_klt*) { TODO("vils_klt_track_device"); }
int vils_klt_download(vils_klt*, float*, uint8_t*, float*) { TODO("vils_klt_download"); }
int vils_klt_last_device_ms(vils_klt*, float*) { TODO("vils_klt_last_device_ms"); }
}
