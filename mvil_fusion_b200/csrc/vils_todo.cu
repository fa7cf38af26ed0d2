// Entry points of include/vils_cabi.h that are not implemented yet return VILS_ERR_BAD_ARG with a message.
#include "common.h"
extern "C" {
#define TODO(name) return vils::fail(VILS_ERR_BAD_ARG, name ": not implemented yet")
int vils_ba_sharded_buffer(vils_ba*, void**, size_t*) { TODO("vils_ba_sharded_buffer"); }
int vils_ba_sharded_linearize(vils_ba*, int32_t) { TODO("vils_ba_sharded_linearize"); }
int vils_ba_sharded_update(vils_ba*, const vils_solve_opts*) { TODO("vils_ba_sharded_update"); }
}
