// lcrt_accel.h — instance-buffer maintenance (host+device portable per-thread functions).
//
// Same behaviour as the reference's two built-in kernels, over the same 80-byte records:
//   update_accel                    src/backends/cuda/cuda_builtin/cuda_builtin_kernels.cu:52-111
//   update_accel_instance_handles   src/backends/cuda/cuda_builtin/cuda_builtin_kernels.cu:42-50
// and as the in-kernel accessors lc_accel_set_instance_* (cuda_device_resource.h:1470-1510).
#pragma once
#include "lcrt_common.h"

namespace lcrt {

LCRT_HD void apply_modification(Instance *instances, const Modification &m) {
    Instance &in = instances[m.index];
    uint32_t flags = in.flags;
    in.sbt_offset = 0u;
    if (m.flags & MOD_PRIMITIVE) {
        in.handle = m.primitive;
        // both triangle meshes and procedural primitives get DISABLE_TRIANGLE_FACE_CULLING (:76-88); curves are
        // out of scope
        flags |= INSTANCE_FLAG_DISABLE_TRIANGLE_FACE_CULLING;
    }
    if (m.flags & MOD_VISIBILITY) in.mask = m.vis_mask;
    if (m.flags & (MOD_OPAQUE_ON | MOD_OPAQUE_OFF)) {
        if (flags & INSTANCE_FLAG_DISABLE_TRIANGLE_FACE_CULLING) {
            flags &= ~(INSTANCE_FLAG_DISABLE_ANYHIT | INSTANCE_FLAG_ENFORCE_ANYHIT);
            flags |= (m.flags & MOD_OPAQUE_ON) ? INSTANCE_FLAG_DISABLE_ANYHIT : INSTANCE_FLAG_ENFORCE_ANYHIT;
        }
    }
    if (m.flags & MOD_USER_ID) in.user_id = m.user_id;
    in.flags = flags;
    if (m.flags & MOD_TRANSFORM)
        for (int k = 0; k < 12; k++) in.affine[k] = m.affine[k];
}

LCRT_HD void apply_handle_update(Instance *instances, const InstanceHandleUpdate &u) { instances[u.index].handle = u.handle; }

// fresh instance slots: identity transform, mask 0xff, opaque (accel.rs:284-295)
LCRT_HD void init_instance(Instance &in) {
    for (int k = 0; k < 12; k++) in.affine[k] = (k == 0 || k == 5 || k == 10) ? 1.f : 0.f;
    in.user_id = 0u, in.sbt_offset = 0u, in.mask = 0xffu, in.flags = INSTANCE_FLAG_DISABLE_ANYHIT;
    in.handle = 0ull, in.pad = 0ull;
}

}// namespace lcrt
