// bvh_build.h -- host BVH container + builder entry points (see bvh_build.cc).
#ifndef MALLIE_B200_BVH_BUILD_H_
#define MALLIE_B200_BVH_BUILD_H_

#include <string>
#include <vector>

#include "mallie_b200.h"

namespace mb200 {

// Reference-layout BVH on the host: what BVHAccel keeps in nodes_/indices_ (bvh_accel.h:83-86).
struct HostBVH {
  std::vector<mb200_bvh_node> nodes;
  std::vector<uint32_t> indices;
  mb200_build_stats stats{0, 0, 0};
};

bool build_bvh(HostBVH &out, const double *vertices, size_t nverts, const uint32_t *faces, size_t nfaces,
               const mb200_build_options &opt, std::string *err);
// Same tree, grown level by level on the GPU (device/bvh_build_gpu.cu).  *cuda_failure tells a CUDA error from
// a rejected argument.
bool build_bvh_device(HostBVH &out, int device, const double *vertices, size_t nverts, const uint32_t *faces,
                      size_t nfaces, const mb200_build_options &opt, std::string *err, bool *cuda_failure);
// BVHAccel::Build + scene upload in one step, all on the device; bvh_out (may be null) receives the
// reference-layout tree.  Returns an mb200_status.
int scene_build_device(mb200_scene **out, int device, const double *vertices, size_t nverts, const uint32_t *faces,
                       size_t nfaces, const uint32_t *material_ids, const double *fv_normals, const double *fv_uvs,
                       const mb200_build_options &opt, HostBVH *bvh_out, std::string *err);
bool dump_bvh(const HostBVH &bvh, const char *path, std::string *err);
bool load_bvh(HostBVH &out, const char *path, std::string *err);

} // namespace mb200

// The opaque C handle is the container itself.
struct mb200_bvh {
  mb200::HostBVH bvh;
};

#endif
