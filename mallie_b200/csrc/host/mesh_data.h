// mesh_data.h -- owning host container for what struct Mesh (mesh.h:7-18) points at.
#ifndef MALLIE_B200_MESH_DATA_H_
#define MALLIE_B200_MESH_DATA_H_

#include <cstddef>
#include <string>
#include <vector>

namespace mb200 {

struct MeshData {
  std::vector<double> vertices;            // [3 * nv]
  std::vector<unsigned int> faces;         // [3 * nf]
  std::vector<unsigned int> material_ids;  // [nf]
  std::vector<double> normals;             // face-varying, [9 * nf]
  std::vector<double> uvs;                 // face-varying, [6 * nf]
  size_t num_shapes = 0;
};

// MeshLoader::LoadObj semantics (mesh_loader.cc): see mesh_loader.cc in this directory.
bool load_obj(MeshData &out, const char *filename, std::string *err);
// MeshLoader::LoadESON semantics (mesh_loader.cc:212-310).
bool load_eson(MeshData &out, const char *filename, std::string *err);
// Scene::Init's vertex transform (scene.cc:112-170): fit to [-1,1]^3 or uniform scale.
void apply_scene_transform(double *vertices, size_t nverts, double scene_scale, bool scene_fit, bool verbose);

} // namespace mb200

// The opaque C handle is the container itself.
struct mb200_mesh {
  mb200::MeshData mesh;
};

#endif
