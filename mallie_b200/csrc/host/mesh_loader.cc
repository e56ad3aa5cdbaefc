// mesh_loader.cc -- MeshLoader::LoadObj: Wavefront .obj -> struct Mesh with the vertex / face
// numbering of the reference loader, so faceIDs, f0..f2 and material ids agree on user files.
//
// Restates the behaviour (not the code) of
//   importers/tiny_obj_loader.cc:60-395,619-850   parsing, fan triangulation, per-export vertex dedupe
//   importers/mesh_loader.cc:15-210               shapes -> one Mesh, face-varying normals / uvs
// Behaviour that matters for parity and is easy to get wrong:
//   * numbers go through the loader's own decimal parser and are narrowed to float
//     (tiny_obj_loader.cc:106-268): mantissa accumulated digit by digit, fraction digits added as
//     d * pow(10,-k), result = ldexp(m * pow(5,e), e).  Kept operation for operation.
//   * the vertex-dedupe cache is passed BY VALUE into the export step (tiny_obj_loader.cc:355-356),
//     so vertices are shared only inside one face group (the faces between two usemtl/g/o lines).
//   * `usemtl` appends the pending faces to the current shape; `g` / `o` push the shape only if
//     faces are pending at that moment, then start a new shape either way (tiny_obj_loader.cc:737-818):
//     faces flushed by a usemtl that is directly followed by g/o are dropped.  Kept.
//   * mtllib is opened relative to the current directory (tiny_obj_loader.cc:604-616); a missing
//     file is not an error, every usemtl then yields material id -1.
//   * polygons become triangle fans (v0, v[k-1], v[k]) in file order (tiny_obj_loader.cc:369-395).
//   * without `vn`, face-varying normals are normalize(cross(v2-v0, v1-v0)) (mesh_loader.cc:15-22),
//     the opposite winding of BuildIntersection's geometric normal.
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <fstream>
#include <iterator>
#include <map>
#include <string>
#include <tuple>
#include <vector>

#include "mallie_api.h"
#include "mesh_data.h"

namespace mb200 {

namespace {

inline bool is_blank(char c) { return c == ' ' || c == '\t'; }
inline bool is_eol(char c) { return c == '\r' || c == '\n' || c == '\0'; }

// The loader's decimal grammar: [sign] digits ['.' digits] [(e|E) [sign] digits].  Returns false
// (value untouched) on a malformed number.
bool parse_decimal(const char *s, const char *end, double *out) {
  if (s >= end) return false;
  double mant = 0.0;
  int expo = 0;
  bool neg = false, eneg = false;
  const char *c = s;
  if (*c == '+' || *c == '-') {
    neg = (*c == '-');
    c++;
  } else if (!isdigit((unsigned char)*c)) {
    return false;
  }
  int nread = 0;
  bool more;
  while ((more = (c != end)) && isdigit((unsigned char)*c)) {
    mant *= 10;
    mant += (int)(*c - '0');
    c++, nread++;
  }
  if (nread == 0) return false;
  if (more) {
    bool to_exp = false;
    if (*c == '.') {
      c++;
      nread = 1;
      while ((more = (c != end)) && isdigit((unsigned char)*c)) {
        mant += (int)(*c - '0') * pow(10, -nread);
        nread++, c++;
      }
      to_exp = more;
    } else if (*c == 'e' || *c == 'E') {
      to_exp = true;
    }
    if (to_exp && (*c == 'e' || *c == 'E')) {
      c++;
      if ((more = (c != end)) && (*c == '+' || *c == '-')) {
        eneg = (*c == '-');
        c++;
      } else if (!isdigit((unsigned char)*c)) {
        return false;
      }
      nread = 0;
      while ((more = (c != end)) && isdigit((unsigned char)*c)) {
        expo *= 10;
        expo += (int)(*c - '0');
        c++, nread++;
      }
      if (eneg) expo = -expo;
      if (nread == 0) return false;
    }
  }
  *out = (neg ? -1 : 1) * ldexp(mant * pow(5, expo), expo);
  return true;
}

float next_float(const char *&tok) {
  tok += strspn(tok, " \t");
  const char *end = tok + strcspn(tok, " \t\r");
  double val = 0.0;
  parse_decimal(tok, end, &val);
  tok = end;
  return (float)val;
}

struct Corner {
  int v, vt, vn;
};

inline int zero_based(int idx, int n) { return idx > 0 ? idx - 1 : (idx == 0 ? 0 : n + idx); }

// i | i/j | i//k | i/j/k
Corner next_corner(const char *&tok, int nv, int nvn, int nvt) {
  Corner c = {-1, -1, -1};
  c.v = zero_based(atoi(tok), nv);
  tok += strcspn(tok, "/ \t\r");
  if (tok[0] != '/') return c;
  tok++;
  if (tok[0] == '/') {
    tok++;
    c.vn = zero_based(atoi(tok), nvn);
    tok += strcspn(tok, "/ \t\r");
    return c;
  }
  c.vt = zero_based(atoi(tok), nvt);
  tok += strcspn(tok, "/ \t\r");
  if (tok[0] != '/') return c;
  tok++;
  c.vn = zero_based(atoi(tok), nvn);
  tok += strcspn(tok, "/ \t\r");
  return c;
}

struct Shape {
  std::vector<float> positions, normals, texcoords;
  std::vector<unsigned int> indices;
  std::vector<int> material_ids;
};

struct Pools {
  std::vector<float> v, vn, vt;
};

// One face group -> appended to `shape`; vertices are shared inside this call only.
bool flush_group(Shape &shape, const Pools &in, const std::vector<std::vector<Corner>> &group, int material) {
  if (group.empty()) return false;
  std::map<std::tuple<int, int, int>, unsigned int> seen;
  auto vertex = [&](const Corner &c) -> unsigned int {
    const auto key = std::make_tuple(c.v, c.vn, c.vt);
    const auto it = seen.find(key);
    if (it != seen.end()) return it->second;
    for (int k = 0; k < 3; k++) {
      const size_t at = (size_t)(3 * c.v + k);
      shape.positions.push_back(c.v >= 0 && at < in.v.size() ? in.v[at] : 0.f);
    }
    if (c.vn >= 0)
      for (int k = 0; k < 3; k++) {
        const size_t at = (size_t)(3 * c.vn + k);
        shape.normals.push_back(at < in.vn.size() ? in.vn[at] : 0.f);
      }
    if (c.vt >= 0)
      for (int k = 0; k < 2; k++) {
        const size_t at = (size_t)(2 * c.vt + k);
        shape.texcoords.push_back(at < in.vt.size() ? in.vt[at] : 0.f);
      }
    const unsigned int idx = (unsigned int)(shape.positions.size() / 3 - 1);
    seen[key] = idx;
    return idx;
  };
  for (const auto &face : group) {
    if (face.size() < 2) continue;
    const Corner first = face[0];
    Corner prev, cur = face[1];
    for (size_t k = 2; k < face.size(); k++) {
      prev = cur;
      cur = face[k];
      const unsigned int a = vertex(first), b = vertex(prev), c = vertex(cur);
      shape.indices.push_back(a);
      shape.indices.push_back(b);
      shape.indices.push_back(c);
      shape.material_ids.push_back(material);
    }
  }
  return true;
}

// Only the material NAMES matter on this path (id = order of appearance).  The reference's reader
// registers a material when the NEXT newmtl arrives (if its name is non-empty) and once more,
// unconditionally, at end of file -- also for a missing file (tiny_obj_loader.cc:403-601).
void read_material_names(const char *path, std::map<std::string, int> &ids, int &count) {
  ids.clear();
  std::ifstream in(path);
  std::string current, line;
  while (in && std::getline(in, line)) {
    while (!line.empty() && (line.back() == '\n' || line.back() == '\r')) line.pop_back();
    const char *tok = line.c_str();
    tok += strspn(tok, " \t");
    if (strncmp(tok, "newmtl", 6) == 0 && is_blank(tok[6])) {
      if (!current.empty()) ids.insert(std::make_pair(current, count++));
      char name[4096] = {0};
      sscanf(tok + 7, "%4095s", name);
      current = name;
    }
  }
  ids.insert(std::make_pair(current, count++));
}

bool parse_obj(const char *filename, std::vector<Shape> &shapes, std::string *err) {
  std::ifstream in(filename);
  if (!in) {
    if (err) *err = std::string("Cannot open file [") + filename + "]";
    return false;
  }
  Pools pool;
  std::vector<std::vector<Corner>> group;
  std::map<std::string, int> material_ids;
  int material_count = 0, material = -1;
  Shape shape;
  std::string line;
  while (std::getline(in, line)) {
    while (!line.empty() && line.back() == '\n') line.pop_back();
    if (!line.empty() && line.back() == '\r') line.pop_back();
    if (line.empty()) continue;
    const char *tok = line.c_str();
    tok += strspn(tok, " \t");
    if (tok[0] == '\0' || tok[0] == '#') continue;

    if (tok[0] == 'v' && is_blank(tok[1])) {
      tok += 2;
      for (int k = 0; k < 3; k++) pool.v.push_back(next_float(tok));
    } else if (tok[0] == 'v' && tok[1] == 'n' && is_blank(tok[2])) {
      tok += 3;
      for (int k = 0; k < 3; k++) pool.vn.push_back(next_float(tok));
    } else if (tok[0] == 'v' && tok[1] == 't' && is_blank(tok[2])) {
      tok += 3;
      for (int k = 0; k < 2; k++) pool.vt.push_back(next_float(tok));
    } else if (tok[0] == 'f' && is_blank(tok[1])) {
      tok += 2;
      tok += strspn(tok, " \t");
      std::vector<Corner> face;
      while (!is_eol(tok[0])) {
        const char *before = tok;
        face.push_back(next_corner(tok, (int)(pool.v.size() / 3), (int)(pool.vn.size() / 3), (int)(pool.vt.size() / 2)));
        tok += strspn(tok, " \t\r");
        if (tok == before) break; // malformed token that cannot be consumed
      }
      group.push_back(face);
    } else if (strncmp(tok, "usemtl", 6) == 0 && is_blank(tok[6])) {
      char name[4096] = {0};
      sscanf(tok + 7, "%4095s", name);
      if (flush_group(shape, pool, group, material)) group.clear();
      const auto it = material_ids.find(name);
      material = (it != material_ids.end()) ? it->second : -1;
    } else if (strncmp(tok, "mtllib", 6) == 0 && is_blank(tok[6])) {
      char name[4096] = {0};
      sscanf(tok + 7, "%4095s", name);
      read_material_names(name, material_ids, material_count);
    } else if ((tok[0] == 'g' || tok[0] == 'o') && is_blank(tok[1])) {
      if (flush_group(shape, pool, group, material)) shapes.push_back(shape);
      shape = Shape();
      group.clear();
    }
  }
  if (flush_group(shape, pool, group, material)) shapes.push_back(shape);
  return true;
}

inline void facet_normal(double n[3], const float *v0, const float *v1, const float *v2) {
  const double a[3] = {(double)v1[0] - (double)v0[0], (double)v1[1] - (double)v0[1], (double)v1[2] - (double)v0[2]};
  const double b[3] = {(double)v2[0] - (double)v0[0], (double)v2[1] - (double)v0[1], (double)v2[2] - (double)v0[2]};
  // cross(v2 - v0, v1 - v0)
  n[0] = b[1] * a[2] - b[2] * a[1];
  n[1] = b[2] * a[0] - b[0] * a[2];
  n[2] = b[0] * a[1] - b[1] * a[0];
  const double len = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
  if (fabs(len) > 1.0e-6) {
    const double inv = 1.0 / len;
    n[0] *= inv, n[1] *= inv, n[2] *= inv;
  }
}

inline float at_or_zero(const std::vector<float> &a, size_t i) { return i < a.size() ? a[i] : 0.f; }

} // namespace

bool load_obj(MeshData &out, const char *filename, std::string *err) {
  std::vector<Shape> shapes;
  if (!parse_obj(filename, shapes, err)) return false;
  size_t nv = 0, nf = 0;
  for (const Shape &s : shapes) nv += s.positions.size() / 3, nf += s.indices.size() / 3;
  out = MeshData();
  out.vertices.resize(3 * nv);
  out.faces.resize(3 * nf);
  out.material_ids.assign(nf, 0u);
  out.normals.resize(9 * nf);
  out.uvs.assign(6 * nf, 0.0);
  out.num_shapes = shapes.size();
  size_t v0 = 0, f0 = 0;
  for (const Shape &s : shapes) {
    const size_t sf = s.indices.size() / 3, sv = s.positions.size() / 3;
    for (size_t f = 0; f < sf; f++) {
      for (int k = 0; k < 3; k++) out.faces[3 * (f0 + f) + k] = s.indices[3 * f + k] + (unsigned int)v0;
      out.material_ids[f0 + f] = (unsigned int)s.material_ids[f];
    }
    for (size_t i = 0; i < 3 * sv; i++) out.vertices[3 * v0 + i] = (double)s.positions[i];
    for (size_t f = 0; f < sf; f++) {
      const unsigned int idx[3] = {s.indices[3 * f], s.indices[3 * f + 1], s.indices[3 * f + 2]};
      double *N = &out.normals[9 * (f0 + f)];
      if (!s.normals.empty()) {
        for (int c = 0; c < 3; c++)
          for (int k = 0; k < 3; k++) N[3 * c + k] = (double)at_or_zero(s.normals, 3 * (size_t)idx[c] + k);
      } else {
        double n[3];
        facet_normal(n, &s.positions[3 * idx[0]], &s.positions[3 * idx[1]], &s.positions[3 * idx[2]]);
        for (int c = 0; c < 3; c++)
          for (int k = 0; k < 3; k++) N[3 * c + k] = n[k];
      }
      if (!s.texcoords.empty()) {
        double *T = &out.uvs[6 * (f0 + f)];
        for (int c = 0; c < 3; c++)
          for (int k = 0; k < 2; k++) T[2 * c + k] = (double)at_or_zero(s.texcoords, 2 * (size_t)idx[c] + k);
      }
    }
    v0 += sv, f0 += sf;
  }
  return true;
}


// ---- ESON -----------------------------------------------------------------------------------------
// LTE's binary container (importers/eson.cc:131-313): i64 total size, then elements
//   u8 tag | NUL-terminated key | payload
// with payload f64 (tag 1), i64 (tag 2), i64 n + n bytes (string 4, binary 6), i64 n + nested element
// (object 7).  MeshLoader::LoadESON (mesh_loader.cc:212-310) reads num_vertices, num_faces (i64),
// vertices (f32[3nv]), faces (i32[3nf]) and optional material_ids (u16[nf]); face-varying normals and
// uvs are left NULL even when present.
namespace {
struct EsonField {
  int tag = 0;
  int64_t i64 = 0;
  const unsigned char *ptr = nullptr;
  int64_t size = 0;
};

bool eson_scan(const std::vector<unsigned char> &buf, std::map<std::string, EsonField> &out, std::string *err) {
  auto fail = [&](const char *m) {
    if (err) *err = m;
    return false;
  };
  if (buf.size() < 8) return fail("ESON: file too short");
  int64_t total;
  memcpy(&total, buf.data(), 8);
  if (total <= 0 || (size_t)total > buf.size()) return fail("ESON: bad total size");
  size_t at = 8;
  while (at < (size_t)total) {
    EsonField f;
    f.tag = buf[at++];
    const void *nul = memchr(buf.data() + at, 0, (size_t)total - at);
    if (!nul) return fail("ESON: unterminated key");
    const std::string key(reinterpret_cast<const char *>(buf.data() + at));
    at += key.size() + 1;
    if (f.tag == 1 || f.tag == 2) {
      if (at + 8 > buf.size()) return fail("ESON: truncated scalar");
      memcpy(&f.i64, buf.data() + at, 8);
      at += 8;
    } else if (f.tag == 4 || f.tag == 6 || f.tag == 7) {
      if (at + 8 > buf.size()) return fail("ESON: truncated length");
      memcpy(&f.size, buf.data() + at, 8);
      at += 8;
      if (f.size < 0 || at + (size_t)f.size > buf.size()) return fail("ESON: truncated payload");
      f.ptr = buf.data() + at;
      at += (size_t)f.size;
    } else {
      return fail("ESON: unsupported element type");
    }
    out[key] = f;
  }
  return true;
}
} // namespace

bool load_eson(MeshData &out, const char *filename, std::string *err) {
  std::ifstream in(filename, std::ios::binary);
  if (!in) {
    if (err) *err = std::string("Failed to load file: ") + filename;
    return false;
  }
  std::vector<unsigned char> buf((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
  std::map<std::string, EsonField> f;
  if (!eson_scan(buf, f, err)) return false;
  auto need = [&](const char *k, int tag) { return f.count(k) && f[k].tag == tag; };
  if (!need("num_vertices", 2) || !need("num_faces", 2) || !need("vertices", 6) || !need("faces", 6)) {
    if (err) *err = "ESON: missing num_vertices / num_faces / vertices / faces";
    return false;
  }
  const int64_t nv = f["num_vertices"].i64, nf = f["num_faces"].i64;
  if (nv < 0 || nf < 0 || f["vertices"].size < nv * 12 || f["faces"].size < nf * 12) {
    if (err) *err = "ESON: array sizes do not match the counts";
    return false;
  }
  out = MeshData();
  out.num_shapes = 1;
  out.vertices.resize(3 * (size_t)nv);
  out.faces.resize(3 * (size_t)nf);
  out.material_ids.assign((size_t)nf, 0u);
  for (size_t i = 0; i < 3 * (size_t)nv; i++) {
    float v;
    memcpy(&v, f["vertices"].ptr + 4 * i, 4);
    out.vertices[i] = (double)v;
  }
  for (size_t i = 0; i < 3 * (size_t)nf; i++) {
    int v;
    memcpy(&v, f["faces"].ptr + 4 * i, 4);
    out.faces[i] = (unsigned int)v;
  }
  if (need("material_ids", 6) && f["material_ids"].size >= nf * 2)
    for (size_t i = 0; i < (size_t)nf; i++) {
      unsigned short v;
      memcpy(&v, f["material_ids"].ptr + 2 * i, 2);
      out.material_ids[i] = v;
    }
  return true;
}

} // namespace mb200

// ---- the reference-facing entry (importers/mesh_loader.h) ---------------------------------------------
// Arrays are allocated with new[] as in the reference: Scene's destructor deletes vertices, faces and
// materialIDs (scene.cc:57-64); the face-varying arrays are released by Scene as well here.
bool MeshLoader::LoadObj(Mesh &mesh, const char *filename) {
  mb200::MeshData d;
  std::string err;
  if (!mb200::load_obj(d, filename, &err)) {
    fprintf(stderr, "%s\n", err.c_str());
    return false;
  }
  printf("[LoadOBJ] # of shapes in .obj : %zu\n", d.num_shapes);
  printf("[LoadOBJ] # of faces: %zu\n", d.faces.size() / 3);
  printf("[LoadOBJ] # of vertices: %zu\n", d.vertices.size() / 3);
  memset(&mesh, 0, sizeof(mesh));
  mesh.numVertices = d.vertices.size() / 3;
  mesh.numFaces = d.faces.size() / 3;
  mesh.vertices = new real[d.vertices.size() + 1];
  mesh.faces = new unsigned int[d.faces.size() + 1];
  mesh.materialIDs = new unsigned int[d.material_ids.size() + 1];
  mesh.facevarying_normals = new real[d.normals.size() + 1];
  mesh.facevarying_uvs = new real[d.uvs.size() + 1];
  if (!d.vertices.empty()) memcpy(mesh.vertices, d.vertices.data(), d.vertices.size() * sizeof(real));
  if (!d.faces.empty()) memcpy(mesh.faces, d.faces.data(), d.faces.size() * sizeof(unsigned int));
  if (!d.material_ids.empty())
    memcpy(mesh.materialIDs, d.material_ids.data(), d.material_ids.size() * sizeof(unsigned int));
  if (!d.normals.empty()) memcpy(mesh.facevarying_normals, d.normals.data(), d.normals.size() * sizeof(real));
  if (!d.uvs.empty()) memcpy(mesh.facevarying_uvs, d.uvs.data(), d.uvs.size() * sizeof(real));
  return true;
}

bool MeshLoader::LoadESON(Mesh &mesh, const char *filename) {
  printf("[LoadESON] %s\n", filename);
  mb200::MeshData d;
  std::string err;
  if (!mb200::load_eson(d, filename, &err)) {
    fprintf(stderr, "%s\n", err.c_str());
    return false;
  }
  printf("# of vertices: %zu\n# of faces   : %zu\n", d.vertices.size() / 3, d.faces.size() / 3);
  memset(&mesh, 0, sizeof(mesh));
  mesh.numVertices = d.vertices.size() / 3;
  mesh.numFaces = d.faces.size() / 3;
  mesh.vertices = new real[d.vertices.size() + 1];
  mesh.faces = new unsigned int[d.faces.size() + 1];
  mesh.materialIDs = new unsigned int[d.material_ids.size() + 1];
  if (!d.vertices.empty()) memcpy(mesh.vertices, d.vertices.data(), d.vertices.size() * sizeof(real));
  if (!d.faces.empty()) memcpy(mesh.faces, d.faces.data(), d.faces.size() * sizeof(unsigned int));
  if (!d.material_ids.empty())
    memcpy(mesh.materialIDs, d.material_ids.data(), d.material_ids.size() * sizeof(unsigned int));
  return true; // face-varying normals / uvs stay NULL (mesh_loader.cc:303-307)
}
