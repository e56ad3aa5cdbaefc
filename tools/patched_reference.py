#!/usr/bin/env python
"""INTEGRATION.md, Option A, carried out: the UNMODIFIED Mallie front end (main.cc, main_console.cc, loaders, JPEG
writer, ...) with the three `#ifdef ENABLE_B200` hunks applied to scene.h / scene.cc / render.cc, linked against
libmallie_b200.so.  Authoring container only (needs /root/reference).  Nothing from the reference is committed: the
sources are copied to oracle/_ref/patched/ (git-ignored), patched there by anchor, and built into
oracle/_ref/mallie_patched, which travels to the GPU box like the other built artefacts.

    python tools/patched_reference.py            # copy + patch + build
    cd <dir with config.json> && <repo>/oracle/_ref/mallie_patched config.json      # on a GPU box: writes output.jpg
"""
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
DST = os.path.join(ROOT, "oracle", "_ref", "patched")
OUT = os.path.join(ROOT, "oracle", "_ref", "mallie_patched")

CXX_SRC = ["main.cc", "main_console.cc", "render.cc", "scene.cc", "bvh_accel.cc", "camera.cc", "matrix.cc", "trackball.cc",
           "prim-plane.cc", "jpge.cc", "script_engine.cc", "filepath_util.cc", "importers/mesh_loader.cc",
           "importers/tiny_obj_loader.cc", "importers/eson.cc", "importers/magicavoxel_loader.cc"]
C_SRC = ["duktape.c", "deps/parson/parson.c"]


def insert_after(text, anchor, block, which=0):
    at = -1
    for _ in range(which + 1):
        at = text.index(anchor, at + 1)
    eol = text.index("\n", at) + 1
    return text[:eol] + block + text[eol:]


def insert_before(text, anchor, block):
    at = text.index(anchor)
    bol = text.rfind("\n", 0, at) + 1
    return text[:bol] + block + text[bol:]


def patch_scene_h(t):
    t = insert_after(t, '#include "bvh_accel.h"', '#ifdef ENABLE_B200\n#include "mallie_b200.h"\n#endif\n')
    t = insert_before(t, "protected:", "#ifdef ENABLE_B200\n  mb200_scene *b200() { return b200_; }\n#endif\n\n")
    t = insert_after(t, "std::vector<Material> materials_;", "#ifdef ENABLE_B200\n  mb200_scene *b200_ = NULL;\n#endif\n")
    return t


def patch_scene_cc(t):
    t = insert_before(t, "delete[] mesh_.vertices;", "#ifdef ENABLE_B200\n  mb200_scene_destroy(b200_);\n#endif\n")
    upload = '''#ifdef ENABLE_B200
  {
    const std::vector<BVHNode> &n = accel_.GetNodes();
    const std::vector<unsigned int> &ix = accel_.GetIndices();
    int rc = mb200_scene_create(&b200_, /*device*/ 0, mesh_.vertices, mesh_.numVertices, mesh_.faces, mesh_.numFaces,
                                mesh_.materialIDs, mesh_.facevarying_normals, mesh_.facevarying_uvs,
                                reinterpret_cast<const mb200_bvh_node *>(&n[0]), n.size(), &ix[0], ix.size());
    if (rc != MB200_OK) {
      printf("Mallie:err\\tmsg:%s\\n", mb200_last_error());
      return false;
    }
  }
#endif
'''
    at = t.index("ret = accel_.Build(&mesh_, options);")
    # after the assert that follows the Build call
    a2 = t.index("assert(ret);", at)
    eol = t.index("\n", a2) + 1
    return t[:eol] + upload + t[eol:]


def patch_render_cc(t):
    hook = '''#ifdef ENABLE_B200
  {
    mb200_render_params p;
    mb200_render_params_default(&p, width, height); // max_path_length 16, jitter on, shader = PathTrace
    for (int c = 0; c < 3; c++) {
      p.frame.origin[c] = origin[c];
      p.frame.corner[c] = corner[c];
      p.frame.du[c] = du[c];
      p.frame.dv[c] = dv[c];
    }
    p.use_plane = gPlane;
    if (gPlane) {
      p.plane[0] = gPlaneObject.m_a, p.plane[1] = gPlaneObject.m_b;
      p.plane[2] = gPlaneObject.m_c, p.plane[3] = gPlaneObject.m_d;
    }
    static unsigned b200_pass = 0;
    p.pass = b200_pass++;
    p.pixel_step = step;
    if (mb200_render_pass(scene.b200(), &p, &image[0], &count[0], NULL) != MB200_OK)
      printf("Mallie:err\\tmsg:%s\\n", mb200_last_error());
  }
#else
'''
    t = insert_before(t, "#if !defined(_OPENMP) // Tasksys version", hook)
    t = insert_after(t, "#endif // !OMP version", "#endif // ENABLE_B200\n")
    return t


def main():
    if not os.path.isdir(REF):
        print("patched_reference: /root/reference is absent; keeping the prebuilt binary (if any)")
        return 0
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    # headers and sources the front end needs (the whole flat directory minus the big vendored trees)
    def ignore(d, names):
        skip = {"SDL2-2.0.3", "extlibs", "gtest-1.7.0", "ptex-master", ".git", "tools", "test"}
        return [n for n in names if n in skip or n.endswith((".dll", ".jpg", ".eson", ".obj", ".vox"))]
    shutil.copytree(REF, DST, ignore=ignore)
    for name, fn in (("scene.h", patch_scene_h), ("scene.cc", patch_scene_cc), ("render.cc", patch_render_cc)):
        p = os.path.join(DST, name)
        os.chmod(p, 0o644)
        text = open(p).read()
        open(p, "w").write(fn(text))
    inc = ["-I" + DST, "-I" + os.path.join(DST, "importers"), "-I" + os.path.join(DST, "deps", "parson"),
           "-I" + os.path.join(DST, "deps", "TinyThread++-1.1", "source"), "-I" + os.path.join(ROOT, "include")]
    flags = ["-O2", "-fopenmp", "-msse2", "-DNDEBUG", "-DENABLE_B200", "-w", "-D__STDC_CONSTANT_MACROS", "-D__STDC_LIMIT_MACROS"]
    objs = []
    os.makedirs(os.path.join(DST, "obj"), exist_ok=True)
    jobs = []
    for src in CXX_SRC + C_SRC:
        o = os.path.join(DST, "obj", os.path.basename(src) + ".o")
        objs.append(o)
        cc = ["g++", "-include", "string"] if src.endswith(".cc") else ["gcc", "-O1"]
        jobs.append(subprocess.Popen(cc + flags + inc + ["-c", os.path.join(DST, src), "-o", o]))
    for j in jobs:
        if j.wait() != 0:
            print("patched_reference: compile failed")
            return 1
    subprocess.check_call(["g++", "-fopenmp", "-o", OUT] + objs +
                          ["-L" + os.path.join(ROOT, "mallie_b200"), "-lmallie_b200", "-Wl,-rpath,$ORIGIN/../../mallie_b200",
                           "-lpthread"])
    if "--keep" not in sys.argv:
        shutil.rmtree(DST)              # only the binary stays (and ships to the GPU box)
    print("patched_reference: built", OUT)
    return 0


if __name__ == "__main__":
    sys.exit(main())
