// host_api_check.cc -- exercises the host C++ mirror of the Mallie API (mallie_api.h) the way a
// Mallie program would, and dumps what it got so tests/test_gpu_host_api.py can compare it with the
// oracle:   host_api_check <obj> <out.bin> <width> <height> [plane] [gpus] [panoramic]
//   1. Scene::Init(obj) (loader + host BVH build), Scene::BoundingBox
//   2. Camera::BuildCameraFrame + Camera::GenerateRay for every pixel, Scene::TraceBatch
//   3. Scene::Trace for a handful of single rays (must equal the batch entries)
//   4. mallie::Render (one pass, step 1), then Render with step 4 (coarse preview)
//   5. (panoramic != 0) Camera::GenerateEnvRay / GenerateStereoEnvRay on a few pixels, RenderPanoramic mono + stereo
// Written against the reference-style headers through the forwarding includes.
#include <string>

#include "scene.h"
#include "camera.h"
#include "render.h"

#include <cstdio>
#include <cstring>
#include <vector>

static void put(FILE *fp, const void *p, size_t n) { fwrite(p, 1, n, fp); }

int main(int argc, char **argv) {
  if (argc < 5) return 64;
  const int W = atoi(argv[3]), H = atoi(argv[4]);
  const bool plane = argc > 5 && atoi(argv[5]) != 0;
  mallie::Scene scene;
  if (!scene.Init(argv[1], "", "", "")) return 1;
  real3 bmin, bmax;
  scene.BoundingBox(bmin, bmax);

  mallie::RenderConfig config;
  config.width = W, config.height = H, config.plane = plane;
  if (argc > 6) config.num_gpus = atoi(argv[6]);
  config.eye[0] = 0.4, config.eye[1] = 0.9, config.eye[2] = 6.0;
  config.lookat[0] = 0.5, config.lookat[1] = 0.8, config.lookat[2] = 0.0;

  mallie::Camera camera(config.eye, config.lookat, config.up);
  double origin[3], corner[3], du[3], dv[3];
  camera.BuildCameraFrame(origin, corner, du, dv, config.fov, config.quat, W, H);
  std::vector<Ray> rays((size_t)W * H);
  for (int y = 0; y < H; y++)
    for (int x = 0; x < W; x++) rays[(size_t)y * W + x] = camera.GenerateRay((double)x, (double)y);
  std::vector<Intersection> isects(rays.size());
  memset(static_cast<void *>(isects.data()), 0, isects.size() * sizeof(Intersection));
  std::vector<unsigned char> mask(rays.size());
  const long nhit = scene.TraceBatch(isects.data(), rays.data(), rays.size(), mask.data());
  if (nhit < 0) return 2;

  // single-ray calls agree with the batch
  int single_bad = 0;
  for (size_t i = 0; i < rays.size(); i += rays.size() / 13 + 1) {
    Intersection one;
    memset(static_cast<void *>(&one), 0, sizeof(one));
    const bool hit = scene.Trace(one, rays[i]);
    if (hit != (mask[i] != 0) || memcmp(&one, &isects[i], hit ? sizeof(one) : 32 - 4) != 0) single_bad++;
  }

  std::vector<float> image((size_t)W * H * 3, -1.f), coarse((size_t)W * H * 3, -1.f);
  std::vector<int> count((size_t)W * H, 0), coarse_count((size_t)W * H, 0);
  mallie::Render(scene, config, image, count, config.eye, config.lookat, config.up, config.quat, 1);
  mallie::Render(scene, config, coarse, coarse_count, config.eye, config.lookat, config.up, config.quat, 4);
  printf("\n");

  FILE *fp = fopen(argv[2], "wb");
  if (!fp) return 3;
  const long long hdr[4] = {W, H, nhit, single_bad};
  put(fp, hdr, sizeof(hdr));
  put(fp, &bmin, sizeof(bmin));
  put(fp, &bmax, sizeof(bmax));
  put(fp, origin, sizeof(origin)), put(fp, corner, sizeof(corner)), put(fp, du, sizeof(du)), put(fp, dv, sizeof(dv));
  put(fp, isects.data(), isects.size() * sizeof(Intersection));
  put(fp, mask.data(), mask.size());
  put(fp, image.data(), image.size() * sizeof(float));
  put(fp, count.data(), count.size() * sizeof(int));
  put(fp, coarse.data(), coarse.size() * sizeof(float));
  put(fp, coarse_count.data(), coarse_count.size() * sizeof(int));
  if (argc > 7 && atoi(argv[7]) != 0) {
    for (int k = 0; k < 5; k++) { // single host rays through the panorama cameras
      const Ray a = camera.GenerateEnvRay(0.37 * W * k / 4.0, 0.61 * H * k / 4.0);
      const Ray b = camera.GenerateStereoEnvRay(0.93 * W * k / 4.0, 0.99 * H * k / 4.0);
      put(fp, &a.org, 48), put(fp, &b.org, 48);
    }
    for (int stereo = 0; stereo < 2; stereo++) {
      std::vector<float> pano((size_t)W * H * 3, -1.f);
      std::vector<int> pcount((size_t)W * H, 0);
      mallie::RenderPanoramic(scene, config, pano, pcount, config.eye, config.lookat, config.up, config.quat, stereo != 0);
      put(fp, pano.data(), pano.size() * sizeof(float));
      put(fp, pcount.data(), pcount.size() * sizeof(int));
    }
    printf("\n");
  }
  fclose(fp);
  printf("host_api_check: %ld hits of %zu rays, %d single-ray mismatches\n", nhit, rays.size(), single_bad);
  return single_bad ? 4 : 0;
}
