// console.cc -- the console front end's frame: DoMainConsole / HDRToLDR (main_console.cc:25-75).
#include <cstdio>
#include <vector>

#include "mallie_api.h"

namespace mallie {

namespace {
// fclamp, main_console.cc:25-32: float * double(255.5), truncated, clamped.
inline unsigned char quantise(float x) {
  const int i = (int)(x * 255.5);
  return (unsigned char)(i < 0 ? 0 : (i > 255 ? 255 : i));
}
} // namespace

void HDRToLDR(std::vector<unsigned char> &out, const std::vector<float> &in, const std::vector<int> &in_count,
              int width, int height) {
  out.resize((size_t)width * height * 3);
  for (size_t i = 0; i < (size_t)width * height * 3; i++) out[i] = quantise(in[i] / in_count[i / 3]);
}

static bool WritePPM(const char *output, const std::vector<unsigned char> &ldr, int width, int height) {
  FILE *fp = fopen(output, "wb");
  if (!fp) {
    printf("Mallie:err\tmsg:cannot write %s\n", output);
    return false;
  }
  fprintf(fp, "P6\n%d %d\n255\n", width, height);
  const bool ok = fwrite(ldr.data(), 1, ldr.size(), fp) == ldr.size();
  fclose(fp);
  printf("[Mallie] Output %s\n", output);
  return ok;
}

bool DoMainConsole(Scene &scene, const RenderConfig &config, const char *output, int passes) {
  printf("[Mallie] Console mode\n");
  const int width = config.width, height = config.height;
  if (width <= 0 || height <= 0) return false;
  std::vector<unsigned char> ldr;
  if (passes > 1 && config.num_gpus <= 1) {
    // Render + HDRToLDR on the device: the float frame never leaves the GPU (mb200_render_frame_ldr)
    mb200_render_stats st;
    const double mrays = RenderLDR(scene, config, ldr, config.eye, config.lookat, config.up, config.quat, passes,
                                   MB200_LDR_RGB8_LINEAR, &st);
    if (ldr.empty() || st.primary_rays == 0) return false;
    printf("[Mallie] %d passes: %.1f Mrays/s (%llu camera, %llu bounce, %llu shadow rays)\n", passes, mrays,
           (unsigned long long)st.primary_rays, (unsigned long long)st.bounce_rays,
           (unsigned long long)st.shadow_rays);
    return WritePPM(output, ldr, width, height);
  }
  std::vector<float> image((size_t)width * height * 3);
  std::vector<int> count((size_t)width * height);
  if (passes <= 1) {
    Render(scene, config, image, count, config.eye, config.lookat, config.up, config.quat, 1);
    printf("\n");
  } else {
    mb200_render_stats st;
    const double mrays = RenderAccumulate(scene, config, image, count, config.eye, config.lookat, config.up,
                                          config.quat, passes, &st);
    printf("[Mallie] %d passes: %.1f Mrays/s (%llu camera, %llu bounce, %llu shadow rays)\n", passes, mrays,
           (unsigned long long)st.primary_rays, (unsigned long long)st.bounce_rays,
           (unsigned long long)st.shadow_rays);
  }
  if (count[0] == 0) return false; // nothing was rendered
  HDRToLDR(ldr, image, count, width, height);
  return WritePPM(output, ldr, width, height);
}

} // namespace mallie
