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

bool DoMainConsole(Scene &scene, const RenderConfig &config, const char *output, int passes) {
  printf("[Mallie] Console mode\n");
  const int width = config.width, height = config.height;
  if (width <= 0 || height <= 0) return false;
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
  std::vector<unsigned char> ldr;
  HDRToLDR(ldr, image, count, width, height);
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

} // namespace mallie
