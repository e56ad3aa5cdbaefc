// main.cc -- `mallie_b200_cli [config.json] [--passes N] [--output file.ppm]`: the reference's console
// front end (main.cc:209-289, main_console.cc:57-75) over the B200 backend: load config.json, Scene::Init,
// one Render() pass (or N accumulated passes), write the image.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "mallie_api.h"

int main(int argc, char **argv) {
  std::string config_filename("config.json"), output("output.ppm");
  int passes = 1;
  for (int i = 1; i < argc; i++) {
    if (strcmp(argv[i], "--help") == 0) {
      printf("Usage: mallie_b200_cli <config.json> [--passes N] [--output out.ppm]\n");
      return 1;
    } else if (strcmp(argv[i], "--passes") == 0 && i + 1 < argc) {
      passes = atoi(argv[++i]);
    } else if (strcmp(argv[i], "--output") == 0 && i + 1 < argc) {
      output = argv[++i];
    } else {
      config_filename = argv[i];
    }
  }
  printf("Mallie:info\tVersion  : %s\n", mb200_version());
  printf("Mallie:info\tPrecision: 64bit double\n");
  printf("Mallie:info\t# of GPUs: %d\n", mb200_device_count());
  printf("Mallie:info\tConfig file: %s\n", config_filename.c_str());
  mallie::RenderConfig config;
  if (!mallie::LoadJSONConfig(config, config_filename)) {
    printf("Mallie:err\tmsg:cannot read %s\n", config_filename.c_str());
    return 2;
  }
  if (mb200_device_count() < 1) {
    printf("Mallie:err\tmsg:no CUDA device: mallie_b200 has no CPU rendering path\n");
    return 3;
  }
  mallie::Scene scene;
  scene.SetDevice(config.device);
  if (!scene.Init(config.obj_filename, config.eson_filename, config.magicavoxel_filename, config.material_filename,
                  config.scene_scale, config.scene_fit))
    return 4;
  printf("Mallie:info\tBegin\n");
  const bool ok = mallie::DoMainConsole(scene, config, output.c_str(), passes);
  printf("Mallie:info\tEnd\n");
  return ok ? 0 : 5;
}
