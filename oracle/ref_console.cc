// ref_console.cc -- TEST INFRASTRUCTURE ONLY (see oracle/Makefile).
//
// HDRToLDR / fclamp of the reference's console front end live in an anonymous namespace of main_console.cc
// (main_console.cc:22-43), so they cannot be linked against.  This translation unit compiles the UNMODIFIED
// reference file where it lies (the #include below pulls in /root/reference/main_console.cc through the
// -I$(REF) search path; nothing is copied into this repository) and exports the one function the tests pin
// the device-side output resolve against.
#include "main_console.cc"

extern "C" void ref_hdr_to_ldr(const float *in, const int *in_count, int width, int height, unsigned char *out) {
  std::vector<float> image(in, in + (size_t)width * height * 3);
  std::vector<int> count(in_count, in_count + (size_t)width * height);
  std::vector<unsigned char> ldr;
  mallie::HDRToLDR(ldr, image, count, width, height);
  for (size_t i = 0; i < ldr.size(); i++) out[i] = ldr[i];
}
