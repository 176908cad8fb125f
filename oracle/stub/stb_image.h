#pragma once
// Stand-in for nothings/stb stb_image.h (fetched from the network by the reference's CMakeLists.txt:17-21, absent
// here).  scene/texture/bitmap.hpp:15,33 only needs these two symbols; oracle/ref_harness.cpp defines them to hand
// back texels that were decoded ahead of time (RTSC texel blob).  Test infrastructure only.
extern "C" unsigned char* stbi_load(const char* filename, int* x, int* y, int* channels_in_file, int desired_channels);
extern "C" void stbi_image_free(void* retval_from_stbi_load);
