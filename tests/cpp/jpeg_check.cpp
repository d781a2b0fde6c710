// Decodes a JPEG file with include/efusion/Tools/JPEGLoader.h and writes "W H\n" + raw RGB (libjpeg's JCS_RGB order) to stdout.
#include <cstdio>
#include <vector>

#include "Tools/JPEGLoader.h"

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  FILE* f = std::fopen(argv[1], "rb");
  if (!f) return 2;
  std::vector<uint8_t> buf;
  uint8_t tmp[65536];
  size_t n;
  while ((n = std::fread(tmp, 1, sizeof(tmp), f)) > 0) buf.insert(buf.end(), tmp, tmp + n);
  std::fclose(f);
  std::vector<uint8_t> rgb;
  int w = 0, h = 0;
  try {
    JPEGLoader::decode(buf.data(), buf.size(), rgb, w, h);
  } catch (const std::exception& e) {
    std::fprintf(stderr, "%s\n", e.what());
    return 1;
  }
  std::printf("%d %d\n", w, h);
  std::fwrite(rgb.data(), 1, rgb.size(), stdout);
  return 0;
}
