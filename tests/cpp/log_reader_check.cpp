// Drives include/efusion/Tools/RawLogReader.h the way MainController.cpp:216-245 drives the reference's reader and prints
// one checksum line per delivered frame. Usage: log_reader_check <file.klg> <w> <h> [peek] [flip]
#include <Tools/RawLogReader.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>

static unsigned long long fnv(const void* p, size_t n) {
  const unsigned char* b = (const unsigned char*)p;
  unsigned long long h = 1469598103934665603ull;
  for (size_t i = 0; i < n; ++i) h = (h ^ b[i]) * 1099511628211ull;
  return h;
}

int main(int argc, char** argv) {
  if (argc < 4) return 2;
  const int w = std::atoi(argv[2]), h = std::atoi(argv[3]);
  bool peek = false, flip = false;
  for (int i = 4; i < argc; ++i) {
    peek |= !std::strcmp(argv[i], "peek");
    flip |= !std::strcmp(argv[i], "flip");
  }
  Resolution::getInstance(w, h);
  RawLogReader reader(argv[1], flip);
  LogReader* log = &reader;  // the interface the reference's GUI holds
  std::printf("FRAMES %d\n", log->getNumFrames());
  unsigned long long expectRgb = 0, expectDepth = 0;
  bool havePeek = false;
  while (log->hasMore()) {
    log->getNext();
    const unsigned long long hr = fnv(log->rgb, (size_t)w * h * 3), hd = fnv(log->depth, (size_t)w * h * 2);
    if (havePeek && (hr != expectRgb || hd != expectDepth)) return 9;  // the peeked frame must be the delivered one
    std::printf("%d %lld %llu %llu\n", log->currentFrame, (long long)log->timestamp, hr, hd);
    havePeek = false;
    if (peek && (log->currentFrame % 2) && reader.peekNext()) {  // look ahead on every other frame
      expectRgb = fnv(reader.nextRgb(), (size_t)w * h * 3);
      expectDepth = fnv(reader.nextDepth(), (size_t)w * h * 2);
      havePeek = true;
    }
  }
  // rewind + fastForward + getBack behave like the reference's
  log->rewind();
  if (!log->rewound() || log->currentFrame != 0) return 10;
  log->fastForward(2);
  log->getNext();
  std::printf("FF %d %lld %llu\n", log->currentFrame, (long long)log->timestamp, fnv(log->depth, (size_t)w * h * 2));
  log->getBack();
  std::printf("BACK %d %lld %llu\n", log->currentFrame, (long long)log->timestamp, fnv(log->depth, (size_t)w * h * 2));
  return 0;
}
