// Image size holder with the reference's call syntax (Core/Utils/Resolution.h: Resolution::getInstance(w, h) once,
// Resolution::getInstance().width() / rows() / numPixels() afterwards). Process-global; the first call that carries a size
// fixes it, later arguments are ignored; reading before a size was given aborts with a message.
#ifndef EFUSION_B200_RESOLUTION_H_
#define EFUSION_B200_RESOLUTION_H_

#include <cstdio>
#include <cstdlib>

class Resolution {
  int dims_[3];  // columns, rows, columns * rows
  Resolution(int w, int h) : dims_{w, h, w * h} {}

  static const Resolution& fixed(int w, int h) {
    if (w <= 0 || h <= 0) {
      std::fprintf(stderr, "Resolution::getInstance(): no image size was given before the first use\n");
      std::abort();
    }
    static const Resolution theOne(w, h);
    return theOne;
  }

 public:
  static const Resolution& getInstance(int width = 0, int height = 0) {
    static const Resolution& ref = fixed(width, height);
    return ref;
  }
  const int& width() const { return dims_[0]; }
  const int& cols() const { return dims_[0]; }
  const int& height() const { return dims_[1]; }
  const int& rows() const { return dims_[1]; }
  const int& numPixels() const { return dims_[2]; }
};

#endif  // EFUSION_B200_RESOLUTION_H_
