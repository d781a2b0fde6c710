// Drop-in for the reference's Core/Utils/Resolution.h (process-global singleton: the first getInstance(w, h) wins,
// later arguments are ignored — reference Core/Utils/Resolution.cpp).
#ifndef EFUSION_B200_RESOLUTION_H_
#define EFUSION_B200_RESOLUTION_H_
#include <cassert>
class Resolution {
 public:
  static const Resolution& getInstance(int width = 0, int height = 0) {
    static const Resolution instance(width, height);
    return instance;
  }
  const int& width() const { return imgWidth; }
  const int& height() const { return imgHeight; }
  const int& cols() const { return imgWidth; }
  const int& rows() const { return imgHeight; }
  const int& numPixels() const { return imgNumPixels; }

 private:
  Resolution(int width, int height) : imgWidth(width), imgHeight(height), imgNumPixels(width * height) {
    assert(width > 0 && height > 0 && "You haven't initialised the Resolution class!");
  }
  const int imgWidth, imgHeight, imgNumPixels;
};
#endif
