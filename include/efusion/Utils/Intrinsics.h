// Drop-in for the reference's Core/Utils/Intrinsics.h (singleton, first call wins — Core/Utils/Intrinsics.cpp).
#ifndef EFUSION_B200_INTRINSICS_H_
#define EFUSION_B200_INTRINSICS_H_
#include <cassert>
class Intrinsics {
 public:
  static const Intrinsics& getInstance(float fx = 0, float fy = 0, float cx = 0, float cy = 0) {
    static const Intrinsics instance(fx, fy, cx, cy);
    return instance;
  }
  const float& fx() const { return fx_; }
  const float& fy() const { return fy_; }
  const float& cx() const { return cx_; }
  const float& cy() const { return cy_; }

 private:
  Intrinsics(float fx, float fy, float cx, float cy) : fx_(fx), fy_(fy), cx_(cx), cy_(cy) {
    assert(fx != 0 && fy != 0 && "You haven't initialised the Intrinsics class!");
  }
  const float fx_, fy_, cx_, cy_;
};
#endif
