// Camera intrinsics holder with the reference's call syntax (Core/Utils/Intrinsics.h: Intrinsics::getInstance(fx, fy, cx, cy)
// once, Intrinsics::getInstance().fx() afterwards). Process-global; the first call that carries values fixes them, later
// arguments are ignored; reading before any values were given is a usage error and aborts with a message.
#ifndef EFUSION_B200_INTRINSICS_H_
#define EFUSION_B200_INTRINSICS_H_

#include <cstdio>
#include <cstdlib>

class Intrinsics {
  struct Values {
    float v[4];  // fx, fy, cx, cy
  };
  Values k_;
  explicit Intrinsics(const Values& k) : k_(k) {}

  static const Intrinsics& fixed(const Values& first) {
    if (!(first.v[0] != 0.f && first.v[1] != 0.f)) {
      std::fprintf(stderr, "Intrinsics::getInstance(): no focal lengths were given before the first use\n");
      std::abort();
    }
    static const Intrinsics theOne(first);
    return theOne;
  }

 public:
  static const Intrinsics& getInstance(float fx = 0.f, float fy = 0.f, float cx = 0.f, float cy = 0.f) {
    static const Intrinsics& ref = fixed(Values{{fx, fy, cx, cy}});
    return ref;
  }
  const float& fx() const { return k_.v[0]; }
  const float& fy() const { return k_.v[1]; }
  const float& cx() const { return k_.v[2]; }
  const float& cy() const { return k_.v[3]; }
};

#endif  // EFUSION_B200_INTRINSICS_H_
