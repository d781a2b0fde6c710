// Ferns — randomised fern encoding of key frames for global loop closure / relocalisation candidates, for programs written
// against the reference's Core/Ferns.h:36-166 (same class, members and method signatures; GPUTexture* = a named device buffer of
// an EfContext). Own implementation of the behaviour of Core/Ferns.cpp:22-420:
//   * the conservatory: `num` ferns, each a pixel of the (W/8 x H/8) frame and four thresholds (R, G, B, depth in mm); a frame's
//     code per fern is 4 bits (Ferns.cpp:100-116), 255 where the vertex map has no depth;
//   * addFrame: Resize::image / vertex to W/8 x H/8 (ef_resize: nearest decimation on the device, only the 80x60 result crosses
//     PCIe), co-occurrence dissimilarity against every stored frame, insert if the most similar one is farther than `threshold`;
//   * findFrame: most similar frame older than 300 ticks, blockHDAware > 0.3, then the tiny geometric registration the
//     reference runs on its third RGBDOdometry instance -- ICP only (icpWeight 100), no pyramid, no SO(3), at W/8 x H/8 -- here on a
//     small EfContext of that size (ef_odom_init_icp_model / _pred / ef_odom_track), the photometric check and 50 sampled
//     surface constraints.
// The deformation solve that consumes the constraints (Deformation::constrain) is not part of this library.
// The reference seeds its generator with time(0); `seed` (default: time(0) as well) makes a run reproducible.
#ifndef EFUSION_B200_FERNS_H_
#define EFUSION_B200_FERNS_H_

#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <limits>
#include <random>
#include <vector>

#include "ElasticFusion.h"

class Ferns {
 public:
  Ferns(int n, int maxDepth, const float photoThresh, unsigned seed = (unsigned)time(0), int device = 0)
      : lastClosest(-1), badCode(255), num(n), factor(8), width(Resolution::getInstance().width() / factor),
        height(Resolution::getInstance().height() / factor), maxDepth(maxDepth), photoThresh(photoThresh), random(seed) {
    std::uniform_int_distribution<int32_t> widthDist(0, width - 1), heightDist(0, height - 1), rgbDist(0, 255), dDist(400, maxDepth);
    for (int i = 0; i < num; i++) {  // generateFerns, Ferns.cpp:64-79
      Fern f;
      f.pos[0] = widthDist(random);
      f.pos[1] = heightDist(random);
      f.rgbd[0] = rgbDist(random);
      f.rgbd[1] = rgbDist(random);
      f.rgbd[2] = rgbDist(random);
      f.rgbd[3] = dDist(random);
      conservatory.push_back(f);
    }
    EfConfig cfg;
    ef_default_config(&cfg, width, height, Intrinsics::getInstance().fx() / factor, Intrinsics::getInstance().fy() / factor,
                      Intrinsics::getInstance().cx() / factor, Intrinsics::getInstance().cy() / factor);
    cfg.capacity = 4096;  // the small context is only used for its tracker
    cfg.device = device;
    cfg.time_delta = std::numeric_limits<int>::max() / 2;
    ef::check(ef_create(&cfg, nullptr, &small_), "Ferns: ef_create");
  }
  virtual ~Ferns() {
    for (size_t i = 0; i < frames.size(); i++) delete frames.at(i);
    if (small_) ef_destroy(small_);
  }
  Ferns(const Ferns&) = delete;
  Ferns& operator=(const Ferns&) = delete;

  class SurfaceConstraint {
   public:
    SurfaceConstraint(const double* s, const double* t) {
      for (int k = 0; k < 4; ++k) {
        sourcePoint[k] = s[k];
        targetPoint[k] = t[k];
      }
    }
    double sourcePoint[4], targetPoint[4];
  };
  class Fern {
   public:
    int pos[2];
    int rgbd[4];
    std::vector<int> ids[16];
  };
  class Frame {
   public:
    Frame(int n, int id, const ef::SE3d& T_wc, int srcTime) : codes((size_t)n, 255), goodCodes(0), id(id), T_wc(T_wc), srcTime(srcTime) {}
    std::vector<uint8_t> codes;
    int goodCodes;
    const int id;
    ef::SE3d T_wc;
    const int srcTime;
    std::vector<uint8_t> initRgb;   // W/8 x H/8 x 3
    std::vector<float> initVerts;   // x 4
    std::vector<float> initNorms;   // x 4
  };

  // Ferns::addFrame, Ferns.cpp:81-158
  bool addFrame(GPUTexture* imageTexture, GPUTexture* vertexTexture, GPUTexture* normalTexture, const ef::SE3d& T_wc, int srcTime, const float threshold) {
    Small s;
    fetch(imageTexture, vertexTexture, normalTexture, s);
    Frame* frame = new Frame(num, (int)frames.size(), T_wc, srcTime);
    frame->initRgb = s.rgb;
    frame->initVerts = s.verts;
    frame->initNorms = s.norms;
    std::vector<int> coOccurrences(frames.size(), 0);
    encode(s, *frame, coOccurrences);
    float minimum = std::numeric_limits<float>::max();
    if (frame->goodCodes > 0)
      for (size_t i = 0; i < frames.size(); i++) {
        const float maxCo = (float)std::min(frame->goodCodes, frames.at(i)->goodCodes);
        const float dissim = (float)(maxCo - coOccurrences[i]) / (float)maxCo;
        if (dissim < minimum) minimum = dissim;
      }
    if ((minimum > threshold || frames.size() == 0) && frame->goodCodes > 0) {
      for (int i = 0; i < num; i++)
        if (frame->codes[i] != badCode) conservatory.at(i).ids[frame->codes[i]].push_back(frame->id);
      frames.push_back(frame);
      return true;
    }
    delete frame;
    return false;
  }

  // Ferns::findFrame, Ferns.cpp:160-298. Returns T_wc_est (identity when nothing matched); lastClosest != -1 on acceptance.
  ef::SE3d findFrame(std::vector<SurfaceConstraint>& constraints, const ef::SE3d& T_wc, GPUTexture* vertexTexture, GPUTexture* normalTexture,
                     GPUTexture* imageTexture, const int time, const bool lost) {
    lastClosest = -1;
    Small s;
    fetch(imageTexture, vertexTexture, normalTexture, s);
    Frame frame(num, 0, ef::SE3d(), 0);
    std::vector<int> coOccurrences(frames.size(), 0);
    encode(s, frame, coOccurrences);
    float minimum = std::numeric_limits<float>::max();
    int minId = -1;
    for (size_t i = 0; i < frames.size(); i++) {
      const float maxCo = (float)std::min(frame.goodCodes, frames.at(i)->goodCodes);
      const float dissim = (float)(maxCo - coOccurrences[i]) / (float)maxCo;
      if (dissim < minimum && time - frames.at(i)->srcTime > 300) {
        minimum = dissim;
        minId = (int)i;
      }
    }
    lastCandidate = minId;
    lastDissimilarity = minimum;
    ef::SE3d T_wc_est;
    if (minId != -1 && blockHDAware(&frame, frames.at(minId)) > 0.3) {
      const Frame& fern = *frames.at(minId);
      const ef::SE3d T_wc_fern = fern.T_wc;
      double Tf[16], Te[16], Tw[16];
      ef::toRowMajor(T_wc_fern, Tf);
      const size_t nb = (size_t)width * height * 16;
      ef::check(ef_upload(small_, EF_BUF_OLD_VERTEX, 0, fern.initVerts.data(), nb), "Ferns: upload");
      ef::check(ef_upload(small_, EF_BUF_OLD_NORMAL, 0, fern.initNorms.data(), nb), "Ferns: upload");
      ef::check(ef_upload(small_, EF_BUF_VERTEX, 0, s.verts.data(), nb), "Ferns: upload");
      ef::check(ef_upload(small_, EF_BUF_NORMAL, 0, s.norms.data(), nb), "Ferns: upload");
      // WARNING initICP* must be called before initRGB* (Ferns.cpp:232-238); the photometric half is commented out there
      ef::check(ef_odom_init_icp_model(small_, 0, (const float*)ptr(EF_BUF_OLD_VERTEX), (const float*)ptr(EF_BUF_OLD_NORMAL), Tf), "Ferns: initICPModel");
      ef::check(ef_odom_init_icp_pred(small_, 0, (const float*)ptr(EF_BUF_VERTEX), (const float*)ptr(EF_BUF_NORMAL)), "Ferns: initICP");
      std::memcpy(Te, Tf, sizeof(Te));
      ef::check(ef_odom_track(small_, 0, Te, 0, 100.0f, 0, 0, 0, nullptr, 0, nullptr), "Ferns: getIncrementalTransformation");
      T_wc_est = ef::fromRowMajor(Te);
      EfOdomStats st;
      ef::check(ef_odom_stats(small_, 0, &st), "Ferns: stats");
      lastICPError = st.lastICPError;
      lastICPCount = st.lastICPCount;
      const float photoError = photometricCheck(s, Te, Tf, fern.initRgb.data());
      lastPhotoError = photoError;
      const int icpCountThresh = lost ? 1400 : 2400;
      if (st.lastICPError < 0.0003 && st.lastICPCount > icpCountThresh && photoError < photoThresh) {
        lastClosest = minId;
        ef::toRowMajor(T_wc, Tw);
        for (int i = 0; i < num; i += num / 50) {
          const float* v = &s.verts[((size_t)conservatory.at(i).pos[1] * width + conservatory.at(i).pos[0]) * 4];
          if (v[2] > 0 && int(v[2] * 1000.0f) < maxDepth) {
            double raw[4] = {0, 0, 0, 1}, model[4] = {0, 0, 0, 1};
            for (int r = 0; r < 3; ++r) {
              raw[r] = Tw[r * 4 + 0] * v[0] + Tw[r * 4 + 1] * v[1] + Tw[r * 4 + 2] * v[2] + Tw[r * 4 + 3];
              model[r] = Te[r * 4 + 0] * v[0] + Te[r * 4 + 1] * v[1] + Te[r * 4 + 2] * v[2] + Te[r * 4 + 3];
            }
            constraints.push_back(SurfaceConstraint(raw, model));
          }
        }
      }
    }
    return T_wc_est;
  }

  std::vector<Fern> conservatory;
  std::vector<Frame*> frames;
  int lastClosest;
  const uint8_t badCode;
  float lastICPError = 0, lastICPCount = 0, lastPhotoError = 0;  // of the last registration (the reference prints them in a comment)
  int lastCandidate = -1;                                        // the frame findFrame tried to register against, accepted or not
  float lastDissimilarity = 0;

 private:
  struct Small {
    std::vector<uint8_t> rgb;
    std::vector<float> verts, norms;
  };
  void* ptr(int id) {
    void* p = nullptr;
    size_t n = 0;
    ef::check(ef_buffer(small_, id, 0, &p, &n), "Ferns: ef_buffer");
    return p;
  }
  void fetch(GPUTexture* image, GPUTexture* vertex, GPUTexture* normal, Small& s) {
    const size_t n = (size_t)width * height;
    std::vector<uint8_t> rgba(n * 4);
    s.rgb.resize(n * 3);
    s.verts.resize(n * 4);
    s.norms.resize(n * 4);
    ef::check(ef_resize(image->context(), image->id(), factor, rgba.data(), rgba.size()), "Ferns: Resize::image");
    ef::check(ef_resize(vertex->context(), vertex->id(), factor, s.verts.data(), n * 16), "Ferns: Resize::vertex");
    ef::check(ef_resize(normal->context(), normal->id(), factor, s.norms.data(), n * 16), "Ferns: Resize::vertex");
    for (size_t i = 0; i < n; ++i) {  // glReadPixels(GL_RGB) of the RGBA8 target
      s.rgb[i * 3 + 0] = rgba[i * 4 + 0];
      s.rgb[i * 3 + 1] = rgba[i * 4 + 1];
      s.rgb[i * 3 + 2] = rgba[i * 4 + 2];
    }
  }
  // codes + co-occurrences with the stored frames, Ferns.cpp:100-122 / 181-203
  void encode(const Small& s, Frame& frame, std::vector<int>& coOccurrences) {
    for (int i = 0; i < num; i++) {
      const Fern& f = conservatory.at(i);
      const size_t px = (size_t)f.pos[1] * width + f.pos[0];
      uint8_t code = badCode;
      const float z = s.verts[px * 4 + 2];
      if (z > 0) {
        const uint8_t* pix = &s.rgb[px * 3];
        code = (uint8_t)((pix[0] > f.rgbd[0]) << 3 | (pix[1] > f.rgbd[1]) << 2 | (pix[2] > f.rgbd[2]) << 1 | (int(z * 1000.0f) > f.rgbd[3]));
        frame.goodCodes++;
        for (size_t j = 0; j < f.ids[code].size(); j++) coOccurrences[f.ids[code].at(j)]++;
      }
      frame.codes[i] = code;
    }
  }
  // Ferns::photometricCheck, Ferns.cpp:300-368 (T_fern_est in float; projection truncated to int as Eigen::Vector2i does)
  float photometricCheck(const Small& s, const double* T_wc_est, const double* T_wc_fern, const uint8_t* fernRgb) {
    const float cx = Intrinsics::getInstance().cx() / factor, cy = Intrinsics::getInstance().cy() / factor;
    const float invfx = 1.0f / float(Intrinsics::getInstance().fx() / factor), invfy = 1.0f / float(Intrinsics::getInstance().fy() / factor);
    double inv[16], M[16];
    for (int r = 0; r < 3; ++r) {  // rigid inverse of T_wc_fern
      for (int c = 0; c < 3; ++c) inv[r * 4 + c] = T_wc_fern[c * 4 + r];
      inv[r * 4 + 3] = -(T_wc_fern[0 * 4 + r] * T_wc_fern[3] + T_wc_fern[1 * 4 + r] * T_wc_fern[7] + T_wc_fern[2 * 4 + r] * T_wc_fern[11]);
    }
    inv[12] = inv[13] = inv[14] = 0;
    inv[15] = 1;
    for (int r = 0; r < 4; ++r)
      for (int c = 0; c < 4; ++c) {
        double a = 0;
        for (int k = 0; k < 4; ++k) a += inv[r * 4 + k] * T_wc_est[k * 4 + c];
        M[r * 4 + c] = a;
      }
    float Mf[16];
    for (int k = 0; k < 16; ++k) Mf[k] = (float)M[k];
    float photoSum = 0;
    int photoCount = 0;
    for (int i = 0; i < num; i++) {
      const Fern& f = conservatory.at(i);
      const size_t px = (size_t)f.pos[1] * width + f.pos[0];
      const float* v = &s.verts[px * 4];
      if (v[2] > 0 && int(v[2] * 1000.0f) < maxDepth) {
        const float x = Mf[0] * v[0] + Mf[1] * v[1] + Mf[2] * v[2] + Mf[3], y = Mf[4] * v[0] + Mf[5] * v[1] + Mf[6] * v[2] + Mf[7],
                    z = Mf[8] * v[0] + Mf[9] * v[1] + Mf[10] * v[2] + Mf[11];
        const int u = (int)(x * (1 / invfx) / z + cx), w = (int)(y * (1 / invfy) / z + cy);
        if (u >= 0 && w >= 0 && u < width && w < height) {
          const uint8_t* q = &fernRgb[((size_t)w * width + u) * 3];
          if (q[0] > 0 || q[1] > 0 || q[2] > 0) {
            const uint8_t* p = &s.rgb[px * 3];
            photoSum += std::abs((int)q[0] - (int)p[0]);
            photoSum += std::abs((int)q[1] - (int)p[1]);
            photoSum += std::abs((int)q[2] - (int)p[2]);
            photoCount++;
          }
        }
      }
    }
    return photoSum / float(photoCount);
  }
  float blockHDAware(const Frame* f1, const Frame* f2) {  // Ferns.cpp:384-400
    int count = 0;
    float val = 0;
    for (int i = 0; i < num; i++)
      if (f1->codes[i] != badCode && f2->codes[i] != badCode) {
        count++;
        if (f1->codes[i] == f2->codes[i]) val += 1.0f;
      }
    return val / (float)count;
  }

  const int num, factor, width, height, maxDepth;
  const float photoThresh;
  std::mt19937 random;
  EfContext* small_ = nullptr;
};

#endif  // EFUSION_B200_FERNS_H_
