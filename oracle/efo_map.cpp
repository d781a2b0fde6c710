// TEST INFRASTRUCTURE ONLY — CPU oracle, mapping half + depth preprocess.
// Restates the reference's GLSL passes (Core/Shaders/*.{vert,geom,frag,glsl}) with the host sequencing of
// Core/GlobalModel.cpp, Core/IndexMap.cpp, Core/Shaders/{FillIn,Resize,FeedbackBuffer,ComputePack}.cpp.
// PARITY PINNED against the reference's own shader files executed on Mesa 18 llvmpipe (oracle/gl/ref_gl_harness.cpp ->
// tests/golden/ref_mapping_160x120.npz, compared pass by pass in tests/test_gl_golden.py). GL semantics encoded here
// (the fixed-point snapping was learnt from that comparison):
//   * nearest sampling + clamp-to-edge: texel = clamp(floor(coord * size), 0, size-1) in fp32
//   * full-screen pass: fragment (i,j) <-> texel (i,j)
//   * window coordinates are snapped to 1/256 px (GL_SUBPIXEL_BITS = 8) before rasterisation
//   * 1-px points: the pixel whose centre lies in the half-open unit square around the snapped position (an integer window
//     coordinate x belongs to pixel x - 1); clipped when the centre leaves the clip volume
//   * point sprites: size clamped to [1, 2047] (NVIDIA; Mesa's limit is 255, not reached by the fixtures); covered pixels are
//     those whose centre lies in the half-open square of the snapped size around the snapped centre
//   * depth: window z quantised to 24 bits, round(z * (2^24-1)); GL_LESS; equal depth -> earlier primitive wins
//   * transform feedback = order-preserving compaction in draw order
//   * all shader arithmetic is IEEE fp32, no contraction (real GPUs use approximate rcp/rsqrt/exp — ulp-level)
//   * mat4*vec4 / mat3*vec3 accumulate left to right: ((m0*x + m1*y) + m2*z) + m3*w
//   * poses reach the shaders as float(double matrix) (the reference builds the matrix from the Sophus quaternion first,
//     GlobalModel.cpp:405 — a ~1e-16 relative difference before the float cast)
#include "ef_oracle.h"
#include "efo_common.h"
#include "efo_linalg.h"

#include <algorithm>
#include <cstdio>

using namespace efo;

namespace {

inline float gmin(float x, float y) { return (y < x) ? y : x; }  // GLSL min
inline float gmax(float x, float y) { return (x < y) ? y : x; }  // GLSL max

inline int texel(float coord, int n) {
  int i = (int)floorf(coord * (float)n);
  return i < 0 ? 0 : (i >= n ? n - 1 : i);
}

struct Cam {
  float cx, cy, fx, fy;
};

struct Pose {
  float m[16];  // row-major 4x4
};

inline Pose to_pose_f(const double* T) {
  Pose p;
  for (int i = 0; i < 16; ++i) p.m[i] = (float)T[i];
  return p;
}
inline Pose to_inv_pose_f(const double* T) {
  double inv[16];
  la::se3_inverse(T, inv);
  return to_pose_f(inv);
}
inline f3 xform(const Pose& p, const f3& v) {
  return mk3(((p.m[0] * v.x + p.m[1] * v.y) + p.m[2] * v.z) + p.m[3], ((p.m[4] * v.x + p.m[5] * v.y) + p.m[6] * v.z) + p.m[7],
             ((p.m[8] * v.x + p.m[9] * v.y) + p.m[10] * v.z) + p.m[11]);
}
inline f3 rot(const Pose& p, const f3& v) {
  return mk3((p.m[0] * v.x + p.m[1] * v.y) + p.m[2] * v.z, (p.m[4] * v.x + p.m[5] * v.y) + p.m[6] * v.z,
             (p.m[8] * v.x + p.m[9] * v.y) + p.m[10] * v.z);
}

// host-side uv buffer value (GlobalModel.cpp:109-117, FeedbackBuffer.cpp:45-53)
inline float uv_coord(int i, int n) { return (float)((double)((float)i / (float)n) + 1.0 / (double)(2 * (float)n)); }

// color.glsl:19-34
inline float encode_color_bytes(uint8_t r, uint8_t g, uint8_t b) {
  int rgb = (int)r;
  rgb = (rgb << 8) + (int)g;
  rgb = (rgb << 8) + (int)b;
  return (float)rgb;
}
inline float encode_color(const f3& c) {
  int rgb = (int)roundf(c.x * 255.0f);
  rgb = (rgb << 8) + (int)roundf(c.y * 255.0f);
  rgb = (rgb << 8) + (int)roundf(c.z * 255.0f);
  return (float)rgb;
}
inline f3 decode_color(float c) {
  int ci = (int)c;
  return mk3((float)(ci >> 16 & 0xFF) / 255.0f, (float)(ci >> 8 & 0xFF) / 255.0f, (float)(ci & 0xFF) / 255.0f);
}

// surfels.glsl:19-34 (cam.z, cam.w = 1/fx, 1/fy)
inline float get_radius(float depth, float norm_z, float inv_fx, float inv_fy) {
  float meanFocal = ((1.0f / fabsf(inv_fx)) + (1.0f / fabsf(inv_fy))) / 2.0f;
  const float sqrt2 = 1.41421356237f;
  float radius = (depth / meanFocal) * sqrt2;
  float radius_n = radius;
  radius_n = radius_n / fabsf(norm_z);
  radius_n = gmin(2.0f * radius, radius_n);
  return radius_n;
}
// surfels.glsl:36-46
inline float confidence(float x, float y, float weighting, float cx, float cy) {
  const float maxRadDist = 400;
  const float twoSigmaSquared = 0.72f;
  float px = x - cx, py = y - cy;
  float radialDist = sqrtf(px * px + py * py) / maxRadDist;
  return expf((-(radialDist * radialDist) / twoSigmaSquared)) * weighting;
}

// geometry.glsl:21-40, float depth sampler. (x,y) are float pixel coords, (ix,iy) the texel the texcoord hits.
inline f3 vertex_f(const float* depth, int rows, int cols, int ix, int iy, float x, float y, const Cam& c, float ifx, float ify) {
  ix = ix < 0 ? 0 : (ix >= cols ? cols - 1 : ix);
  iy = iy < 0 ? 0 : (iy >= rows ? rows - 1 : iy);
  float z = depth[iy * cols + ix];
  return mk3((x - c.cx) * z * ifx, (y - c.cy) * z * ify, z);
}
inline f3 normal_central(const float* depth, int rows, int cols, int ix, int iy, float x, float y, const f3& vPos,
                         const Cam& c, float ifx, float ify) {
  f3 xf = vertex_f(depth, rows, cols, ix + 1, iy, x + 1, y, c, ifx, ify);
  f3 xb = vertex_f(depth, rows, cols, ix - 1, iy, x - 1, y, c, ifx, ify);
  f3 yf = vertex_f(depth, rows, cols, ix, iy + 1, x, y + 1, c, ifx, ify);
  f3 yb = vertex_f(depth, rows, cols, ix, iy - 1, x, y - 1, c, ifx, ify);
  f3 del_x = mk3((xb.x + vPos.x) / 2 - (xf.x + vPos.x) / 2, (xb.y + vPos.y) / 2 - (xf.y + vPos.y) / 2,
                 (xb.z + vPos.z) / 2 - (xf.z + vPos.z) / 2);
  f3 del_y = mk3((yb.x + vPos.x) / 2 - (yf.x + vPos.x) / 2, (yb.y + vPos.y) / 2 - (yf.y + vPos.y) / 2,
                 (yb.z + vPos.z) / 2 - (yf.z + vPos.z) / 2);
  return normalized(cross(del_x, del_y));
}
// geometry.glsl:42-60, ushort mm sampler, integer pixel coords, forward difference
inline f3 vertex_u(const uint16_t* depth, int rows, int cols, int ix, int iy, int x, int y, const Cam& c, float ifx, float ify) {
  ix = ix < 0 ? 0 : (ix >= cols ? cols - 1 : ix);
  iy = iy < 0 ? 0 : (iy >= rows ? rows - 1 : iy);
  float z = (float)depth[iy * cols + ix] / 1000.0f;
  return mk3(((float)x - c.cx) * z * ifx, ((float)y - c.cy) * z * ify, z);
}

// fixed-point window coordinate of a point / sprite centre: round((w - 0.5) * 256), the pixel-centre offset removed
inline int snap256(float w) { return (int)lrintf((w - 0.5f) * 256.0f); }
// the pixel whose centre lies in the half-open unit square around the snapped coordinate
inline int point_pixel(float w) { return (snap256(w) + 127) >> 8; }
// pixel range [p0, p1] whose centres lie in the half-open square of side `size` around the snapped centre
inline void sprite_range(float w, float size, int& p0, int& p1) {
  int fw = (int)lrintf(size * 256.0f);
  if (fw < 256) fw = 256;
  const int x0 = snap256(w) - fw / 2;
  p0 = (x0 + 255) >> 8;
  p1 = ((x0 + fw + 255) >> 8) - 1;
}
inline uint32_t depth24(float zw) {
  if (!(zw > 0.f)) zw = 0.f;
  if (zw > 1.f) zw = 1.f;
  return (uint32_t)rintf(zw * 16777215.0f);
}

inline void atomic_min_u64(uint64_t* addr, uint64_t v) {
  uint64_t old = __atomic_load_n(addr, __ATOMIC_RELAXED);
  while (v < old && !__atomic_compare_exchange_n(addr, &old, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {
  }
}

const uint64_t kEmptyKey = ~0ull;

}  // namespace

// ---------------------------------------------------------------------------------------------
// depth preprocess
// ---------------------------------------------------------------------------------------------

// depth_bilateral.frag:30-75 via ElasticFusion::filterDepth (ElasticFusion.cpp:665-673)
extern "C" void efo_bilateral(const uint16_t* depth, int rows, int cols, float max_d, uint16_t* out) {
  const uint32_t hi = (uint32_t)(max_d * 1000.0f);
#pragma omp parallel for schedule(dynamic, 4)
  for (int y = 0; y < rows; ++y)
    for (int x = 0; x < cols; ++x) {
      uint32_t value = depth[y * cols + x];
      if (value > hi || value < 300U) {
        out[y * cols + x] = 0;
        continue;
      }
      const float sigma_space2_inv_half = 0.024691358f;
      const float sigma_color2_inv_half = 0.000555556f;
      const int R = 6, D = R * 2 + 1;
      int tx = imin(x - D / 2 + D, cols);
      int ty = imin(y - D / 2 + D, rows);
      float sum1 = 0, sum2 = 0;
      for (int cy = imax(y - D / 2, 0); cy < ty; ++cy)
        for (int cx = imax(x - D / 2, 0); cx < tx; ++cx) {
          // texture(gSampler, vec2(cx/cols, cy/rows)) — texel edge, resolves to texel (cx,cy) (App. A-27)
          uint32_t tmp = depth[cy * cols + cx];
          float space2 = ((float)x - (float)cx) * ((float)x - (float)cx) + ((float)y - (float)cy) * ((float)y - (float)cy);
          float color2 = ((float)value - (float)tmp) * ((float)value - (float)tmp);
          float weight = expf(-(space2 * sigma_space2_inv_half + color2 * sigma_color2_inv_half));
          sum1 += (float)tmp * weight;
          sum2 += weight;
        }
      out[y * cols + x] = (uint16_t)(uint32_t)roundf(sum1 / sum2);
    }
}

// depth_metric.frag:28-39
extern "C" void efo_metric(const uint16_t* depth, int rows, int cols, float max_d, float* out) {
  const uint32_t hi = (uint32_t)(max_d * 1000.0f);
  for (int i = 0; i < rows * cols; ++i) {
    uint32_t value = depth[i];
    out[i] = (value > hi || value < 300U) ? 0.f : (float)value / 1000.0f;
  }
}

// ---------------------------------------------------------------------------------------------
// first-frame map initialisation
// ---------------------------------------------------------------------------------------------

// vertex_feedback.vert/.geom via FeedbackBuffer::compute (FeedbackBuffer.cpp:81-138). cam4 = cx,cy,fx,fy.
// Emits in uv-buffer order (x-major: for i<width, for j<height). Returns the number of surfels written.
extern "C" int efo_feedback_buffer(const uint8_t* rgb, const float* depth, int rows, int cols, const float* cam4,
                                   int time, float max_depth, float* out) {
  Cam c{cam4[0], cam4[1], cam4[2], cam4[3]};
  const float ifx = 1.0f / c.fx, ify = 1.0f / c.fy;
  int n = 0;
  for (int i = 0; i < cols; ++i)
    for (int j = 0; j < rows; ++j) {
      float tcx = uv_coord(i, cols), tcy = uv_coord(j, rows);
      float x = tcx * (float)cols, y = tcy * (float)rows;
      f3 vPos = vertex_f(depth, rows, cols, i, j, x, y, c, ifx, ify);
      f3 nrm = normal_central(depth, rows, cols, i, j, x, y, vPos, c, ifx, ify);
      float rad = get_radius(vPos.z, nrm.z, ifx, ify);
      if (vPos.z <= 0 || vPos.z > max_depth) continue;
      const uint8_t* px = rgb + (size_t)(j * cols + i) * 3;
      float* s = out + (size_t)n * 12;
      s[0] = vPos.x;
      s[1] = vPos.y;
      s[2] = vPos.z;
      s[3] = confidence(x, y, 1.0f, c.cx, c.cy);
      s[4] = encode_color_bytes(px[0], px[1], px[2]);
      s[5] = 0;
      s[6] = (float)px[2] / 255.0f;  // vColor.z keeps the blue channel (vertex_feedback.vert:40-65)
      s[7] = (float)time;
      s[8] = nrm.x;
      s[9] = nrm.y;
      s[10] = nrm.z;
      s[11] = rad;
      ++n;
    }
  return n;
}

// init_unstable.vert via GlobalModel::initialise (GlobalModel.cpp:229-284): attributes 0,1 from the RAW feedback
// buffer, attribute 2 from the FILTERED one, paired by index; raw_count vertices are drawn (App. A-29).
extern "C" int efo_map_initialise(const float* raw_fb, int raw_count, const float* filt_fb, int filt_count,
                                  int cap_pixels, float* map) {
  (void)cap_pixels;
  for (int k = 0; k < raw_count; ++k) {
    float* s = map + (size_t)k * 12;
    const float* r = raw_fb + (size_t)k * 12;
    for (int q = 0; q < 8; ++q) s[q] = r[q];
    s[5] = 0;
    s[6] = 1;
    if (k < filt_count) {
      const float* f = filt_fb + (size_t)k * 12;
      for (int q = 8; q < 12; ++q) s[q] = f[q];
    } else {
      s[8] = s[9] = s[10] = s[11] = 0;  // zero-initialised VBO tail (FeedbackBuffer.cpp:30-38)
    }
  }
  return raw_count;
}

// ---------------------------------------------------------------------------------------------
// index map (1-px point splat)
// ---------------------------------------------------------------------------------------------

// index_map.vert/.frag via IndexMap::predictIndices (IndexMap.cpp:190-258)
extern "C" void efo_predict_indices(const float* map, int count, const double* T_wc, int time, float max_depth,
                                    int time_delta, int rows, int cols, const float* cam4, uint32_t* index,
                                    float* vert_conf4, float* color_time4, float* norm_rad4) {
  Cam c{cam4[0], cam4[1], cam4[2], cam4[3]};
  const Pose t_inv = to_inv_pose_f(T_wc);
  const size_t n = (size_t)rows * cols;
  std::vector<uint64_t> zbuf(n, kEmptyKey);
  const float fcols = (float)cols, frows = (float)rows;
#pragma omp parallel for schedule(static)
  for (int id = 0; id < count; ++id) {
    const float* s = map + (size_t)id * 12;
    f3 h = xform(t_inv, mk3(s[0], s[1], s[2]));
    if (h.z > max_depth || h.z < 0 || (float)time - s[7] > (float)time_delta) continue;
    float xn = ((((c.fx * h.x) / h.z) + c.cx) - (fcols * 0.5f)) / (fcols * 0.5f);
    float yn = ((((c.fy * h.y) / h.z) + c.cy) - (frows * 0.5f)) / (frows * 0.5f);
    float zn = h.z / max_depth;
    if (!(xn >= -1.f && xn <= 1.f && yn >= -1.f && yn <= 1.f && zn >= -1.f && zn <= 1.f)) continue;
    float xw = (xn + 1.0f) * (fcols * 0.5f);
    float yw = (yn + 1.0f) * (frows * 0.5f);
    // Window coordinates are snapped to 1/256 pixel before rasterisation (GL_SUBPIXEL_BITS = 8 on Mesa llvmpipe and on NVIDIA
    // hardware), and a size-1 point is the half-open square [xs - 0.5, xs + 0.5): pinned against the reference's shaders
    // run on Mesa (tests/golden/ref_mapping_*.npz).
    const int px = point_pixel(xw), py = point_pixel(yw);
    if (px < 0 || py < 0 || px >= cols || py >= rows) continue;
    uint32_t d24 = depth24(0.5f * zn + 0.5f);
    if (d24 >= 16777215u) continue;  // GL_LESS against the cleared depth 1.0
    uint64_t key = ((uint64_t)d24 << 32) | (uint32_t)id;
    atomic_min_u64(&zbuf[(size_t)py * cols + px], key);
  }
#pragma omp parallel for schedule(static)
  for (int p = 0; p < (int)n; ++p) {
    float* vc = vert_conf4 + (size_t)p * 4;
    float* ct = color_time4 + (size_t)p * 4;
    float* nr = norm_rad4 + (size_t)p * 4;
    if (zbuf[p] == kEmptyKey) {
      index[p] = 0;
      for (int q = 0; q < 4; ++q) vc[q] = ct[q] = nr[q] = 0.f;
      continue;
    }
    uint32_t id = (uint32_t)(zbuf[p] & 0xffffffffu);
    const float* s = map + (size_t)id * 12;
    f3 h = xform(t_inv, mk3(s[0], s[1], s[2]));
    f3 nn = normalized(rot(t_inv, mk3(s[8], s[9], s[10])));
    index[p] = id;
    vc[0] = h.x;
    vc[1] = h.y;
    vc[2] = h.z;
    vc[3] = s[3];
    ct[0] = s[4];
    ct[1] = s[5];
    ct[2] = s[6];
    ct[3] = s[7];
    nr[0] = nn.x;
    nr[1] = nn.y;
    nr[2] = nn.z;
    nr[3] = s[11];
  }
}

// ---------------------------------------------------------------------------------------------
// fuse: data association (data.vert/.geom/.frag) + merge (update.vert)
// ---------------------------------------------------------------------------------------------

namespace {
struct Meas {
  float pos[4], col[4], nr[4];
  f3 vPosLocal, vNormLocal;
};
}  // namespace

extern "C" int efo_fuse(float* map, int count, const double* T_wc, int time, const uint8_t* rgb,
                        const float* depth_raw, const float* depth_filt, const uint32_t* index,
                        const float* vert_conf4, const float* color_time4, const float* norm_rad4, float max_depth,
                        float weighting, int rows, int cols, const float* cam4, float* new_unstable) {
  (void)color_time4;
  Cam c{cam4[0], cam4[1], cam4[2], cam4[3]};
  const float ifx = (float)(1.0 / (double)c.fx), ify = (float)(1.0 / (double)c.fy);  // GlobalModel.cpp:396-401
  const Pose pose = to_pose_f(T_wc);
  const float fcols = (float)cols, frows = (float)rows;
  const float ftime = (float)time;
  const size_t n = (size_t)rows * cols;
  // per draw index d = i*rows + j: 0 = nothing, 1 = matched (best id), 2 = new unstable
  std::vector<uint8_t> kind(n, 0);
  std::vector<uint32_t> best_id(n, 0);
  std::vector<Meas> meas(n);
  std::vector<uint32_t> winner((size_t)count + 1, 0xffffffffu);

#pragma omp parallel for schedule(dynamic, 8)
  for (int i = 0; i < cols; ++i)
    for (int j = 0; j < rows; ++j) {
      const uint32_t d = (uint32_t)i * rows + j;
      float tcx = uv_coord(i, cols), tcy = uv_coord(j, rows);
      float x = tcx * fcols, y = tcy * frows;
      if (!((int)x % 2 == (int)ftime % 2 && (int)y % 2 == (int)ftime % 2)) continue;
      f3 vPosLocal = vertex_f(depth_raw, rows, cols, i, j, x, y, c, ifx, ify);
      // checkNeighbours (data.vert:50-69), clamp-to-edge
      auto dr = [&](int ii, int jj) {
        ii = ii < 0 ? 0 : (ii >= cols ? cols - 1 : ii);
        jj = jj < 0 ? 0 : (jj >= rows ? rows - 1 : jj);
        return depth_raw[jj * cols + ii];
      };
      if (dr(i - 1, j) == 0 || dr(i, j - 1) == 0 || dr(i + 1, j) == 0 || dr(i, j + 1) == 0) continue;
      if (!(vPosLocal.z > 0 && vPosLocal.z <= max_depth)) continue;

      Meas& m = meas[d];
      f3 vg = xform(pose, vPosLocal);
      f3 vPos_f = vertex_f(depth_filt, rows, cols, i, j, x, y, c, ifx, ify);
      f3 vNormLocal = normal_central(depth_filt, rows, cols, i, j, x, y, vPos_f, c, ifx, ify);
      f3 ng = rot(pose, vNormLocal);
      const uint8_t* px = rgb + (size_t)(j * cols + i) * 3;
      m.pos[0] = vg.x;
      m.pos[1] = vg.y;
      m.pos[2] = vg.z;
      m.pos[3] = confidence(x, y, weighting, c.cx, c.cy);
      m.col[0] = encode_color_bytes(px[0], px[1], px[2]);
      m.col[1] = 0;
      m.col[2] = ftime;
      m.col[3] = 0;
      m.nr[0] = ng.x;
      m.nr[1] = ng.y;
      m.nr[2] = ng.z;
      m.nr[3] = get_radius(vPos_f.z, vNormLocal.z, ifx, ify);

      int counter = 0;
      uint32_t best = 0;
      const float scale = 1.0f;  // IndexMap::FACTOR
      float indexXStep = (1.0f / (fcols * scale)) * 0.5f;
      float indexYStep = (1.0f / (frows * scale)) * 0.5f;
      float bestDist = 1000;
      const float windowMultiplier = 2;
      float xl = (x - c.cx) * ifx;
      float yl = (y - c.cy) * ify;
      float lambda = sqrtf(xl * xl + yl * yl + 1);
      f3 ray = mk3(xl, yl, 1);
      for (float ii = tcx - (scale * indexXStep * windowMultiplier); ii < tcx + (scale * indexXStep * windowMultiplier); ii += indexXStep)
        for (float jj = tcy - (scale * indexYStep * windowMultiplier); jj < tcy + (scale * indexYStep * windowMultiplier); jj += indexYStep) {
          int p = texel(jj, rows) * cols + texel(ii, cols);
          uint32_t current = index[p];
          if (current > 0U) {
            const float* vc = vert_conf4 + (size_t)p * 4;
            if (fabsf((vc[2] * lambda) - (vPosLocal.z * lambda)) < 0.05f) {
              f3 vcp = mk3(vc[0], vc[1], vc[2]);
              float dist = norm(cross(ray, vcp)) / norm(ray);
              const float* nr = norm_rad4 + (size_t)p * 4;
              f3 nrm = mk3(nr[0], nr[1], nr[2]);
              float ang = acosf(dot(nrm, vNormLocal) / (norm(nrm) * norm(vNormLocal)));
              if (dist < bestDist && (fabsf(nr[2]) < 0.75f || fabsf(ang) < 0.5f)) {
                counter++;
                bestDist = dist;
                best = current;
              }
            }
          }
        }
      if (counter > 0) {
        kind[d] = 1;
        best_id[d] = best;
        m.col[3] = -1;
      } else {
        kind[d] = 2;
        m.col[3] = -2;
      }
    }

  // update-map rasterisation: same texel, equal depth, GL_LESS -> lowest draw index wins (App. A-17)
  for (size_t d = 0; d < n; ++d)
    if (kind[d] == 1 && best_id[d] < (uint32_t)count && (uint32_t)d < winner[best_id[d]]) winner[best_id[d]] = (uint32_t)d;

  // update.vert:36-92 — only surfels with an update texel (newColor.w == -1) change
#pragma omp parallel for schedule(static)
  for (int id = 0; id < count; ++id) {
    if (winner[id] == 0xffffffffu) continue;
    const Meas& m = meas[winner[id]];
    float* s = map + (size_t)id * 12;
    float c_k = s[3];
    float a = m.pos[3];
    if (m.nr[3] < (1.0f + 0.5f) * s[11]) {
      float ck_a = c_k + a;
      s[0] = ((c_k * s[0]) + (a * m.pos[0])) / ck_a;
      s[1] = ((c_k * s[1]) + (a * m.pos[1])) / ck_a;
      s[2] = ((c_k * s[2]) + (a * m.pos[2])) / ck_a;
      s[3] = ck_a;
      f3 oldCol = decode_color(s[4]);
      f3 newCol = decode_color(m.col[0]);
      f3 avg = mk3(((c_k * oldCol.x) + (a * newCol.x)) / ck_a, ((c_k * oldCol.y) + (a * newCol.y)) / ck_a,
                   ((c_k * oldCol.z) + (a * newCol.z)) / ck_a);
      s[4] = encode_color(avg);
      s[7] = (float)time;
      f3 nn = mk3(((c_k * s[8]) + (a * m.nr[0])) / ck_a, ((c_k * s[9]) + (a * m.nr[1])) / ck_a,
                  ((c_k * s[10]) + (a * m.nr[2])) / ck_a);
      float rr = ((c_k * s[11]) + (a * m.nr[3])) / ck_a;
      nn = normalized(nn);
      s[8] = nn.x;
      s[9] = nn.y;
      s[10] = nn.z;
      s[11] = rr;
    } else {
      s[3] = c_k + a;
      s[7] = (float)time;
    }
  }

  // transform feedback of the data pass, draw order; matched points (w == -1) are captured too but are culled
  // as "degenerate" by copy_unstable.vert:120 — only the new (w == -2) ones are kept here.
  int nn = 0;
  for (size_t d = 0; d < n; ++d)
    if (kind[d] == 2) {
      float* o = new_unstable + (size_t)nn * 12;
      memcpy(o, meas[d].pos, 16);
      memcpy(o + 4, meas[d].col, 16);
      memcpy(o + 8, meas[d].nr, 16);
      ++nn;
    }
  return nn;
}

// ---------------------------------------------------------------------------------------------
// clean: copy_unstable.vert/.geom (no deformation graph: nodes == 0)
// ---------------------------------------------------------------------------------------------

namespace {
inline bool clean_test(float* s /* may update s[7] */, const Pose& t_inv, const Cam& c, int time, const uint32_t* index,
                       const float* vert_conf4, const float* color_time4, float conf_threshold, int time_delta,
                       int rows, int cols) {
  const float fcols = (float)cols, frows = (float)rows;
  int test = 1;
  f3 localPos = xform(t_inv, mk3(s[0], s[1], s[2]));
  float x = ((c.fx * localPos.x) / localPos.z) + c.cx;
  float y = ((c.fy * localPos.y) / localPos.z) + c.cy;
  f3 localNorm = normalized(rot(t_inv, mk3(s[8], s[9], s[10])));
  const float scale = 1.0f;
  float indexXStep = (1.0f / (fcols * scale)) * 0.5f;
  float indexYStep = (1.0f / (frows * scale)) * 0.5f;
  const float windowMultiplier = 2;
  int count = 0, zCount = 0;
  if ((float)time - s[7] < (float)time_delta && localPos.z > 0 && x > 0 && y > 0 && x < fcols && y < frows) {
    for (float i = x / fcols - (scale * indexXStep * windowMultiplier); i < x / fcols + (scale * indexXStep * windowMultiplier); i += indexXStep)
      for (float j = y / frows - (scale * indexYStep * windowMultiplier); j < y / frows + (scale * indexYStep * windowMultiplier); j += indexYStep) {
        int p = texel(j, rows) * cols + texel(i, cols);
        uint32_t current = index[p];
        if (current > 0U) {
          const float* vc = vert_conf4 + (size_t)p * 4;
          const float* ct = color_time4 + (size_t)p * 4;
          float dx = vc[0] - localPos.x, dy = vc[1] - localPos.y;
          if (ct[2] < s[6] && vc[3] > conf_threshold && vc[2] > localPos.z && vc[2] - localPos.z < 0.01f &&
              sqrtf(dx * dx + dy * dy) < s[11] * 1.4f)
            count++;
          if (ct[3] == (float)time && vc[3] > conf_threshold && vc[2] > localPos.z && vc[2] - localPos.z > 0.01f &&
              fabsf(localNorm.z) > 0.85f)
            zCount++;
        }
      }
  }
  if (count > 8 || zCount > 4) test = 0;
  if (s[7] == -2) s[7] = (float)time;
  if (s[7] == -1 || (((float)time - s[7]) > 20 && s[3] < conf_threshold)) test = 0;
  if (s[7] > 0 && (float)time - s[7] > (float)time_delta) test = 1;
  return test > 0;
}
}  // namespace

// copy_unstable.vert:132-322 — the deformation-graph branch of clean: the surfel is moved by the weighted rigid motions of
// the k = 4 nearest of <= 20 temporally neighbouring graph nodes (binary search on the node time stamps, 10 back / up to 20 total),
// its normal by the inverse-transpose rotations, and its lastTime is refreshed when the moved surfel is stable and lies in
// front of (or within 10 cm behind) the synthesised model depth at its new projection. nodes: 16 floats each — position 3,
// rotation 9 (column-major, Eigen storage order), translation 3, time (Deformation.cpp:175-189; the node texture is sampled
// at texel centres with nearest filtering, i.e. plain array reads). GLSL pow(x, 2) is evaluated as x * x (x >= 0 here).
static void deform_surfel(float* s, const float* nodes, int n_nodes, const Pose& t_inv, const Cam& c, int time, float conf_threshold,
                          float max_depth, int is_fern, const float* depth, int rows, int cols) {
  const int k = 4, lookBack = 20;
  int nearNodes[lookBack];
  float nearDists[lookBack];
  for (int i = 0; i < lookBack; i++) {
    nearNodes[i] = -1;
    nearDists[i] = 16777216.0f;
  }
  auto ntime = [&](int i) { return (int)nodes[(size_t)i * 16 + 15]; };
  auto npos = [&](int i) { return mk3(nodes[(size_t)i * 16 + 0], nodes[(size_t)i * 16 + 1], nodes[(size_t)i * 16 + 2]); };
  const int poseTime = (int)s[6];
  int foundIndex = 0;
  int imin = 0, imax = n_nodes - 1, imid = (imin + imax) / 2;
  while (imax >= imin) {
    imid = (imin + imax) / 2;
    const int nodeTime = ntime(imid);
    if (nodeTime < poseTime)
      imin = imid + 1;
    else if (nodeTime > poseTime)
      imax = imid - 1;
    else
      break;
  }
  imin = std::min(imin, n_nodes - 1);
  // imax can reach -1 (every node is later than the surfel): the texture fetch clamps to texel 0 (CLAMP_TO_EDGE)
  const int nodeMin = ntime(imin), nodeMid = ntime(imid), nodeMax = ntime(std::max(imax, 0));
  if (std::abs(nodeMin - poseTime) <= std::abs(nodeMid - poseTime) && std::abs(nodeMin - poseTime) <= std::abs(nodeMax - poseTime))
    foundIndex = imin;
  else if (std::abs(nodeMid - poseTime) <= std::abs(nodeMin - poseTime) && std::abs(nodeMid - poseTime) <= std::abs(nodeMax - poseTime))
    foundIndex = imid;
  else
    foundIndex = imax;
  if (foundIndex == n_nodes) foundIndex = n_nodes - 1;
  const f3 v = mk3(s[0], s[1], s[2]);
  int nearNodeIndex = 0, distanceBack = 0;
  for (int j = foundIndex; j >= 0; j--) {
    const f3 d = v - npos(j);
    nearNodes[nearNodeIndex] = j;
    nearDists[nearNodeIndex] = sqrtf(dot(d, d));
    nearNodeIndex++;
    if (++distanceBack == lookBack / 2) break;
  }
  for (int j = foundIndex + 1; j < n_nodes; j++) {
    const f3 d = v - npos(j);
    nearNodes[nearNodeIndex] = j;
    nearDists[nearNodeIndex] = sqrtf(dot(d, d));
    nearNodeIndex++;
    if (++distanceBack == lookBack) break;
  }
  for (int i = 0; i < lookBack - 1; ++i)
    for (int j = i + 1; j < lookBack; ++j)
      if (nearDists[j] < nearDists[i]) {
        std::swap(nearDists[i], nearDists[j]);
        std::swap(nearNodes[i], nearNodes[j]);
      }
  const float dMax = nearDists[k];
  float w[k], wsum = 0;
  for (int j = 0; j < k; j++) {
    const f3 d = v - npos(std::max(nearNodes[j], 0));
    const float b = 1.0f - (sqrtf(dot(d, d)) / dMax);
    w[j] = b * b;
    wsum += w[j];
  }
  for (int j = 0; j < k; j++) w[j] /= wsum;
  f3 newPos = mk3(0, 0, 0), newNorm = mk3(0, 0, 0);
  const f3 nrm = mk3(s[8], s[9], s[10]);
  for (int i = 0; i < k; i++) {
    const float* nd = nodes + (size_t)std::max(nearNodes[i], 0) * 16;
    const f3 position = mk3(nd[0], nd[1], nd[2]);
    const f3 c0 = mk3(nd[3], nd[4], nd[5]), c1 = mk3(nd[6], nd[7], nd[8]), c2 = mk3(nd[9], nd[10], nd[11]);  // columns
    const f3 translation = mk3(nd[12], nd[13], nd[14]);
    const f3 d = v - position;
    const f3 rd = mk3(c0.x * d.x + c1.x * d.y + c2.x * d.z, c0.y * d.x + c1.y * d.y + c2.y * d.z, c0.z * d.x + c1.z * d.y + c2.z * d.z);
    newPos = newPos + ((rd + position) + translation) * w[i];
    // transpose(inverse(R)) = cofactor(R) / det(R): its columns are the cross products of R's columns
    const f3 k0 = cross(c1, c2), k1 = cross(c2, c0), k2 = cross(c0, c1);
    const float det = dot(c0, k0);
    const f3 tn = mk3((k0.x * nrm.x + k1.x * nrm.y + k2.x * nrm.z) / det, (k0.y * nrm.x + k1.y * nrm.y + k2.y * nrm.z) / det,
                      (k0.z * nrm.x + k1.z * nrm.y + k2.z * nrm.z) / det);
    newNorm = newNorm + tn * w[i];
  }
  s[0] = newPos.x;
  s[1] = newPos.y;
  s[2] = newPos.z;
  const f3 nn = normalized(newNorm);
  s[8] = nn.x;
  s[9] = nn.y;
  s[10] = nn.z;
  if (s[3] > conf_threshold && is_fern == 0) {
    const f3 lp = xform(t_inv, newPos);
    const float x = ((c.fx * lp.x) / lp.z) + c.cx, y = ((c.fy * lp.y) / lp.z) + c.cy;
    if (lp.z > 0 && lp.z < max_depth && x > 0 && y > 0 && x < (float)cols && y < (float)rows) {
      const float currentDepth = depth[(size_t)texel(y / (float)rows, rows) * cols + texel(x / (float)cols, cols)];
      if (currentDepth > 0.0f && lp.z < currentDepth + 0.1f) s[7] = (float)time;
    }
  }
}

// clean with a deformation graph (GlobalModel.cpp:527-671 with graph.size() > 0). depth = IndexMap::depthTex()
// (synthesizeDepth, ElasticFusion.cpp:559-569). nodes == nullptr / n_nodes == 0 reduces to efo_clean.
extern "C" int efo_clean_deform(const float* map, int count, const float* new_unstable, int new_count, const double* T_wc, int time,
                                const uint32_t* index, const float* vert_conf4, const float* color_time4, float conf_threshold,
                                int time_delta, float max_depth, int rows, int cols, const float* cam4, const float* nodes, int n_nodes,
                                const float* depth, int is_fern, float* out) {
  Cam c{cam4[0], cam4[1], cam4[2], cam4[3]};
  const Pose t_inv = to_inv_pose_f(T_wc);
  const int total = count + new_count;
  std::vector<uint8_t> keep(total, 0);
  std::vector<float> tmp((size_t)total * 12);
#pragma omp parallel for schedule(static)
  for (int k = 0; k < total; ++k) {
    float* s = tmp.data() + (size_t)k * 12;
    memcpy(s, (k < count) ? map + (size_t)k * 12 : new_unstable + (size_t)(k - count) * 12, 48);
    const bool t = clean_test(s, t_inv, c, time, index, vert_conf4, color_time4, conf_threshold, time_delta, rows, cols);
    keep[k] = t ? 1 : 0;
    if (t && n_nodes > 0 && s[6] != (float)time) deform_surfel(s, nodes, n_nodes, t_inv, c, time, conf_threshold, max_depth, is_fern, depth, rows, cols);
  }
  int n = 0;
  for (int k = 0; k < total; ++k)
    if (keep[k]) {
      memcpy(out + (size_t)n * 12, tmp.data() + (size_t)k * 12, 48);
      ++n;
    }
  return n;
}

extern "C" int efo_clean(const float* map, int count, const float* new_unstable, int new_count, const double* T_wc,
                         int time, const uint32_t* index, const float* vert_conf4, const float* color_time4,
                         const float* norm_rad4, float conf_threshold, int time_delta, float max_depth, int rows,
                         int cols, const float* cam4, float* out) {
  (void)norm_rad4;
  (void)max_depth;
  Cam c{cam4[0], cam4[1], cam4[2], cam4[3]};
  const Pose t_inv = to_inv_pose_f(T_wc);
  const int total = count + new_count;
  std::vector<uint8_t> keep(total, 0);
  std::vector<float> lastTime(total, 0.f);
#pragma omp parallel for schedule(static)
  for (int k = 0; k < total; ++k) {
    float s[12];
    memcpy(s, (k < count) ? map + (size_t)k * 12 : new_unstable + (size_t)(k - count) * 12, sizeof(s));
    keep[k] = clean_test(s, t_inv, c, time, index, vert_conf4, color_time4, conf_threshold, time_delta, rows, cols) ? 1 : 0;
    lastTime[k] = s[7];
  }
  int n = 0;
  for (int k = 0; k < total; ++k)
    if (keep[k]) {
      float* o = out + (size_t)n * 12;
      memcpy(o, (k < count) ? map + (size_t)k * 12 : new_unstable + (size_t)(k - count) * 12, 48);
      o[7] = lastTime[k];
      ++n;
    }
  return n;
}

// ---------------------------------------------------------------------------------------------
// model raycast: splat.vert + combo_splat.frag / depth_splat.frag
// ---------------------------------------------------------------------------------------------

namespace {
struct Splat {
  bool visible;
  f3 pos;      // camera frame
  float conf;
  f3 nrm;      // camera frame, normalised
  float rad;
  float xw, yw, size;
};

inline f3 project_image(const Cam& c, const f3& p) { return mk3(((c.fx * p.x) / p.z) + c.cx, ((c.fy * p.y) / p.z) + c.cy, p.z); }

inline Splat splat_vertex(const float* s, const Pose& t_inv, const Cam& c, float max_depth, float conf_threshold,
                          int time, int max_time, int time_delta, int rows, int cols) {
  Splat sp;
  sp.visible = false;
  const float fcols = (float)cols, frows = (float)rows;
  f3 h = xform(t_inv, mk3(s[0], s[1], s[2]));
  if (h.z > max_depth || h.z < 0 || s[3] < conf_threshold || (float)time - s[7] > (float)time_delta || s[7] > (float)max_time)
    return sp;
  float xn = ((((c.fx * h.x) / h.z) + c.cx) - (fcols * 0.5f)) / (fcols * 0.5f);
  float yn = ((((c.fy * h.y) / h.z) + c.cy) - (frows * 0.5f)) / (frows * 0.5f);
  float zn = h.z / max_depth;
  if (!(xn >= -1.f && xn <= 1.f && yn >= -1.f && yn <= 1.f && zn >= -1.f && zn <= 1.f)) return sp;
  sp.pos = h;
  sp.conf = s[3];
  sp.nrm = normalized(rot(t_inv, mk3(s[8], s[9], s[10])));
  sp.rad = s[11];
  f3 x1 = normalized(mk3((sp.nrm.y - sp.nrm.z), -sp.nrm.x, sp.nrm.x)) * sp.rad * 1.41421356f;
  f3 y1 = cross(sp.nrm, x1);
  f3 p1 = project_image(c, h + x1), p2 = project_image(c, h + y1), p3 = project_image(c, h - y1), p4 = project_image(c, h - x1);
  float xs0 = gmin(p1.x, gmin(p2.x, gmin(p3.x, p4.x))), xs1 = gmax(p1.x, gmax(p2.x, gmax(p3.x, p4.x)));
  float ys0 = gmin(p1.y, gmin(p2.y, gmin(p3.y, p4.y))), ys1 = gmax(p1.y, gmax(p2.y, gmax(p3.y, p4.y)));
  float xDiff = fabsf(xs1 - xs0), yDiff = fabsf(ys1 - ys0);
  float size = gmax(0.f, gmax(xDiff, yDiff));
  // point size clamp [1, 2047] (ALIASED_POINT_SIZE_RANGE on the reference's target GPUs); NaN -> 1
  if (!(size >= 1.0f)) size = 1.0f;
  if (size > 2047.0f) size = 2047.0f;
  sp.size = size;
  sp.xw = (xn + 1.0f) * (fcols * 0.5f);
  sp.yw = (yn + 1.0f) * (frows * 0.5f);
  sp.visible = true;
  return sp;
}

// combo_splat.frag:33-48 — returns false on discard
inline bool splat_fragment(const Splat& sp, const Cam& c, int px, int py, f3& corrected) {
  float fxc = (float)px + 0.5f, fyc = (float)py + 0.5f;
  f3 l = normalized(mk3((fxc - c.cx) / c.fx, (fyc - c.cy) / c.fy, 1.0f));
  corrected = l * (dot(sp.pos, sp.nrm) / dot(l, sp.nrm));
  float sqrRad = sp.rad * sp.rad;
  f3 diff = corrected - sp.pos;
  if (dot(diff, diff) > sqrRad) return false;
  return true;
}

inline void splat_bounds(const Splat& sp, int rows, int cols, int& x0, int& x1, int& y0, int& y1) {
  float half = sp.size * 0.5f;
  (void)half;
  sprite_range(sp.xw, sp.size, x0, x1);
  sprite_range(sp.yw, sp.size, y0, y1);
  if (x0 < 0) x0 = 0;
  if (y0 < 0) y0 = 0;
  if (x1 >= cols) x1 = cols - 1;
  if (y1 >= rows) y1 = rows - 1;
}
}  // namespace

extern "C" void efo_combined_predict(const float* map, int count, const double* T_wc, float max_depth,
                                     float conf_threshold, int time, int max_time, int time_delta, int rows, int cols,
                                     const float* cam4, uint8_t* image4, float* vertex4, float* normal4,
                                     uint16_t* time_out, float* depth_out, int depth_only) {
  Cam c{cam4[0], cam4[1], cam4[2], cam4[3]};
  const Pose t_inv = to_inv_pose_f(T_wc);
  const size_t n = (size_t)rows * cols;
  std::vector<uint64_t> zbuf(n, kEmptyKey);
#pragma omp parallel for schedule(dynamic, 1024)
  for (int id = 0; id < count; ++id) {
    Splat sp = splat_vertex(map + (size_t)id * 12, t_inv, c, max_depth, conf_threshold, time, max_time, time_delta, rows, cols);
    if (!sp.visible) continue;
    int x0, x1, y0, y1;
    splat_bounds(sp, rows, cols, x0, x1, y0, y1);
    for (int py = y0; py <= y1; ++py)
      for (int px = x0; px <= x1; ++px) {
        f3 cp;
        if (!splat_fragment(sp, c, px, py, cp)) continue;
        float fd = (cp.z / (2 * max_depth)) + 0.5f;
        uint64_t key = ((uint64_t)depth24(fd) << 32) | (uint32_t)id;
        // GL_LESS against the cleared depth 1.0: a fragment at depth 1.0 never passes
        if (depth24(fd) >= 16777215u) continue;
        atomic_min_u64(&zbuf[(size_t)py * cols + px], key);
      }
  }
#pragma omp parallel for schedule(static)
  for (int p = 0; p < (int)n; ++p) {
    int px = p % cols, py = p / cols;
    if (zbuf[p] == kEmptyKey) {
      if (depth_only) {
        depth_out[p] = 0.f;
      } else {
        for (int q = 0; q < 4; ++q) {
          image4[(size_t)p * 4 + q] = 0;
          vertex4[(size_t)p * 4 + q] = 0.f;
          normal4[(size_t)p * 4 + q] = 0.f;
        }
        time_out[p] = 0;
      }
      continue;
    }
    uint32_t id = (uint32_t)(zbuf[p] & 0xffffffffu);
    const float* s = map + (size_t)id * 12;
    Splat sp = splat_vertex(s, t_inv, c, max_depth, conf_threshold, time, max_time, time_delta, rows, cols);
    f3 cp;
    splat_fragment(sp, c, px, py, cp);
    if (depth_only) {
      depth_out[p] = cp.z;
      continue;
    }
    f3 col = decode_color(s[4]);
    image4[(size_t)p * 4 + 0] = (uint8_t)(int)rintf(col.x * 255.0f);
    image4[(size_t)p * 4 + 1] = (uint8_t)(int)rintf(col.y * 255.0f);
    image4[(size_t)p * 4 + 2] = (uint8_t)(int)rintf(col.z * 255.0f);
    image4[(size_t)p * 4 + 3] = 255;
    float z = cp.z;
    float fxc = (float)px + 0.5f, fyc = (float)py + 0.5f;
    vertex4[(size_t)p * 4 + 0] = (fxc - c.cx) * z * (1.f / c.fx);
    vertex4[(size_t)p * 4 + 1] = (fyc - c.cy) * z * (1.f / c.fy);
    vertex4[(size_t)p * 4 + 2] = z;
    vertex4[(size_t)p * 4 + 3] = sp.conf;
    normal4[(size_t)p * 4 + 0] = sp.nrm.x;
    normal4[(size_t)p * 4 + 1] = sp.nrm.y;
    normal4[(size_t)p * 4 + 2] = sp.nrm.z;
    normal4[(size_t)p * 4 + 3] = sp.rad;
    time_out[p] = (uint16_t)(uint32_t)s[6];
  }
}

// ---------------------------------------------------------------------------------------------
// fill-in and density test
// ---------------------------------------------------------------------------------------------

// fill_vertex.frag via FillIn::vertex (FillIn.cpp:98-140). cam4 = cx,cy,fx,fy
extern "C" void efo_fill_vertex(const float* existing4, const uint16_t* raw_depth, int passthrough, int rows, int cols,
                                const float* cam4, float* out4) {
  Cam c{cam4[0], cam4[1], cam4[2], cam4[3]};
  const float ifx = 1.0f / c.fx, ify = 1.0f / c.fy;
#pragma omp parallel for schedule(static)
  for (int y = 0; y < rows; ++y)
    for (int x = 0; x < cols; ++x) {
      const float* s = existing4 + (size_t)(y * cols + x) * 4;
      float* o = out4 + (size_t)(y * cols + x) * 4;
      if (s[2] == 0 || passthrough == 1) {
        f3 v = vertex_u(raw_depth, rows, cols, x, y, x, y, c, ifx, ify);
        o[0] = v.x;
        o[1] = v.y;
        o[2] = v.z;
        o[3] = 1;
      } else {
        memcpy(o, s, 16);
      }
    }
}

// fill_normal.frag via FillIn::normal (FillIn.cpp:142-184)
extern "C" void efo_fill_normal(const float* existing4, const uint16_t* raw_depth, int passthrough, int rows, int cols,
                                const float* cam4, float* out4) {
  Cam c{cam4[0], cam4[1], cam4[2], cam4[3]};
  const float ifx = 1.0f / c.fx, ify = 1.0f / c.fy;
#pragma omp parallel for schedule(static)
  for (int y = 0; y < rows; ++y)
    for (int x = 0; x < cols; ++x) {
      const float* s = existing4 + (size_t)(y * cols + x) * 4;
      float* o = out4 + (size_t)(y * cols + x) * 4;
      if (s[2] == 0 || passthrough == 1) {
        f3 v = vertex_u(raw_depth, rows, cols, x, y, x, y, c, ifx, ify);
        f3 vx = vertex_u(raw_depth, rows, cols, x + 1, y, x + 1, y, c, ifx, ify);
        f3 vy = vertex_u(raw_depth, rows, cols, x, y + 1, x, y + 1, c, ifx, ify);
        f3 nn = normalized(cross(vx - v, vy - v));
        o[0] = nn.x;
        o[1] = nn.y;
        o[2] = nn.z;
        o[3] = 1;
      } else {
        memcpy(o, s, 16);
      }
    }
}

// fill_rgb.frag via FillIn::image (FillIn.cpp:62-96): existing RGBA8, raw RGB8 (alpha reads 1.0)
extern "C" void efo_fill_image(const uint8_t* existing4, const uint8_t* rgb, int passthrough, int rows, int cols,
                               uint8_t* out4) {
#pragma omp parallel for schedule(static)
  for (int p = 0; p < rows * cols; ++p) {
    const uint8_t* s = existing4 + (size_t)p * 4;
    uint8_t* o = out4 + (size_t)p * 4;
    if (((int)s[0] + (int)s[1] + (int)s[2] == 0) || passthrough == 1) {
      o[0] = rgb[(size_t)p * 3 + 0];
      o[1] = rgb[(size_t)p * 3 + 1];
      o[2] = rgb[(size_t)p * 3 + 2];
      o[3] = 255;
    } else {
      memcpy(o, s, 4);
    }
  }
}

// Resize::image (Resize.cpp:50-79, nearest decimation by `factor`) + ElasticFusion::denseEnough (ElasticFusion.cpp:256-268)
extern "C" int efo_dense_enough(const uint8_t* image4, int rows, int cols, int factor) {
  const int drows = rows / factor, dcols = cols / factor;
  int sum = 0;
  for (int j = 0; j < drows; ++j)
    for (int i = 0; i < dcols; ++i) {
      int sx = texel(((float)i + 0.5f) / (float)dcols, cols);
      int sy = texel(((float)j + 0.5f) / (float)drows, rows);
      const uint8_t* s = image4 + (size_t)(sy * cols + sx) * 4;
      sum += (s[0] > 0 && s[1] > 0 && s[2] > 0) ? 1 : 0;
    }
  return ((float)sum / (float)(drows * dcols) > 0.75f) ? 1 : 0;
}
