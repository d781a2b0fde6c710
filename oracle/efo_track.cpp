// TEST INFRASTRUCTURE ONLY — CPU oracle, tracking half.
// Restates Core/Cuda/cudafuncs.cu, Core/Cuda/reduce.cu and Core/Utils/RGBDOdometry.cpp of the
// reference. Per-pixel arithmetic is fp32 in the reference's operation order; the grid reductions
// accumulate in double (the reference sums fp32 in a thread/warp/block tree, reduce.cu:57-140 — the
// double sum is the order-independent value that tree approximates to ~1e-6 relative).
#include "ef_oracle.h"
#include "efo_common.h"
#include "efo_linalg.h"

#include <cassert>
#include <cfloat>
#include <cstdio>
#include <limits>

using namespace efo;

// ---------------------------------------------------------------------------------------------
// image kernels
// ---------------------------------------------------------------------------------------------

// cudafuncs.cu:75-121 pyrDownGaussKernel / pyrDown (sigma_color = 30)
extern "C" void efo_pyr_down_u16(const uint16_t* src, int srows, int scols, uint16_t* dst) {
  const int drows = srows / 2, dcols = scols / 2;
  const float sigma_color = 30.f;
  const float weights[3] = {0.375f, 0.25f, 0.0625f};
#pragma omp parallel for schedule(static)
  for (int y = 0; y < drows; ++y) {
    for (int x = 0; x < dcols; ++x) {
      const int D = 5;
      int center = src[(2 * y) * scols + 2 * x];
      int x_mi = imax(0, 2 * x - D / 2) - 2 * x;
      int y_mi = imax(0, 2 * y - D / 2) - 2 * y;
      int x_ma = imin(scols, 2 * x - D / 2 + D) - 2 * x;
      int y_ma = imin(srows, 2 * y - D / 2 + D) - 2 * y;
      float sum = 0, wall = 0;
      for (int yi = y_mi; yi < y_ma; ++yi)
        for (int xi = x_mi; xi < x_ma; ++xi) {
          int val = src[(2 * y + yi) * scols + 2 * x + xi];
          if ((float)abs(val - center) < 3 * sigma_color) {
            sum += val * weights[abs(xi)] * weights[abs(yi)];
            wall += weights[abs(xi)] * weights[abs(yi)];
          }
        }
      dst[y * dcols + x] = (uint16_t)f2i_rz(sum / wall);
    }
  }
}

// cudafuncs.cu:123-168 computeVmapKernel / createVMap. Invalid -> NaN in the x plane only (y,z untouched).
extern "C" void efo_create_vmap(const uint16_t* depth, int rows, int cols, float fx, float fy, float cx, float cy,
                                float depth_cutoff, float* vmap) {
  const float fx_inv = 1.f / fx, fy_inv = 1.f / fy;
#pragma omp parallel for schedule(static)
  for (int v = 0; v < rows; ++v)
    for (int u = 0; u < cols; ++u) {
      float z = depth[v * cols + u] / 1000.f;
      if (z != 0 && z < depth_cutoff) {
        float vx = z * (u - cx) * fx_inv;
        float vy = z * (v - cy) * fy_inv;
        vmap[v * cols + u] = vx;
        vmap[(v + rows) * cols + u] = vy;
        vmap[(v + 2 * rows) * cols + u] = z;
      } else {
        vmap[v * cols + u] = qnan();
      }
    }
}

// cudafuncs.cu:170-219 computeNmapKernel / createNMap
extern "C" void efo_create_nmap(const float* vmap, int rows, int cols, float* nmap) {
#pragma omp parallel for schedule(static)
  for (int v = 0; v < rows; ++v)
    for (int u = 0; u < cols; ++u) {
      if (u == cols - 1 || v == rows - 1) {
        nmap[v * cols + u] = qnan();
        continue;
      }
      f3 v00, v01, v10;
      v00.x = vmap[v * cols + u];
      v01.x = vmap[v * cols + u + 1];
      v10.x = vmap[(v + 1) * cols + u];
      if (!std::isnan(v00.x) && !std::isnan(v01.x) && !std::isnan(v10.x)) {
        v00.y = vmap[(v + rows) * cols + u];
        v01.y = vmap[(v + rows) * cols + u + 1];
        v10.y = vmap[(v + 1 + rows) * cols + u];
        v00.z = vmap[(v + 2 * rows) * cols + u];
        v01.z = vmap[(v + 2 * rows) * cols + u + 1];
        v10.z = vmap[(v + 1 + 2 * rows) * cols + u];
        f3 r = normalized(cross(v01 - v00, v10 - v00));
        nmap[v * cols + u] = r.x;
        nmap[(v + rows) * cols + u] = r.y;
        nmap[(v + 2 * rows) * cols + u] = r.z;
      } else {
        nmap[v * cols + u] = qnan();
      }
    }
}

static inline m33 load_m33(const float* R) {
  m33 m;
  m.r[0] = mk3(R[0], R[1], R[2]);
  m.r[1] = mk3(R[3], R[4], R[5]);
  m.r[2] = mk3(R[6], R[7], R[8]);
  return m;
}

// cudafuncs.cu:221-293 tranformMapsKernel / tranformMaps — the reference calls it in place (RGBDOdometry.cpp:199-207)
extern "C" void efo_transform_maps(float* vmap, float* nmap, int rows, int cols, const float* R, const float* t) {
  const m33 Rm = load_m33(R);
  const f3 tv = mk3(t[0], t[1], t[2]);
#pragma omp parallel for schedule(static)
  for (int y = 0; y < rows; ++y)
    for (int x = 0; x < cols; ++x) {
      f3 vsrc, vdst = mk3(qnan(), qnan(), qnan());
      vsrc.x = vmap[y * cols + x];
      if (!std::isnan(vsrc.x)) {
        vsrc.y = vmap[(y + rows) * cols + x];
        vsrc.z = vmap[(y + 2 * rows) * cols + x];
        vdst = mul(Rm, vsrc) + tv;
        vmap[(y + rows) * cols + x] = vdst.y;
        vmap[(y + 2 * rows) * cols + x] = vdst.z;
      }
      vmap[y * cols + x] = vdst.x;
      f3 nsrc, ndst = mk3(qnan(), qnan(), qnan());
      nsrc.x = nmap[y * cols + x];
      if (!std::isnan(nsrc.x)) {
        nsrc.y = nmap[(y + rows) * cols + x];
        nsrc.z = nmap[(y + 2 * rows) * cols + x];
        ndst = mul(Rm, nsrc);
        nmap[(y + rows) * cols + x] = ndst.y;
        nmap[(y + 2 * rows) * cols + x] = ndst.z;
      }
      nmap[y * cols + x] = ndst.x;
    }
}

// cudafuncs.cu:295-381 copyMapsKernelTex / copyMaps: float4 textures -> vmaps_tmp (AoS) + SoA maps, z==0 -> NaN.
extern "C" void efo_copy_maps(const float* vtx4, const float* nrm4, int rows, int cols, float* vmaps_tmp,
                              float* vmap, float* nmap) {
#pragma omp parallel for schedule(static)
  for (int y = 0; y < rows; ++y)
    for (int x = 0; x < cols; ++x) {
      const float* vs = vtx4 + (size_t)(y * cols + x) * 4;
      const float* ns = nrm4 + (size_t)(y * cols + x) * 4;
      if (vmaps_tmp) {
        float* tmp = vmaps_tmp + (size_t)(y * cols + x) * 4;
        tmp[0] = vs[0];
        tmp[1] = vs[1];
        tmp[2] = vs[2];
        tmp[3] = vs[3];
      }
      f3 vdst = mk3(qnan(), qnan(), qnan()), ndst = vdst;
      if (!(vs[2] == 0)) {
        vdst = mk3(vs[0], vs[1], vs[2]);
        ndst = mk3(ns[0], ns[1], ns[2]);
      }
      vmap[y * cols + x] = vdst.x;
      vmap[(y + rows) * cols + x] = vdst.y;
      vmap[(y + 2 * rows) * cols + x] = vdst.z;
      nmap[y * cols + x] = ndst.x;
      nmap[(y + rows) * cols + x] = ndst.y;
      nmap[(y + 2 * rows) * cols + x] = ndst.z;
    }
}

// cudafuncs.cu:413-490 resizeMapKernel<normalize>
extern "C" void efo_resize_map(const float* in, int srows, int scols, float* out, int normalize) {
  const int drows = srows / 2, dcols = scols / 2;
#pragma omp parallel for schedule(static)
  for (int y = 0; y < drows; ++y)
    for (int x = 0; x < dcols; ++x) {
      int xs = x * 2, ys = y * 2;
      float x00 = in[(ys + 0) * scols + xs + 0];
      float x01 = in[(ys + 0) * scols + xs + 1];
      float x10 = in[(ys + 1) * scols + xs + 0];
      float x11 = in[(ys + 1) * scols + xs + 1];
      if (std::isnan(x00) || std::isnan(x01) || std::isnan(x10) || std::isnan(x11)) {
        out[y * dcols + x] = qnan();
        continue;
      }
      f3 n;
      n.x = (x00 + x01 + x10 + x11) / 4;
      float y00 = in[(ys + srows + 0) * scols + xs + 0];
      float y01 = in[(ys + srows + 0) * scols + xs + 1];
      float y10 = in[(ys + srows + 1) * scols + xs + 0];
      float y11 = in[(ys + srows + 1) * scols + xs + 1];
      n.y = (y00 + y01 + y10 + y11) / 4;
      float z00 = in[(ys + 2 * srows + 0) * scols + xs + 0];
      float z01 = in[(ys + 2 * srows + 0) * scols + xs + 1];
      float z10 = in[(ys + 2 * srows + 1) * scols + xs + 0];
      float z11 = in[(ys + 2 * srows + 1) * scols + xs + 1];
      n.z = (z00 + z01 + z10 + z11) / 4;
      if (normalize) n = normalized(n);
      out[y * dcols + x] = n.x;
      out[(y + drows) * dcols + x] = n.y;
      out[(y + 2 * drows) * dcols + x] = n.z;
    }
}

static const float kGauss25[25] = {1, 4, 6, 4, 1, 4, 16, 24, 16, 4, 6, 24, 36, 24, 6, 4, 16, 24, 16, 4, 1, 4, 6, 4, 1};

// cudafuncs.cu:383-411,492-510 pyrDownKernelGaussF (window/index quirks)
extern "C" void efo_pyr_down_gauss_f(const float* src, int srows, int scols, float* dst) {
  const int drows = srows / 2, dcols = scols / 2;
#pragma omp parallel for schedule(static)
  for (int y = 0; y < drows; ++y)
    for (int x = 0; x < dcols; ++x) {
      const int D = 5;
      int tx = imin(2 * x - D / 2 + D, scols - 1);
      int ty = imin(2 * y - D / 2 + D, srows - 1);
      int cy = imax(0, 2 * y - D / 2);
      float sum = 0;
      int count = 0;
      for (; cy < ty; ++cy)
        for (int cx = imax(0, 2 * x - D / 2); cx < tx; ++cx) {
          float s = src[cy * scols + cx];
          if (!std::isnan(s)) {
            float w = kGauss25[(ty - cy - 1) * 5 + (tx - cx - 1)];
            sum += s * w;
            count = f2i_rz((float)count + w);
          }
        }
      dst[y * dcols + x] = (float)(sum / (float)count);
    }
}

// cudafuncs.cu:512-562 pyrDownKernelIntensityGauss (skips zeros, float->u8 truncation; 0/0 NaN -> 0 as cvt does)
extern "C" void efo_pyr_down_u8(const uint8_t* src, int srows, int scols, uint8_t* dst) {
  const int drows = srows / 2, dcols = scols / 2;
#pragma omp parallel for schedule(static)
  for (int y = 0; y < drows; ++y)
    for (int x = 0; x < dcols; ++x) {
      const int D = 5;
      int tx = imin(2 * x - D / 2 + D, scols - 1);
      int ty = imin(2 * y - D / 2 + D, srows - 1);
      int cy = imax(0, 2 * y - D / 2);
      float sum = 0;
      int count = 0;
      for (; cy < ty; ++cy)
        for (int cx = imax(0, 2 * x - D / 2); cx < tx; ++cx) {
          uint8_t s = src[cy * scols + cx];
          if (s > 0) {
            float w = kGauss25[(ty - cy - 1) * 5 + (tx - cx - 1)];
            sum += s * w;
            count = f2i_rz((float)count + w);
          }
        }
      int v = f2i_rz(sum / (float)count);
      dst[y * dcols + x] = (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
    }
}

// cudafuncs.cu:564-582 verticesToDepthKernel
extern "C" void efo_vertices_to_depth(const float* vmaps_tmp, int rows, int cols, float cutoff, float* dst) {
#pragma omp parallel for schedule(static)
  for (int y = 0; y < rows; ++y)
    for (int x = 0; x < cols; ++x) {
      float z = vmaps_tmp[(size_t)y * cols * 4 + (x * 4) + 2];
      dst[y * cols + x] = (z > cutoff || z <= 0) ? qnan() : z;
    }
}

// cudafuncs.cu:584-610 bgr2IntensityKernel: int(0.114*x + 0.299*y + 0.587*z) on the RGBA8 texel
extern "C" void efo_rgba_to_intensity(const uint8_t* rgba, int rows, int cols, uint8_t* dst) {
#pragma omp parallel for schedule(static)
  for (int i = 0; i < rows * cols; ++i) {
    const uint8_t* s = rgba + (size_t)i * 4;
    int value = f2i_rz((float)s[0] * 0.114f + (float)s[1] * 0.299f + (float)s[2] * 0.587f);
    dst[i] = (uint8_t)value;
  }
}

// cudafuncs.cu:612-668 applyKernel (scaled Sobel, decrementing kernelIndex quirk App. A-10)
extern "C" void efo_sobel(const uint8_t* src, int rows, int cols, int16_t* dx, int16_t* dy) {
  // the reference initialises these from double literals narrowed to float (cudafuncs.cu:646-650)
  const float gsx[9] = {(float)0.52201, (float)0.00000, (float)-0.52201, (float)0.79451, (float)-0.00000,
                        (float)-0.79451, (float)0.52201, (float)0.00000, (float)-0.52201};
  const float gsy[9] = {(float)0.52201, (float)0.79451, (float)0.52201, (float)0.00000, (float)0.00000,
                        (float)0.00000, (float)-0.52201, (float)-0.79451, (float)-0.52201};
#pragma omp parallel for schedule(static)
  for (int y = 0; y < rows; ++y)
    for (int x = 0; x < cols; ++x) {
      float dxVal = 0, dyVal = 0;
      int kernelIndex = 8;
      for (int j = imax(y - 1, 0); j <= imin(y + 1, rows - 1); j++)
        for (int i = imax(x - 1, 0); i <= imin(x + 1, cols - 1); i++) {
          dxVal += (float)src[j * cols + i] * gsx[kernelIndex];
          dyVal += (float)src[j * cols + i] * gsy[kernelIndex];
          --kernelIndex;
        }
      dx[y * cols + x] = (int16_t)f2i_rz(dxVal);
      dy[y * cols + x] = (int16_t)f2i_rz(dyVal);
    }
}

// cudafuncs.cu:670-709 projectPointsKernel
extern "C" void efo_project_points(const float* depth, int rows, int cols, float fx, float fy, float cx, float cy,
                                   float* cloud3) {
  const float invFx = 1.0f / fx, invFy = 1.0f / fy;
#pragma omp parallel for schedule(static)
  for (int y = 0; y < rows; ++y)
    for (int x = 0; x < cols; ++x) {
      float z = depth[y * cols + x];
      float* c = cloud3 + (size_t)(y * cols + x) * 3;
      c[0] = (float)((x - cx) * z * invFx);
      c[1] = (float)((y - cy) * z * invFy);
      c[2] = z;
    }
}

// ---------------------------------------------------------------------------------------------
// reductions
// ---------------------------------------------------------------------------------------------

static inline void accumulate_row7(const float row[7], bool found, double acc[29]) {
  int k = 0;
  for (int i = 0; i < 6; ++i)
    for (int j = i; j < 7; ++j) acc[k++] += (double)(row[i] * row[j]);
  acc[27] += (double)(row[6] * row[6]);
  acc[28] += (double)(found ? 1.0f : 0.0f);
}

// host unpack, reduce.cu:388-400
static inline void unpack_se3(const double acc[29], float* A36, float* b6, float* residual2) {
  int shift = 0;
  for (int i = 0; i < 6; ++i)
    for (int j = i; j < 7; ++j) {
      float value = (float)acc[shift++];
      if (j == 6)
        b6[i] = value;
      else
        A36[j * 6 + i] = A36[i * 6 + j] = value;
    }
  if (residual2) {
    residual2[0] = (float)acc[27];
    residual2[1] = (float)acc[28];
  }
}

// reduce.cu:204-401 ICPReduction::{search,getProducts} / icpStep
extern "C" void efo_icp_step(const float* Rcurr_, const float* tcurr_, const float* vmap_curr, const float* nmap_curr,
                             const float* Rprev_inv_, const float* tprev_, float fx, float fy, float cx, float cy,
                             const float* vmap_g_prev, const float* nmap_g_prev, float dist_thres, float angle_thres,
                             int rows, int cols, float* A36, float* b6, float* residual2) {
  const m33 Rcurr = load_m33(Rcurr_), Rprev_inv = load_m33(Rprev_inv_);
  const f3 tcurr = mk3(tcurr_[0], tcurr_[1], tcurr_[2]), tprev = mk3(tprev_[0], tprev_[1], tprev_[2]);
  std::vector<double> partial((size_t)rows * 29, 0.0);
#pragma omp parallel for schedule(static)
  for (int y = 0; y < rows; ++y) {
    double* acc = &partial[(size_t)y * 29];
    for (int x = 0; x < cols; ++x) {
      f3 vcurr = mk3(vmap_curr[y * cols + x], vmap_curr[(y + rows) * cols + x], vmap_curr[(y + 2 * rows) * cols + x]);
      f3 vcurr_g = mul(Rcurr, vcurr) + tcurr;
      f3 vcurr_cp = mul(Rprev_inv, vcurr_g - tprev);
      int ux = f2i_rn(vcurr_cp.x * fx / vcurr_cp.z + cx);
      int uy = f2i_rn(vcurr_cp.y * fy / vcurr_cp.z + cy);
      float row[7] = {0, 0, 0, 0, 0, 0, 0};
      bool found = false;
      if (!(ux < 0 || uy < 0 || ux >= cols || uy >= rows || vcurr_cp.z < 0)) {
        f3 vprev_g = mk3(vmap_g_prev[uy * cols + ux], vmap_g_prev[(uy + rows) * cols + ux],
                         vmap_g_prev[(uy + 2 * rows) * cols + ux]);
        f3 ncurr = mk3(nmap_curr[y * cols + x], nmap_curr[(y + rows) * cols + x], nmap_curr[(y + 2 * rows) * cols + x]);
        f3 ncurr_g = mul(Rcurr, ncurr);
        f3 nprev_g = mk3(nmap_g_prev[uy * cols + ux], nmap_g_prev[(uy + rows) * cols + ux],
                         nmap_g_prev[(uy + 2 * rows) * cols + ux]);
        float dist = norm(vprev_g - vcurr_g);
        float sine = norm(cross(ncurr_g, nprev_g));
        found = (sine < angle_thres && dist <= dist_thres && !std::isnan(ncurr.x) && !std::isnan(nprev_g.x));
        if (found) {
          f3 s_cp = mul(Rprev_inv, vcurr_g - tprev);
          f3 d_cp = mul(Rprev_inv, vprev_g - tprev);
          f3 n_cp = mul(Rprev_inv, nprev_g);
          f3 c = cross(s_cp, n_cp);
          row[0] = n_cp.x;
          row[1] = n_cp.y;
          row[2] = n_cp.z;
          row[3] = c.x;
          row[4] = c.y;
          row[5] = c.z;
          row[6] = dot(n_cp, s_cp - d_cp);
        }
      }
      accumulate_row7(row, found, acc);
    }
  }
  double acc[29] = {0};
  for (int y = 0; y < rows; ++y)
    for (int k = 0; k < 29; ++k) acc[k] += partial[(size_t)y * 29 + k];
  unpack_se3(acc, A36, b6, residual2);
}

// reduce.cu:603-787 RGBResidual::getProducts / computeRgbResidual
extern "C" void efo_rgb_residual(float min_scale, const int16_t* dIdx, const int16_t* dIdy, const float* last_depth,
                                 const float* next_depth, const uint8_t* last_image, const uint8_t* next_image,
                                 EfoDataTerm* corres_img, float max_depth_delta, const float* kt3,
                                 const float* krkinv9, int rows, int cols, int* sigma_sum, int* count_out) {
  const m33 krkinv = load_m33(krkinv9);
  const f3 kt = mk3(kt3[0], kt3[1], kt3[2]);
  std::vector<uint32_t> pc(rows, 0), ps(rows, 0);
#pragma omp parallel for schedule(static)
  for (int i = 0; i < rows; ++i) {
    uint32_t cnt = 0, sig = 0;
    for (int j0 = 0; j0 < cols; ++j0) {
      EfoDataTerm corres;
      memset(&corres, 0, sizeof(corres));
      corres.valid = 0;
      if (j0 < cols - 5 && i < rows - 1) {
        bool valid = true;
        for (int u = imax(i - 2, 0); u < imin(i + 2, rows); u++)
          for (int v = imax(j0 - 2, 0); v < imin(j0 + 2, cols); v++) valid = valid && (next_image[u * cols + v] > 0);
        if (valid) {
          int16_t valx = dIdx[i * cols + j0];
          int16_t valy = dIdy[i * cols + j0];
          float mTwo = (float)((valx * valx) + (valy * valy));
          if (mTwo >= min_scale) {
            int y = i, x = j0;
            float d1 = next_depth[y * cols + x];
            if (!std::isnan(d1)) {
              float transformed_d1 = (float)(d1 * (krkinv.r[2].x * x + krkinv.r[2].y * y + krkinv.r[2].z) + kt.z);
              int u0 = f2i_rn((d1 * (krkinv.r[0].x * x + krkinv.r[0].y * y + krkinv.r[0].z) + kt.x) / transformed_d1);
              int v0 = f2i_rn((d1 * (krkinv.r[1].x * x + krkinv.r[1].y * y + krkinv.r[1].z) + kt.y) / transformed_d1);
              if (u0 >= 0 && v0 >= 0 && u0 < cols && v0 < rows) {
                float d0 = last_depth[v0 * cols + u0];
                if (d0 > 0 && std::fabs(transformed_d1 - d0) <= max_depth_delta && last_image[v0 * cols + u0] != 0) {
                  corres.zero_x = (int16_t)u0;
                  corres.zero_y = (int16_t)v0;
                  corres.one_x = (int16_t)x;
                  corres.one_y = (int16_t)y;
                  corres.diff = (float)next_image[y * cols + x] - (float)last_image[v0 * cols + u0];
                  corres.valid = 1;
                  cnt += 1;
                  sig += (uint32_t)f2i_rz(corres.diff * corres.diff);
                }
              }
            }
          }
        }
      }
      corres_img[i * cols + j0] = corres;
    }
    pc[i] = cnt;
    ps[i] = sig;
  }
  uint32_t c = 0, s = 0;  // int accumulators wrap like the reference's int2 sums
  for (int i = 0; i < rows; ++i) {
    c += pc[i];
    s += ps[i];
  }
  *count_out = (int)c;
  *sigma_sum = (int)s;
}

// reduce.cu:403-550 RGBReduction::getProducts / rgbStep
extern "C" void efo_rgb_step(const EfoDataTerm* corres_img, float sigma, const float* cloud3, float fx, float fy,
                             const int16_t* dIdx, const int16_t* dIdy, float sobel_scale, int rows, int cols,
                             float* A36, float* b6) {
  std::vector<double> partial((size_t)rows * 29, 0.0);
#pragma omp parallel for schedule(static)
  for (int y = 0; y < rows; ++y) {
    double* acc = &partial[(size_t)y * 29];
    for (int x = 0; x < cols; ++x) {
      const EfoDataTerm& corresp = corres_img[y * cols + x];
      bool found = corresp.valid != 0;
      float row[7] = {0, 0, 0, 0, 0, 0, 0};
      if (found) {
        float w = sigma + std::fabs(corresp.diff);
        w = w > 1.19209290E-07F ? 1.0f / w : 1.0f;
        if (sigma == -1) w = 1;
        row[6] = -w * corresp.diff;
        const float* cp = cloud3 + (size_t)(corresp.zero_y * cols + corresp.zero_x) * 3;
        f3 cloudPoint = mk3(cp[0], cp[1], cp[2]);
        float invz = (float)(1.0 / (double)cloudPoint.z);  // `1.0 / cloudPoint.z` is a double division narrowed
        float dI_dx_val = w * sobel_scale * dIdx[corresp.one_y * cols + corresp.one_x];
        float dI_dy_val = w * sobel_scale * dIdy[corresp.one_y * cols + corresp.one_x];
        float v0 = dI_dx_val * fx * invz;
        float v1 = dI_dy_val * fy * invz;
        float v2 = -(v0 * cloudPoint.x + v1 * cloudPoint.y) * invz;
        row[0] = v0;
        row[1] = v1;
        row[2] = v2;
        row[3] = -cloudPoint.z * v1 + cloudPoint.y * v2;
        row[4] = cloudPoint.z * v0 - cloudPoint.x * v2;
        row[5] = -cloudPoint.y * v0 + cloudPoint.x * v1;
      }
      accumulate_row7(row, found, acc);
    }
  }
  double acc[29] = {0};
  for (int y = 0; y < rows; ++y)
    for (int k = 0; k < 29; ++k) acc[k] += partial[(size_t)y * 29 + k];
  unpack_se3(acc, A36, b6, nullptr);
}

static inline void so3_gradient(const uint8_t* img, int cols, int x, int y, float& gx, float& gy) {
  float actu = (float)img[y * cols + x];
  float back = (float)img[y * cols + x - 1];
  float fore = (float)img[y * cols + x + 1];
  gx = ((back + actu) / 2.0f) - ((fore + actu) / 2.0f);
  back = (float)img[(y - 1) * cols + x];
  fore = (float)img[(y + 1) * cols + x];
  gy = ((back + actu) / 2.0f) - ((fore + actu) / 2.0f);
}

// reduce.cu:789-973 SO3Reduction::getProducts / so3Step
extern "C" void efo_so3_step(const uint8_t* last_image, const uint8_t* next_image, const float* image_basis9,
                             const float* kinv9, const float* krlr9, int rows, int cols, float* A9, float* b3,
                             float* residual2) {
  const m33 imageBasis = load_m33(image_basis9), kinv = load_m33(kinv9), krlr = load_m33(krlr9);
  std::vector<double> partial((size_t)rows * 11, 0.0);
#pragma omp parallel for schedule(static)
  for (int y = 0; y < rows; ++y) {
    double* acc = &partial[(size_t)y * 11];
    for (int x = 0; x < cols; ++x) {
      bool found = false;
      f3 unwarped = mk3((float)x, (float)y, 1.0f);
      f3 warped = mul(imageBasis, unwarped);
      int wx = f2i_rn(warped.x / warped.z);
      int wy = f2i_rn(warped.y / warped.z);
      if (wx >= 1 && wx < cols - 1 && wy >= 1 && wy < rows - 1 && x >= 1 && x < cols - 1 && y >= 1 && y < rows - 1)
        found = true;
      float row[4] = {0, 0, 0, 0};
      if (found) {
        float gnx, gny, glx, gly;
        so3_gradient(next_image, cols, wx, wy, gnx, gny);
        so3_gradient(last_image, cols, x, y, glx, gly);
        float gx = (gnx + glx) / 2.0f;
        float gy = (gny + gly) / 2.0f;
        f3 point = mul(kinv, unwarped);
        float z2 = point.z * point.z;
        float a = krlr.r[0].x, b = krlr.r[0].y, c = krlr.r[0].z;
        float d = krlr.r[1].x, e = krlr.r[1].y, f = krlr.r[1].z;
        float g = krlr.r[2].x, h = krlr.r[2].y, i = krlr.r[2].z;
        f3 leftProduct = mk3(((point.z * (d * gy + a * gx)) - (gy * g * y) - (gx * g * x)) / z2,
                             ((point.z * (e * gy + b * gx)) - (gy * h * y) - (gx * h * x)) / z2,
                             ((point.z * (f * gy + c * gx)) - (gy * i * y) - (gx * i * x)) / z2);
        f3 jacRow = cross(leftProduct, point);
        row[0] = jacRow.x;
        row[1] = jacRow.y;
        row[2] = jacRow.z;
        row[3] = -((float)next_image[wy * cols + wx] - (float)last_image[y * cols + x]);
      }
      int k = 0;
      for (int i = 0; i < 3; ++i)
        for (int j = i; j < 4; ++j) acc[k++] += (double)(row[i] * row[j]);
      acc[9] += (double)(row[3] * row[3]);
      acc[10] += (double)(found ? 1.0f : 0.0f);
    }
  }
  double acc[11] = {0};
  for (int y = 0; y < rows; ++y)
    for (int k = 0; k < 11; ++k) acc[k] += partial[(size_t)y * 11 + k];
  int shift = 0;
  for (int i = 0; i < 3; ++i)
    for (int j = i; j < 4; ++j) {
      float value = (float)acc[shift++];
      if (j == 3)
        b3[i] = value;
      else
        A9[j * 3 + i] = A9[i * 3 + j] = value;
    }
  residual2[0] = (float)acc[9];
  residual2[1] = (float)acc[10];
}

// ---------------------------------------------------------------------------------------------
// RGBDOdometry host logic (Core/Utils/RGBDOdometry.cpp)
// ---------------------------------------------------------------------------------------------

static const int NUM_PYRS = 3;

struct EfoOdometry {
  int width, height;
  float cx, cy, fx, fy;
  float distThres, angleThres;
  float sobelScale, maxDepthDeltaRGB, maxDepthRGB;
  float minGrad[NUM_PYRS];
  int prow[NUM_PYRS], pcol[NUM_PYRS];

  std::vector<uint16_t> depth_tmp[NUM_PYRS];
  std::vector<float> vmaps_tmp;
  std::vector<float> vmaps_g_prev[NUM_PYRS], nmaps_g_prev[NUM_PYRS], vmaps_curr[NUM_PYRS], nmaps_curr[NUM_PYRS];
  std::vector<float> lastDepth[NUM_PYRS], nextDepth[NUM_PYRS];
  std::vector<uint8_t> lastImage[NUM_PYRS], nextImage[NUM_PYRS], lastNextImage[NUM_PYRS];
  std::vector<int16_t> dIdx[NUM_PYRS], dIdy[NUM_PYRS];
  std::vector<EfoDataTerm> corresImg[NUM_PYRS];
  std::vector<float> pointClouds[NUM_PYRS];

  float lastICPError, lastICPCount, lastRGBError, lastRGBCount, lastSO3Error, lastSO3Count;
  double lastA[36], lastb[6];
};

// CameraModel::operator()(level), types.cuh:92-95
static inline void intr_level(const EfoOdometry* o, int level, float& fx, float& fy, float& cx, float& cy) {
  int div = 1 << level;
  fx = o->fx / div;
  fy = o->fy / div;
  cx = o->cx / div;
  cy = o->cy / div;
}

// RGBDOdometry.cpp:22-117
extern "C" EfoOdometry* efo_odom_create(int width, int height, float cx, float cy, float fx, float fy,
                                        float dist_thresh, float angle_thresh) {
  EfoOdometry* o = new EfoOdometry();
  o->width = width;
  o->height = height;
  o->cx = cx;
  o->cy = cy;
  o->fx = fx;
  o->fy = fy;
  o->distThres = dist_thresh;
  o->angleThres = angle_thresh;
  o->sobelScale = (float)(1.0 / pow(2.0, 3));
  o->maxDepthDeltaRGB = 0.07f;
  o->maxDepthRGB = 6.0f;
  o->minGrad[0] = 5;
  o->minGrad[1] = 3;
  o->minGrad[2] = 1;
  o->lastICPError = 0;
  o->lastICPCount = (float)(width * height);
  o->lastRGBError = 0;
  o->lastRGBCount = (float)(width * height);
  o->lastSO3Error = 0;
  o->lastSO3Count = (float)(width * height);
  memset(o->lastA, 0, sizeof(o->lastA));
  memset(o->lastb, 0, sizeof(o->lastb));
  for (int i = 0; i < NUM_PYRS; ++i) {
    int r = height >> i, c = width >> i;
    o->prow[i] = r;
    o->pcol[i] = c;
    size_t n = (size_t)r * c;
    o->depth_tmp[i].assign(n, 0);
    // the reference's cudaMalloc'd maps are uninitialised; NaN-fill so stale reads are well defined
    o->vmaps_g_prev[i].assign(3 * n, qnan());
    o->nmaps_g_prev[i].assign(3 * n, qnan());
    o->vmaps_curr[i].assign(3 * n, qnan());
    o->nmaps_curr[i].assign(3 * n, qnan());
    o->lastDepth[i].assign(n, qnan());
    o->nextDepth[i].assign(n, qnan());
    o->lastImage[i].assign(n, 0);
    o->nextImage[i].assign(n, 0);
    o->lastNextImage[i].assign(n, 0);
    o->dIdx[i].assign(n, 0);
    o->dIdy[i].assign(n, 0);
    o->corresImg[i].resize(n);
    o->pointClouds[i].assign(3 * n, 0.f);
  }
  o->vmaps_tmp.assign((size_t)width * height * 4, 0.f);
  return o;
}

extern "C" void efo_odom_destroy(EfoOdometry* o) { delete o; }

// RGBDOdometry.cpp:121-147 initICP(GPUTexture* filteredDepth, depthCutoff) — does NOT touch vmaps_tmp (App. A-2)
extern "C" void efo_odom_init_icp_depth(EfoOdometry* o, const uint16_t* filtered_depth, float depth_cutoff) {
  memcpy(o->depth_tmp[0].data(), filtered_depth, sizeof(uint16_t) * o->width * o->height);
  for (int i = 1; i < NUM_PYRS; ++i)
    efo_pyr_down_u16(o->depth_tmp[i - 1].data(), o->prow[i - 1], o->pcol[i - 1], o->depth_tmp[i].data());
  for (int i = 0; i < NUM_PYRS; ++i) {
    float fx, fy, cx, cy;
    intr_level(o, i, fx, fy, cx, cy);
    efo_create_vmap(o->depth_tmp[i].data(), o->prow[i], o->pcol[i], fx, fy, cx, cy, depth_cutoff,
                    o->vmaps_curr[i].data());
    efo_create_nmap(o->vmaps_curr[i].data(), o->prow[i], o->pcol[i], o->nmaps_curr[i].data());
  }
}

// RGBDOdometry.cpp:149-169 initICP(predictedVertices, predictedNormals)
extern "C" void efo_odom_init_icp_pred(EfoOdometry* o, const float* vtx4, const float* nrm4) {
  efo_copy_maps(vtx4, nrm4, o->height, o->width, o->vmaps_tmp.data(), o->vmaps_curr[0].data(),
                o->nmaps_curr[0].data());
  for (int i = 1; i < NUM_PYRS; ++i) {
    efo_resize_map(o->vmaps_curr[i - 1].data(), o->prow[i - 1], o->pcol[i - 1], o->vmaps_curr[i].data(), 0);
    efo_resize_map(o->nmaps_curr[i - 1].data(), o->prow[i - 1], o->pcol[i - 1], o->nmaps_curr[i].data(), 1);
  }
}

// RGBDOdometry.cpp:171-210 initICPModel
extern "C" void efo_odom_init_icp_model(EfoOdometry* o, const float* vtx4, const float* nrm4, const double* T) {
  efo_copy_maps(vtx4, nrm4, o->height, o->width, o->vmaps_tmp.data(), o->vmaps_g_prev[0].data(),
                o->nmaps_g_prev[0].data());
  for (int i = 1; i < NUM_PYRS; ++i) {
    efo_resize_map(o->vmaps_g_prev[i - 1].data(), o->prow[i - 1], o->pcol[i - 1], o->vmaps_g_prev[i].data(), 0);
    efo_resize_map(o->nmaps_g_prev[i - 1].data(), o->prow[i - 1], o->pcol[i - 1], o->nmaps_g_prev[i].data(), 1);
  }
  float R[9], t[3];
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) R[r * 3 + c] = (float)T[r * 4 + c];
    t[r] = (float)T[r * 4 + 3];
  }
  for (int i = 0; i < NUM_PYRS; ++i)
    efo_transform_maps(o->vmaps_g_prev[i].data(), o->nmaps_g_prev[i].data(), o->prow[i], o->pcol[i], R, t);
}

// RGBDOdometry.cpp:212-234 populateRGBDData
static void populate_rgbd(EfoOdometry* o, const uint8_t* rgba, std::vector<float>* destDepths,
                          std::vector<uint8_t>* destImages) {
  efo_vertices_to_depth(o->vmaps_tmp.data(), o->height, o->width, o->maxDepthRGB, destDepths[0].data());
  for (int i = 0; i + 1 < NUM_PYRS; i++)
    efo_pyr_down_gauss_f(destDepths[i].data(), o->prow[i], o->pcol[i], destDepths[i + 1].data());
  efo_rgba_to_intensity(rgba, o->height, o->width, destImages[0].data());
  for (int i = 0; i + 1 < NUM_PYRS; i++)
    efo_pyr_down_u8(destImages[i].data(), o->prow[i], o->pcol[i], destImages[i + 1].data());
}

extern "C" void efo_odom_init_rgb_model(EfoOdometry* o, const uint8_t* rgba) {
  populate_rgbd(o, rgba, o->lastDepth, o->lastImage);
}
extern "C" void efo_odom_init_rgb(EfoOdometry* o, const uint8_t* rgba) {
  populate_rgbd(o, rgba, o->nextDepth, o->nextImage);
}
// RGBDOdometry.cpp:246-257
extern "C" void efo_odom_init_first_rgb(EfoOdometry* o, const uint8_t* rgba) {
  efo_rgba_to_intensity(rgba, o->height, o->width, o->lastNextImage[0].data());
  for (int i = 0; i + 1 < NUM_PYRS; i++)
    efo_pyr_down_u8(o->lastNextImage[i].data(), o->prow[i], o->pcol[i], o->lastNextImage[i + 1].data());
}

// RGBDOdometry.cpp:259-571 getIncrementalTransformation
extern "C" int efo_odom_track(EfoOdometry* o, double* T_wc, int rgbOnly, float icpWeight, int pyramid, int fastOdom,
                              int so3, EfoTrace* trace, int max_trace) {
  int ntrace = 0;
  const bool icp = !rgbOnly && icpWeight > 0;
  const bool rgb = rgbOnly || icpWeight < 100;

  float Rprev[9], tprev[3], Rcurr[9], tcurr[3];
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) Rprev[r * 3 + c] = (float)T_wc[r * 4 + c];
    tprev[r] = (float)T_wc[r * 4 + 3];
  }
  memcpy(Rcurr, Rprev, sizeof(Rprev));
  memcpy(tcurr, tprev, sizeof(tprev));

  if (rgb)
    for (int i = 0; i < NUM_PYRS; i++)
      efo_sobel(o->nextImage[i].data(), o->prow[i], o->pcol[i], o->dIdx[i].data(), o->dIdy[i].data());

  double resultR[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};

  if (so3) {
    const int pyramidLevel = 2;
    float R_lr[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    float lfx, lfy, lcx, lcy;
    intr_level(o, pyramidLevel, lfx, lfy, lcx, lcy);
    double K[9] = {lfx, 0, lcx, 0, lfy, lcy, 0, 0, 1};
    double Kinv[9];
    la::inv3(K, Kinv);
    float lastError = std::numeric_limits<float>::max() / 2;
    float lastCount = std::numeric_limits<float>::max() / 2;
    double lastResultR[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};

    for (int i = 0; i < 10; i++) {
      float jtj[9], jtr[3];
      double tmp[9], homography[9], K_R_lr[9];
      la::mul3(K, resultR, tmp);
      la::mul3(tmp, Kinv, homography);
      la::mul3(K, resultR, K_R_lr);
      float imageBasis[9], kinv[9], krlr[9];
      for (int k = 0; k < 9; ++k) {
        imageBasis[k] = (float)homography[k];
        kinv[k] = (float)Kinv[k];
        krlr[k] = (float)K_R_lr[k];
      }
      float residual[2];
      efo_so3_step(o->lastNextImage[pyramidLevel].data(), o->nextImage[pyramidLevel].data(), imageBasis, kinv, krlr,
                   o->prow[pyramidLevel], o->pcol[pyramidLevel], jtj, jtr, residual);
      if (trace && ntrace < max_trace) {
        EfoTrace& t = trace[ntrace++];
        memset(&t, 0, sizeof(t));
        t.kind = 1;
        t.level = pyramidLevel;
        t.iter = i;
        memcpy(t.A_so3, jtj, sizeof(jtj));
        memcpy(t.b_so3, jtr, sizeof(jtr));
        t.so3_residual[0] = residual[0];
        t.so3_residual[1] = residual[1];
      }
      o->lastSO3Error = sqrtf(residual[0]) / residual[1];
      o->lastSO3Count = residual[1];

      if (o->lastSO3Error < lastError && lastCount == o->lastSO3Count) {
        break;
      } else if (o->lastSO3Error > lastError + 0.001)  // double comparison, as in the reference
      {
        o->lastSO3Error = lastError;
        o->lastSO3Count = lastCount;
        memcpy(resultR, lastResultR, sizeof(resultR));
        break;
      }
      lastError = o->lastSO3Error;
      lastCount = o->lastSO3Count;
      memcpy(lastResultR, resultR, sizeof(resultR));

      float delta[3];
      la::solve_sym3f(jtj, jtr, delta);
      double dd[3] = {delta[0], delta[1], delta[2]};
      double rotUpdate[9];
      la::rodrigues(dd, rotUpdate);
      float ru[9], nr[9];
      for (int k = 0; k < 9; ++k) ru[k] = (float)rotUpdate[k];
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) {
          // Eigen float product: plain left-to-right accumulation of three terms
          nr[r * 3 + c] = ru[r * 3 + 0] * R_lr[0 * 3 + c] + ru[r * 3 + 1] * R_lr[1 * 3 + c] + ru[r * 3 + 2] * R_lr[2 * 3 + c];
        }
      memcpy(R_lr, nr, sizeof(nr));
      for (int k = 0; k < 9; ++k) resultR[k] = R_lr[k];
    }
  }

  int iterations[NUM_PYRS];
  iterations[0] = fastOdom ? 3 : 10;
  iterations[1] = pyramid ? 5 : 0;
  iterations[2] = pyramid ? 4 : 0;

  float Rprev_inv[9];
  la::inv3f(Rprev, Rprev_inv);

  double resultRt[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
  if (so3)
    for (int x = 0; x < 3; x++)
      for (int y = 0; y < 3; y++) resultRt[x * 4 + y] = resultR[x * 3 + y];

  for (int i = NUM_PYRS - 1; i >= 0; i--) {
    float lfx, lfy, lcx, lcy;
    intr_level(o, i, lfx, lfy, lcx, lcy);
    if (rgb)
      efo_project_points(o->lastDepth[i].data(), o->prow[i], o->pcol[i], lfx, lfy, lcx, lcy, o->pointClouds[i].data());

    double K[9] = {lfx, 0, lcx, 0, lfy, lcy, 0, 0, 1};
    double Kinv[9];
    la::inv3(K, Kinv);

    o->lastRGBError = std::numeric_limits<float>::max();

    for (int j = 0; j < iterations[i]; j++) {
      double Rt[16];
      la::inv4(resultRt, Rt);
      double R[9] = {Rt[0], Rt[1], Rt[2], Rt[4], Rt[5], Rt[6], Rt[8], Rt[9], Rt[10]};
      double tmp[9], KRK_inv[9];
      la::mul3(K, R, tmp);
      la::mul3(tmp, Kinv, KRK_inv);
      float krkInv[9];
      for (int k = 0; k < 9; ++k) krkInv[k] = (float)KRK_inv[k];
      double Ktv[3] = {Rt[3], Rt[7], Rt[11]};
      double Kt[3];
      la::mulv3(K, Ktv, Kt);
      float kt[3] = {(float)Kt[0], (float)Kt[1], (float)Kt[2]};

      int sigma = 0, rgbSize = 0;
      if (rgb) {
        float minScale = (float)(pow((double)o->minGrad[i], 2.0) / pow((double)o->sobelScale, 2.0));
        efo_rgb_residual(minScale, o->dIdx[i].data(), o->dIdy[i].data(), o->lastDepth[i].data(),
                         o->nextDepth[i].data(), o->lastImage[i].data(), o->nextImage[i].data(),
                         o->corresImg[i].data(), o->maxDepthDeltaRGB, kt, krkInv, o->prow[i], o->pcol[i], &sigma,
                         &rgbSize);
      }

      // RGBDOdometry.cpp:442 — precedence quirk (App. A-1): sqrt( ((float)sigma / rgbSize == 0) ? 1 : rgbSize )
      float sigmaVal = std::sqrt((float)sigma / rgbSize == 0 ? 1 : rgbSize);
      float rgbError = (float)(std::sqrt((double)sigma) / (rgbSize == 0 ? 1 : rgbSize));

      if (rgbOnly && rgbError > o->lastRGBError) break;

      o->lastRGBError = rgbError;
      o->lastRGBCount = (float)rgbSize;
      if (rgbOnly) sigmaVal = -1;

      float A_icp[36], b_icp[6], A_rgbd[36], b_rgbd[6];
      memset(A_icp, 0, sizeof(A_icp));
      memset(b_icp, 0, sizeof(b_icp));
      memset(A_rgbd, 0, sizeof(A_rgbd));
      memset(b_rgbd, 0, sizeof(b_rgbd));
      float residual[2] = {0, 0};  // uninitialised in the reference when !icp (App. A-3)

      if (icp)
        efo_icp_step(Rcurr, tcurr, o->vmaps_curr[i].data(), o->nmaps_curr[i].data(), Rprev_inv, tprev, lfx, lfy, lcx,
                     lcy, o->vmaps_g_prev[i].data(), o->nmaps_g_prev[i].data(), o->distThres, o->angleThres,
                     o->prow[i], o->pcol[i], A_icp, b_icp, residual);

      o->lastICPError = sqrtf(residual[0]) / residual[1];
      o->lastICPCount = residual[1];

      if (rgb)
        efo_rgb_step(o->corresImg[i].data(), sigmaVal, o->pointClouds[i].data(), lfx, lfy, o->dIdx[i].data(),
                     o->dIdy[i].data(), o->sobelScale, o->prow[i], o->pcol[i], A_rgbd, b_rgbd);

      double result[6];
      if (icp && rgb) {
        double w = icpWeight;
        for (int k = 0; k < 36; ++k) o->lastA[k] = (double)A_rgbd[k] + w * w * (double)A_icp[k];
        for (int k = 0; k < 6; ++k) o->lastb[k] = (double)b_rgbd[k] + w * (double)b_icp[k];
      } else if (icp) {
        for (int k = 0; k < 36; ++k) o->lastA[k] = A_icp[k];
        for (int k = 0; k < 6; ++k) o->lastb[k] = b_icp[k];
      } else {
        for (int k = 0; k < 36; ++k) o->lastA[k] = A_rgbd[k];
        for (int k = 0; k < 6; ++k) o->lastb[k] = b_rgbd[k];
      }
      la::solve_sym6(o->lastA, o->lastb, result);

      if (trace && ntrace < max_trace) {
        EfoTrace& t = trace[ntrace++];
        memset(&t, 0, sizeof(t));
        t.kind = 0;
        t.level = i;
        t.iter = j;
        t.rgb_count = rgbSize;
        t.rgb_sigma = sigma;
        t.sigma_val = sigmaVal;
        memcpy(t.A_icp, A_icp, sizeof(A_icp));
        memcpy(t.b_icp, b_icp, sizeof(b_icp));
        t.icp_residual[0] = residual[0];
        t.icp_residual[1] = residual[1];
        memcpy(t.A_rgb, A_rgbd, sizeof(A_rgbd));
        memcpy(t.b_rgb, b_rgbd, sizeof(b_rgbd));
        memcpy(t.lastA, o->lastA, sizeof(t.lastA));
        memcpy(t.lastb, o->lastb, sizeof(t.lastb));
        memcpy(t.result, result, sizeof(result));
      }

      // OdometryProvider::computeUpdateSE3 (OdometryProvider.h:73-96)
      double rvec[3] = {result[3], result[4], result[5]};
      double Rd[9];
      la::rodrigues(rvec, Rd);
      double upd[16] = {Rd[0], Rd[1], Rd[2], result[0], Rd[3], Rd[4], Rd[5], result[1],
                        Rd[6], Rd[7], Rd[8], result[2], 0,     0,     0,     1};
      double nrt[16];
      la::mul4(upd, resultRt, nrt);
      memcpy(resultRt, nrt, sizeof(nrt));

      // rgbOdom (float isometry) and currentT = T_prev * rgbOdom^-1 (float), RGBDOdometry.cpp:543-551
      float oR[9], ot[3];
      for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) oR[r * 3 + c] = (float)resultRt[r * 4 + c];
        ot[r] = (float)resultRt[r * 4 + 3];
      }
      // Isometry inverse: R^T, -R^T t
      float iR[9], it[3];
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) iR[r * 3 + c] = oR[c * 3 + r];
      for (int r = 0; r < 3; ++r) it[r] = -(iR[r * 3 + 0] * ot[0] + iR[r * 3 + 1] * ot[1] + iR[r * 3 + 2] * ot[2]);
      // affine product: R = Rprev*iR ; t = Rprev*it + tprev
      for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c)
          Rcurr[r * 3 + c] = Rprev[r * 3 + 0] * iR[0 * 3 + c] + Rprev[r * 3 + 1] * iR[1 * 3 + c] + Rprev[r * 3 + 2] * iR[2 * 3 + c];
        tcurr[r] = (Rprev[r * 3 + 0] * it[0] + Rprev[r * 3 + 1] * it[1] + Rprev[r * 3 + 2] * it[2]) + tprev[r];
      }
    }
  }

  if (rgb) {
    float dx = tcurr[0] - tprev[0], dy = tcurr[1] - tprev[1], dz = tcurr[2] - tprev[2];
    if (sqrtf(dx * dx + dy * dy + dz * dz) > 0.3) {
      memcpy(Rcurr, Rprev, sizeof(Rprev));
      memcpy(tcurr, tprev, sizeof(tprev));
    }
  }

  if (so3)
    for (int i = 0; i < NUM_PYRS; i++) std::swap(o->lastNextImage[i], o->nextImage[i]);

  // JacobiSVD U*V^T == orthogonal polar factor of Rcurr (RGBDOdometry.cpp:566-570)
  double Rc[9], Rorth[9];
  for (int k = 0; k < 9; ++k) Rc[k] = Rcurr[k];
  la::polar_orthogonal(Rc, Rorth);
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) T_wc[r * 4 + c] = Rorth[r * 3 + c];
    T_wc[r * 4 + 3] = (double)tcurr[r];
  }
  T_wc[12] = T_wc[13] = T_wc[14] = 0;
  T_wc[15] = 1;
  return ntrace;
}

extern "C" void efo_odom_stats(const EfoOdometry* o, float* out8) {
  out8[0] = o->lastICPError;
  out8[1] = o->lastICPCount;
  out8[2] = o->lastRGBError;
  out8[3] = o->lastRGBCount;
  out8[4] = o->lastSO3Error;
  out8[5] = o->lastSO3Count;
  out8[6] = out8[7] = 0;
}
extern "C" void efo_odom_last_system(const EfoOdometry* o, double* A36, double* b6) {
  memcpy(A36, o->lastA, sizeof(o->lastA));
  memcpy(b6, o->lastb, sizeof(o->lastb));
}
// RGBDOdometry.cpp:573-575 lastA.lu().inverse()
extern "C" void efo_odom_covariance(const EfoOdometry* o, double* cov36) { la::inv_n(o->lastA, cov36, 6); }

extern "C" const void* efo_odom_buffer(const EfoOdometry* o, int which, int level) {
  switch (which) {
    case 0: return o->vmaps_curr[level].data();
    case 1: return o->nmaps_curr[level].data();
    case 2: return o->vmaps_g_prev[level].data();
    case 3: return o->nmaps_g_prev[level].data();
    case 4: return o->lastDepth[level].data();
    case 5: return o->nextDepth[level].data();
    case 6: return o->lastImage[level].data();
    case 7: return o->nextImage[level].data();
    case 8: return o->lastNextImage[level].data();
    case 9: return o->dIdx[level].data();
    case 10: return o->dIdy[level].data();
    case 11: return o->depth_tmp[level].data();
    case 12: return o->corresImg[level].data();
    case 13: return o->pointClouds[level].data();
    case 14: return o->vmaps_tmp.data();
  }
  return nullptr;
}
