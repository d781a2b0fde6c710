// Tracking half of the hot path: pyramid construction and the coarse-to-fine geometric + photometric
// Gauss-Newton loop, as hand-written sm_100a kernels.
//
// Behavioural specification = the reference's Core/Cuda/{cudafuncs,reduce}.cu and Core/Utils/RGBDOdometry.cpp
// (cited per kernel). Structure is B200-first rather than a translation:
//   * every reduction is ONE launch: per-CTA warp-shuffle tree -> per-CTA partial in HBM/L2 -> the CTA that takes
//     the last ticket sums the partials in double and finishes the job (the reference uses 64 CTAs + a second
//     1-CTA kernel + cudaDeviceSynchronize + a blocking D2H per step, reduce.cu:378-386);
//   * the 6x6 / 3x3 solves, the SE(3) update and the next iteration's warp matrices are computed by that last
//     CTA, so the whole SO(3) + 19-iteration SE(3) schedule is a stream of launches with no host round trip;
//   * geometric and photometric systems of one iteration are reduced by the same launch;
//   * grids are sized from the SM count (148 on B200), not the reference's fixed 64 CTAs (types.cuh:64-65);
//   * the back-projected point cloud is recomputed in the photometric step instead of being stored.
#include <float.h>
#include <stddef.h>
#include <stdio.h>

#include "ef_device.cuh"
#include "ef_dmath.cuh"
#include "ef_internal.h"

using namespace ef;

// =============================================================================================
// pyramid / image kernels
// =============================================================================================

// reference pyrDownGaussKernel, cudafuncs.cu:75-121 (sigma_color 30, centre-relative gate, truncating store)
__global__ void k_pyr_down_u16(const uint16_t* __restrict__ src, int srows, int scols, uint16_t* __restrict__ dst) {
  const int drows = srows / 2, dcols = scols / 2;
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= dcols || y >= drows) return;
  const int D = 5;
  const float sigma_color = 30.f;
  const float weights[3] = {0.375f, 0.25f, 0.0625f};
  const int center = src[(2 * y) * scols + 2 * x];
  const int x_mi = max(0, 2 * x - D / 2) - 2 * x;
  const int y_mi = max(0, 2 * y - D / 2) - 2 * y;
  const int x_ma = min(scols, 2 * x - D / 2 + D) - 2 * x;
  const int y_ma = min(srows, 2 * y - D / 2 + D) - 2 * y;
  float sum = 0, wall = 0;
  for (int yi = y_mi; yi < y_ma; ++yi)
    for (int xi = x_mi; xi < x_ma; ++xi) {
      const int val = src[(2 * y + yi) * scols + 2 * x + xi];
      if ((float)abs(val - center) < 3 * sigma_color) {
        sum += val * weights[abs(xi)] * weights[abs(yi)];
        wall += weights[abs(xi)] * weights[abs(yi)];
      }
    }
  dst[y * dcols + x] = (uint16_t)__float2int_rz(sum / wall);
}

// createVMap + createNMap fused (cudafuncs.cu:123-219): the normal is built from the three depths it needs with the
// same arithmetic createVMap would have used, so vmap never has to be re-read. Invalid -> NaN (x flags validity).
__device__ __forceinline__ bool vertex_from_depth(const uint16_t* depth, int cols, int u, int v, float fx_inv, float fy_inv,
                                                  float cx, float cy, float cutoff, f3& out) {
  const float z = depth[v * cols + u] / 1000.f;
  if (z != 0 && z < cutoff) {
    out = mk3(z * (u - cx) * fx_inv, z * (v - cy) * fy_inv, z);
    return true;
  }
  return false;
}

__global__ void k_vmap_nmap(const uint16_t* __restrict__ depth, int rows, int cols, float fx_inv, float fy_inv, float cx,
                            float cy, float cutoff, float* __restrict__ vmap, float* __restrict__ nmap) {
  const int u = blockIdx.x * blockDim.x + threadIdx.x;
  const int v = blockIdx.y * blockDim.y + threadIdx.y;
  if (u >= cols || v >= rows) return;
  const size_t plane = (size_t)rows * cols;
  const size_t p = (size_t)v * cols + u;
  f3 v00;
  const bool ok00 = vertex_from_depth(depth, cols, u, v, fx_inv, fy_inv, cx, cy, cutoff, v00);
  if (ok00) {
    vmap[p] = v00.x;
    vmap[p + plane] = v00.y;
    vmap[p + 2 * plane] = v00.z;
  } else {
    vmap[p] = qnan();
    vmap[p + plane] = qnan();
    vmap[p + 2 * plane] = qnan();
  }
  f3 n = mk3(qnan(), qnan(), qnan());
  if (ok00 && u != cols - 1 && v != rows - 1) {
    f3 v01, v10;
    if (vertex_from_depth(depth, cols, u + 1, v, fx_inv, fy_inv, cx, cy, cutoff, v01) &&
        vertex_from_depth(depth, cols, u, v + 1, fx_inv, fy_inv, cx, cy, cutoff, v10))
      n = normalized(cross(v01 - v00, v10 - v00));
  }
  nmap[p] = n.x;
  nmap[p + plane] = n.y;
  nmap[p + 2 * plane] = n.z;
}

// copyMaps (cudafuncs.cu:295-381): predicted float4 vertex/normal maps -> vmaps_tmp (AoS copy) + SoA planes; z==0 -> NaN
// (vtxB, nrmB, dense_flag): optional fill-in alternates chosen on the device when *dense_flag == 0
// (ElasticFusion.cpp:304-313 decides this on the host after a glReadPixels)
__global__ void k_copy_maps(const float4* __restrict__ vtxA, const float4* __restrict__ nrmA, const float4* __restrict__ vtxB,
                            const float4* __restrict__ nrmB, const int* __restrict__ dense_flag, int rows, int cols,
                            float4* __restrict__ vmaps_tmp, float* __restrict__ vmap, float* __restrict__ nmap) {
  const size_t n = (size_t)rows * cols;
  const bool useB = dense_flag && (*dense_flag == 0);
  const float4* __restrict__ vtx = useB ? vtxB : vtxA;
  const float4* __restrict__ nrm = useB ? nrmB : nrmA;
  for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (size_t)gridDim.x * blockDim.x) {
    const float4 vs = vtx[p];
    const float4 ns = nrm[p];
    if (vmaps_tmp) vmaps_tmp[p] = vs;
    f3 vd = mk3(qnan(), qnan(), qnan()), nd = vd;
    if (!(vs.z == 0)) {
      vd = mk3(vs.x, vs.y, vs.z);
      nd = mk3(ns.x, ns.y, ns.z);
    }
    vmap[p] = vd.x;
    vmap[p + n] = vd.y;
    vmap[p + 2 * n] = vd.z;
    nmap[p] = nd.x;
    nmap[p + n] = nd.y;
    nmap[p + 2 * n] = nd.z;
  }
}

// resizeVMap + resizeNMap (cudafuncs.cu:413-490) in one launch: 2x2 box average, any NaN -> NaN, normals renormalised
__global__ void k_resize_maps(const float* __restrict__ vin, const float* __restrict__ nin, int srows, int scols,
                              float* __restrict__ vout, float* __restrict__ nout) {
  const int drows = srows / 2, dcols = scols / 2;
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= dcols || y >= drows) return;
  const size_t splane = (size_t)srows * scols, dplane = (size_t)drows * dcols;
  const size_t q = (size_t)y * dcols + x;
  const int xs = x * 2, ys = y * 2;
#pragma unroll
  for (int which = 0; which < 2; ++which) {
    const float* in = which ? nin : vin;
    float* out = which ? nout : vout;
    const float x00 = in[(size_t)(ys + 0) * scols + xs + 0], x01 = in[(size_t)(ys + 0) * scols + xs + 1];
    const float x10 = in[(size_t)(ys + 1) * scols + xs + 0], x11 = in[(size_t)(ys + 1) * scols + xs + 1];
    if (isnan(x00) || isnan(x01) || isnan(x10) || isnan(x11)) {
      out[q] = qnan();
      out[q + dplane] = qnan();
      out[q + 2 * dplane] = qnan();
      continue;
    }
    f3 n;
    n.x = (x00 + x01 + x10 + x11) / 4;
    const float* iy = in + splane;
    n.y = (iy[(size_t)(ys + 0) * scols + xs + 0] + iy[(size_t)(ys + 0) * scols + xs + 1] + iy[(size_t)(ys + 1) * scols + xs + 0] +
           iy[(size_t)(ys + 1) * scols + xs + 1]) / 4;
    const float* iz = in + 2 * splane;
    n.z = (iz[(size_t)(ys + 0) * scols + xs + 0] + iz[(size_t)(ys + 0) * scols + xs + 1] + iz[(size_t)(ys + 1) * scols + xs + 0] +
           iz[(size_t)(ys + 1) * scols + xs + 1]) / 4;
    if (which) n = normalized(n);
    out[q] = n.x;
    out[q + dplane] = n.y;
    out[q + 2 * dplane] = n.z;
  }
}

// tranformMaps (cudafuncs.cu:221-293), all three levels in one launch (blockIdx.y = level), in place, pose from gn->T_wc
struct XformArgs {
  float* v[NUM_PYRS];
  float* n[NUM_PYRS];
  int rows[NUM_PYRS], cols[NUM_PYRS];
};
__global__ void k_transform_maps(XformArgs a, const GNState* __restrict__ gn) {
  const int lv = blockIdx.y;
  const size_t np = (size_t)a.rows[lv] * a.cols[lv];
  float R[9], t[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
#pragma unroll
    for (int c = 0; c < 3; ++c) R[r * 3 + c] = (float)gn->T_wc[r * 4 + c];
    t[r] = (float)gn->T_wc[r * 4 + 3];
  }
  const m33 Rm = load_m33(R);
  const f3 tv = mk3(t[0], t[1], t[2]);
  float* vm = a.v[lv];
  float* nm = a.n[lv];
  for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < np; p += (size_t)gridDim.x * blockDim.x) {
    const float vx = vm[p];
    if (!isnan(vx)) {
      const f3 d = mul(Rm, mk3(vx, vm[p + np], vm[p + 2 * np])) + tv;
      vm[p] = d.x;
      vm[p + np] = d.y;
      vm[p + 2 * np] = d.z;
    }
    const float nx = nm[p];
    if (!isnan(nx)) {
      const f3 d = mul(Rm, mk3(nx, nm[p + np], nm[p + 2 * np]));
      nm[p] = d.x;
      nm[p + np] = d.y;
      nm[p + 2 * np] = d.z;
    }
  }
}

// verticesToDepth + imageBGRToIntensity fused (cudafuncs.cu:564-610): level-0 depth from vmaps_tmp.z (cutoff 6 m),
// level-0 intensity int(0.114 x + 0.299 y + 0.587 z) from the RGBA8 texel. Either output may be NULL.
__global__ void k_depth_intensity_l0(const float4* __restrict__ vmaps_tmp, const uchar4* __restrict__ rgbaA,
                                     const uchar4* __restrict__ rgbaB, const int* __restrict__ dense_flag, int forceB, size_t n,
                                     float cutoff, float* __restrict__ depth, uint8_t* __restrict__ image) {
  const uchar4* __restrict__ rgba = (forceB || (dense_flag && *dense_flag == 0)) ? rgbaB : rgbaA;
  for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (size_t)gridDim.x * blockDim.x) {
    if (depth) {
      const float z = vmaps_tmp[p].z;
      depth[p] = (z > cutoff || z <= 0) ? qnan() : z;
    }
    if (image) {
      const uchar4 s = rgba[p];
      const int value = __float2int_rz((float)s.x * 0.114f + (float)s.y * 0.299f + (float)s.z * 0.587f);
      image[p] = (uint8_t)value;
    }
  }
}

__device__ __constant__ float c_gauss25[25] = {1, 4, 6, 4, 1, 4, 16, 24, 16, 4, 6, 24, 36, 24, 6, 4, 16, 24, 16, 4, 1, 4, 6, 4, 1};

// pyrDownKernelGaussF + pyrDownKernelIntensityGauss fused (cudafuncs.cu:383-411,512-562): window [2x-2, min(2x+3, n-1)),
// flipped/shifted tap index, int normaliser, NaN / zero skipping, truncating u8 store. Either pair may be NULL.
__global__ void k_pyr_down_depth_image(const float* __restrict__ dsrc, float* __restrict__ ddst, const uint8_t* __restrict__ isrc,
                                       uint8_t* __restrict__ idst, int srows, int scols) {
  const int drows = srows / 2, dcols = scols / 2;
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= dcols || y >= drows) return;
  const int D = 5;
  const int tx = min(2 * x - D / 2 + D, scols - 1);
  const int ty = min(2 * y - D / 2 + D, srows - 1);
  const int cy0 = max(0, 2 * y - D / 2), cx0 = max(0, 2 * x - D / 2);
  if (dsrc) {
    float sum = 0;
    int count = 0;
    for (int cy = cy0; cy < ty; ++cy)
      for (int cx = cx0; cx < tx; ++cx) {
        const float s = dsrc[(size_t)cy * scols + cx];
        if (!isnan(s)) {
          const float w = c_gauss25[(ty - cy - 1) * 5 + (tx - cx - 1)];
          sum += s * w;
          count = __float2int_rz((float)count + w);
        }
      }
    ddst[(size_t)y * dcols + x] = (float)(sum / (float)count);
  }
  if (isrc) {
    float sum = 0;
    int count = 0;
    for (int cy = cy0; cy < ty; ++cy)
      for (int cx = cx0; cx < tx; ++cx) {
        const uint8_t s = isrc[(size_t)cy * scols + cx];
        if (s > 0) {
          const float w = c_gauss25[(ty - cy - 1) * 5 + (tx - cx - 1)];
          sum += s * w;
          count = __float2int_rz((float)count + w);
        }
      }
    int v = __float2int_rz(sum / (float)count);  // NaN -> 0
    idst[(size_t)y * dcols + x] = (uint8_t)min(max(v, 0), 255);
  }
}

// computeDerivativeImages / applyKernel (cudafuncs.cu:612-668) for all three levels in one launch
struct SobelArgs {
  const uint8_t* src[NUM_PYRS];
  int16_t* dx[NUM_PYRS];
  int16_t* dy[NUM_PYRS];
  int rows[NUM_PYRS], cols[NUM_PYRS];
};
__global__ void k_sobel(SobelArgs a) {
  const int lv = blockIdx.y;
  const int rows = a.rows[lv], cols = a.cols[lv];
  const size_t np = (size_t)rows * cols;
  const float gsx[9] = {(float)0.52201, (float)0.00000, (float)-0.52201, (float)0.79451, (float)-0.00000,
                        (float)-0.79451, (float)0.52201, (float)0.00000, (float)-0.52201};
  const float gsy[9] = {(float)0.52201, (float)0.79451, (float)0.52201, (float)0.00000, (float)0.00000,
                        (float)0.00000, (float)-0.52201, (float)-0.79451, (float)-0.52201};
  const uint8_t* src = a.src[lv];
  for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < np; p += (size_t)gridDim.x * blockDim.x) {
    const int y = (int)(p / cols), x = (int)(p - (size_t)y * cols);
    float dxVal = 0, dyVal = 0;
    int kernelIndex = 8;
    for (int j = max(y - 1, 0); j <= min(y + 1, rows - 1); j++)
      for (int i = max(x - 1, 0); i <= min(x + 1, cols - 1); i++) {
        const float s = (float)src[(size_t)j * cols + i];
        dxVal += s * gsx[kernelIndex];
        dyVal += s * gsy[kernelIndex];
        --kernelIndex;
      }
    a.dx[lv][p] = (int16_t)__float2int_rz(dxVal);
    a.dy[lv][p] = (int16_t)__float2int_rz(dyVal);
  }
}

// =============================================================================================
// Gauss-Newton state machine (device side)
// =============================================================================================

__device__ __forceinline__ void level_intr(const GNState* gn, int level, float& fx, float& fy, float& cx, float& cy) {
  const int div = 1 << level;  // CameraModel::operator()(level), reference types.cuh:92-95
  fx = gn->fx / div;
  fy = gn->fy / div;
  cx = gn->cx / div;
  cy = gn->cy / div;
}

// warp matrices for the next photometric residual pass (RGBDOdometry.cpp:407-417)
__device__ void gn_prepare_warp(GNState* gn, int level) {
  float lfx, lfy, lcx, lcy;
  level_intr(gn, level, lfx, lfy, lcx, lcy);
  double K[9] = {lfx, 0, lcx, 0, lfy, lcy, 0, 0, 1};
  double Kinv[9], Rt[16];
  efm::inv3<double>(K, Kinv);
  efm::inv_n<4>(gn->resultRt, Rt);
  double R[9] = {Rt[0], Rt[1], Rt[2], Rt[4], Rt[5], Rt[6], Rt[8], Rt[9], Rt[10]};
  double tmp[9], KRK_inv[9];
  efm::mul3(K, R, tmp);
  efm::mul3(tmp, Kinv, KRK_inv);
  for (int k = 0; k < 9; ++k) gn->krkinv[k] = (float)KRK_inv[k];
  double tv[3] = {Rt[3], Rt[7], Rt[11]}, Kt[3];
  efm::mulv3(K, tv, Kt);
  for (int k = 0; k < 3; ++k) gn->kt[k] = (float)Kt[k];
}

// homography etc. for the next SO3 pass (RGBDOdometry.cpp:309-321)
__device__ void so3_prepare(GNState* gn) {
  float lfx, lfy, lcx, lcy;
  level_intr(gn, 2, lfx, lfy, lcx, lcy);
  double K[9] = {lfx, 0, lcx, 0, lfy, lcy, 0, 0, 1};
  double Kinv[9], tmp[9], H[9];
  efm::inv3<double>(K, Kinv);
  efm::mul3(K, gn->resultR, tmp);
  efm::mul3(tmp, Kinv, H);
  for (int k = 0; k < 9; ++k) {
    gn->imageBasis[k] = (float)H[k];
    gn->kinv[k] = (float)Kinv[k];
    gn->krlr[k] = (float)tmp[k];
  }
}

// start of getIncrementalTransformation (RGBDOdometry.cpp:266-273,284-303)
__global__ void k_gn_begin(GNState* gn, int rgbOnly, float icpWeight, int so3) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  gn->rgbOnly = rgbOnly;
  gn->icpWeight = icpWeight;
  gn->icp = (!rgbOnly && icpWeight > 0) ? 1 : 0;
  gn->rgb = (rgbOnly || icpWeight < 100) ? 1 : 0;
  gn->so3 = so3;
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) gn->Rprev[r * 3 + c] = (float)gn->T_wc[r * 4 + c];
    gn->tprev[r] = (float)gn->T_wc[r * 4 + 3];
  }
  for (int k = 0; k < 9; ++k) gn->Rcurr[k] = gn->Rprev[k];
  for (int k = 0; k < 3; ++k) gn->tcurr[k] = gn->tprev[k];
  efm::inv3<float>(gn->Rprev, gn->Rprev_inv);
  const double I3[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  for (int k = 0; k < 9; ++k) {
    gn->resultR[k] = I3[k];
    gn->lastResultR[k] = I3[k];
    gn->R_lr[k] = (float)I3[k];
  }
  gn->so3_lastError = FLT_MAX / 2;
  gn->so3_lastCount = FLT_MAX / 2;
  gn->so3_done = 0;
  gn->break_level = -1;
  gn->trace_n = 0;
  if (so3) so3_prepare(gn);
}

// after the SO3 loop: seed resultRt (RGBDOdometry.cpp:379-388) and prepare the first SE3 iteration
__global__ void k_gn_seed(GNState* gn, int first_level) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  for (int k = 0; k < 16; ++k) gn->resultRt[k] = (k % 5 == 0) ? 1.0 : 0.0;
  if (gn->so3)
    for (int x = 0; x < 3; x++)
      for (int y = 0; y < 3; y++) gn->resultRt[x * 4 + y] = gn->resultR[x * 3 + y];
  gn->lastRGBError = FLT_MAX;
  if (!gn->rgb) {
    gn->rgbSize = 0;
    gn->sigma = 0;
    gn->sigmaVal = 0.f;  // sqrt((0.f/0 == 0) ? 1 : 0)
    gn->lastRGBError = 0.f;
    gn->lastRGBCount = 0.f;
  }
  gn_prepare_warp(gn, first_level);
}

// end of getIncrementalTransformation (RGBDOdometry.cpp:555-570) + velocity weighting (ElasticFusion.cpp:369-383)
__global__ void k_gn_finish(GNState* gn, float weightMultiplier, int have_track) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double Tprev[16];
  for (int k = 0; k < 16; ++k) Tprev[k] = gn->T_wc[k];
  if (have_track) {
    if (gn->rgb) {
      const float dx = gn->tcurr[0] - gn->tprev[0], dy = gn->tcurr[1] - gn->tprev[1], dz = gn->tcurr[2] - gn->tprev[2];
      if (sqrtf(dx * dx + dy * dy + dz * dz) > 0.3) {
        for (int k = 0; k < 9; ++k) gn->Rcurr[k] = gn->Rprev[k];
        for (int k = 0; k < 3; ++k) gn->tcurr[k] = gn->tprev[k];
      }
    }
    double Rc[9], Ro[9];
    for (int k = 0; k < 9; ++k) Rc[k] = gn->Rcurr[k];
    efm::polar_orthogonal(Rc, Ro);
    for (int r = 0; r < 3; ++r) {
      for (int c = 0; c < 3; ++c) gn->T_wc[r * 4 + c] = Ro[r * 3 + c];
      gn->T_wc[r * 4 + 3] = (double)gn->tcurr[r];
    }
    gn->T_wc[12] = gn->T_wc[13] = gn->T_wc[14] = 0;
    gn->T_wc[15] = 1;
  }
  // weighting from T_curr_prev = T_wc_curr^-1 * T_wc_prev; Tprev is reconstructed from Rprev/tprev only when tracking
  // ran — otherwise the caller stored the previous pose in resultRt before overwriting T_wc.
  double inv[16], Tcp[16];
  efm::se3_inverse(gn->T_wc, inv);
  if (!have_track)
    for (int k = 0; k < 16; ++k) Tprev[k] = gn->resultRt[k];
  efm::mul4(inv, Tprev, Tcp);
  const double tn = sqrt(Tcp[3] * Tcp[3] + Tcp[7] * Tcp[7] + Tcp[11] * Tcp[11]);
  const double ln = efm::se3_log_norm(Tcp);
  float weighting = (float)fmax(tn, ln);
  const float largest = 0.01f, minWeight = 0.5f;
  if (weighting > largest) weighting = largest;
  gn->weighting = fmaxf(1.0f - (weighting / largest), minWeight) * weightMultiplier;
}

// k_gn_finish needs the pre-tracking pose; stash it (tracking overwrites T_wc only at the end, so this is only needed
// for the in_T_wc path where the host replaces the pose).
__global__ void k_set_pose(GNState* gn, const double* T_new) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  for (int k = 0; k < 16; ++k) {
    gn->resultRt[k] = gn->T_wc[k];
    gn->T_wc[k] = T_new[k];
  }
}

// unpack the reference's 29-float JtJJtrSE3 into A (6x6, symmetric) and b (reduce.cu:388-400)
__device__ __forceinline__ void unpack_se3(const float* h, float* A, float* b) {
  int shift = 0;
  for (int i = 0; i < 6; ++i)
    for (int j = i; j < 7; ++j) {
      const float value = h[shift++];
      if (j == 6)
        b[i] = value;
      else
        A[j * 6 + i] = A[i * 6 + j] = value;
    }
}

// one SE3 Gauss-Newton update, executed by a single thread of the last CTA (RGBDOdometry.cpp:492-551)
__device__ void gn_update(const OdomDev& od, int level, int iter, int next_level) {
  GNState* gn = od.gn;
  float A_icp[36], b_icp[6], A_rgb[36], b_rgb[6];
  unpack_se3(gn->sum_icp, A_icp, b_icp);
  unpack_se3(gn->sum_rgb, A_rgb, b_rgb);
  const float res0 = gn->sum_icp[27], res1 = gn->sum_icp[28];
  gn->lastICPError = sqrtf(res0) / res1;
  gn->lastICPCount = res1;

  double result[6];
  if (gn->icp && gn->rgb) {
    const double w = gn->icpWeight;
    for (int k = 0; k < 36; ++k) gn->lastA[k] = (double)A_rgb[k] + w * w * (double)A_icp[k];
    for (int k = 0; k < 6; ++k) gn->lastb[k] = (double)b_rgb[k] + w * (double)b_icp[k];
  } else if (gn->icp) {
    for (int k = 0; k < 36; ++k) gn->lastA[k] = A_icp[k];
    for (int k = 0; k < 6; ++k) gn->lastb[k] = b_icp[k];
  } else {
    for (int k = 0; k < 36; ++k) gn->lastA[k] = A_rgb[k];
    for (int k = 0; k < 6; ++k) gn->lastb[k] = b_rgb[k];
  }
  efm::solve_sym6(gn->lastA, gn->lastb, result);

  if (od.trace && gn->trace_n < MAX_TRACE) {
    EfSolveTrace& t = od.trace[gn->trace_n++];
    t.kind = 0;
    t.level = level;
    t.iter = iter;
    t.rgb_count = gn->rgbSize;
    t.rgb_sigma = gn->sigma;
    t.sigma_val = gn->sigmaVal;
    for (int k = 0; k < 36; ++k) {
      t.A_icp[k] = A_icp[k];
      t.A_rgb[k] = A_rgb[k];
      t.lastA[k] = gn->lastA[k];
    }
    for (int k = 0; k < 6; ++k) {
      t.b_icp[k] = b_icp[k];
      t.b_rgb[k] = b_rgb[k];
      t.lastb[k] = gn->lastb[k];
      t.result[k] = result[k];
    }
    t.icp_residual[0] = res0;
    t.icp_residual[1] = res1;
  }

  // OdometryProvider::computeUpdateSE3 (OdometryProvider.h:73-96)
  double rvec[3] = {result[3], result[4], result[5]}, Rd[9];
  efm::rodrigues(rvec, Rd);
  double upd[16] = {Rd[0], Rd[1], Rd[2], result[0], Rd[3], Rd[4], Rd[5], result[1],
                    Rd[6], Rd[7], Rd[8], result[2], 0,     0,     0,     1};
  double nrt[16];
  efm::mul4(upd, gn->resultRt, nrt);
  for (int k = 0; k < 16; ++k) gn->resultRt[k] = nrt[k];

  // currentT = T_prev * rgbOdom^-1 in float (RGBDOdometry.cpp:543-551)
  float oR[9], ot[3], iR[9], it[3];
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) oR[r * 3 + c] = (float)gn->resultRt[r * 4 + c];
    ot[r] = (float)gn->resultRt[r * 4 + 3];
  }
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) iR[r * 3 + c] = oR[c * 3 + r];
  for (int r = 0; r < 3; ++r) it[r] = -(iR[r * 3 + 0] * ot[0] + iR[r * 3 + 1] * ot[1] + iR[r * 3 + 2] * ot[2]);
  const float* Rp = gn->Rprev;
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c)
      gn->Rcurr[r * 3 + c] = Rp[r * 3 + 0] * iR[0 * 3 + c] + Rp[r * 3 + 1] * iR[1 * 3 + c] + Rp[r * 3 + 2] * iR[2 * 3 + c];
    gn->tcurr[r] = (Rp[r * 3 + 0] * it[0] + Rp[r * 3 + 1] * it[1] + Rp[r * 3 + 2] * it[2]) + gn->tprev[r];
  }
  if (next_level >= 0) gn_prepare_warp(gn, next_level);
}

// =============================================================================================
// reductions
// =============================================================================================

// final cross-CTA sum of `nvals` (<=32) floats at `off` inside each CTA's partial; result (float) to dst[0..nvals)
__device__ __forceinline__ void final_sum(const float* partials, int nblocks, int off, int nvals, float* dst, double* dsm) {
  const int v = threadIdx.x & 31, s = threadIdx.x >> 5;  // 8 slices of CTAs
  double acc = 0;
  if (v < nvals)
    for (int b = s; b < nblocks; b += RED_THREADS / 32) acc += (double)partials[(size_t)b * PARTIAL_STRIDE + off + v];
  dsm[s * 32 + v] = acc;
  __syncthreads();
  if (s == 0 && v < nvals) {
    double t = 0;
#pragma unroll
    for (int k = 0; k < RED_THREADS / 32; ++k) t += dsm[k * 32 + v];
    dst[v] = (float)t;
  }
  __syncthreads();
}

// ---- geometric row: ICPReduction::search/getProducts (reduce.cu:224-331) ------------------------------------
struct IcpFrame {
  m33 Rcurr, Rprev_inv;
  f3 tcurr, tprev;
  float fx, fy, cx, cy;
  float distThres, angleThres;
};

__device__ __forceinline__ bool icp_row(const IcpFrame& F, const float* __restrict__ vmap_curr, const float* __restrict__ nmap_curr,
                                        const float* __restrict__ vmap_g_prev, const float* __restrict__ nmap_g_prev, int rows,
                                        int cols, size_t plane, int x, int y, float row[7]) {
  const size_t p = (size_t)y * cols + x;
  const f3 vcurr = mk3(vmap_curr[p], vmap_curr[p + plane], vmap_curr[p + 2 * plane]);
  const f3 vcurr_g = mul(F.Rcurr, vcurr) + F.tcurr;
  const f3 vcurr_cp = mul(F.Rprev_inv, vcurr_g - F.tprev);
  const int ux = __float2int_rn(vcurr_cp.x * F.fx / vcurr_cp.z + F.cx);
  const int uy = __float2int_rn(vcurr_cp.y * F.fy / vcurr_cp.z + F.cy);
  if (ux < 0 || uy < 0 || ux >= cols || uy >= rows || vcurr_cp.z < 0) return false;
  const size_t q = (size_t)uy * cols + ux;
  const f3 vprev_g = mk3(vmap_g_prev[q], vmap_g_prev[q + plane], vmap_g_prev[q + 2 * plane]);
  const f3 ncurr = mk3(nmap_curr[p], nmap_curr[p + plane], nmap_curr[p + 2 * plane]);
  const f3 ncurr_g = mul(F.Rcurr, ncurr);
  const f3 nprev_g = mk3(nmap_g_prev[q], nmap_g_prev[q + plane], nmap_g_prev[q + 2 * plane]);
  const float dist = norm(vprev_g - vcurr_g);
  const float sine = norm(cross(ncurr_g, nprev_g));
  if (!(sine < F.angleThres && dist <= F.distThres && !isnan(ncurr.x) && !isnan(nprev_g.x))) return false;
  const f3 s_cp = mul(F.Rprev_inv, vcurr_g - F.tprev);
  const f3 d_cp = mul(F.Rprev_inv, vprev_g - F.tprev);
  const f3 n_cp = mul(F.Rprev_inv, nprev_g);
  const f3 c = cross(s_cp, n_cp);
  row[0] = n_cp.x;
  row[1] = n_cp.y;
  row[2] = n_cp.z;
  row[3] = c.x;
  row[4] = c.y;
  row[5] = c.z;
  row[6] = dot(n_cp, s_cp - d_cp);
  return true;
}

__device__ __forceinline__ void accumulate29(const float row[7], float (&acc)[29]) {
  int k = 0;
#pragma unroll
  for (int i = 0; i < 6; ++i)
#pragma unroll
    for (int j = i; j < 7; ++j) acc[k++] += row[i] * row[j];
  acc[27] += row[6] * row[6];
  acc[28] += 1.0f;
}

// ---- photometric row: RGBReduction::getProducts (reduce.cu:419-480); cloud point recomputed from lastDepth
//      with projectPointsKernel's arithmetic (cudafuncs.cu:670-688) ----------------------------------------
__device__ __forceinline__ bool rgb_row(const DataTerm& corresp, float sigma, const float* __restrict__ lastDepth, int cols,
                                        float fx, float fy, float cx, float cy, const int16_t* __restrict__ dIdx,
                                        const int16_t* __restrict__ dIdy, float sobelScale, float row[7]) {
  if (!corresp.valid) return false;
  float w = sigma + fabsf(corresp.diff);
  w = w > 1.19209290E-07F ? 1.0f / w : 1.0f;
  if (sigma == -1) w = 1;
  row[6] = -w * corresp.diff;
  const int zx = corresp.zero_x, zy = corresp.zero_y;
  const float z = lastDepth[(size_t)zy * cols + zx];
  const float invFx = 1.0f / fx, invFy = 1.0f / fy;
  const f3 cp = mk3((float)((zx - cx) * z * invFx), (float)((zy - cy) * z * invFy), z);
  const float invz = (float)(1.0 / (double)cp.z);
  const size_t o = (size_t)corresp.one_y * cols + corresp.one_x;
  const float dI_dx_val = w * sobelScale * dIdx[o];
  const float dI_dy_val = w * sobelScale * dIdy[o];
  const float v0 = dI_dx_val * fx * invz;
  const float v1 = dI_dy_val * fy * invz;
  const float v2 = -(v0 * cp.x + v1 * cp.y) * invz;
  row[0] = v0;
  row[1] = v1;
  row[2] = v2;
  row[3] = -cp.z * v1 + cp.y * v2;
  row[4] = cp.z * v0 - cp.x * v2;
  row[5] = -cp.y * v0 + cp.x * v1;
  return true;
}

// One Gauss-Newton step: geometric and/or photometric 6x6 systems reduced in ONE launch, then (solve != 0) the last
// CTA solves, updates the pose and prepares the next iteration. Replaces icpStep + rgbStep + host solve
// (reduce.cu:333-401,502-550; RGBDOdometry.cpp:470-551).
__global__ void __launch_bounds__(RED_THREADS) k_se3_step(OdomDev od, int level, int iter, int next_level, int do_icp, int do_rgb, int solve) {
  GNState* gn = od.gn;
  if (solve && gn->break_level == level) {
    // rgbOnly `break` (RGBDOdometry.cpp:445-447): the rest of this level is skipped, but the first iteration of the
    // next level still needs its warp matrices
    if (blockIdx.x == 0 && threadIdx.x == 0 && next_level >= 0 && next_level != level) gn_prepare_warp(gn, next_level);
    return;
  }
  __shared__ float sred[29 * (RED_THREADS / 32)];
  __shared__ double dsm[(RED_THREADS / 32) * 32];
  const int rows = od.rows[level], cols = od.cols[level];
  const int N = rows * cols;
  const size_t plane = (size_t)N;
  float lfx, lfy, lcx, lcy;
  level_intr(gn, level, lfx, lfy, lcx, lcy);
  float* my_partial = od.partials + (size_t)blockIdx.x * PARTIAL_STRIDE;

  if (do_icp) {
    IcpFrame F;
    F.Rcurr = load_m33(gn->Rcurr);
    F.Rprev_inv = load_m33(gn->Rprev_inv);
    F.tcurr = mk3(gn->tcurr[0], gn->tcurr[1], gn->tcurr[2]);
    F.tprev = mk3(gn->tprev[0], gn->tprev[1], gn->tprev[2]);
    F.fx = lfx;
    F.fy = lfy;
    F.cx = lcx;
    F.cy = lcy;
    F.distThres = od.distThres;
    F.angleThres = od.angleThres;
    float acc[29];
#pragma unroll
    for (int k = 0; k < 29; ++k) acc[k] = 0.f;
    for (int i = blockIdx.x * RED_THREADS + threadIdx.x; i < N; i += gridDim.x * RED_THREADS) {
      const int y = i / cols, x = i - y * cols;
      float row[7];
      if (icp_row(F, od.vmap_curr[level], od.nmap_curr[level], od.vmap_g_prev[level], od.nmap_g_prev[level], rows, cols, plane, x, y, row))
        accumulate29(row, acc);
    }
    block_reduce_sum<29, RED_THREADS>(acc, sred);
    if (threadIdx.x == 0) {
#pragma unroll
      for (int k = 0; k < 29; ++k) my_partial[k] = acc[k];
    }
    __syncthreads();
  }
  if (do_rgb) {
    const float sigma = gn->sigmaVal;
    const DataTerm* corres = od.corres[level];
    float acc[29];
#pragma unroll
    for (int k = 0; k < 29; ++k) acc[k] = 0.f;
    for (int i = blockIdx.x * RED_THREADS + threadIdx.x; i < N; i += gridDim.x * RED_THREADS) {
      const DataTerm c = corres[i];
      float row[7];
      if (rgb_row(c, sigma, od.lastDepth[level], cols, lfx, lfy, lcx, lcy, od.dIdx[level], od.dIdy[level], od.sobelScale, row))
        accumulate29(row, acc);
    }
    block_reduce_sum<29, RED_THREADS>(acc, sred);
    if (threadIdx.x == 0) {
#pragma unroll
      for (int k = 0; k < 29; ++k) my_partial[32 + k] = acc[k];
    }
  }
  if (!last_block_done(od.counter)) return;

  if (do_icp)
    final_sum(od.partials, gridDim.x, 0, 29, gn->sum_icp, dsm);
  else if (threadIdx.x < 32)
    gn->sum_icp[threadIdx.x] = 0.f;
  if (do_rgb)
    final_sum(od.partials, gridDim.x, 32, 29, gn->sum_rgb, dsm);
  else if (threadIdx.x < 32)
    gn->sum_rgb[threadIdx.x] = 0.f;
  __syncthreads();
  if (threadIdx.x == 0) {
    *od.counter = 0;
    if (solve) gn_update(od, level, iter, next_level);
  }
}

// Photometric correspondences + {count, sum diff^2}: RGBResidual / computeRgbResidual (reduce.cu:603-787).
// The last CTA also derives sigma / rgbError exactly as the host does (RGBDOdometry.cpp:442-455, incl. the
// operator-precedence quirk) so the following step launch needs nothing from the host.
__global__ void __launch_bounds__(RED_THREADS) k_rgb_residual(OdomDev od, int level, int iter, int finalize) {
  GNState* gn = od.gn;
  if (finalize && gn->break_level == level) return;
  __shared__ int sred_i[2 * (RED_THREADS / 32)];
  const int rows = od.rows[level], cols = od.cols[level];
  const int N = rows * cols;
  const m33 krkinv = load_m33(gn->krkinv);
  const f3 kt = mk3(gn->kt[0], gn->kt[1], gn->kt[2]);
  const float minScale = od.minScale[level];
  const int16_t* dIdx = od.dIdx[level];
  const int16_t* dIdy = od.dIdy[level];
  const float* lastDepth = od.lastDepth[level];
  const float* nextDepth = od.nextDepth[level];
  const uint8_t* lastImage = od.lastImage[level];
  const uint8_t* nextImage = od.nextImage[level];
  DataTerm* corresImg = od.corres[level];
  unsigned int cnt = 0, sig = 0;
  for (int k = blockIdx.x * RED_THREADS + threadIdx.x; k < N; k += gridDim.x * RED_THREADS) {
    const int i = k / cols, j0 = k - i * cols;
    DataTerm corres;
    corres.zero_x = corres.zero_y = corres.one_x = corres.one_y = 0;
    corres.diff = 0.f;
    corres.valid = 0;
    if (j0 < cols - 5 && i < rows - 1) {
      bool valid = true;
      for (int u = max(i - 2, 0); u < min(i + 2, rows); u++)
        for (int v = max(j0 - 2, 0); v < min(j0 + 2, cols); v++) valid = valid && (nextImage[(size_t)u * cols + v] > 0);
      if (valid) {
        const int valx = dIdx[k], valy = dIdy[k];
        const float mTwo = (float)((valx * valx) + (valy * valy));
        if (mTwo >= minScale) {
          const int y = i, x = j0;
          const float d1 = nextDepth[k];
          if (!isnan(d1)) {
            const float transformed_d1 = (float)(d1 * (krkinv.r[2].x * x + krkinv.r[2].y * y + krkinv.r[2].z) + kt.z);
            const int u0 = __float2int_rn((d1 * (krkinv.r[0].x * x + krkinv.r[0].y * y + krkinv.r[0].z) + kt.x) / transformed_d1);
            const int v0 = __float2int_rn((d1 * (krkinv.r[1].x * x + krkinv.r[1].y * y + krkinv.r[1].z) + kt.y) / transformed_d1);
            if (u0 >= 0 && v0 >= 0 && u0 < cols && v0 < rows) {
              const float d0 = lastDepth[(size_t)v0 * cols + u0];
              const uint8_t li = lastImage[(size_t)v0 * cols + u0];
              if (d0 > 0 && fabsf(transformed_d1 - d0) <= od.maxDepthDeltaRGB && li != 0) {
                corres.zero_x = (short)u0;
                corres.zero_y = (short)v0;
                corres.one_x = (short)x;
                corres.one_y = (short)y;
                corres.diff = (float)nextImage[k] - (float)li;
                corres.valid = 1;
                cnt += 1;
                sig += (unsigned int)__float2int_rz(corres.diff * corres.diff);
              }
            }
          }
        }
      }
    }
    corresImg[k] = corres;
  }
  // block reduce two ints (wrapping adds, like the reference's int2 sums)
  {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      cnt += __shfl_down_sync(0xffffffffu, cnt, off);
      sig += __shfl_down_sync(0xffffffffu, sig, off);
    }
    if (lane == 0) {
      sred_i[wid * 2] = (int)cnt;
      sred_i[wid * 2 + 1] = (int)sig;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned int c = 0, s = 0;
      for (int w = 0; w < RED_THREADS / 32; ++w) {
        c += (unsigned int)sred_i[w * 2];
        s += (unsigned int)sred_i[w * 2 + 1];
      }
      od.partials_i[blockIdx.x * 2] = (int)c;
      od.partials_i[blockIdx.x * 2 + 1] = (int)s;
    }
  }
  if (!last_block_done(od.counter)) return;
  if (threadIdx.x == 0) {
    unsigned int c = 0, s = 0;
    for (int b = 0; b < (int)gridDim.x; ++b) {
      c += (unsigned int)od.partials_i[b * 2];
      s += (unsigned int)od.partials_i[b * 2 + 1];
    }
    *od.counter = 0;
    const int rgbSize = (int)c, sigma = (int)s;
    gn->sum_res[0] = rgbSize;
    gn->sum_res[1] = sigma;
    if (finalize) {
      gn->rgbSize = rgbSize;
      gn->sigma = sigma;
      // reference: std::sqrt((float)sigma / rgbSize == 0 ? 1 : rgbSize)  (RGBDOdometry.cpp:442, App. A-1)
      float sigmaVal = (float)sqrt((double)(((float)sigma / rgbSize == 0) ? 1 : rgbSize));
      const float rgbError = (float)(sqrt((double)sigma) / (rgbSize == 0 ? 1 : rgbSize));
      const float prevError = (iter == 0) ? FLT_MAX : gn->lastRGBError;  // RGBDOdometry.cpp:404
      if (gn->rgbOnly && rgbError > prevError) {
        gn->break_level = level;
      } else {
        gn->lastRGBError = rgbError;
        gn->lastRGBCount = (float)rgbSize;
        if (gn->rgbOnly) sigmaVal = -1;
        gn->sigmaVal = sigmaVal;
      }
    }
  }
}

// SO3 pre-alignment step: SO3Reduction / so3Step + the host loop body (reduce.cu:789-973, RGBDOdometry.cpp:305-368)
__device__ __forceinline__ void so3_gradient(const uint8_t* img, int cols, int x, int y, float& gx, float& gy) {
  const float actu = (float)img[(size_t)y * cols + x];
  float back = (float)img[(size_t)y * cols + x - 1];
  float fore = (float)img[(size_t)y * cols + x + 1];
  gx = ((back + actu) / 2.0f) - ((fore + actu) / 2.0f);
  back = (float)img[(size_t)(y - 1) * cols + x];
  fore = (float)img[(size_t)(y + 1) * cols + x];
  gy = ((back + actu) / 2.0f) - ((fore + actu) / 2.0f);
}

__global__ void __launch_bounds__(RED_THREADS) k_so3_step(OdomDev od, int iter, int solve) {
  GNState* gn = od.gn;
  if (solve && gn->so3_done) return;
  __shared__ float sred[11 * (RED_THREADS / 32)];
  __shared__ double dsm[(RED_THREADS / 32) * 32];
  const int level = 2;
  const int rows = od.rows[level], cols = od.cols[level];
  const int N = rows * cols;
  const m33 imageBasis = load_m33(gn->imageBasis), kinv = load_m33(gn->kinv), krlr = load_m33(gn->krlr);
  const uint8_t* lastImage = od.lastNextImage[level];
  const uint8_t* nextImage = od.nextImage[level];
  float acc[11];
#pragma unroll
  for (int k = 0; k < 11; ++k) acc[k] = 0.f;
  for (int k = blockIdx.x * RED_THREADS + threadIdx.x; k < N; k += gridDim.x * RED_THREADS) {
    const int y = k / cols, x = k - y * cols;
    const f3 unwarped = mk3((float)x, (float)y, 1.0f);
    const f3 warped = mul(imageBasis, unwarped);
    const int wx = __float2int_rn(warped.x / warped.z);
    const int wy = __float2int_rn(warped.y / warped.z);
    if (wx >= 1 && wx < cols - 1 && wy >= 1 && wy < rows - 1 && x >= 1 && x < cols - 1 && y >= 1 && y < rows - 1) {
      float gnx, gny, glx, gly;
      so3_gradient(nextImage, cols, wx, wy, gnx, gny);
      so3_gradient(lastImage, cols, x, y, glx, gly);
      const float gx = (gnx + glx) / 2.0f;
      const float gy = (gny + gly) / 2.0f;
      const f3 point = mul(kinv, unwarped);
      const float z2 = point.z * point.z;
      const float a = krlr.r[0].x, b = krlr.r[0].y, c = krlr.r[0].z;
      const float d = krlr.r[1].x, e = krlr.r[1].y, f = krlr.r[1].z;
      const float g = krlr.r[2].x, h = krlr.r[2].y, i = krlr.r[2].z;
      const f3 leftProduct = mk3(((point.z * (d * gy + a * gx)) - (gy * g * y) - (gx * g * x)) / z2,
                                 ((point.z * (e * gy + b * gx)) - (gy * h * y) - (gx * h * x)) / z2,
                                 ((point.z * (f * gy + c * gx)) - (gy * i * y) - (gx * i * x)) / z2);
      const f3 jacRow = cross(leftProduct, point);
      float row[4];
      row[0] = jacRow.x;
      row[1] = jacRow.y;
      row[2] = jacRow.z;
      row[3] = -((float)nextImage[(size_t)wy * cols + wx] - (float)lastImage[k]);
      int q = 0;
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int s = r; s < 4; ++s) acc[q++] += row[r] * row[s];
      acc[9] += row[3] * row[3];
      acc[10] += 1.0f;
    }
  }
  block_reduce_sum<11, RED_THREADS>(acc, sred);
  if (threadIdx.x == 0) {
    float* my_partial = od.partials + (size_t)blockIdx.x * PARTIAL_STRIDE;
#pragma unroll
    for (int k = 0; k < 11; ++k) my_partial[k] = acc[k];
  }
  if (!last_block_done(od.counter)) return;
  final_sum(od.partials, gridDim.x, 0, 11, gn->sum_so3, dsm);
  if (threadIdx.x != 0) return;
  *od.counter = 0;
  if (!solve) return;

  float jtj[9], jtr[3];
  {
    int shift = 0;
    for (int i = 0; i < 3; ++i)
      for (int j = i; j < 4; ++j) {
        const float value = gn->sum_so3[shift++];
        if (j == 3)
          jtr[i] = value;
        else
          jtj[j * 3 + i] = jtj[i * 3 + j] = value;
      }
  }
  const float res0 = gn->sum_so3[9], res1 = gn->sum_so3[10];
  if (od.trace && gn->trace_n < MAX_TRACE) {
    EfSolveTrace& t = od.trace[gn->trace_n++];
    t.kind = 1;
    t.level = 2;
    t.iter = iter;
    for (int k = 0; k < 9; ++k) t.A_so3[k] = jtj[k];
    for (int k = 0; k < 3; ++k) t.b_so3[k] = jtr[k];
    t.so3_residual[0] = res0;
    t.so3_residual[1] = res1;
  }
  gn->lastSO3Error = sqrtf(res0) / res1;
  gn->lastSO3Count = res1;
  if (gn->lastSO3Error < gn->so3_lastError && gn->so3_lastCount == gn->lastSO3Count) {
    gn->so3_done = 1;  // converged
    return;
  } else if ((double)gn->lastSO3Error > (double)gn->so3_lastError + 0.001) {  // diverging
    gn->lastSO3Error = gn->so3_lastError;
    gn->lastSO3Count = gn->so3_lastCount;
    for (int k = 0; k < 9; ++k) gn->resultR[k] = gn->lastResultR[k];
    gn->so3_done = 1;
    return;
  }
  gn->so3_lastError = gn->lastSO3Error;
  gn->so3_lastCount = gn->lastSO3Count;
  for (int k = 0; k < 9; ++k) gn->lastResultR[k] = gn->resultR[k];
  float delta[3];
  efm::solve_sym3f(jtj, jtr, delta);
  double dd[3] = {delta[0], delta[1], delta[2]}, rotUpdate[9];
  efm::rodrigues(dd, rotUpdate);
  float ru[9], nr[9];
  for (int k = 0; k < 9; ++k) ru[k] = (float)rotUpdate[k];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c)
      nr[r * 3 + c] = ru[r * 3 + 0] * gn->R_lr[0 * 3 + c] + ru[r * 3 + 1] * gn->R_lr[1 * 3 + c] + ru[r * 3 + 2] * gn->R_lr[2 * 3 + c];
  for (int k = 0; k < 9; ++k) {
    gn->R_lr[k] = nr[k];
    gn->resultR[k] = nr[k];
  }
  so3_prepare(gn);
}

// =============================================================================================
// host side: RGBDOdometry mirror
// =============================================================================================

namespace {

inline int red_blocks(const EfContext* ctx, int n_items, int per_thread) {
  int b = (n_items + RED_THREADS * per_thread - 1) / (RED_THREADS * per_thread);
  int cap = ctx->num_sms * 4;
  if (cap > MAX_RED_BLOCKS) cap = MAX_RED_BLOCKS;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return b;
}
inline dim3 grid2d(int cols, int rows) { return dim3((cols + 31) / 32, (rows + 7) / 8); }
inline int flat_blocks(const EfContext* ctx, size_t n) {
  size_t b = (n + 255) / 256;
  size_t cap = (size_t)ctx->num_sms * 8;
  return (int)(b < cap ? (b ? b : 1) : cap);
}

#define EF_CHECK_LAST()                          \
  do {                                           \
    cudaError_t e__ = cudaGetLastError();        \
    if (e__ != cudaSuccess) return (int)e__;     \
  } while (0)

}  // namespace

namespace ef {

int odom_init_icp_depth(EfContext* ctx, int which, const uint16_t* depth_dev, float cutoff) {
  OdomDev& od = ctx->odom[which];
  cudaError_t e = cudaMemcpyAsync(od.depth_tmp[0], depth_dev, sizeof(uint16_t) * od.width * od.height, cudaMemcpyDeviceToDevice, ctx->stream);
  if (e != cudaSuccess) return (int)e;
  const dim3 block(32, 8);
  for (int i = 1; i < NUM_PYRS; ++i)
    EF_LAUNCH(ctx, k_pyr_down_u16, grid2d(od.cols[i], od.rows[i]), block, 0, od.depth_tmp[i - 1], od.rows[i - 1], od.cols[i - 1], od.depth_tmp[i]);
  for (int i = 0; i < NUM_PYRS; ++i) {
    const int div = 1 << i;
    const float fx = ctx->cfg.fx / div, fy = ctx->cfg.fy / div, cx = ctx->cfg.cx / div, cy = ctx->cfg.cy / div;
    EF_LAUNCH(ctx, k_vmap_nmap, grid2d(od.cols[i], od.rows[i]), block, 0, od.depth_tmp[i], od.rows[i], od.cols[i], 1.f / fx, 1.f / fy, cx, cy,
              cutoff, od.vmap_curr[i], od.nmap_curr[i]);
  }
  EF_CHECK_LAST();
  return 0;
}

static int copy_and_resize(EfContext* ctx, OdomDev& od, const float* vtx4, const float* nrm4, float** vm, float** nm,
                           const float* vtxB = nullptr, const float* nrmB = nullptr, const int* flag = nullptr) {
  const size_t n = (size_t)od.width * od.height;
  EF_LAUNCH(ctx, k_copy_maps, flat_blocks(ctx, n), 256, 0, (const float4*)vtx4, (const float4*)nrm4, (const float4*)vtxB, (const float4*)nrmB,
            flag, od.height, od.width, (float4*)od.vmaps_tmp, vm[0], nm[0]);
  const dim3 block(32, 8);
  for (int i = 1; i < NUM_PYRS; ++i)
    EF_LAUNCH(ctx, k_resize_maps, grid2d(od.cols[i], od.rows[i]), block, 0, vm[i - 1], nm[i - 1], od.rows[i - 1], od.cols[i - 1], vm[i], nm[i]);
  EF_CHECK_LAST();
  return 0;
}

int odom_init_icp_pred(EfContext* ctx, int which, const float* vtx4, const float* nrm4) {
  OdomDev& od = ctx->odom[which];
  return copy_and_resize(ctx, od, vtx4, nrm4, od.vmap_curr, od.nmap_curr);
}

// pose is taken from gn->T_wc (device) — callers that pass an explicit pose upload it first
int odom_init_icp_model(EfContext* ctx, int which, const float* vtx4, const float* nrm4, const float* vtxB = nullptr,
                        const float* nrmB = nullptr, const int* flag = nullptr) {
  OdomDev& od = ctx->odom[which];
  int rc = copy_and_resize(ctx, od, vtx4, nrm4, od.vmap_g_prev, od.nmap_g_prev, vtxB, nrmB, flag);
  if (rc) return rc;
  XformArgs a;
  for (int i = 0; i < NUM_PYRS; ++i) {
    a.v[i] = od.vmap_g_prev[i];
    a.n[i] = od.nmap_g_prev[i];
    a.rows[i] = od.rows[i];
    a.cols[i] = od.cols[i];
  }
  EF_LAUNCH(ctx, k_transform_maps, dim3(flat_blocks(ctx, (size_t)od.width * od.height), NUM_PYRS), 256, 0, a, (const GNState*)od.gn);
  EF_CHECK_LAST();
  return 0;
}

// populateRGBDData (RGBDOdometry.cpp:212-234); with_depth=0 is initFirstRGB (:246-257)
int odom_populate(EfContext* ctx, int which, const uint8_t* rgba, float** destDepths, uint8_t** destImages, bool with_depth,
                  const uint8_t* rgbaB = nullptr, const int* flag = nullptr, bool forceB = false) {
  OdomDev& od = ctx->odom[which];
  const size_t n = (size_t)od.width * od.height;
  EF_LAUNCH(ctx, k_depth_intensity_l0, flat_blocks(ctx, n), 256, 0, (const float4*)od.vmaps_tmp, (const uchar4*)rgba, (const uchar4*)rgbaB, flag,
            forceB ? 1 : 0, n, od.maxDepthRGB, with_depth ? destDepths[0] : (float*)nullptr, destImages[0]);
  const dim3 block(32, 8);
  for (int i = 0; i + 1 < NUM_PYRS; ++i)
    EF_LAUNCH(ctx, k_pyr_down_depth_image, grid2d(od.cols[i + 1], od.rows[i + 1]), block, 0, with_depth ? destDepths[i] : (const float*)nullptr,
              with_depth ? destDepths[i + 1] : (float*)nullptr, destImages[i], destImages[i + 1], od.rows[i], od.cols[i]);
  EF_CHECK_LAST();
  return 0;
}

// frameToModel.initICPModel + initRGBModel with the reference's fill-in choice (ElasticFusion.cpp:302-315) resolved on
// the device: predicted maps when the model view is dense enough, FillIn maps otherwise.
int map_select_model_inputs(EfContext* ctx, const float**, const float**, const uint8_t**) {
  OdomDev& od = ctx->odom[0];
  Textures& t = ctx->tex;
  int rc = odom_init_icp_model(ctx, 0, (const float*)t.vertex, (const float*)t.normal, (const float*)t.fill_vertex, (const float*)t.fill_normal,
                               ctx->map.dense_flag);
  if (rc) return rc;
  return odom_populate(ctx, 0, (const uint8_t*)t.image, od.lastDepth, od.lastImage, true, (const uint8_t*)t.fill_image, ctx->map.dense_flag,
                       ctx->frame_to_frame_rgb);
}

// the device-resident Gauss-Newton schedule; T_wc in/out lives in gn->T_wc
int odom_track_async(EfContext* ctx, int which, bool rgbOnly, float icpWeight, bool pyramid, bool fastOdom, bool so3) {
  OdomDev& od = ctx->odom[which];
  const bool icp = !rgbOnly && icpWeight > 0;
  const bool rgb = rgbOnly || icpWeight < 100;
  if (rgb) {
    SobelArgs a;
    for (int i = 0; i < NUM_PYRS; ++i) {
      a.src[i] = od.nextImage[i];
      a.dx[i] = od.dIdx[i];
      a.dy[i] = od.dIdy[i];
      a.rows[i] = od.rows[i];
      a.cols[i] = od.cols[i];
    }
    EF_LAUNCH(ctx, k_sobel, dim3(flat_blocks(ctx, (size_t)od.width * od.height), NUM_PYRS), 256, 0, a);
  }
  EF_LAUNCH(ctx, k_gn_begin, 1, 32, 0, od.gn, rgbOnly ? 1 : 0, icpWeight, so3 ? 1 : 0);
  if (so3) {
    const int nb = red_blocks(ctx, od.rows[2] * od.cols[2], 1);
    for (int i = 0; i < 10; ++i) EF_LAUNCH(ctx, k_so3_step, nb, RED_THREADS, 0, od, i, 1);
  }
  int iterations[NUM_PYRS] = {fastOdom ? 3 : 10, pyramid ? 5 : 0, pyramid ? 4 : 0};
  // static schedule of (level, iter)
  int sched_level[32], sched_iter[32], ns = 0;
  for (int i = NUM_PYRS - 1; i >= 0; --i)
    for (int j = 0; j < iterations[i]; ++j) {
      sched_level[ns] = i;
      sched_iter[ns] = j;
      ++ns;
    }
  EF_LAUNCH(ctx, k_gn_seed, 1, 32, 0, od.gn, ns ? sched_level[0] : 0);
  for (int s = 0; s < ns; ++s) {
    const int lv = sched_level[s];
    const int nb = red_blocks(ctx, od.rows[lv] * od.cols[lv], 2);
    if (rgb) EF_LAUNCH(ctx, k_rgb_residual, nb, RED_THREADS, 0, od, lv, sched_iter[s], 1);
    EF_LAUNCH(ctx, k_se3_step, nb, RED_THREADS, 0, od, lv, sched_iter[s], (s + 1 < ns) ? sched_level[s + 1] : -1, icp ? 1 : 0, rgb ? 1 : 0, 1);
  }
  if (so3)
    for (int i = 0; i < NUM_PYRS; ++i) {  // RGBDOdometry.cpp:560-564: handle swap
      uint8_t* t = od.lastNextImage[i];
      od.lastNextImage[i] = od.nextImage[i];
      od.nextImage[i] = t;
    }
  EF_CHECK_LAST();
  return 0;
}

int odom_finish_async(EfContext* ctx, int which, float weightMultiplier, bool have_track) {
  OdomDev& od = ctx->odom[which];
  EF_LAUNCH(ctx, k_gn_finish, 1, 32, 0, od.gn, weightMultiplier, have_track ? 1 : 0);
  EF_CHECK_LAST();
  return 0;
}

int odom_set_pose_async(EfContext* ctx, int which, const double* T_dev) {
  OdomDev& od = ctx->odom[which];
  EF_LAUNCH(ctx, k_set_pose, 1, 32, 0, od.gn, T_dev);
  EF_CHECK_LAST();
  return 0;
}

// stand-alone reduction launches for the stage API
int launch_se3_step_raw(EfContext* ctx, int which, int level, bool do_icp, bool do_rgb) {
  OdomDev& od = ctx->odom[which];
  const int nb = red_blocks(ctx, od.rows[level] * od.cols[level], 2);
  EF_LAUNCH(ctx, k_se3_step, nb, RED_THREADS, 0, od, level, 0, -1, do_icp ? 1 : 0, do_rgb ? 1 : 0, 0);
  EF_CHECK_LAST();
  return 0;
}
int launch_rgb_residual_raw(EfContext* ctx, int which, int level) {
  OdomDev& od = ctx->odom[which];
  const int nb = red_blocks(ctx, od.rows[level] * od.cols[level], 2);
  EF_LAUNCH(ctx, k_rgb_residual, nb, RED_THREADS, 0, od, level, 0, 0);
  EF_CHECK_LAST();
  return 0;
}
int launch_so3_raw(EfContext* ctx, int which) {
  OdomDev& od = ctx->odom[which];
  const int nb = red_blocks(ctx, od.rows[2] * od.cols[2], 1);
  EF_LAUNCH(ctx, k_so3_step, nb, RED_THREADS, 0, od, 0, 0);
  EF_CHECK_LAST();
  return 0;
}
int launch_sobel(EfContext* ctx, int which) {
  OdomDev& od = ctx->odom[which];
  SobelArgs a;
  for (int i = 0; i < NUM_PYRS; ++i) {
    a.src[i] = od.nextImage[i];
    a.dx[i] = od.dIdx[i];
    a.dy[i] = od.dIdy[i];
    a.rows[i] = od.rows[i];
    a.cols[i] = od.cols[i];
  }
  EF_LAUNCH(ctx, k_sobel, dim3(flat_blocks(ctx, (size_t)od.width * od.height), NUM_PYRS), 256, 0, a);
  EF_CHECK_LAST();
  return 0;
}

}  // namespace ef
