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

// reference pyrDownGaussKernel, cudafuncs.cu:75-121 (sigma_color 30, centre-relative gate, truncating store).
// `fetch(y, x)` abstracts the source so that level 2 can be produced in the same launch as level 1 by recomputing the
// level-1 values it needs from level 0 (identical arithmetic -> identical values, one launch instead of two).
template <typename Fetch>
__device__ __forceinline__ uint16_t pyr_down_u16_at(Fetch fetch, int srows, int scols, int x, int y) {
  const int D = 5;
  const float sigma_color = 30.f;
  const float weights[3] = {0.375f, 0.25f, 0.0625f};
  const int center = fetch(2 * y, 2 * x);
  const int x_mi = max(0, 2 * x - D / 2) - 2 * x;
  const int y_mi = max(0, 2 * y - D / 2) - 2 * y;
  const int x_ma = min(scols, 2 * x - D / 2 + D) - 2 * x;
  const int y_ma = min(srows, 2 * y - D / 2 + D) - 2 * y;
  float sum = 0, wall = 0;
  // fixed 5x5 trip count with predication (same accumulation order): all taps are requested at once
#pragma unroll
  for (int yi = -2; yi <= 2; ++yi)
#pragma unroll
    for (int xi = -2; xi <= 2; ++xi) {
      const bool in = (yi >= y_mi) && (yi < y_ma) && (xi >= x_mi) && (xi < x_ma);
      const int val = in ? fetch(2 * y + yi, 2 * x + xi) : 0;
      if (in && (float)abs(val - center) < 3 * sigma_color) {
        sum += val * weights[abs(xi)] * weights[abs(yi)];
        wall += weights[abs(xi)] * weights[abs(yi)];
      }
    }
  return (uint16_t)__float2int_rz(sum / wall);
}

__global__ void k_pyr_down_u16(const uint16_t* __restrict__ src, int srows, int scols, uint16_t* __restrict__ dst) {
  pdl_enter();
  const int r1 = srows / 2, c1 = scols / 2;
  auto f0 = [&](int yy, int xx) { return (int)src[(size_t)yy * scols + xx]; };
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < r1 * c1; t += gridDim.x * blockDim.x) {
    const int y = t / c1, x = t - y * c1;
    dst[t] = pyr_down_u16_at(f0, srows, scols, x, y);
  }
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

struct VmapArgs {
  const uint16_t* depth[NUM_PYRS];
  float* vmap[NUM_PYRS];
  float* nmap[NUM_PYRS];
  int rows[NUM_PYRS], cols[NUM_PYRS], start[NUM_PYRS + 1];
  float fx_inv[NUM_PYRS], fy_inv[NUM_PYRS], cx[NUM_PYRS], cy[NUM_PYRS];
};
__device__ __forceinline__ void vmap_nmap_at(const uint16_t* __restrict__ depth, int rows, int cols, float fx_inv, float fy_inv, float cx,
                                             float cy, float cutoff, float* __restrict__ vmap, float* __restrict__ nmap, int u, int v) {
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
__global__ void k_vmap_nmap(VmapArgs a, float cutoff) {
  pdl_enter();
  const int total = a.start[NUM_PYRS];
  for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < total; f += gridDim.x * blockDim.x) {
    const int lv = (f >= a.start[2]) ? 2 : (f >= a.start[1] ? 1 : 0);
    const int p = f - a.start[lv];
    const int v = p / a.cols[lv], u = p - v * a.cols[lv];
    vmap_nmap_at(a.depth[lv], a.rows[lv], a.cols[lv], a.fx_inv[lv], a.fy_inv[lv], a.cx[lv], a.cy[lv], cutoff, a.vmap[lv], a.nmap[lv], u, v);
  }
}

// copyMaps (cudafuncs.cu:295-381): predicted float4 vertex/normal maps -> vmaps_tmp (AoS copy) + SoA planes; z==0 -> NaN
// (vtxB, nrmB, dense_flag): optional fill-in alternates chosen on the device when *dense_flag == 0
// (ElasticFusion.cpp:304-313 decides this on the host after a glReadPixels)
__global__ void k_copy_maps(const float4* __restrict__ vtxA, const float4* __restrict__ nrmA, const float4* __restrict__ vtxB,
                            const float4* __restrict__ nrmB, const int* __restrict__ dense_flag, int rows, int cols,
                            float4* __restrict__ vmaps_tmp, float* __restrict__ vmap, float* __restrict__ nmap) {
  pdl_enter();
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
  pdl_enter();
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

// tranformMaps (cudafuncs.cu:221-293), all three levels in one launch (blockIdx.y = level), pose from gn->T_wc. The
// reference transforms in place; here the camera-frame maps are kept (the ICP kernel reads those) and the world-frame
// copy is only produced for the stage API / inspection.
struct XformArgs {
  const float* sv[NUM_PYRS];
  const float* sn[NUM_PYRS];
  float* v[NUM_PYRS];
  float* n[NUM_PYRS];
  int rows[NUM_PYRS], cols[NUM_PYRS];
};
__global__ void k_transform_maps(XformArgs a, const GNState* __restrict__ gn) {
  pdl_enter();
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
  const float* __restrict__ sv = a.sv[lv];
  const float* __restrict__ sn = a.sn[lv];
  float* __restrict__ vm = a.v[lv];
  float* __restrict__ nm = a.n[lv];
  for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < np; p += (size_t)gridDim.x * blockDim.x) {
    f3 v = mk3(sv[p], sv[p + np], sv[p + 2 * np]);
    if (!isnan(v.x)) v = mul(Rm, v) + tv;
    vm[p] = v.x;
    vm[p + np] = v.y;
    vm[p + 2 * np] = v.z;
    f3 n = mk3(sn[p], sn[p + np], sn[p + 2 * np]);
    if (!isnan(n.x)) n = mul(Rm, n);
    nm[p] = n.x;
    nm[p + np] = n.y;
    nm[p + 2 * np] = n.z;
  }
}

// verticesToDepth + imageBGRToIntensity fused (cudafuncs.cu:564-610): level-0 depth from vmaps_tmp.z (cutoff 6 m),
// level-0 intensity int(0.114 x + 0.299 y + 0.587 z) from the RGBA8 texel. Either output may be NULL.
__global__ void k_depth_intensity_l0(const float4* __restrict__ vmaps_tmp, const uchar4* __restrict__ rgbaA,
                                     const uchar4* __restrict__ rgbaB, const int* __restrict__ dense_flag, int forceB, size_t n,
                                     float cutoff, float* __restrict__ depth, uint8_t* __restrict__ image) {
  pdl_enter();
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

// pyrDownKernelGaussF + pyrDownKernelIntensityGauss (cudafuncs.cu:383-411,512-562): window [2x-2, min(2x+3, n-1)),
// flipped/shifted tap index, int normaliser, NaN / zero skipping, truncating u8 store. As above, level 2 is produced in the
// same launch by recomputing the level-1 taps it needs.
template <typename Fetch>
__device__ __forceinline__ float pyr_down_f_at(Fetch fetch, int srows, int scols, int x, int y) {
  const int D = 5;
  const int tx = min(2 * x - D / 2 + D, scols - 1);
  const int ty = min(2 * y - D / 2 + D, srows - 1);
  float sum = 0;
  int count = 0;
  const int cy0 = max(0, 2 * y - D / 2), cx0 = max(0, 2 * x - D / 2);
#pragma unroll
  for (int dy = 0; dy < 5; ++dy)
#pragma unroll
    for (int dx = 0; dx < 5; ++dx) {
      const int cy = cy0 + dy, cx = cx0 + dx;
      const bool in = (cy < ty) && (cx < tx);
      const float s = in ? fetch(cy, cx) : 0.f;
      if (in && !isnan(s)) {
        const float w = c_gauss25[(ty - cy - 1) * 5 + (tx - cx - 1)];
        sum += s * w;
        count = __float2int_rz((float)count + w);
      }
    }
  return (float)(sum / (float)count);
}
template <typename Fetch>
__device__ __forceinline__ uint8_t pyr_down_u8_at(Fetch fetch, int srows, int scols, int x, int y) {
  const int D = 5;
  const int tx = min(2 * x - D / 2 + D, scols - 1);
  const int ty = min(2 * y - D / 2 + D, srows - 1);
  float sum = 0;
  int count = 0;
  const int cy0 = max(0, 2 * y - D / 2), cx0 = max(0, 2 * x - D / 2);
#pragma unroll
  for (int dy = 0; dy < 5; ++dy)
#pragma unroll
    for (int dx = 0; dx < 5; ++dx) {
      const int cy = cy0 + dy, cx = cx0 + dx;
      const bool in = (cy < ty) && (cx < tx);
      const int s = in ? fetch(cy, cx) : 0;
      if (in && s > 0) {
        const float w = c_gauss25[(ty - cy - 1) * 5 + (tx - cx - 1)];
        sum += s * w;
        count = __float2int_rz((float)count + w);
      }
    }
  const int v = __float2int_rz(sum / (float)count);  // NaN -> 0
  return (uint8_t)min(max(v, 0), 255);
}

// depth (fp32) and intensity (u8) pyramid level in one launch. Either pair may be NULL.
__global__ void k_pyr_down_depth_image(const float* __restrict__ d0, float* __restrict__ d1, const uint8_t* __restrict__ i0,
                                       uint8_t* __restrict__ i1, int srows, int scols) {
  pdl_enter();
  const int r1 = srows / 2, c1 = scols / 2;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < r1 * c1; t += gridDim.x * blockDim.x) {
    const int y = t / c1, x = t - y * c1;
    if (d0) {
      auto f0 = [&](int yy, int xx) { return d0[(size_t)yy * scols + xx]; };
      d1[t] = pyr_down_f_at(f0, srows, scols, x, y);
    }
    if (i0) {
      auto g0 = [&](int yy, int xx) { return (int)i0[(size_t)yy * scols + xx]; };
      i1[t] = pyr_down_u8_at(g0, srows, scols, x, y);
    }
  }
}

// The model side of a frame (frameToModel.initICPModel + initRGBModel, ElasticFusion.cpp:302-322) in three launches instead of six:
// level 0 = k_copy_maps + k_depth_intensity_l0 in one pass over the predicted textures (the depth IS the vertex z just loaded);
// each coarser level = k_resize_maps + k_pyr_down_depth_image for the same output pixel. Same arithmetic, same outputs; the
// separate kernels remain for the stage API and the trackers that initialise only one half.
__global__ void k_model_level0(const float4* __restrict__ vtxA, const float4* __restrict__ nrmA, const float4* __restrict__ vtxB,
                               const float4* __restrict__ nrmB, const uchar4* __restrict__ rgbaA, const uchar4* __restrict__ rgbaB,
                               const int* __restrict__ dense_flag, int forceB_rgb, int rows, int cols, float cutoff, float4* __restrict__ vmaps_tmp,
                               float* __restrict__ vmap, float* __restrict__ nmap, float* __restrict__ depth, uint8_t* __restrict__ image) {
  pdl_enter();
  const size_t n = (size_t)rows * cols;
  const bool useB = dense_flag && (*dense_flag == 0);
  const float4* __restrict__ vtx = useB ? vtxB : vtxA;
  const float4* __restrict__ nrm = useB ? nrmB : nrmA;
  const uchar4* __restrict__ rgba = (forceB_rgb || useB) ? rgbaB : rgbaA;
  for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (size_t)gridDim.x * blockDim.x) {
    const float4 vs = vtx[p];
    const float4 ns = nrm[p];
    const uchar4 s = rgba[p];
    vmaps_tmp[p] = vs;
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
    depth[p] = (vs.z > cutoff || vs.z <= 0) ? qnan() : vs.z;
    image[p] = (uint8_t)__float2int_rz((float)s.x * 0.114f + (float)s.y * 0.299f + (float)s.z * 0.587f);
  }
}

__global__ void k_model_level_down(const float* __restrict__ vin, const float* __restrict__ nin, const float* __restrict__ d0,
                                   const uint8_t* __restrict__ i0, int srows, int scols, float* __restrict__ vout, float* __restrict__ nout,
                                   float* __restrict__ d1, uint8_t* __restrict__ i1) {
  pdl_enter();
  const int drows = srows / 2, dcols = scols / 2;
  const size_t splane = (size_t)srows * scols, dplane = (size_t)drows * dcols;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < drows * dcols; t += gridDim.x * blockDim.x) {
    const int y = t / dcols, x = t - y * dcols;
    const int xs = x * 2, ys = y * 2;
    const size_t q = (size_t)t;
    // resizeVMap + resizeNMap (cudafuncs.cu:413-490): 2x2 box average, any NaN -> NaN, normals renormalised
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
    // pyrDownGaussF + pyrDownUcharGauss for the same output pixel
    auto f0 = [&](int yy, int xx) { return d0[(size_t)yy * scols + xx]; };
    d1[t] = pyr_down_f_at(f0, srows, scols, x, y);
    auto g0 = [&](int yy, int xx) { return (int)i0[(size_t)yy * scols + xx]; };
    i1[t] = pyr_down_u8_at(g0, srows, scols, x, y);
  }
}

// computeDerivativeImages / applyKernel (cudafuncs.cu:612-668) for all three levels in one launch, fused with the
// pose-independent gates of computeRgbResidual (reduce.cu:641-660): a pixel is a photometric *candidate* when
// j < cols-5, i < rows-1, its 4x4 neighbourhood of the live image is non-zero, its gradient magnitude passes minScale and
// its depth is finite. None of this changes across the 19 Gauss-Newton iterations, so it is evaluated once per frame and
// the iterations only visit the compacted candidate list (typically ~10 % of the pixels at level 0).
struct SobelArgs {
  const uint8_t* src[NUM_PYRS];
  const float* depth[NUM_PYRS];
  int16_t* dx[NUM_PYRS];
  int16_t* dy[NUM_PYRS];
  int rows[NUM_PYRS], cols[NUM_PYRS];
  int start[NUM_PYRS + 1];
  float minScale[NUM_PYRS];
};
__global__ void k_sobel_cand(SobelArgs a, uint8_t* __restrict__ flags) {
  pdl_enter();
  const float gsx[9] = {(float)0.52201, (float)0.00000, (float)-0.52201, (float)0.79451, (float)-0.00000,
                        (float)-0.79451, (float)0.52201, (float)0.00000, (float)-0.52201};
  const float gsy[9] = {(float)0.52201, (float)0.79451, (float)0.52201, (float)0.00000, (float)0.00000,
                        (float)0.00000, (float)-0.52201, (float)-0.79451, (float)-0.52201};
  const int total = a.start[NUM_PYRS];
  for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < total; f += gridDim.x * blockDim.x) {
    const int lv = (f >= a.start[2]) ? 2 : (f >= a.start[1] ? 1 : 0);
    const int rows = a.rows[lv], cols = a.cols[lv];
    const int p = f - a.start[lv];
    const uint8_t* __restrict__ src = a.src[lv];
    const int y = p / cols, x = p - y * cols;
    float dxVal = 0, dyVal = 0;
    int kernelIndex = 8;
    for (int j = max(y - 1, 0); j <= min(y + 1, rows - 1); j++)
      for (int i = max(x - 1, 0); i <= min(x + 1, cols - 1); i++) {
        const float s = (float)src[(size_t)j * cols + i];
        dxVal += s * gsx[kernelIndex];
        dyVal += s * gsy[kernelIndex];
        --kernelIndex;
      }
    const int valx = (int16_t)__float2int_rz(dxVal), valy = (int16_t)__float2int_rz(dyVal);
    a.dx[lv][p] = (int16_t)valx;
    a.dy[lv][p] = (int16_t)valy;
    bool ok = (x < cols - 5 && y < rows - 1);
    if (ok) {
      const float mTwo = (float)((valx * valx) + (valy * valy));
      ok = (mTwo >= a.minScale[lv]) && !isnan(a.depth[lv][p]);
    }
    if (ok) {
      for (int u = max(y - 2, 0); u < min(y + 2, rows); u++)
        for (int v = max(x - 2, 0); v < min(x + 2, cols); v++) ok = ok && (src[(size_t)u * cols + v] > 0);
    }
    flags[f] = ok ? 1 : 0;
  }
}

__global__ void k_cand_scatter(SobelArgs a, const uint8_t* __restrict__ flags, const int* __restrict__ offsets, int4* __restrict__ cand,
                               GNState* gn) {
  pdl_enter();
  const int total = a.start[NUM_PYRS];
  for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < total; f += gridDim.x * blockDim.x) {
    const int lv = (f >= a.start[2]) ? 2 : (f >= a.start[1] ? 1 : 0);
    const int o = offsets[f];
    if (f == a.start[lv]) gn->cand_base[lv] = o;  // exclusive prefix at the first pixel of the level
    if (!flags[f]) continue;
    const int p = f - a.start[lv];
    const int g = ((int)(uint16_t)a.dx[lv][p]) | (((int)(uint16_t)a.dy[lv][p]) << 16);
    cand[o] = make_int4(p, __float_as_int(a.depth[lv][p]), g, (int)a.src[lv][p]);
  }
}

// =============================================================================================
// host side: RGBDOdometry mirror
// =============================================================================================

namespace {

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
int run_scan(EfContext* ctx, const uint8_t* flags, const int* n_a, const int* n_b, size_t max_items, int* offsets, int* total);
void scan_scratch(EfContext* ctx, uint8_t** flags, int** offsets);

int odom_init_icp_depth(EfContext* ctx, int which, const uint16_t* depth_dev, float cutoff) {
  OdomDev& od = ctx->odom[which];
  ctx->maps_dirty[which] = true;
  cudaError_t e = cudaMemcpyAsync(od.depth_tmp[0], depth_dev, sizeof(uint16_t) * od.width * od.height, cudaMemcpyDeviceToDevice, ctx->stream);
  if (e != cudaSuccess) return (int)e;
  for (int i = 1; i < NUM_PYRS; ++i)
    EF_LAUNCH(ctx, k_pyr_down_u16, flat_blocks(ctx, (size_t)od.rows[i] * od.cols[i] * 2), 128, 0, od.depth_tmp[i - 1], od.rows[i - 1], od.cols[i - 1],
              od.depth_tmp[i]);
  VmapArgs va;
  for (int i = 0; i < NUM_PYRS; ++i) {
    const int div = 1 << i;
    const float fx = ctx->cfg.fx / div, fy = ctx->cfg.fy / div, cx = ctx->cfg.cx / div, cy = ctx->cfg.cy / div;
    va.depth[i] = od.depth_tmp[i];
    va.vmap[i] = od.vmap_curr[i];
    va.nmap[i] = od.nmap_curr[i];
    va.rows[i] = od.rows[i];
    va.cols[i] = od.cols[i];
    va.start[i] = od.level_start[i];
    va.fx_inv[i] = 1.f / fx;
    va.fy_inv[i] = 1.f / fy;
    va.cx[i] = cx;
    va.cy[i] = cy;
  }
  va.start[NUM_PYRS] = od.level_start[NUM_PYRS];
  EF_LAUNCH(ctx, k_vmap_nmap, flat_blocks(ctx, (size_t)od.level_start[NUM_PYRS]), 256, 0, va, cutoff);
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
  ctx->maps_dirty[which] = true;
  return copy_and_resize(ctx, od, vtx4, nrm4, od.vmap_curr, od.nmap_curr);
}

// pose is taken from gn->T_wc (device) — callers that pass an explicit pose upload it first
int odom_init_icp_model(EfContext* ctx, int which, const float* vtx4, const float* nrm4, const float* vtxB = nullptr,
                        const float* nrmB = nullptr, const int* flag = nullptr, bool with_global = true) {
  OdomDev& od = ctx->odom[which];
  int rc = copy_and_resize(ctx, od, vtx4, nrm4, od.vmap_c_prev, od.nmap_c_prev, vtxB, nrmB, flag);
  if (rc) return rc;
  if (!with_global) return 0;  // the frame loop never reads the world-frame copy
  XformArgs a;
  for (int i = 0; i < NUM_PYRS; ++i) {
    a.sv[i] = od.vmap_c_prev[i];
    a.sn[i] = od.nmap_c_prev[i];
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
                  const uint8_t* rgbaB = nullptr, const int* flag = nullptr, bool forceB = false, bool with_image = true) {
  OdomDev& od = ctx->odom[which];
  const size_t n = (size_t)od.width * od.height;
  EF_LAUNCH(ctx, k_depth_intensity_l0, flat_blocks(ctx, n), 256, 0, (const float4*)od.vmaps_tmp, (const uchar4*)rgba, (const uchar4*)rgbaB, flag,
            forceB ? 1 : 0, n, od.maxDepthRGB, with_depth ? destDepths[0] : (float*)nullptr, with_image ? destImages[0] : (uint8_t*)nullptr);
  for (int i = 0; i + 1 < NUM_PYRS; ++i)
    EF_LAUNCH(ctx, k_pyr_down_depth_image, flat_blocks(ctx, (size_t)od.rows[i + 1] * od.cols[i + 1] * 2), 128, 0,
              with_depth ? destDepths[i] : (const float*)nullptr, with_depth ? destDepths[i + 1] : (float*)nullptr,
              with_image ? (const uint8_t*)destImages[i] : (const uint8_t*)nullptr, with_image ? destImages[i + 1] : (uint8_t*)nullptr, od.rows[i],
              od.cols[i]);
  EF_CHECK_LAST();
  return 0;
}

// frameToModel.initICPModel + initRGBModel with the reference's fill-in choice (ElasticFusion.cpp:302-315) resolved on
// the device: predicted maps when the model view is dense enough, FillIn maps otherwise.
int map_select_model_inputs(EfContext* ctx, const float**, const float**, const uint8_t**) {
  OdomDev& od = ctx->odom[0];
  Textures& t = ctx->tex;
  if (ctx->fused_model_side) {
    const size_t n = (size_t)od.width * od.height;
    EF_LAUNCH(ctx, k_model_level0, flat_blocks(ctx, n), 256, 0, (const float4*)t.vertex, (const float4*)t.normal, (const float4*)t.fill_vertex,
              (const float4*)t.fill_normal, (const uchar4*)t.image, (const uchar4*)t.fill_image, (const int*)ctx->map.dense_flag,
              ctx->frame_to_frame_rgb ? 1 : 0, od.height, od.width, od.maxDepthRGB, (float4*)od.vmaps_tmp, od.vmap_c_prev[0], od.nmap_c_prev[0],
              od.lastDepth[0], od.lastImage[0]);
    for (int i = 0; i + 1 < NUM_PYRS; ++i)
      EF_LAUNCH(ctx, k_model_level_down, flat_blocks(ctx, (size_t)od.rows[i + 1] * od.cols[i + 1] * 2), 128, 0, (const float*)od.vmap_c_prev[i],
                (const float*)od.nmap_c_prev[i], (const float*)od.lastDepth[i], (const uint8_t*)od.lastImage[i], od.rows[i], od.cols[i],
                od.vmap_c_prev[i + 1], od.nmap_c_prev[i + 1], od.lastDepth[i + 1], od.lastImage[i + 1]);
    EF_CHECK_LAST();
    return 0;
  }
  int rc = odom_init_icp_model(ctx, 0, (const float*)t.vertex, (const float*)t.normal, (const float*)t.fill_vertex, (const float*)t.fill_normal,
                               ctx->map.dense_flag, false);
  if (rc) return rc;
  return odom_populate(ctx, 0, (const uint8_t*)t.image, od.lastDepth, od.lastImage, true, (const uint8_t*)t.fill_image, ctx->map.dense_flag,
                       ctx->frame_to_frame_rgb);
}

int launch_sobel(EfContext* ctx, int which) {
  OdomDev& od = ctx->odom[which];
  SobelArgs a;
  for (int i = 0; i < NUM_PYRS; ++i) {
    a.src[i] = od.nextImage[i];
    a.depth[i] = od.nextDepth[i];
    a.dx[i] = od.dIdx[i];
    a.dy[i] = od.dIdy[i];
    a.rows[i] = od.rows[i];
    a.cols[i] = od.cols[i];
    a.start[i] = od.level_start[i];
    a.minScale[i] = od.minScale[i];
  }
  a.start[NUM_PYRS] = od.level_start[NUM_PYRS];
  const size_t flat = (size_t)od.level_start[NUM_PYRS];
  uint8_t* flags;
  int* offsets;
  scan_scratch(ctx, &flags, &offsets);
  EF_LAUNCH(ctx, k_sobel_cand, flat_blocks(ctx, flat), 256, 0, a, flags);
  int rc = run_scan(ctx, flags, &od.gn->flat_n, nullptr, flat, offsets, &od.gn->cand_base[NUM_PYRS]);
  if (rc) return rc;
  EF_LAUNCH(ctx, k_cand_scatter, flat_blocks(ctx, flat), 256, 0, a, (const uint8_t*)flags, (const int*)offsets, od.cand, od.gn);
  ctx->maps_dirty[which] = true;  // k_iter1 / k_iter2 read the list ahead of their dependency wait: fence before the next one
  EF_CHECK_LAST();
  return 0;
}

}  // namespace ef
