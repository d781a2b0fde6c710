// Reductions of the tracking half: the geometric (ICP) and photometric 6x6 systems, the photometric correspondence
// pass and the SO(3) pre-alignment, each as ONE launch (per-CTA shuffle tree -> partials -> last-CTA final sum in double),
// plus the device-resident Gauss-Newton state machine that replaces the reference's host loop
// (Core/Utils/RGBDOdometry.cpp:259-571; kernels Core/Cuda/reduce.cu).
//
// This translation unit is compiled WITH fused multiply-add contraction (the reference build has it on as well): the
// sums are order-dependent anyway, and FMA removes about a third of the instructions of the per-pixel rows. The
// per-pixel image/pyramid kernels live in ef_track.cu, compiled with --fmad=false for bit-reproducibility.
#include <float.h>
#include <stddef.h>
#include <stdio.h>

#include "ef_device.cuh"
#include <stdlib.h>
#include <string.h>

#include "ef_dmath.cuh"
#include "ef_internal.h"

using namespace ef;

#ifdef EF_PROFILE_PHASES
__device__ __forceinline__ long long ef_gtime() {
  long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define EF_STAMP(gn, slot, cond) do { if (cond) (gn)->dbg[slot] = ef_gtime(); } while (0)
#else
#define EF_STAMP(gn, slot, cond) do { } while (0)
#endif

namespace ef {
int launch_sobel(EfContext* ctx, int which);
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

// K and K^-1 of a pyramid level in double; the pinhole inverse is closed-form (no general 3x3 inverse on the device)
__device__ __forceinline__ void level_K(const GNState* gn, int level, double* K, double* Kinv) {
  // filled once at context creation (ef_api.cu): K and its closed-form inverse for each pyramid level
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    K[k] = gn->Kd[level][k];
    Kinv[k] = gn->Kinvd[level][k];
  }
}

// warp matrices for the next photometric residual pass (RGBDOdometry.cpp:407-417). resultRt is a rigid transform
// (products of Rodrigues rotations and translations), so its inverse is [R^T | -R^T t]: no 4x4 elimination.
__device__ void gn_prepare_warp(GNState* gn, int level) {
  double K[9], Kinv[9], Rt[16];
  level_K(gn, level, K, Kinv);
  efm::se3_inverse(gn->resultRt, Rt);
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
__device__ void so3_prepare(So3State* s, const GNState* gn) {
  double K[9], Kinv[9], tmp[9], H[9];
  level_K(gn, 2, K, Kinv);  // (constant after context creation)
  efm::mul3(K, s->resultR, tmp);
  efm::mul3(tmp, Kinv, H);
  for (int k = 0; k < 9; ++k) {
    s->imageBasis[k] = (float)H[k];
    s->kinv[k] = (float)Kinv[k];
    s->krlr[k] = (float)tmp[k];
  }
}
// start of the SO(3) loop (RGBDOdometry.cpp:284-303)
__global__ void k_so3_begin(So3State* s, const GNState* gn) {
  pdl_enter();
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const double I3[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  for (int k = 0; k < 9; ++k) {
    s->resultR[k] = I3[k];
    s->lastResultR[k] = I3[k];
    s->R_lr[k] = (float)I3[k];
  }
  s->so3_lastError = FLT_MAX / 2;
  s->so3_lastCount = FLT_MAX / 2;
  s->so3_done = 0;
  s->trace_n = 0;
  so3_prepare(s, gn);
}

// start of getIncrementalTransformation (RGBDOdometry.cpp:266-273,284-303)
__device__ void gn_begin_body(GNState* gn, int rgbOnly, float icpWeight, int so3) {
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
  for (int k = 0; k < 9; ++k) gn->Mcp[k] = (k % 4 == 0) ? 1.f : 0.f;  // Rcurr = Rprev, tcurr = tprev
  for (int k = 0; k < 3; ++k) gn->tcp[k] = 0.f;
  efm::inv3<float>(gn->Rprev, gn->Rprev_inv);
  gn->break_level = -1;
  gn->trace_n = 0;
}
// One CTA: thread 0 starts getIncrementalTransformation; all threads copy the records of the SO(3) loop (which ran before,
// possibly on the look-ahead stream) into the tracker's trace (a few hundred words: one or two per thread, so the copy is
// one memory round trip instead of a serial chain).
constexpr int GN_BEGIN_THREADS = 256;
__device__ void gn_seed_body(GNState* gn, int first_level, const So3State* s);
// first_level >= 0: thread 0 also seeds the SE(3) loop (resultRt, first warp matrices): one launch instead of two.
__global__ void __launch_bounds__(GN_BEGIN_THREADS) k_gn_begin(GNState* gn, int rgbOnly, float icpWeight, int so3, const So3State* s,
                                                               EfSolveTrace* trace, int first_level) {
  pdl_enter();
  if (blockIdx.x != 0) return;
  if (threadIdx.x == 0) {
    gn_begin_body(gn, rgbOnly, icpWeight, so3);
    if (so3) {
      gn->lastSO3Error = s->lastSO3Error;
      gn->lastSO3Count = s->lastSO3Count;
      gn->trace_n = trace ? s->trace_n : 0;
    }
    if (first_level >= 0) gn_seed_body(gn, first_level, s);
  }
  if (so3 && trace) {
    const int words = s->trace_n * (int)(sizeof(EfSolveTrace) / 4);
    const int* src = reinterpret_cast<const int*>(s->trace);
    int* dst = reinterpret_cast<int*>(trace);
    for (int k = threadIdx.x; k < words; k += GN_BEGIN_THREADS) dst[k] = src[k];
  }
}

// after the SO3 loop: seed resultRt (RGBDOdometry.cpp:379-388) and prepare the first SE3 iteration
__device__ void gn_seed_body(GNState* gn, int first_level, const So3State* s) {
  for (int k = 0; k < 16; ++k) gn->resultRt[k] = (k % 5 == 0) ? 1.0 : 0.0;
  if (gn->so3)
    for (int x = 0; x < 3; x++)
      for (int y = 0; y < 3; y++) gn->resultRt[x * 4 + y] = s->resultR[x * 3 + y];
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
__global__ void k_gn_finish(GNState* gn, float weightMultiplier, int have_track, MapPose* map_pose) {
  pdl_enter();
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
  if (map_pose) {
    // the map kernels' pose "uniforms" (GlobalModel.cpp:405,562): float casts of T_wc and of its rigid inverse
    for (int k = 0; k < 16; ++k) {
      map_pose->pose[k] = (float)gn->T_wc[k];
      map_pose->t_inv[k] = (float)inv[k];
    }
  }
}

// k_gn_finish needs the pre-tracking pose; stash it (tracking overwrites T_wc only at the end, so this is only needed
// for the in_T_wc path where the host replaces the pose).
__global__ void k_set_pose(GNState* gn, const double* T_new) {
  pdl_enter();
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

// Shared-memory scratch of the update kernel
struct GnScratch {
  double dsm[20][64];  // per-slice partial sums (threads/32 slices x 64 values)
  float sums[64];      // reduced systems: [0,29) geometric, [32,61) photometric
  float A_icp[36], b_icp[8], A_rgb[36], b_rgb[8];
  double lastA[36], lastb[8], result[8];
  double rRt[16], nrt[16], upd[16];
  double K[9], Kinv[9], Rt[16], tmp[9];
  float Rprev[9], tprev[3], iR[9], it[3];
  int flags[2];
  double w;
};

// packed index k (reference JtJJtrSE3 order, types.cuh:98-104) -> (i, j) with j == 6 meaning the b column
__device__ __forceinline__ void se3_unpack_index(int k, int& i, int& j) {
  int row = 0, start = 0;
#pragma unroll
  for (int r = 0; r < 6; ++r) {
    const int len = 7 - r;
    if (row == r && k >= start + len) {
      start += len;
      row = r + 1;
    }
  }
  i = row;
  j = row + (k - start);
}

// One SE3 Gauss-Newton update (RGBDOdometry.cpp:492-551) executed cooperatively by one warp: the independent pieces
// (unpacking, lastA/lastb, 4x4 and 3x3 products, trace) are spread over the lanes; only the 6x6 LDL^T and the Rodrigues
// formula run on lane 0. All operands live in shared memory.
__device__ __noinline__ void gn_update_warp(const OdomDev& od, GnScratch& S, int level, int iter, int next_level) {
  GNState* gn = od.gn;
  const int lane = threadIdx.x;
  // (resultRt, Rprev, tprev were staged into S at the top of k_iter2, long before this CTA took the last ticket)
  const int icp = S.flags[0], rgb = S.flags[1];
  const double w = S.w;
  if (lane < 27) {
    int i, j;
    se3_unpack_index(lane, i, j);
    const float vi = S.sums[lane], vr = S.sums[32 + lane];
    if (j == 6) {
      S.b_icp[i] = vi;
      S.b_rgb[i] = vr;
    } else {
      S.A_icp[i * 6 + j] = S.A_icp[j * 6 + i] = vi;
      S.A_rgb[i * 6 + j] = S.A_rgb[j * 6 + i] = vr;
    }
  }
  __syncwarp();
  for (int k = lane; k < 36; k += 32) {
    double a;
    if (icp && rgb)
      a = (double)S.A_rgb[k] + w * w * (double)S.A_icp[k];
    else if (icp)
      a = S.A_icp[k];
    else
      a = S.A_rgb[k];
    S.lastA[k] = a;
    gn->lastA[k] = a;
  }
  if (lane < 6) {
    double b;
    if (icp && rgb)
      b = (double)S.b_rgb[lane] + w * (double)S.b_icp[lane];
    else if (icp)
      b = S.b_icp[lane];
    else
      b = S.b_rgb[lane];
    S.lastb[lane] = b;
    gn->lastb[lane] = b;
  }
  __syncwarp();
  const float res0 = S.sums[27], res1 = S.sums[28];
  EF_STAMP(gn, 13, lane == 0 && level == 0);
  if (lane == 0) {
    gn->lastICPError = sqrtf(res0) / res1;
    gn->lastICPCount = res1;
    double A[36], b[6], x[6];
#pragma unroll
    for (int k = 0; k < 36; ++k) A[k] = S.lastA[k];
#pragma unroll
    for (int k = 0; k < 6; ++k) b[k] = S.lastb[k];
    efm::ldlt_solve_unrolled<6>(A, b, x);
#pragma unroll
    for (int k = 0; k < 6; ++k) S.result[k] = x[k];
    EF_STAMP(gn, 14, level == 0);
    // OdometryProvider::computeUpdateSE3 (OdometryProvider.h:73-96)
    double rvec[3] = {x[3], x[4], x[5]}, Rd[9];
    efm::rodrigues(rvec, Rd);
    const double upd[16] = {Rd[0], Rd[1], Rd[2], x[0], Rd[3], Rd[4], Rd[5], x[1], Rd[6], Rd[7], Rd[8], x[2], 0, 0, 0, 1};
#pragma unroll
    for (int k = 0; k < 16; ++k) S.upd[k] = upd[k];
    EF_STAMP(gn, 15, level == 0);
  }
  __syncwarp();
  if (lane < 16) {
    const int r = lane >> 2, c = lane & 3;
    double acc = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) acc += S.upd[r * 4 + k] * S.rRt[k * 4 + c];
    S.nrt[lane] = acc;
    gn->resultRt[lane] = acc;
  }
  // trace (independent of the rest)
  int slot = -1;
  if (od.trace) {
    const int tn = gn->trace_n;
    if (tn < MAX_TRACE) slot = tn;
  }
  __syncwarp();
  if (slot >= 0) {
    EfSolveTrace& t = od.trace[slot];
    for (int k = lane; k < 36; k += 32) {
      t.A_icp[k] = S.A_icp[k];
      t.A_rgb[k] = S.A_rgb[k];
      t.lastA[k] = S.lastA[k];
    }
    if (lane < 6) {
      t.b_icp[lane] = S.b_icp[lane];
      t.b_rgb[lane] = S.b_rgb[lane];
      t.lastb[lane] = S.lastb[lane];
      t.result[lane] = S.result[lane];
    }
    if (lane == 0) {
      t.kind = 0;
      t.level = level;
      t.iter = iter;
      t.rgb_count = gn->rgbSize;
      t.rgb_sigma = gn->sigma;
      t.sigma_val = gn->sigmaVal;
      t.icp_residual[0] = res0;
      t.icp_residual[1] = res1;
      gn->trace_n = slot + 1;
    }
  }
  // currentT = T_prev * rgbOdom^-1 in float (RGBDOdometry.cpp:543-551): iR = R^T, it = -(R^T t)
  if (lane < 9) {
    const int r = lane / 3, c = lane % 3;
    S.iR[lane] = (float)S.nrt[c * 4 + r];
  }
  __syncwarp();
  if (lane < 3) {
    const float ot0 = (float)S.nrt[3], ot1 = (float)S.nrt[7], ot2 = (float)S.nrt[11];
    S.it[lane] = -(S.iR[lane * 3 + 0] * ot0 + S.iR[lane * 3 + 1] * ot1 + S.iR[lane * 3 + 2] * ot2);
  }
  __syncwarp();
  if (lane >= 12 && lane < 21) gn->Mcp[lane - 12] = S.iR[lane - 12];  // Rprev^-1 Rcurr = the inverse increment itself
  if (lane >= 21 && lane < 24) gn->tcp[lane - 21] = S.it[lane - 21];
  if (lane < 9) {
    const int r = lane / 3, c = lane % 3;
    gn->Rcurr[lane] = S.Rprev[r * 3 + 0] * S.iR[0 * 3 + c] + S.Rprev[r * 3 + 1] * S.iR[1 * 3 + c] + S.Rprev[r * 3 + 2] * S.iR[2 * 3 + c];
  } else if (lane < 12) {
    const int r = lane - 9;
    gn->tcurr[r] = (S.Rprev[r * 3 + 0] * S.it[0] + S.Rprev[r * 3 + 1] * S.it[1] + S.Rprev[r * 3 + 2] * S.it[2]) + S.tprev[r];
  }
  // next iteration's warp matrices (RGBDOdometry.cpp:407-417): KRK^-1 and K t of resultRt^-1 = [R^T | -R^T t]
  if (next_level >= 0) {
    if (lane == 0) level_K(gn, next_level, S.K, S.Kinv);
    if (lane < 9) {
      const int r = lane / 3, c = lane % 3;
      S.Rt[r * 4 + c] = S.nrt[c * 4 + r];
    }
    __syncwarp();
    if (lane < 3) S.Rt[lane * 4 + 3] = -(S.Rt[lane * 4 + 0] * S.nrt[3] + S.Rt[lane * 4 + 1] * S.nrt[7] + S.Rt[lane * 4 + 2] * S.nrt[11]);
    __syncwarp();
    if (lane < 9) {
      const int r = lane / 3, c = lane % 3;
      S.tmp[lane] = S.K[r * 3 + 0] * S.Rt[0 * 4 + c] + S.K[r * 3 + 1] * S.Rt[1 * 4 + c] + S.K[r * 3 + 2] * S.Rt[2 * 4 + c];
    }
    __syncwarp();
    if (lane < 9) {
      const int r = lane / 3, c = lane % 3;
      gn->krkinv[lane] = (float)(S.tmp[r * 3 + 0] * S.Kinv[0 * 3 + c] + S.tmp[r * 3 + 1] * S.Kinv[1 * 3 + c] + S.tmp[r * 3 + 2] * S.Kinv[2 * 3 + c]);
    } else if (lane < 12) {
      const int r = lane - 9;
      gn->kt[r] = (float)(S.K[r * 3 + 0] * S.Rt[3] + S.K[r * 3 + 1] * S.Rt[7] + S.K[r * 3 + 2] * S.Rt[11]);
    }
  }
}

// =============================================================================================
// reductions
// =============================================================================================

constexpr int IT1_THREADS = 128;  // 640x480/4 pixel groups = 600 CTAs of 128: one resident wave at 5 CTAs/SM (<=102 registers)
constexpr int IT1_CTAS_PER_SM = 5;
constexpr int IT2_THREADS = 256;

// ---- geometric row: ICPReduction::search/getProducts (reduce.cu:224-331) ------------------------------------
// The reference moves the live vertex to the world (Rcurr, tcurr), back into the previous camera (Rprev^-1, tprev) to
// project it, gathers the model point from WORLD-frame maps and rotates that back into the previous camera as well
// (three 3x3 transforms + two extra for the normals, per pixel and iteration). All of it happens in one rigid frame
// here: M = Rprev^-1 Rcurr and t' = Rprev^-1 (tcurr - tprev) are formed once per iteration (they are the inverse
// increment the solver already has), the model maps stay in the previous camera's frame (no tranformMaps pass), and the
// distance / angle gates compare squared norms (both are rotation invariant). Same correspondences, rows and sums up to
// float rounding (parity tests: 1e-4 relative on A and b, as north_star asks; inlier counts within borderline flips).
struct IcpFrame {
  m33 M;
  f3 t;
  float fx, fy, cx, cy;
  float distThres2, angleThres2;
};

__device__ __forceinline__ void accumulate29(const float row[7], float (&acc)[29]) {
  int k = 0;
#pragma unroll
  for (int i = 0; i < 6; ++i)
#pragma unroll
    for (int j = i; j < 7; ++j) acc[k++] += row[i] * row[j];
  acc[27] += row[6] * row[6];
  acc[28] += 1.0f;
}

// a / z and b / z, both correctly rounded: the instruction sequence of an IEEE fp32 division (reciprocal estimate, one
// Newton step, quotient, remainder, correction) with the refined reciprocal shared by the two quotients. Exact for finite
// operands away from the denormal / overflow ranges, which is where depths in metres and pixel coordinates live; z == 0
// (where the reference divides by zero) is rejected by the caller.
__device__ __forceinline__ void div2_rn(float a, float b, float z, float& qa, float& qb) {
  float r0;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(z));
  const float e = __fmaf_rn(-z, r0, 1.0f);
  const float r = __fmaf_rn(r0, e, r0);
  const float qa0 = __fmul_rn(a, r), qb0 = __fmul_rn(b, r);
  qa = __fmaf_rn(r, __fmaf_rn(-z, qa0, a), qa0);
  qb = __fmaf_rn(r, __fmaf_rn(-z, qb0, b), qb0);
}

// projective association of one live vertex: s = its position in the previous camera; returns the model pixel or -1
__device__ __forceinline__ int icp_project(const IcpFrame& F, const f3& vcurr, int rows, int cols, f3& s) {
  s = mul(F.M, vcurr) + F.t;
  float px, py;
  div2_rn(s.x * F.fx, s.y * F.fy, s.z, px, py);
  const int ux = __float2int_rn(px + F.cx);
  const int uy = __float2int_rn(py + F.cy);
  if (ux < 0 || uy < 0 || ux >= cols || uy >= rows || !(s.z > 0)) return -1;
  return uy * cols + ux;
}

// s: live vertex, ncurr: live normal (current camera), d / n: model vertex and normal (previous camera)
__device__ __forceinline__ void icp_accumulate(const IcpFrame& F, const f3& s, const f3& ncurr, const f3& d, const f3& n, float (&acc)[29]) {
  const f3 nc = mul(F.M, ncurr);
  const f3 e = s - d;
  const f3 cr = cross(nc, n);
  if (!(dot(cr, cr) < F.angleThres2 && dot(e, e) <= F.distThres2 && !isnan(ncurr.x) && !isnan(n.x))) return;
  const f3 c = cross(s, n);
  const float row[7] = {n.x, n.y, n.z, c.x, c.y, c.z, dot(n, e)};
  accumulate29(row, acc);
}

// First launch of a Gauss-Newton iteration. Two independent jobs share the launch so their memory round trips overlap:
//  (a) photometric correspondences for this pose over the per-frame candidate list (RGBResidual::getProducts,
//      reduce.cu:661-697; the pose-independent gates were applied when the list was built): writes one compact term per
//      candidate and the CTA's {count, sum int(diff^2)};
//  (b) the dense geometric rows (ICPReduction, reduce.cu:224-331) reduced to one 29-float partial per CTA.
// vb / nvb: index and number of the (virtual) CTAs sharing the pass -- the launch grid for the stand-alone kernel, the
// persistent grid for k_gn_loop. sred: 32 * THREADS/32 floats of shared memory.
// First grid-stride round of the live maps of one thread (six 16-byte pieces), copied global -> shared asynchronously
// (cp.async / LDGSTS: no registers are held while the copy is in flight) before griddepcontrol.wait. pre == nullptr: not staged.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

template <int THREADS>
__device__ __forceinline__ void iter1_body(const OdomDev& od, int level, int do_res, int do_icp, int vb, int nvb, float* sred, const float4* pre /* [6][THREADS] in shared memory, or nullptr */,
                                           int pre_cand_have = 0, int pre_base = 0, int pre_ncand = 0, int4 pre_cand = make_int4(0, 0, 0, 0)) {
  GNState* gn = od.gn;
  const int rows = od.rows[level], cols = od.cols[level];
  const int N = rows * cols;
  const size_t plane = (size_t)N;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  // work items are dealt to warps round-robin ACROSS the CTAs (warp w of CTA b is global warp w * nvb + b), so a pass
  // smaller than the grid still occupies every SM evenly
  const int gid = (wid * nvb + vb) * 32 + lane, gstride = nvb * THREADS;

  [[maybe_unused]] const bool stamp = (vb == 0 && threadIdx.x == 0 && level == 0);
  EF_STAMP(gn, 0, stamp);
  unsigned int cnt = 0, sig = 0;
  const bool vec = (cols & 3) == 0;
  if (do_res) {
    // (bounds and this thread's first candidate may have been read ahead of the dependency wait: see k_iter1)
    const int base = pre_cand_have ? pre_base : gn->cand_base[level];
    const int ncand = pre_cand_have ? pre_ncand : gn->cand_base[level + 1] - base;
    const m33 krkinv = load_m33(gn->krkinv);
    const f3 kt = mk3(gn->kt[0], gn->kt[1], gn->kt[2]);
    const float* __restrict__ lastDepth = od.lastDepth[level];
    const uint8_t* __restrict__ lastImage = od.lastImage[level];
    const int4* __restrict__ cand = od.cand + base;
    int4* terms = od.terms + base;
    for (int c = gid; c < ncand; c += gstride) {
      const int4 cr = (pre_cand_have && c == gid) ? pre_cand : cand[c];
      const int k = cr.x;
      const int y = k / cols, x = k - y * cols;
      const float d1 = __int_as_float(cr.y);
      const float transformed_d1 = (float)(d1 * (krkinv.r[2].x * x + krkinv.r[2].y * y + krkinv.r[2].z) + kt.z);
      const int u0 = __float2int_rn((d1 * (krkinv.r[0].x * x + krkinv.r[0].y * y + krkinv.r[0].z) + kt.x) / transformed_d1);
      const int v0 = __float2int_rn((d1 * (krkinv.r[1].x * x + krkinv.r[1].y * y + krkinv.r[1].z) + kt.y) / transformed_d1);
      int4 out = make_int4(-1, 0, cr.z, 0);
      if (u0 >= 0 && v0 >= 0 && u0 < cols && v0 < rows) {
        const float d0 = lastDepth[(size_t)v0 * cols + u0];
        const int li = lastImage[(size_t)v0 * cols + u0];
        if (d0 > 0 && fabsf(transformed_d1 - d0) <= od.maxDepthDeltaRGB && li != 0) {
          const float diff = (float)cr.w - (float)li;
          out.x = (u0 & 0xffff) | (v0 << 16);
          out.y = __float_as_int(diff);
          out.w = __float_as_int(d0);
          cnt += 1;
          sig += (unsigned int)__float2int_rz(diff * diff);
        }
      }
      terms[c] = out;
    }
  }

  EF_STAMP(gn, 1, stamp);
  if (do_icp) {
    IcpFrame F;
    F.M = load_m33(gn->Mcp);
    F.t = mk3(gn->tcp[0], gn->tcp[1], gn->tcp[2]);
    {
      const int div = 1 << level;
      F.fx = gn->fx / div;
      F.fy = gn->fy / div;
      F.cx = gn->cx / div;
      F.cy = gn->cy / div;
    }
    F.distThres2 = od.distThres * od.distThres;
    F.angleThres2 = od.angleThres * od.angleThres;
    float acc[29];
#pragma unroll
    for (int k = 0; k < 29; ++k) acc[k] = 0.f;
    const float* __restrict__ vc = od.vmap_curr[level];
    const float* __restrict__ nc = od.nmap_curr[level];
    const float* __restrict__ vp = od.vmap_c_prev[level];
    const float* __restrict__ np_ = od.nmap_c_prev[level];
    if (vec) {
      // 4 consecutive pixels per thread: six 128-bit coalesced loads for the live maps, the model maps gathered in pairs
      const int ngroups = N >> 2;
      // all four projections first, then all 24 gathers in flight together (the pass is bound by memory round trips, not
      // by issue slots: measured 17.8 -> 15.9 us at 1280x960 against gathering pixel pairs)
      auto group = [&](const float4& vx4, const float4& vy4, const float4& vz4, const float4& nx4, const float4& ny4, const float4& nz4) {
        const float vxs[4] = {vx4.x, vx4.y, vx4.z, vx4.w}, vys[4] = {vy4.x, vy4.y, vy4.z, vy4.w}, vzs[4] = {vz4.x, vz4.y, vz4.z, vz4.w};
        const float nxs[4] = {nx4.x, nx4.y, nx4.z, nx4.w}, nys[4] = {ny4.x, ny4.y, ny4.z, ny4.w}, nzs[4] = {nz4.x, nz4.y, nz4.z, nz4.w};
        f3 sv[4];
        int qv[4];
        float gm[4][6];
#pragma unroll
        for (int h = 0; h < 4; ++h) qv[h] = icp_project(F, mk3(vxs[h], vys[h], vzs[h]), rows, cols, sv[h]);
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          const int a = qv[h] < 0 ? 0 : qv[h];
          gm[h][0] = __ldg(vp + a);
          gm[h][1] = __ldg(vp + plane + a);
          gm[h][2] = __ldg(vp + 2 * plane + a);
          gm[h][3] = __ldg(np_ + a);
          gm[h][4] = __ldg(np_ + plane + a);
          gm[h][5] = __ldg(np_ + 2 * plane + a);
        }
#pragma unroll
        for (int h = 0; h < 4; ++h)
          if (qv[h] >= 0) icp_accumulate(F, sv[h], mk3(nxs[h], nys[h], nzs[h]), mk3(gm[h][0], gm[h][1], gm[h][2]), mk3(gm[h][3], gm[h][4], gm[h][5]), acc);
      };
      int g = gid;
      if (pre && g < ngroups) {  // this thread's first round was staged in shared memory ahead of the dependency wait
        cp_async_wait_all();
        group(pre[0 * THREADS + threadIdx.x], pre[1 * THREADS + threadIdx.x], pre[2 * THREADS + threadIdx.x], pre[3 * THREADS + threadIdx.x],
              pre[4 * THREADS + threadIdx.x], pre[5 * THREADS + threadIdx.x]);
        g += gstride;
      }
      // (the remaining rounds -- three more per thread at 1280x960 -- run the loop without the staging test in it)
      for (; g < ngroups; g += gstride) {
        const int i0 = g << 2;
        group(*reinterpret_cast<const float4*>(vc + i0), *reinterpret_cast<const float4*>(vc + plane + i0),
              *reinterpret_cast<const float4*>(vc + 2 * plane + i0), *reinterpret_cast<const float4*>(nc + i0),
              *reinterpret_cast<const float4*>(nc + plane + i0), *reinterpret_cast<const float4*>(nc + 2 * plane + i0));
      }
    } else {
      for (int i = gid; i < N; i += gstride) {
        f3 sp;
        const int q = icp_project(F, mk3(vc[i], vc[i + plane], vc[i + 2 * plane]), rows, cols, sp);
        if (q >= 0)
          icp_accumulate(F, sp, mk3(nc[i], nc[i + plane], nc[i + 2 * plane]), mk3(vp[q], vp[q + plane], vp[q + 2 * plane]),
                         mk3(np_[q], np_[q + plane], np_[q + 2 * plane]), acc);
      }
    }
    EF_STAMP(gn, 2, stamp);
    if (do_res) {
      // warp totals of the two ints parked behind the float partials; the barrier inside block_reduce_sum publishes them
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        cnt += __shfl_down_sync(0xffffffffu, cnt, off);
        sig += __shfl_down_sync(0xffffffffu, sig, off);
      }
    }
    __shared__ unsigned int s_stat[2 * 32];
    if (do_res && lane == 0) {
      s_stat[wid] = cnt;
      s_stat[32 + wid] = sig;
    }
    block_reduce_sum<29, THREADS>(acc, sred);
    EF_STAMP(gn, 3, stamp);
    if (threadIdx.x < 29) od.partials[(size_t)vb * PARTIAL_STRIDE + threadIdx.x] = acc[0];
    if (do_res && threadIdx.x == 32) {
      // one integer atomic pair per CTA (wrapping adds like the reference's int2 sums; integer addition is order
      // independent, so the totals stay deterministic)
      unsigned int c = 0, g = 0;
#pragma unroll
      for (int w = 0; w < THREADS / 32; ++w) {
        c += s_stat[w];
        g += s_stat[32 + w];
      }
      if (c | g) {
        atomicAdd(&gn->res_acc[0], c);
        atomicAdd(&gn->res_acc[1], g);
      }
    }
  } else if (do_res) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      cnt += __shfl_down_sync(0xffffffffu, cnt, off);
      sig += __shfl_down_sync(0xffffffffu, sig, off);
    }
    if (lane == 0 && (cnt | sig)) {
      atomicAdd(&gn->res_acc[0], cnt);
      atomicAdd(&gn->res_acc[1], sig);
    }
  }
}

// prefetch != 0: the live vertex / normal maps of the thread's first grid-stride round are loaded BEFORE griddepcontrol.wait,
// i.e. while the predecessor (the previous iteration's k_iter2, whose tail is a single CTA summing and solving) is still
// running. Nothing in flight can be writing them: they were produced by the frame's input side, and the host puts one plain
// (non-programmatic) launch between that and the first k_iter1 (odom_track_async: k_gn_begin), so every kernel that wrote
// them had completed before any kernel of the Gauss-Newton loop could become resident.
__global__ void __launch_bounds__(IT1_THREADS, IT1_CTAS_PER_SM) k_iter1(OdomDev od, int level, int do_res, int do_icp, int solve, int prefetch) {
  pdl_launch();
  __shared__ __align__(16) float4 s_pre[6 * IT1_THREADS];
  const float4* pre = nullptr;
  if (prefetch && do_icp && (od.cols[level] & 3) == 0) {
    const int N = od.rows[level] * od.cols[level];
    const int gid = ((threadIdx.x >> 5) * gridDim.x + blockIdx.x) * 32 + (threadIdx.x & 31);
    pre = s_pre;  // (uniform across the CTA; threads beyond the last pixel group never take the first round)
    if (gid < (N >> 2)) {
      const float* __restrict__ vc = od.vmap_curr[level];
      const float* __restrict__ nc = od.nmap_curr[level];
      const size_t plane = (size_t)N;
      const int i0 = gid << 2;
      cp_async16(&s_pre[0 * IT1_THREADS + threadIdx.x], vc + i0);
      cp_async16(&s_pre[1 * IT1_THREADS + threadIdx.x], vc + plane + i0);
      cp_async16(&s_pre[2 * IT1_THREADS + threadIdx.x], vc + 2 * plane + i0);
      cp_async16(&s_pre[3 * IT1_THREADS + threadIdx.x], nc + i0);
      cp_async16(&s_pre[4 * IT1_THREADS + threadIdx.x], nc + plane + i0);
      cp_async16(&s_pre[5 * IT1_THREADS + threadIdx.x], nc + 2 * plane + i0);
    }
  }
  // ... and so were the photometric candidate list and its per-level bounds (launch_sobel runs before that fence): the bounds and
  // this thread's first candidate are loaded here as well, which takes two dependent round trips out of the chain after the wait
  int pre_base = 0, pre_ncand = 0;
  int4 pre_cand = make_int4(0, 0, 0, 0);
  const int pre_cand_have = (prefetch && do_res) ? 1 : 0;
  if (pre_cand_have) {
    const int* __restrict__ cb = od.cand_base;
    pre_base = cb[level];
    pre_ncand = cb[level + 1] - pre_base;
    const int c0 = ((threadIdx.x >> 5) * gridDim.x + blockIdx.x) * 32 + (threadIdx.x & 31);
    if (c0 < pre_ncand) pre_cand = od.cand[pre_base + c0];
  }
  pdl_wait();
  if (solve && od.gn->break_level == level) return;  // rgbOnly `break`: rest of the level is skipped
  __shared__ float sred[32 * (IT1_THREADS / 32)];
  iter1_body<IT1_THREADS>(od, level, do_res, do_icp, blockIdx.x, gridDim.x, sred, pre, pre_cand_have, pre_base, pre_ncand, pre_cand);
}

// ---- photometric row: RGBReduction::getProducts (reduce.cu:419-480); the cloud point is recomputed from the gathered
//      lastDepth with projectPointsKernel's arithmetic (cudafuncs.cu:670-688) -------------------------------------
__device__ __forceinline__ void rgb_accumulate(const int4& term, float sigma, float fx, float fy, float cx, float cy, float sobelScale,
                                               float (&acc)[29]) {
  const float diff = __int_as_float(term.y);
  const int zx = (int)(short)(term.x & 0xffff), zy = term.x >> 16;
  const float z = __int_as_float(term.w);
  const float gx = (float)(short)(term.z & 0xffff), gy = (float)(short)(term.z >> 16);
  float w = sigma + fabsf(diff);
  w = w > 1.19209290E-07F ? 1.0f / w : 1.0f;
  if (sigma == -1) w = 1;
  float row[7];
  row[6] = -w * diff;
  const float invFx = 1.0f / fx, invFy = 1.0f / fy;
  const f3 cp = mk3((float)((zx - cx) * z * invFx), (float)((zy - cy) * z * invFy), z);
  const float invz = (float)(1.0 / (double)cp.z);
  const float dI_dx_val = w * sobelScale * gx;
  const float dI_dy_val = w * sobelScale * gy;
  const float v0 = dI_dx_val * fx * invz;
  const float v1 = dI_dy_val * fy * invz;
  const float v2 = -(v0 * cp.x + v1 * cp.y) * invz;
  row[0] = v0;
  row[1] = v1;
  row[2] = v2;
  row[3] = -cp.z * v1 + cp.y * v2;
  row[4] = cp.z * v0 - cp.x * v2;
  row[5] = -cp.y * v0 + cp.x * v1;
  accumulate29(row, acc);
}

// Second launch of a Gauss-Newton iteration: every CTA first finishes the correspondence statistics of k_iter1 (sigma,
// rgbError and the rgbOnly break decision, RGBDOdometry.cpp:442-455 incl. the operator-precedence quirk), then the
// photometric rows over the candidate terms are reduced; the CTA that takes the last ticket sums all partials in double
// and its first warp solves and updates the pose. mode bits: 1 = rgb rows, 2 = icp partials present, 4 = solve,
// 8 = correspondence statistics present, 16 = use sigma_override.
struct Iter2Shared {
  GnScratch S;
  float sred[32 * 20];  // up to 640 threads
  float sigma;
  int brk;
};

// Everything of the second phase up to the ticket: correspondence statistics (every CTA, redundantly), this CTA's share of
// the dense-pass partials and the photometric rows over its candidates. Returns the rgbOnly `break` decision.
template <int THREADS>
__device__ __forceinline__ bool iter2_rows(const OdomDev& od, Iter2Shared& sh, int level, int iter, int next_level, int nblocks1, int mode,
                                           float sigma_override, int vb, int nvb, int pre_have = 0, int pre_base = 0, int pre_ncand = 0) {
  GnScratch& S = sh.S;
  GNState* gn = od.gn;
  const bool do_rgb = mode & 1, do_icp = mode & 2, solve = mode & 4, have_res = mode & 8;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  [[maybe_unused]] const bool stamp = (threadIdx.x == 0 && level == 0);
  EF_STAMP(gn, 8, stamp && vb == 0);

  // issue the loads that do not depend on sigma first: this CTA's candidate terms and its share of the dense partials
  const int base = pre_have ? pre_base : gn->cand_base[level];
  const int ncand = do_rgb ? (pre_have ? pre_ncand : gn->cand_base[level + 1] - base) : 0;
  const int4* terms = od.terms + base;
  const int c0 = (wid * nvb + vb) * 32 + lane, cstride = nvb * THREADS;
  int4 t0 = make_int4(-1, 0, 0, 0);
  if (c0 < ncand) t0 = terms[c0];
  if (solve && wid == 1) {  // state of the solve: not written by anything in this phase before the last CTA's update
    if (lane < 16) S.rRt[lane] = gn->resultRt[lane];
    if (lane < 9) S.Rprev[lane] = gn->Rprev[lane];
    if (lane < 3) S.tprev[lane] = gn->tprev[lane];
    if (lane == 0) {
      S.flags[0] = gn->icp;
      S.flags[1] = gn->rgb;
      S.w = gn->icpWeight;
    }
  }
  double presum = 0;
  if (do_icp && threadIdx.x < 32) {
    const int per = (nblocks1 + nvb - 1) / nvb;
    const int b0 = vb * per, b1 = min(b0 + per, nblocks1);
    for (int bb = b0; bb < b1; bb += 8) {  // 8 independent loads in flight
      float x[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) x[k] = (bb + k < b1) ? od.partials[(size_t)(bb + k) * PARTIAL_STRIDE + threadIdx.x] : 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) presum += (double)x[k];
    }
  }

  if (threadIdx.x == 0) {
    sh.brk = 0;
    float sig_val = (mode & 16) ? sigma_override : gn->sigmaVal;
    if (have_res) {
      const int rgbSize = (int)gn->res_acc[0], sigma = (int)gn->res_acc[1];
      // reference: std::sqrt((float)sigma / rgbSize == 0 ? 1 : rgbSize)  (RGBDOdometry.cpp:442, App. A-1)
      float sigmaVal = (float)sqrt((double)(((float)sigma / rgbSize == 0) ? 1 : rgbSize));
      const float rgbError = (float)(sqrt((double)sigma) / (rgbSize == 0 ? 1 : rgbSize));
      const float prevError = (iter == 0) ? FLT_MAX : gn->rgbErrBuf[(iter + 1) & 1];  // RGBDOdometry.cpp:404
      const bool brk = solve && gn->rgbOnly && rgbError > prevError;
      if (gn->rgbOnly) sigmaVal = -1;
      if (!(mode & 16)) sig_val = sigmaVal;
      sh.brk = brk ? 1 : 0;
      if (vb == 0) {
        gn->sum_res[0] = rgbSize;
        gn->sum_res[1] = sigma;
        if (solve) {
          gn->rgbSize = rgbSize;
          gn->sigma = sigma;
          if (!brk) {
            gn->rgbErrBuf[iter & 1] = rgbError;
            gn->lastRGBError = rgbError;
            gn->lastRGBCount = (float)rgbSize;
            gn->sigmaVal = sigmaVal;
          }
        }
      }
    }
    sh.sigma = sig_val;
  }
  __syncthreads();
  EF_STAMP(gn, 9, stamp && vb == 0);
  const bool brk = sh.brk != 0;

  if (!brk) {
    if (do_icp && threadIdx.x < 32) od.partials2[vb * 32 + threadIdx.x] = presum;
    if (do_rgb) {
      const float sigma = sh.sigma;
      float lfx, lfy, lcx, lcy;
      level_intr(gn, level, lfx, lfy, lcx, lcy);
      float acc[29];
#pragma unroll
      for (int k = 0; k < 29; ++k) acc[k] = 0.f;
      if (t0.x != -1) rgb_accumulate(t0, sigma, lfx, lfy, lcx, lcy, od.sobelScale, acc);
      for (int c = c0 + cstride; c < ncand; c += cstride) {
        const int4 t = terms[c];
        if (t.x != -1) rgb_accumulate(t, sigma, lfx, lfy, lcx, lcy, od.sobelScale, acc);
      }
      block_reduce_sum<29, THREADS>(acc, sh.sred);
      if (threadIdx.x < 29) od.partials_rgb[vb * 32 + threadIdx.x] = acc[0];
    }
  }
  EF_STAMP(gn, 10, stamp && vb == 0);
  return brk;
}

// The CTA that took the last ticket: re-arm the accumulators, sum all partials in double (fixed order: THREADS/32 slices
// x 32 values) and let its first warp solve and update the pose.
template <int THREADS>
__device__ __forceinline__ void iter2_final(const OdomDev& od, Iter2Shared& sh, int level, int iter, int next_level, int mode, int nvb, bool brk) {
  GnScratch& S = sh.S;
  GNState* gn = od.gn;
  const bool do_rgb = mode & 1, do_icp = mode & 2, solve = mode & 4;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  [[maybe_unused]] const bool stamp = (threadIdx.x == 0 && level == 0);
  EF_STAMP(gn, 11, stamp);
  if (threadIdx.x == 0) {
    *od.counter = 0;
    gn->res_acc[0] = 0u;
    gn->res_acc[1] = 0u;
  }
  if (brk) {
    // rgbOnly `break` (RGBDOdometry.cpp:452-455). Published only here, by the CTA that took the LAST ticket: every CTA of
    // this launch has passed the entry check of k_iter2 by now, so none can see the flag and skip its ticket (the counter
    // and the residual accumulators above are always re-armed exactly once per launch).
    if (solve && threadIdx.x == 0) {
      gn->break_level = level;
      if (next_level >= 0 && next_level != level) gn_prepare_warp(gn, next_level);
    }
    return;
  }
  {
    const int v = lane, sl = wid;
    double a0 = 0, a1 = 0;
    constexpr int SL = THREADS / 32;
    // every row of this thread's slice in ONE batch of independent loads (<= 160 CTA rows / 8 slices = 20 x 2 loads): the sums
    // wait for one L2 round trip instead of three (2.7 -> ~1 us of the serial tail of every iteration)
    constexpr int UN = (MAX_RGB_BLOCKS + SL - 1) / SL;
    for (int bb = sl; bb < nvb; bb += UN * SL) {
      double x[UN];
      float y[UN];
#pragma unroll
      for (int k = 0; k < UN; ++k) {
        const int b = bb + k * SL;
        const bool in = b < nvb;
        x[k] = (in && do_icp) ? od.partials2[b * 32 + v] : 0.0;
        y[k] = (in && do_rgb) ? od.partials_rgb[b * 32 + v] : 0.f;
      }
#pragma unroll
      for (int k = 0; k < UN; ++k) {
        a0 += x[k];
        a1 += (double)y[k];
      }
    }
    S.dsm[sl][v] = a0;
    S.dsm[sl][32 + v] = a1;
  }
  __syncthreads();
  if (threadIdx.x < 64) {
    double t = 0;
#pragma unroll
    for (int k = 0; k < THREADS / 32; ++k) t += S.dsm[k][threadIdx.x];
    const float f = (float)t;
    S.sums[threadIdx.x] = f;
    if (threadIdx.x < 32)
      gn->sum_icp[threadIdx.x] = f;
    else
      gn->sum_rgb[threadIdx.x - 32] = f;
  }
  __syncthreads();
  EF_STAMP(gn, 12, stamp);
  if (!solve || threadIdx.x >= 32) return;
  gn_update_warp(od, S, level, iter, next_level);
  EF_STAMP(gn, 16, stamp);
}

// Second launch of a Gauss-Newton iteration: every CTA first finishes the correspondence statistics of k_iter1 (sigma,
// rgbError and the rgbOnly break decision, RGBDOdometry.cpp:442-455 incl. the operator-precedence quirk), then the
// photometric rows over the candidate terms are reduced; the CTA that takes the last ticket sums all partials in double
// and its first warp solves and updates the pose. mode bits: 1 = rgb rows, 2 = icp partials present, 4 = solve,
// 8 = correspondence statistics present, 16 = use sigma_override.
// 32 = the candidate bounds are final (the launch is part of the tracking loop, behind its fence): read them before the wait.
__global__ void __launch_bounds__(IT2_THREADS) k_iter2(OdomDev od, int level, int iter, int next_level, int nblocks1, int mode, float sigma_override) {
  pdl_launch();
  int pre_base = 0, pre_ncand = 0;
  const int pre_have = (mode & 32) ? 1 : 0;
  if (pre_have) {
    const int* __restrict__ cb = od.cand_base;
    pre_base = cb[level];
    pre_ncand = cb[level + 1] - pre_base;
  }
  pdl_wait();
  __shared__ Iter2Shared sh;
  GNState* gn = od.gn;
  if ((mode & 4) && gn->break_level == level) {
    // rgbOnly `break`: the first iteration of the next level still needs its warp matrices
    if (blockIdx.x == 0 && threadIdx.x == 0 && next_level >= 0 && next_level != level) gn_prepare_warp(gn, next_level);
    return;
  }
  const bool brk = iter2_rows<IT2_THREADS>(od, sh, level, iter, next_level, nblocks1, mode, sigma_override, blockIdx.x, gridDim.x, pre_have, pre_base, pre_ncand);
  // every CTA takes a ticket (also on `break`, so the accumulators of k_iter1 get re-armed exactly once)
  if (!last_block_done(od.counter)) return;
  iter2_final<IT2_THREADS>(od, sh, level, iter, next_level, mode, gridDim.x, brk);
}


// =============================================================================================
// Gauss-Newton iterations inside ONE thread-block cluster
// =============================================================================================
// The coarse pyramid levels are latency chains, not throughput problems: 19 k / 77 k pixels per pass, yet every iteration of
// the two-kernel path above pays two kernel boundaries, a gpu-scope ticket and an L2 round trip for the partials (~16 us per
// iteration for ~3 us of work). Here a whole run of iterations executes in one launch of a single cluster of up to 16 CTAs:
//  * the per-iteration hand-offs are hardware cluster barriers (barrier.cluster.arrive/wait) instead of kernel boundaries;
//  * the correspondence statistics and the per-CTA 29-term partial systems travel through DISTRIBUTED SHARED MEMORY
//    (remote atomics / stores into the leader CTA's shared memory), not through L2;
//  * the statistics barrier is split: CTAs arrive after the photometric correspondence pass, run the dense geometric pass, and
//    only then wait -- the photometric rows need sigma, the geometric rows do not;
//  * the leader sums the <= 16 partials in rank order in double, its first warp solves (gn_update_warp, unchanged) and
//    publishes the 25 floats the next iteration needs through its shared memory.
// Same per-pixel arithmetic as k_iter1 / k_iter2 (icp_project, icp_accumulate, rgb_accumulate); only the grouping of the float
// partial sums differs (<= 16 CTA partials instead of 600), which moves A and b by float rounding.
struct GnSched {
  int n;
  signed char level[24], iter[24];
};
constexpr int GC_THREADS = 512;
constexpr int GC_MAX_CL = 16;

struct GcParams {  // what one iteration needs from the solver
  float Mcp[9], tcp[3], krkinv[9], kt[3];
  int break_level;
};
constexpr int GC_PARAM_WORDS = 25;
struct GcShared {
  GnScratch S;                 // leader: operands of the solve
  float slots[GC_MAX_CL][64];  // leader: per-CTA partial systems, [0,29) geometric and [32,61) photometric
  unsigned int stat[2];        // leader: {count, sum int(diff^2)} of the correspondence pass
  GcParams P;                  // leader: published parameters
  GcParams Pl;                 // every CTA: its copy for the running iteration
  float sred_a[32 * (GC_THREADS / 32)], sred_b[32 * (GC_THREADS / 32)];
  unsigned int s_stat[2];
  float sigma;
  int brk;
};

__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_sync_all() {
  cluster_arrive();
  cluster_wait();
}
__device__ __forceinline__ unsigned int cluster_rank() {
  unsigned int r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ unsigned int cluster_size() {
  unsigned int r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
// generic address of `p` (a shared-memory object of this CTA) in the shared memory of CTA `rank` of the cluster
template <typename T>
__device__ __forceinline__ T* cluster_map(T* p, unsigned int rank) {
  unsigned long long in = (unsigned long long)p, out;
  asm volatile("mapa.u64 %0, %1, %2;" : "=l"(out) : "l"(in), "r"(rank));
  return (T*)out;
}

__global__ void __launch_bounds__(GC_THREADS, 1) k_gn_cluster(OdomDev od, GnSched sched, int s_begin, int s_end, int do_rgb, int do_icp) {
  pdl_enter();
  __shared__ GcShared sh;
  GNState* gn = od.gn;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int rank = (int)cluster_rank(), C = (int)cluster_size();
  GcShared* lead = cluster_map(&sh, 0u);

  // parameters of the first iteration (k_gn_begin or the previous launch left them in the global state)
  if (tid < 9) {
    sh.Pl.Mcp[tid] = gn->Mcp[tid];
    sh.Pl.krkinv[tid] = gn->krkinv[tid];
  } else if (tid < 12) {
    sh.Pl.tcp[tid - 9] = gn->tcp[tid - 9];
    sh.Pl.kt[tid - 9] = gn->kt[tid - 9];
  } else if (tid == 12) {
    sh.Pl.break_level = gn->break_level;
    sh.s_stat[0] = sh.s_stat[1] = 0u;
    sh.stat[0] = sh.stat[1] = 0u;
  }
  if (rank == 0 && wid == 1) {
    GnScratch& S = sh.S;
    if (lane < 16) S.rRt[lane] = gn->resultRt[lane];
    if (lane < 9) S.Rprev[lane] = gn->Rprev[lane];
    if (lane < 3) S.tprev[lane] = gn->tprev[lane];
    if (lane == 0) {
      S.flags[0] = gn->icp;
      S.flags[1] = gn->rgb;
      S.w = gn->icpWeight;
    }
  }
  __syncthreads();
  cluster_sync_all();  // every CTA's shared memory is initialised before anyone touches it remotely

  for (int s = s_begin; s < s_end; ++s) {
    const int level = sched.level[s], iter = sched.iter[s];
    const int next_level = (s + 1 < sched.n) ? sched.level[s + 1] : -1;
    if (sh.Pl.break_level == level) {
      // rgbOnly `break` (RGBDOdometry.cpp:452-455): the rest of the level is skipped; the first iteration of the next level
      // still needs its warp matrices
      if (next_level >= 0 && next_level != level) {
        if (rank == 0 && tid == 0) {
          gn_prepare_warp(gn, next_level);
          for (int k = 0; k < 9; ++k) sh.P.krkinv[k] = gn->krkinv[k];
          for (int k = 0; k < 3; ++k) sh.P.kt[k] = gn->kt[k];
        }
        cluster_sync_all();
        if (tid < 9)
          sh.Pl.krkinv[tid] = lead->P.krkinv[tid];
        else if (tid < 12)
          sh.Pl.kt[tid - 9] = lead->P.kt[tid - 9];
        __syncthreads();
      }
      continue;
    }
    const int rows = od.rows[level], cols = od.cols[level];
    const int N = rows * cols;
    const size_t plane = (size_t)N;
    // work items are dealt to warps round-robin across the CTAs of the cluster
    const int gid = (wid * C + rank) * 32 + lane, gstride = C * GC_THREADS;
    const int base = gn->cand_base[level], ncand = do_rgb ? gn->cand_base[level + 1] - base : 0;
    int4* terms = od.terms + base;

    // ---- (a) photometric correspondences for this pose (RGBResidual::getProducts, reduce.cu:661-697)
    if (do_rgb) {
      unsigned int cnt = 0, sig = 0;
      const m33 krkinv = load_m33(sh.Pl.krkinv);
      const f3 kt = mk3(sh.Pl.kt[0], sh.Pl.kt[1], sh.Pl.kt[2]);
      const float* __restrict__ lastDepth = od.lastDepth[level];
      const uint8_t* __restrict__ lastImage = od.lastImage[level];
      const int4* __restrict__ cand = od.cand + base;
      for (int c = gid; c < ncand; c += gstride) {
        const int4 cr = cand[c];
        const int k = cr.x;
        const int y = k / cols, x = k - y * cols;
        const float d1 = __int_as_float(cr.y);
        const float transformed_d1 = (float)(d1 * (krkinv.r[2].x * x + krkinv.r[2].y * y + krkinv.r[2].z) + kt.z);
        const int u0 = __float2int_rn((d1 * (krkinv.r[0].x * x + krkinv.r[0].y * y + krkinv.r[0].z) + kt.x) / transformed_d1);
        const int v0 = __float2int_rn((d1 * (krkinv.r[1].x * x + krkinv.r[1].y * y + krkinv.r[1].z) + kt.y) / transformed_d1);
        int4 out = make_int4(-1, 0, cr.z, 0);
        if (u0 >= 0 && v0 >= 0 && u0 < cols && v0 < rows) {
          const float d0 = lastDepth[(size_t)v0 * cols + u0];
          const int li = lastImage[(size_t)v0 * cols + u0];
          if (d0 > 0 && fabsf(transformed_d1 - d0) <= od.maxDepthDeltaRGB && li != 0) {
            const float diff = (float)cr.w - (float)li;
            out.x = (u0 & 0xffff) | (v0 << 16);
            out.y = __float_as_int(diff);
            out.w = __float_as_int(d0);
            cnt += 1;
            sig += (unsigned int)__float2int_rz(diff * diff);
          }
        }
        terms[c] = out;
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        cnt += __shfl_down_sync(0xffffffffu, cnt, off);
        sig += __shfl_down_sync(0xffffffffu, sig, off);
      }
      if (lane == 0 && (cnt | sig)) {
        atomicAdd(&sh.s_stat[0], cnt);
        atomicAdd(&sh.s_stat[1], sig);
      }
      __syncthreads();
      if (tid == 0) {  // one pair of remote integer atomics per CTA (wrapping adds: order independent, deterministic)
        atomicAdd(&lead->stat[0], sh.s_stat[0]);
        atomicAdd(&lead->stat[1], sh.s_stat[1]);
        sh.s_stat[0] = sh.s_stat[1] = 0u;
      }
    }
    cluster_arrive();

    // ---- (b) dense geometric rows (ICPReduction, reduce.cu:224-331) while the statistics settle
    float acc[29];
#pragma unroll
    for (int k = 0; k < 29; ++k) acc[k] = 0.f;
    if (do_icp) {
      IcpFrame F;
      F.M = load_m33(sh.Pl.Mcp);
      F.t = mk3(sh.Pl.tcp[0], sh.Pl.tcp[1], sh.Pl.tcp[2]);
      level_intr(gn, level, F.fx, F.fy, F.cx, F.cy);
      F.distThres2 = od.distThres * od.distThres;
      F.angleThres2 = od.angleThres * od.angleThres;
      const float* __restrict__ vc = od.vmap_curr[level];
      const float* __restrict__ nc = od.nmap_curr[level];
      const float* __restrict__ vp = od.vmap_c_prev[level];
      const float* __restrict__ np_ = od.nmap_c_prev[level];
      if ((cols & 3) == 0) {
        const int ngroups = N >> 2;
        for (int g = gid; g < ngroups; g += gstride) {
          const int i0 = g << 2;
          const float4 vx4 = *reinterpret_cast<const float4*>(vc + i0);
          const float4 vy4 = *reinterpret_cast<const float4*>(vc + plane + i0);
          const float4 vz4 = *reinterpret_cast<const float4*>(vc + 2 * plane + i0);
          const float4 nx4 = *reinterpret_cast<const float4*>(nc + i0);
          const float4 ny4 = *reinterpret_cast<const float4*>(nc + plane + i0);
          const float4 nz4 = *reinterpret_cast<const float4*>(nc + 2 * plane + i0);
          const float vxs[4] = {vx4.x, vx4.y, vx4.z, vx4.w}, vys[4] = {vy4.x, vy4.y, vy4.z, vy4.w}, vzs[4] = {vz4.x, vz4.y, vz4.z, vz4.w};
          const float nxs[4] = {nx4.x, nx4.y, nx4.z, nx4.w}, nys[4] = {ny4.x, ny4.y, ny4.z, ny4.w}, nzs[4] = {nz4.x, nz4.y, nz4.z, nz4.w};
          f3 sv[4];
          int qv[4];
          float gm[4][6];
#pragma unroll
          for (int h = 0; h < 4; ++h) qv[h] = icp_project(F, mk3(vxs[h], vys[h], vzs[h]), rows, cols, sv[h]);
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            const int a = qv[h] < 0 ? 0 : qv[h];
            gm[h][0] = __ldg(vp + a);
            gm[h][1] = __ldg(vp + plane + a);
            gm[h][2] = __ldg(vp + 2 * plane + a);
            gm[h][3] = __ldg(np_ + a);
            gm[h][4] = __ldg(np_ + plane + a);
            gm[h][5] = __ldg(np_ + 2 * plane + a);
          }
#pragma unroll
          for (int h = 0; h < 4; ++h)
            if (qv[h] >= 0) icp_accumulate(F, sv[h], mk3(nxs[h], nys[h], nzs[h]), mk3(gm[h][0], gm[h][1], gm[h][2]), mk3(gm[h][3], gm[h][4], gm[h][5]), acc);
        }
      } else {
        for (int i = gid; i < N; i += gstride) {
          f3 sp;
          const int q = icp_project(F, mk3(vc[i], vc[i + plane], vc[i + 2 * plane]), rows, cols, sp);
          if (q >= 0)
            icp_accumulate(F, sp, mk3(nc[i], nc[i + plane], nc[i + 2 * plane]), mk3(vp[q], vp[q + plane], vp[q + 2 * plane]),
                           mk3(np_[q], np_[q + plane], np_[q + 2 * plane]), acc);
        }
      }
    }
    cluster_wait();

    // ---- (c) photometric rows (RGBReduction::getProducts, reduce.cu:419-480) with this iteration's sigma
    float acc2[29];
#pragma unroll
    for (int k = 0; k < 29; ++k) acc2[k] = 0.f;
    if (do_rgb) {
      if (tid == 0) {
        const int rgbSize = (int)lead->stat[0], sigma = (int)lead->stat[1];
        // reference: std::sqrt((float)sigma / rgbSize == 0 ? 1 : rgbSize)  (RGBDOdometry.cpp:442, App. A-1)
        float sigmaVal = (float)sqrt((double)(((float)sigma / rgbSize == 0) ? 1 : rgbSize));
        if (gn->rgbOnly) sigmaVal = -1;
        sh.sigma = sigmaVal;
      }
      __syncthreads();
      const float sigma = sh.sigma;
      float lfx, lfy, lcx, lcy;
      level_intr(gn, level, lfx, lfy, lcx, lcy);
      for (int c = gid; c < ncand; c += gstride) {
        const int4 t = terms[c];
        if (t.x != -1) rgb_accumulate(t, sigma, lfx, lfy, lcx, lcy, od.sobelScale, acc2);
      }
    }
    block_reduce_sum<29, GC_THREADS>(acc, sh.sred_a);
    block_reduce_sum<29, GC_THREADS>(acc2, sh.sred_b);
    if (tid < 29) {
      lead->slots[rank][tid] = acc[0];
      lead->slots[rank][32 + tid] = acc2[0];
    }
    cluster_sync_all();

    // ---- (d) leader: statistics bookkeeping, sums in rank order, solve, publish
    if (rank == 0) {
      GnScratch& S = sh.S;
      if (tid == 0) {
        sh.brk = 0;
        if (do_rgb) {
          const int rgbSize = (int)sh.stat[0], sigma = (int)sh.stat[1];
          float sigmaVal = (float)sqrt((double)(((float)sigma / rgbSize == 0) ? 1 : rgbSize));
          const float rgbError = (float)(sqrt((double)sigma) / (rgbSize == 0 ? 1 : rgbSize));
          const float prevError = (iter == 0) ? FLT_MAX : gn->rgbErrBuf[(iter + 1) & 1];  // RGBDOdometry.cpp:404
          const bool brk = gn->rgbOnly && rgbError > prevError;
          if (gn->rgbOnly) sigmaVal = -1;
          sh.brk = brk ? 1 : 0;
          gn->sum_res[0] = rgbSize;
          gn->sum_res[1] = sigma;
          gn->rgbSize = rgbSize;
          gn->sigma = sigma;
          if (!brk) {
            gn->rgbErrBuf[iter & 1] = rgbError;
            gn->lastRGBError = rgbError;
            gn->lastRGBCount = (float)rgbSize;
            gn->sigmaVal = sigmaVal;
          }
          sh.stat[0] = sh.stat[1] = 0u;
        }
      }
      if (tid >= 64 && tid < 128) {
        const int v = tid - 64;
        double t = 0;
        for (int r = 0; r < C; ++r) t += (double)sh.slots[r][v];
        const float f = (float)t;
        S.sums[v] = f;
        if (v < 32)
          gn->sum_icp[v] = f;
        else
          gn->sum_rgb[v - 32] = f;
      }
      __syncthreads();
      if (sh.brk) {
        if (tid == 0) {
          gn->break_level = level;
          if (next_level >= 0 && next_level != level) gn_prepare_warp(gn, next_level);
        }
      } else if (wid == 0) {
        gn_update_warp(od, S, level, iter, next_level);
        __syncwarp();
        if (lane < 16) S.rRt[lane] = S.nrt[lane];
      }
      __syncthreads();
      if (tid < 9) {
        sh.P.Mcp[tid] = gn->Mcp[tid];
        sh.P.krkinv[tid] = gn->krkinv[tid];
      } else if (tid < 12) {
        sh.P.tcp[tid - 9] = gn->tcp[tid - 9];
        sh.P.kt[tid - 9] = gn->kt[tid - 9];
      } else if (tid == 12) {
        sh.P.break_level = gn->break_level;
      }
    }
    cluster_sync_all();
    if (tid < GC_PARAM_WORDS) reinterpret_cast<int*>(&sh.Pl)[tid] = reinterpret_cast<const int*>(&lead->P)[tid];
    __syncthreads();
  }
  cluster_sync_all();  // nobody leaves while a peer may still read its shared memory
}

// expands the compact per-candidate terms into the reference's dense DataTerm image (inspection / stage API only)
__global__ void k_terms_expand(OdomDev od, int level) {
  pdl_enter();
  const GNState* gn = od.gn;
  const int cols = od.cols[level];
  const int base = gn->cand_base[level], ncand = gn->cand_base[level + 1] - base;
  DataTerm* out = od.corres[level];
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < ncand; c += gridDim.x * blockDim.x) {
    const int4 cr = od.cand[base + c], t = od.terms[base + c];
    if (t.x == -1) continue;
    DataTerm d;
    d.zero_x = (short)(t.x & 0xffff);
    d.zero_y = (short)(t.x >> 16);
    d.one_x = (short)(cr.x % cols);
    d.one_y = (short)(cr.x / cols);
    d.diff = __int_as_float(t.y);
    d.valid = 1;
    out[cr.x] = d;
  }
}

template <int THREADS>
__device__ __forceinline__ void so3_final_sum(const float* partials, int nblocks, float* dst, double* dsm) {
  const int v = threadIdx.x & 31, s = threadIdx.x >> 5;
  double acc = 0;
  if (v < 11)
    for (int b = s; b < nblocks; b += THREADS / 32) acc += (double)partials[(size_t)b * PARTIAL_STRIDE + v];
  dsm[s * 32 + v] = acc;
  __syncthreads();
  if (s == 0 && v < 11) {
    double t = 0;
#pragma unroll
    for (int k = 0; k < THREADS / 32; ++k) t += dsm[k * 32 + v];
    dst[v] = (float)t;
  }
  __syncthreads();
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

// per-CTA partial of one SO3 step (11 floats into od.partials[vb])
// rows of one SO3 step accumulated by this thread: pixels vb * THREADS + tid, stride nvb * THREADS
template <int THREADS>
__device__ __forceinline__ void so3_rows(const OdomDev& od, const m33& imageBasis, const m33& kinv, const m33& krlr, int vb, int nvb, float (&acc)[11]) {
  const int level = 2;
  const int rows = od.rows[level], cols = od.cols[level];
  const int N = rows * cols;
  const uint8_t* lastImage = od.lastNextImage[level];
  const uint8_t* nextImage = od.nextImage[level];
#pragma unroll
  for (int k = 0; k < 11; ++k) acc[k] = 0.f;
  for (int k = vb * THREADS + threadIdx.x; k < N; k += nvb * THREADS) {
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
}

template <int THREADS>
__device__ __forceinline__ void so3_partial(const OdomDev& od, int vb, int nvb, float* sred) {
  const So3State* gn = od.so3s;
  float acc[11];
  so3_rows<THREADS>(od, load_m33(gn->imageBasis), load_m33(gn->kinv), load_m33(gn->krlr), vb, nvb, acc);
  block_reduce_sum<11, THREADS>(acc, sred);
  if (threadIdx.x < 11) od.so3_partials[(size_t)vb * PARTIAL_STRIDE + threadIdx.x] = acc[0];
}

// solve + convergence logic of one SO3 step on the reduced system in gn->sum_so3 (one thread)
__device__ void so3_finish(const OdomDev& od, int iter) {
  So3State* gn = od.so3s;
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
  if (gn->trace_n < SO3_MAX_ITER) {
    EfSolveTrace& t = gn->trace[gn->trace_n++];
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
  so3_prepare(gn, od.gn);
}

__global__ void __launch_bounds__(RED_THREADS) k_so3_step(OdomDev od, int iter, int solve) {
  pdl_enter();
  So3State* st = od.so3s;
  if (solve && st->so3_done) return;
  __shared__ float sred[32 * (RED_THREADS / 32)];
  __shared__ double dsm[(RED_THREADS / 32) * 32];
  so3_partial<RED_THREADS>(od, blockIdx.x, gridDim.x, sred);
  if (!last_block_done(od.so3_counter)) return;
  so3_final_sum<RED_THREADS>(od.so3_partials, gridDim.x, st->sum_so3, dsm);
  if (threadIdx.x != 0) return;
  *od.so3_counter = 0;
  if (!solve) return;
  so3_finish(od, iter);
}

// The whole SO(3) pre-alignment loop (k_so3_begin + up to 10 x k_so3_step) in one launch of one cluster: same protocol as
// k_gn_cluster -- per-CTA 11-term partials into the leader's shared memory, the leader sums them in rank order in double, its
// thread 0 runs the unchanged solve / convergence logic (so3_finish) and publishes the next iteration's three matrices and the
// exit flag through its shared memory. 160x120 pixels of work per iteration; the ten-launch version pays a kernel boundary and a
// gpu-scope ticket for each, also for the iterations after convergence (47 us in total, on the critical path of every caller that
// does not use look-ahead).
struct So3Shared {
  float slots[GC_MAX_CL][16];
  float P[28];   // imageBasis 9, kinv 9, krlr 9, done flag
  float Pl[28];
  float sred[32 * (GC_THREADS / 32)];
};
__global__ void __launch_bounds__(GC_THREADS, 1) k_so3_cluster(OdomDev od) {
  pdl_enter();
  __shared__ So3Shared sh;
  So3State* st = od.so3s;
  const int tid = threadIdx.x;
  const int rank = (int)cluster_rank(), C = (int)cluster_size();
  So3Shared* lead = cluster_map(&sh, 0u);
  if (rank == 0 && tid == 0) {
    // start of the SO(3) loop (RGBDOdometry.cpp:284-303), k_so3_begin's body
    const double I3[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    for (int k = 0; k < 9; ++k) {
      st->resultR[k] = I3[k];
      st->lastResultR[k] = I3[k];
      st->R_lr[k] = (float)I3[k];
    }
    st->so3_lastError = FLT_MAX / 2;
    st->so3_lastCount = FLT_MAX / 2;
    st->so3_done = 0;
    st->trace_n = 0;
    so3_prepare(st, od.gn);
    for (int k = 0; k < 9; ++k) {
      sh.P[k] = st->imageBasis[k];
      sh.P[9 + k] = st->kinv[k];
      sh.P[18 + k] = st->krlr[k];
    }
    sh.P[27] = 0.f;
  }
  __syncthreads();
  cluster_sync_all();
  for (int iter = 0; iter < SO3_MAX_ITER; ++iter) {
    if (tid < 28) sh.Pl[tid] = lead->P[tid];
    __syncthreads();
    if (sh.Pl[27] != 0.f) break;  // converged or diverging: the rest of the loop is skipped (uniform over the cluster)
    float acc[11];
    so3_rows<GC_THREADS>(od, load_m33(sh.Pl), load_m33(sh.Pl + 9), load_m33(sh.Pl + 18), rank, C, acc);
    block_reduce_sum<11, GC_THREADS>(acc, sh.sred);
    if (tid < 11) lead->slots[rank][tid] = acc[0];
    cluster_sync_all();
    if (rank == 0) {
      if (tid < 11) {
        double t = 0;
        for (int r = 0; r < C; ++r) t += (double)sh.slots[r][tid];
        st->sum_so3[tid] = (float)t;
      }
      __syncthreads();
      if (tid == 0) {
        so3_finish(od, iter);
        for (int k = 0; k < 9; ++k) {
          sh.P[k] = st->imageBasis[k];
          sh.P[9 + k] = st->kinv[k];
          sh.P[18 + k] = st->krlr[k];
        }
        sh.P[27] = st->so3_done ? 1.f : 0.f;
      }
    }
    cluster_sync_all();
  }
  cluster_sync_all();  // nobody leaves while a peer may still read its shared memory
}

namespace {

inline int red_blocks(const EfContext* ctx, int n_items, int per_thread, int threads, int ctas_per_sm) {
  int b = (n_items + threads * per_thread - 1) / (threads * per_thread);
  int cap = ctx->num_sms * ctas_per_sm;  // one resident wave
  if (cap > MAX_RED_BLOCKS) cap = MAX_RED_BLOCKS;
  if (b > cap) {
    // more work than one wave: every thread makes the same number of grid-stride rounds (1280x960: 600 CTAs x 4 rounds
    // instead of 740 CTAs of which a quarter would run a 4th round alone)
    const int rounds = (b + cap - 1) / cap;
    b = (b + rounds - 1) / rounds;
  }
  if (b < 1) b = 1;
  return b;
}

inline int iter2_blocks(const EfContext* ctx, int npx, bool rgb, int nb1) {
  int b = rgb ? (npx / 8 + IT2_THREADS - 1) / IT2_THREADS : 1;  // candidates are typically <= 1/8 of the pixels; the loop strides anyway
  const int b_icp = (nb1 + 7) / 8;                              // <= 8 dense-pass partials pre-summed per CTA
  if (b_icp > b) b = b_icp;
  return b < 1 ? 1 : (b > ctx->it2_max_blocks ? ctx->it2_max_blocks : b);
}

#define EF_CHECK_LAST()                          \
  do {                                           \
    cudaError_t e__ = cudaGetLastError();        \
    if (e__ != cudaSuccess) return (int)e__;     \
  } while (0)

}  // namespace

namespace ef {

// the device-resident Gauss-Newton schedule; T_wc in/out lives in gn->T_wc
// The SO(3) pre-alignment loop of tracker `which` on ctx->stream (its state block, partials and ticket are its own)
// one cluster of `cluster` CTAs (grid = cluster), programmatic dependent launch like every other kernel
template <typename... KArgs, typename... Args>
static void ef_launch_cluster(EfContext* ctx, void (*kernel)(KArgs...), int cluster, int block, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(cluster);
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = ctx->stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (ctx->pdl && !ctx->plain_next) ? 2 : 1;
  ctx->plain_next = false;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
  ctx->launches++;
}

// Largest cluster (16, else 8) of k_gn_cluster the device can co-schedule; 0 when clusters are unavailable. Called once per context.
int odom_cluster_size(int want) {
  if (want <= 0) return 0;
  cudaFuncSetAttribute(k_gn_cluster, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  cudaFuncSetAttribute(k_so3_cluster, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  for (int c = want > 8 ? 16 : 8; c >= 8; c >>= 1) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(c);
    cfg.blockDim = dim3(GC_THREADS);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = c;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, k_gn_cluster, &cfg) == cudaSuccess && n >= 1) return c;
    cudaGetLastError();
  }
  return 0;
}

int odom_so3_async(EfContext* ctx, int which) {
  OdomDev& od = ctx->odom[which];
  if (ctx->gn_cluster > 0 && ctx->so3_cluster) {
    ef_launch_cluster(ctx, k_so3_cluster, ctx->gn_cluster, GC_THREADS, od);
    EF_CHECK_LAST();
    return 0;
  }
  EF_LAUNCH(ctx, k_so3_begin, 1, 32, 0, od.so3s, (const GNState*)od.gn);
  const int nb = red_blocks(ctx, od.rows[2] * od.cols[2], 1, RED_THREADS, 2);
  for (int i = 0; i < SO3_MAX_ITER; ++i) EF_LAUNCH(ctx, k_so3_step, nb, RED_THREADS, 0, od, i, 1);
  EF_CHECK_LAST();
  return 0;
}

int odom_track_async(EfContext* ctx, int which, bool rgbOnly, float icpWeight, bool pyramid, bool fastOdom, bool so3) {
  OdomDev& od = ctx->odom[which];
  const bool icp = !rgbOnly && icpWeight > 0;
  const bool rgb = rgbOnly || icpWeight < 100;
  if (rgb) {
    int rc = launch_sobel(ctx, which);
    if (rc) return rc;
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
  if (so3) {
    // already done with the frame's input side (frame loop: on the look-ahead stream when the frame was prefetched)?
    const bool ready = (which == 0) && ctx->so3_ready;
    if (!ready) {
      int rc = odom_so3_async(ctx, which);
      if (rc) return rc;
    }
  }
  if (which == 0) ctx->so3_ready = false;
  // One plain launch between the frame's input side and the Gauss-Newton loop: k_gn_begin starts only when everything before it
  // has completed, so k_iter1 may read the live maps ahead of its dependency wait (see k_iter1).
  const int prefetch = ctx->it1_prefetch ? 1 : 0;
  if (prefetch) EF_PLAIN_NEXT(ctx);
  EF_LAUNCH(ctx, k_gn_begin, 1, GN_BEGIN_THREADS, 0, od.gn, rgbOnly ? 1 : 0, icpWeight, so3 ? 1 : 0, (const So3State*)od.so3s, od.trace,
            ns ? sched_level[0] : 0);
  ctx->maps_dirty[which] = false;
  ef_stage(ctx, 4);
  // the coarse levels (ctx->gn_cluster_levels of them, from the top of the pyramid) run inside one cluster launch
  int s0 = 0;
  if (ctx->gn_cluster > 0 && ns > 0) {
    GnSched sched;
    sched.n = ns;
    for (int s = 0; s < ns; ++s) {
      sched.level[s] = (signed char)sched_level[s];
      sched.iter[s] = (signed char)sched_iter[s];
    }
    while (s0 < ns && sched_level[s0] > NUM_PYRS - 1 - ctx->gn_cluster_levels) ++s0;
    if (s0 > 0) ef_launch_cluster(ctx, k_gn_cluster, ctx->gn_cluster, GC_THREADS, od, sched, 0, s0, rgb ? 1 : 0, icp ? 1 : 0);
  }
  for (int s = s0; s < ns; ++s) {
    const int lv = sched_level[s];
    const int npx = od.rows[lv] * od.cols[lv];
    const int next_lv = (s + 1 < ns) ? sched_level[s + 1] : -1;
    const int nb1 = red_blocks(ctx, npx, 4, IT1_THREADS, IT1_CTAS_PER_SM);
    EF_LAUNCH(ctx, k_iter1, nb1, IT1_THREADS, 0, od, lv, rgb ? 1 : 0, icp ? 1 : 0, 1, prefetch);
    const int nb2 = iter2_blocks(ctx, npx, rgb, icp ? nb1 : 0);
    EF_LAUNCH(ctx, k_iter2, nb2, IT2_THREADS, 0, od, lv, sched_iter[s], next_lv, nb1, (rgb ? 1 | 8 : 0) | (icp ? 2 : 0) | 4 | (prefetch ? 32 : 0), 0.f);
  }
  ef_stage(ctx, 5);
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
  EF_LAUNCH(ctx, k_gn_finish, 1, 32, 0, od.gn, weightMultiplier, have_track ? 1 : 0, which == 0 ? ctx->map.pose : (MapPose*)nullptr);
  EF_CHECK_LAST();
  return 0;
}

int odom_set_pose_async(EfContext* ctx, int which, const double* T_dev) {
  OdomDev& od = ctx->odom[which];
  EF_LAUNCH(ctx, k_set_pose, 1, 32, 0, od.gn, T_dev);
  EF_CHECK_LAST();
  return 0;
}

// Stage-API launches of k_iter1: the live maps may have been written by a kernel that is still in flight (an init stage
// called just before). The first such launch is a plain stream-ordered one (a full barrier); until an init stage runs again
// the maps are static and later launches may read them ahead of their dependency wait.
static int stage_prefetch(EfContext* ctx, int which) {
  if (!ctx->it1_prefetch) return 0;
  if (ctx->maps_dirty[which]) {
    EF_PLAIN_NEXT(ctx);
    ctx->maps_dirty[which] = false;
  }
  return 1;
}

// stand-alone reduction launches for the stage API
int launch_se3_step_raw(EfContext* ctx, int which, int level, bool do_icp, bool do_rgb, float sigma) {
  OdomDev& od = ctx->odom[which];
  const int npx = od.rows[level] * od.cols[level];
  const int nb1 = red_blocks(ctx, npx, 4, IT1_THREADS, IT1_CTAS_PER_SM);
  if (do_icp) {
    const int prefetch = stage_prefetch(ctx, which);
    EF_LAUNCH(ctx, k_iter1, nb1, IT1_THREADS, 0, od, level, 0, 1, 0, prefetch);
  }
  EF_LAUNCH(ctx, k_iter2, iter2_blocks(ctx, npx, do_rgb, do_icp ? nb1 : 0), IT2_THREADS, 0, od, level, 0, -1, nb1, (do_rgb ? 1 | 16 : 0) | (do_icp ? 2 : 0), sigma);
  EF_CHECK_LAST();
  return 0;
}
int launch_icp_dense_only(EfContext* ctx, int which, int level) {
  OdomDev& od = ctx->odom[which];
  const int nb1 = red_blocks(ctx, od.rows[level] * od.cols[level], 4, IT1_THREADS, IT1_CTAS_PER_SM);
  const int prefetch = stage_prefetch(ctx, which);
  EF_LAUNCH(ctx, k_iter1, nb1, IT1_THREADS, 0, od, level, 0, 1, 0, prefetch);
  EF_CHECK_LAST();
  return 0;
}
int launch_rgb_residual_raw(EfContext* ctx, int which, int level) {
  OdomDev& od = ctx->odom[which];
  const int npx = od.rows[level] * od.cols[level];
  const int nb1 = red_blocks(ctx, npx, 4, IT1_THREADS, IT1_CTAS_PER_SM);
  EF_LAUNCH(ctx, k_iter1, nb1, IT1_THREADS, 0, od, level, 1, 0, 0, 0);
  EF_LAUNCH(ctx, k_iter2, 1, IT2_THREADS, 0, od, level, 0, -1, nb1, 8, 0.f);
  cudaError_t e = cudaMemsetAsync(od.corres[level], 0, (size_t)npx * sizeof(DataTerm), ctx->stream);
  if (e != cudaSuccess) return (int)e;
  EF_LAUNCH(ctx, k_terms_expand, 128, 256, 0, od, level);
  EF_CHECK_LAST();
  return 0;
}
int launch_so3_raw(EfContext* ctx, int which) {
  OdomDev& od = ctx->odom[which];
  const int nb = red_blocks(ctx, od.rows[2] * od.cols[2], 1, RED_THREADS, 2);
  EF_LAUNCH(ctx, k_so3_step, nb, RED_THREADS, 0, od, 0, 0);
  EF_CHECK_LAST();
  return 0;
}
}  // namespace ef
