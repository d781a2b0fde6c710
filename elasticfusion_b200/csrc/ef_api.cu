// C ABI of libefusion.so (include/efusion_b200.h): context management, named buffers, the RGBDOdometry stage API and
// the whole-frame orchestration that mirrors ElasticFusion::processFrame (reference Core/ElasticFusion.cpp:270-607).
#include <math.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <new>
#include <utility>
#include <vector>

#include "ef_internal.h"

using namespace ef;

// ---- kernels' host entry points (ef_track.cu, ef_preprocess.cu, ef_map.cu) --------------------------------------
namespace ef {
int odom_init_icp_depth(EfContext* ctx, int which, const uint16_t* depth_dev, float cutoff);
int odom_init_icp_pred(EfContext* ctx, int which, const float* vtx4, const float* nrm4);
int odom_init_icp_model(EfContext* ctx, int which, const float* vtx4, const float* nrm4, const float* vtxB = nullptr,
                        const float* nrmB = nullptr, const int* flag = nullptr, bool with_global = true);
int odom_populate(EfContext* ctx, int which, const uint8_t* rgba, float** destDepths, uint8_t** destImages, bool with_depth,
                  const uint8_t* rgbaB = nullptr, const int* flag = nullptr, bool forceB = false, bool with_image = true);
int odom_cluster_size(int want);
int odom_track_async(EfContext* ctx, int which, bool rgbOnly, float icpWeight, bool pyramid, bool fastOdom, bool so3);
int odom_finish_async(EfContext* ctx, int which, float weightMultiplier, bool have_track);
int odom_set_pose_async(EfContext* ctx, int which, const double* T_dev);
int launch_se3_step_raw(EfContext* ctx, int which, int level, bool do_icp, bool do_rgb, float sigma);
int launch_rgb_residual_raw(EfContext* ctx, int which, int level);
int launch_icp_dense_only(EfContext* ctx, int which, int level);
int launch_so3_raw(EfContext* ctx, int which);
int launch_sobel(EfContext* ctx, int which);
int odom_so3_async(EfContext* ctx, int which);
int preprocess_depth(EfContext* ctx, const uint16_t* raw, float cutoff, uint16_t* filtered, float* metric, float* metric_filtered);
int rgb_to_rgba(EfContext* ctx, const uint8_t* rgb, uint8_t* rgba);

int map_initialise_async(EfContext* ctx);
int map_update_pose_async(EfContext* ctx, const double* T_host_or_null);
int map_predict_indices_async(EfContext* ctx, int time_or_neg, float max_depth, int time_delta, int vis_mode = 0);
int map_fuse_async(EfContext* ctx, int time_or_neg, float max_depth, float weighting_or_neg);
int map_clean_async(EfContext* ctx, int time_or_neg, float conf_threshold, int time_delta, float max_depth, int n_nodes = 0, bool is_fern = false);
int map_set_graph(EfContext* ctx, const float* nodes16, int n_nodes);
int map_loop_constraints_async(EfContext* ctx, int count_thresh, float err_thresh, float cov_thresh);
int map_loop_reset_async(EfContext* ctx);
int odom_copy_pose_async(EfContext* ctx, int dst, int src);
int map_resize_to_host(EfContext* ctx, const void* src_dev, int elem, int factor, void* host_out);
int map_raycast_async(EfContext* ctx, float max_depth, float conf_threshold, int time, int max_time, int time_delta, int mode,
                      bool use_device_tick);
int map_fill_in_async(EfContext* ctx, bool passthrough_geometry, bool passthrough_image);
int map_dense_enough_async(EfContext* ctx);
int map_tick_increment_async(EfContext* ctx);
int map_select_model_inputs(EfContext* ctx, const float** vtx, const float** nrm, const uint8_t** img);
void map_free_host(EfContext* ctx);
}  // namespace ef

#define CU(x)                                  \
  do {                                         \
    cudaError_t e__ = (x);                     \
    if (e__ != cudaSuccess) return (int)e__;   \
  } while (0)
#define RC(x)              \
  do {                     \
    int rc__ = (x);        \
    if (rc__) return rc__; \
  } while (0)

namespace {

struct Arena {
  std::vector<void*> blocks;
  template <typename T>
  cudaError_t alloc(T** p, size_t n) {
    void* q = nullptr;
    cudaError_t e = cudaMalloc(&q, n * sizeof(T) + 256);
    if (e != cudaSuccess) return e;
    blocks.push_back(q);
    *p = (T*)q;
    return cudaSuccess;
  }
};


struct CtxExtra {
  Arena arena;
  std::vector<EfSolveTrace> trace_host;
};

}  // namespace

struct EfContextFull : EfContext {
  CtxExtra extra;
};

static Arena* arena(EfContext* ctx) { return &static_cast<EfContextFull*>(ctx)->extra.arena; }

extern "C" void ef_default_config(EfConfig* c, int width, int height, float fx, float fy, float cx, float cy) {
  memset(c, 0, sizeof(*c));
  c->width = width;
  c->height = height;
  c->fx = fx;
  c->fy = fy;
  c->cx = cx;
  c->cy = cy;
  // ElasticFusion ctor defaults, reference Core/ElasticFusion.h:42-58
  c->time_delta = 200;
  c->count_thresh = 35000;
  c->err_thresh = 5e-05f;
  c->cov_thresh = 1e-05f;
  c->close_loops = 0;
  c->iclnuim = 0;
  c->reloc = 0;
  c->photo_thresh = 115;
  c->confidence = 10;
  c->depth_cutoff = 3;
  c->icp_weight = 10;
  c->fast_odom = 0;
  c->fern_thresh = 0.3095f;
  c->so3 = 1;
  c->frame_to_frame_rgb = 0;
  c->capacity = 3072 * 3072;  // GlobalModel::MAX_VERTICES, reference Core/GlobalModel.cpp:22-24
  c->device = 0;
  c->skip_mid_predict = 1;
}

extern "C" const char* ef_error_string(int code) {
  if (code == 0) return "ok";
  if (code == EF_EINVAL) return "invalid argument";
  if (code == EF_ENOMEM) return "out of memory";
  if (code == EF_ESTATE) return "invalid state";
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  return "unknown error";
}

static int alloc_odom(EfContext* ctx, OdomDev& od) {
  Arena* A = arena(ctx);
  const EfConfig& c = ctx->cfg;
  memset(&od, 0, sizeof(od));
  od.width = c.width;
  od.height = c.height;
  // RGBDOdometry ctor, reference Core/Utils/RGBDOdometry.cpp:22-117 and RGBDOdometry.h:41-42
  od.distThres = 0.10f;
  od.angleThres = sinf(20.f * 3.14159254f / 180.f);
  od.sobelScale = (float)(1.0 / pow(2.0, 3));
  od.maxDepthDeltaRGB = 0.07f;
  od.maxDepthRGB = 6.0f;
  od.minGrad[0] = 5;
  od.minGrad[1] = 3;
  od.minGrad[2] = 1;
  for (int i = 0; i < NUM_PYRS; ++i) {
    od.minScale[i] = (float)(pow((double)od.minGrad[i], 2.0) / pow((double)od.sobelScale, 2.0));
    od.rows[i] = c.height >> i;
    od.cols[i] = c.width >> i;
    const size_t n = (size_t)od.rows[i] * od.cols[i];
    CU(A->alloc(&od.depth_tmp[i], n));
    CU(A->alloc(&od.vmap_g_prev[i], 3 * n));
    CU(A->alloc(&od.nmap_g_prev[i], 3 * n));
    CU(A->alloc(&od.vmap_c_prev[i], 3 * n));
    CU(A->alloc(&od.nmap_c_prev[i], 3 * n));
    CU(A->alloc(&od.vmap_curr[i], 3 * n));
    CU(A->alloc(&od.nmap_curr[i], 3 * n));
    CU(A->alloc(&od.lastDepth[i], n));
    CU(A->alloc(&od.nextDepth[i], n));
    CU(A->alloc(&od.lastImage[i], n));
    CU(A->alloc(&od.nextImage[i], n));
    CU(A->alloc(&od.lastNextImage[i], n));
    CU(A->alloc(&od.dIdx[i], n));
    CU(A->alloc(&od.dIdy[i], n));
    CU(A->alloc(&od.corres[i], n));
    // the reference's cudaMalloc'd maps start uninitialised; NaN / zero fill keeps every first read defined
    CU(cudaMemsetAsync(od.vmap_g_prev[i], 0xff, 3 * n * sizeof(float), ctx->stream));
    CU(cudaMemsetAsync(od.nmap_g_prev[i], 0xff, 3 * n * sizeof(float), ctx->stream));
    CU(cudaMemsetAsync(od.vmap_c_prev[i], 0xff, 3 * n * sizeof(float), ctx->stream));
    CU(cudaMemsetAsync(od.nmap_c_prev[i], 0xff, 3 * n * sizeof(float), ctx->stream));
    CU(cudaMemsetAsync(od.vmap_curr[i], 0xff, 3 * n * sizeof(float), ctx->stream));
    CU(cudaMemsetAsync(od.nmap_curr[i], 0xff, 3 * n * sizeof(float), ctx->stream));
    CU(cudaMemsetAsync(od.lastDepth[i], 0xff, n * sizeof(float), ctx->stream));
    CU(cudaMemsetAsync(od.nextDepth[i], 0xff, n * sizeof(float), ctx->stream));
    CU(cudaMemsetAsync(od.lastImage[i], 0, n, ctx->stream));
    CU(cudaMemsetAsync(od.nextImage[i], 0, n, ctx->stream));
    CU(cudaMemsetAsync(od.lastNextImage[i], 0, n, ctx->stream));
    CU(cudaMemsetAsync(od.dIdx[i], 0, n * 2, ctx->stream));
    CU(cudaMemsetAsync(od.dIdy[i], 0, n * 2, ctx->stream));
    CU(cudaMemsetAsync(od.corres[i], 0, n * sizeof(DataTerm), ctx->stream));
    CU(cudaMemsetAsync(od.depth_tmp[i], 0, n * 2, ctx->stream));
  }
  const size_t n0 = (size_t)c.width * c.height;
  CU(A->alloc(&od.vmaps_tmp, 4 * n0));
  CU(cudaMemsetAsync(od.vmaps_tmp, 0, 4 * n0 * sizeof(float), ctx->stream));
  CU(A->alloc(&od.gn, 1));
  od.cand_base = reinterpret_cast<const int*>(reinterpret_cast<const char*>(od.gn) + offsetof(GNState, cand_base));
  CU(A->alloc(&od.so3s, 1));
  CU(A->alloc(&od.so3_partials, (size_t)MAX_RGB_BLOCKS * PARTIAL_STRIDE));
  CU(A->alloc(&od.so3_counter, 4));
  CU(cudaMemsetAsync(od.so3s, 0, sizeof(So3State), ctx->stream));
  CU(cudaMemsetAsync(od.so3_counter, 0, 16, ctx->stream));
  CU(A->alloc(&od.partials, (size_t)MAX_RED_BLOCKS * PARTIAL_STRIDE));
  CU(A->alloc(&od.partials_rgb, (size_t)MAX_RGB_BLOCKS * 32));
  CU(A->alloc(&od.partials2, (size_t)MAX_RGB_BLOCKS * 32));
  // (the reductions read whole 32-float rows of these; the kernels write the 29 / 11 terms of a system, so the padding lanes
  // are defined once here)
  CU(cudaMemsetAsync(od.so3_partials, 0, (size_t)MAX_RGB_BLOCKS * PARTIAL_STRIDE * sizeof(float), ctx->stream));
  CU(cudaMemsetAsync(od.partials, 0, (size_t)MAX_RED_BLOCKS * PARTIAL_STRIDE * sizeof(float), ctx->stream));
  CU(cudaMemsetAsync(od.partials_rgb, 0, (size_t)MAX_RGB_BLOCKS * 32 * sizeof(float), ctx->stream));
  CU(cudaMemsetAsync(od.partials2, 0, (size_t)MAX_RGB_BLOCKS * 32 * sizeof(double), ctx->stream));
  {
    size_t flat = 0;
    for (int i = 0; i < NUM_PYRS; ++i) {
      od.level_start[i] = (int)flat;
      flat += (size_t)od.rows[i] * od.cols[i];
    }
    od.level_start[NUM_PYRS] = (int)flat;
    CU(A->alloc(&od.cand, flat));
    CU(A->alloc(&od.terms, flat));
    CU(cudaMemsetAsync(od.cand, 0, flat * sizeof(int4), ctx->stream));
    CU(cudaMemsetAsync(od.terms, 0, flat * sizeof(int4), ctx->stream));
  }
  CU(A->alloc(&od.partials_i, (size_t)MAX_RED_BLOCKS * 2));
  CU(A->alloc(&od.counter, 4));
  CU(A->alloc(&od.trace, MAX_TRACE));
  CU(cudaMemsetAsync(od.counter, 0, 16, ctx->stream));
  CU(cudaMemsetAsync(od.trace, 0, sizeof(EfSolveTrace) * MAX_TRACE, ctx->stream));
  GNState g;
  memset(&g, 0, sizeof(g));
  for (int k = 0; k < 16; ++k) g.T_wc[k] = (k % 5 == 0) ? 1.0 : 0.0;
  g.lastICPCount = g.lastRGBCount = g.lastSO3Count = (float)(c.width * c.height);
  g.fx = c.fx;
  g.fy = c.fy;
  g.cx = c.cx;
  g.cy = c.cy;
  for (int lv = 0; lv < NUM_PYRS; ++lv) {
    const int div = 1 << lv;  // CameraModel::operator()(level), reference Core/Cuda/types.cuh:92-95
    const double fx = (double)(c.fx / div), fy = (double)(c.fy / div), cx = (double)(c.cx / div), cy = (double)(c.cy / div);
    const double K[9] = {fx, 0, cx, 0, fy, cy, 0, 0, 1};
    const double ifx = 1.0 / fx, ify = 1.0 / fy;
    const double Ki[9] = {ifx, 0, -cx * ifx, 0, ify, -cy * ify, 0, 0, 1};
    for (int k = 0; k < 9; ++k) {
      g.Kd[lv][k] = K[k];
      g.Kinvd[lv][k] = Ki[k];
    }
  }
  g.break_level = -1;
  g.weighting = 1.0f;
  g.flat_n = od.level_start[NUM_PYRS];
  CU(cudaMemcpyAsync(od.gn, &g, sizeof(g), cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return 0;
}

namespace ef {
int alloc_map(EfContext* ctx);  // ef_map.cu
}
namespace ef {
template <typename T>
cudaError_t ctx_alloc(EfContext* ctx, T** p, size_t n) {
  return arena(ctx)->alloc(p, n);
}
template cudaError_t ctx_alloc<float4>(EfContext*, float4**, size_t);
template cudaError_t ctx_alloc<float>(EfContext*, float**, size_t);
template cudaError_t ctx_alloc<int>(EfContext*, int**, size_t);
template cudaError_t ctx_alloc<unsigned int>(EfContext*, unsigned int**, size_t);
template cudaError_t ctx_alloc<unsigned long long>(EfContext*, unsigned long long**, size_t);
template cudaError_t ctx_alloc<uint8_t>(EfContext*, uint8_t**, size_t);
template cudaError_t ctx_alloc<uint16_t>(EfContext*, uint16_t**, size_t);
template cudaError_t ctx_alloc<uchar4>(EfContext*, uchar4**, size_t);
template cudaError_t ctx_alloc<MapPose>(EfContext*, MapPose**, size_t);
template cudaError_t ctx_alloc<int4>(EfContext*, int4**, size_t);
template cudaError_t ctx_alloc<LoopDev>(EfContext*, LoopDev**, size_t);
}  // namespace ef

extern "C" int ef_create(const EfConfig* cfg, void* stream, EfContext** out) {
  if (!cfg || !out || cfg->width <= 0 || cfg->height <= 0 || cfg->capacity <= 0) return EF_EINVAL;
  // close_loops = 1 runs the LOCAL loop closure front half every frame (ElasticFusion.cpp:447-505; results through
  // ef_local_loop_result). Ferns / relocalisation and the deformation solve stay outside this library.
  if (cfg->reloc) return EF_EINVAL;
  if ((cfg->width >> 2) < 8 || (cfg->height >> 2) < 8) return EF_EINVAL;
  CU(cudaSetDevice(cfg->device));
  EfContextFull* full = new (std::nothrow) EfContextFull();
  if (!full) return EF_ENOMEM;
  EfContext* ctx = full;
  ctx->cfg = *cfg;
  ctx->device = cfg->device;
  {
    int sms = 0;
    cudaError_t e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, cfg->device);
    ctx->num_sms = sms;
    ctx->own_stream = false;
    if (e == cudaSuccess) {
      if (stream) {
        ctx->stream = (cudaStream_t)stream;
      } else {
        e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
        ctx->own_stream = (e == cudaSuccess);
      }
    }
    if (e != cudaSuccess) {  // nothing else has been allocated yet
      delete full;
      return (int)e;
    }
  }
  ctx->launches = 0;
  ctx->so3_ready = false;
  {
    const char* e = getenv("EF_NO_PDL");
    ctx->pdl = !(e && e[0] == '1');
    e = getenv("EF_IT1_PREFETCH");
    ctx->it1_prefetch = !(e && e[0] == '0');
    e = getenv("EF_IT2_MAXBLOCKS");
    ctx->it2_max_blocks = (e && atoi(e) > 0) ? atoi(e) : MAX_RGB_BLOCKS;
    if (ctx->it2_max_blocks > MAX_RGB_BLOCKS) ctx->it2_max_blocks = MAX_RGB_BLOCKS;
    ctx->plain_next = false;
    ctx->maps_dirty[0] = ctx->maps_dirty[1] = true;
    // Gauss-Newton iterations of the coarse pyramid levels inside one thread-block cluster (k_gn_cluster): EF_GN_CLUSTER = wanted
    // cluster size (16 default, 8, or 0 = off), EF_GN_CLUSTER_LEVELS = how many levels from the top of the pyramid (default 1: the 160x120
    // level; measured 313 / 322 / 446 us for the whole loop with 1 / 2 / 3 levels against 338 without -- 16 SMs are too few for the
    // dense pass of the finer levels)
    e = getenv("EF_VISIBLE_LIST");
    ctx->visible_list = !(e && e[0] == '0');
    ctx->vis_pending = false;
    e = getenv("EF_FUSED_MODEL");
    ctx->fused_model_side = !(e && e[0] == '0');
    e = getenv("EF_GN_CLUSTER");
    ctx->gn_cluster = odom_cluster_size(e ? atoi(e) : 16);
    e = getenv("EF_SO3_CLUSTER");
    ctx->so3_cluster = !(e && e[0] == '0');
    e = getenv("EF_GN_CLUSTER_LEVELS");
    ctx->gn_cluster_levels = e ? atoi(e) : 1;
    if (ctx->gn_cluster_levels < 0) ctx->gn_cluster_levels = 0;
    if (ctx->gn_cluster_levels > NUM_PYRS) ctx->gn_cluster_levels = NUM_PYRS;
    e = getenv("EF_STAGE_TIMING");
    ctx->stage_timing = (e && e[0] == '1');
    ctx->stage_n = 0;
    if (ctx->stage_timing)
      for (int i = 0; i < 16; ++i) cudaEventCreate(&ctx->stage_ev[i]);
  }
  ctx->tick = 1;
  for (int k = 0; k < 16; ++k) ctx->T_wc[k] = (k % 5 == 0) ? 1.0 : 0.0;
  ctx->rgb_only = false;
  ctx->icp_weight = cfg->icp_weight;
  ctx->pyramid = true;
  ctx->fast_odom = cfg->fast_odom != 0;
  ctx->so3 = cfg->so3 != 0;
  ctx->frame_to_frame_rgb = cfg->frame_to_frame_rgb != 0;
  ctx->confidence = cfg->confidence;
  ctx->depth_cutoff = cfg->depth_cutoff;
  ctx->max_depth_processed = 20.0f;  // reference Core/ElasticFusion.cpp:83
  ctx->host_count = 0;
  ctx->frame_open = false;

  int rc = alloc_odom(ctx, ctx->odom[0]);
  if (!rc) rc = alloc_odom(ctx, ctx->odom[1]);
  const size_t n = (size_t)cfg->width * cfg->height;
  Arena* A = arena(ctx);
  Textures& t = ctx->tex;
  memset(&t, 0, sizeof(t));
#define TA(p, cnt)                               \
  if (!rc) {                                     \
    cudaError_t e_ = A->alloc(&(p), (cnt));      \
    if (e_ != cudaSuccess) rc = (int)e_;         \
    else cudaMemsetAsync((p), 0, (cnt) * sizeof(*(p)), ctx->stream); \
  }
  TA(t.rgb, n * 3);
  TA(t.rgba, n * 4);
  TA(t.depth_raw, n);
  TA(t.depth_filtered, n);
  TA(t.depth_metric, n);
  TA(t.depth_metric_filtered, n);
  TA(t.index, n);
  TA(t.vert_conf, n);
  TA(t.color_time, n);
  TA(t.norm_rad, n);
  TA(t.image, n);
  TA(t.old_image, n);
  TA(t.fill_image, n);
  TA(t.vertex, n);
  TA(t.normal, n);
  TA(t.old_vertex, n);
  TA(t.old_normal, n);
  TA(t.fill_vertex, n);
  TA(t.fill_normal, n);
  TA(t.time, n);
  TA(t.old_time, n);
  TA(t.synth_depth, n);
  {
    // spare input-side set + side stream of the frame look-ahead
    Lookahead& la = ctx->la;
    memset(&la, 0, sizeof(la));
    TA(la.rgb, n * 3);
    TA(la.rgba, n * 4);
    TA(la.depth_raw, n);
    TA(la.depth_filtered, n);
    TA(la.depth_metric, n);
    TA(la.depth_metric_filtered, n);
    for (int i = 0; i < NUM_PYRS; ++i) {
      const size_t ni = (size_t)ctx->odom[0].rows[i] * ctx->odom[0].cols[i];
      TA(la.depth_tmp[i], ni);
      TA(la.vmap_curr[i], 3 * ni);
      TA(la.nmap_curr[i], 3 * ni);
      TA(la.image[i], ni);
      if (!rc) {
        cudaMemsetAsync(la.vmap_curr[i], 0xff, 3 * ni * sizeof(float), ctx->stream);
        cudaMemsetAsync(la.nmap_curr[i], 0xff, 3 * ni * sizeof(float), ctx->stream);
      }
    }
    TA(la.so3s, 1);
    TA(la.so3_partials, (size_t)MAX_RGB_BLOCKS * PARTIAL_STRIDE);
    TA(la.so3_counter, 4);
    if (!rc) cudaMemsetAsync(la.so3_partials, 0, (size_t)MAX_RGB_BLOCKS * PARTIAL_STRIDE * sizeof(float), ctx->stream);
    if (!rc) {
      cudaError_t e = cudaStreamCreateWithFlags(&la.stream, cudaStreamNonBlocking);
      if (e == cudaSuccess) e = cudaEventCreateWithFlags(&la.ready, cudaEventDisableTiming);
      if (e == cudaSuccess) e = cudaEventCreateWithFlags(&la.spare_free, cudaEventDisableTiming);
      if (e == cudaSuccess) e = cudaEventCreateWithFlags(&la.h2d_done, cudaEventDisableTiming);
      if (e == cudaSuccess) e = cudaEventCreateWithFlags(&la.image_ready, cudaEventDisableTiming);
      if (e == cudaSuccess) e = cudaMallocHost((void**)&la.pin_rgb, n * 3);
      if (e == cudaSuccess) e = cudaMallocHost((void**)&la.pin_depth, n * 2);
      if (e != cudaSuccess) rc = (int)e;
    }
  }
#undef TA
  if (!rc) rc = alloc_map(ctx);
  if (!rc) {
    cudaError_t e = cudaMallocHost((void**)&ctx->pin_rgb, n * 3);
    if (e == cudaSuccess) e = cudaMallocHost((void**)&ctx->pin_depth, n * 2);
    if (e == cudaSuccess) e = cudaMallocHost(&ctx->pin_small, 65536);
    if (e == cudaSuccess) e = cudaMalloc(&ctx->dev_small, 65536);
    if (e != cudaSuccess) rc = (int)e;
  }
  if (!rc) {
    cudaError_t e = cudaEventRecord(ctx->la.spare_free, ctx->stream);
    if (e == cudaSuccess) e = cudaEventRecord(ctx->la.h2d_done, ctx->la.stream);
    if (e == cudaSuccess) e = cudaEventRecord(ctx->la.image_ready, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) rc = (int)e;
  }
  if (rc) {
    ef_destroy(ctx);
    return rc;
  }
  *out = ctx;
  return 0;
}

extern "C" int ef_destroy(EfContext* ctx) {
  if (!ctx) return EF_EINVAL;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->la.stream) {
    cudaStreamSynchronize(ctx->la.stream);
    cudaStreamDestroy(ctx->la.stream);
  }
  if (ctx->la.ready) cudaEventDestroy(ctx->la.ready);
  if (ctx->la.spare_free) cudaEventDestroy(ctx->la.spare_free);
  if (ctx->la.h2d_done) cudaEventDestroy(ctx->la.h2d_done);
  if (ctx->la.image_ready) cudaEventDestroy(ctx->la.image_ready);
  if (ctx->la.pin_rgb) cudaFreeHost(ctx->la.pin_rgb);
  if (ctx->la.pin_depth) cudaFreeHost(ctx->la.pin_depth);
  map_free_host(ctx);
  for (void* p : arena(ctx)->blocks) cudaFree(p);
  if (ctx->pin_rgb) cudaFreeHost(ctx->pin_rgb);
  if (ctx->pin_depth) cudaFreeHost(ctx->pin_depth);
  if (ctx->pin_small) cudaFreeHost(ctx->pin_small);
  if (ctx->dev_small) cudaFree(ctx->dev_small);
  if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
  delete static_cast<EfContextFull*>(ctx);
  return 0;
}

extern "C" void* ef_stream(EfContext* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
extern "C" int ef_sync(EfContext* ctx) {
  if (!ctx) return EF_EINVAL;
  CU(cudaStreamSynchronize(ctx->stream));
  CU(cudaStreamSynchronize(ctx->la.stream));
  return 0;
}
extern "C" int ef_launch_count(EfContext* ctx, int64_t* n) {
  if (!ctx || !n) return EF_EINVAL;
  *n = ctx->launches;
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// named buffers
// ---------------------------------------------------------------------------------------------------------------
extern "C" int ef_buffer(EfContext* ctx, int32_t id, int32_t level, void** dev_ptr, size_t* bytes) {
  if (!ctx) return EF_EINVAL;
  const size_t n = (size_t)ctx->cfg.width * ctx->cfg.height;
  void* p = nullptr;
  size_t b = 0;
  Textures& t = ctx->tex;
  if (id < 40) {
    switch (id) {
      case EF_BUF_RGB: p = t.rgb; b = n * 3; break;
      case EF_BUF_RGBA: p = t.rgba; b = n * 4; break;
      case EF_BUF_DEPTH_RAW: p = t.depth_raw; b = n * 2; break;
      case EF_BUF_DEPTH_FILTERED: p = t.depth_filtered; b = n * 2; break;
      case EF_BUF_DEPTH_METRIC: p = t.depth_metric; b = n * 4; break;
      case EF_BUF_DEPTH_METRIC_FILTERED: p = t.depth_metric_filtered; b = n * 4; break;
      case EF_BUF_INDEX: p = t.index; b = n * 4; break;
      case EF_BUF_VERT_CONF: p = t.vert_conf; b = n * 16; break;
      case EF_BUF_COLOR_TIME: p = t.color_time; b = n * 16; break;
      case EF_BUF_NORM_RAD: p = t.norm_rad; b = n * 16; break;
      case EF_BUF_IMAGE: p = t.image; b = n * 4; break;
      case EF_BUF_VERTEX: p = t.vertex; b = n * 16; break;
      case EF_BUF_NORMAL: p = t.normal; b = n * 16; break;
      case EF_BUF_TIME: p = t.time; b = n * 2; break;
      case EF_BUF_OLD_IMAGE: p = t.old_image; b = n * 4; break;
      case EF_BUF_OLD_VERTEX: p = t.old_vertex; b = n * 16; break;
      case EF_BUF_OLD_NORMAL: p = t.old_normal; b = n * 16; break;
      case EF_BUF_OLD_TIME: p = t.old_time; b = n * 2; break;
      case EF_BUF_SYNTH_DEPTH: p = t.synth_depth; b = n * 4; break;
      case EF_BUF_FILL_IMAGE: p = t.fill_image; b = n * 4; break;
      case EF_BUF_FILL_VERTEX: p = t.fill_vertex; b = n * 16; break;
      case EF_BUF_FILL_NORMAL: p = t.fill_normal; b = n * 16; break;
      default: return EF_EINVAL;
    }
  } else {
    const int which = id / 100, base = id % 100;
    if (which < 0 || which > 1 || level < 0 || level >= NUM_PYRS) return EF_EINVAL;
    OdomDev& od = ctx->odom[which];
    const size_t nl = (size_t)od.rows[level] * od.cols[level];
    switch (base) {
      case EF_BUF_VMAP_CURR: p = od.vmap_curr[level]; b = nl * 12; break;
      case EF_BUF_NMAP_CURR: p = od.nmap_curr[level]; b = nl * 12; break;
      case EF_BUF_VMAP_G_PREV: p = od.vmap_g_prev[level]; b = nl * 12; break;
      case EF_BUF_NMAP_G_PREV: p = od.nmap_g_prev[level]; b = nl * 12; break;
      case EF_BUF_LAST_DEPTH: p = od.lastDepth[level]; b = nl * 4; break;
      case EF_BUF_NEXT_DEPTH: p = od.nextDepth[level]; b = nl * 4; break;
      case EF_BUF_LAST_IMAGE: p = od.lastImage[level]; b = nl; break;
      case EF_BUF_NEXT_IMAGE: p = od.nextImage[level]; b = nl; break;
      case EF_BUF_LAST_NEXT_IMAGE: p = od.lastNextImage[level]; b = nl; break;
      case EF_BUF_DIDX: p = od.dIdx[level]; b = nl * 2; break;
      case EF_BUF_DIDY: p = od.dIdy[level]; b = nl * 2; break;
      case EF_BUF_DEPTH_TMP: p = od.depth_tmp[level]; b = nl * 2; break;
      case EF_BUF_CORRES: p = od.corres[level]; b = nl * 16; break;
      case EF_BUF_VMAPS_TMP: p = od.vmaps_tmp; b = n * 16; break;
      default: return EF_EINVAL;
    }
  }
  if (dev_ptr) *dev_ptr = p;
  if (bytes) *bytes = b;
  return 0;
}

extern "C" int ef_resize(EfContext* ctx, int32_t id, int32_t factor, void* host_out, size_t bytes) {
  if (!ctx || !host_out || factor < 1) return EF_EINVAL;
  void* p;
  size_t b;
  RC(ef_buffer(ctx, id, 0, &p, &b));
  if (id >= 40) return EF_EINVAL;  // full-resolution image attachments only
  const size_t n = (size_t)ctx->cfg.width * ctx->cfg.height;
  const int elem = (int)(b / n);
  if (elem != 2 && elem != 4 && elem != 16) return EF_EINVAL;
  if (bytes < (size_t)(ctx->cfg.width / factor) * (ctx->cfg.height / factor) * elem) return EF_EINVAL;
  return map_resize_to_host(ctx, p, elem, factor, host_out);
}
extern "C" int ef_upload(EfContext* ctx, int32_t id, int32_t level, const void* host, size_t bytes) {
  void* p;
  size_t b;
  RC(ef_buffer(ctx, id, level, &p, &b));
  if (bytes > b || !host) return EF_EINVAL;
  CU(cudaMemcpyAsync(p, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return 0;
}
extern "C" int ef_download(EfContext* ctx, int32_t id, int32_t level, void* host, size_t bytes) {
  void* p;
  size_t b;
  RC(ef_buffer(ctx, id, level, &p, &b));
  if (bytes > b || !host) return EF_EINVAL;
  CU(cudaMemcpyAsync(host, p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// tracker stage API
// ---------------------------------------------------------------------------------------------------------------
static int upload_gn(EfContext* ctx, int which, size_t offset, const void* src, size_t bytes) {
  CU(cudaMemcpyAsync((char*)ctx->odom[which].gn + offset, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
  return 0;
}
static int download_gn(EfContext* ctx, int which, GNState* g) {
  CU(cudaMemcpyAsync(g, ctx->odom[which].gn, sizeof(GNState), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return 0;
}
#define WHICH_OK(w) ((w) == 0 || (w) == 1)

extern "C" int ef_odom_init_icp_depth(EfContext* ctx, int which, const uint16_t* depth_dev, float cutoff) {
  if (!ctx || !WHICH_OK(which) || !depth_dev) return EF_EINVAL;
  return odom_init_icp_depth(ctx, which, depth_dev, cutoff);
}
extern "C" int ef_odom_init_icp_pred(EfContext* ctx, int which, const float* vtx4, const float* nrm4) {
  if (!ctx || !WHICH_OK(which) || !vtx4 || !nrm4) return EF_EINVAL;
  return odom_init_icp_pred(ctx, which, vtx4, nrm4);
}
extern "C" int ef_odom_init_icp_model(EfContext* ctx, int which, const float* vtx4, const float* nrm4, const double* T) {
  if (!ctx || !WHICH_OK(which) || !vtx4 || !nrm4) return EF_EINVAL;
  if (T) {
    RC(upload_gn(ctx, which, offsetof(GNState, T_wc), T, sizeof(double) * 16));
    CU(cudaStreamSynchronize(ctx->stream));  // T is caller memory
  }
  return odom_init_icp_model(ctx, which, vtx4, nrm4);
}
extern "C" int ef_odom_init_rgb(EfContext* ctx, int which, const uint8_t* rgba) {
  if (!ctx || !WHICH_OK(which) || !rgba) return EF_EINVAL;
  OdomDev& od = ctx->odom[which];
  return odom_populate(ctx, which, rgba, od.nextDepth, od.nextImage, true);
}
extern "C" int ef_odom_init_rgb_model(EfContext* ctx, int which, const uint8_t* rgba) {
  if (!ctx || !WHICH_OK(which) || !rgba) return EF_EINVAL;
  OdomDev& od = ctx->odom[which];
  return odom_populate(ctx, which, rgba, od.lastDepth, od.lastImage, true);
}
extern "C" int ef_odom_init_first_rgb(EfContext* ctx, int which, const uint8_t* rgba) {
  if (!ctx || !WHICH_OK(which) || !rgba) return EF_EINVAL;
  OdomDev& od = ctx->odom[which];
  return odom_populate(ctx, which, rgba, nullptr, od.lastNextImage, false);
}

extern "C" int ef_odom_track(EfContext* ctx, int which, double* T_wc, int32_t rgb_only, float icp_weight, int32_t pyramid,
                             int32_t fast_odom, int32_t so3, EfSolveTrace* trace, int32_t max_trace, int32_t* n_trace) {
  if (!ctx || !WHICH_OK(which) || !T_wc) return EF_EINVAL;
  RC(upload_gn(ctx, which, offsetof(GNState, T_wc), T_wc, sizeof(double) * 16));
  CU(cudaStreamSynchronize(ctx->stream));
  RC(odom_track_async(ctx, which, rgb_only != 0, icp_weight, pyramid != 0, fast_odom != 0, so3 != 0));
  RC(odom_finish_async(ctx, which, 1.0f, true));
  GNState g;
  RC(download_gn(ctx, which, &g));
  memcpy(T_wc, g.T_wc, sizeof(double) * 16);
  if (n_trace) *n_trace = g.trace_n < max_trace ? g.trace_n : max_trace;
  if (trace && max_trace > 0) {
    const int n = g.trace_n < max_trace ? g.trace_n : max_trace;
    CU(cudaMemcpyAsync(trace, ctx->odom[which].trace, sizeof(EfSolveTrace) * n, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
  }
  return 0;
}

extern "C" int ef_odom_stats(EfContext* ctx, int which, EfOdomStats* out) {
  if (!ctx || !WHICH_OK(which) || !out) return EF_EINVAL;
  GNState g;
  RC(download_gn(ctx, which, &g));
  out->lastICPError = g.lastICPError;
  out->lastICPCount = g.lastICPCount;
  out->lastRGBError = g.lastRGBError;
  out->lastRGBCount = g.lastRGBCount;
  out->lastSO3Error = g.lastSO3Error;
  out->lastSO3Count = g.lastSO3Count;
  memcpy(out->lastA, g.lastA, sizeof(g.lastA));
  memcpy(out->lastb, g.lastb, sizeof(g.lastb));
  return 0;
}

static void inv6_host(const double* m, double* o) {
  double a[6][12];
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) {
      a[i][j] = m[i * 6 + j];
      a[i][6 + j] = (i == j) ? 1.0 : 0.0;
    }
  for (int c = 0; c < 6; ++c) {
    int p = c;
    for (int r = c + 1; r < 6; ++r)
      if (fabs(a[r][c]) > fabs(a[p][c])) p = r;
    if (p != c)
      for (int j = 0; j < 12; ++j) {
        double t = a[c][j];
        a[c][j] = a[p][j];
        a[p][j] = t;
      }
    double d = a[c][c];
    for (int j = 0; j < 12; ++j) a[c][j] /= d;
    for (int r = 0; r < 6; ++r)
      if (r != c) {
        double f = a[r][c];
        if (f != 0)
          for (int j = 0; j < 12; ++j) a[r][j] -= f * a[c][j];
      }
  }
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) o[i * 6 + j] = a[i][6 + j];
}

extern "C" int ef_odom_covariance(EfContext* ctx, int which, double* cov36) {
  if (!ctx || !WHICH_OK(which) || !cov36) return EF_EINVAL;
  GNState g;
  RC(download_gn(ctx, which, &g));
  inv6_host(g.lastA, cov36);  // lastA.lu().inverse(), reference RGBDOdometry.cpp:573-575
  return 0;
}

static void unpack_se3_host(const float* h, float* A, float* b) {
  int shift = 0;
  for (int i = 0; i < 6; ++i)
    for (int j = i; j < 7; ++j) {
      float value = h[shift++];
      if (j == 6)
        b[i] = value;
      else
        A[j * 6 + i] = A[i * 6 + j] = value;
    }
}

extern "C" int ef_icp_step_async(EfContext* ctx, int which, int level, const float* Rcurr, const float* tcurr, const float* Rprev_inv,
                                 const float* tprev) {
  if (!ctx || !WHICH_OK(which) || level < 0 || level >= NUM_PYRS) return EF_EINVAL;
  if (Rcurr) {
    float* s = (float*)ctx->pin_small;
    CU(cudaStreamSynchronize(ctx->stream));  // the H2D copies of the previous call may still be reading the staging buffer
    memcpy(s, Rcurr, 36);
    memcpy(s + 9, tcurr, 12);
    memcpy(s + 12, Rprev_inv, 36);
    memcpy(s + 21, tprev, 12);
    RC(upload_gn(ctx, which, offsetof(GNState, Rcurr), s, 36));
    RC(upload_gn(ctx, which, offsetof(GNState, tcurr), s + 9, 12));
    RC(upload_gn(ctx, which, offsetof(GNState, Rprev_inv), s + 12, 36));
    RC(upload_gn(ctx, which, offsetof(GNState, tprev), s + 21, 12));
    // the kernel works in the previous camera's frame: M = R_prev^-1 R_curr, t' = R_prev^-1 (t_curr - t_prev)
    float* M = s + 24;
    for (int r = 0; r < 3; ++r) {
      for (int c = 0; c < 3; ++c)
        M[r * 3 + c] = Rprev_inv[r * 3 + 0] * Rcurr[0 * 3 + c] + Rprev_inv[r * 3 + 1] * Rcurr[1 * 3 + c] + Rprev_inv[r * 3 + 2] * Rcurr[2 * 3 + c];
      M[9 + r] = Rprev_inv[r * 3 + 0] * (tcurr[0] - tprev[0]) + Rprev_inv[r * 3 + 1] * (tcurr[1] - tprev[1]) + Rprev_inv[r * 3 + 2] * (tcurr[2] - tprev[2]);
    }
    RC(upload_gn(ctx, which, offsetof(GNState, Mcp), M, 48));
  }
  return launch_se3_step_raw(ctx, which, level, true, false, 0.f);
}

extern "C" int ef_icp_dense_pass_async(EfContext* ctx, int which, int level) {
  if (!ctx || !WHICH_OK(which) || level < 0 || level >= NUM_PYRS) return EF_EINVAL;
  return launch_icp_dense_only(ctx, which, level);
}

extern "C" int ef_icp_step(EfContext* ctx, int which, int level, const float* Rcurr, const float* tcurr, const float* Rprev_inv,
                           const float* tprev, float* A36, float* b6, float* residual2) {
  if (!Rcurr || !tcurr || !Rprev_inv || !tprev || !A36 || !b6 || !residual2) return EF_EINVAL;
  RC(ef_icp_step_async(ctx, which, level, Rcurr, tcurr, Rprev_inv, tprev));
  GNState g;
  RC(download_gn(ctx, which, &g));
  unpack_se3_host(g.sum_icp, A36, b6);
  residual2[0] = g.sum_icp[27];
  residual2[1] = g.sum_icp[28];
  return 0;
}

extern "C" int ef_rgb_residual(EfContext* ctx, int which, int level, const float* krkinv, const float* kt, int32_t* sigma_sum,
                               int32_t* count) {
  if (!ctx || !WHICH_OK(which) || level < 0 || level >= NUM_PYRS || !krkinv || !kt) return EF_EINVAL;
  RC(upload_gn(ctx, which, offsetof(GNState, krkinv), krkinv, 36));
  RC(upload_gn(ctx, which, offsetof(GNState, kt), kt, 12));
  CU(cudaStreamSynchronize(ctx->stream));
  RC(launch_sobel(ctx, which));
  RC(launch_rgb_residual_raw(ctx, which, level));
  GNState g;
  RC(download_gn(ctx, which, &g));
  if (count) *count = g.sum_res[0];
  if (sigma_sum) *sigma_sum = g.sum_res[1];
  return 0;
}

extern "C" int ef_rgb_step(EfContext* ctx, int which, int level, float sigma, float* A36, float* b6) {
  if (!ctx || !WHICH_OK(which) || level < 0 || level >= NUM_PYRS || !A36 || !b6) return EF_EINVAL;
  RC(launch_se3_step_raw(ctx, which, level, false, true, sigma));
  GNState g;
  RC(download_gn(ctx, which, &g));
  unpack_se3_host(g.sum_rgb, A36, b6);
  return 0;
}

extern "C" int ef_so3_step(EfContext* ctx, int which, const float* image_basis, const float* kinv, const float* krlr, float* A9, float* b3,
                           float* residual2) {
  if (!ctx || !WHICH_OK(which) || !image_basis || !kinv || !krlr || !A9 || !b3 || !residual2) return EF_EINVAL;
  char* sdev = (char*)ctx->odom[which].so3s;
  CU(cudaMemcpyAsync(sdev + offsetof(So3State, imageBasis), image_basis, 36, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(sdev + offsetof(So3State, kinv), kinv, 36, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(sdev + offsetof(So3State, krlr), krlr, 36, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  RC(launch_so3_raw(ctx, which));
  struct {
    float sum_so3[12];
  } g;
  CU(cudaMemcpyAsync(g.sum_so3, sdev + offsetof(So3State, sum_so3), sizeof(g.sum_so3), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  int shift = 0;
  for (int i = 0; i < 3; ++i)
    for (int j = i; j < 4; ++j) {
      float value = g.sum_so3[shift++];
      if (j == 3)
        b3[i] = value;
      else
        A9[j * 3 + i] = A9[i * 3 + j] = value;
    }
  residual2[0] = g.sum_so3[9];
  residual2[1] = g.sum_so3[10];
  return 0;
}

extern "C" int ef_preprocess_depth(EfContext* ctx, const uint16_t* raw, float cutoff, uint16_t* filtered, float* metric,
                                   float* metric_filtered) {
  if (!ctx || !raw) return EF_EINVAL;
  return preprocess_depth(ctx, raw, cutoff, filtered, metric, metric_filtered);
}

// ---------------------------------------------------------------------------------------------------------------
// setters / getters
// ---------------------------------------------------------------------------------------------------------------
extern "C" int ef_get_pose(EfContext* ctx, double* T) {
  if (!ctx || !T) return EF_EINVAL;
  CU(cudaMemcpyAsync(T, (char*)ctx->odom[0].gn + offsetof(GNState, T_wc), sizeof(double) * 16, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  memcpy(ctx->T_wc, T, sizeof(double) * 16);
  return 0;
}
extern "C" int ef_set_pose(EfContext* ctx, const double* T) {
  if (!ctx || !T) return EF_EINVAL;
  RC(upload_gn(ctx, 0, offsetof(GNState, T_wc), T, sizeof(double) * 16));
  CU(cudaStreamSynchronize(ctx->stream));
  memcpy(ctx->T_wc, T, sizeof(double) * 16);
  return 0;
}
extern "C" int ef_get_tick(EfContext* ctx, int32_t* tick) {
  if (!ctx || !tick) return EF_EINVAL;
  *tick = ctx->tick;
  return 0;
}
extern "C" int ef_set_tick(EfContext* ctx, int32_t tick) {
  if (!ctx) return EF_EINVAL;
  ctx->tick = tick;
  CU(cudaMemcpyAsync(ctx->map.tick, &ctx->tick, 4, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return 0;
}
extern "C" int ef_set_rgb_only(EfContext* ctx, int32_t v) { if (!ctx) return EF_EINVAL; ctx->rgb_only = v != 0; return 0; }
extern "C" int ef_set_icp_weight(EfContext* ctx, float v) { if (!ctx) return EF_EINVAL; ctx->icp_weight = v; return 0; }
extern "C" int ef_set_pyramid(EfContext* ctx, int32_t v) { if (!ctx) return EF_EINVAL; ctx->pyramid = v != 0; return 0; }
extern "C" int ef_set_fast_odom(EfContext* ctx, int32_t v) { if (!ctx) return EF_EINVAL; ctx->fast_odom = v != 0; return 0; }
extern "C" int ef_set_so3(EfContext* ctx, int32_t v) { if (!ctx) return EF_EINVAL; ctx->so3 = v != 0; return 0; }
extern "C" int ef_set_frame_to_frame_rgb(EfContext* ctx, int32_t v) { if (!ctx) return EF_EINVAL; ctx->frame_to_frame_rgb = v != 0; return 0; }
extern "C" int ef_set_confidence_threshold(EfContext* ctx, float v) { if (!ctx) return EF_EINVAL; ctx->confidence = v; return 0; }
extern "C" int ef_set_depth_cutoff(EfContext* ctx, float v) { if (!ctx) return EF_EINVAL; ctx->depth_cutoff = v; return 0; }

// ---------------------------------------------------------------------------------------------------------------
// map stage API
// ---------------------------------------------------------------------------------------------------------------
namespace ef {
int map_download(EfContext* ctx, const float4* a, const float4* b, const float4* c, int n, float* out);
int map_upload(EfContext* ctx, const float* in, int n);
int map_upload_range(EfContext* ctx, const float* in, int first, int n);
void map_free_host(EfContext* ctx);
}

static int read_count(EfContext* ctx, const int* dev, int* out) {
  CU(cudaMemcpyAsync(out, dev, 4, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return 0;
}

extern "C" int ef_map_initialise(EfContext* ctx) {
  if (!ctx) return EF_EINVAL;
  return map_initialise_async(ctx);
}
extern "C" int ef_map_predict_indices(EfContext* ctx, const double* T, int32_t time, float max_depth, int32_t time_delta) {
  if (!ctx) return EF_EINVAL;
  RC(map_update_pose_async(ctx, T));
  return map_predict_indices_async(ctx, time, max_depth, time_delta);
}
extern "C" int ef_map_fuse(EfContext* ctx, const double* T, int32_t time, float max_depth, float weighting) {
  if (!ctx) return EF_EINVAL;
  RC(map_update_pose_async(ctx, T));
  return map_fuse_async(ctx, time, max_depth, weighting);
}
extern "C" int ef_map_clean(EfContext* ctx, const double* T, int32_t time, float conf_threshold, int32_t time_delta, float max_depth) {
  if (!ctx) return EF_EINVAL;
  RC(map_update_pose_async(ctx, T));
  return map_clean_async(ctx, time, conf_threshold, time_delta, max_depth);
}
extern "C" int ef_map_clean_deform(EfContext* ctx, const double* T, int32_t time, float conf_threshold, int32_t time_delta, float max_depth,
                                   const float* graph_nodes16, int32_t n_nodes, int32_t is_fern) {
  if (!ctx || n_nodes < 0 || (n_nodes > 0 && !graph_nodes16)) return EF_EINVAL;
  RC(map_set_graph(ctx, graph_nodes16, n_nodes));
  RC(map_update_pose_async(ctx, T));
  return map_clean_async(ctx, time, conf_threshold, time_delta, max_depth, n_nodes, is_fern != 0);
}
extern "C" int ef_map_raycast(EfContext* ctx, const double* T, float max_depth, float conf_threshold, int32_t time, int32_t max_time,
                              int32_t time_delta, int32_t mode) {
  if (!ctx || mode < 0 || mode > 2) return EF_EINVAL;
  RC(map_update_pose_async(ctx, T));
  return map_raycast_async(ctx, max_depth, conf_threshold, time, max_time, time_delta, mode, false);
}
extern "C" int ef_map_fill_in(EfContext* ctx, int32_t pass_geom, int32_t pass_img) {
  if (!ctx) return EF_EINVAL;
  return map_fill_in_async(ctx, pass_geom != 0, pass_img != 0);
}
extern "C" int ef_dense_enough(EfContext* ctx, int32_t* out) {
  if (!ctx || !out) return EF_EINVAL;
  RC(map_dense_enough_async(ctx));
  return read_count(ctx, ctx->map.dense_flag, out);
}
extern "C" int ef_map_count(EfContext* ctx, int32_t* count) {
  if (!ctx || !count) return EF_EINVAL;
  RC(read_count(ctx, ctx->map.count, count));
  ctx->host_count = *count;
  return 0;
}
extern "C" int ef_map_download(EfContext* ctx, float* out12, int32_t max_surfels, int32_t* count) {
  if (!ctx || !out12) return EF_EINVAL;
  int n = 0;
  RC(read_count(ctx, ctx->map.count, &n));
  ctx->host_count = n;
  if (count) *count = n;
  if (n > max_surfels) n = max_surfels;
  return map_download(ctx, ctx->map.pos_conf, ctx->map.color_time, ctx->map.norm_rad, n, out12);
}
extern "C" int ef_map_download_new(EfContext* ctx, float* out12, int32_t max_surfels, int32_t* count) {
  if (!ctx || !out12) return EF_EINVAL;
  int n = 0;
  RC(read_count(ctx, ctx->map.new_count, &n));
  if (count) *count = n;
  if (n > max_surfels) n = max_surfels;
  return map_download(ctx, ctx->map.new_pos, ctx->map.new_col, ctx->map.new_nr, n, out12);
}
extern "C" int ef_map_upload_range(EfContext* ctx, const float* in12, int32_t first, int32_t count) {
  if (!ctx || !in12 || first < 0 || count <= 0) return EF_EINVAL;
  return map_upload_range(ctx, in12, first, count);
}
extern "C" int ef_map_upload(EfContext* ctx, const float* in12, int32_t count) {
  if (!ctx || (!in12 && count > 0) || count < 0) return EF_EINVAL;
  return map_upload(ctx, in12, count);
}

// ---------------------------------------------------------------------------------------------------------------
// whole frame
// ---------------------------------------------------------------------------------------------------------------
// ElasticFusion::predict, reference Core/ElasticFusion.cpp:621-653 (lost == false, lastFrameRecovery == false)
static int predict_async(EfContext* ctx) {
  RC(map_raycast_async(ctx, ctx->max_depth_processed, ctx->confidence, ctx->tick, ctx->tick, ctx->cfg.time_delta, 0, false));
  return map_fill_in_async(ctx, false, ctx->frame_to_frame_rgb);
}

extern "C" int ef_predict(EfContext* ctx) {
  if (!ctx) return EF_EINVAL;
  RC(map_update_pose_async(ctx, nullptr));
  return predict_async(ctx);
}

// Input side of a frame: everything that depends neither on the map nor on the pose (ElasticFusion.cpp:278-285 and the
// live half of RGBDOdometry::initICP / initRGB). Runs on ctx->stream into the live buffer set.
static int frame_input_side(EfContext* ctx, const uint8_t* rgb_dev, const uint16_t* depth_dev) {
  const size_t n = (size_t)ctx->cfg.width * ctx->cfg.height;
  Textures& t = ctx->tex;
  // texture uploads, reference ElasticFusion.cpp:278-280
  if (rgb_dev != t.rgb) CU(cudaMemcpyAsync(t.rgb, rgb_dev, n * 3, cudaMemcpyDeviceToDevice, ctx->stream));
  if (depth_dev != t.depth_raw) CU(cudaMemcpyAsync(t.depth_raw, depth_dev, n * 2, cudaMemcpyDeviceToDevice, ctx->stream));
  RC(rgb_to_rgba(ctx, t.rgb, t.rgba));
  // filterDepth + metriciseDepth, ElasticFusion.cpp:284-285
  RC(preprocess_depth(ctx, t.depth_raw, ctx->depth_cutoff, t.depth_filtered, t.depth_metric, t.depth_metric_filtered));
  if (ctx->stream != ctx->la.stream) ef_stage(ctx, 1);  // (stage events live on the main stream only)
  // frameToModel.initICP(filtered depth) and the intensity half of initRGB, ElasticFusion.cpp:318-319
  RC(odom_init_icp_depth(ctx, 0, t.depth_filtered, ctx->max_depth_processed));
  RC(odom_populate(ctx, 0, t.rgba, nullptr, ctx->odom[0].nextImage, false));
  // SO(3) pre-alignment (RGBDOdometry.cpp:305-368): needs only this intensity pyramid and the previous frame's
  // (lastNextImage), so it belongs to the input side. Not for the first frame (nothing is tracked, ElasticFusion.cpp:290).
  ctx->so3_ready = false;
  if (ctx->so3 && ctx->tick > 1) {
    RC(odom_so3_async(ctx, 0));
    ctx->so3_ready = true;
  }
  CU(cudaEventRecord(ctx->la.image_ready, ctx->stream));  // this pyramid is the next frame's lastNextImage
  return 0;
}

static void swap_sides(EfContext* ctx) {
  Lookahead& la = ctx->la;
  Textures& t = ctx->tex;
  OdomDev& od = ctx->odom[0];
  std::swap(t.rgb, la.rgb);
  std::swap(t.rgba, la.rgba);
  std::swap(t.depth_raw, la.depth_raw);
  std::swap(t.depth_filtered, la.depth_filtered);
  std::swap(t.depth_metric, la.depth_metric);
  std::swap(t.depth_metric_filtered, la.depth_metric_filtered);
  for (int i = 0; i < NUM_PYRS; ++i) {
    std::swap(od.depth_tmp[i], la.depth_tmp[i]);
    std::swap(od.vmap_curr[i], la.vmap_curr[i]);
    std::swap(od.nmap_curr[i], la.nmap_curr[i]);
    std::swap(od.nextImage[i], la.image[i]);
  }
  std::swap(od.so3s, la.so3s);
  std::swap(od.so3_partials, la.so3_partials);
  std::swap(od.so3_counter, la.so3_counter);
  std::swap(ctx->so3_ready, la.so3_ready);
}

// rgb/depth: device pointers, or pinned host staging when from_host
static int prefetch_common(EfContext* ctx, const uint8_t* rgb, const uint16_t* depth, bool from_host) {
  Lookahead& la = ctx->la;
  if (la.pending) return EF_ESTATE;
  const size_t n = (size_t)ctx->cfg.width * ctx->cfg.height;
  cudaStream_t main_stream = ctx->stream;
  swap_sides(ctx);  // the helpers below write the live pointers: make the spare set live while enqueueing
  ctx->stream = la.stream;
  int rc = 0;
  cudaError_t e = cudaStreamWaitEvent(la.stream, la.spare_free, 0);
  if (e == cudaSuccess) e = cudaStreamWaitEvent(la.stream, la.image_ready, 0);  // previous frame's intensity pyramid (SO(3) input)
  if (e == cudaSuccess && from_host) {
    e = cudaMemcpyAsync(ctx->tex.rgb, rgb, n * 3, cudaMemcpyHostToDevice, la.stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(ctx->tex.depth_raw, depth, n * 2, cudaMemcpyHostToDevice, la.stream);
    if (e == cudaSuccess) e = cudaEventRecord(la.h2d_done, la.stream);
    rgb = ctx->tex.rgb;
    depth = ctx->tex.depth_raw;
  }
  if (e != cudaSuccess) rc = (int)e;
  if (!rc) rc = frame_input_side(ctx, rgb, depth);
  if (!rc) {
    e = cudaEventRecord(la.ready, la.stream);
    if (e != cudaSuccess) rc = (int)e;
  }
  ctx->stream = main_stream;
  swap_sides(ctx);
  if (!rc) la.pending = true;
  return rc;
}

extern "C" int ef_prefetch_frame_device(EfContext* ctx, const uint8_t* rgb_dev, const uint16_t* depth_dev) {
  if (!ctx || !rgb_dev || !depth_dev) return EF_EINVAL;
  return prefetch_common(ctx, rgb_dev, depth_dev, false);
}

extern "C" int ef_prefetch_frame(EfContext* ctx, const uint8_t* rgb, const uint16_t* depth) {
  if (!ctx || !rgb || !depth) return EF_EINVAL;
  if (ctx->la.pending) return EF_ESTATE;
  const size_t n = (size_t)ctx->cfg.width * ctx->cfg.height;
  CU(cudaEventSynchronize(ctx->la.h2d_done));  // the staging buffers of the previous prefetch have been read
  memcpy(ctx->la.pin_rgb, rgb, n * 3);
  memcpy(ctx->la.pin_depth, depth, n * 2);
  return prefetch_common(ctx, ctx->la.pin_rgb, ctx->la.pin_depth, true);
}

// Local loop closure front half (ElasticFusion.cpp:447-505, the branch taken when no fern matched): INACTIVE prediction,
// modelToModel in the reference's order (initICPModel -> initRGBModel -> initICP(pred, pred) -> initRGB, App. A-2),
// getIncrementalTransformation(rgbOnly = false, icpWeight = 10, so3 = false), acceptance test and constraint sampling.
// Everything stays on the device (LoopDev); nothing is applied to the map or the pose.
static int local_loop_async(EfContext* ctx) {
  Textures& t = ctx->tex;
  OdomDev& od = ctx->odom[1];
  RC(map_raycast_async(ctx, ctx->max_depth_processed, ctx->confidence, 0, ctx->tick - ctx->cfg.time_delta, ctx->cfg.time_delta, 1, false));
  RC(odom_copy_pose_async(ctx, 1, 0));
  RC(odom_init_icp_model(ctx, 1, (const float*)t.old_vertex, (const float*)t.old_normal, nullptr, nullptr, nullptr, false));
  RC(odom_populate(ctx, 1, (const uint8_t*)t.old_image, od.lastDepth, od.lastImage, true));
  RC(odom_init_icp_pred(ctx, 1, (const float*)t.vertex, (const float*)t.normal));
  RC(odom_populate(ctx, 1, (const uint8_t*)t.image, od.nextDepth, od.nextImage, true));
  RC(odom_track_async(ctx, 1, false, 10.0f, ctx->pyramid, ctx->fast_odom, false));
  RC(odom_finish_async(ctx, 1, 1.0f, true));
  return map_loop_constraints_async(ctx, ctx->cfg.count_thresh, ctx->cfg.err_thresh, ctx->cfg.cov_thresh);
}

// First half of processFrame (ElasticFusion.cpp:270-534): input side, first-frame map / tracking, velocity weighting, the
// mid-frame predict() and the local loop closure front half. Asynchronous.
static int frame_begin_device(EfContext* ctx, const uint8_t* rgb_dev, const uint16_t* depth_dev, float weight_multiplier, const double* in_T_wc) {
  Textures& t = ctx->tex;
  Lookahead& la = ctx->la;
  if (ctx->frame_open) return EF_ESTATE;  // ef_process_frame_end has not been called for the previous frame
  ef_stage(ctx, 0);
  if (!rgb_dev) {
    // consume the prefetched frame: its buffer set becomes live, the previous live set becomes the spare one
    if (!la.pending) return EF_ESTATE;
    swap_sides(ctx);
    CU(cudaEventRecord(la.spare_free, ctx->stream));
    CU(cudaStreamWaitEvent(ctx->stream, la.ready, 0));
    la.pending = false;
  } else {
    if (la.pending) return EF_ESTATE;  // a prefetched frame must be consumed first (pass NULL, NULL)
    RC(frame_input_side(ctx, rgb_dev, depth_dev));
  }
  ef_stage(ctx, 2);
  ctx->frame_open = true;
  if (ctx->cfg.close_loops) RC(map_loop_reset_async(ctx));

  if (ctx->tick == 1) {
    // ElasticFusion.cpp:290-296; initFirstRGB: the intensity pyramid of the first frame is the SO(3) "last" image
    RC(map_initialise_async(ctx));
    for (int i = 0; i < NUM_PYRS; ++i) std::swap(ctx->odom[0].nextImage[i], ctx->odom[0].lastNextImage[i]);
    return 0;
  }
  OdomDev& od = ctx->odom[0];
  if (!in_T_wc) {
    // ElasticFusion.cpp:302-323. The fill-in decision stays on the device: both candidate inputs are handed to the
    // pyramid kernels together with the flag.
    RC(map_dense_enough_async(ctx));
    RC(map_select_model_inputs(ctx, nullptr, nullptr, nullptr));
    // initRGB's depth half (populateRGBDData -> verticesToDepth(vmaps_tmp), RGBDOdometry.cpp:212-222) reads the SAME
    // vmaps_tmp initICPModel just filled, so in frame-to-model mode nextDepth is identical to lastDepth: alias it for
    // the tracking call instead of building the pyramid twice.
    float* saved[NUM_PYRS];
    const bool alias = !ctx->frame_to_frame_rgb;
    if (alias) {
      for (int i = 0; i < NUM_PYRS; ++i) {
        saved[i] = od.nextDepth[i];
        od.nextDepth[i] = od.lastDepth[i];
      }
    } else {
      RC(odom_populate(ctx, 0, t.rgba, od.nextDepth, od.nextImage, true, nullptr, nullptr, false, false));
    }
    ef_stage(ctx, 3);
    int rc = odom_track_async(ctx, 0, ctx->rgb_only, ctx->icp_weight, ctx->pyramid, ctx->fast_odom, ctx->so3);
    if (alias)
      for (int i = 0; i < NUM_PYRS; ++i) od.nextDepth[i] = saved[i];
    RC(rc);
    RC(odom_finish_async(ctx, 0, weight_multiplier, true));
  } else {
    CU(cudaStreamSynchronize(ctx->stream));
    memcpy((char*)ctx->pin_small + 4096, in_T_wc, sizeof(double) * 16);
    CU(cudaMemcpyAsync((char*)ctx->dev_small + 4096, (char*)ctx->pin_small + 4096, sizeof(double) * 16, cudaMemcpyHostToDevice, ctx->stream));
    ctx->so3_ready = false;  // no tracking for this frame: the SO(3) result of its input side is not used
    RC(odom_set_pose_async(ctx, 0, (const double*)((char*)ctx->dev_small + 4096)));
    RC(odom_finish_async(ctx, 0, weight_multiplier, false));
  }
  // (k_gn_finish also refreshed the map kernels' float pose + inverse from the new T_wc)
  ef_stage(ctx, 6);
  // ElasticFusion.cpp:387: only loop closure reads this prediction
  if (!ctx->cfg.skip_mid_predict || ctx->cfg.close_loops) RC(predict_async(ctx));
  if (ctx->cfg.close_loops && !ctx->rgb_only) RC(local_loop_async(ctx));
  return 0;
}

// Second half of processFrame (ElasticFusion.cpp:536-607): index map, fuse, index map, clean (with the deformation graph stored by
// ef_set_deformation_graph when n_nodes > 0), predict, tick++.
static int frame_end_device(EfContext* ctx, int n_nodes, bool fern_accepted) {
  if (!ctx->frame_open) return EF_ESTATE;
  if (ctx->tick > 1 && !ctx->rgb_only) {
    RC(map_predict_indices_async(ctx, ctx->tick, ctx->max_depth_processed, ctx->cfg.time_delta, 1));
    ef_stage(ctx, 7);
    RC(map_fuse_async(ctx, ctx->tick, ctx->max_depth_processed, -1.0f));
    ef_stage(ctx, 8);
    RC(map_predict_indices_async(ctx, ctx->tick, ctx->max_depth_processed, ctx->cfg.time_delta, 2));
    ef_stage(ctx, 9);
    if (n_nodes > 0 && !fern_accepted)  // ElasticFusion.cpp:559-569: the time-stamp refresh of deformed surfels reads this depth
      RC(map_raycast_async(ctx, ctx->max_depth_processed, ctx->confidence, ctx->tick, ctx->tick - ctx->cfg.time_delta, 65535, 2, false));
    RC(map_clean_async(ctx, ctx->tick, ctx->confidence, ctx->cfg.time_delta, ctx->max_depth_processed, n_nodes, fern_accepted));
    ef_stage(ctx, 10);
  }
  if (ctx->tick == 1) RC(map_update_pose_async(ctx, nullptr));  // later frames: done by k_gn_finish, pose unchanged since
  RC(predict_async(ctx));  // ElasticFusion.cpp:599
  ef_stage(ctx, 11);
  ctx->tick++;
  ctx->frame_open = false;
  return 0;
}

extern "C" int ef_process_frame_device(EfContext* ctx, const uint8_t* rgb_dev, const uint16_t* depth_dev, int64_t timestamp,
                                       float weight_multiplier, const double* in_T_wc) {
  (void)timestamp;
  if (!ctx || ((rgb_dev == nullptr) != (depth_dev == nullptr))) return EF_EINVAL;
  RC(frame_begin_device(ctx, rgb_dev, depth_dev, weight_multiplier, in_T_wc));
  return frame_end_device(ctx, 0, false);
}

// processFrame split at the point where the reference hands control to its CPU deformation solver (ElasticFusion.cpp:505-526):
// begin = everything up to and including the local loop closure front half, end = fuse / clean / predict. Between the two
// the host may read ef_local_loop_result, run Deformation::constrain (unchanged reference code) and hand its output back:
// T_wc_override (T_wc_curr = T_wc_est) and the graph nodes applied inside clean.
extern "C" int ef_process_frame_begin(EfContext* ctx, const uint8_t* rgb, const uint16_t* depth, int64_t timestamp, float weight_multiplier,
                                      const double* in_T_wc) {
  (void)timestamp;
  if (!ctx || !rgb || !depth) return EF_EINVAL;
  if (ctx->la.pending || ctx->frame_open) return EF_ESTATE;
  const size_t n = (size_t)ctx->cfg.width * ctx->cfg.height;
  CU(cudaStreamSynchronize(ctx->stream));
  memcpy(ctx->pin_rgb, rgb, n * 3);
  memcpy(ctx->pin_depth, depth, n * 2);
  CU(cudaMemcpyAsync(ctx->tex.rgb, ctx->pin_rgb, n * 3, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(ctx->tex.depth_raw, ctx->pin_depth, n * 2, cudaMemcpyHostToDevice, ctx->stream));
  RC(frame_begin_device(ctx, ctx->tex.rgb, ctx->tex.depth_raw, weight_multiplier, in_T_wc));
  return ef_finish_frame(ctx);  // pose (and the loop-closure result) are final on return
}

extern "C" int ef_process_frame_end(EfContext* ctx, const double* T_wc_override, const float* graph_nodes16, int32_t n_nodes, int32_t fern_accepted) {
  if (!ctx || n_nodes < 0 || (n_nodes > 0 && !graph_nodes16)) return EF_EINVAL;
  if (!ctx->frame_open) return EF_ESTATE;
  if (T_wc_override) {
    RC(ef_set_pose(ctx, T_wc_override));
    RC(map_update_pose_async(ctx, nullptr));
  }
  RC(map_set_graph(ctx, graph_nodes16, n_nodes));
  RC(frame_end_device(ctx, n_nodes, fern_accepted != 0));
  return ef_finish_frame(ctx);
}

extern "C" int ef_local_loop_result(EfContext* ctx, EfLoopResult* out, double* src3, double* dst3, int32_t* times, int32_t max_constraints,
                                    int32_t* n_out) {
  if (!ctx || !out || max_constraints < 0) return EF_EINVAL;
  LoopDev* L = ctx->map.loop;
  struct Head {
    int ran, accepted, n_constraints;
    float lastICPError, lastICPCount;
    double cov_diag[6];
    double T_wc_est[16];
  } h;
  static_assert(offsetof(LoopDev, src) >= sizeof(Head), "LoopDev header layout");
  CU(cudaMemcpyAsync(&h, L, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  out->ran = h.ran;
  out->accepted = h.accepted;
  out->n_constraints = h.n_constraints;
  out->lastICPError = h.lastICPError;
  out->lastICPCount = h.lastICPCount;
  memcpy(out->cov_diag, h.cov_diag, sizeof(h.cov_diag));
  memcpy(out->T_wc_est, h.T_wc_est, sizeof(h.T_wc_est));
  int n = h.n_constraints < max_constraints ? h.n_constraints : max_constraints;
  if (n_out) *n_out = n;
  if (n > 0) {
    if (src3) CU(cudaMemcpyAsync(src3, (char*)L + offsetof(LoopDev, src), sizeof(double) * 3 * n, cudaMemcpyDeviceToHost, ctx->stream));
    if (dst3) CU(cudaMemcpyAsync(dst3, (char*)L + offsetof(LoopDev, dst), sizeof(double) * 3 * n, cudaMemcpyDeviceToHost, ctx->stream));
    if (times) CU(cudaMemcpyAsync(times, (char*)L + offsetof(LoopDev, times), sizeof(int) * n, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
  }
  return 0;
}

// Makes the main stream wait for the side stream's staged frame (a no-op without a pending frame): after this call every
// kernel enqueued so far, on either stream, precedes whatever the caller records on ef_stream() next.
extern "C" int ef_join_lookahead(EfContext* ctx) {
  if (!ctx) return EF_EINVAL;
  if (ctx->la.pending) CU(cudaStreamWaitEvent(ctx->stream, ctx->la.ready, 0));
  return 0;
}

// Completes the frame enqueued by ef_process_frame_device: pose and surfel count are read back and final on return.
extern "C" int ef_finish_frame(EfContext* ctx) {
  if (!ctx) return EF_EINVAL;
  char* s = (char*)ctx->pin_small + 8192;
  CU(cudaMemcpyAsync(s, (char*)ctx->odom[0].gn + offsetof(GNState, T_wc), sizeof(double) * 16, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaMemcpyAsync(s + 128, ctx->map.count, 4, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  memcpy(ctx->T_wc, s, sizeof(double) * 16);
  ctx->host_count = *(int*)(s + 128);
  return 0;
}

extern "C" int ef_process_frame(EfContext* ctx, const uint8_t* rgb, const uint16_t* depth, int64_t timestamp, float weight_multiplier,
                                const double* in_T_wc) {
  if (!ctx || ((rgb == nullptr) != (depth == nullptr))) return EF_EINVAL;
  if (!rgb) {
    RC(ef_process_frame_device(ctx, nullptr, nullptr, timestamp, weight_multiplier, in_T_wc));  // prefetched frame
    return ef_finish_frame(ctx);
  }
  if (ctx->la.pending) return EF_ESTATE;  // a prefetched frame must be consumed first; nothing has been touched yet
  const size_t n = (size_t)ctx->cfg.width * ctx->cfg.height;
  // the staging buffers may still be in flight from the previous frame
  CU(cudaStreamSynchronize(ctx->stream));
  memcpy(ctx->pin_rgb, rgb, n * 3);
  memcpy(ctx->pin_depth, depth, n * 2);
  CU(cudaMemcpyAsync(ctx->tex.rgb, ctx->pin_rgb, n * 3, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(ctx->tex.depth_raw, ctx->pin_depth, n * 2, cudaMemcpyHostToDevice, ctx->stream));
  RC(ef_process_frame_device(ctx, ctx->tex.rgb, ctx->tex.depth_raw, timestamp, weight_multiplier, in_T_wc));
  // results the caller can observe (get_T_wc, counts) are final on return
  return ef_finish_frame(ctx);
}

// debug exports (phase profiling builds)
// EF_STAGE_TIMING=1: milliseconds between consecutive stage events of the last frame (0 start, 1 upload + RGBA + bilateral/metric,
// 2 live pyramids + SO(3) loop, 3 model pyramids, 4 sobel + candidates + begin, 5 Gauss-Newton loop, 6 finish, 7 index map, 8 fuse,
// 9 index map, 10 clean, 11 predict); returns the event count. Stages 2..6 are what the reference does inside
// RGBDOdometry::init* + getIncrementalTransformation.
extern "C" int ef_debug_stage_ms(EfContext* ctx, float* out) {
  if (!ctx || !ctx->stage_timing) return 0;
  cudaStreamSynchronize(ctx->stream);
  int prev = -1;
  for (int i = 0; i < 16; ++i) {
    out[i] = 0.f;
    if (!(ctx->stage_n >> i & 1)) continue;
    if (prev >= 0) cudaEventElapsedTime(&out[i], ctx->stage_ev[prev], ctx->stage_ev[i]);
    prev = i;
  }
  return 12;
}
extern "C" void* ef_debug_gn(EfContext* ctx, int which) { return ctx ? (void*)ctx->odom[which].gn : nullptr; }
extern "C" int ef_debug_gn_size() { return (int)sizeof(GNState); }
extern "C" int ef_debug_dbg_offset() { return (int)offsetof(GNState, dbg); }
