/* efusion_b200.h — C ABI of libefusion.so, the B200-native (sm_100a) implementation of ElasticFusion's per-frame
 * tracking + surfel fuse/predict hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types. Each entry point
 * names the reference interface it replaces (paths relative to the reference tree). The C++ classes in
 * include/efusion/ (ElasticFusion, RGBDOdometry, GlobalModel, IndexMap) are thin wrappers over these calls.
 *
 * Conventions
 *  - every function returns 0 on success, a cudaError_t value (>0) on a CUDA failure, or a negative EF_E* code.
 *    (The reference prints and exit(0)s on CUDA errors, Core/Cuda/convenience.cuh:64-70; the C++ wrappers keep
 *    that policy, the C ABI reports instead.)
 *  - one EfContext == one device + one CUDA stream; contexts on different devices may be driven from different
 *    host threads concurrently. Host buffers are borrowed only for the duration of a call.
 *  - poses are row-major double[16] camera-to-world (Sophus::SE3d::matrix()); 3x3 matrices row-major float[9].
 *  - all device work is asynchronous on the context's stream; calls that return results to host memory
 *    synchronise that stream before returning.
 */
#ifndef EFUSION_B200_H_
#define EFUSION_B200_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EF_OK 0
#define EF_EINVAL (-1)
#define EF_ENOMEM (-2)
#define EF_ESTATE (-3)

typedef struct EfContext EfContext;

/* Constructor arguments of ElasticFusion (Core/ElasticFusion.h:42-58) plus what the reference keeps in the
 * Resolution / Intrinsics singletons (Core/Utils/Resolution.h, Intrinsics.h) and a runtime surfel capacity
 * (the reference hard-codes 3072^2, Core/GlobalModel.cpp:22-24). */
typedef struct {
  int32_t width, height;
  float fx, fy, cx, cy;
  int32_t time_delta;     /* 200 */
  int32_t count_thresh;   /* 35000  (loop closure only; kept for API parity) */
  float err_thresh;       /* 5e-05  (loop closure only) */
  float cov_thresh;       /* 1e-05  (loop closure only) */
  int32_t close_loops;    /* 1: run the local loop closure front half every frame (results: ef_local_loop_result); the
                             deformation solve and Ferns stay with the host */
  int32_t iclnuim;
  int32_t reloc;          /* must be 0 */
  float photo_thresh;     /* 115 (ferns only) */
  float confidence;       /* 10 */
  float depth_cutoff;     /* 3 */
  float icp_weight;       /* 10 */
  int32_t fast_odom;      /* 0 */
  float fern_thresh;      /* 0.3095 (ferns only) */
  int32_t so3;            /* 1 */
  int32_t frame_to_frame_rgb; /* 0 */
  int32_t capacity;       /* max surfels resident in HBM (48 B each) */
  int32_t device;         /* CUDA device ordinal */
  int32_t skip_mid_predict; /* 1: skip the predict() at Core/ElasticFusion.cpp:387 whose outputs only loop closure reads */
} EfConfig;

void ef_default_config(EfConfig* cfg, int width, int height, float fx, float fy, float cx, float cy);

/* ElasticFusion::ElasticFusion / ~ElasticFusion (Core/ElasticFusion.cpp:22-163). `stream` is a cudaStream_t
 * (NULL: the context creates its own non-blocking stream). */
int ef_create(const EfConfig* cfg, void* stream, EfContext** out);
int ef_destroy(EfContext* ctx);
void* ef_stream(EfContext* ctx);
const char* ef_error_string(int code);

/* ---- whole frame: ElasticFusion::processFrame (Core/ElasticFusion.cpp:270-607) ------------------------------- */
/* rgb: uint8 RGB row-major W*H*3, depth: uint16 millimetres W*H (HOST memory, consumed before return);
 * in_T_wc may be NULL (track) or a pose to use instead of tracking. Returns after get_T_wc-observable state is final. */
int ef_process_frame(EfContext* ctx, const uint8_t* rgb, const uint16_t* depth, int64_t timestamp,
                     float weight_multiplier, const double* in_T_wc);
/* same with inputs already resident in HBM; fully asynchronous (no host sync) — call ef_sync before reading results */
int ef_process_frame_device(EfContext* ctx, const uint8_t* rgb_dev, const uint16_t* depth_dev, int64_t timestamp,
                            float weight_multiplier, const double* in_T_wc);
int ef_sync(EfContext* ctx);
/* Frame look-ahead (no reference counterpart: the reference uploads and filters a frame inside processFrame, each GL
 * pass followed by glFinish — Core/ElasticFusion.cpp:278-285). Everything of the NEXT frame that depends neither on the
 * map nor on the pose (upload, RGBA expansion, bilateral filter + metric depth, depth pyramid + vertex/normal maps,
 * intensity pyramid) is enqueued on a side stream into a spare buffer set, so it overlaps the Gauss-Newton loop and the
 * map update of the frame in flight. One frame may be pending; it is consumed by passing rgb = depth = NULL to
 * ef_process_frame / ef_process_frame_device (EF_ESTATE if none is pending, or if a frame is passed while one is).
 * Results are identical to the non-prefetched call. Typical loop:
 *   ef_prefetch_frame(f0); for (i...) { ef_process_frame_device(NULL, NULL, ts[i], ...); ef_prefetch_frame(f[i+1]);
 *   ef_finish_frame(); }                                                                                          */
int ef_prefetch_frame(EfContext* ctx, const uint8_t* rgb, const uint16_t* depth);                /* HOST buffers   */
int ef_prefetch_frame_device(EfContext* ctx, const uint8_t* rgb_dev, const uint16_t* depth_dev); /* DEVICE buffers */
/* orders the main stream after the side stream's staged frame (so an event recorded on ef_stream() afterwards covers the
 * look-ahead work too); no-op when nothing is staged */
int ef_join_lookahead(EfContext* ctx);
/* waits for the frame enqueued by ef_process_frame_device and refreshes the host mirrors (pose, surfel count) */
int ef_finish_frame(EfContext* ctx);
/* processFrame split where the reference hands control to its CPU deformation solver (Core/ElasticFusion.cpp:505-526).
 * ef_process_frame_begin: upload, filter, track, velocity weighting, mid-frame predict and -- with cfg.close_loops -- the local
 * loop closure FRONT HALF (:447-505: INACTIVE prediction, modelToModel registration, acceptance test, constraint sampling);
 * results are final on return. ef_process_frame_end: optional pose override (T_wc_curr = T_wc_est, :524) and deformation graph
 * (16 floats per node: position 3, rotation 9 column-major, translation 3, time -- Core/Deformation.cpp:175-189) applied inside
 * clean (Core/Shaders/copy_unstable.vert:132-322), then index map / fuse / index map / clean / predict, tick++.
 * ef_process_frame == begin + end(NULL, NULL, 0, 0). EF_ESTATE if the calls are not paired. */
int ef_process_frame_begin(EfContext* ctx, const uint8_t* rgb, const uint16_t* depth, int64_t timestamp, float weight_multiplier,
                           const double* in_T_wc);
int ef_process_frame_end(EfContext* ctx, const double* T_wc_override, const float* graph_nodes16, int32_t n_nodes,
                         int32_t fern_accepted);
/* result of the last frame's local loop closure front half (cfg.close_loops = 1); src3 / dst3: vert_w_curr / vert_w_est of
 * each constraint (what localDeformation.addConstraint receives, ElasticFusion.cpp:493-503), times: the INACTIVE view's stamp */
typedef struct {
  int32_t ran;            /* 0: first frame, rgbOnly, or close_loops off */
  int32_t accepted;       /* covariance diagonal <= cov_thresh && lastICPCount > count_thresh && lastICPError < err_thresh */
  int32_t n_constraints;
  float lastICPError, lastICPCount;
  double cov_diag[6];
  double T_wc_est[16];
} EfLoopResult;
int ef_local_loop_result(EfContext* ctx, EfLoopResult* out, double* src3, double* dst3, int32_t* times, int32_t max_constraints,
                         int32_t* n_out);
/* ElasticFusion::predict (Core/ElasticFusion.cpp:621-653) */
int ef_predict(EfContext* ctx);

/* getters/setters of ElasticFusion (Core/ElasticFusion.h:120-213) */
int ef_get_pose(EfContext* ctx, double* T_wc16);             /* get_T_wc */
int ef_set_pose(EfContext* ctx, const double* T_wc16);
int ef_get_tick(EfContext* ctx, int32_t* tick);              /* getTick */
int ef_set_tick(EfContext* ctx, int32_t tick);               /* setTick */
int ef_set_rgb_only(EfContext* ctx, int32_t v);              /* setRgbOnly */
int ef_set_icp_weight(EfContext* ctx, float v);              /* setIcpWeight */
int ef_set_pyramid(EfContext* ctx, int32_t v);               /* setPyramid */
int ef_set_fast_odom(EfContext* ctx, int32_t v);             /* setFastOdom */
int ef_set_so3(EfContext* ctx, int32_t v);                   /* setSo3 */
int ef_set_frame_to_frame_rgb(EfContext* ctx, int32_t v);    /* setFrameToFrameRGB */
int ef_set_confidence_threshold(EfContext* ctx, float v);    /* setConfidenceThreshold */
int ef_set_depth_cutoff(EfContext* ctx, float v);            /* setDepthCutoff */

/* ---- tracker: RGBDOdometry (Core/Utils/RGBDOdometry.h:31-79) and the free functions of
 *      Core/Cuda/cudafuncs.cuh:61-169 it drives -------------------------------------------------------------- */
typedef struct {
  int32_t kind;  /* 0 SE3 Gauss-Newton iteration, 1 SO3 pre-alignment iteration */
  int32_t level, iter;
  int32_t rgb_count, rgb_sigma;
  float sigma_val;
  float A_icp[36], b_icp[6], icp_residual[2];
  float A_rgb[36], b_rgb[6];
  float A_so3[9], b_so3[3], so3_residual[2];
  double lastA[36], lastb[6], result[6];
} EfSolveTrace;

/* public result fields of RGBDOdometry (RGBDOdometry.h:71-79) */
typedef struct {
  float lastICPError, lastICPCount, lastRGBError, lastRGBCount, lastSO3Error, lastSO3Count;
  double lastA[36], lastb[6];
} EfOdomStats;

/* which tracker instance: 0 = frameToModel (the only one driven by ef_process_frame), 1 = modelToModel */
/* RGBDOdometry::initICP(GPUTexture* filteredDepth, depthCutoff) — RGBDOdometry.cpp:121-147. depth: DEVICE u16 */
int ef_odom_init_icp_depth(EfContext* ctx, int which, const uint16_t* depth_dev, float depth_cutoff);
/* RGBDOdometry::initICP(predictedVertices, predictedNormals) — RGBDOdometry.cpp:149-169. DEVICE float4 maps */
int ef_odom_init_icp_pred(EfContext* ctx, int which, const float* vtx4_dev, const float* nrm4_dev);
/* RGBDOdometry::initICPModel — RGBDOdometry.cpp:171-210 */
int ef_odom_init_icp_model(EfContext* ctx, int which, const float* vtx4_dev, const float* nrm4_dev,
                           const double* T_wc16);
/* RGBDOdometry::initRGB / initRGBModel / initFirstRGB — RGBDOdometry.cpp:212-257. DEVICE RGBA8 image */
int ef_odom_init_rgb(EfContext* ctx, int which, const uint8_t* rgba_dev);
int ef_odom_init_rgb_model(EfContext* ctx, int which, const uint8_t* rgba_dev);
int ef_odom_init_first_rgb(EfContext* ctx, int which, const uint8_t* rgba_dev);
/* RGBDOdometry::getIncrementalTransformation — RGBDOdometry.cpp:259-571. The whole SO3 + 3-level Gauss-Newton loop
 * runs on the device with no host round trip; trace (HOST, may be NULL) receives one record per iteration. */
int ef_odom_track(EfContext* ctx, int which, double* T_wc16, int32_t rgb_only, float icp_weight, int32_t pyramid,
                  int32_t fast_odom, int32_t so3, EfSolveTrace* trace, int32_t max_trace, int32_t* n_trace);
int ef_odom_stats(EfContext* ctx, int which, EfOdomStats* out);
/* RGBDOdometry::getCovariance — RGBDOdometry.cpp:573-575 */
int ef_odom_covariance(EfContext* ctx, int which, double* cov36);

/* single reduction steps on the tracker's current pyramids, with explicit poses and HOST results, exactly the
 * argument meaning of the reference's free functions (used by the parity tests and the roofline bench):
 * icpStep (Core/Cuda/reduce.cu:333-401), computeRgbResidual (:723-787), rgbStep (:502-550), so3Step (:919-973) */
int ef_icp_step(EfContext* ctx, int which, int level, const float* Rcurr9, const float* tcurr3,
                const float* Rprev_inv9, const float* tprev3, float* A36, float* b6, float* residual2);
int ef_rgb_residual(EfContext* ctx, int which, int level, const float* krkinv9, const float* kt3,
                    int32_t* sigma_sum, int32_t* count);
int ef_rgb_step(EfContext* ctx, int which, int level, float sigma, float* A36, float* b6);
int ef_so3_step(EfContext* ctx, int which, const float* image_basis9, const float* kinv9, const float* krlr9,
                float* A9, float* b3, float* residual2);
/* asynchronous variant for benchmarking the ICP reduction: result stays on the device. Poses may be NULL (reuse the last ones).
 * ef_icp_dense_pass_async launches only the dense residual+Jacobian+per-CTA reduction kernel (k_iter1), without the 1-CTA
 * final sum — the launch bench.py's roofline object times. */
int ef_icp_step_async(EfContext* ctx, int which, int level, const float* Rcurr9, const float* tcurr3,
                      const float* Rprev_inv9, const float* tprev3);
int ef_icp_dense_pass_async(EfContext* ctx, int which, int level);

/* ---- depth preprocess: ElasticFusion::filterDepth + metriciseDepth (Core/ElasticFusion.cpp:655-673,
 *      Core/Shaders/depth_bilateral.frag, depth_metric.frag). DEVICE in/out, any out may be NULL ------------- */
int ef_preprocess_depth(EfContext* ctx, const uint16_t* depth_raw_dev, float depth_cutoff, uint16_t* filtered_dev,
                        float* metric_dev, float* metric_filtered_dev);

/* ---- surfel map: GlobalModel (Core/GlobalModel.h:34-89), IndexMap (Core/IndexMap.h:33-142),
 *      FillIn (Core/Shaders/FillIn.h), Resize (Core/Shaders/Resize.h) -------------------------------------- */
/* GlobalModel::initialise + FeedbackBuffer::compute (GlobalModel.cpp:229-284, FeedbackBuffer.cpp:81-138): builds the
 * first-frame map from the context's current RGB / metric depth textures */
int ef_map_initialise(EfContext* ctx);
/* IndexMap::predictIndices — IndexMap.cpp:190-258 */
int ef_map_predict_indices(EfContext* ctx, const double* T_wc16, int32_t time, float max_depth, int32_t time_delta);
/* GlobalModel::fuse — GlobalModel.cpp:356-525 (uses the context's RGB, metric depth and index-map textures) */
int ef_map_fuse(EfContext* ctx, const double* T_wc16, int32_t time, float max_depth, float weighting);
/* GlobalModel::clean — GlobalModel.cpp:527-671 (no deformation graph) */
int ef_map_clean(EfContext* ctx, const double* T_wc16, int32_t time, float conf_threshold, int32_t time_delta,
                 float max_depth);
/* GlobalModel::clean with a deformation graph (GlobalModel.cpp:527-671, graph.size() > 0; copy_unstable.vert:132-322); the
 * time-stamp refresh reads EF_BUF_SYNTH_DEPTH (ef_map_raycast mode 2). graph_nodes16: HOST, 16 floats per node. */
int ef_map_clean_deform(EfContext* ctx, const double* T_wc16, int32_t time, float conf_threshold, int32_t time_delta,
                        float max_depth, const float* graph_nodes16, int32_t n_nodes, int32_t is_fern);
/* IndexMap::combinedPredict (mode 0 ACTIVE, 1 INACTIVE) / synthesizeDepth (mode 2) — IndexMap.cpp:293-476 */
int ef_map_raycast(EfContext* ctx, const double* T_wc16, float max_depth, float conf_threshold, int32_t time,
                   int32_t max_time, int32_t time_delta, int32_t mode);
/* FillIn::vertex/normal/image — FillIn.cpp:62-191 */
int ef_map_fill_in(EfContext* ctx, int32_t passthrough_geometry, int32_t passthrough_image);
/* Resize::image + ElasticFusion::denseEnough — Resize.cpp:50-79, ElasticFusion.cpp:256-268 */
int ef_dense_enough(EfContext* ctx, int32_t* out);
/* GlobalModel::lastCount / downloadMap (GlobalModel.cpp:673-706): out = count*12 floats, reference Vertex layout
 * (Core/Shaders/Vertex.cpp:22-41) */
int ef_map_count(EfContext* ctx, int32_t* count);
int ef_map_download(EfContext* ctx, float* out12, int32_t max_surfels, int32_t* count);
int ef_map_upload(EfContext* ctx, const float* in12, int32_t count);
/* overwrites surfels [first, first+count) of the resident map in place (count unchanged; EF_EINVAL if the range leaves it).
 * No reference counterpart (the reference never edits its VBO from the host); used to build large maps piecewise. */
int ef_map_upload_range(EfContext* ctx, const float* in12, int32_t first, int32_t count);
/* unstable surfels appended by the last ef_map_fuse (the reference's newUnstableVbo) */
int ef_map_download_new(EfContext* ctx, float* out12, int32_t max_surfels, int32_t* count);

/* ---- named device buffers (the reference's GPUTexture / DeviceArray handles) ------------------------------ */
enum {
  /* input / preprocess textures (ElasticFusion::textures, GPUTexture.cpp:22-27) */
  EF_BUF_RGB = 0,                 /* u8 x3  */
  EF_BUF_DEPTH_RAW = 1,           /* u16    */
  EF_BUF_DEPTH_FILTERED = 2,      /* u16    */
  EF_BUF_DEPTH_METRIC = 3,        /* f32    */
  EF_BUF_DEPTH_METRIC_FILTERED = 4,
  EF_BUF_RGBA = 5,                /* u8 x4: the RGB texture as CUDA sees it */
  /* IndexMap attachments (IndexMap.h:74-142) */
  EF_BUF_INDEX = 10,              /* u32    */
  EF_BUF_VERT_CONF = 11,          /* f32 x4 */
  EF_BUF_COLOR_TIME = 12,
  EF_BUF_NORM_RAD = 13,
  EF_BUF_IMAGE = 14,              /* u8 x4  */
  EF_BUF_VERTEX = 15,             /* f32 x4 */
  EF_BUF_NORMAL = 16,
  EF_BUF_TIME = 17,               /* u16    */
  EF_BUF_OLD_IMAGE = 18,
  EF_BUF_OLD_VERTEX = 19,
  EF_BUF_OLD_NORMAL = 20,
  EF_BUF_OLD_TIME = 21,
  EF_BUF_SYNTH_DEPTH = 22,        /* f32    */
  /* FillIn textures */
  EF_BUF_FILL_IMAGE = 30,
  EF_BUF_FILL_VERTEX = 31,
  EF_BUF_FILL_NORMAL = 32,
  /* RGBDOdometry pyramids: id + 100*which, `level` selects the pyramid level */
  EF_BUF_VMAP_CURR = 40,          /* f32, 3 planes stacked ((3*rows) x cols) */
  EF_BUF_NMAP_CURR = 41,
  EF_BUF_VMAP_G_PREV = 42,
  EF_BUF_NMAP_G_PREV = 43,
  EF_BUF_LAST_DEPTH = 44,         /* f32 */
  EF_BUF_NEXT_DEPTH = 45,
  EF_BUF_LAST_IMAGE = 46,         /* u8 */
  EF_BUF_NEXT_IMAGE = 47,
  EF_BUF_LAST_NEXT_IMAGE = 48,
  EF_BUF_DIDX = 49,               /* i16 */
  EF_BUF_DIDY = 50,
  EF_BUF_DEPTH_TMP = 51,          /* u16 */
  EF_BUF_CORRES = 52,             /* DataTerm 16 B */
  EF_BUF_VMAPS_TMP = 53           /* f32 x4, level 0 only */
};
/* device pointer + byte size of a named buffer (id + 100*which for tracker buffers) */
int ef_buffer(EfContext* ctx, int32_t id, int32_t level, void** dev_ptr, size_t* bytes);
int ef_upload(EfContext* ctx, int32_t id, int32_t level, const void* host, size_t bytes);
int ef_download(EfContext* ctx, int32_t id, int32_t level, void* host, size_t bytes);
/* Resize::image / vertex / time (Core/Shaders/Resize.cpp:50-159): the named full-resolution attachment (RGBA8, RGBA32F or
 * R16UI) sampled on the (W/factor) x (H/factor) grid of texel centres with nearest filtering, into HOST memory, tightly packed. */
int ef_resize(EfContext* ctx, int32_t id, int32_t factor, void* host_out, size_t bytes);
/* number of kernels launched by this context since creation (bench.py's gpu_launches) */
int ef_launch_count(EfContext* ctx, int64_t* n);
/* EF_STAGE_TIMING=1 in the environment at ef_create: milliseconds between the stage events of the last frame, out[16]:
 * [1] upload + RGBA + bilateral/metric, [2] live pyramids + SO(3), [3] model pyramids, [4] sobel + candidates, [5] Gauss-Newton
 * loop, [6] finish, [7] index map, [8] fuse, [9] index map, [10] clean, [11] predict. Returns the number of slots (0 if off). */
int ef_debug_stage_ms(EfContext* ctx, float* out16);

#ifdef __cplusplus
}
#endif
#endif /* EFUSION_B200_H_ */
