/* emvs_b200.h — C-ABI of the B200-native DSI ray-voting engine.
 *
 * This is the drop-in boundary for the mapping hot path of tub-rip/dvs_mcemvs.  The
 * reference has no FFI layer: its boundary is the C++ class API of libcartesian3dgrid
 * (class Grid3D) and libmapper_emvs_stereo (class EMVS::MapperEMVS).  The header-only host
 * classes in dvs_mcemvs_b200/host/ keep those class names and signatures and forward to the
 * entry points below; every entry point cites the reference interface it replaces
 * (paths relative to the reference root):
 *   MAP = mapper_emvs_stereo/src/mapper_emvs_stereo.cpp
 *   MHP = mapper_emvs_stereo/include/mapper_emvs_stereo/mapper_emvs_stereo.hpp
 *   G3H = cartesian3dgrid/include/cartesian3dgrid/cartesian3dgrid.h
 *   G3C = cartesian3dgrid/src/cartesian3dgrid.cpp
 *   DV  = mapper_emvs_stereo/include/mapper_emvs_stereo/depth_vector.hpp
 *   TRJ = mapper_emvs_stereo/include/mapper_emvs_stereo/trajectory.hpp
 *   P1/P2 = mapper_emvs_stereo/src/process1.cpp / process2.cpp
 *
 * Conventions: plain pointers and sizes only; every function returns an int status
 * (EMVS_OK == 0) and never throws or aborts across the boundary; emvs_last_error() gives the
 * message of the last failure on the calling thread.  There is NO CPU fallback: every compute
 * entry point needs a CUDA device of compute capability 10.x and fails with EMVS_ERR_CUDA
 * otherwise.  Calls on one context are serialised by the caller (the reference's evaluateDSI
 * is not re-entrant either, MAP:79,83); different contexts may be driven from different
 * threads.  Unless a function says otherwise, pointer arguments are HOST pointers that are
 * only borrowed for the duration of the call.  Execution model: every context is one in-order
 * pipeline (a CUDA stream).  Calls that only produce DEVICE state (build, evaluate_dsi, grid ops,
 * reset, copy, allreduce_async) return as soon as their host inputs have been consumed; calls
 * that produce HOST output (download, collapse, counts, mean_square, timers) return after the
 * output is complete, which implies everything issued before them.  emvs_context_sync waits
 * for the whole pipeline (use it before reading a wall clock).
 */
#ifndef EMVS_B200_H_
#define EMVS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EMVS_ABI_VERSION 2

#if defined(__GNUC__)
#define EMVS_API __attribute__((visibility("default")))
#else
#define EMVS_API
#endif

/* ---- status codes -------------------------------------------------------------------- */
enum {
  EMVS_OK = 0,
  EMVS_ERR_INVALID = 1,     /* bad argument (null pointer, shape mismatch, bad id)          */
  EMVS_ERR_CUDA = 2,        /* CUDA runtime error or no usable device                       */
  EMVS_ERR_TOO_FEW = 3,     /* fewer than EMVS_PACKET_SIZE events: evaluateDSI -> false (MAP:71-75) */
  EMVS_ERR_NCCL = 4,        /* NCCL unavailable or a collective failed                      */
  EMVS_ERR_STATE = 5        /* call sequence error (e.g. LUT not set, comm not initialised) */
};

#define EMVS_PACKET_SIZE 1024 /* MHP:153 packet_size_ */

/* ---- POD types ------------------------------------------------------------------------ */

/* dvs_msgs/Event {uint16 x; uint16 y; ros::Time ts; bool polarity} as roscpp lays it out
 * (16 bytes).  Only x, y are read on the device; ts is read by the host packet stage. */
typedef struct emvs_event {
  uint16_t x, y;
  uint32_t sec, nsec;
  uint8_t polarity;
  uint8_t pad_[3];
} emvs_event;

/* Structure-of-arrays event list, as the HDF5 event files of DSEC / TUM-VIE store them (events/x, events/y, events/t)
 * and as a caller can cheaply keep them: only x and y travel to the device (4 bytes per event instead of the 16-byte
 * dvs_msgs::Event), the timestamps stay on the host, where the packet stage reads ONE of them per 1024 events
 * (MAP:91).  t_ns = ros::Time::toNSec() of the event, non-decreasing like the reference's sorted event vector. */
typedef struct emvs_events_soa {
  const uint16_t* x;
  const uint16_t* y;
  const int64_t* t_ns;
  size_t n;
} emvs_events_soa;

/* kindr::minimal::QuatTransformation: unit quaternion (w,x,y,z) + position, double. */
typedef struct emvs_pose {
  double q[4];
  double t[3];
} emvs_pose;

/* One control pose of a LinearTrajectory (std::map<ros::Time, Transformation>, TRJ:11). */
typedef struct emvs_stamped_pose {
  uint32_t sec, nsec;
  emvs_pose T;
} emvs_stamped_pose;

/* Output of the packet stage of evaluateDSI (MAP:88-126): the pixel homography H_z0_px
 * (row-major) that maps a rectified event pixel onto plane Z0 of the reference view, the
 * camera centre C in the reference view, and the index of the packet's first event. */
typedef struct emvs_packet {
  float H[9];
  float C[3];
  uint64_t first_event;
} emvs_packet;

/* EMVS::ShapeDSI (MHP:40-65) + the quantities MapperEMVS::setupDSI derives (MAP:208-241). */
typedef struct emvs_shape {
  uint32_t dimX, dimY, dimZ;   /* 0 for dimX/dimY -> sensor size (MAP:216-217)              */
  float min_depth, max_depth;
  float fov_deg;               /* < 10 -> use the camera's fx (MAP:220-224)                 */
  int32_t inverse_depth;       /* 0: LinearDepthVector (build default), 1: InverseDepthVector */
} emvs_shape;

/* What the reference takes from image_geometry::PinholeCameraModel (MAP:34-48): full
 * resolution and the PROJECTION-matrix intrinsics fx(), fy(), cx(), cy(). */
typedef struct emvs_camera {
  uint32_t width, height;
  float fx, fy, cx, cy;
} emvs_camera;

/* Fusion ids as used by --stereo_fusion / --temporal_fusion (P1:136-158, docs/running.md). */
enum {
  EMVS_FUSE_MIN = 1, EMVS_FUSE_HM = 2, EMVS_FUSE_GM = 3,
  EMVS_FUSE_AM = 4, EMVS_FUSE_RMS = 5, EMVS_FUSE_MAX = 6
};

/* Pairwise / in-place voxel ops of Grid3D (G3H:64-192). */
enum {
  EMVS_OP_ADD = 0,            /* addTwoGrids            G3H:64-70    a += b                 */
  EMVS_OP_MIN = 1,            /* minTwoGrids            G3H:111-117                          */
  EMVS_OP_HM = 2,             /* harmonicMeanTwoGrids   G3H:119-127  2ab/(a+b+eps)          */
  EMVS_OP_GM = 3,             /* geometricMeanTwoGrids  G3H:150-156  sqrt(ab)               */
  EMVS_OP_AM = 4,             /* arithmeticMeanTwoGrids G3H:158-164  0.5(a+b)               */
  EMVS_OP_RMS = 5,            /* rmsTwoGrids            G3H:141-148  sqrt(0.5(a^2+b^2))     */
  EMVS_OP_MAX = 6,            /* maxTwoGrids            G3H:186-192                          */
  EMVS_OP_HM_N = 7,           /* harmonicMeanTwoGrids(g,n,eps) G3H:130-139                   */
  EMVS_OP_ADD_INV = 8,        /* addInverseOfTwoGrids   G3H:72-78    a += 1/(eps+b)         */
  EMVS_OP_HM_FROM_SUMINV = 9, /* computeHMfromSumOfInv  G3H:80-86    a = n/a   (b unused)   */
  EMVS_OP_AM_FROM_SUM = 10    /* computeAMfromSum       G3H:87-93    a = a/n   (b unused)   */
};

/* emvs_mapper_build flags */
enum {
  EMVS_BUILD_RESET = 0,       /* dsi_.resetGrid() then vote (MAP:145-146) — the reference behaviour */
  EMVS_BUILD_ACCUMULATE = 1,  /* vote on top of the current contents (sub-interval sharding)  */
  EMVS_BUILD_ALLREDUCE = 2,   /* multi-GPU: this build is one rank's shard; every Z-slab is summed over
                                 the ranks (ncclAllReduce) as soon as it is voted, overlapped with the
                                 votes of the next slab; the vote counts are summed too.  Needs
                                 emvs_comm_init; every rank must issue the same sequence of builds.      */
  EMVS_BUILD_PEER_REDUCE = 4  /* multi-GPU, fused exchange (emvs_exchange_begin ... emvs_exchange_fuse_collapse):
                                 every Z-slab, once merged, is announced to the peers and this rank's row band of
                                 it is summed over all ranks straight from their HBM (NVLink peer loads) into a
                                 local band buffer, on a side stream under the votes of the next slab.  May be
                                 combined with EMVS_BUILD_ACCUMULATE for the LAST of several builds into one DSI
                                 (its merged slabs are the ones announced as final).                              */
};

typedef struct emvs_context emvs_context;  /* one CUDA device + stream + scratch              */
typedef struct emvs_grid emvs_grid;        /* a device-resident DSI volume    == Grid3D        */
typedef struct emvs_mapper emvs_mapper;    /* camera + DSI shape + its grid   == MapperEMVS    */

/* ---- library / context ---------------------------------------------------------------- */
EMVS_API int emvs_abi_version(void);
EMVS_API const char* emvs_last_error(void);

/* Create a context on CUDA device `device`.  Fails (EMVS_ERR_CUDA) when no sm_10x device. */
EMVS_API int emvs_context_create(int device, emvs_context** out);
EMVS_API int emvs_context_destroy(emvs_context* ctx);
EMVS_API int emvs_context_sync(emvs_context* ctx);
/* Tuning: number of Z-planes voted per pass over the event list (0 = automatic). */
EMVS_API int emvs_context_set_slab(emvs_context* ctx, uint32_t planes_per_slab);
/* Tuning of emvs_mapper_evaluate_dsi on an idle pipeline: the first `percent` % of the event list is uploaded and
 * voted first while the rest is still crossing PCIe, then the rest is voted into the same DSI (votes add);
 * $EMVS_UPLOAD_PIECES = 3 / 4 cuts the rest again (pieces grow by 2.2 x; as complete builds measured slower).  Lists
 * shorter than `min_events` are built in one piece; percent = 0 disables.  Defaults: 15 %, 2^20 events
 * ($EMVS_UPLOAD_SPLIT overrides the percentage at context creation).
 * Builds that take the single multi-slab vote launch and are not part of a peer exchange use the deferred
 * form instead, with its own geometry ($EMVS_UPLOAD_DEFER_MERGE=1, _DEFER_SPLIT=6, _DEFER_PIECES=4): the pieces before the
 * last one only vote — their votes stay in the per-slab scratch —, the last one votes on top and merges once.  `percent`
 * = 0 disables that form too; `min_events` applies to both. */
EMVS_API int emvs_context_set_upload_split(emvs_context* ctx, uint32_t percent, uint64_t min_events);
/* Streaming callers (consecutive windows, main.cpp:177-431 full_seq loop): announce the event list of the NEXT
 * emvs_mapper_evaluate_dsi / emvs_mapper_build call on this context.  Its host->device copy starts now, on the copy
 * stream, under whatever the context is computing; that next call recognises the list by (pointer, n_events) and
 * skips its own upload.  Returns immediately.  Lifetime rule: the list must stay valid and unchanged until the
 * consuming call has returned or the prefetch has been cancelled.  One prefetch can be pending per context (a new
 * one replaces it); the next host-buffer evaluate / build call either consumes it (same list) or DROPS it (any
 * other list), so an announcement never outlives that call and can never be matched to an unrelated later list
 * that happens to be allocated at the same address.  (Builds from device-resident inputs leave it pending.) */
EMVS_API int emvs_context_prefetch_events(emvs_context* ctx, const emvs_event* events, size_t n_events);
/* Generation number of the pending prefetch (every prefetch call gets a new one), 0 when none is pending — lets a
 * caller assert that the announcement it made is the one about to be consumed. */
EMVS_API int emvs_context_prefetch_pending(emvs_context* ctx, uint64_t* generation);
/* Withdraws the pending prefetch, if any: waits for its copies, after which the caller owns its arrays again. */
EMVS_API int emvs_context_prefetch_cancel(emvs_context* ctx);
/* Device self-test of the vote kernel's shared-divisor division (a prepared reciprocal per (plane, packet) and three
 * FFMAs per numerator instead of a full IEEE division): runs about n_pairs pseudo-random and adversarial operand
 * pairs inside the prepared range through both and returns how many quotients differ in any bit from __fdiv_rn. */
EMVS_API int emvs_selftest_division(emvs_context* ctx, uint64_t n_pairs, uint32_t seed, uint64_t* mismatches);
/* The same for a whole later emvs_mapper_evaluate_dsi call: the event upload starts now AND the host packet stage
 * runs now (in the calling thread, while the device is busy with earlier work), its packets follow the events to
 * the device.  The later emvs_mapper_evaluate_dsi[_flags] call with the same (mapper, events, n_events, traj, n_poses,
 * *T_rv_w) only launches kernels; with any other arguments it falls back to what emvs_context_prefetch_events gives
 * (or to the ordinary path).  `events` and `traj` must stay valid and unchanged until that call has returned. */
EMVS_API int emvs_mapper_prefetch_dsi(emvs_mapper* m, const emvs_event* events, size_t n_events,
                             const emvs_stamped_pose* traj, size_t n_poses, const emvs_pose* T_rv_w);
EMVS_API int emvs_mapper_prefetch_dsi_soa(emvs_mapper* m, const emvs_events_soa* events, const emvs_stamped_pose* traj,
                                 size_t n_poses, const emvs_pose* T_rv_w);
/* Number of kernels this library launched on the context so far (bench `gpu_launches`). */
EMVS_API int emvs_context_launch_count(emvs_context* ctx, uint64_t* out);
/* Per-launch device timing of the vote kernel (the dominant kernel; bench.py's roofline):
 * enable -> every later vote launch is bracketed by a cudaEvent pair on the context's stream;
 * emvs_context_vote_time syncs and returns the summed duration and the launch count since the
 * last enable. */
EMVS_API int emvs_context_profile_vote(emvs_context* ctx, int enable);
EMVS_API int emvs_context_vote_time(emvs_context* ctx, double* total_ms, uint64_t* n_launches);
/* The context's cudaStream_t, for callers that time with CUDA events. */
EMVS_API int emvs_context_stream(emvs_context* ctx, void** out_stream);

/* Device-side timing on the context's stream (cudaEvent pair). */
typedef struct emvs_timer emvs_timer;
EMVS_API int emvs_timer_create(emvs_context* ctx, emvs_timer** out);
EMVS_API int emvs_timer_destroy(emvs_timer* t);
EMVS_API int emvs_timer_start(emvs_timer* t);                   /* record the start event */
EMVS_API int emvs_timer_stop(emvs_timer* t);                    /* record the stop event  */
EMVS_API int emvs_timer_elapsed_ms(emvs_timer* t, float* ms);   /* waits for the stop event */

/* Pinned host memory helpers (cudaHostAlloc / cudaFreeHost). */
EMVS_API int emvs_host_alloc(size_t bytes, void** out);
EMVS_API int emvs_host_free(void* p);

/* ---- host-side geometry (no GPU needed) ------------------------------------------------- */

/* raw_depths_vec_ = [cellIndexToDepth(i)]  (DV:88-103 / DV:131-148, MAP:213-214). */
EMVS_API int emvs_depth_vector(const emvs_shape* shape, float* out_depths /* dimZ */);
/* Virtual camera of the DSI (MAP:219-239): out = {fx, fy, cx, cy}. */
EMVS_API int emvs_virtual_camera(const emvs_camera* cam, const emvs_shape* shape, float out[4]);
/* MapperEMVS::precomputeRectifiedPoints (MAP:244-299): the raw-pixel -> rectified-pixel LUT, out =
 * width*height interleaved (x,y) floats indexed y*width + x.  K (3x3), R (3x3), P (3x4) row-major as in
 * sensor_msgs/CameraInfo; D = k1,k2,p1,p2[,k3[,k4,k5,k6[,s1..s4[,tx,ty]]]] for plumb_bob (as
 * image_geometry::rectifyPoint -> cv::undistortPoints; all-zero D returns the raw pixel) or k1..k4 for fisheye
 * (cv::fisheye::undistortPoints).  Any other model is the reference's "Distortion model not set properly!". */
enum { EMVS_DISTORTION_NONE = 0, EMVS_DISTORTION_PLUMB_BOB = 1, EMVS_DISTORTION_FISHEYE = 2 };
EMVS_API int emvs_rectify_lut(int distortion_model, const double K[9], const double* D, int n_d, const double R[9],
                     const double P[12], uint32_t width, uint32_t height, float* out_lut_xy);
/* LinearTrajectory::getPoseAt (TRJ:92-127).  *found = 0 when t is outside the control poses. */
EMVS_API int emvs_trajectory_pose_at(const emvs_stamped_pose* traj, size_t n_poses, uint32_t sec, uint32_t nsec,
                            emvs_pose* out, int* found);
EMVS_API int emvs_pose_compose(const emvs_pose* a, const emvs_pose* b, emvs_pose* out); /* a * b */
EMVS_API int emvs_pose_inverse(const emvs_pose* a, emvs_pose* out);
/* Packet stage of evaluateDSI (MAP:86-126): groups events in packets of 1024, one pose per
 * packet at the timestamp of its middle event, skipping one event on a pose miss.  Writes up
 * to max_packets packets and the number produced to *n_packets.  Returns EMVS_ERR_TOO_FEW when
 * n_events < 1024, and EMVS_ERR_INVALID (with the packets written so far still valid) when max_packets is too
 * small for the list — n_events / 1024 + 1 always suffices. */
EMVS_API int emvs_packetize(const emvs_event* events, size_t n_events,
                   const emvs_stamped_pose* traj, size_t n_poses, const emvs_pose* T_rv_w,
                   const emvs_camera* cam, const float virt[4], float z0,
                   emvs_packet* out, size_t max_packets, size_t* n_packets);
/* The same packet stage for a structure-of-arrays list (timestamps from events->t_ns). */
EMVS_API int emvs_packetize_soa(const emvs_events_soa* events, const emvs_stamped_pose* traj, size_t n_poses,
                       const emvs_pose* T_rv_w, const emvs_camera* cam, const float virt[4], float z0,
                       emvs_packet* out, size_t max_packets, size_t* n_packets);
/* Resumable form for streaming callers: continues the same packet loop from event index *cursor and stops before
 * the first packet that would reach past `event_limit` (events [0, event_limit) are known so far; the full list
 * has n_events).  *cursor is advanced; successive calls with growing limits produce exactly the packets of one
 * emvs_packetize call over the whole list. */
EMVS_API int emvs_packetize_range(const emvs_event* events, size_t n_events,
                   const emvs_stamped_pose* traj, size_t n_poses, const emvs_pose* T_rv_w,
                   const emvs_camera* cam, const float virt[4], float z0, size_t* cursor, size_t event_limit,
                   emvs_packet* out, size_t max_packets, size_t* n_packets);

/* ---- Grid3D ------------------------------------------------------------------------------ */
EMVS_API int emvs_grid_create(emvs_context* ctx, uint32_t dimX, uint32_t dimY, uint32_t dimZ, emvs_grid** out); /* G3C:30-45 (zeroed) */
EMVS_API int emvs_grid_destroy(emvs_grid* g);
EMVS_API int emvs_grid_dims(const emvs_grid* g, uint32_t* dimX, uint32_t* dimY, uint32_t* dimZ);               /* G3H:230-235 */
EMVS_API int emvs_grid_reset(emvs_grid* g);                                                                     /* G3C:67-70 */
/* a = op(a, b[, n, eps]);  b may be NULL for the two finalisers.  (G3H:64-192) */
EMVS_API int emvs_grid_op(emvs_grid* a, const emvs_grid* b, int op, int n, float eps);
/* a = b (device copy) — the `resetGrid(); addTwoGrids(g)` initialisation idiom, P1:126-127. */
EMVS_API int emvs_grid_copy(emvs_grid* dst, const emvs_grid* src);
/* Host <-> device transfer of the whole volume, layout x + dimX*(y + dimY*z) (G3H:34-35). */
EMVS_API int emvs_grid_download(const emvs_grid* g, float* host_out);
EMVS_API int emvs_grid_upload(emvs_grid* g, const float* host_in);
/* Grid3D::computeMeanSquare (G3C:164-174): sum(x^2)/N in double. */
EMVS_API int emvs_grid_mean_square(const emvs_grid* g, double* out);
/* Grid3D::collapseMaxZSlice (G3C:115-137) + convertDepthIndicesToValues (MAP:302-313).
 * conf: dimY*dimX float; idx: dimY*dimX of uint8 when dimZ <= 256 (reference CV_8U) else
 * uint16 (extension); depth (may be NULL, needs `depths`): depths[idx].  First maximum wins. */
EMVS_API int emvs_grid_collapse_max(const emvs_grid* g, const float* depths, float* conf, void* idx, float* depth);
/* Fast path of process_1 step 2+3 (P1:126-191, MAP:367-369): n-ary fusion of `n` grids with
 * stereo-fusion id `method` (1..6) folded left-to-right exactly like the reference's pairwise
 * calls (HM with n == 3 uses the HM_N step for the third grid, P1:176; GM/AM/RMS with n > 2 are
 * n-ary extensions folded pairwise), then Z-argmax, without materialising the fused volume.
 * If fused_out != NULL the fused volume is also written there. */
EMVS_API int emvs_fuse_collapse(emvs_grid* const* grids, int n, int method, const float* depths,
                       emvs_grid* fused_out, float* conf, void* idx, float* depth);
/* Same sweep with DEVICE output pointers and a DEVICE depth table (emvs_mapper_depths_device);
 * asynchronous on the context's stream: nothing is copied to the host (bench `value`). */
EMVS_API int emvs_fuse_collapse_device(emvs_grid* const* grids, int n, int method, const float* d_depths,
                              emvs_grid* fused_out, float* d_conf, void* d_idx, float* d_depth);
/* Options of MapperEMVS::getDepthMapFromDSI that the device post-processing reads
 * (OptionsDepthMap, MHP:68-82; defaults of main.cpp:73-75,97). */
typedef struct emvs_depthmap_options {
  int32_t adaptive_threshold_kernel_size;  /* 3, 5 or 7 (every shipped .conf uses the default 5)  */
  double adaptive_threshold_c;
  double max_confidence;                   /* 0: the reference's default                           */
  int32_t median_filter_size;              /* odd, >= 1                                            */
} emvs_depthmap_options;
/* MapperEMVS::getDepthMapFromDSI (MAP:339-436) for method = -1 WITHOUT the Telea inpainting:
 * n-ary fusion (n == 1: none) + collapseMaxZSlice, conf(0,0) = max_confidence, min-max
 * normalisation to 8 bit, adaptive Gaussian threshold -> mask, masked Huang median of the depth
 * indices, border removal, depth = depths[filtered index].  Host outputs, dimY*dimX each:
 * depth_map f32, confidence_map f32 (with (0,0) overwritten like the reference leaves it),
 * mask u8 (0/1), idx_filtered u8 (may be NULL).  dimZ must be <= 256 (the reference's limit). */
EMVS_API int emvs_depth_map_from_dsi(emvs_grid* const* grids, int n, int method, const float* depths,
                            const emvs_depthmap_options* opt, float* depth_map, float* confidence_map,
                            uint8_t* mask, uint8_t* idx_filtered);
/* The post-processing alone on HOST confidence / index maps (rows x cols), same outputs. */
EMVS_API int emvs_depth_map_postprocess(emvs_context* ctx, const float* conf_in, const uint8_t* idx_in, uint32_t rows,
                               uint32_t cols, const float* depths, uint32_t n_depths, const emvs_depthmap_options* opt,
                               float* depth_map, float* confidence_map, uint8_t* mask, uint8_t* idx_filtered,
                               uint8_t* conf8 /* may be NULL */);
/* Raw device pointer of the volume (for collectives run by the caller, e.g. torch.distributed). */
EMVS_API int emvs_grid_device_ptr(const emvs_grid* g, void** out);

/* ---- MapperEMVS ---------------------------------------------------------------------------- */
/* MapperEMVS::MapperEMVS(cam, shape) (MAP:29-64): derives the depth table and the virtual
 * camera, allocates the DSI.  The rectification LUT (MAP:256-299, built by OpenCV /
 * image_geometry in the reference) is an INPUT: set it with emvs_mapper_set_lut. */
EMVS_API int emvs_mapper_create(emvs_context* ctx, const emvs_camera* cam, const emvs_shape* shape, emvs_mapper** out);
EMVS_API int emvs_mapper_destroy(emvs_mapper* m);
/* precomputed_rectified_points_: interleaved (x,y) float pairs indexed y*width + x (MAP:275,296). */
EMVS_API int emvs_mapper_set_lut(emvs_mapper* m, const float* lut_xy, size_t n_pixels);
EMVS_API int emvs_mapper_shape(const emvs_mapper* m, emvs_shape* shape_out, float virt_out[4]);
EMVS_API int emvs_mapper_depths(const emvs_mapper* m, float* out_depths /* dimZ */);
EMVS_API int emvs_mapper_grid(emvs_mapper* m, emvs_grid** out);   /* the public member dsi_ (MHP:116)   */
EMVS_API int emvs_mapper_depths_device(const emvs_mapper* m, const float** out_d_depths); /* device copy of the table */
/* Event stage + resetGrid + fillVoxelGrid of evaluateDSI (MAP:129-205) for packets already
 * computed by the packet stage.  events/packets are HOST pointers (pinned memory uploads at full
 * PCIe rate); the call returns once they are uploaded, so the upload of the next camera's events
 * overlaps this camera's vote kernels. */
EMVS_API int emvs_mapper_build(emvs_mapper* m, const emvs_event* events, size_t n_events,
                      const emvs_packet* packets, size_t n_packets, int flags);
/* Same with DEVICE-resident events/packets (bench `value`: inputs already in HBM). */
EMVS_API int emvs_mapper_build_device(emvs_mapper* m, const void* d_events, size_t n_events,
                             const void* d_packets, size_t n_packets, int flags);
/* Whole MapperEMVS::evaluateDSI (MAP:67-148): packet stage on the host, the rest on the GPU.
 * Returns EMVS_ERR_TOO_FEW (-> `false`) when n_events < 1024. */
EMVS_API int emvs_mapper_evaluate_dsi(emvs_mapper* m, const emvs_event* events, size_t n_events,
                             const emvs_stamped_pose* traj, size_t n_poses, const emvs_pose* T_rv_w);
/* emvs_mapper_evaluate_dsi with build flags (EMVS_BUILD_ALLREDUCE / EMVS_BUILD_PEER_REDUCE for a rank's shard). */
EMVS_API int emvs_mapper_evaluate_dsi_flags(emvs_mapper* m, const emvs_event* events, size_t n_events,
                                   const emvs_stamped_pose* traj, size_t n_poses, const emvs_pose* T_rv_w, int flags);
/* MapperEMVS::evaluateDSI (MAP:67-148) for a structure-of-arrays event list: same packets, same votes, same DSI as the
 * dvs_msgs::Event form with equal x, y and timestamps; a quarter of the PCIe traffic.  flags as above. */
EMVS_API int emvs_mapper_evaluate_dsi_soa(emvs_mapper* m, const emvs_events_soa* events, const emvs_stamped_pose* traj,
                                 size_t n_poses, const emvs_pose* T_rv_w, int flags);
/* Build-defined integer observable: accepted (event, plane k) votes of the last
 * build(s) since the last RESET, one uint64 per plane. */
EMVS_API int emvs_mapper_counts(const emvs_mapper* m, uint64_t* per_plane /* dimZ */);

/* ---- multi-GPU (one process per GPU) --------------------------------------------------------- */
/* NCCL is resolved at run time (dlopen of the libnccl already loaded in the process, else
 * libnccl.so.2).  Rank 0 calls emvs_comm_unique_id and ships the 128 bytes to the others. */
EMVS_API int emvs_comm_unique_id(uint8_t out_id[128]);
EMVS_API int emvs_comm_init(emvs_context* ctx, const uint8_t id[128], int n_ranks, int rank);
EMVS_API int emvs_comm_destroy(emvs_context* ctx);
/* In-place sum of a partial DSI over all ranks (ncclAllReduce float32 sum, chunked by Z-slab). */
EMVS_API int emvs_grid_allreduce(emvs_grid* g);
/* Same, enqueued on the context's stream without waiting for completion. */
EMVS_API int emvs_grid_allreduce_async(emvs_grid* g);
/* Sum of the per-plane vote counts over all ranks. */
EMVS_API int emvs_mapper_counts_allreduce(emvs_mapper* m);


/* ---- multi-GPU, fused: reduce + fuse + argmax over NVLink peer memory ----------------------------
 * Alternative to allreduce + emvs_fuse_collapse when only the depth / confidence maps are needed:
 * rank r owns a band of image rows and reads the partial voxels of ALL ranks for that band
 * directly from their HBM (CUDA IPC mappings, loads through NVSwitch), sums them in rank order,
 * fuses the cameras, takes the Z-argmax and stores the result into the map buffers of every rank.
 * The per-camera DSIs stay partial on every rank; no summed volume is written anywhere.
 *
 * Setup (collective, once): every rank creates an exchange over its n_cams partial grids (same
 * shapes, same camera order), exports a blob of emvs_exchange_blob_bytes() bytes, the caller
 * all-gathers the blobs (e.g. torch.distributed) and every rank imports the concatenation in rank
 * order.  Per step (collective): after the local builds have been issued,
 * emvs_exchange_fuse_collapse enqueues {signal "built", the fused sweep, signal "done", wait for
 * all "done"} on the context's stream; when the stream reaches the end of that sequence the local
 * map buffers hold the full maps and no peer reads this rank's DSIs any more (so they may be
 * rebuilt).  A peer that never arrives is reported as EMVS_ERR_STATE by the next
 * emvs_exchange_download after a timeout instead of hanging the GPU. */
typedef struct emvs_exchange emvs_exchange;
EMVS_API int emvs_exchange_create(emvs_context* ctx, emvs_grid* const* grids, int n_cams, int n_ranks, int rank,
                         emvs_exchange** out);
EMVS_API int emvs_exchange_destroy(emvs_exchange* ex);
EMVS_API int emvs_exchange_blob_bytes(const emvs_exchange* ex, size_t* out);
EMVS_API int emvs_exchange_export(emvs_exchange* ex, uint8_t* blob);
EMVS_API int emvs_exchange_import(emvs_exchange* ex, const uint8_t* all_blobs /* n_ranks * blob_bytes */);
/* Camera x sub-interval sharding ("give each camera G/C GPUs and split its packet list").  builds[c *
 * n_ranks + r] != 0 iff rank r builds (a sub-interval of) camera c; the same table on every rank.  Camera c is then
 * summed over ITS ranks only, nobody reads or waits for the others' grid of that camera, and a rank builds only its
 * own cameras between emvs_exchange_begin and emvs_exchange_fuse_collapse.  Default: every rank builds every camera.
 * Every camera needs at least one rank and every rank at least one camera. */
EMVS_API int emvs_exchange_set_participants(emvs_exchange* ex, const uint8_t* builds);
/* Optional overlapped form of a round: emvs_exchange_begin, then build every camera of the exchange with
 * EMVS_BUILD_PEER_REDUCE (same order on every rank), then emvs_exchange_fuse_collapse, which in that case only
 * runs a local sweep over the already-reduced band and distributes it.  Without begin, fuse_collapse does the
 * whole reduction in one exposed sweep after the builds. */
EMVS_API int emvs_exchange_begin(emvs_exchange* ex);
/* method: EMVS_FUSE_*; d_depths: DEVICE depth table (emvs_mapper_depths_device) or NULL. */
EMVS_API int emvs_exchange_fuse_collapse(emvs_exchange* ex, int method, const float* d_depths);
/* Device pointers of this rank's map buffers (conf f32, idx u8|u16, depth f32; dimY*dimX each). */
EMVS_API int emvs_exchange_maps(const emvs_exchange* ex, float** d_conf, void** d_idx, float** d_depth);
/* Waits for the pipeline and copies the maps to the host (any pointer may be NULL). */
EMVS_API int emvs_exchange_download(emvs_exchange* ex, float* conf, void* idx, float* depth);

#ifdef __cplusplus
}
#endif
#endif /* EMVS_B200_H_ */
