// emvs_engine.cu — context / grid / mapper objects and the C-ABI of include/emvs_b200.h.
// Host orchestration only; device code lives in emvs_kernels.cuh.  No CPU fallback: every
// compute entry point launches CUDA kernels or fails.

#include "emvs_internal.h"
#include "emvs_kernels.cuh"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

namespace emvs {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...)
{
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
}

}  // namespace emvs

using namespace emvs;

#define CUDA_TRY(expr)                                                                          \
  do {                                                                                          \
    cudaError_t e_ = (expr);                                                                    \
    if (e_ != cudaSuccess) {                                                                    \
      set_error("CUDA error %s at %s:%d: %s", cudaGetErrorName(e_), __FILE__, __LINE__, #expr); \
      return EMVS_ERR_CUDA;                                                                     \
    }                                                                                           \
  } while (0)

#define REQUIRE(cond, code, msg)  \
  do {                            \
    if (!(cond)) {                \
      set_error("%s", msg);       \
      return code;                \
    }                             \
  } while (0)

// ---------------------------------------------------------------------------------------------
// objects
// ---------------------------------------------------------------------------------------------
struct emvs_context {
  // Grids, mappers and timers keep their context alive: emvs_context_destroy only drops the
  // caller's reference, the teardown happens when the last dependent object is destroyed (so a
  // garbage-collected binding may release objects in any order).
  std::atomic<int> refs{1};
  int device = 0;
  cudaStream_t stream = nullptr;
  int sm_count = 148;
  uint64_t launches = 0;
  uint32_t slab_override = 0;
  // Quad scratch: two buffers of one slab of planes each (invariant: all zero between builds).
  // Slab s is voted into buffer s&1 on `stream` while buffer (s-1)&1 is merged into the DSI and
  // re-zeroed on `aux_stream` (EMVS_OVERLAP=0 serialises everything on `stream` with one buffer).
  float4* quad[2] = {nullptr, nullptr};
  size_t quad_bytes[2] = {0, 0};
  cudaStream_t aux_stream = nullptr;
  cudaEvent_t ev_vote[2] = {nullptr, nullptr};    // slab voted into buffer b
  cudaEvent_t ev_merge[2] = {nullptr, nullptr};   // buffer b merged + zeroed again
  bool merge_pending[2] = {false, false};
  bool overlap = true;
  // grow-only device staging
  // two event staging buffers: the list of the next build (or a prefetched list) is uploaded into one while the
  // event stage of the current build may still be reading the other
  void* d_events[2] = {nullptr, nullptr};  size_t events_cap[2] = {0, 0};
  unsigned upload_next = 0;            // staging buffer the next upload goes to
  int cur_events = 0;                  // staging buffer the current / last host-buffer build reads
  struct {                             // emvs_context_prefetch_events: a list already on its way to d_events[buf]
    const void* host = nullptr;        // the caller's event array (AoS) or x array (SoA): the key the later call is matched by
    bool soa = false;
    size_t n = 0;
    int buf = 0;
    bool valid = false;
    uint64_t generation = 0;
    // emvs_mapper_prefetch_dsi: its packet stage has run too and the packets are on their way to d_packets[par]
    bool has_packets = false;
    size_t n_pk = 0;
    unsigned par = 0;
    const struct emvs_mapper* mapper = nullptr;
    const emvs_stamped_pose* traj = nullptr;
    size_t n_poses = 0;
    emvs_pose T_rv_w{};
  } prefetch;
  emvs_packet* h_packets_pf = nullptr; size_t h_packets_pf_cap = 0;   // pinned packets of the pending prefetch
  cudaEvent_t ev_prefetched = nullptr;   // the prefetch's copies (events, packets) have landed
  bool prefetched_recorded = false;
  cudaEvent_t ev_pk_free[2] = {nullptr, nullptr};   // the last build reading d_packets[i] has finished voting
  bool pk_free_recorded[2] = {false, false};
  unsigned cur_packets = 0;            // d_packets buffer of the current host-buffer build
  void* d_packets[2] = {nullptr, nullptr}; size_t packets_cap[2] = {0, 0};  // alternate per build (see build_from_host)
  // host->device staging runs on its own stream so that the upload of the next camera's events
  // overlaps the vote kernels of the previous one
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_copied = nullptr;     // inputs of the current build are in HBM
  cudaEvent_t ev_consumed[2] = {nullptr, nullptr};   // k_warp_events of the last build on d_events[i] has read it
  bool consumed_recorded[2] = {false, false};
  bool mark_consumed = false;          // build_on_device records ev_consumed after the event stage
  float2* d_xy0 = nullptr;   size_t xy0_cap = 0;
  // k_vote_tma: one work counter per vote launch of a build (packets are handed out dynamically), zeroed when the
  // build is issued; the persistent grid leaves one 256-thread slot per SM to the merge / exchange kernels
  unsigned int* d_work = nullptr;
  uint32_t vote_ctas_per_sm = 7;       // EMVS_VOTE_CTAS_PER_SM: resident CTAs per SM of the persistent vote grid (1..8)
  int vote_split = -1;                 // EMVS_VOTE_SPLIT: log2 of work items per packet (0..2); -1: automatic
  // EMVS_VOTE_MULTISLAB: ONE vote launch per build walks all Z-slabs (one scratch buffer per slab, merges started from
  // per-slab completion counters) instead of one launch per slab; used when the slab buffers fit multislab_budget
  // (EMVS_MULTISLAB_BUDGET_MB, default 4096: 640x480x256 takes 1.26 GB).  Measured 7.52 against 7.83 ms per stereo
  // window with device-resident inputs, 9.13 against 10.02 ms through the host-buffer calls (profiles/r2_ab_knobs.md)
  bool vote_multislab = true;
  size_t multislab_budget = (size_t)4 << 30;
  float4* quad_ms = nullptr;  size_t quad_ms_bytes = 0;
  bool ms_check_pending = false;       // a multi-slab build was issued since the error word was last read (emvs_context_sync)
  cudaEvent_t ev_ms_start = nullptr;
  bool vote_tma = true;                // EMVS_VOTE_KERNEL=classic selects k_vote_grouped (one CTA per packet, A/B baseline)
  // tuning / experiment knobs, read from the environment when the context is created (tools/ab_bench.py)
  int zero_ctas = 592;                 // EMVS_ZERO_CTAS: grid of the re-zero kernel (its stores carry the evict_last hint); 0: cudaMemsetAsync
  int peer_reduce_ctas = 1024;           // EMVS_PEER_REDUCE_CTAS: persistent grid of the slab-wise peer reduce (0: one thread per voxel)
  bool fc_v4 = true;                   // EMVS_FC_V4: four pixels per thread in the fuse + collapse sweep (0: one pixel per thread)
  int fc_zsplit = 8;                   // EMVS_FC_ZSPLIT: plane chunks of the fuse + collapse sweep
  bool dbg_skip_merge = false, dbg_skip_zero = false;   // EMVS_DEBUG_SKIP_MERGE / _ZERO: timing experiments, WRONG results
  // L2 eviction hints (0 none, 1 evict_first, 2 evict_last): EMVS_HINT_XY0 (event-tile bulk copies), EMVS_HINT_RED (vote
  // REDs, 0 / 2), EMVS_HINT_DSI (1: streaming stores of the merged planes), EMVS_HINT_ZERO (re-zero kernel stores, needs
  // EMVS_ZERO_CTAS > 0)
  // Defaults measured in profiles/r2_l2_hints.md: -2.5 % per device-resident step, -1.4 % end to end and at N = 2.
  int hint_xy0 = 1, hint_red = 2, hint_dsi = 1, hint_zero = 2;
  uint64_t prefetch_generation = 0;    // bumped by every prefetch; emvs_context_prefetch_pending reports the pending one
  void* d_out = nullptr;     size_t out_cap = 0;      // conf | depth | idx of a collapse
  void* d_fc_part = nullptr; size_t fc_part_cap = 0;  // per-chunk (max, index) of the Z-split sweep
  double* d_partial = nullptr;                        // 1024 partial sums + 1 result
  // pinned host staging for packets produced by evaluate_dsi
  emvs_packet* h_packets = nullptr; size_t h_packets_cap = 0;
  // evaluate_dsi on an idle pipeline: the head of the event list (split_percent %) is uploaded and voted first
  // while the tail is still crossing PCIe (see emvs_mapper_evaluate_dsi_flags); 0 disables
  uint32_t split_percent = 15;
  int split_pieces = 2;                // EMVS_UPLOAD_PIECES (2..4): pieces of p %, 2.2 p %, 2.2^2 p % ... of the list, the last one takes the rest
  // EMVS_UPLOAD_DEFER_MERGE: the pieces of a split upload before the last one only VOTE (into the per-slab scratch of the
  // multi-slab launch, which persists between launches); the last piece votes on top and merges once.  Without it every
  // piece is a complete build (merge + re-zero of all slabs, accumulating from the second piece on).
  // Used for builds that take the multi-slab launch and are not part of a peer exchange, with its own geometry
  // (defer_percent / defer_pieces); every other build splits into split_pieces complete builds.  Measured: stock e2e 8.69
  // against 9.15 ms per stereo window (profiles/r2_e2e.md, trip 23).
  bool split_defer_merge = true;
  uint32_t defer_percent = 6;          // EMVS_UPLOAD_DEFER_SPLIT
  int defer_pieces = 4;                // EMVS_UPLOAD_DEFER_PIECES (2..4)
  bool deferred_votes = false;         // the scratch of quad_ms holds votes of a piece whose merge is still to come
  size_t split_min_events = (size_t)1 << 20;
  // NCCL
  void* comm = nullptr;
  int n_ranks = 1, rank = 0;
  cudaStream_t comm_stream = nullptr;      // slab allreduces / peer reduces run here, overlapped with the next slab's votes
  struct emvs_exchange* active_exchange = nullptr;   // set by emvs_exchange_begin for EMVS_BUILD_PEER_REDUCE builds
  std::vector<cudaEvent_t> slab_events;    // slab s merged -> its allreduce may start
  cudaEvent_t ev_comm_done = nullptr;
  // optional per-launch timing of the vote kernel (bench roofline): event pairs on `stream`
  bool profile = false;
  std::vector<cudaEvent_t> prof_events;   // pairs (start, stop), prof_used of them recorded
  size_t prof_used = 0;
};

struct emvs_grid {
  emvs_context* ctx = nullptr;
  uint32_t dimX = 0, dimY = 0, dimZ = 0;
  size_t n_cells = 0;
  float* d = nullptr;
};

struct emvs_mapper {
  emvs_context* ctx = nullptr;
  emvs_camera cam{};
  emvs_shape shape{};   // resolved (dimX/dimY filled in)
  float virt[4] = {0, 0, 0, 0};
  std::vector<float> depths;
  float* d_depths = nullptr;
  float2* d_lut = nullptr;
  bool lut_set = false;
  emvs_grid* grid = nullptr;
  unsigned long long* d_counts = nullptr;
};

constexpr uint32_t kMaxSlabs = 256;            // per camera, for the slab-wise peer reduce
constexpr uint32_t kFlagWordsPhase = 32;        // [2][kMaxPeerRanks] phase epochs, word 16 = error, rest padding
struct emvs_exchange {
  emvs_context* ctx = nullptr;
  int n_cams = 0, n_ranks = 1, rank = 0;
  uint32_t dimX = 0, dimY = 0, dimZ = 0;
  uint32_t row_lo = 0, row_hi = 0;               // the row band this rank owns
  const float* local_dsi[kMaxPeerCams] = {};
  char* maps = nullptr;            // conf | depth | idx of this rank (one allocation, IPC-exported)
  size_t off_depth = 0, off_idx = 0, maps_bytes = 0;
  // IPC-exported flag words: [0..15] phase epochs [2][kMaxPeerRanks], [16] error,
  // [kFlagWordsPhase + ((cam * kMaxSlabs + slab) * kMaxPeerRanks + rank)] slab epochs
  unsigned int* flags = nullptr;
  float* band_buf = nullptr;       // [n_cams][dimZ][band pixels]: this rank's band summed over the ranks
  PeerArgs args{};
  FlagPtrs flag_ptrs{};
  std::vector<void*> opened;       // cudaIpcOpenMemHandle mappings to close
  unsigned int epoch = 0;
  bool imported = false;
  bool band_round = false;         // emvs_exchange_begin was called: builds reduce their slabs into band_buf
  int first_local_cam = 0;         // lowest camera this rank builds (emvs_exchange_set_participants)
};

namespace {

struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev)
  {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != dev) ok = cudaSetDevice(dev) == cudaSuccess;
  }
  ~DeviceGuard()
  {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

int grow(void** p, size_t* cap, size_t need)
{
  if (need <= *cap) return EMVS_OK;
  if (*p) CUDA_TRY(cudaFree(*p));
  *p = nullptr;
  *cap = 0;
  CUDA_TRY(cudaMalloc(p, need));
  *cap = need;
  return EMVS_OK;
}

inline uint32_t ceil_div(uint32_t a, uint32_t b) { return (a + b - 1) / b; }

int env_int(const char* name, int dflt)
{
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

// Planes voted per pass over the event list (the slab) and planes whose quads are interleaved in the scratch (the
// plane group G of k_vote_grouped).  Measured on B200 (126 MB L2) at 640x480, profiles/r1_vote_group.md:
//   * G = 8 with the events staged in shared memory and the grouped merge is the fastest (1235 Mevents/s against
//     1097 at G = 4, 811 at G = 2, 739 ungrouped); G = 16 is no better;
//   * two scratch buffers of 16 planes (79 MB each; slab s is voted while slab s-1 is merged and re-zeroed) beat
//     2 x 8 (1200) and one buffer of 16 without overlap (1175), although together they exceed the L2.
// The slab is a multiple of G (idle lanes otherwise) and at least one full group.
struct SlabPlan {
  uint32_t slab, group;
};

SlabPlan choose_slab(const emvs_context* ctx, uint32_t dimX, uint32_t dimY, uint32_t dimZ)
{
  static const uint32_t group_env = [] {
    const char* e = getenv("EMVS_VOTE_GROUP");
    const int g = e ? atoi(e) : 0;
    return (uint32_t)((g == 1 || g == 2 || g == 4 || g == 8 || g == 16 || g == 32) ? g : 0);
  }();
  uint32_t s = ctx->slab_override;
  if (!s) {
    if (const char* env = getenv("EMVS_SLAB")) s = (uint32_t)atoi(env);
  }
  const bool fixed = s != 0;
  if (!s) {
    const size_t plane_bytes = (size_t)ceil_div(dimX, 2) * ceil_div(dimY, 2) * 64;
    const size_t budget = (size_t)80 << 20;   // per buffer; two buffers when overlapped
    s = (uint32_t)std::max<size_t>(1, budget / std::max<size_t>(plane_bytes, 1));
  }
  s = std::min(s, dimZ);
  s = std::min<uint32_t>(s, 1024);
  s = std::max<uint32_t>(s, 1);
  uint32_t g = group_env;
  if (!g) {
    g = 8;
    while (g > 1 && g > (fixed ? s : dimZ)) g >>= 1;   // tiny volumes / tiny explicit slabs
  }
  // a full group of 8 planes pays even when its scratch no longer fits the L2: 1024x1024 planes (16.8 MB of quads
  // each) run 13 % faster with 2 x 8 planes (268 MB) than with the 2 x 4 planes the budget allows (trip 28)
  if (!fixed) s = std::min(dimZ, std::max(g, s / g * g));
  return SlabPlan{s, g};
}

int ensure_quad(emvs_context* ctx, int b, size_t bytes)
{
  if (bytes <= ctx->quad_bytes[b]) return EMVS_OK;
  if (ctx->quad[b]) CUDA_TRY(cudaFree(ctx->quad[b]));   // cudaFree waits for the device: no kernel still uses it
  ctx->quad[b] = nullptr;
  ctx->quad_bytes[b] = 0;
  CUDA_TRY(cudaMalloc((void**)&ctx->quad[b], bytes));
  ctx->quad_bytes[b] = bytes;
  CUDA_TRY(cudaMemsetAsync(ctx->quad[b], 0, bytes, ctx->stream));
  return EMVS_OK;
}

int ncclAllReduce_checked(const NcclApi* api, emvs_context* ctx, float* p, size_t count, cudaStream_t st)
{
  const int r = api->AllReduce(p, p, count, /*ncclFloat32*/ 7, /*ncclSum*/ 0, ctx->comm, (void*)st);
  if (r != 0) set_error("ncclAllReduce(float32, sum) failed: %s", api->GetErrorString(r));
  return r;
}

int ncclAllReduceU64_checked(const NcclApi* api, emvs_context* ctx, unsigned long long* p, size_t count, cudaStream_t st)
{
  const int r = api->AllReduce(p, p, count, /*ncclUint64*/ 5, /*ncclSum*/ 0, ctx->comm, (void*)st);
  if (r != 0) set_error("ncclAllReduce(uint64, sum) failed: %s", api->GetErrorString(r));
  return r;
}

// Slab-wise reduce-scatter over NVLink peer memory (EMVS_BUILD_PEER_REDUCE).  `after` is the stream on which
// planes [k0, k0+nk) of camera `cam` become final on this rank (the merge stream): the "slab built" epoch is
// published to every rank from there, and the band reduce is ordered behind it on the communication stream.
// launches the band reduce of (cam, slab) on the communication stream (it waits for the slab flags of the ranks that
// build `cam`)
void launch_band_reduce(emvs_context* ctx, emvs_exchange* ex, int cam, uint32_t slab_idx, uint32_t k0, uint32_t nk)
{
  const uint32_t word = kFlagWordsPhase + ((uint32_t)cam * kMaxSlabs + slab_idx) * kMaxPeerRanks;
  const uint32_t p_lo = ex->row_lo * ex->dimX, p_hi = ex->row_hi * ex->dimX, band = p_hi - p_lo;
  if (!band) return;
  float* out = ex->band_buf + ((size_t)cam * ex->dimZ + k0) * band;
  // EMVS_PEER_REDUCE_CTAS: size of the grid-stride reduce grid (0: one thread per voxel)
  const int reduce_ctas = ctx->peer_reduce_ctas;
  const uint32_t n_pix = ex->dimX * ex->dimY;
  if (reduce_ctas > 0 && p_lo % 4 == 0 && band % 4 == 0 && n_pix % 4 == 0) {
    k_peer_reduce_band_v4<<<(unsigned)reduce_ctas, 256, 0, ctx->comm_stream>>>(ex->args, cam, ex->flags + word, ex->epoch,
                                                                               40000000000LL, ex->flags + 16, p_lo, p_hi, n_pix,
                                                                               k0, nk, out);
  } else {
    const dim3 grid((band + 255) / 256, nk);
    k_peer_reduce_band<<<grid, 256, 0, ctx->comm_stream>>>(ex->args, cam, ex->flags + word, ex->epoch, 40000000000LL,
                                                           ex->flags + 16, p_lo, p_hi, n_pix, k0, nk, out);
  }
  ctx->launches++;
}

int peer_reduce_slab(emvs_context* ctx, emvs_exchange* ex, int cam, uint32_t slab_idx, uint32_t k0, uint32_t nk,
                     cudaStream_t after)
{
  REQUIRE(slab_idx < kMaxSlabs, EMVS_ERR_INVALID, "peer_reduce: slab index out of range");
  const uint32_t word = kFlagWordsPhase + ((uint32_t)cam * kMaxSlabs + slab_idx) * kMaxPeerRanks;
  k_flag_signal_word<<<1, 32, 0, after>>>(ex->flag_ptrs, ex->n_ranks, word + (uint32_t)ex->rank, ex->epoch);
  ctx->launches++;
  while (ctx->slab_events.size() <= slab_idx) {
    cudaEvent_t e;
    CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    ctx->slab_events.push_back(e);
  }
  CUDA_TRY(cudaEventRecord(ctx->slab_events[slab_idx], after));
  CUDA_TRY(cudaStreamWaitEvent(ctx->comm_stream, ctx->slab_events[slab_idx], 0));
  launch_band_reduce(ctx, ex, cam, slab_idx, k0, nk);
  // Camera x sub-interval sharding: the cameras this rank does not build are reduced for ITS row band too.  Their
  // slab reduces ride on the slab loop of the rank's first own camera (every rank walks the same slabs at about the
  // same pace, so the flags they wait for are published around the same time; a rank only ever waits for slabs
  // that its peers publish from their own vote / merge pipeline, so the slowest rank never blocks).
  if (cam == ex->first_local_cam)
    for (int c = 0; c < ex->n_cams; ++c)
      if (!((ex->args.cam_ranks[c] >> ex->rank) & 1u)) launch_band_reduce(ctx, ex, c, slab_idx, k0, nk);
  return EMVS_OK;
}

// Device-resident event list: the caller's 16-byte dvs_msgs::Event structs, or (emvs_events_soa) separate uint16
// x / y arrays.
struct EventSrc {
  const emvs_event* aos = nullptr;
  const uint16_t* x = nullptr;
  const uint16_t* y = nullptr;
};

constexpr uint32_t kMaxWorkCounters = 4096;    // vote launches (slabs) per build
// internal build flags of the split upload (never part of the C-ABI's EMVS_BUILD_* values)
constexpr int kPublicBuildFlags = EMVS_BUILD_ACCUMULATE | EMVS_BUILD_ALLREDUCE | EMVS_BUILD_PEER_REDUCE;
constexpr int kBuildDeferMerge = 1 << 16;      // vote only: the votes stay in the multi-slab scratch
constexpr int kBuildContinue = 1 << 17;        // the scratch holds votes of earlier pieces: vote on top, merge everything

// gentle re-zeroing of a merged scratch buffer: 256-thread CTAs that fit beside the persistent vote grid
__global__ void __launch_bounds__(256) k_zero_f4(float4* __restrict__ p, size_t n, int hint)
{
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  if (hint) {
    const uint64_t pol = l2_policy(hint);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
      asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %1, %1, %1}, %2;" ::"l"(p + i), "f"(0.f), "l"(pol) : "memory");
    return;
  }
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) p[i] = z;
}

// Would a build into a DSI of this size take the single multi-slab vote launch (one scratch buffer per slab)?
static bool multislab_eligible(const emvs_context* ctx, uint32_t dimX, uint32_t dimY, uint32_t dimZ)
{
  const SlabPlan plan = choose_slab(ctx, dimX, dimY, dimZ);
  const uint32_t G = plan.group, slab = plan.slab;
  if (!(ctx->vote_tma && ctx->vote_multislab && ctx->overlap && G > 1 && slab % G == 0)) return false;
  const uint32_t n_slabs = (dimZ + slab - 1) / slab;
  const size_t slab_bytes = (size_t)((slab + G - 1) / G * G) * ceil_div(dimX, 2) * ceil_div(dimY, 2) * 4 * sizeof(float4);
  return n_slabs > 1 && (size_t)n_slabs * slab_bytes <= ctx->multislab_budget;
}

// Device part of evaluateDSI: event stage, (reset), slab loop of {vote, merge, re-zero[, allreduce | peer reduce]}.
int build_on_device(emvs_mapper* m, const EventSrc& d_ev, size_t n_events, const emvs_packet* d_pk,
                    size_t n_packets, int flags)
{
  emvs_context* ctx = m->ctx;
  emvs_grid* g = m->grid;
  const bool accumulate = (flags & EMVS_BUILD_ACCUMULATE) != 0;
  const bool reduce = (flags & EMVS_BUILD_ALLREDUCE) != 0;
  // internal (split upload): vote only, leave the votes in the per-slab scratch / vote on top of such votes and merge
  const bool want_defer = (flags & kBuildDeferMerge) != 0;
  const bool cont = (flags & kBuildContinue) != 0;
  cudaStream_t st = ctx->stream;
  const bool peer = (flags & EMVS_BUILD_PEER_REDUCE) != 0;
  emvs_exchange* ex = peer ? ctx->active_exchange : nullptr;
  int peer_cam = -1;
  if (peer) {
    // (with ACCUMULATE this must be the LAST build into the DSI of the round: its merged slabs are announced as final)
    REQUIRE(!reduce, EMVS_ERR_INVALID, "build: EMVS_BUILD_PEER_REDUCE excludes EMVS_BUILD_ALLREDUCE");
    REQUIRE(ex && ex->band_round, EMVS_ERR_STATE, "build: EMVS_BUILD_PEER_REDUCE needs emvs_exchange_begin");
    for (int c = 0; c < ex->n_cams; ++c)
      if (ex->local_dsi[c] == g->d) peer_cam = c;
    REQUIRE(peer_cam >= 0, EMVS_ERR_INVALID, "build: this mapper's DSI is not part of the active exchange");
    REQUIRE((ex->args.cam_ranks[peer_cam] >> ex->rank) & 1u, EMVS_ERR_INVALID,
            "build: this rank is not a participant of that camera (emvs_exchange_set_participants)");
  }
  const NcclApi* nccl = nullptr;
  if (reduce) {
    REQUIRE(!accumulate, EMVS_ERR_INVALID, "build: EMVS_BUILD_ALLREDUCE cannot be combined with EMVS_BUILD_ACCUMULATE");
    REQUIRE(ctx->comm, EMVS_ERR_STATE, "build: EMVS_BUILD_ALLREDUCE needs emvs_comm_init");
    nccl = nccl_api();
    if (!nccl) return EMVS_ERR_NCCL;
  }
  if (!accumulate && !cont) {
    CUDA_TRY(cudaMemsetAsync(m->d_counts, 0, sizeof(unsigned long long) * g->dimZ, st));
  }
  if (n_packets == 0 && !cont) {
    if (!accumulate) CUDA_TRY(cudaMemsetAsync(g->d, 0, g->n_cells * sizeof(float), st));
    if (reduce) {  // this rank has no packets but must issue the SAME sequence of collectives as the others
      const uint32_t zslab = choose_slab(ctx, g->dimX, g->dimY, g->dimZ).slab;
      const size_t plane = (size_t)g->dimX * g->dimY;
      for (uint32_t k0 = 0; k0 < g->dimZ; k0 += zslab) {
        const uint32_t nk = std::min(zslab, g->dimZ - k0);
        if (ncclAllReduce_checked(nccl, ctx, g->d + (size_t)k0 * plane, (size_t)nk * plane, st)) return EMVS_ERR_NCCL;
      }
      if (ncclAllReduceU64_checked(nccl, ctx, m->d_counts, g->dimZ, st)) return EMVS_ERR_NCCL;
    }
    if (peer) {    // an all-zero partial DSI: announce and reduce the same slabs as the ranks that do vote
      const uint32_t zslab = choose_slab(ctx, g->dimX, g->dimY, g->dimZ).slab;
      for (uint32_t k0 = 0; k0 < g->dimZ; k0 += zslab) {
        const int rc = peer_reduce_slab(ctx, ex, peer_cam, k0 / zslab, k0, std::min(zslab, g->dimZ - k0), st);
        if (rc) return rc;   // (emvs_exchange_fuse_collapse waits for the communication stream)
      }
    }
    return EMVS_OK;
  }
  (void)n_events;   // the packets carry the event indices; n_events is validated by the host-buffer entry points
  const size_t n_voted = n_packets * (size_t)EMVS_PACKET_SIZE;
  {
    const int rc = grow((void**)&ctx->d_xy0, &ctx->xy0_cap, n_voted * sizeof(float2));
    if (rc) return rc;
  }
  const uint32_t dimX = g->dimX, dimY = g->dimY, dimZ = g->dimZ;
  const uint32_t QW = ceil_div(dimX, 2), QH = ceil_div(dimY, 2);
  const SlabPlan plan = choose_slab(ctx, dimX, dimY, dimZ);
  const uint32_t slab = plan.slab;
  // plane-grouped scratch layout + k_vote_grouped<G> (EMVS_VOTE_GROUP=1 selects the plain one-plane-per-instruction kernel)
  const uint32_t G = plan.group;
  auto round_up_g = [&](uint32_t n) { return (n + G - 1) / G * G; };
  const size_t slab_bytes = (size_t)round_up_g(slab) * QW * QH * 4 * sizeof(float4);
  const bool overlap = ctx->overlap;
  for (int b = 0; b < (overlap ? 2 : 1); ++b) {
    const int rc = ensure_quad(ctx, b, slab_bytes);
    if (rc) return rc;
  }
  // the vote kernels index a plane group with 32 bits
  REQUIRE((uint64_t)QW * QH * 4 * G < (1ull << 32), EMVS_ERR_INVALID, "build: plane too large for the 32-bit quad index");
  const uint32_t n_slabs = (dimZ + slab - 1) / slab;
  const bool use_tma = ctx->vote_tma && G > 1;
  if (use_tma) {
    REQUIRE(n_slabs <= kMaxWorkCounters, EMVS_ERR_INVALID, "build: too many slabs (raise the slab size)");
    REQUIRE(n_packets < (1ull << 31), EMVS_ERR_INVALID, "build: too many packets for one build");
    // [0, K): work counters (one per vote launch), [K, 2K): per-slab completion counters of a multi-slab launch, [2K]: error word
    if (!ctx->d_work) {
      CUDA_TRY(cudaMalloc((void**)&ctx->d_work, sizeof(unsigned int) * (2 * kMaxWorkCounters + 16)));
      CUDA_TRY(cudaMemsetAsync(ctx->d_work, 0, sizeof(unsigned int) * (2 * kMaxWorkCounters + 16), st));
    }
    CUDA_TRY(cudaMemsetAsync(ctx->d_work, 0, sizeof(unsigned int) * n_slabs, st));
    CUDA_TRY(cudaMemsetAsync(ctx->d_work + kMaxWorkCounters, 0, sizeof(unsigned int) * n_slabs, st));
  }
  // one vote launch for all slabs of this build?
  bool multislab = multislab_eligible(ctx, dimX, dimY, dimZ);
  REQUIRE(!cont || (multislab && ctx->deferred_votes && (size_t)n_slabs * slab_bytes <= ctx->quad_ms_bytes), EMVS_ERR_STATE,
          "build: no deferred votes to continue from");
  if (multislab && !cont && ctx->deferred_votes) {
    // a split upload was abandoned between its pieces (an error return): its votes must not leak into this build
    CUDA_TRY(cudaMemsetAsync(ctx->quad_ms, 0, ctx->quad_ms_bytes, st));
    ctx->deferred_votes = false;
  }
  if (multislab && (size_t)n_slabs * slab_bytes > ctx->quad_ms_bytes) {
    if (ctx->quad_ms) CUDA_TRY(cudaFree(ctx->quad_ms));
    ctx->quad_ms = nullptr;
    ctx->quad_ms_bytes = 0;
    if (cudaMalloc((void**)&ctx->quad_ms, (size_t)n_slabs * slab_bytes) != cudaSuccess) {
      (void)cudaGetLastError();
      multislab = false;          // not enough memory for one buffer per slab: per-slab launches
    } else {
      ctx->quad_ms_bytes = (size_t)n_slabs * slab_bytes;
      CUDA_TRY(cudaMemsetAsync(ctx->quad_ms, 0, ctx->quad_ms_bytes, st));
    }
  }
  // vote-only piece of a split upload: possible when this build takes the single multi-slab launch (else it is a plain build
  // and the caller continues with EMVS_BUILD_ACCUMULATE: ctx->deferred_votes stays false)
  const bool defer = want_defer && multislab;

  if (n_packets) {
    const unsigned blocks = (unsigned)((n_voted + 255) / 256);
    if (d_ev.aos)
      k_warp_events<<<blocks, 256, 0, st>>>(d_ev.aos, d_pk, m->d_lut, m->cam.width, m->cam.height, ctx->d_xy0,
                                            (unsigned long long)n_voted);
    else
      k_warp_events_soa<<<blocks, 256, 0, st>>>(d_ev.x, d_ev.y, d_pk, m->d_lut, m->cam.width, m->cam.height, ctx->d_xy0,
                                                (unsigned long long)n_voted);
    ctx->launches++;
    if (ctx->mark_consumed) {
      CUDA_TRY(cudaEventRecord(ctx->ev_consumed[ctx->cur_events], st));
      ctx->consumed_recorded[ctx->cur_events] = true;
    }
  }

  VoteParams P;
  P.vfx = m->virt[0]; P.vfy = m->virt[1]; P.vcx = m->virt[2]; P.vcy = m->virt[3];
  P.z0 = m->depths[0];
  P.xmax = (float)((int)dimX - 1);
  P.ymax = (float)((int)dimY - 1);
  P.QW = QW; P.QH = QH;

  // persistent vote grid (k_vote_tma): vote_ctas_per_sm CTAs per SM take work items from a device-side counter; their
  // event tiles arrive by cp.async.bulk (TMA) into two shared-memory stages.
  // 7 of the 8 CTA slots per SM: the eighth (and, at 32 registers per thread, an eighth of the register file) stays
  // free for the merge / re-zero / peer-reduce kernels that run beside the votes.  Measured (profiles/r2_ab_knobs.md):
  // 6 per SM is 2 % faster with device-resident inputs but 5 % slower through the host-buffer calls, 8 is the reverse
  const size_t resident = (size_t)ctx->sm_count * ctx->vote_ctas_per_sm;
  // work items per packet: whole packets when every resident CTA gets several of them, halves / quarters when a
  // build is short (head of a split upload, a small shard) so that the dynamic queue can still balance the SMs
  uint32_t sub = 0;
  if (ctx->vote_split >= 0) sub = (uint32_t)ctx->vote_split;
  else while (sub < 2 && (n_packets << sub) < 8 * resident) ++sub;
  while (sub > 0 && (EMVS_PACKET_SIZE >> sub) / (kVoteThreads / std::max<uint32_t>(G, 1)) < 8) --sub;   // keep the unrolled event loop whole
  const size_t n_items = n_packets << sub;
  unsigned int* const d_done = ctx->d_work ? ctx->d_work + kMaxWorkCounters : nullptr;
  unsigned int* const d_ms_error = ctx->d_work ? ctx->d_work + 2 * kMaxWorkCounters : nullptr;
#define LAUNCH_VOTE_T(GG, GRID, SMEM, K0, SLAB_PLANES, N_SLABS, DIMZ, QUAD, STRIDE, WC, DONE)                                \
  do {                                                                                                                      \
    if (ctx->hint_red == 2)                                                                                                 \
      k_vote_tma<GG, 2><<<GRID, kVoteThreads, SMEM, st>>>(ctx->d_xy0, d_pk, m->d_depths, K0, SLAB_PLANES, N_SLABS, DIMZ,     \
                                                          (uint32_t)n_items, sub, P, QUAD, STRIDE, m->d_counts, WC, DONE,  \
                                                          ctx->hint_xy0);                                                   \
    else                                                                                                                    \
      k_vote_tma<GG, 0><<<GRID, kVoteThreads, SMEM, st>>>(ctx->d_xy0, d_pk, m->d_depths, K0, SLAB_PLANES, N_SLABS, DIMZ,     \
                                                          (uint32_t)n_items, sub, P, QUAD, STRIDE, m->d_counts, WC, DONE,  \
                                                          ctx->hint_xy0);                                                   \
  } while (0)
#define LAUNCH_VOTE_T_ANY(...)                        \
  switch (G) {                                        \
    case 2: LAUNCH_VOTE_T(2, __VA_ARGS__); break;     \
    case 4: LAUNCH_VOTE_T(4, __VA_ARGS__); break;     \
    case 8: LAUNCH_VOTE_T(8, __VA_ARGS__); break;     \
    case 16: LAUNCH_VOTE_T(16, __VA_ARGS__); break;   \
    default: LAUNCH_VOTE_T(32, __VA_ARGS__); break;   \
  }
  auto profile_begin = [&](cudaEvent_t* pe1) -> int {
    *pe1 = nullptr;
    if (!ctx->profile) return EMVS_OK;
    if (ctx->prof_used + 2 > ctx->prof_events.size()) {
      cudaEvent_t a, b;
      CUDA_TRY(cudaEventCreate(&a));
      CUDA_TRY(cudaEventCreate(&b));
      ctx->prof_events.push_back(a);
      ctx->prof_events.push_back(b);
    }
    cudaEvent_t pe0 = ctx->prof_events[ctx->prof_used];
    *pe1 = ctx->prof_events[ctx->prof_used + 1];
    ctx->prof_used += 2;
    CUDA_TRY(cudaEventRecord(pe0, st));
    return EMVS_OK;
  };
  const size_t slab_f4 = slab_bytes / sizeof(float4);
  if (multislab) {
    // ONE launch votes every slab of this build, slab s into its own scratch buffer; the merges below are queued behind
    // k_wait_count on the slab's completion counter.  The previous build's merges (which re-zero the same buffers) are
    // already ordered before this point on `st` (end-of-build wait below).
    CUDA_TRY(cudaEventRecord(ctx->ev_ms_start, st));              // the counters are zero from here on
    CUDA_TRY(cudaStreamWaitEvent(ctx->aux_stream, ctx->ev_ms_start, 0));
    cudaEvent_t pe1 = nullptr;
    { const int rc = profile_begin(&pe1); if (rc) return rc; }
    const unsigned grid = (unsigned)std::min<size_t>(n_items, resident);
    if (n_items) {
      // a deferred piece publishes nothing: the launch of the piece that merges is ordered after it on the stream
      LAUNCH_VOTE_T_ANY(grid, vote_tma_smem_bytes(slab), 0u, slab, n_slabs, dimZ, ctx->quad_ms, slab_f4, ctx->d_work,
                        defer ? (unsigned int*)nullptr : d_done);
      ctx->launches++;
    }
    ctx->ms_check_pending = true;
    if (pe1) CUDA_TRY(cudaEventRecord(pe1, st));
    ctx->deferred_votes = defer;
  }

  for (uint32_t k0 = 0; k0 < dimZ && !defer; k0 += slab) {
    const uint32_t nk = std::min(slab, dimZ - k0);
    const size_t smem = nk * (sizeof(float4) + sizeof(unsigned int));
    const int b = (overlap && !multislab) ? (int)((k0 / slab) & 1u) : 0;
    cudaStream_t ms = overlap ? ctx->aux_stream : st;   // merge + re-zero stream
    float4* const qbuf = multislab ? ctx->quad_ms + (size_t)(k0 / slab) * slab_f4 : ctx->quad[b];
    if (multislab) {
      // the merge of this slab may start when its last vote has landed: n_items items counted by the vote CTAs
      k_wait_count<<<1, 32, 0, ms>>>(d_done + k0 / slab, (unsigned int)n_items, 20000000000LL, d_ms_error);
      ctx->launches++;
    }
    if (!multislab && overlap && ctx->merge_pending[b]) {             // buffer b must be merged and zero again (slab s-2)
      CUDA_TRY(cudaStreamWaitEvent(st, ctx->ev_merge[b], 0));
      ctx->merge_pending[b] = false;
    }
    cudaEvent_t pe0 = nullptr, pe1 = nullptr;
    if (ctx->profile && !multislab) {
      if (ctx->prof_used + 2 > ctx->prof_events.size()) {
        cudaEvent_t a, b;
        CUDA_TRY(cudaEventCreate(&a));
        CUDA_TRY(cudaEventCreate(&b));
        ctx->prof_events.push_back(a);
        ctx->prof_events.push_back(b);
      }
      pe0 = ctx->prof_events[ctx->prof_used];
      pe1 = ctx->prof_events[ctx->prof_used + 1];
      ctx->prof_used += 2;
      CUDA_TRY(cudaEventRecord(pe0, st));
    }
    const size_t smem_g = smem + (EMVS_PACKET_SIZE + nk) * sizeof(float2);   // + the packet's warped events + prepared reciprocals
    static const bool fastdiv = [] { const char* e = getenv("EMVS_VOTE_FASTDIV"); return e ? atoi(e) != 0 : false; }();
    if (multislab) {
      // voted by the build's single launch above
    } else if (use_tma) {
      const unsigned grid = (unsigned)std::min<size_t>(n_items, resident);
      LAUNCH_VOTE_T_ANY(grid, vote_tma_smem_bytes(nk), k0, nk, 1u, k0 + nk, ctx->quad[b], (size_t)0, ctx->d_work + k0 / slab,
                        (unsigned int*)nullptr);
    } else {
#define LAUNCH_VOTE_GF(GG, FF)                                                                                         \
  k_vote_grouped<GG, FF><<<(unsigned)n_packets, kVoteThreads, smem_g, st>>>(ctx->d_xy0, d_pk, m->d_depths, k0, nk, P, ctx->quad[b], \
                                                                            m->d_counts)
#define LAUNCH_VOTE_G(GG)                  \
  do {                                     \
    if (fastdiv) LAUNCH_VOTE_GF(GG, true); \
    else LAUNCH_VOTE_GF(GG, false);        \
  } while (0)
      switch (G) {
        case 2: LAUNCH_VOTE_G(2); break;
        case 4: LAUNCH_VOTE_G(4); break;
        case 8: LAUNCH_VOTE_G(8); break;
        case 16: LAUNCH_VOTE_G(16); break;
        case 32: LAUNCH_VOTE_G(32); break;
        default:
          k_vote<<<(unsigned)n_packets, kVoteThreads, smem, st>>>(ctx->d_xy0, d_pk, m->d_depths, k0, nk, P, ctx->quad[b], m->d_counts);
      }
#undef LAUNCH_VOTE_G
#undef LAUNCH_VOTE_GF
    }
    if (!multislab) ctx->launches++;
    if (pe1) CUDA_TRY(cudaEventRecord(pe1, st));
    if (overlap && !multislab) {
      CUDA_TRY(cudaEventRecord(ctx->ev_vote[b], st));
      CUDA_TRY(cudaStreamWaitEvent(ms, ctx->ev_vote[b], 0));
    }
    static const bool merge_grouped = [] { const char* e = getenv("EMVS_MERGE_GROUPED"); return e ? atoi(e) != 0 : true; }();
    // timing experiments only (profiles/r2_interference.md): the DSI is WRONG with either of them set
    const bool dbg_skip_merge = ctx->dbg_skip_merge, dbg_skip_zero = ctx->dbg_skip_zero;
    if (dbg_skip_merge) {
    } else if (G > 1 && merge_grouped) {
      const dim3 mg(ceil_div(QW, 256 / G), QH, ceil_div(nk, G));
      float* dst = g->d + (size_t)k0 * dimX * dimY;
#define LAUNCH_MERGE_G(GG) \
  k_merge_quads_grouped<GG><<<mg, 256, 0, ms>>>(qbuf, dst, dimX, dimY, QW, QH, nk, accumulate ? 1 : 0, ctx->hint_dsi)
      switch (G) {
        case 2: LAUNCH_MERGE_G(2); break;
        case 4: LAUNCH_MERGE_G(4); break;
        case 8: LAUNCH_MERGE_G(8); break;
        case 16: LAUNCH_MERGE_G(16); break;
        default: LAUNCH_MERGE_G(32); break;
      }
#undef LAUNCH_MERGE_G
    } else {
      dim3 mb(32, 8, 1), mg(ceil_div(QW, 32), ceil_div(QH, 8), nk);
      k_merge_quads<<<mg, mb, 0, ms>>>(qbuf, g->d + (size_t)k0 * dimX * dimY, dimX, dimY, QW, QH,
                                       accumulate ? 1 : 0, (int)G);
    }
    ctx->launches++;
    {
      // re-zero the merged buffer: a grid-stride kernel whose stores carry the L2 evict_last hint, like the vote REDs that
      // will land on these lines next (without the hints cudaMemsetAsync is the faster way, profiles/r2_l2_hints.md)
      const int zero_ctas = ctx->zero_ctas;
      const size_t n_f4 = (size_t)round_up_g(nk) * QW * QH * 4;
      if (dbg_skip_zero) {
      } else if (zero_ctas > 0) {
        k_zero_f4<<<(unsigned)zero_ctas, 256, 0, ms>>>(qbuf, n_f4, ctx->hint_zero);
        ctx->launches++;
      } else {
        CUDA_TRY(cudaMemsetAsync(qbuf, 0, n_f4 * sizeof(float4), ms));
      }
    }
    if (overlap) {
      CUDA_TRY(cudaEventRecord(ctx->ev_merge[b], ms));
      ctx->merge_pending[b] = true;
    }
    if (peer) {
      // planes [k0, k0+nk) are final on this rank: tell every peer, then sum this rank's row band of those
      // planes over all ranks into the band buffer on the communication stream, under the next slab's votes
      const int rc = peer_reduce_slab(ctx, ex, peer_cam, (uint32_t)(k0 / slab), k0, nk, ms);
      if (rc) return rc;
    }
    if (reduce) {
      // planes [k0, k0+nk) are final on this rank: sum them over the ranks on the communication
      // stream while the next slab is being voted
      const size_t si = k0 / slab;
      while (ctx->slab_events.size() <= si) {
        cudaEvent_t e;
        CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ctx->slab_events.push_back(e);
      }
      CUDA_TRY(cudaEventRecord(ctx->slab_events[si], ms));
      CUDA_TRY(cudaStreamWaitEvent(ctx->comm_stream, ctx->slab_events[si], 0));
      if (ncclAllReduce_checked(nccl, ctx, g->d + (size_t)k0 * dimX * dimY, (size_t)nk * dimX * dimY, ctx->comm_stream))
        return EMVS_ERR_NCCL;
    }
  }
#undef LAUNCH_VOTE_T_ANY
#undef LAUNCH_VOTE_T
  // everything later on `stream` (fusion, download, the next build) sees the finished DSI
  for (int b = 0; b < 2; ++b)
    if (ctx->merge_pending[b]) {
      CUDA_TRY(cudaStreamWaitEvent(st, ctx->ev_merge[b], 0));
      ctx->merge_pending[b] = false;
    }
  // (peer mode: the band reduces keep running on the communication stream under the next camera's votes;
  //  emvs_exchange_fuse_collapse waits for them)
  if (reduce) {
    // the vote counters are final once the last vote kernel ran (recorded by the last slab event)
    if (ncclAllReduceU64_checked(nccl, ctx, m->d_counts, dimZ, ctx->comm_stream)) return EMVS_ERR_NCCL;
    CUDA_TRY(cudaEventRecord(ctx->ev_comm_done, ctx->comm_stream));
    CUDA_TRY(cudaStreamWaitEvent(st, ctx->ev_comm_done, 0));
  }
  if (ctx->mark_consumed) {   // host-buffer build: its packet buffer may be overwritten once these votes are done
    CUDA_TRY(cudaEventRecord(ctx->ev_pk_free[ctx->cur_packets], st));
    ctx->pk_free_recorded[ctx->cur_packets] = true;
  }
  CUDA_TRY(cudaGetLastError());
  return EMVS_OK;
}

template <int METHOD>
int launch_fuse_collapse_n(emvs_context* ctx, const FuseArgs& A, uint32_t n_pix, uint32_t dimZ, const float* d_depths,
                           float* fused, float* conf, void* idx, int idx_bytes, float* depth)
{
  const unsigned blocks = (n_pix + 127) / 128;
  cudaStream_t st = ctx->stream;
  // (A float4-per-thread variant with the Z range split over warps was measured slower than this
  // one-pixel-per-thread sweep — 0.234 vs 0.209 ms for two 640x480x256 volumes, profiles/r1_fuse_collapse.md.)
  // The planes are split over 4 CTAs per pixel tile + a combine pass: 0.196 vs 0.209 ms (EMVS_FC_ZSPLIT overrides).
  const uint32_t n_chunks = (uint32_t)std::max(1, std::min<int>(ctx->fc_zsplit, (int)(dimZ / 16)));   // at least 16 planes per chunk
  float* part_best = nullptr;
  uint32_t* part_k = nullptr;
  uint32_t per_chunk = dimZ;
  if (n_chunks > 1) {
    const int rc = grow(&ctx->d_fc_part, &ctx->fc_part_cap, (size_t)n_chunks * n_pix * 8);
    if (rc) return rc;
    part_best = (float*)ctx->d_fc_part;
    part_k = (uint32_t*)((char*)ctx->d_fc_part + (size_t)n_chunks * n_pix * 4);
    per_chunk = (dimZ + n_chunks - 1) / n_chunks;
  }
  const uint32_t used_chunks = n_chunks > 1 ? (dimZ + per_chunk - 1) / per_chunk : 1;
  // EMVS_FC_V4=1: four pixels per thread (16-byte loads); needs 16-byte aligned planes
  const bool v4 = ctx->fc_v4 && n_chunks > 1 && n_pix % 4 == 0;
  const unsigned blocks4 = (n_pix / 4 + 127) / 128;
#define LAUNCH_M(M, N)                                                                                                     \
  do {                                                                                                                     \
    if (v4) {                                                                                                              \
      k_fuse_collapse_zsplit_v4<M, N><<<dim3(blocks4, used_chunks), 128, 0, st>>>(A, n_pix, dimZ, per_chunk, fused, part_best, part_k); \
      k_fc_combine<<<(n_pix + 255) / 256, 256, 0, st>>>(part_best, part_k, used_chunks, n_pix, d_depths, conf, idx, idx_bytes, depth); \
      ctx->launches++;                                                                                                     \
    } else if (n_chunks > 1) {                                                                                             \
      k_fuse_collapse_zsplit<M, N><<<dim3(blocks, used_chunks), 128, 0, st>>>(A, n_pix, dimZ, per_chunk, fused, part_best, part_k); \
      k_fc_combine<<<(n_pix + 255) / 256, 256, 0, st>>>(part_best, part_k, used_chunks, n_pix, d_depths, conf, idx, idx_bytes, depth); \
      ctx->launches++;                                                                                                     \
    } else {                                                                                                               \
      k_fuse_collapse<M, N><<<blocks, 128, 0, st>>>(A, n_pix, dimZ, d_depths, fused, conf, idx, idx_bytes, depth);          \
    }                                                                                                                      \
  } while (0)
#define LAUNCH(N) LAUNCH_M(METHOD, N)
  switch (A.n) {
    case 1: LAUNCH_M(EMVS_FUSE_MAX, 1); break;
    case 2: LAUNCH(2); break;
    case 3: LAUNCH(3); break;
    case 4: LAUNCH(4); break;
    case 5: LAUNCH(5); break;
    case 6: LAUNCH(6); break;
    case 7: LAUNCH(7); break;
    case 8: LAUNCH(8); break;
    default: set_error("fuse_collapse: need 1..8 grids"); return EMVS_ERR_INVALID;
  }
#undef LAUNCH
#undef LAUNCH_M
  ctx->launches++;
  CUDA_TRY(cudaGetLastError());
  return EMVS_OK;
}

int launch_fuse_collapse(emvs_context* ctx, const FuseArgs& A, uint32_t n_pix, uint32_t dimZ, const float* d_depths,
                         float* fused, float* conf, void* idx, int idx_bytes, float* depth)
{
#define CASE(M) case M: return launch_fuse_collapse_n<M>(ctx, A, n_pix, dimZ, d_depths, fused, conf, idx, idx_bytes, depth)
  switch (A.method) {
    CASE(EMVS_FUSE_MIN);
    CASE(EMVS_FUSE_HM);
    CASE(EMVS_FUSE_GM);
    CASE(EMVS_FUSE_AM);
    CASE(EMVS_FUSE_RMS);
    CASE(EMVS_FUSE_MAX);
    default: set_error("Improper fusion method selected (%d)", A.method); return EMVS_ERR_INVALID;
  }
#undef CASE
}

// Collapse (optionally fused with an n-ary fusion) into host buffers.
int collapse_to_host(emvs_context* ctx, const FuseArgs& A, uint32_t dimX, uint32_t dimY, uint32_t dimZ,
                     const float* h_depths, float* fused, float* conf, void* idx, float* depth)
{
  REQUIRE(conf && idx, EMVS_ERR_INVALID, "collapse: conf and idx must not be NULL");
  REQUIRE(!depth || h_depths, EMVS_ERR_INVALID, "collapse: depth output needs the depth table");
  const uint32_t n_pix = dimX * dimY;
  const size_t idx_sz = dimZ <= 256 ? 1 : 2;
  const size_t off_depth = (size_t)n_pix * 4, off_idx = (size_t)n_pix * 8;
  const size_t off_tab = ((size_t)n_pix * 10 + 15) & ~(size_t)15;
  const size_t total = off_tab + (size_t)dimZ * 4;
  int rc = grow(&ctx->d_out, &ctx->out_cap, total);
  if (rc) return rc;
  char* base = (char*)ctx->d_out;
  float* d_conf = (float*)base;
  float* d_depth = depth ? (float*)(base + off_depth) : nullptr;
  void* d_idx = base + off_idx;
  float* d_tab = (float*)(base + off_tab);
  cudaStream_t st = ctx->stream;
  if (depth) CUDA_TRY(cudaMemcpyAsync(d_tab, h_depths, (size_t)dimZ * 4, cudaMemcpyHostToDevice, st));
  rc = launch_fuse_collapse(ctx, A, n_pix, dimZ, d_tab, fused, d_conf, d_idx, (int)idx_sz, d_depth);
  if (rc) return rc;
  CUDA_TRY(cudaMemcpyAsync(conf, d_conf, (size_t)n_pix * 4, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(idx, d_idx, (size_t)n_pix * idx_sz, cudaMemcpyDeviceToHost, st));
  if (depth) CUDA_TRY(cudaMemcpyAsync(depth, d_depth, (size_t)n_pix * 4, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  return EMVS_OK;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// C-ABI
// ---------------------------------------------------------------------------------------------
extern "C" {

int emvs_abi_version(void) { return EMVS_ABI_VERSION; }
const char* emvs_last_error(void) { return g_err; }

int emvs_context_create(int device, emvs_context** out)
{
  REQUIRE(out, EMVS_ERR_INVALID, "context_create: out is NULL");
  *out = nullptr;
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) {
    cudaGetLastError();
    set_error("no CUDA device available: this library has no CPU fallback");
    return EMVS_ERR_CUDA;
  }
  REQUIRE(device >= 0 && device < n_dev, EMVS_ERR_INVALID, "context_create: bad device index");
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_error("device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
    return EMVS_ERR_CUDA;
  }
  DeviceGuard guard(device);
  REQUIRE(guard.ok, EMVS_ERR_CUDA, "cudaSetDevice failed");
  emvs_context* ctx = new (std::nothrow) emvs_context;
  REQUIRE(ctx, EMVS_ERR_INVALID, "out of host memory");
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  cudaError_t e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->aux_stream, cudaStreamNonBlocking);
  for (int b = 0; b < 2 && e == cudaSuccess; ++b) {
    e = cudaEventCreateWithFlags(&ctx->ev_vote[b], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_merge[b], cudaEventDisableTiming);
  }
  if (const char* env = getenv("EMVS_OVERLAP")) ctx->overlap = atoi(env) != 0;
  if (const char* env = getenv("EMVS_VOTE_KERNEL")) ctx->vote_tma = strcmp(env, "classic") != 0;
  if (const char* env = getenv("EMVS_VOTE_CTAS_PER_SM")) ctx->vote_ctas_per_sm = (uint32_t)std::min(8, std::max(1, atoi(env)));
  ctx->vote_split = std::min(2, std::max(-1, env_int("EMVS_VOTE_SPLIT", ctx->vote_split)));
  ctx->vote_multislab = env_int("EMVS_VOTE_MULTISLAB", ctx->vote_multislab ? 1 : 0) != 0;
  ctx->multislab_budget = (size_t)std::max(0, env_int("EMVS_MULTISLAB_BUDGET_MB", (int)(ctx->multislab_budget >> 20))) << 20;
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_ms_start, cudaEventDisableTiming);
  ctx->zero_ctas = std::max(0, env_int("EMVS_ZERO_CTAS", ctx->zero_ctas));
  ctx->peer_reduce_ctas = std::max(0, env_int("EMVS_PEER_REDUCE_CTAS", ctx->peer_reduce_ctas));
  ctx->fc_v4 = env_int("EMVS_FC_V4", ctx->fc_v4 ? 1 : 0) != 0;
  ctx->fc_zsplit = std::max(1, env_int("EMVS_FC_ZSPLIT", ctx->fc_zsplit));
  ctx->dbg_skip_merge = env_int("EMVS_DEBUG_SKIP_MERGE", 0) != 0;
  ctx->dbg_skip_zero = env_int("EMVS_DEBUG_SKIP_ZERO", 0) != 0;
  ctx->hint_xy0 = std::min(2, std::max(0, env_int("EMVS_HINT_XY0", ctx->hint_xy0)));
  ctx->hint_red = env_int("EMVS_HINT_RED", ctx->hint_red) == 2 ? 2 : 0;
  ctx->hint_dsi = env_int("EMVS_HINT_DSI", ctx->hint_dsi) != 0 ? 1 : 0;
  ctx->hint_zero = std::min(2, std::max(0, env_int("EMVS_HINT_ZERO", ctx->hint_zero)));
  if (const char* env = getenv("EMVS_UPLOAD_SPLIT")) ctx->split_percent = (uint32_t)std::min(90, std::max(0, atoi(env)));
  ctx->split_pieces = std::min(4, std::max(2, env_int("EMVS_UPLOAD_PIECES", ctx->split_pieces)));
  ctx->split_defer_merge = env_int("EMVS_UPLOAD_DEFER_MERGE", ctx->split_defer_merge ? 1 : 0) != 0;
  ctx->defer_percent = (uint32_t)std::min(60, std::max(1, env_int("EMVS_UPLOAD_DEFER_SPLIT", (int)ctx->defer_percent)));
  ctx->defer_pieces = std::min(4, std::max(2, env_int("EMVS_UPLOAD_DEFER_PIECES", ctx->defer_pieces)));
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_copied, cudaEventDisableTiming);
  for (int b = 0; b < 2 && e == cudaSuccess; ++b) e = cudaEventCreateWithFlags(&ctx->ev_consumed[b], cudaEventDisableTiming);
  for (int b = 0; b < 2 && e == cudaSuccess; ++b) e = cudaEventCreateWithFlags(&ctx->ev_pk_free[b], cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_prefetched, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaMalloc((void**)&ctx->d_partial, sizeof(double) * 1025);
  if (e != cudaSuccess) {
    set_error("context_create: %s", cudaGetErrorString(e));
    delete ctx;
    return EMVS_ERR_CUDA;
  }
  *out = ctx;
  return EMVS_OK;
}

static void context_release(emvs_context* ctx);

int emvs_context_destroy(emvs_context* ctx)
{
  if (!ctx) return EMVS_OK;
  context_release(ctx);
  return EMVS_OK;
}

static void context_retain(emvs_context* ctx) { ctx->refs.fetch_add(1); }

static void context_release(emvs_context* ctx)
{
  if (ctx->refs.fetch_sub(1) != 1) return;
  DeviceGuard guard(ctx->device);
  if (ctx->comm) emvs_comm_destroy(ctx);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->aux_stream) cudaStreamSynchronize(ctx->aux_stream);
  cudaFree(ctx->quad[0]);
  cudaFree(ctx->quad[1]);
  for (int b = 0; b < 2; ++b) {
    if (ctx->ev_vote[b]) cudaEventDestroy(ctx->ev_vote[b]);
    if (ctx->ev_merge[b]) cudaEventDestroy(ctx->ev_merge[b]);
  }
  if (ctx->aux_stream) cudaStreamDestroy(ctx->aux_stream);
  if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
  cudaFree(ctx->d_events[0]);
  cudaFree(ctx->d_events[1]);
  cudaFree(ctx->d_packets[0]);
  cudaFree(ctx->d_packets[1]);
  for (cudaEvent_t e : ctx->slab_events) cudaEventDestroy(e);
  if (ctx->ev_comm_done) cudaEventDestroy(ctx->ev_comm_done);
  if (ctx->comm_stream) cudaStreamDestroy(ctx->comm_stream);
  if (ctx->ev_copied) cudaEventDestroy(ctx->ev_copied);
  for (int b = 0; b < 2; ++b)
    if (ctx->ev_consumed[b]) cudaEventDestroy(ctx->ev_consumed[b]);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  cudaFree(ctx->d_xy0);
  cudaFree(ctx->d_work);
  cudaFree(ctx->quad_ms);
  if (ctx->ev_ms_start) cudaEventDestroy(ctx->ev_ms_start);
  cudaFree(ctx->d_out);
  cudaFree(ctx->d_fc_part);
  cudaFree(ctx->d_partial);
  if (ctx->h_packets) cudaFreeHost(ctx->h_packets);
  if (ctx->h_packets_pf) cudaFreeHost(ctx->h_packets_pf);
  if (ctx->ev_prefetched) cudaEventDestroy(ctx->ev_prefetched);
  for (int b = 0; b < 2; ++b)
    if (ctx->ev_pk_free[b]) cudaEventDestroy(ctx->ev_pk_free[b]);
  for (cudaEvent_t e : ctx->prof_events) cudaEventDestroy(e);
  cudaStreamDestroy(ctx->stream);
  delete ctx;
}

int emvs_context_sync(emvs_context* ctx)
{
  REQUIRE(ctx, EMVS_ERR_INVALID, "context is NULL");
  DeviceGuard guard(ctx->device);
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  if (ctx->ms_check_pending && ctx->d_work) {
    // k_wait_count gives up after ~10 s instead of hanging the device; a merge that ran on an unfinished slab is an error
    ctx->ms_check_pending = false;
    unsigned int err = 0;
    CUDA_TRY(cudaMemcpy(&err, ctx->d_work + 2 * kMaxWorkCounters, sizeof(err), cudaMemcpyDeviceToHost));
    if (err) {
      CUDA_TRY(cudaMemset(ctx->d_work + 2 * kMaxWorkCounters, 0, sizeof(err)));
      set_error("a slab's vote completion counter timed out: the DSIs built since the last emvs_context_sync are incomplete");
      return EMVS_ERR_CUDA;
    }
  }
  return EMVS_OK;
}

int emvs_context_set_slab(emvs_context* ctx, uint32_t planes)
{
  REQUIRE(ctx, EMVS_ERR_INVALID, "context is NULL");
  ctx->slab_override = planes;
  return EMVS_OK;
}

int emvs_context_set_upload_split(emvs_context* ctx, uint32_t percent, uint64_t min_events)
{
  REQUIRE(ctx, EMVS_ERR_INVALID, "context is NULL");
  REQUIRE(percent <= 90, EMVS_ERR_INVALID, "set_upload_split: percent must be 0..90");
  ctx->split_percent = percent;
  ctx->split_min_events = (size_t)min_events;
  return EMVS_OK;
}

int emvs_selftest_division(emvs_context* ctx, uint64_t n_pairs, uint32_t seed, uint64_t* mismatches)
{
  REQUIRE(ctx && mismatches, EMVS_ERR_INVALID, "selftest_division: NULL argument");
  DeviceGuard guard(ctx->device);
  unsigned long long* d_bad = reinterpret_cast<unsigned long long*>(ctx->d_partial);   // scratch of the context
  CUDA_TRY(cudaMemsetAsync(d_bad, 0, sizeof(unsigned long long), ctx->stream));
  const unsigned blocks = (unsigned)ctx->sm_count * 8u, threads = 256u;
  const uint64_t per_thread = std::max<uint64_t>(1, n_pairs / ((uint64_t)blocks * threads));
  REQUIRE(per_thread <= 0x3fffffffull, EMVS_ERR_INVALID, "selftest_division: n_pairs too large");
  k_selftest_division<<<blocks, threads, 0, ctx->stream>>>((uint32_t)per_thread, seed, d_bad);
  ctx->launches++;
  unsigned long long bad = 0;
  CUDA_TRY(cudaMemcpyAsync(&bad, d_bad, sizeof bad, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  *mismatches = bad;
  return EMVS_OK;
}

int emvs_context_launch_count(emvs_context* ctx, uint64_t* out)
{
  REQUIRE(ctx && out, EMVS_ERR_INVALID, "launch_count: NULL argument");
  *out = ctx->launches;
  return EMVS_OK;
}

int emvs_context_profile_vote(emvs_context* ctx, int enable)
{
  REQUIRE(ctx, EMVS_ERR_INVALID, "context is NULL");
  ctx->profile = enable != 0;
  ctx->prof_used = 0;
  return EMVS_OK;
}

int emvs_context_vote_time(emvs_context* ctx, double* total_ms, uint64_t* n_launches)
{
  REQUIRE(ctx && total_ms && n_launches, EMVS_ERR_INVALID, "vote_time: NULL argument");
  DeviceGuard guard(ctx->device);
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  double sum = 0.;
  for (size_t i = 0; i + 1 < ctx->prof_used; i += 2) {
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, ctx->prof_events[i], ctx->prof_events[i + 1]));
    sum += ms;
  }
  *total_ms = sum;
  *n_launches = ctx->prof_used / 2;
  return EMVS_OK;
}

int emvs_context_stream(emvs_context* ctx, void** out_stream)
{
  REQUIRE(ctx && out_stream, EMVS_ERR_INVALID, "context_stream: NULL argument");
  *out_stream = (void*)ctx->stream;
  return EMVS_OK;
}

struct emvs_timer {
  emvs_context* ctx;
  cudaEvent_t a, b;
};

int emvs_timer_create(emvs_context* ctx, emvs_timer** out)
{
  REQUIRE(ctx && out, EMVS_ERR_INVALID, "timer_create: NULL argument");
  DeviceGuard guard(ctx->device);
  emvs_timer* t = new (std::nothrow) emvs_timer;
  REQUIRE(t, EMVS_ERR_INVALID, "out of host memory");
  t->ctx = ctx;
  context_retain(ctx);
  CUDA_TRY(cudaEventCreate(&t->a));
  CUDA_TRY(cudaEventCreate(&t->b));
  *out = t;
  return EMVS_OK;
}

int emvs_timer_destroy(emvs_timer* t)
{
  if (!t) return EMVS_OK;
  cudaEventDestroy(t->a);
  cudaEventDestroy(t->b);
  context_release(t->ctx);
  delete t;
  return EMVS_OK;
}

int emvs_timer_start(emvs_timer* t)
{
  REQUIRE(t, EMVS_ERR_INVALID, "timer is NULL");
  DeviceGuard guard(t->ctx->device);
  CUDA_TRY(cudaEventRecord(t->a, t->ctx->stream));
  return EMVS_OK;
}

int emvs_timer_stop(emvs_timer* t)
{
  REQUIRE(t, EMVS_ERR_INVALID, "timer is NULL");
  DeviceGuard guard(t->ctx->device);
  CUDA_TRY(cudaEventRecord(t->b, t->ctx->stream));
  return EMVS_OK;
}

int emvs_timer_elapsed_ms(emvs_timer* t, float* ms)
{
  REQUIRE(t && ms, EMVS_ERR_INVALID, "timer_elapsed_ms: NULL argument");
  DeviceGuard guard(t->ctx->device);
  CUDA_TRY(cudaEventSynchronize(t->b));
  CUDA_TRY(cudaEventElapsedTime(ms, t->a, t->b));
  return EMVS_OK;
}

int emvs_host_alloc(size_t bytes, void** out)
{
  REQUIRE(out, EMVS_ERR_INVALID, "host_alloc: out is NULL");
  CUDA_TRY(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault));
  return EMVS_OK;
}

int emvs_host_free(void* p)
{
  if (p) CUDA_TRY(cudaFreeHost(p));
  return EMVS_OK;
}

// ---- host geometry ---------------------------------------------------------------------------
static int check_shape(const emvs_shape* s)
{
  REQUIRE(s, EMVS_ERR_INVALID, "shape is NULL");
  REQUIRE(s->dimZ >= 1, EMVS_ERR_INVALID, "shape: dimZ must be >= 1");
  REQUIRE(s->min_depth > 0.f, EMVS_ERR_INVALID, "shape: min_depth must be > 0");             // MAP:210
  REQUIRE(s->max_depth > s->min_depth, EMVS_ERR_INVALID, "shape: max_depth must exceed min_depth");  // MAP:211
  return EMVS_OK;
}

int emvs_depth_vector(const emvs_shape* shape, float* out)
{
  int rc = check_shape(shape);
  if (rc) return rc;
  REQUIRE(out, EMVS_ERR_INVALID, "depth_vector: out is NULL");
  host_depth_vector(*shape, out);
  return EMVS_OK;
}

int emvs_virtual_camera(const emvs_camera* cam, const emvs_shape* shape, float out[4])
{
  REQUIRE(cam && shape && out, EMVS_ERR_INVALID, "virtual_camera: NULL argument");
  host_virtual_camera(*cam, *shape, out);
  return EMVS_OK;
}

int emvs_rectify_lut(int model, const double K[9], const double* D, int n_d, const double R[9], const double P[12],
                     uint32_t width, uint32_t height, float* out)
{
  REQUIRE(K && R && P && out && (D || n_d == 0), EMVS_ERR_INVALID, "rectify_lut: NULL argument");
  REQUIRE(width && height && n_d >= 0 && n_d <= 14, EMVS_ERR_INVALID, "rectify_lut: bad size");
  REQUIRE(host_rectify_lut(model, K, D, n_d, R, P, width, height, out) == 0, EMVS_ERR_INVALID,
          "Distortion model not set properly!");   // MAP:293
  return EMVS_OK;
}

int emvs_trajectory_pose_at(const emvs_stamped_pose* traj, size_t n, uint32_t sec, uint32_t nsec, emvs_pose* out,
                            int* found)
{
  REQUIRE(traj && out && found, EMVS_ERR_INVALID, "pose_at: NULL argument");
  REQUIRE(n >= 2, EMVS_ERR_INVALID, "At least two poses need to be provided");  // TRJ:89
  *found = host_pose_at(traj, n, sec, nsec, out) ? 1 : 0;
  return EMVS_OK;
}

int emvs_pose_compose(const emvs_pose* a, const emvs_pose* b, emvs_pose* out)
{
  REQUIRE(a && b && out, EMVS_ERR_INVALID, "pose_compose: NULL argument");
  host_pose_compose(*a, *b, out);
  return EMVS_OK;
}

int emvs_pose_inverse(const emvs_pose* a, emvs_pose* out)
{
  REQUIRE(a && out, EMVS_ERR_INVALID, "pose_inverse: NULL argument");
  host_pose_inverse(*a, out);
  return EMVS_OK;
}

int emvs_packetize(const emvs_event* events, size_t n_events, const emvs_stamped_pose* traj, size_t n_poses,
                   const emvs_pose* T_rv_w, const emvs_camera* cam, const float virt[4], float z0, emvs_packet* out,
                   size_t max_packets, size_t* n_packets)
{
  REQUIRE(events && traj && T_rv_w && cam && virt && n_packets, EMVS_ERR_INVALID, "packetize: NULL argument");
  REQUIRE(n_poses >= 2, EMVS_ERR_INVALID, "At least two poses need to be provided");
  *n_packets = 0;
  if (n_events < EMVS_PACKET_SIZE) {
    set_error("Number of events (%zu) < packet size (%d)", n_events, EMVS_PACKET_SIZE);
    return EMVS_ERR_TOO_FEW;
  }
  REQUIRE(out || max_packets == 0, EMVS_ERR_INVALID, "packetize: out is NULL");
  bool truncated = false;
  *n_packets = host_packetize(times_of(events), n_events, traj, n_poses, *T_rv_w, *cam, virt, z0, out, max_packets, &truncated);
  REQUIRE(!truncated, EMVS_ERR_INVALID, "packetize: max_packets is too small for this list (n_events / 1024 + 1 always suffices)");
  return EMVS_OK;
}

// ---- Grid3D ----------------------------------------------------------------------------------
int emvs_packetize_range(const emvs_event* events, size_t n_events, const emvs_stamped_pose* traj, size_t n_poses,
                         const emvs_pose* T_rv_w, const emvs_camera* cam, const float virt[4], float z0, size_t* cursor,
                         size_t event_limit, emvs_packet* out, size_t max_packets, size_t* n_packets)
{
  REQUIRE(events && traj && T_rv_w && cam && virt && n_packets && cursor, EMVS_ERR_INVALID, "packetize_range: NULL argument");
  REQUIRE(n_poses >= 2, EMVS_ERR_INVALID, "At least two poses need to be provided");
  REQUIRE(event_limit <= n_events && *cursor <= n_events, EMVS_ERR_INVALID, "packetize_range: cursor / limit past the list");
  REQUIRE(out || max_packets == 0, EMVS_ERR_INVALID, "packetize_range: out is NULL");
  bool truncated = false;
  *n_packets = host_packetize_range(times_of(events), n_events, traj, n_poses, *T_rv_w, *cam, virt, z0, cursor, event_limit, out,
                                    max_packets, &truncated);
  REQUIRE(!truncated, EMVS_ERR_INVALID, "packetize_range: max_packets is too small for the events up to event_limit");
  return EMVS_OK;
}

int emvs_grid_create(emvs_context* ctx, uint32_t dimX, uint32_t dimY, uint32_t dimZ, emvs_grid** out)
{
  REQUIRE(ctx && out, EMVS_ERR_INVALID, "grid_create: NULL argument");
  *out = nullptr;
  REQUIRE(dimX && dimY && dimZ, EMVS_ERR_INVALID, "grid_create: zero dimension");
  REQUIRE((uint64_t)dimX * dimY < (1ull << 31), EMVS_ERR_INVALID, "grid_create: plane too large");
  DeviceGuard guard(ctx->device);
  emvs_grid* g = new (std::nothrow) emvs_grid;
  REQUIRE(g, EMVS_ERR_INVALID, "out of host memory");
  g->ctx = ctx;
  g->dimX = dimX; g->dimY = dimY; g->dimZ = dimZ;
  g->n_cells = (size_t)dimX * dimY * dimZ;
  cudaError_t e = cudaMalloc((void**)&g->d, g->n_cells * sizeof(float));
  if (e == cudaSuccess) e = cudaMemsetAsync(g->d, 0, g->n_cells * sizeof(float), ctx->stream);
  if (e != cudaSuccess) {
    set_error("grid_create: %s", cudaGetErrorString(e));
    delete g;
    return EMVS_ERR_CUDA;
  }
  context_retain(ctx);
  *out = g;
  return EMVS_OK;
}

int emvs_grid_destroy(emvs_grid* g)
{
  if (!g) return EMVS_OK;
  DeviceGuard guard(g->ctx->device);
  cudaStreamSynchronize(g->ctx->stream);
  cudaFree(g->d);
  context_release(g->ctx);
  delete g;
  return EMVS_OK;
}

int emvs_grid_dims(const emvs_grid* g, uint32_t* dimX, uint32_t* dimY, uint32_t* dimZ)
{
  REQUIRE(g, EMVS_ERR_INVALID, "grid is NULL");
  if (dimX) *dimX = g->dimX;
  if (dimY) *dimY = g->dimY;
  if (dimZ) *dimZ = g->dimZ;
  return EMVS_OK;
}

int emvs_grid_reset(emvs_grid* g)
{
  REQUIRE(g, EMVS_ERR_INVALID, "grid is NULL");
  DeviceGuard guard(g->ctx->device);
  CUDA_TRY(cudaMemsetAsync(g->d, 0, g->n_cells * sizeof(float), g->ctx->stream));
  return EMVS_OK;
}

static bool same_dims(const emvs_grid* a, const emvs_grid* b)
{
  return a->dimX == b->dimX && a->dimY == b->dimY && a->dimZ == b->dimZ && a->ctx == b->ctx;
}

int emvs_grid_op(emvs_grid* a, const emvs_grid* b, int op, int n, float eps)
{
  REQUIRE(a, EMVS_ERR_INVALID, "grid_op: a is NULL");
  REQUIRE(op >= EMVS_OP_ADD && op <= EMVS_OP_AM_FROM_SUM, EMVS_ERR_INVALID, "grid_op: unknown op");
  const bool unary = op == EMVS_OP_HM_FROM_SUMINV || op == EMVS_OP_AM_FROM_SUM;
  REQUIRE(unary || b, EMVS_ERR_INVALID, "grid_op: b is NULL");
  REQUIRE(unary || same_dims(a, b), EMVS_ERR_INVALID, "grid_op: grids differ in shape or context");
  emvs_context* ctx = a->ctx;
  DeviceGuard guard(ctx->device);
  const unsigned blocks = (unsigned)std::min<size_t>((a->n_cells + 255) / 256, (size_t)ctx->sm_count * 16);
  k_grid_op<<<blocks, 256, 0, ctx->stream>>>(a->d, unary ? nullptr : b->d, a->n_cells, op, n, eps);
  ctx->launches++;
  CUDA_TRY(cudaGetLastError());
  return EMVS_OK;
}

int emvs_grid_copy(emvs_grid* dst, const emvs_grid* src)
{
  REQUIRE(dst && src, EMVS_ERR_INVALID, "grid_copy: NULL argument");
  REQUIRE(same_dims(dst, src), EMVS_ERR_INVALID, "grid_copy: grids differ in shape or context");
  DeviceGuard guard(dst->ctx->device);
  CUDA_TRY(cudaMemcpyAsync(dst->d, src->d, dst->n_cells * sizeof(float), cudaMemcpyDeviceToDevice, dst->ctx->stream));
  return EMVS_OK;
}

int emvs_grid_download(const emvs_grid* g, float* host_out)
{
  REQUIRE(g && host_out, EMVS_ERR_INVALID, "grid_download: NULL argument");
  DeviceGuard guard(g->ctx->device);
  CUDA_TRY(cudaMemcpyAsync(host_out, g->d, g->n_cells * sizeof(float), cudaMemcpyDeviceToHost, g->ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(g->ctx->stream));
  return EMVS_OK;
}

int emvs_grid_upload(emvs_grid* g, const float* host_in)
{
  REQUIRE(g && host_in, EMVS_ERR_INVALID, "grid_upload: NULL argument");
  DeviceGuard guard(g->ctx->device);
  CUDA_TRY(cudaMemcpyAsync(g->d, host_in, g->n_cells * sizeof(float), cudaMemcpyHostToDevice, g->ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(g->ctx->stream));
  return EMVS_OK;
}

int emvs_grid_mean_square(const emvs_grid* g, double* out)
{
  REQUIRE(g && out, EMVS_ERR_INVALID, "grid_mean_square: NULL argument");
  emvs_context* ctx = g->ctx;
  DeviceGuard guard(ctx->device);
  const int blocks = (int)std::min<size_t>((g->n_cells + 255) / 256, 1024);
  k_sumsq_partial<<<blocks, 256, 0, ctx->stream>>>(g->d, g->n_cells, ctx->d_partial);
  k_sumsq_final<<<1, 256, 0, ctx->stream>>>(ctx->d_partial, blocks, ctx->d_partial + 1024);
  ctx->launches += 2;
  double sum = 0.;
  CUDA_TRY(cudaMemcpyAsync(&sum, ctx->d_partial + 1024, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  *out = sum / (double)g->n_cells;
  return EMVS_OK;
}

int emvs_grid_collapse_max(const emvs_grid* g, const float* depths, float* conf, void* idx, float* depth)
{
  REQUIRE(g, EMVS_ERR_INVALID, "grid is NULL");
  DeviceGuard guard(g->ctx->device);
  FuseArgs A{};
  A.g[0] = g->d;
  A.n = 1;
  A.method = EMVS_FUSE_MAX;
  return collapse_to_host(g->ctx, A, g->dimX, g->dimY, g->dimZ, depths, nullptr, conf, idx, depth);
}

int emvs_fuse_collapse(emvs_grid* const* grids, int n, int method, const float* depths, emvs_grid* fused_out,
                       float* conf, void* idx, float* depth)
{
  REQUIRE(grids && n >= 1 && n <= kMaxFuse, EMVS_ERR_INVALID, "fuse_collapse: need 1..8 grids");
  REQUIRE(method >= EMVS_FUSE_MIN && method <= EMVS_FUSE_MAX, EMVS_ERR_INVALID, "Improper fusion method selected");
  FuseArgs A{};
  A.n = n;
  A.method = method;
  for (int i = 0; i < n; ++i) {
    REQUIRE(grids[i], EMVS_ERR_INVALID, "fuse_collapse: NULL grid");
    REQUIRE(same_dims(grids[0], grids[i]), EMVS_ERR_INVALID, "fuse_collapse: grids differ in shape or context");
    A.g[i] = grids[i]->d;
  }
  REQUIRE(!fused_out || same_dims(grids[0], fused_out), EMVS_ERR_INVALID, "fuse_collapse: fused_out shape mismatch");
  const emvs_grid* g = grids[0];
  DeviceGuard guard(g->ctx->device);
  return collapse_to_host(g->ctx, A, g->dimX, g->dimY, g->dimZ, depths, fused_out ? fused_out->d : nullptr, conf, idx,
                          depth);
}

int emvs_fuse_collapse_device(emvs_grid* const* grids, int n, int method, const float* d_depths, emvs_grid* fused_out,
                              float* d_conf, void* d_idx, float* d_depth)
{
  REQUIRE(grids && n >= 1 && n <= kMaxFuse, EMVS_ERR_INVALID, "fuse_collapse: need 1..8 grids");
  REQUIRE(method >= EMVS_FUSE_MIN && method <= EMVS_FUSE_MAX, EMVS_ERR_INVALID, "Improper fusion method selected");
  REQUIRE(d_conf && d_idx, EMVS_ERR_INVALID, "fuse_collapse_device: conf and idx must not be NULL");
  REQUIRE(!d_depth || d_depths, EMVS_ERR_INVALID, "fuse_collapse_device: depth output needs the depth table");
  FuseArgs A{};
  A.n = n;
  A.method = method;
  for (int i = 0; i < n; ++i) {
    REQUIRE(grids[i], EMVS_ERR_INVALID, "fuse_collapse: NULL grid");
    REQUIRE(same_dims(grids[0], grids[i]), EMVS_ERR_INVALID, "fuse_collapse: grids differ in shape or context");
    A.g[i] = grids[i]->d;
  }
  REQUIRE(!fused_out || same_dims(grids[0], fused_out), EMVS_ERR_INVALID, "fuse_collapse: fused_out shape mismatch");
  const emvs_grid* g = grids[0];
  DeviceGuard guard(g->ctx->device);
  return launch_fuse_collapse(g->ctx, A, g->dimX * g->dimY, g->dimZ, d_depths, fused_out ? fused_out->d : nullptr,
                              d_conf, d_idx, g->dimZ <= 256 ? 1 : 2, d_depth);
}

// ---- depth-map post-processing ------------------------------------------------------------------
static int check_post_options(const emvs_depthmap_options* o)
{
  REQUIRE(o, EMVS_ERR_INVALID, "depth map options are NULL");
  REQUIRE(o->adaptive_threshold_kernel_size == 3 || o->adaptive_threshold_kernel_size == 5 ||
              o->adaptive_threshold_kernel_size == 7,
          EMVS_ERR_INVALID, "adaptive_threshold_kernel_size must be 3, 5 or 7 (the dyadic sigma=0 Gaussian kernels)");
  REQUIRE(o->median_filter_size >= 1 && (o->median_filter_size % 2) == 1, EMVS_ERR_INVALID,
          "median_filter_size must be odd");   // CHECK_EQ(patch_size % 2, 1), median_filtering.cpp:43
  return EMVS_OK;
}

// Device buffers inside ctx->d_out, laid out by post_layout(); conf and idx are already there.
struct PostBuffers {
  float* conf; float* depth; uint8_t* idx; uint8_t* conf8; uint8_t* mask; uint8_t* idx_f; float* tab;
  float2* partial; PostScale* scale;
  size_t total;
};

static PostBuffers post_layout(char* base, size_t n_pix, size_t n_depths)
{
  auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
  PostBuffers b;
  size_t o = 0;
  b.conf = (float*)(base + o); o = up(o + n_pix * 4);
  b.depth = (float*)(base + o); o = up(o + n_pix * 4);
  b.idx = (uint8_t*)(base + o); o = up(o + n_pix * 2);
  b.conf8 = (uint8_t*)(base + o); o = up(o + n_pix);
  b.mask = (uint8_t*)(base + o); o = up(o + n_pix);
  b.idx_f = (uint8_t*)(base + o); o = up(o + n_pix);
  b.tab = (float*)(base + o); o = up(o + n_depths * 4);
  b.partial = (float2*)(base + o); o = up(o + 256 * sizeof(float2));
  b.scale = (PostScale*)(base + o); o = up(o + sizeof(PostScale));
  b.total = o;
  return b;
}

static int run_post(emvs_context* ctx, const PostBuffers& b, uint32_t rows, uint32_t cols, const emvs_depthmap_options* opt)
{
  cudaStream_t st = ctx->stream;
  const uint32_t n_pix = rows * cols;
  const int nb = (int)std::min<uint32_t>((n_pix + 255) / 256, 256);
  k_post_minmax<<<nb, 256, 0, st>>>(b.conf, n_pix, (float)opt->max_confidence, b.partial);
  k_post_minmax_final<<<1, 256, 0, st>>>(b.partial, nb, b.scale);
  k_post_to_u8<<<(n_pix + 255) / 256, 256, 0, st>>>(b.conf, n_pix, b.scale, b.conf8);
  const dim3 tb(32, 8), tg((cols + 31) / 32, (rows + 7) / 8);
  const int idelta = (int)std::ceil(-opt->adaptive_threshold_c);   // cvCeil(delta) with delta = -c (THRESH_BINARY)
  switch (opt->adaptive_threshold_kernel_size) {
    case 3: k_post_adaptive_threshold<3><<<tg, tb, 0, st>>>(b.conf8, (int)rows, (int)cols, idelta, b.mask); break;
    case 5: k_post_adaptive_threshold<5><<<tg, tb, 0, st>>>(b.conf8, (int)rows, (int)cols, idelta, b.mask); break;
    default: k_post_adaptive_threshold<7><<<tg, tb, 0, st>>>(b.conf8, (int)rows, (int)cols, idelta, b.mask); break;
  }
  k_post_masked_median<<<tg, tb, 0, st>>>(b.idx, b.mask, (int)rows, (int)cols, opt->median_filter_size / 2, b.idx_f);
  const int border = std::max(opt->adaptive_threshold_kernel_size / 2, 1);
  k_post_finalize<<<tg, tb, 0, st>>>(b.mask, b.idx_f, (int)rows, (int)cols, border, b.tab, b.depth);
  ctx->launches += 6;
  CUDA_TRY(cudaGetLastError());
  return EMVS_OK;
}

static int post_download(emvs_context* ctx, const PostBuffers& b, size_t n_pix, float* depth_map, float* confidence_map,
                         uint8_t* mask, uint8_t* idx_filtered, uint8_t* conf8)
{
  cudaStream_t st = ctx->stream;
  CUDA_TRY(cudaMemcpyAsync(depth_map, b.depth, n_pix * 4, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(confidence_map, b.conf, n_pix * 4, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(mask, b.mask, n_pix, cudaMemcpyDeviceToHost, st));
  if (idx_filtered) CUDA_TRY(cudaMemcpyAsync(idx_filtered, b.idx_f, n_pix, cudaMemcpyDeviceToHost, st));
  if (conf8) CUDA_TRY(cudaMemcpyAsync(conf8, b.conf8, n_pix, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  return EMVS_OK;
}

int emvs_depth_map_from_dsi(emvs_grid* const* grids, int n, int method, const float* depths,
                            const emvs_depthmap_options* opt, float* depth_map, float* confidence_map, uint8_t* mask,
                            uint8_t* idx_filtered)
{
  REQUIRE(grids && n >= 1 && n <= kMaxFuse, EMVS_ERR_INVALID, "depth_map_from_dsi: need 1..8 grids");
  REQUIRE(n == 1 || (method >= EMVS_FUSE_MIN && method <= EMVS_FUSE_MAX), EMVS_ERR_INVALID, "Improper fusion method selected");
  REQUIRE(depths && depth_map && confidence_map && mask, EMVS_ERR_INVALID, "depth_map_from_dsi: NULL argument");
  int rc = check_post_options(opt);
  if (rc) return rc;
  FuseArgs A{};
  A.n = n;
  A.method = n == 1 ? EMVS_FUSE_MAX : method;
  for (int i = 0; i < n; ++i) {
    REQUIRE(grids[i], EMVS_ERR_INVALID, "depth_map_from_dsi: NULL grid");
    REQUIRE(same_dims(grids[0], grids[i]), EMVS_ERR_INVALID, "depth_map_from_dsi: grids differ in shape or context");
    A.g[i] = grids[i]->d;
  }
  const emvs_grid* g = grids[0];
  REQUIRE(g->dimZ <= 256, EMVS_ERR_INVALID, "depth_map_from_dsi: dimZ > 256 (8-bit depth indices, main.cpp:156)");
  emvs_context* ctx = g->ctx;
  DeviceGuard guard(ctx->device);
  const size_t n_pix = (size_t)g->dimX * g->dimY;
  const size_t need = post_layout(nullptr, n_pix, g->dimZ).total;
  rc = grow(&ctx->d_out, &ctx->out_cap, need);
  if (rc) return rc;
  const PostBuffers b = post_layout((char*)ctx->d_out, n_pix, g->dimZ);
  CUDA_TRY(cudaMemcpyAsync(b.tab, depths, (size_t)g->dimZ * 4, cudaMemcpyHostToDevice, ctx->stream));
  rc = launch_fuse_collapse(ctx, A, (uint32_t)n_pix, g->dimZ, b.tab, nullptr, b.conf, b.idx, 1, nullptr);
  if (rc) return rc;
  rc = run_post(ctx, b, g->dimY, g->dimX, opt);
  if (rc) return rc;
  return post_download(ctx, b, n_pix, depth_map, confidence_map, mask, idx_filtered, nullptr);
}

int emvs_depth_map_postprocess(emvs_context* ctx, const float* conf_in, const uint8_t* idx_in, uint32_t rows, uint32_t cols,
                               const float* depths, uint32_t n_depths, const emvs_depthmap_options* opt, float* depth_map,
                               float* confidence_map, uint8_t* mask, uint8_t* idx_filtered, uint8_t* conf8)
{
  REQUIRE(ctx && conf_in && idx_in && depths && depth_map && confidence_map && mask, EMVS_ERR_INVALID,
          "depth_map_postprocess: NULL argument");
  REQUIRE(rows && cols && n_depths >= 1 && n_depths <= 256, EMVS_ERR_INVALID, "depth_map_postprocess: bad size");
  int rc = check_post_options(opt);
  if (rc) return rc;
  DeviceGuard guard(ctx->device);
  const size_t n_pix = (size_t)rows * cols;
  rc = grow(&ctx->d_out, &ctx->out_cap, post_layout(nullptr, n_pix, n_depths).total);
  if (rc) return rc;
  const PostBuffers b = post_layout((char*)ctx->d_out, n_pix, n_depths);
  CUDA_TRY(cudaMemcpyAsync(b.tab, depths, (size_t)n_depths * 4, cudaMemcpyHostToDevice, ctx->stream));
  CUDA_TRY(cudaMemcpyAsync(b.conf, conf_in, n_pix * 4, cudaMemcpyHostToDevice, ctx->stream));
  CUDA_TRY(cudaMemcpyAsync(b.idx, idx_in, n_pix, cudaMemcpyHostToDevice, ctx->stream));
  rc = run_post(ctx, b, rows, cols, opt);
  if (rc) return rc;
  return post_download(ctx, b, n_pix, depth_map, confidence_map, mask, idx_filtered, conf8);
}

int emvs_grid_device_ptr(const emvs_grid* g, void** out)
{
  REQUIRE(g && out, EMVS_ERR_INVALID, "grid_device_ptr: NULL argument");
  *out = g->d;
  return EMVS_OK;
}

// ---- MapperEMVS ------------------------------------------------------------------------------
int emvs_mapper_create(emvs_context* ctx, const emvs_camera* cam, const emvs_shape* shape, emvs_mapper** out)
{
  REQUIRE(ctx && cam && out, EMVS_ERR_INVALID, "mapper_create: NULL argument");
  *out = nullptr;
  int rc = check_shape(shape);
  if (rc) return rc;
  REQUIRE(cam->width && cam->height, EMVS_ERR_INVALID, "mapper_create: empty sensor");
  REQUIRE(cam->width <= 65536 && cam->height <= 65536, EMVS_ERR_INVALID, "mapper_create: sensor exceeds uint16 event coordinates");
  // geometry_utils.hpp:36-41 CHECKs on the virtual camera
  REQUIRE(cam->fx > 0.f && cam->fy > 0.f && cam->cx > 0.f && cam->cy > 0.f, EMVS_ERR_INVALID,
          "mapper_create: fx, fy, cx, cy must be > 0");
  DeviceGuard guard(ctx->device);
  emvs_mapper* m = new (std::nothrow) emvs_mapper;
  REQUIRE(m, EMVS_ERR_INVALID, "out of host memory");
  m->ctx = ctx;
  context_retain(ctx);
  m->cam = *cam;
  m->shape = *shape;
  if (!m->shape.dimX) m->shape.dimX = cam->width;    // MAP:216
  if (!m->shape.dimY) m->shape.dimY = cam->height;   // MAP:217
  host_virtual_camera(*cam, m->shape, m->virt);
  m->depths.resize(shape->dimZ);
  host_depth_vector(m->shape, m->depths.data());
  rc = emvs_grid_create(ctx, m->shape.dimX, m->shape.dimY, m->shape.dimZ, &m->grid);
  if (rc) { context_release(ctx); delete m; return rc; }
  cudaError_t e = cudaMalloc((void**)&m->d_depths, sizeof(float) * shape->dimZ);
  if (e == cudaSuccess) e = cudaMemcpyAsync(m->d_depths, m->depths.data(), sizeof(float) * shape->dimZ, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaMalloc((void**)&m->d_lut, sizeof(float2) * (size_t)cam->width * cam->height);
  if (e == cudaSuccess) e = cudaMalloc((void**)&m->d_counts, sizeof(unsigned long long) * shape->dimZ);
  if (e == cudaSuccess) e = cudaMemsetAsync(m->d_counts, 0, sizeof(unsigned long long) * shape->dimZ, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) {
    set_error("mapper_create: %s", cudaGetErrorString(e));
    emvs_mapper_destroy(m);
    return EMVS_ERR_CUDA;
  }
  *out = m;
  return EMVS_OK;
}

int emvs_mapper_destroy(emvs_mapper* m)
{
  if (!m) return EMVS_OK;
  DeviceGuard guard(m->ctx->device);
  cudaStreamSynchronize(m->ctx->stream);
  // packets prefetched for this mapper must not be matched by a later mapper allocated at the same address
  // (the prefetched events stay usable: they do not depend on the mapper)
  if (m->ctx->prefetch.mapper == m) {
    m->ctx->prefetch.has_packets = false;
    m->ctx->prefetch.mapper = nullptr;
  }
  emvs_grid_destroy(m->grid);
  cudaFree(m->d_depths);
  cudaFree(m->d_lut);
  cudaFree(m->d_counts);
  context_release(m->ctx);
  delete m;
  return EMVS_OK;
}

int emvs_mapper_set_lut(emvs_mapper* m, const float* lut_xy, size_t n_pixels)
{
  REQUIRE(m && lut_xy, EMVS_ERR_INVALID, "set_lut: NULL argument");
  REQUIRE(n_pixels == (size_t)m->cam.width * m->cam.height, EMVS_ERR_INVALID, "set_lut: size must be width*height");
  DeviceGuard guard(m->ctx->device);
  CUDA_TRY(cudaMemcpyAsync(m->d_lut, lut_xy, n_pixels * sizeof(float2), cudaMemcpyHostToDevice, m->ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(m->ctx->stream));
  m->lut_set = true;
  return EMVS_OK;
}

int emvs_mapper_shape(const emvs_mapper* m, emvs_shape* shape_out, float virt_out[4])
{
  REQUIRE(m, EMVS_ERR_INVALID, "mapper is NULL");
  if (shape_out) *shape_out = m->shape;
  if (virt_out) memcpy(virt_out, m->virt, sizeof m->virt);
  return EMVS_OK;
}

int emvs_mapper_depths(const emvs_mapper* m, float* out)
{
  REQUIRE(m && out, EMVS_ERR_INVALID, "mapper_depths: NULL argument");
  memcpy(out, m->depths.data(), m->depths.size() * sizeof(float));
  return EMVS_OK;
}

int emvs_mapper_depths_device(const emvs_mapper* m, const float** out)
{
  REQUIRE(m && out, EMVS_ERR_INVALID, "mapper_depths_device: NULL argument");
  *out = m->d_depths;
  return EMVS_OK;
}

int emvs_mapper_grid(emvs_mapper* m, emvs_grid** out)
{
  REQUIRE(m && out, EMVS_ERR_INVALID, "mapper_grid: NULL argument");
  *out = m->grid;
  return EMVS_OK;
}

static int check_packets(const emvs_packet* pk, size_t n_packets, size_t n_events)
{
  for (size_t j = 0; j < n_packets; ++j)
    if (pk[j].first_event + EMVS_PACKET_SIZE > n_events) {
      set_error("packet %zu reaches past the event list (first_event=%llu, n_events=%zu)", j,
                (unsigned long long)pk[j].first_event, n_events);
      return EMVS_ERR_INVALID;
    }
  return EMVS_OK;
}

// A caller's event list in host memory: the 16-byte dvs_msgs::Event structs (the reference's
// std::vector<dvs_msgs::Event>), or an emvs_events_soa (separate x / y / t arrays).
struct HostEvents {
  const emvs_event* aos = nullptr;
  const uint16_t* x = nullptr;
  const uint16_t* y = nullptr;
  const int64_t* t_ns = nullptr;
  size_t n = 0;
  bool soa() const { return aos == nullptr; }
  const void* key() const { return aos ? (const void*)aos : (const void*)x; }
  EventTimes times() const
  {
    EventTimes t;
    t.aos = aos;
    t.t_ns = t_ns;
    return t;
  }
};

static HostEvents host_aos(const emvs_event* ev, size_t n)
{
  HostEvents h;
  h.aos = ev;
  h.n = n;
  return h;
}

// SoA lists are staged as [x: n uint16 | pad to 256 B | y: n uint16]
static size_t soa_pitch(size_t n) { return (n * sizeof(uint16_t) + 255) & ~(size_t)255; }
static size_t staging_bytes(const HostEvents& ev) { return ev.soa() ? 2 * soa_pitch(ev.n) : ev.n * sizeof(emvs_event); }

static EventSrc device_src(const emvs_context* ctx, const HostEvents& ev, int buf)
{
  EventSrc d;
  char* base = (char*)ctx->d_events[buf];
  if (ev.soa()) {
    d.x = (const uint16_t*)base;
    d.y = (const uint16_t*)(base + soa_pitch(ev.n));
  } else {
    d.aos = (const emvs_event*)base;
  }
  return d;
}

// Host-buffer build.  Staging protocol (one context = one in-order pipeline):
//   copy_stream:  wait(ev_consumed of the staging buffer's previous build) -> H2D events, packets -> record ev_copied
//   stream:       wait(ev_copied) -> k_warp_events -> record ev_consumed -> slab loop
//   host:         waits for ev_copied only, so the caller may reuse its buffers and issue the next
//                 camera's build, whose upload then overlaps this build's vote kernels.
// d_events is read by k_warp_events only; d_packets is read by every vote launch, hence two
// alternating packet buffers: build N+2 can only upload after k_warp_events of build N+1 ran,
// which is stream-ordered behind the last vote of build N.
// Picks the staging buffer of the next upload (never the one that holds a pending prefetch), sizes it for the
// list and orders the copy stream behind the last event stage that read it.  It becomes the current buffer.
static int stage_events(emvs_context* ctx, size_t bytes)
{
  int u = (int)(ctx->upload_next & 1u);
  if (ctx->prefetch.valid && ctx->prefetch.buf == u) u ^= 1;
  ctx->upload_next = (unsigned)u + 1u;
  if (bytes > ctx->events_cap[u] && ctx->consumed_recorded[u]) {
    // growing frees the old buffer: the event stage that may still be reading it must have finished (cudaFree would
    // wait for the whole device anyway; this keeps the wait to the one kernel concerned)
    CUDA_TRY(cudaEventSynchronize(ctx->ev_consumed[u]));
  }
  const int rc = grow(&ctx->d_events[u], &ctx->events_cap[u], bytes);
  if (rc) return rc;
  if (ctx->consumed_recorded[u]) CUDA_TRY(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_consumed[u], 0));
  ctx->cur_events = u;
  return EMVS_OK;
}

static bool prefetch_matches(const emvs_context* ctx, const HostEvents& ev)
{
  return ctx->prefetch.valid && ctx->prefetch.host == ev.key() && ctx->prefetch.n == ev.n && ctx->prefetch.soa == ev.soa();
}

// A pending prefetch is consumed by the NEXT host-buffer build / evaluate call on the context if that call names the
// same list, and dropped otherwise: a stale announcement can never be matched to some later, unrelated list that
// happens to live at the same address (its staging buffer is simply reused).
static void drop_unmatched_prefetch(emvs_context* ctx, const HostEvents& ev)
{
  if (ctx->prefetch.valid && !prefetch_matches(ctx, ev)) {
    ctx->prefetch.valid = false;
    ctx->prefetch.has_packets = false;
  }
}

// A list announced with emvs_context_prefetch_events is already in (or on its way to) a staging buffer.
static bool take_prefetch(emvs_context* ctx, const HostEvents& ev)
{
  if (!prefetch_matches(ctx, ev)) return false;
  ctx->prefetch.valid = false;
  ctx->prefetch.has_packets = false;
  ctx->cur_events = ctx->prefetch.buf;
  return true;
}

// Packet buffer of the next host-buffer build: alternates, never the one a pending prefetch has filled, and the
// copy stream is ordered behind the votes of the last build that read it.
static int stage_packets(emvs_context* ctx, size_t n_packets_cap, unsigned* par_out)
{
  unsigned par = ctx->cur_packets ^ 1u;   // not the one the latest build is voting from
  if (ctx->prefetch.valid && ctx->prefetch.has_packets && ctx->prefetch.par == par) par ^= 1u;
  if (n_packets_cap * sizeof(emvs_packet) > ctx->packets_cap[par] && ctx->pk_free_recorded[par])
    CUDA_TRY(cudaEventSynchronize(ctx->ev_pk_free[par]));   // see stage_events
  const int rc = grow(&ctx->d_packets[par], &ctx->packets_cap[par], n_packets_cap * sizeof(emvs_packet));
  if (rc) return rc;
  if (ctx->pk_free_recorded[par]) CUDA_TRY(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_pk_free[par], 0));
  *par_out = par;
  return EMVS_OK;
}

// enqueue the host -> device copy of events [lo, hi) into the current staging buffer
static int copy_event_range(emvs_context* ctx, const HostEvents& ev, size_t lo, size_t hi)
{
  if (hi <= lo) return EMVS_OK;
  char* base = (char*)ctx->d_events[ctx->cur_events];
  if (ev.soa()) {
    CUDA_TRY(cudaMemcpyAsync(base + lo * 2, ev.x + lo, (hi - lo) * 2, cudaMemcpyHostToDevice, ctx->copy_stream));
    CUDA_TRY(cudaMemcpyAsync(base + soa_pitch(ev.n) + lo * 2, ev.y + lo, (hi - lo) * 2, cudaMemcpyHostToDevice, ctx->copy_stream));
  } else {
    CUDA_TRY(cudaMemcpyAsync(base + lo * sizeof(emvs_event), ev.aos + lo, (hi - lo) * sizeof(emvs_event), cudaMemcpyHostToDevice,
                             ctx->copy_stream));
  }
  return EMVS_OK;
}

static int upload_events(emvs_context* ctx, const HostEvents& ev, size_t lo, size_t hi)
{
  const int rc = stage_events(ctx, staging_bytes(ev));
  if (rc) return rc;
  return copy_event_range(ctx, ev, lo, hi);
}

// tail_lo < tail_hi: events [tail_lo, tail_hi) are enqueued for upload right after this build's packets and before
// its kernels are launched (split upload of evaluate_dsi: the copy engine never idles while the host launches
// the head's kernels).
static int build_from_host(emvs_mapper* m, const HostEvents& ev, const emvs_packet* packets, size_t n_packets, int flags,
                           bool events_uploaded, size_t tail_lo = 0, size_t tail_hi = 0)
{
  emvs_context* ctx = m->ctx;
  unsigned par = 0;
  {
    // sized for the whole list so that head / tail / whole-list builds alternating over the two buffers never regrow them
    const int rc = stage_packets(ctx, std::max(n_packets, ev.n / EMVS_PACKET_SIZE + 1), &par);
    if (rc) return rc;
  }
  ctx->cur_packets = par;
  if (n_packets) {
    if (!events_uploaded && !take_prefetch(ctx, ev)) {
      size_t lo = ev.n, last = 0;   // only the span of events that packets reference has to travel
      for (size_t j = 0; j < n_packets; ++j) {
        lo = std::min<size_t>(lo, packets[j].first_event);
        last = std::max<size_t>(last, packets[j].first_event + EMVS_PACKET_SIZE);
      }
      const int rc = upload_events(ctx, ev, lo, last);
      if (rc) return rc;
    }
    CUDA_TRY(cudaMemcpyAsync(ctx->d_packets[par], packets, n_packets * sizeof(emvs_packet), cudaMemcpyHostToDevice,
                             ctx->copy_stream));
    CUDA_TRY(cudaEventRecord(ctx->ev_copied, ctx->copy_stream));
    CUDA_TRY(cudaStreamWaitEvent(ctx->stream, ctx->ev_copied, 0));
  }
  {
    const int rc = copy_event_range(ctx, ev, tail_lo, tail_hi);
    if (rc) return rc;
  }
  ctx->mark_consumed = true;
  const int rc = build_on_device(m, device_src(ctx, ev, ctx->cur_events), ev.n, (const emvs_packet*)ctx->d_packets[par], n_packets,
                                 flags);
  ctx->mark_consumed = false;
  if (rc) return rc;
  if (n_packets) CUDA_TRY(cudaEventSynchronize(ctx->ev_copied));
  return EMVS_OK;
}

static int ensure_pinned_packets(emvs_packet** p, size_t* cap, size_t need)
{
  if (need <= *cap) return EMVS_OK;
  if (*p) CUDA_TRY(cudaFreeHost(*p));
  *p = nullptr;
  *cap = 0;
  CUDA_TRY(cudaHostAlloc((void**)p, need * sizeof(emvs_packet), cudaHostAllocDefault));
  *cap = need;
  return EMVS_OK;
}

static int prefetch_events_impl(emvs_context* ctx, const HostEvents& ev)
{
  ctx->prefetch.valid = false;          // an earlier, unconsumed prefetch is dropped: its buffer is free again
  ctx->prefetch.has_packets = false;
  const int keep = ctx->cur_events;     // stage_events moves cur_events; the current build keeps its own buffer
  int rc = stage_events(ctx, staging_bytes(ev));
  const int buf = ctx->cur_events;
  if (!rc) rc = copy_event_range(ctx, ev, 0, ev.n);
  ctx->cur_events = keep;
  if (rc) return rc;
  ctx->prefetch.host = ev.key();
  ctx->prefetch.soa = ev.soa();
  ctx->prefetch.n = ev.n;
  ctx->prefetch.buf = buf;
  ctx->prefetch.valid = true;
  ctx->prefetch.generation = ++ctx->prefetch_generation;
  return EMVS_OK;
}

static int prefetch_dsi_impl(emvs_mapper* m, const HostEvents& ev, const emvs_stamped_pose* traj, size_t n_poses,
                             const emvs_pose* T_rv_w)
{
  emvs_context* ctx = m->ctx;
  // the pinned packet buffer of an earlier prefetch may still be in flight
  if (ctx->prefetched_recorded) CUDA_TRY(cudaEventSynchronize(ctx->ev_prefetched));
  int rc = prefetch_events_impl(ctx, ev);
  if (rc) return rc;
  const size_t max_pk = ev.n / EMVS_PACKET_SIZE + 1;
  rc = ensure_pinned_packets(&ctx->h_packets_pf, &ctx->h_packets_pf_cap, max_pk);
  if (rc) return rc;
  const size_t n_pk = host_packetize(ev.times(), ev.n, traj, n_poses, *T_rv_w, m->cam, m->virt, m->depths[0], ctx->h_packets_pf,
                                     max_pk);
  unsigned par = 0;
  rc = stage_packets(ctx, max_pk, &par);
  if (rc) return rc;
  if (n_pk)
    CUDA_TRY(cudaMemcpyAsync(ctx->d_packets[par], ctx->h_packets_pf, n_pk * sizeof(emvs_packet), cudaMemcpyHostToDevice,
                             ctx->copy_stream));
  CUDA_TRY(cudaEventRecord(ctx->ev_prefetched, ctx->copy_stream));
  ctx->prefetched_recorded = true;
  ctx->prefetch.has_packets = true;
  ctx->prefetch.n_pk = n_pk;
  ctx->prefetch.par = par;
  ctx->prefetch.mapper = m;
  ctx->prefetch.traj = traj;
  ctx->prefetch.n_poses = n_poses;
  ctx->prefetch.T_rv_w = *T_rv_w;
  return EMVS_OK;
}

// Whole evaluateDSI for an event list in host memory (AoS or SoA).
static int evaluate_dsi_impl(emvs_mapper* m, const HostEvents& ev, const emvs_stamped_pose* traj, size_t n_poses,
                             const emvs_pose* T_rv_w, int flags)
{
  emvs_context* ctx = m->ctx;
  const size_t n_events = ev.n;
  const size_t max_pk = n_events / EMVS_PACKET_SIZE + 1;
  int rc = ensure_pinned_packets(&ctx->h_packets, &ctx->h_packets_cap, max_pk);
  if (rc) return rc;
  drop_unmatched_prefetch(ctx, ev);
  const EventTimes times = ev.times();
  if (prefetch_matches(ctx, ev) && ctx->prefetch.has_packets && ctx->prefetch.mapper == m && ctx->prefetch.traj == traj &&
      ctx->prefetch.n_poses == n_poses && std::memcmp(&ctx->prefetch.T_rv_w, T_rv_w, sizeof(emvs_pose)) == 0) {
    // emvs_mapper_prefetch_dsi ran for exactly this call: events and packets are in HBM (or landing), only the
    // kernels are left
    ctx->prefetch.valid = false;
    ctx->prefetch.has_packets = false;
    ctx->cur_events = ctx->prefetch.buf;
    ctx->cur_packets = ctx->prefetch.par;
    CUDA_TRY(cudaStreamWaitEvent(ctx->stream, ctx->ev_prefetched, 0));
    ctx->mark_consumed = true;
    rc = build_on_device(m, device_src(ctx, ev, ctx->cur_events), n_events, (const emvs_packet*)ctx->d_packets[ctx->cur_packets],
                         ctx->prefetch.n_pk, flags);
    ctx->mark_consumed = false;
    if (rc) return rc;
    CUDA_TRY(cudaEventSynchronize(ctx->ev_prefetched));   // the caller may reuse its list on return
    return EMVS_OK;
  }
  if (take_prefetch(ctx, ev)) {
    // the list was announced earlier (emvs_context_prefetch_events): nothing to upload, the packet stage is all
    // that stands between the call and the first vote
    const size_t n_pk = host_packetize(times, n_events, traj, n_poses, *T_rv_w, m->cam, m->virt, m->depths[0], ctx->h_packets, max_pk);
    rc = build_from_host(m, ev, ctx->h_packets, n_pk, flags, true);
    if (rc) return rc;
    if (n_pk == 0) CUDA_TRY(cudaStreamSynchronize(ctx->copy_stream));
    return EMVS_OK;
  }
  // Split upload.  When the pipeline is idle nothing hides this call's event upload (80 MB per 5 M events, ~1.6 ms
  // over PCIe 5): upload and vote the HEAD of the list first (a build of its packets), and let the TAIL cross
  // PCIe under the head's vote kernels; the tail is then voted with EMVS_BUILD_ACCUMULATE into the same DSI
  // (voting is a sum over events, so head + tail == whole list up to float summation order; the per-plane
  // counters add exactly).  When the stream is busy (the previous camera is still voting) the whole upload is
  // already hidden and the list is built in one piece.  The slab-wise exchanges need final slabs: with the peer
  // reduce the head is a plain build and only the tail build (the last one into the DSI) announces its slabs; the
  // NCCL allreduce form is not split.
  const bool exchange = (flags & EMVS_BUILD_ALLREDUCE) != 0;
  const int head_flags = flags & ~EMVS_BUILD_PEER_REDUCE;
  size_t n_head = 0;
  // vote-only early pieces (see split_defer_merge) where the build takes the multi-slab launch; not under a peer exchange
  // (that combination has not been run on hardware: such builds keep the head + tail form)
  const bool defer_mode = ctx->split_defer_merge && !(flags & EMVS_BUILD_PEER_REDUCE) &&
                          multislab_eligible(ctx, m->grid->dimX, m->grid->dimY, m->grid->dimZ);
  const uint32_t split_percent = defer_mode ? ctx->defer_percent : ctx->split_percent;
  const int split_pieces = defer_mode ? ctx->defer_pieces : ctx->split_pieces;
  if (ctx->split_percent && !exchange && n_events >= ctx->split_min_events && n_events >= 4 * (size_t)EMVS_PACKET_SIZE) {
    const cudaError_t q = cudaStreamQuery(ctx->stream);
    if (q == cudaSuccess) n_head = (n_events / 100 * split_percent) / EMVS_PACKET_SIZE * EMVS_PACKET_SIZE;
    else if (q == cudaErrorNotReady) (void)cudaGetLastError();   // "busy" is an answer, not an error: do not leave it behind
    else CUDA_TRY(q);
  }
  if (n_head >= EMVS_PACKET_SIZE) {
    // Pieces [0, c0), [c0, c1), ..., [c_last, n).  Default: head + tail.  With split_pieces = 3 or 4 the pieces grow by a
    // factor 2.2 (p, 2.2 p, 4.84 p % of the list, the last one takes the rest): PCIe delivers events about 2.4 times faster
    // than they are voted, so every piece's upload hides under the previous piece's votes.  As complete builds extra pieces
    // were measured SLOWER (10.1 vs 9.1 ms per stock step with 3 pieces, profiles/r2_e2e.md: each is one more merge + re-zero
    // pass over all slabs); with split_defer_merge the pieces before the last only vote.
    size_t cuts[5];
    int n_cuts = 0;
    cuts[n_cuts++] = n_head;
    {
      double piece = (double)split_percent, cum = piece;
      while (n_cuts < split_pieces - 1) {
        piece *= 2.2;
        cum += piece;
        if (cum > 75.0) break;
        const size_t c = (size_t)((double)n_events * cum / 100.0) / EMVS_PACKET_SIZE * EMVS_PACKET_SIZE;
        if (c <= cuts[n_cuts - 1] + EMVS_PACKET_SIZE || c + EMVS_PACKET_SIZE >= n_events) break;
        cuts[n_cuts++] = c;
      }
    }
    cuts[n_cuts] = n_events;   // sentinel: the end of the list
    rc = upload_events(ctx, ev, 0, cuts[0]);
    if (rc) return rc;
    size_t cur = 0, n_pk_done = 0;
    bool built = false;
    for (int k = 0; k <= n_cuts; ++k) {
      const bool last = k == n_cuts;
      const size_t limit = cuts[k];
      const size_t n_pk = host_packetize_range(times, n_events, traj, n_poses, *T_rv_w, m->cam, m->virt, m->depths[0], &cur, limit,
                                               ctx->h_packets + n_pk_done, max_pk - n_pk_done);
      const size_t next_lo = last ? 0 : cuts[k], next_hi = last ? 0 : cuts[k + 1];
      // only the LAST build into the DSI announces its slabs to the peers; an empty last piece still has to
      const bool must_build = n_pk > 0 || (last && (!built || ctx->deferred_votes || (flags & EMVS_BUILD_PEER_REDUCE)));
      if (must_build) {
        int f = last ? flags : head_flags;
        // earlier pieces left their votes in the scratch (deferred merge): vote on top; else they are in the DSI: accumulate
        if (built) f |= ctx->deferred_votes ? kBuildContinue : EMVS_BUILD_ACCUMULATE;
        if (!last && defer_mode && (!built || ctx->deferred_votes)) f |= kBuildDeferMerge;
        rc = build_from_host(m, ev, ctx->h_packets + n_pk_done, n_pk, f, true, next_lo, next_hi);
        if (rc) return rc;
        built = true;
      } else if (!last) {
        rc = copy_event_range(ctx, ev, next_lo, next_hi);   // nothing to vote in this piece: just keep the upload going
        if (rc) return rc;
      }
      n_pk_done += n_pk;
    }
    CUDA_TRY(cudaStreamSynchronize(ctx->copy_stream));  // the caller may reuse its list on return
    return EMVS_OK;
  }
  // start the event upload first: the host packet stage (one pose + one 3x3 inverse per 1024
  // events) then runs while the copy engine is busy
  rc = upload_events(ctx, ev, 0, n_events);
  if (rc) return rc;
  const size_t n_pk = host_packetize(times, n_events, traj, n_poses, *T_rv_w, m->cam, m->virt, m->depths[0], ctx->h_packets, max_pk);
  rc = build_from_host(m, ev, ctx->h_packets, n_pk, flags, true);
  if (rc) return rc;
  if (n_pk == 0) CUDA_TRY(cudaStreamSynchronize(ctx->copy_stream));  // nothing waited for the event upload
  return EMVS_OK;
}

static int check_soa(const emvs_events_soa* ev, const char* who)
{
  if (!ev || !ev->x || !ev->y || !ev->t_ns) {
    set_error("%s: NULL event arrays", who);
    return EMVS_ERR_INVALID;
  }
  return EMVS_OK;
}

static HostEvents host_soa(const emvs_events_soa* ev)
{
  HostEvents h;
  h.x = ev->x;
  h.y = ev->y;
  h.t_ns = ev->t_ns;
  h.n = ev->n;
  return h;
}

int emvs_context_prefetch_events(emvs_context* ctx, const emvs_event* events, size_t n_events)
{
  REQUIRE(ctx && events && n_events, EMVS_ERR_INVALID, "prefetch_events: NULL argument or empty list");
  DeviceGuard guard(ctx->device);
  return prefetch_events_impl(ctx, host_aos(events, n_events));
}

int emvs_context_prefetch_pending(emvs_context* ctx, uint64_t* generation)
{
  REQUIRE(ctx && generation, EMVS_ERR_INVALID, "prefetch_pending: NULL argument");
  *generation = ctx->prefetch.valid ? ctx->prefetch.generation : 0;
  return EMVS_OK;
}

int emvs_context_prefetch_cancel(emvs_context* ctx)
{
  REQUIRE(ctx, EMVS_ERR_INVALID, "context is NULL");
  DeviceGuard guard(ctx->device);
  if (ctx->prefetch.valid) {
    // the copies may still be reading the caller's arrays: after this call they are the caller's again
    CUDA_TRY(cudaStreamSynchronize(ctx->copy_stream));
    ctx->prefetch.valid = false;
    ctx->prefetch.has_packets = false;
  }
  return EMVS_OK;
}

int emvs_mapper_prefetch_dsi(emvs_mapper* m, const emvs_event* events, size_t n_events, const emvs_stamped_pose* traj,
                             size_t n_poses, const emvs_pose* T_rv_w)
{
  REQUIRE(m && events && traj && T_rv_w, EMVS_ERR_INVALID, "prefetch_dsi: NULL argument");
  REQUIRE(n_poses >= 2, EMVS_ERR_INVALID, "At least two poses need to be provided");
  if (n_events < EMVS_PACKET_SIZE) return EMVS_OK;   // the later evaluateDSI returns false without touching the device
  DeviceGuard guard(m->ctx->device);
  return prefetch_dsi_impl(m, host_aos(events, n_events), traj, n_poses, T_rv_w);
}

int emvs_mapper_prefetch_dsi_soa(emvs_mapper* m, const emvs_events_soa* events, const emvs_stamped_pose* traj, size_t n_poses,
                                 const emvs_pose* T_rv_w)
{
  REQUIRE(m && traj && T_rv_w, EMVS_ERR_INVALID, "prefetch_dsi_soa: NULL argument");
  int rc = check_soa(events, "prefetch_dsi_soa");
  if (rc) return rc;
  REQUIRE(n_poses >= 2, EMVS_ERR_INVALID, "At least two poses need to be provided");
  if (events->n < EMVS_PACKET_SIZE) return EMVS_OK;
  DeviceGuard guard(m->ctx->device);
  return prefetch_dsi_impl(m, host_soa(events), traj, n_poses, T_rv_w);
}

int emvs_mapper_build(emvs_mapper* m, const emvs_event* events, size_t n_events, const emvs_packet* packets,
                      size_t n_packets, int flags)
{
  REQUIRE(m, EMVS_ERR_INVALID, "mapper is NULL");
  REQUIRE(m->lut_set, EMVS_ERR_STATE, "mapper_build: rectification LUT not set (emvs_mapper_set_lut)");
  REQUIRE((flags & ~kPublicBuildFlags) == 0, EMVS_ERR_INVALID, "mapper_build: unknown build flag");
  REQUIRE(n_packets == 0 || (events && packets), EMVS_ERR_INVALID, "mapper_build: NULL events/packets");
  int rc = check_packets(packets, n_packets, n_events);
  if (rc) return rc;
  DeviceGuard guard(m->ctx->device);
  const HostEvents ev = host_aos(events, n_events);
  drop_unmatched_prefetch(m->ctx, ev);
  return build_from_host(m, ev, packets, n_packets, flags, false);
}

int emvs_mapper_build_device(emvs_mapper* m, const void* d_events, size_t n_events, const void* d_packets,
                             size_t n_packets, int flags)
{
  REQUIRE(m, EMVS_ERR_INVALID, "mapper is NULL");
  REQUIRE(m->lut_set, EMVS_ERR_STATE, "mapper_build_device: rectification LUT not set (emvs_mapper_set_lut)");
  REQUIRE((flags & ~kPublicBuildFlags) == 0, EMVS_ERR_INVALID, "mapper_build_device: unknown build flag");
  REQUIRE(n_packets == 0 || (d_events && d_packets), EMVS_ERR_INVALID, "mapper_build_device: NULL events/packets");
  DeviceGuard guard(m->ctx->device);
  EventSrc src;
  src.aos = (const emvs_event*)d_events;
  return build_on_device(m, src, n_events, (const emvs_packet*)d_packets, n_packets, flags);
}

int emvs_mapper_evaluate_dsi(emvs_mapper* m, const emvs_event* events, size_t n_events,
                             const emvs_stamped_pose* traj, size_t n_poses, const emvs_pose* T_rv_w)
{
  return emvs_mapper_evaluate_dsi_flags(m, events, n_events, traj, n_poses, T_rv_w, EMVS_BUILD_RESET);
}

int emvs_mapper_evaluate_dsi_flags(emvs_mapper* m, const emvs_event* events, size_t n_events,
                                   const emvs_stamped_pose* traj, size_t n_poses, const emvs_pose* T_rv_w, int flags)
{
  REQUIRE(m && events && traj && T_rv_w, EMVS_ERR_INVALID, "evaluate_dsi: NULL argument");
  REQUIRE(n_poses >= 2, EMVS_ERR_INVALID, "At least two poses need to be provided");
  if (n_events < EMVS_PACKET_SIZE) {
    set_error("Number of events (%zu) < packet size (%d)", n_events, EMVS_PACKET_SIZE);
    return EMVS_ERR_TOO_FEW;
  }
  REQUIRE(m->lut_set, EMVS_ERR_STATE, "evaluate_dsi: rectification LUT not set (emvs_mapper_set_lut)");
  REQUIRE((flags & ~kPublicBuildFlags) == 0, EMVS_ERR_INVALID, "evaluate_dsi: unknown build flag");
  DeviceGuard guard(m->ctx->device);
  return evaluate_dsi_impl(m, host_aos(events, n_events), traj, n_poses, T_rv_w, flags);
}

int emvs_mapper_evaluate_dsi_soa(emvs_mapper* m, const emvs_events_soa* events, const emvs_stamped_pose* traj, size_t n_poses,
                                 const emvs_pose* T_rv_w, int flags)
{
  REQUIRE(m && traj && T_rv_w, EMVS_ERR_INVALID, "evaluate_dsi_soa: NULL argument");
  int rc = check_soa(events, "evaluate_dsi_soa");
  if (rc) return rc;
  REQUIRE(n_poses >= 2, EMVS_ERR_INVALID, "At least two poses need to be provided");
  if (events->n < EMVS_PACKET_SIZE) {
    set_error("Number of events (%zu) < packet size (%d)", events->n, EMVS_PACKET_SIZE);
    return EMVS_ERR_TOO_FEW;
  }
  REQUIRE(m->lut_set, EMVS_ERR_STATE, "evaluate_dsi_soa: rectification LUT not set (emvs_mapper_set_lut)");
  REQUIRE((flags & ~kPublicBuildFlags) == 0, EMVS_ERR_INVALID, "evaluate_dsi_soa: unknown build flag");
  DeviceGuard guard(m->ctx->device);
  return evaluate_dsi_impl(m, host_soa(events), traj, n_poses, T_rv_w, flags);
}

int emvs_packetize_soa(const emvs_events_soa* events, const emvs_stamped_pose* traj, size_t n_poses, const emvs_pose* T_rv_w,
                       const emvs_camera* cam, const float virt[4], float z0, emvs_packet* out, size_t max_packets,
                       size_t* n_packets)
{
  REQUIRE(traj && T_rv_w && cam && virt && n_packets, EMVS_ERR_INVALID, "packetize_soa: NULL argument");
  int rc = check_soa(events, "packetize_soa");
  if (rc) return rc;
  REQUIRE(n_poses >= 2, EMVS_ERR_INVALID, "At least two poses need to be provided");
  *n_packets = 0;
  if (events->n < EMVS_PACKET_SIZE) {
    set_error("Number of events (%zu) < packet size (%d)", events->n, EMVS_PACKET_SIZE);
    return EMVS_ERR_TOO_FEW;
  }
  REQUIRE(out || max_packets == 0, EMVS_ERR_INVALID, "packetize_soa: out is NULL");
  bool truncated = false;
  *n_packets = host_packetize(host_soa(events).times(), events->n, traj, n_poses, *T_rv_w, *cam, virt, z0, out, max_packets, &truncated);
  REQUIRE(!truncated, EMVS_ERR_INVALID, "packetize_soa: max_packets is too small for this list (n_events / 1024 + 1 always suffices)");
  return EMVS_OK;
}

int emvs_mapper_counts(const emvs_mapper* m, uint64_t* per_plane)
{
  REQUIRE(m && per_plane, EMVS_ERR_INVALID, "mapper_counts: NULL argument");
  DeviceGuard guard(m->ctx->device);
  CUDA_TRY(cudaMemcpyAsync(per_plane, m->d_counts, sizeof(uint64_t) * m->shape.dimZ, cudaMemcpyDeviceToHost,
                           m->ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(m->ctx->stream));
  return EMVS_OK;
}

// ---- multi-GPU --------------------------------------------------------------------------------
#define NCCL_TRY(api, expr)                                                          \
  do {                                                                               \
    int r_ = (expr);                                                                 \
    if (r_ != 0) {                                                                   \
      set_error("NCCL error at %s:%d: %s", __FILE__, __LINE__, (api)->GetErrorString(r_)); \
      return EMVS_ERR_NCCL;                                                          \
    }                                                                                \
  } while (0)

int emvs_comm_unique_id(uint8_t out_id[128])
{
  REQUIRE(out_id, EMVS_ERR_INVALID, "comm_unique_id: NULL argument");
  const NcclApi* api = nccl_api();
  if (!api) return EMVS_ERR_NCCL;
  NcclId id;
  NCCL_TRY(api, api->GetUniqueId(&id));
  memcpy(out_id, id.b, 128);
  return EMVS_OK;
}

int emvs_comm_init(emvs_context* ctx, const uint8_t id_bytes[128], int n_ranks, int rank)
{
  REQUIRE(ctx && id_bytes, EMVS_ERR_INVALID, "comm_init: NULL argument");
  REQUIRE(n_ranks >= 1 && rank >= 0 && rank < n_ranks, EMVS_ERR_INVALID, "comm_init: bad rank / n_ranks");
  REQUIRE(!ctx->comm, EMVS_ERR_STATE, "comm_init: communicator already initialised");
  const NcclApi* api = nccl_api();
  if (!api) return EMVS_ERR_NCCL;
  DeviceGuard guard(ctx->device);
  NcclId id;
  memcpy(id.b, id_bytes, 128);
  NCCL_TRY(api, api->CommInitRank(&ctx->comm, n_ranks, id, rank));
  if (!ctx->comm_stream) CUDA_TRY(cudaStreamCreateWithFlags(&ctx->comm_stream, cudaStreamNonBlocking));
  if (!ctx->ev_comm_done) CUDA_TRY(cudaEventCreateWithFlags(&ctx->ev_comm_done, cudaEventDisableTiming));
  ctx->n_ranks = n_ranks;
  ctx->rank = rank;
  return EMVS_OK;
}

int emvs_comm_destroy(emvs_context* ctx)
{
  REQUIRE(ctx, EMVS_ERR_INVALID, "context is NULL");
  if (!ctx->comm) return EMVS_OK;
  const NcclApi* api = nccl_api();
  if (!api) return EMVS_ERR_NCCL;
  DeviceGuard guard(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->comm_stream) cudaStreamSynchronize(ctx->comm_stream);
  api->CommDestroy(ctx->comm);
  ctx->comm = nullptr;
  ctx->n_ranks = 1;
  ctx->rank = 0;
  return EMVS_OK;
}

int emvs_grid_allreduce(emvs_grid* g)
{
  const int rc = emvs_grid_allreduce_async(g);
  if (rc) return rc;
  DeviceGuard guard(g->ctx->device);
  CUDA_TRY(cudaStreamSynchronize(g->ctx->stream));
  return EMVS_OK;
}

int emvs_grid_allreduce_async(emvs_grid* g)
{
  REQUIRE(g, EMVS_ERR_INVALID, "grid is NULL");
  emvs_context* ctx = g->ctx;
  REQUIRE(ctx->comm, EMVS_ERR_STATE, "grid_allreduce: communicator not initialised (emvs_comm_init)");
  const NcclApi* api = nccl_api();
  if (!api) return EMVS_ERR_NCCL;
  DeviceGuard guard(ctx->device);
  // chunk by Z-slab so the ring pipeline starts on the first planes while later ones queue
  const size_t plane = (size_t)g->dimX * g->dimY;
  const size_t planes_per_chunk = std::max<size_t>(1, ((size_t)64 << 20) / (plane * sizeof(float)));
  NCCL_TRY(api, api->GroupStart());
  for (size_t k = 0; k < g->dimZ; k += planes_per_chunk) {
    const size_t nk = std::min(planes_per_chunk, (size_t)g->dimZ - k);
    float* p = g->d + k * plane;
    NCCL_TRY(api, api->AllReduce(p, p, nk * plane, /*ncclFloat32*/ 7, /*ncclSum*/ 0, ctx->comm, (void*)ctx->stream));
  }
  NCCL_TRY(api, api->GroupEnd());
  return EMVS_OK;
}

int emvs_mapper_counts_allreduce(emvs_mapper* m)
{
  REQUIRE(m, EMVS_ERR_INVALID, "mapper is NULL");
  emvs_context* ctx = m->ctx;
  REQUIRE(ctx->comm, EMVS_ERR_STATE, "counts_allreduce: communicator not initialised (emvs_comm_init)");
  const NcclApi* api = nccl_api();
  if (!api) return EMVS_ERR_NCCL;
  DeviceGuard guard(ctx->device);
  NCCL_TRY(api, api->AllReduce(m->d_counts, m->d_counts, m->shape.dimZ, /*ncclUint64*/ 5, /*ncclSum*/ 0, ctx->comm,
                               (void*)ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  return EMVS_OK;
}

// ---- fused multi-GPU sweep over peer memory ------------------------------------------------------
static const size_t kIpcBytes = sizeof(cudaIpcMemHandle_t);

int emvs_exchange_create(emvs_context* ctx, emvs_grid* const* grids, int n_cams, int n_ranks, int rank, emvs_exchange** out)
{
  REQUIRE(ctx && grids && out, EMVS_ERR_INVALID, "exchange_create: NULL argument");
  *out = nullptr;
  REQUIRE(n_cams >= 1 && n_cams <= kMaxPeerCams, EMVS_ERR_INVALID, "exchange_create: need 1..4 cameras");
  REQUIRE(n_ranks >= 1 && n_ranks <= kMaxPeerRanks && rank >= 0 && rank < n_ranks, EMVS_ERR_INVALID,
          "exchange_create: need 1..8 ranks and 0 <= rank < n_ranks");
  for (int c = 0; c < n_cams; ++c) {
    REQUIRE(grids[c] && grids[c]->ctx == ctx, EMVS_ERR_INVALID, "exchange_create: grid of another context");
    REQUIRE(same_dims(grids[0], grids[c]), EMVS_ERR_INVALID, "exchange_create: grids differ in shape");
  }
  DeviceGuard guard(ctx->device);
  emvs_exchange* ex = new (std::nothrow) emvs_exchange;
  REQUIRE(ex, EMVS_ERR_INVALID, "out of host memory");
  ex->ctx = ctx;
  ex->n_cams = n_cams; ex->n_ranks = n_ranks; ex->rank = rank;
  for (int c = 0; c < kMaxPeerCams; ++c) ex->args.cam_ranks[c] = c < n_cams ? ((1u << n_ranks) - 1u) : 0u;   // every rank builds every camera
  ex->dimX = grids[0]->dimX; ex->dimY = grids[0]->dimY; ex->dimZ = grids[0]->dimZ;
  for (int c = 0; c < n_cams; ++c) ex->local_dsi[c] = grids[c]->d;
  const size_t n_pix = (size_t)ex->dimX * ex->dimY;
  ex->off_depth = n_pix * 4;
  ex->off_idx = n_pix * 8;
  ex->maps_bytes = n_pix * 10;
  {  // rows owned by this rank (balanced to within one row)
    const uint32_t base = ex->dimY / n_ranks, extra = ex->dimY % n_ranks;
    ex->row_lo = rank * base + std::min<uint32_t>(rank, extra);
    ex->row_hi = ex->row_lo + base + ((uint32_t)rank < extra ? 1u : 0u);
  }
  const size_t flag_words = kFlagWordsPhase + (size_t)kMaxPeerCams * kMaxSlabs * kMaxPeerRanks;
  const size_t band_bytes = std::max<size_t>(16, (size_t)n_cams * ex->dimZ * (ex->row_hi - ex->row_lo) * ex->dimX * sizeof(float));
  cudaError_t e = cudaMalloc((void**)&ex->maps, ex->maps_bytes);
  if (e == cudaSuccess) e = cudaMalloc((void**)&ex->flags, sizeof(unsigned int) * flag_words);
  if (e == cudaSuccess) e = cudaMalloc((void**)&ex->band_buf, band_bytes);
  if (e == cudaSuccess) e = cudaMemset(ex->flags, 0, sizeof(unsigned int) * flag_words);
  if (e == cudaSuccess) e = cudaMemset(ex->maps, 0, ex->maps_bytes);
  if (e == cudaSuccess && !ctx->comm_stream) e = cudaStreamCreateWithFlags(&ctx->comm_stream, cudaStreamNonBlocking);
  if (e == cudaSuccess && !ctx->ev_comm_done) e = cudaEventCreateWithFlags(&ctx->ev_comm_done, cudaEventDisableTiming);
  if (e != cudaSuccess) {
    set_error("exchange_create: %s", cudaGetErrorString(e));
    cudaFree(ex->maps);
    cudaFree(ex->flags);
    cudaFree(ex->band_buf);
    delete ex;
    return EMVS_ERR_CUDA;
  }
  context_retain(ctx);
  *out = ex;
  return EMVS_OK;
}

int emvs_exchange_destroy(emvs_exchange* ex)
{
  if (!ex) return EMVS_OK;
  DeviceGuard guard(ex->ctx->device);
  cudaStreamSynchronize(ex->ctx->stream);
  if (ex->ctx->comm_stream) cudaStreamSynchronize(ex->ctx->comm_stream);
  if (ex->ctx->active_exchange == ex) ex->ctx->active_exchange = nullptr;
  for (void* p : ex->opened) cudaIpcCloseMemHandle(p);
  cudaFree(ex->maps);
  cudaFree(ex->flags);
  cudaFree(ex->band_buf);
  context_release(ex->ctx);
  delete ex;
  return EMVS_OK;
}

int emvs_exchange_set_participants(emvs_exchange* ex, const uint8_t* builds /* n_cams * n_ranks */)
{
  REQUIRE(ex && builds, EMVS_ERR_INVALID, "exchange_set_participants: NULL argument");
  REQUIRE(!ex->band_round, EMVS_ERR_STATE, "exchange_set_participants: a round is in progress");
  uint32_t masks[kMaxPeerCams] = {};
  int first_local = -1;
  for (int c = 0; c < ex->n_cams; ++c) {
    for (int r = 0; r < ex->n_ranks; ++r)
      if (builds[(size_t)c * ex->n_ranks + r]) masks[c] |= 1u << r;
    REQUIRE(masks[c] != 0u, EMVS_ERR_INVALID, "exchange_set_participants: a camera that no rank builds");
    if (first_local < 0 && ((masks[c] >> ex->rank) & 1u)) first_local = c;
  }
  REQUIRE(first_local >= 0, EMVS_ERR_INVALID, "exchange_set_participants: this rank builds no camera");
  for (int c = 0; c < ex->n_cams; ++c) ex->args.cam_ranks[c] = masks[c];
  ex->first_local_cam = first_local;
  return EMVS_OK;
}

int emvs_exchange_blob_bytes(const emvs_exchange* ex, size_t* out)
{
  REQUIRE(ex && out, EMVS_ERR_INVALID, "exchange_blob_bytes: NULL argument");
  *out = kIpcBytes * (size_t)(ex->n_cams + 2);
  return EMVS_OK;
}

int emvs_exchange_export(emvs_exchange* ex, uint8_t* blob)
{
  REQUIRE(ex && blob, EMVS_ERR_INVALID, "exchange_export: NULL argument");
  DeviceGuard guard(ex->ctx->device);
  cudaIpcMemHandle_t h;
  for (int c = 0; c < ex->n_cams; ++c) {
    CUDA_TRY(cudaIpcGetMemHandle(&h, (void*)ex->local_dsi[c]));
    memcpy(blob + kIpcBytes * c, &h, kIpcBytes);
  }
  CUDA_TRY(cudaIpcGetMemHandle(&h, ex->maps));
  memcpy(blob + kIpcBytes * ex->n_cams, &h, kIpcBytes);
  CUDA_TRY(cudaIpcGetMemHandle(&h, ex->flags));
  memcpy(blob + kIpcBytes * (ex->n_cams + 1), &h, kIpcBytes);
  return EMVS_OK;
}

int emvs_exchange_import(emvs_exchange* ex, const uint8_t* all)
{
  REQUIRE(ex && all, EMVS_ERR_INVALID, "exchange_import: NULL argument");
  REQUIRE(!ex->imported, EMVS_ERR_STATE, "exchange_import: already imported");
  DeviceGuard guard(ex->ctx->device);
  const size_t per_rank = kIpcBytes * (size_t)(ex->n_cams + 2);
  PeerArgs& A = ex->args;
  A.n_cams = ex->n_cams;
  A.n_ranks = ex->n_ranks;
  for (int r = 0; r < ex->n_ranks; ++r) {
    char* maps = nullptr;
    unsigned int* flags = nullptr;
    if (r == ex->rank) {
      for (int c = 0; c < ex->n_cams; ++c) A.dsi[c][r] = ex->local_dsi[c];
      maps = ex->maps;
      flags = ex->flags;
    } else {
      const uint8_t* b = all + per_rank * r;
      cudaIpcMemHandle_t h;
      void* p = nullptr;
      for (int c = 0; c < ex->n_cams; ++c) {
        memcpy(&h, b + kIpcBytes * c, kIpcBytes);
        CUDA_TRY(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        ex->opened.push_back(p);
        A.dsi[c][r] = (const float*)p;
      }
      memcpy(&h, b + kIpcBytes * ex->n_cams, kIpcBytes);
      CUDA_TRY(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
      ex->opened.push_back(p);
      maps = (char*)p;
      memcpy(&h, b + kIpcBytes * (ex->n_cams + 1), kIpcBytes);
      CUDA_TRY(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
      ex->opened.push_back(p);
      flags = (unsigned int*)p;
    }
    A.conf[r] = (float*)maps;
    A.depth[r] = (float*)(maps + ex->off_depth);
    A.idx[r] = maps + ex->off_idx;
    ex->flag_ptrs.p[r] = flags;
  }
  ex->imported = true;
  return EMVS_OK;
}

}  // extern "C" (templates cannot have C linkage)

template <int METHOD>
static int launch_peer(emvs_exchange* ex, const float* d_depths, uint32_t p_lo, uint32_t p_hi, long long timeout)
{
  emvs_context* ctx = ex->ctx;
  const uint32_t n_pix = ex->dimX * ex->dimY;
  const uint32_t band = p_hi - p_lo;
  const unsigned blocks = (band + 127) / 128;
  const int idx_bytes = ex->dimZ <= 256 ? 1 : 2;
  unsigned int* err = ex->flags + 2 * kMaxPeerRanks;
  // plane chunks: enough CTAs to cover the NVLink latency (EMVS_PEER_ZSPLIT overrides), >= 8 planes per chunk
  static const int zs_env = [] { const char* e = getenv("EMVS_PEER_ZSPLIT"); return e ? atoi(e) : 8; }();
  const uint32_t n_chunks = (uint32_t)std::max(1, std::min<int>(zs_env, (int)(ex->dimZ / 8)));
  const uint32_t per_chunk = (ex->dimZ + n_chunks - 1) / n_chunks;
  const uint32_t used = (ex->dimZ + per_chunk - 1) / per_chunk;
  int rc = grow(&ctx->d_fc_part, &ctx->fc_part_cap, (size_t)used * band * 8);
  if (rc) return rc;
  float* part_best = (float*)ctx->d_fc_part;
  uint32_t* part_k = (uint32_t*)((char*)ctx->d_fc_part + (size_t)used * band * 4);
  const dim3 grid(blocks, used);
#define LAUNCH(M, N)                                                                                                       \
  k_fuse_collapse_peer<M, N><<<grid, 128, 0, ctx->stream>>>(ex->args, ex->flags, ex->epoch, timeout, err, p_lo, p_hi, n_pix, \
                                                            ex->dimZ, per_chunk, part_best, part_k)
  switch (ex->n_cams) {
    case 1: LAUNCH(EMVS_FUSE_MAX, 1); break;
    case 2: LAUNCH(METHOD, 2); break;
    case 3: LAUNCH(METHOD, 3); break;
    case 4: LAUNCH(METHOD, 4); break;
    default: break;
  }
#undef LAUNCH
  k_peer_combine_store<<<(band + 255) / 256, 256, 0, ctx->stream>>>(ex->args, part_best, part_k, used, p_lo, p_hi, d_depths,
                                                                      idx_bytes, err);
  ctx->launches += 2;
  return EMVS_OK;
}

// Band round (emvs_exchange_begin + EMVS_BUILD_PEER_REDUCE builds): the band is already summed over the ranks
// in band_buf [cam][Z][band]; a purely LOCAL fuse + argmax sweep, then the band goes to every rank's maps.
template <int METHOD>
static int launch_band_sweep(emvs_exchange* ex, const float* d_depths)
{
  emvs_context* ctx = ex->ctx;
  const uint32_t p_lo = ex->row_lo * ex->dimX, p_hi = ex->row_hi * ex->dimX, band = p_hi - p_lo;
  const int idx_bytes = ex->dimZ <= 256 ? 1 : 2;
  const uint32_t n_chunks = std::max<uint32_t>(1, std::min<uint32_t>(8, ex->dimZ / 8));
  const uint32_t per_chunk = (ex->dimZ + n_chunks - 1) / n_chunks;
  const uint32_t used = (ex->dimZ + per_chunk - 1) / per_chunk;
  int rc = grow(&ctx->d_fc_part, &ctx->fc_part_cap, (size_t)used * band * 8);
  if (rc) return rc;
  float* part_best = (float*)ctx->d_fc_part;
  uint32_t* part_k = (uint32_t*)((char*)ctx->d_fc_part + (size_t)used * band * 4);
  FuseArgs A{};
  A.n = ex->n_cams;
  A.method = METHOD;
  for (int c = 0; c < ex->n_cams; ++c) A.g[c] = ex->band_buf + (size_t)c * ex->dimZ * band;
  const dim3 grid((band + 127) / 128, used);
#define LAUNCH(M, N) k_fuse_collapse_zsplit<M, N><<<grid, 128, 0, ctx->stream>>>(A, band, ex->dimZ, per_chunk, nullptr, part_best, part_k)
  switch (ex->n_cams) {
    case 1: LAUNCH(EMVS_FUSE_MAX, 1); break;
    case 2: LAUNCH(METHOD, 2); break;
    case 3: LAUNCH(METHOD, 3); break;
    case 4: LAUNCH(METHOD, 4); break;
    default: break;
  }
#undef LAUNCH
  k_peer_combine_store<<<(band + 255) / 256, 256, 0, ctx->stream>>>(ex->args, part_best, part_k, used, p_lo, p_hi, d_depths,
                                                                      idx_bytes, ex->flags + 16);
  ctx->launches += 2;
  return EMVS_OK;
}

extern "C" {

int emvs_exchange_begin(emvs_exchange* ex)
{
  REQUIRE(ex, EMVS_ERR_INVALID, "exchange is NULL");
  REQUIRE(ex->imported, EMVS_ERR_STATE, "exchange_begin: peers not imported (emvs_exchange_import)");
  REQUIRE(!ex->band_round, EMVS_ERR_STATE, "exchange_begin: the previous round was not finished with emvs_exchange_fuse_collapse");
  ex->epoch++;
  ex->band_round = true;
  ex->ctx->active_exchange = ex;
  return EMVS_OK;
}

int emvs_exchange_fuse_collapse(emvs_exchange* ex, int method, const float* d_depths)
{
  REQUIRE(ex, EMVS_ERR_INVALID, "exchange is NULL");
  REQUIRE(ex->imported, EMVS_ERR_STATE, "exchange_fuse_collapse: peers not imported (emvs_exchange_import)");
  REQUIRE(method >= EMVS_FUSE_MIN && method <= EMVS_FUSE_MAX, EMVS_ERR_INVALID, "Improper fusion method selected");
  emvs_context* ctx = ex->ctx;
  DeviceGuard guard(ctx->device);
  cudaStream_t st = ctx->stream;
  ex->args.method = method;
  // ~20 s at 2 GHz: a peer that never arrives raises the error word instead of hanging the GPU
  const long long timeout = 40000000000LL;
  const uint32_t row_lo = ex->row_lo, row_hi = ex->row_hi;
  if (ex->band_round) {
    // the builds of this round already reduced every slab of this rank's band into band_buf (stream-ordered
    // before this point): local sweep, then store the band into every rank's maps
    ex->band_round = false;
    ctx->active_exchange = nullptr;
    CUDA_TRY(cudaEventRecord(ctx->ev_comm_done, ctx->comm_stream));   // every slab reduce of this round
    CUDA_TRY(cudaStreamWaitEvent(st, ctx->ev_comm_done, 0));
    if (row_hi > row_lo) {
      int rc = EMVS_OK;
      switch (method) {
        case EMVS_FUSE_MIN: rc = launch_band_sweep<EMVS_FUSE_MIN>(ex, d_depths); break;
        case EMVS_FUSE_HM: rc = launch_band_sweep<EMVS_FUSE_HM>(ex, d_depths); break;
        case EMVS_FUSE_GM: rc = launch_band_sweep<EMVS_FUSE_GM>(ex, d_depths); break;
        case EMVS_FUSE_AM: rc = launch_band_sweep<EMVS_FUSE_AM>(ex, d_depths); break;
        case EMVS_FUSE_RMS: rc = launch_band_sweep<EMVS_FUSE_RMS>(ex, d_depths); break;
        default: rc = launch_band_sweep<EMVS_FUSE_MAX>(ex, d_depths); break;
      }
      if (rc) return rc;
    }
  } else {
  ex->epoch++;
  k_flag_signal<<<1, 32, 0, st>>>(ex->flag_ptrs, ex->n_ranks, ex->rank, 0, ex->epoch);   // "my partial DSIs are built"
  ctx->launches++;
  if (row_hi > row_lo) {
    int rc = EMVS_OK;
    switch (method) {
      case EMVS_FUSE_MIN: rc = launch_peer<EMVS_FUSE_MIN>(ex, d_depths, row_lo * ex->dimX, row_hi * ex->dimX, timeout); break;
      case EMVS_FUSE_HM: rc = launch_peer<EMVS_FUSE_HM>(ex, d_depths, row_lo * ex->dimX, row_hi * ex->dimX, timeout); break;
      case EMVS_FUSE_GM: rc = launch_peer<EMVS_FUSE_GM>(ex, d_depths, row_lo * ex->dimX, row_hi * ex->dimX, timeout); break;
      case EMVS_FUSE_AM: rc = launch_peer<EMVS_FUSE_AM>(ex, d_depths, row_lo * ex->dimX, row_hi * ex->dimX, timeout); break;
      case EMVS_FUSE_RMS: rc = launch_peer<EMVS_FUSE_RMS>(ex, d_depths, row_lo * ex->dimX, row_hi * ex->dimX, timeout); break;
      default: rc = launch_peer<EMVS_FUSE_MAX>(ex, d_depths, row_lo * ex->dimX, row_hi * ex->dimX, timeout); break;
    }
    if (rc) return rc;
  }
  }
  // "done": my peer loads of this round are finished and my band is stored everywhere; then wait for everyone
  k_flag_signal<<<1, 32, 0, st>>>(ex->flag_ptrs, ex->n_ranks, ex->rank, 1, ex->epoch);
  k_flag_wait<<<1, 32, 0, st>>>(ex->flags, ex->n_ranks, 1, ex->epoch, timeout, ex->flags + 2 * kMaxPeerRanks);
  ctx->launches += 2;
  CUDA_TRY(cudaGetLastError());
  return EMVS_OK;
}

int emvs_exchange_maps(const emvs_exchange* ex, float** d_conf, void** d_idx, float** d_depth)
{
  REQUIRE(ex, EMVS_ERR_INVALID, "exchange is NULL");
  if (d_conf) *d_conf = (float*)ex->maps;
  if (d_depth) *d_depth = (float*)(ex->maps + ex->off_depth);
  if (d_idx) *d_idx = ex->maps + ex->off_idx;
  return EMVS_OK;
}

int emvs_exchange_download(emvs_exchange* ex, float* conf, void* idx, float* depth)
{
  REQUIRE(ex, EMVS_ERR_INVALID, "exchange is NULL");
  emvs_context* ctx = ex->ctx;
  DeviceGuard guard(ctx->device);
  const size_t n_pix = (size_t)ex->dimX * ex->dimY;
  unsigned int err = 0;
  if (conf) CUDA_TRY(cudaMemcpyAsync(conf, ex->maps, n_pix * 4, cudaMemcpyDeviceToHost, ctx->stream));
  if (depth) CUDA_TRY(cudaMemcpyAsync(depth, ex->maps + ex->off_depth, n_pix * 4, cudaMemcpyDeviceToHost, ctx->stream));
  if (idx) CUDA_TRY(cudaMemcpyAsync(idx, ex->maps + ex->off_idx, n_pix * (ex->dimZ <= 256 ? 1 : 2), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(cudaMemcpyAsync(&err, ex->flags + 2 * kMaxPeerRanks, sizeof err, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  REQUIRE(err == 0, EMVS_ERR_STATE, "exchange: timed out waiting for a peer rank");
  return EMVS_OK;
}

}  // extern "C"
