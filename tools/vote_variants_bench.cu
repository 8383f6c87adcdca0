// vote_variants_bench.cu — micro-benchmarks behind the design choices of the vote kernel: what the
// B200 gives for every way of getting a bilinear 2x2 vote (four float32 weights) into the DSI.  No geometry, no
// divisions: only the memory side of a vote, so every number is a CEILING for a vote kernel built that way.
//
//   red_lines      the product's pattern: one red.global.add.v4.f32 per vote, the 8 lanes that vote one event on the 8
//                  planes of a group hit ONE 128-byte line (4 lines per warp-level RED), footprint L2-resident.
//                  -> the RED-path ceiling that bench.py's roofline uses (profiles/red_port_ceiling.json)
//   red_random     the same instruction with 32 unrelated quads per warp (the ungrouped layout of round 1)
//   match_any      red_lines preceded by warp-level aggregation of equal addresses (__match_any_sync + shuffles, one RED
//                  per distinct quad); `dup` = fraction of lanes that share their quad with another lane of the warp
//   smem_f32       shared-memory privatisation: a CTA-private tile (8 planes x 48 x 32 voxels, 48 KB) receives the four
//                  float atomicAdds of a vote, and is flushed to global once at the end (red.global.add.v4.f32 per 4
//                  voxels).  atomicAdd(float) on shared memory compiles to a CAS loop on sm_100a (ATOMS.CAST.SPIN).
//   smem_u32       the same tile with 32-bit fixed-point integer atomics (native ATOMS.ADD)
//   stage_ldg / stage_tma   8 KB event tiles into shared memory: 256 threads x ld.global.nc + st.shared, against ONE
//                  cp.async.bulk (TMA, UBLKCP) per tile with an mbarrier, double-buffered
//
// Usage: vote_variants_bench [csv]      (prints one line per variant; `csv` = machine-readable)
// Build: make -C tools          ncu: ncu --metrics l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed,... ./vote_variants_bench
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x)                                                                                  \
  do {                                                                                         \
    cudaError_t e_ = (x);                                                                      \
    if (e_ != cudaSuccess) {                                                                   \
      fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(1);                                                                                 \
    }                                                                                          \
  } while (0)

__device__ __forceinline__ uint32_t hash32(uint32_t x)
{
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

__device__ __forceinline__ void red_add_v4(float4* addr, float a, float b, float c, float d)
{
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

constexpr int kThreads = 256;

// ---- RED straight to L2 ----------------------------------------------------------------------------------------------
// mode 0: 8 lanes share a 128-byte line (the product's grouped layout); mode 1: every lane its own random quad
__global__ void __launch_bounds__(kThreads) k_red(float4* quad, uint32_t n_lines, int iters, int mode)
{
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t group = tid >> 3, h = tid & 7u;
  const float w = 0.25f;
  for (int i = 0; i < iters; ++i) {
    const uint32_t key = (mode == 0 ? group : tid) * 0x9e3779b9u + (uint32_t)i;
    const uint32_t line = hash32(key) % n_lines;
    float4* q = quad + (size_t)line * 8 + (mode == 0 ? h : (hash32(key ^ 0x5bd1e995u) & 7u));
    red_add_v4(q, w, w, w, w);
  }
}

// warp-level aggregation of equal quad addresses before the RED.  dup_per_1024: how many lanes out of 1024 copy the
// quad of the lane 8 positions below them (same plane of another event of the warp)
__global__ void __launch_bounds__(kThreads) k_red_match_any(float4* quad, uint32_t n_lines, int iters, uint32_t dup_per_1024)
{
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t h = tid & 7u;
  for (int i = 0; i < iters; ++i) {
    uint32_t group = tid >> 3;
    if (lane >= 8 && (hash32(tid * 31u + (uint32_t)i) & 1023u) < dup_per_1024) group -= 1;   // same quad as the event "before"
    const uint32_t line = hash32(group * 0x9e3779b9u + (uint32_t)i) % n_lines;
    const unsigned long long addr = (unsigned long long)(quad + (size_t)line * 8 + h);
    float4 w = make_float4(0.25f, 0.25f, 0.25f, 0.25f);
    const unsigned peers = __match_any_sync(0xffffffffu, addr);
    const int leader = __ffs(peers) - 1;
    // fold the peers' weights onto the leader (at most 4 lanes can share a (line, plane) in this pattern)
    unsigned rest = peers & ~(1u << leader);
    while (rest) {
      const int src = __ffs(rest) - 1;
      const float x = __shfl_sync(peers, w.x, src), y = __shfl_sync(peers, w.y, src), z = __shfl_sync(peers, w.z, src),
                  t = __shfl_sync(peers, w.w, src);
      if ((int)lane == leader) { w.x += x; w.y += y; w.z += z; w.w += t; }
      rest &= rest - 1;
    }
    if ((int)lane == leader) red_add_v4(reinterpret_cast<float4*>(addr), w.x, w.y, w.z, w.w);
  }
}

// ---- shared-memory privatisation ---------------------------------------------------------------------------------------
constexpr int kTileX = 48, kTileY = 32, kTileZ = 8;
constexpr int kTileVox = kTileX * kTileY * kTileZ;   // 12288 voxels = 48 KB

template <typename T>
__global__ void __launch_bounds__(kThreads) k_smem_tile(float* out, int iters)
{
  extern __shared__ unsigned char s_raw[];
  T* tile = reinterpret_cast<T*>(s_raw);
  for (int i = threadIdx.x; i < kTileVox; i += kThreads) tile[i] = T(0);
  __syncthreads();
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t h = threadIdx.x & 7u;
  for (int i = 0; i < iters; ++i) {
    const uint32_t r = hash32((tid >> 3) * 0x9e3779b9u + (uint32_t)i);
    const uint32_t x = r % (kTileX - 1), y = (r >> 8) % (kTileY - 1);
    T* p = tile + (h * kTileY + y) * kTileX + x;
    const T w0 = T(1 + ((r >> 20) & 3u)), w1 = T(1 + ((r >> 22) & 3u));   // data-dependent weights (a constant would let
    atomicAdd(p, w0);                                                      // ptxas turn the integer case into a warp-aggregated
    atomicAdd(p + 1, w1);                                                  // ATOMS.POPC.INC)
    atomicAdd(p + kTileX, w1);
    atomicAdd(p + kTileX + 1, w0);
  }
  __syncthreads();
  // flush: the tile's voxels are added to the CTA's region of the global volume
  float* dst = out + (size_t)blockIdx.x * kTileVox;
  for (int i = threadIdx.x * 4; i < kTileVox; i += kThreads * 4)
    red_add_v4(reinterpret_cast<float4*>(dst + i), (float)tile[i], (float)tile[i + 1], (float)tile[i + 2], (float)tile[i + 3]);
}

// ---- staging 8 KB tiles into shared memory --------------------------------------------------------------------------------
constexpr uint32_t kTileBytes = 8192;

__global__ void __launch_bounds__(kThreads) k_stage_ldg(const float2* src, uint32_t n_tiles, float* sink)
{
  __shared__ float2 s[1024];
  float acc = 0.f;
  for (uint32_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    for (int i = 0; i < 4; ++i) {
      float2 v;
      const float2* p = src + (size_t)t * 1024 + i * kThreads + threadIdx.x;
      asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
      s[i * kThreads + threadIdx.x] = v;
    }
    __syncthreads();
    acc += s[(threadIdx.x * 37u) & 1023u].x;
    __syncthreads();
  }
  if (acc == 123.456f) sink[0] = acc;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(kThreads) k_stage_tma(const float2* src, uint32_t n_tiles, float* sink)
{
  __shared__ __align__(128) float2 s[2][1024];
  __shared__ uint64_t bar[2];
  if (threadIdx.x == 0) {
    for (int b = 0; b < 2; ++b) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[b])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](uint32_t t, int b) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[b])), "r"(kTileBytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(s[b])),
                 "l"(src + (size_t)t * 1024), "r"(kTileBytes), "r"(smem_u32(&bar[b]))
                 : "memory");
  };
  float acc = 0.f;
  uint32_t phase[2] = {0, 0};
  int b = 0;
  if (threadIdx.x == 0 && blockIdx.x < n_tiles) issue(blockIdx.x, 0);
  for (uint32_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    if (threadIdx.x == 0 && t + gridDim.x < n_tiles) issue(t + gridDim.x, b ^ 1);
    asm volatile(
        "{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}" ::"r"(
            smem_u32(&bar[b])),
        "r"(phase[b])
        : "memory");
    phase[b] ^= 1u;
    acc += s[b][(threadIdx.x * 37u) & 1023u].x;
    __syncthreads();
    b ^= 1;
  }
  if (acc == 123.456f) sink[0] = acc;
}

// -----------------------------------------------------------------------------------------------------------------------
struct Timer {
  cudaEvent_t a, b;
  Timer() { CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b)); }
  template <typename F>
  float ms(F f, int reps = 5)
  {
    f();   // warm-up
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
      CK(cudaEventRecord(a));
      f();
      CK(cudaEventRecord(b));
      CK(cudaEventSynchronize(b));
      float t;
      CK(cudaEventElapsedTime(&t, a, b));
      best = t < best ? t : best;
    }
    CK(cudaGetLastError());
    return best;
  }
};

int main(int argc, char** argv)
{
  const bool csv = argc > 1 && !strcmp(argv[1], "csv");
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  const int grid = sms * 8, iters = 512;
  const double votes = (double)grid * kThreads * iters;
  const size_t scratch_bytes = (size_t)79 << 20;   // one slab of the product: 16 planes x 4.9 MB, L2-resident
  const uint32_t n_lines = (uint32_t)(scratch_bytes / 128);
  float4* quad;
  CK(cudaMalloc(&quad, scratch_bytes));
  CK(cudaMemset(quad, 0, scratch_bytes));
  float* out;
  CK(cudaMalloc(&out, (size_t)grid * kTileVox * sizeof(float)));
  CK(cudaMemset(out, 0, (size_t)grid * kTileVox * sizeof(float)));
  Timer T;
  if (csv) printf("# %s, %d SMs; %d CTAs x %d threads x %d votes = %.1f Mvotes per launch, scratch %zu MB\nvariant,param,ms,Gvotes_per_s,payload_GB_per_s\n",
                  prop.name, sms, grid, kThreads, iters, votes / 1e6, scratch_bytes >> 20);
  auto report = [&](const char* name, const char* param, float ms, double n_votes) {
    const double gv = n_votes / (ms * 1e-3) / 1e9;
    if (csv) printf("%s,%s,%.4f,%.2f,%.1f\n", name, param, ms, gv, gv * 16.0);
    else printf("%-12s %-10s %8.4f ms  %8.2f Gvotes/s  %8.1f GB/s of vote payload\n", name, param, ms, gv, gv * 16.0);
  };
  report("red_lines", "-", T.ms([&] { k_red<<<grid, kThreads>>>(quad, n_lines, iters, 0); }), votes);
  report("red_random", "-", T.ms([&] { k_red<<<grid, kThreads>>>(quad, n_lines, iters, 1); }), votes);
  for (uint32_t dup : {0u, 128u, 256u, 512u}) {
    char p[32];
    snprintf(p, sizeof p, "dup=%.3f", dup / 1024.0);
    report("match_any", p, T.ms([&] { k_red_match_any<<<grid, kThreads>>>(quad, n_lines, iters, dup); }), votes);
  }
  CK(cudaFuncSetAttribute(k_smem_tile<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTileVox * 4));
  CK(cudaFuncSetAttribute(k_smem_tile<unsigned int>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTileVox * 4));
  // 4 CTAs per SM fit (4 x 48 KB); same number of votes per launch
  const int grid_s = sms * 4, iters_s = iters * 2;
  report("smem_f32", "-", T.ms([&] { k_smem_tile<float><<<grid_s, kThreads, kTileVox * 4>>>(out, iters_s); }), votes);
  report("smem_u32", "-", T.ms([&] { k_smem_tile<unsigned int><<<grid_s, kThreads, kTileVox * 4>>>(out, iters_s); }), votes);
  // staging: 4882 tiles of 8 KB (one camera's packets at 5 M events)
  const uint32_t n_tiles = 4882;
  float2* src;
  CK(cudaMalloc(&src, (size_t)n_tiles * kTileBytes));
  CK(cudaMemset(src, 0, (size_t)n_tiles * kTileBytes));
  float* sink;
  CK(cudaMalloc(&sink, 4));
  {
    const float ms = T.ms([&] { k_stage_ldg<<<sms * 8, kThreads>>>(src, n_tiles, sink); });
    if (csv) printf("stage_ldg,tiles=%u,%.4f,,%.1f\n", n_tiles, ms, n_tiles * (double)kTileBytes / (ms * 1e-3) / 1e9);
    else printf("%-12s %-10s %8.4f ms  %8.1f GB/s staged\n", "stage_ldg", "8KB tiles", ms, n_tiles * (double)kTileBytes / (ms * 1e-3) / 1e9);
  }
  {
    const float ms = T.ms([&] { k_stage_tma<<<sms * 8, kThreads>>>(src, n_tiles, sink); });
    if (csv) printf("stage_tma,tiles=%u,%.4f,,%.1f\n", n_tiles, ms, n_tiles * (double)kTileBytes / (ms * 1e-3) / 1e9);
    else printf("%-12s %-10s %8.4f ms  %8.1f GB/s staged\n", "stage_tma", "8KB tiles", ms, n_tiles * (double)kTileBytes / (ms * 1e-3) / 1e9);
  }
  return 0;
}
