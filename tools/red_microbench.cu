// red_microbench.cu — how fast can a B200 resolve bilinear-vote read-modify-writes?
// Decides the layout of the vote kernel: four scalar RED.F32 on the canonical
// layout vs two RED.F32x2 vs one RED.F32x4 on the quad layout, as a function of the footprint
// the votes are spread over (L2-resident slab vs HBM-resident volume) and of warp locality.
//
//   ./red_microbench [votes_per_thread]
// prints one CSV line per (variant, footprint, pattern): Gvotes/s.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t hash32(uint32_t x)
{
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}

// pattern 0: every lane an independent random voxel; 1: lanes of a warp hit 32 consecutive
// quads / pixels (best-case locality); 2: random within a 64x64 window per warp
template <int VARIANT, int PATTERN>
__global__ void __launch_bounds__(256) k_red(float* buf, uint32_t W, uint32_t H, uint32_t planes, int iters)
{
  const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t lane = threadIdx.x & 31, warp = gtid >> 5;
  const size_t plane = (size_t)W * H;
  for (int it = 0; it < iters; ++it) {
    uint32_t x, y, k;
    if (PATTERN == 0) {
      const uint32_t h = hash32(gtid * 9781u + it * 6271u + 1u), h2 = hash32(h + 0x9e3779b9u);
      x = h % (W - 2); y = (h >> 16) % (H - 2); k = h2 % planes;
    } else if (PATTERN == 1) {
      const uint32_t h = hash32(warp * 9781u + it * 6271u + 1u), h2 = hash32(h + 0x9e3779b9u);
      x = (h % (W - 66)) + 2 * lane; y = (h >> 16) % (H - 2); k = h2 % planes;
    } else {
      const uint32_t h = hash32(warp * 9781u + it * 6271u + 1u), h2 = hash32(h + 0x9e3779b9u);
      const uint32_t hl = hash32(gtid * 31u + it * 17u + 7u);
      x = (h % (W - 66)) + (hl & 63); y = ((h >> 16) % (H - 66)) + ((hl >> 8) & 63); k = h2 % planes;
    }
    const float w = 0.25f;
    if (VARIANT == 0) {            // canonical layout, 4 scalar REDs
      float* g = buf + k * plane + (size_t)y * W + x;
      atomicAdd(g, w); atomicAdd(g + 1, w); atomicAdd(g + W, w); atomicAdd(g + W + 1, w);
    } else if (VARIANT == 1) {     // canonical layout, 2 x v2 (x forced even)
      float* g = buf + k * plane + (size_t)y * W + (x & ~1u);
      asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(g), "f"(w), "f"(w) : "memory");
      asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(g + W), "f"(w), "f"(w) : "memory");
    } else if (VARIANT == 2) {     // quad layout: 4 parity copies interleaved, 1 x v4
      const uint32_t QW = W / 2;
      float* g = buf + k * plane * 4 + ((((size_t)(y >> 1) * QW + (x >> 1)) * 4 + ((x & 1) | ((y & 1) << 1))) * 4);
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(g), "f"(w), "f"(w), "f"(w), "f"(w) : "memory");
    } else if (VARIANT == 3) {     // quad layout: 4 parity copies in separate sub-planes, 1 x v4
      const uint32_t QW = W / 2, QH = H / 2;
      const uint32_t c = (x & 1) | ((y & 1) << 1);
      float* g = buf + k * plane * 4 + ((size_t)c * QW * QH + (size_t)(y >> 1) * QW + (x >> 1)) * 4;
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(g), "f"(w), "f"(w), "f"(w), "f"(w) : "memory");
    }
  }
}

template <int V, int P>
float run(float* buf, uint32_t W, uint32_t H, uint32_t planes, int iters, int blocks)
{
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  k_red<V, P><<<blocks, 256>>>(buf, W, H, planes, 4);
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < 3; ++r) {
    CK(cudaEventRecord(a));
    k_red<V, P><<<blocks, 256>>>(buf, W, H, planes, iters);
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    if (ms < best) best = ms;
  }
  CK(cudaGetLastError());
  return best;
}

int main(int argc, char** argv)
{
  const int iters = argc > 1 ? atoi(argv[1]) : 128;
  const uint32_t W = 640, H = 480;
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int blocks = prop.multiProcessorCount * 8;
  const double votes = (double)blocks * 256 * iters;
  printf("# %s, %d SMs, L2 %d MB; %d blocks x 256 thr x %d votes = %.1f Mvotes per launch\n", prop.name,
         prop.multiProcessorCount, prop.l2CacheSize >> 20, blocks, iters, votes * 1e-6);
  const size_t max_bytes = (size_t)1300 << 20;
  float* buf; CK(cudaMalloc(&buf, max_bytes)); CK(cudaMemset(buf, 0, max_bytes));
  printf("variant,pattern,planes,footprint_MB,ms,Gvotes_per_s\n");
  const char* vn[] = {"scalar4", "v2x2", "v4_interleaved", "v4_split"};
  const char* pn[] = {"random", "warp_row", "warp_window"};
  const uint32_t plane_counts[] = {1, 4, 8, 16, 32, 64, 256};
  for (uint32_t planes : plane_counts) {
    for (int v = 0; v < 4; ++v) {
      const size_t per_plane = (size_t)W * H * 4 * (v >= 2 ? 4 : 1);
      if (per_plane * planes > max_bytes) continue;
      for (int p = 0; p < 3; ++p) {
        float ms = 0;
#define RUN(V, P) if (v == V && p == P) ms = run<V, P>(buf, W, H, planes, iters, blocks);
        RUN(0, 0) RUN(0, 1) RUN(0, 2) RUN(1, 0) RUN(1, 1) RUN(1, 2) RUN(2, 0) RUN(2, 1) RUN(2, 2) RUN(3, 0) RUN(3, 1) RUN(3, 2)
#undef RUN
        printf("%s,%s,%u,%.1f,%.3f,%.2f\n", vn[v], pn[p], planes, per_plane * planes / 1048576.0, ms, votes / ms * 1e-6);
        fflush(stdout);
      }
    }
  }
  return 0;
}
