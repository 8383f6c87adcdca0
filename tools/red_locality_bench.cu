// red_locality_bench.cu — does red.global.add.v4.f32 get cheaper when the lanes of one warp
// instruction share 32-byte sectors or even addresses?  Decides whether spatially sorting the
// events can lift the vote kernel above the L2 atomic-sector rate.
//
// Quad scratch layouts (one plane = QW x QH quad positions x 4 parity copies x 16 B):
//   L0 interleaved  float4 index = (qy*QW + qx)*4 + c          (sector = copies {0,1} or {2,3} of one quad position)
//   L1 split        float4 index = (c*QH + qy)*QW + qx         (sector = quad positions qx, qx^1 of one copy)
// Patterns (per warp instruction; centre random per warp and iteration, footprint = `planes` planes):
//   P0 random        32 independent random pixels
//   P1 window8x4     32 pixels random inside an 8x4-pixel window
//   P2 window4x2     32 pixels random inside a 4x2-pixel window (heavy duplicates)
//   P3 same          all 32 lanes the same pixel
//   P4 window16x8    32 pixels random inside a 16x8 window
//   P5 window32x8
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

template <int LAYOUT, int PATTERN>
__global__ void __launch_bounds__(256) k_red(float4* buf, uint32_t W, uint32_t H, uint32_t planes, int iters)
{
  const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t warp = gtid >> 5;
  const uint32_t QW = W / 2, QH = H / 2;
  for (int it = 0; it < iters; ++it) {
    const uint32_t hw = hash32(warp * 9781u + it * 6271u + 1u), hw2 = hash32(hw + 0x9e3779b9u);
    const uint32_t hl = hash32(gtid * 31u + it * 17u + 7u);
    uint32_t x, y;
    const uint32_t k = hw2 % planes;
    if (PATTERN == 0) { x = hl % (W - 2); y = (hl >> 16) % (H - 2); }
    else {
      const uint32_t wx = PATTERN == 1 ? 8 : PATTERN == 2 ? 4 : PATTERN == 3 ? 1 : PATTERN == 4 ? 16 : 32;
      const uint32_t wy = PATTERN == 1 ? 4 : PATTERN == 2 ? 2 : PATTERN == 3 ? 1 : 8;
      x = (hw % (W - 2 - wx)) + (hl % wx);
      y = ((hw >> 16) % (H - 2 - wy)) + ((hl >> 8) % wy);
    }
    const uint32_t c = (x & 1) | ((y & 1) << 1), qx = x >> 1, qy = y >> 1;
    size_t idx;
    if (LAYOUT == 0) idx = ((size_t)k * QH * QW + (size_t)qy * QW + qx) * 4 + c;
    else idx = (((size_t)k * 4 + c) * QH + qy) * QW + qx;
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(buf + idx), "f"(0.25f), "f"(0.25f), "f"(0.25f), "f"(0.25f) : "memory");
  }
}

template <int L, int P>
float run(float4* buf, uint32_t W, uint32_t H, uint32_t planes, int iters, int blocks)
{
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  k_red<L, P><<<blocks, 256>>>(buf, W, H, planes, 4);
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < 3; ++r) {
    CK(cudaEventRecord(a));
    k_red<L, P><<<blocks, 256>>>(buf, W, H, planes, iters);
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
  const uint32_t planes = 8;  // 37.5 MB: L2-resident
  float4* buf; CK(cudaMalloc(&buf, (size_t)W * H * 16 * planes)); CK(cudaMemset(buf, 0, (size_t)W * H * 16 * planes));
  printf("# %s; %d blocks x 256 thr x %d votes; footprint %u planes\nlayout,pattern,ms,Gvotes_per_s\n", prop.name, blocks, iters, planes);
  const char* ln[] = {"interleaved", "split"};
  const char* pn[] = {"random", "window8x4", "window4x2", "same_pixel", "window16x8", "window32x8"};
  for (int l = 0; l < 2; ++l)
    for (int p = 0; p < 6; ++p) {
      float ms = 0;
#define RUN(L, P) if (l == L && p == P) ms = run<L, P>(buf, W, H, planes, iters, blocks);
      RUN(0, 0) RUN(0, 1) RUN(0, 2) RUN(0, 3) RUN(0, 4) RUN(0, 5) RUN(1, 0) RUN(1, 1) RUN(1, 2) RUN(1, 3) RUN(1, 4) RUN(1, 5)
#undef RUN
      printf("%s,%s,%.3f,%.2f\n", ln[l], pn[p], ms, votes / ms * 1e-6);
      fflush(stdout);
    }
  return 0;
}
