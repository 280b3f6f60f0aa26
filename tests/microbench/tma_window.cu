// Microbenchmark 6: does TMA accept an OVERLAPPING-window tensor map (stride of dimension 1 = 16 bytes < the 128-byte
// extent of dimension 0)?  View of X8[B][H][W+8][8 fp16] as {64 elements = 8 consecutive pixels x 8 channels, W, H, B}.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1);} } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k(const __grid_constant__ CUtensorMap map, int x0, int y0, int b, __half* out) {
  extern __shared__ uint8_t sm_raw[];
  const uint32_t base = (smem_u32(sm_raw) + 1023u) & ~1023u;
  uint8_t* sm = sm_raw + (base - smem_u32(sm_raw));
  __shared__ uint64_t bar;
  const uint32_t b0 = smem_u32(&bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b0));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b0), "r"(16 * 8 * 128) : "memory");
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(base), "l"((uint64_t)&map), "r"(b0), "r"(0), "r"(x0), "r"(y0), "r"(b) : "memory");
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(b0), "r"(0) : "memory");
  }
  __syncthreads();
  // un-swizzle: row r (128 B) chunk c (16 B) is stored at chunk c ^ (r & 7)
  for (int i = threadIdx.x; i < 128 * 64; i += blockDim.x) {
    const int r = i / 64, e = i % 64, c = e / 8;
    out[i] = reinterpret_cast<const __half*>(sm + r * 128 + ((c ^ (r & 7)) * 16))[e % 8];
  }
}
int main() {
  CK(cudaSetDevice(0));
  const int B = 2, H = 16, W = 32, WP = W + 8;
  std::vector<__half> h((size_t)B * H * WP * 8);
  for (int b = 0; b < B; ++b) for (int y = 0; y < H; ++y) for (int x = 0; x < WP; ++x) for (int c = 0; c < 8; ++c)
    h[(((size_t)b * H + y) * WP + x) * 8 + c] = __float2half((float)(b * 1000 + y * 40 + x) + c * 0.0625f);
  __half* d; CK(cudaMalloc(&d, h.size() * 2)); CK(cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice));
  __half* out; CK(cudaMalloc(&out, 128 * 64 * 2));
  void* fn = nullptr; cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  auto enc = (CUresult(*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                          const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                          CUtensorMapL2promotion, CUtensorMapFloatOOBfill))fn;
  cuuint64_t gdim[4] = {64, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t gstr[3] = {16, (cuuint64_t)WP * 16, (cuuint64_t)H * WP * 16};
  cuuint32_t box[4] = {64, 16, 8, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUtensorMap map;
  CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode: %d\n", (int)r);
  if (r != CUDA_SUCCESS) return 0;
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 1024));
  const int x0 = 8, y0 = -1, b = 1;
  k<<<1, 128, 20 * 1024>>>(map, x0, y0, b, out);
  CK(cudaDeviceSynchronize());
  std::vector<__half> o(128 * 64);
  CK(cudaMemcpy(o.data(), out, o.size() * 2, cudaMemcpyDeviceToHost));
  int bad = 0;
  for (int ry = 0; ry < 8; ++ry) for (int rx = 0; rx < 16; ++rx) for (int e = 0; e < 64; ++e) {
    const int y = y0 + ry, x = x0 + rx;   // window of pixel x: physical pixels x .. x+7
    float want = 0.f;
    if (y >= 0 && y < H) want = (float)(b * 1000 + y * 40 + (x + e / 8)) + (e % 8) * 0.0625f;
    const float got = __half2float(o[(ry * 16 + rx) * 64 + e]);
    if (got != __half2float(__float2half(want))) { if (bad < 5) printf("mismatch row (%d,%d) e %d: got %f want %f\n", ry, rx, e, got, want); ++bad; }
  }
  printf("mismatches: %d of %d\n", bad, 128 * 64);
  return 0;
}
