// Microbenchmark 3: cost (cycles, issuing thread) of the individual producer-loop instructions.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1);} } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t b) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(b) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t ph) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(ph) : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t ph) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(ph) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void tma3(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(dst), "l"((uint64_t)m), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__global__ void __launch_bounds__(64) k(const __grid_constant__ CUtensorMap map, long long* out) {
  extern __shared__ uint8_t sm_raw[];
  const uint32_t base = (smem_u32(sm_raw) + 1023u) & ~1023u;
  __shared__ uint64_t bars[32];
  const uint32_t bar0 = smem_u32(bars);
  if (threadIdx.x == 0) {
    for (int s = 0; s < 32; ++s) mbar_init(bar0 + 8 * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&map) : "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    long long t[12];
    // (a) 16 TMA ops (8 KB each) back to back onto one barrier armed up front
    t[0] = clock64();
    mbar_expect_tx(bar0, 16 * 8192);
    t[1] = clock64();
    for (int i = 0; i < 16; ++i) tma3(base + i * 8192, &map, bar0, 0, 0, i);
    t[2] = clock64();
    while (!mbar_try(bar0, 0)) {}
    t[3] = clock64();
    // (b) 16 try_wait on the completed phase
    uint32_t acc = 0;
    for (int i = 0; i < 16; ++i) acc += mbar_try(bar0, 0);
    t[4] = clock64();
    for (int i = 0; i < 16; ++i) acc += mbar_test(bar0, 0);
    t[5] = clock64();
    // (c) 16 expect_tx on 16 different barriers
    for (int i = 1; i < 17; ++i) mbar_expect_tx(bar0 + 8 * i, 8192);
    t[6] = clock64();
    // (d) those 16 barriers each get one TMA op; then wait all
    for (int i = 1; i < 17; ++i) tma3(base + (i - 1) * 8192, &map, bar0 + 8 * i, 0, 0, 32 + i);
    t[7] = clock64();
    for (int i = 1; i < 17; ++i) while (!mbar_try(bar0 + 8 * i, 0)) {}
    t[8] = clock64();
    // (e) interleaved: expect_tx + tma per barrier (second phase of the same barriers)
    for (int i = 1; i < 17; ++i) { mbar_expect_tx(bar0 + 8 * i, 8192); tma3(base + (i - 1) * 8192, &map, bar0 + 8 * i, 0, 0, 64 + i); }
    t[9] = clock64();
    for (int i = 1; i < 17; ++i) while (!mbar_try(bar0 + 8 * i, 1)) {}
    t[10] = clock64();
    // (f) plain arrive
    for (int i = 17; i < 32; ++i) mbar_arrive(bar0 + 8 * i);
    t[11] = clock64();
    if (blockIdx.x == 0) { for (int i = 0; i < 12; ++i) out[i] = t[i]; out[12] = acc; }
    // (g) sustained ring: 8 slots, wait -> arm -> issue, timestamps per iteration; barriers 17..24 (phase 1 now: one arrive above)
    {
      const int S = 8;
      const long long g0 = clock64();
      for (int it = 0; it < 48; ++it) {
        const int s = it % S;
        const uint32_t b = bar0 + 8 * (17 + s);
        if (it >= S) while (!mbar_try(b, (((it / S) - 1) & 1) ^ 1)) {}
        mbar_expect_tx(b, 8192);
        tma3(base + s * 8192, &map, b, 0, 0, 100 + it);
        if (blockIdx.x == 0) out[16 + it] = clock64() - g0;
      }
    }
  }
}
int main() {
  CK(cudaSetDevice(0));
  const long long bytes = 1ll << 26;
  void* buf; CK(cudaMalloc(&buf, bytes)); CK(cudaMemset(buf, 1, bytes));
  long long* out; CK(cudaMalloc(&out, 128 * 8));
  void* fn = nullptr; cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  auto enc = (CUresult(*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                          const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                          CUtensorMapL2promotion, CUtensorMapFloatOOBfill))fn;
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  const int R = 64;
  cuuint64_t gdim[3] = {64, (cuuint64_t)R, (cuuint64_t)(bytes / (R * 128))};
  cuuint64_t gstr[2] = {128, (cuuint64_t)R * 128};
  cuuint32_t box[3] = {64, (cuuint32_t)R, 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUtensorMap map;
  enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, buf, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  long long h[128];
  for (int rep = 0; rep < 3; ++rep) {
    k<<<1, 64, 140 * 1024>>>(map, out);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h, out, 64 * 8, cudaMemcpyDeviceToHost));
    printf("rep %d: expect_tx(1) %lld | 16 tma same bar %lld (%.0f/op) | wait all %lld | 16 try_wait(done) %lld (%.0f) | 16 test_wait %lld (%.0f) | 16 expect_tx %lld (%.0f) | 16 tma 16 bars %lld (%.0f) | wait %lld | 16x(expect+tma) %lld (%.0f) | wait %lld | 15 arrive %lld (%.0f)\n",
           rep, h[1] - h[0], h[2] - h[1], (h[2] - h[1]) / 16.0, h[3] - h[2], h[4] - h[3], (h[4] - h[3]) / 16.0, h[5] - h[4],
           (h[5] - h[4]) / 16.0, h[6] - h[5], (h[6] - h[5]) / 16.0, h[7] - h[6], (h[7] - h[6]) / 16.0, h[8] - h[7], h[9] - h[8],
           (h[9] - h[8]) / 16.0, h[10] - h[9], h[11] - h[10], (h[11] - h[10]) / 15.0);
    printf("ring:"); for (int i = 0; i < 48; ++i) printf(" %lld", h[16 + i] - (i ? h[15 + i] : 0)); printf("\n");
  }
  return 0;
}
