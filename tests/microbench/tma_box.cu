// Microbenchmark 4: TMA latency / sustained rate of the 4-D activation boxes the low-resolution convolutions use
// ([B][H][W][C] fp16, box {64, TW, TH, TB}, taps shifted by -1..+1 -> partially out of bounds), single issuing thread.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1);} } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t b) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(b) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t ph) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(ph) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void tma4(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
               ::"r"(dst), "l"((uint64_t)m), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
// nops boxes issued back to back (one barrier each), then waited in order: reports issue time per op and completion times
__global__ void __launch_bounds__(64) k(const __grid_constant__ CUtensorMap map, int nops, int dx, int dy, int cchunks, int box_bytes, long long* out) {
  extern __shared__ uint8_t sm_raw[];
  const uint32_t base = (smem_u32(sm_raw) + 1023u) & ~1023u;
  __shared__ uint64_t bars[16];
  const uint32_t bar0 = smem_u32(bars);
  if (threadIdx.x == 0) {
    for (int s = 0; s < 16; ++s) mbar_init(bar0 + 8 * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&map) : "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    long long ti[9], tc[9];
    for (int rep = 0; rep < 2; ++rep) {   // second repetition = warm
      const long long t0 = clock64();
      for (int i = 0; i < nops; ++i) {
        mbar_expect_tx(bar0 + 8 * i, box_bytes);
        tma4(base + i * box_bytes, &map, bar0 + 8 * i, (i % cchunks) * 64, dx, dy, 2 * (blockIdx.x % 4));
        ti[i] = clock64() - t0;
      }
      for (int i = 0; i < nops; ++i) {
        while (!mbar_try(bar0 + 8 * i, rep & 1)) {}
        tc[i] = clock64() - t0;
      }
    }
    if (blockIdx.x == 0) for (int i = 0; i < nops; ++i) { out[i] = ti[i]; out[8 + i] = tc[i]; }
  }
}
int main() {
  CK(cudaSetDevice(0));
  void* fn = nullptr; cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  auto enc = (CUresult(*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                          const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                          CUtensorMapL2promotion, CUtensorMapFloatOOBfill))fn;
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  long long* out; CK(cudaMalloc(&out, 16 * 8));
  struct Cfg { int B, H, W, C, TW, TH, TB; const char* name; };
  Cfg cfgs[] = {{8, 8, 8, 384, 8, 8, 2, "8x8 C384 box{64,8,8,2}"}, {8, 16, 16, 320, 16, 8, 1, "16x16 C320 box{64,16,8,1}"},
                {8, 128, 128, 128, 16, 8, 1, "128x128 C128 box{64,16,8,1}"}, {8, 8, 8, 384, 8, 8, 1, "8x8 C384 box{64,8,8,1} (64 rows)"}};
  for (const Cfg& c : cfgs) {
    const size_t bytes = (size_t)c.B * c.H * c.W * c.C * 2;
    void* buf; CK(cudaMalloc(&buf, bytes)); CK(cudaMemset(buf, 1, bytes));
    cuuint64_t gdim[4] = {(cuuint64_t)c.C, (cuuint64_t)c.W, (cuuint64_t)c.H, (cuuint64_t)c.B};
    cuuint64_t gstr[3] = {(cuuint64_t)c.C * 2, (cuuint64_t)c.W * c.C * 2, (cuuint64_t)c.H * c.W * c.C * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)c.TW, (cuuint32_t)c.TH, (cuuint32_t)c.TB};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUtensorMap map;
    CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, buf, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); continue; }
    const int box_bytes = 128 * c.TW * c.TH * c.TB;
    for (int ctas : {1, 148})
      for (int sh = 0; sh < 2; ++sh) {
        const int dx = sh ? -1 : 0, dy = sh ? -1 : 0;
        long long h[16];
        for (int rep = 0; rep < 2; ++rep) { k<<<ctas, 64, 8 * box_bytes + 1024>>>(map, 8, dx, dy, c.C / 64, box_bytes, out); CK(cudaDeviceSynchronize()); }
        CK(cudaMemcpy(h, out, 16 * 8, cudaMemcpyDeviceToHost));
        printf("%-36s ctas %3d shift %2d | issued at:", c.name, ctas, dx);
        for (int i = 0; i < 8; ++i) printf(" %5lld", h[i]);
        printf(" | complete at:");
        for (int i = 0; i < 8; ++i) printf(" %5lld", h[8 + i]);
        printf("\n");
      }
    CK(cudaFree(buf));
  }
  return 0;
}
