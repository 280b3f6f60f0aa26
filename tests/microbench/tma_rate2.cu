// Microbenchmark 2: does the per-op TMA issue cost (~220 cycles) scale out over several issuing warps of one CTA?
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1);} } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t b) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(b) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t ph) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(ph) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t ph) { while (!mbar_try(bar, ph)) {} }
__device__ __forceinline__ void tma3(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(dst), "l"((uint64_t)m), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
struct P { int R, stages, iters, window, cta_stride, ops, mode; };
// mode 0: each of W warps runs its own ring (own barriers, own smem).  mode 1: one ring; a stage's `ops` boxes are
// issued by `ops` different warps (warp w issues box w), all signalling the stage's barrier; warp 0 waits + arms.
__global__ void __launch_bounds__(128) k(const __grid_constant__ CUtensorMap map, P p, long long* out) {
  extern __shared__ uint8_t sm_raw[];
  const uint32_t base = (smem_u32(sm_raw) + 1023u) & ~1023u;
  __shared__ uint64_t bars[64];
  __shared__ volatile int go[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, W = blockDim.x >> 5;
  const uint32_t bar0 = smem_u32(bars);
  const int box_bytes = p.R * 128;
  if (threadIdx.x == 0) {
    for (int s = 0; s < 64; ++s) mbar_init(bar0 + 8 * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&map) : "memory");
  }
  __syncthreads();
  const long long t0 = clock64();
  if (p.mode == 0) {
    if (lane == 0) {
      const int stage_bytes = box_bytes * p.ops;
      const uint32_t my = base + warp * p.stages * stage_bytes;
      const uint32_t mb = bar0 + 8 * warp * 8;
      const int blk0 = blockIdx.x * p.cta_stride + warp * (p.cta_stride / W);
      int off = 0, issued = 0;
      for (int done = -p.stages; done < p.iters; ++done) {
        if (done >= 0) mbar_wait(mb + 8 * (done % p.stages), (done / p.stages) & 1);
        if (issued < p.iters) {
          const int s = issued % p.stages;
          mbar_expect_tx(mb + 8 * s, stage_bytes);
          for (int o = 0; o < p.ops; ++o) {
            tma3(my + s * stage_bytes + o * box_bytes, &map, mb + 8 * s, 0, 0, blk0 + off);
            off = (off + 1) % p.window;
          }
          ++issued;
        }
      }
    }
  } else {
    // one ring, ops boxes per stage, box o issued by warp o; every warp tracks the ring on its own (waits the stage's
    // barrier before re-using the slot), warp 0 arms the barrier with the full byte count first
    if (lane == 0 && warp < p.ops) {
      const int stage_bytes = box_bytes * p.ops;
      const int blk0 = blockIdx.x * p.cta_stride + warp * (p.cta_stride / p.ops);
      int off = 0, issued = 0;
      for (int done = -p.stages; done < p.iters; ++done) {
        if (done >= 0) mbar_wait(bar0 + 8 * (done % p.stages), (done / p.stages) & 1);
        if (issued < p.iters) {
          const int s = issued % p.stages;
          if (warp == 0) mbar_expect_tx(bar0 + 8 * s, stage_bytes);   // tx-count may go transiently negative: allowed
          tma3(base + s * stage_bytes + warp * box_bytes, &map, bar0 + 8 * s, 0, 0, blk0 + off);
          off = (off + 1) % p.window;
          ++issued;
        }
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = clock64() - t0;
}
int main() {
  CK(cudaSetDevice(0));
  const long long bytes = 1ll << 30;
  void* buf; CK(cudaMalloc(&buf, bytes)); CK(cudaMemset(buf, 1, bytes));
  long long* out; CK(cudaMalloc(&out, 148 * 8));
  void* fn = nullptr; cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  auto enc = (CUresult(*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                          const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                          CUtensorMapL2promotion, CUtensorMapFloatOOBfill))fn;
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  printf("%4s %5s %3s %3s %3s %4s %5s | %10s %10s %10s\n", "mode", "boxKB", "W", "ops", "stg", "ctas", "src", "cyc/box", "B/clk/SM", "chipTB/s");
  for (int mode = 0; mode < 2; ++mode)
    for (int src = 0; src < 2; ++src)
      for (int ctas : {148, 24})
        for (int boxkb : {8, 16, 32})
          for (int W : {1, 2, 4})
            for (int ops : {1, 2, 4}) {
              if (mode == 1 && (W != 4 || ops == 1)) continue;
              if (mode == 0 && ops == 4) continue;
              for (int stages : {2, 4, 6, 8}) {
              const int ring_kb = (mode == 0 ? W : 1) * stages * ops * boxkb;
              if (ring_kb > 200) continue;
              if (src == 1 || ctas == 148) continue;
              const int R = boxkb * 8;
              cuuint64_t gdim[3] = {64, (cuuint64_t)R, (cuuint64_t)(bytes / (R * 128))};
              cuuint64_t gstr[2] = {128, (cuuint64_t)R * 128};
              cuuint32_t box[3] = {64, (cuuint32_t)R, 1};
              cuuint32_t es[3] = {1, 1, 1};
              CUtensorMap map;
              CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, buf, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
              if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); continue; }
              P p; p.R = R; p.stages = stages; p.iters = 128; p.ops = ops; p.mode = mode;
              const long long units = bytes / (R * 128);
              p.cta_stride = (int)(units / 148);
              p.window = src == 0 ? (128 / boxkb) : p.cta_stride / 4;
              std::vector<long long> h(148);
              if (mode == 1) for (int rep = 0; rep < 3; ++rep) { k<<<ctas, 128 * 1, (size_t)ring_kb * 1024 + 1024>>>(map, p, out); CK(cudaDeviceSynchronize()); }
              // note: W warps used = blockDim/32 = 4 always in mode 0?  launch with W warps instead
              if (mode == 0) { for (int rep = 0; rep < 3; ++rep) { k<<<ctas, 32 * W, (size_t)ring_kb * 1024 + 1024>>>(map, p, out); CK(cudaDeviceSynchronize()); } }
              CK(cudaMemcpy(h.data(), out, ctas * 8, cudaMemcpyDeviceToHost));
              double mx = 0; for (int i = 0; i < ctas; ++i) mx = h[i] > mx ? h[i] : mx;
              const double nbox = (double)p.iters * ops * (mode == 0 ? W : 1);
              const double bpc = nbox * boxkb * 1024 / mx;
              printf("%4d %5d %3d %3d %3d %4d %5s | %10.0f %10.1f %10.2f  cyc/stage %6.0f\n", mode, boxkb, W, ops, stages, ctas, src ? "hbm" : "l2", mx / nbox, bpc, bpc * ctas * 1.965e9 / 1e12, mx / p.iters);
              }
            }
  return 0;
}
