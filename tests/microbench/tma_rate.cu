// Microbenchmark: per-SM TMA load throughput as a function of the box size, the number of boxes in flight and the
// residency of the source (L2 hit vs HBM).  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lcuda tma_rate.cu -o tma_rate
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

// tensor: [blocks][rows_per_block][64] fp16; box {64, R, Z}: Z consecutive blocks of R rows
struct P { int R, Z, stages, iters, window, cta_stride_blocks, ops_per_stage; };

__global__ void __launch_bounds__(64) k(const __grid_constant__ CUtensorMap map, P p, long long* out) {
  extern __shared__ uint8_t sm_raw[];
  const uint32_t base = (smem_u32(sm_raw) + 1023u) & ~1023u;
  __shared__ uint64_t bars[16];
  const uint32_t bar0 = smem_u32(bars);
  const int box_bytes = p.R * p.Z * 128;
  const int stage_bytes = box_bytes * p.ops_per_stage;
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) mbar_init(bar0 + 8 * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&map) : "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    long long t_issue = 0;
    const long long t0 = clock64();
    const int blk0 = blockIdx.x * p.cta_stride_blocks;
    int off = 0;
#define NEXT_BLK() (blk0 + off); off = (off + p.Z) % p.window
    // prime
    int issued = 0, done = 0;
    for (; issued < p.stages && issued < p.iters; ++issued) {
      const long long a = clock64();
      mbar_expect_tx(bar0 + 8 * (issued % p.stages), stage_bytes);
      for (int o = 0; o < p.ops_per_stage; ++o) {
        { const int blk = NEXT_BLK(); tma3(base + (issued % p.stages) * stage_bytes + o * box_bytes, &map, bar0 + 8 * (issued % p.stages), 0, 0, blk); }
      }
      t_issue += clock64() - a;
    }
    for (; done < p.iters; ++done) {
      const int s = done % p.stages;
      mbar_wait(bar0 + 8 * s, (done / p.stages) & 1);
      if (issued < p.iters) {
        const long long a = clock64();
        mbar_expect_tx(bar0 + 8 * s, stage_bytes);
        for (int o = 0; o < p.ops_per_stage; ++o) {
          { const int blk = NEXT_BLK(); tma3(base + s * stage_bytes + o * box_bytes, &map, bar0 + 8 * s, 0, 0, blk); }
        }
        t_issue += clock64() - a;
        ++issued;
      }
    }
    const long long t1 = clock64();
    out[blockIdx.x * 2] = t1 - t0;
    out[blockIdx.x * 2 + 1] = t_issue;
  }
}

int main() {
  CK(cudaSetDevice(0));
  const int rows_per_block = 64;           // one "block" = 64 rows x 128 B = 8 KB
  const long long nblocks = 1 << 17;       // 1 GiB
  void* buf;
  CK(cudaMalloc(&buf, nblocks * rows_per_block * 128));
  CK(cudaMemset(buf, 1, nblocks * rows_per_block * 128));
  long long* out;
  CK(cudaMalloc(&out, 148 * 2 * 8));
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  auto enc = (CUresult(*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                          const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                          CUtensorMapL2promotion, CUtensorMapFloatOOBfill))fn;
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  printf("%6s %3s %3s %4s %5s %8s | %10s %10s %10s %10s\n", "boxKB", "ops", "stg", "ctas", "src", "stageKB", "cyc/stage", "B/clk/SM", "issue/op", "chipTB/s");
  for (int src = 0; src < 2; ++src)
    for (int ctas : {148, 24})
      for (int ops : {1, 2})
        for (int boxkb : {8, 16, 32, 64}) {
          for (int stages : {2, 4}) {
            int R = boxkb >= 32 ? 256 : boxkb * 8, Z = boxkb >= 32 ? boxkb / 32 : 1;
            // tensor view for this box: dim1 = R rows (stride 128 B), dim2 = blocks of R rows
            if ((long long)stages * ops * boxkb > 200) continue;
            cuuint64_t gdim[3] = {64, (cuuint64_t)R, (cuuint64_t)(nblocks * rows_per_block / R)};
            cuuint64_t gstr[2] = {128, (cuuint64_t)R * 128};
            cuuint32_t box[3] = {64, (cuuint32_t)R, (cuuint32_t)Z};
            cuuint32_t es[3] = {1, 1, 1};
            CUtensorMap map;
            CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, buf, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); continue; }
            P p;
            p.R = R; p.Z = Z; p.stages = stages; p.iters = 64; p.ops_per_stage = ops;
            const long long blocks_total = nblocks * rows_per_block / R;   // units of R rows
            const int unit_kb = R * 128 / 1024;
            p.cta_stride_blocks = (int)(blocks_total / 148);
            p.window = src == 0 ? (256 / unit_kb) : p.cta_stride_blocks;   // L2 case: 256 KB private window per CTA
            if (p.window < Z) p.window = Z;
            p.window -= p.window % Z;
            std::vector<long long> h(148 * 2);
            for (int rep = 0; rep < 3; ++rep) {
              k<<<ctas, 64, (size_t)stages * ops * boxkb * 1024 + 1024>>>(map, p, out);
              CK(cudaDeviceSynchronize());
            }
            CK(cudaMemcpy(h.data(), out, ctas * 16, cudaMemcpyDeviceToHost));
            double mx = 0, iss = 0;
            for (int i = 0; i < ctas; ++i) { mx = h[2 * i] > mx ? h[2 * i] : mx; iss += h[2 * i + 1]; }
            const double cyc_stage = mx / p.iters;
            const double bpc = (double)ops * boxkb * 1024 / cyc_stage;
            printf("%6d %3d %3d %4d %5s %8d | %10.0f %10.1f %10.0f %10.2f\n", boxkb, ops, stages, ctas, src ? "hbm" : "l2", ops * boxkb,
                   cyc_stage, bpc, iss / ctas / p.iters / ops, bpc * ctas * 1.965e9 / 1e12);
          }
        }
  return 0;
}
