// Microbenchmark 5: issue cost and execution rate of tcgen05.mma (cta_group::1, kind::f16, M=128) from one thread.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1);} } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c)); }
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t ph) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(ph) : "memory");
  return ok != 0;
}
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ uint32_t make_idesc_f16(int n) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24); }
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// mode 0: nmma MMAs back to back, one commit.  mode 1: groups of 4 MMAs + commit each (like a pipeline stage)
__global__ void __launch_bounds__(128) k(int N, int nmma, int mode, long long* out) {
  extern __shared__ uint8_t sm_raw[];
  const uint32_t base = (smem_u32(sm_raw) + 1023u) & ~1023u;
  __shared__ uint64_t bars[8];
  __shared__ uint32_t s_tmem;
  const uint32_t bar0 = smem_u32(bars);
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(bar0 + 8 * i, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = s_tmem;
  if (threadIdx.x == 32) {
    const uint32_t idesc = make_idesc_f16(N);
    const uint64_t da = make_desc_sw128(base), db = make_desc_sw128(base + 32768);
    long long t_issue_end = 0;
    const long long t0 = clock64();
    if (mode == 0) {
      for (int i = 0; i < nmma; ++i) umma(tmem + (i & 1) * 256, da + (i & 3) * 2, db + (i & 3) * 2, idesc, 1u);
      t_issue_end = clock64();
      commit(bar0);
    } else {
      for (int i = 0; i < nmma; i += 4) {
#pragma unroll
        for (int j = 0; j < 4; ++j) umma(tmem, da + j * 2, db + j * 2, idesc, 1u);
        commit(bar0 + 8 * (1 + ((i >> 2) & 3)));
      }
      t_issue_end = clock64();
      commit(bar0);
    }
    while (!mbar_try(bar0, 0)) {}
    const long long t1 = clock64();
    out[blockIdx.x * 2] = t_issue_end - t0;
    out[blockIdx.x * 2 + 1] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512)); }
}
int main() {
  CK(cudaSetDevice(0));
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  long long* out; CK(cudaMalloc(&out, 148 * 16));
  long long h[4];
  for (int mode = 0; mode < 2; ++mode)
    for (int N : {64, 128, 192, 256})
      for (int nmma : {4, 16, 64, 256}) {
        for (int rep = 0; rep < 2; ++rep) { k<<<1, 128, 96 * 1024>>>(N, nmma, mode, out); CK(cudaDeviceSynchronize()); }
        CK(cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost));
        printf("mode %d N %3d nmma %3d: issue %6lld cyc (%5.1f / mma)  total %6lld cyc (%5.1f / mma; tensor floor %d)\n", mode, N, nmma, h[0],
               (double)h[0] / nmma, h[1], (double)h[1] / nmma, N / 2);
      }
  return 0;
}
