// Implicit-GEMM convolution on the Blackwell tensor cores: TMA -> shared memory -> tcgen05.mma -> TMEM.
//
// Same GEMM view and fused epilogues as igemm_hmma.cuh (reference: Block / ResnetBlock / Upsample /
// res_conv / attention output, epsilonparam/modules/network_components.py:34-42, 83-114, 117-139),
// for stride-1 convolutions:
//   * an output tile is 128 pixels = TB x TH x TW (images x rows x cols); one 4-D TMA box
//     {64 channels, TW, TH, TB} per (tap, 64-channel chunk) lands as a K-major SWIZZLE_128B operand
//     [128 rows][64 x fp16]; conv padding and ragged tile edges are TMA out-of-bounds zero fill.
//   * the weight tile [C_out rows][64 x fp16] of the same K chunk is a 2-D TMA box from the repacked
//     weights [chunk][C_out][64].
//   * one elected thread issues tcgen05.mma (cta_group::1, kind::f16, M=128, N=C_out split into pieces of
//     <= 256) accumulating fp32 in TMEM; tcgen05.commit releases smem stages / publishes the accumulator.
//   * 4 epilogue warps own one pixel row each (TMEM lane == GEMM row): bias, channel LayerNorm (exact
//     two-pass variance from the fp32 accumulator), ReLU, timestep shift or residual, fp16 NHWC store.
//   * persistent CTAs walk tiles with a static stride; the accumulator is double-buffered in TMEM when
//     2*C_out <= 512 columns so the epilogue of tile i overlaps the MMAs of tile i+1.
//   * pipeline stages carry one activation box and 1..3 weight tiles: vertical reuse (one box of TH+kh-1 tile rows
//     feeds all kh vertical taps) or a "dual" stage (x_hi feeds the W_hi and W_lo tiles of a 3-pass convolution).
//   * K segments may accumulate into a SECOND TMEM accumulator: a ResnetBlock's res_conv, or the identity residual
//     as x_hi*I + x_lo*I — the epilogue then takes the residual from TMEM in fp32 (RT) instead of global memory.
//   * epilogue global I/O is row-contiguous through per-warp staging buffers; arithmetic in packed fp32 pairs; 64-
//     column CTAs keep the whole row in registers (N64); low-resolution layers run as 64-column slices whose CTAs form a
//     thread-block cluster and exchange LayerNorm statistics over distributed shared memory.
// Warp roles: 0 = TMA producer, 1 = MMA issuer (+TMEM alloc), 2..5 = epilogue (TMEM lane quadrant = warp%4).
#pragma once
#include <cuda.h>

#include <type_traits>

#include "common.cuh"
#include "igemm_hmma.cuh"  // EpiKind

namespace cdc {

struct TcSeg {
  int cpt;       // 64-channel chunks per tap
  int kh, kw;
  int dy0, dx0;
  int nchunk;
  int vr;        // vertical reuse: 1, or kh = one activation box of TH+kh-1 tile rows feeds all kh vertical taps
  int a_bytes;   // bytes of this segment's activation box
  int q0;        // first weight chunk of the segment
  int acc;       // TMEM accumulator this segment adds to (0 = convolution, 1 = fused res_conv)
  int wshared;   // per-image-weight ops only: this segment's weights are shared by all images (identity residual)
  // weight tiles per pipeline stage and the distance (descriptor units) between the activation views they multiply:
  //   vertical reuse: nw = vr, avstep = one tile row of pixels;   dual: nw = 2, avstep = 0 — the W_hi and W_lo passes of
  //   a 3-pass ("trunk") convolution share one activation load (x_hi) and issue two MMAs from it
  int nw, avstep, dual;
};

// Division by a launch-constant divisor without the ~35-instruction integer-division sequence (the tile loops of all
// three warp roles decode a tile index with five of them):  q = (umulhi(n, mul) + n) >> shr,  n < 2^31.
struct FastDiv {
  uint32_t mul, shr;
};
inline FastDiv make_fastdiv(uint32_t d) {
  FastDiv f{1u, 0u};
  if (d <= 1) return f;
  uint32_t s = 0;
  while ((1ull << s) < d) ++s;
  f.shr = s;
  f.mul = (uint32_t)((((1ull << s) - d) << 32) / d + 1ull);
  return f;
}
__device__ __forceinline__ int fdiv(int n, FastDiv d) {
  return (int)((__umulhi((uint32_t)n, d.mul) + (uint32_t)n) >> d.shr);
}

struct TcConvParams {
  int TW, TH, TB;            // tile = TB x TH x TW <= 128 pixels (rows beyond it are masked)
  int b_off;                 // byte offset of the weight tiles inside a pipeline stage (>= largest activation box)
  int vr_max;                // weight tiles per stage
  int total_sc;              // pipeline stages' worth of work per tile ("super-chunks"), split by k_splits
  int tiles_x, tiles_y, tiles_b;
  FastDiv fd_upt, fd_ks, fd_tpp, fd_txy, fd_tx;   // units per tile, K splits, tiles per phase, tiles_x*tiles_y, tiles_x
  int B, H, W;               // tile-space (output) extents; source pixel = stride * tile pixel + tap offset
  int stride;
  int nseg;
  TcSeg seg[kMaxSeg];
  int total_chunks;
  int Ntot;                  // C_out
  int n_split, n_piece;      // UMMA N pieces (n_piece <= 256, n_piece % 16 == 0)
  int nbuf;                  // TMEM accumulator buffers (1 or 2)
  int stages;                // smem pipeline depth
  int phases;                // 1, or 4 for the transposed-conv output phases
  int w_rows_per_phase;      // weight rows between phases ( = total_chunks * Ntot )
  int w_rows_per_image;      // per-image weights (attention): rows between images, else 0
  int epi;
  __half* out;               // NHWC fp16 [B, out_H, out_W, Ntot]
  __half* out_lo;            // optional compensation tensor (see ConvParams)
  int out_H, out_W, out_sy, out_sx;
  const float* bias;
  const float* ln_g;
  const float* ln_b;
  const float* shift;
  int shift_stride;
  const __half* res;
  int res_C0;
  const __half* res2;
  const __half* res_lo;
  const __half* res2_lo;
  const float2* stats_in;
  const float* aff_u;        // [B][aff_parts][Ntot] (per image): partial sums, added in order by the epilogue
  const float* aff_c;
  int aff_parts;
  float2* stats_out;
  // Sliced mode (low-resolution layers: few output tiles, long K, wide N).  The work of one output tile is spread
  // over n_slices x k_splits CTAs: each computes columns [slice*Nc, +Nc) over a K sub-range and stores its fp32
  // partial tile to raw[k_split][out_pixel][Ntot]; ln_rows_kernel then sums the K partials in a fixed order and
  // applies the fused epilogue (deterministic, no atomics).  Nc == Ntot, n_slices == k_splits == 1 otherwise.
  long long* dbg_clk;        // profiling aid (CDC_DBG_CLK=1 + cdc_engine_profile_ops): per-role cycle counters of CTA 0
  int Nc;
  int n_slices, k_splits;
  // Fused column slices (epi != EPI_RAW, n_slices > 1, k_splits == 1): the CTA with blockIdx % n_slices == s computes
  // columns [s*Nc, +Nc) of every tile it visits with the fused epilogue.  For the LayerNorm epilogues the n_slices CTAs
  // of a tile form a thread-block cluster (cluster_n == n_slices) and exchange per-row partial statistics through
  // distributed shared memory (st.async + mbarrier), so no fp32 partial tile ever goes to HBM.
  // EPI_BIAS, C_out == 64 only: additionally (or, with skip_out, instead) store fp16(LayerNorm(row) * ln_g + ln_b) to
  // ln_out — the last Upsample feeds nothing but the final LayerNorm + 7x7 convolution, which then needs no
  // normalisation pass of its own (final_tc.cuh).
  __half* ln_out;
  int skip_out;
  // Fused res_conv (64-column CTAs, EPI_LN_RES): segments with acc == 1 accumulate the ResnetBlock's 1x1 shortcut
  // convolution into TMEM columns [Nc, 2Nc) of the tile's buffer; the epilogue adds them (+ res_bias) as the residual in
  // fp32 — no separate res_conv launch, no residual round trip through HBM.
  int res_acc;
  const float* res_bias;
  int cluster_n;
  int xchg_stats;            // second exchange for stats_out (row statistics of the stored LayerNorm output)
  float* raw;
  long long raw_split_stride;
  // Fused finish of the K-split form (tile_ctr != nullptr): after storing its fp32 partial tile a CTA signals the tile's
  // arrival counter; once all n_slices * k_splits units of the tile have arrived (all CTAs of the launch are co-resident:
  // the grid never exceeds the SM slots), every unit finishes its share of the tile's rows — K partials summed in a fixed
  // order, then the fused epilogue `fin_epi` (bias | LayerNorm+ReLU+shift | LayerNorm+ReLU+residual) exactly as
  // ln_rows_kernel does — so the separate ln_rows_kernel launch disappears.  tile_ctr[2t] = arrivals, [2t+1] = departures;
  // the last departing unit zeroes both (the counters are launch-invariant, CUDA-graph replays included).
  unsigned int* tile_ctr;
  int fin_epi;
  int dbg_skip_epi;          // timing experiment only (CDC_DBG_EPI=1, results are garbage): the epilogue releases the
                             // accumulator and skips its arithmetic and stores — the ceiling of any epilogue optimisation
};

constexpr int EPI_RAW = 4;

struct TcMaps {
  CUtensorMap a[kMaxSeg];   // activations of segment s: {64 channels, W, H, B}
  CUtensorMap b[kMaxSeg];   // weights of segment s: {64, kw*cpt*C_out, kh, phase | image}
};

namespace tc {

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a pipeline bug traps (reported as a CUDA error) instead of hanging the GPU.
__device__ __noinline__ void mbar_wait_slow(uint32_t bar, uint32_t parity) {
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 6000000000LL) {  // ~3 s at 1.9 GHz
      printf("cdc igemm_tc: mbarrier wait timed out (block %d thread %d bar 0x%x parity %u)\n", blockIdx.x,
             threadIdx.x, bar, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  mbar_wait_slow(bar, parity);
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
// Predicated forms (issue iff pred != 0): the elected lane issues without a divergent branch — no BSSY / BSYNC / WARPSYNC
// bookkeeping per pipeline stage in the producer and MMA loops, which are single-warp latency chains.
__device__ __forceinline__ void mbar_expect_tx_if(uint32_t pred, uint32_t bar, uint32_t bytes) {
  asm volatile(
      "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %0, 0;\n\t"
      "@q mbarrier.arrive.expect_tx.shared::cta.b64 _, [%1], %2;\n\t}"
      ::"r"(pred), "r"(bar), "r"(bytes)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d_if(uint32_t pred, uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1,
                                               int c2, int c3) {
  asm volatile(
      "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %0, 0;\n\t"
      "@q cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%1], [%2, {%4, %5, %6, %7}], [%3];\n\t}"
      ::"r"(pred), "r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

// K-major SWIZZLE_128B operand descriptor: rows of 128 bytes, 8-row swizzle atoms 1024 bytes apart.
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);  // start address, bits [0,14)
  d |= (uint64_t)1 << 16;                       // leading byte offset (unused for swizzled K-major) = 16 B
  d |= (uint64_t)(1024 >> 4) << 32;             // stride byte offset between 8-row groups
  d |= (uint64_t)1 << 46;                       // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                       // SWIZZLE_128B
  return d;
}
// kind::f16 instruction descriptor: fp16 A/B (K-major), fp32 accumulate, M = 128, N = n.
__device__ __forceinline__ uint32_t make_idesc_f16(int n) {
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// the same instruction with the descriptors given as their low words (start address | LBO) over a shared high word
__device__ __forceinline__ void umma_f16_lo(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi,
                                            uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %3};\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}"
      ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_f16_lo_if(uint32_t pred, uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi,
                                               uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %3};\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "setp.ne.b32 q, %6, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}"
      ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate), "r"(pred)
      : "memory");
}
__device__ __forceinline__ void umma_commit_if(uint32_t pred, uint32_t bar) {
  asm volatile(
      "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %0, 0;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%1];\n\t}"
      ::"r"(pred), "r"(bar)
      : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void prefetch_l2(const void* ptr) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr));
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,"
      "%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---- thread-block cluster helpers (distributed shared memory exchange of the sliced LayerNorm statistics) ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa(uint32_t local_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
// 8-byte remote store that signals the destination CTA's mbarrier with its byte count
__device__ __forceinline__ void st_async_f2(uint32_t remote_addr, float a, float b, uint32_t remote_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f32 [%0], {%1, %2}, [%3];"
               ::"r"(remote_addr), "f"(a), "f"(b), "r"(remote_bar)
               : "memory");
}

// issue-only variant (pair with tmem_wait_ld) so that several loads are in flight before the single wait
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,"
      "%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Packed fp32 pairs (Blackwell FADD2 / FMUL2 / FFMA2): halves the instruction count of the epilogue arithmetic.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk(float a, float b) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ f32x2 pku(uint32_t a, uint32_t b) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(a), "r"(b));
  return r;
}
__device__ __forceinline__ float2 upk(f32x2 v) {
  float2 r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
  return r;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}

}  // namespace tc

// Role cycle counters (profiling aid, see TcConvParams::dbg_clk) cost ~10 % of the epilogue's stall samples: compiled in
// only with -DCDC_ROLE_CLK (CDC_ROLE_CLK=1 python -m cdc_compression_b200.build).
#ifdef CDC_ROLE_CLK
#define CDC_CLK() clock64()
#else
#define CDC_CLK() 0LL
#endif

constexpr int kTcThreads = 192;
constexpr int kTcMaxStages = 8;

// Epilogue staging: every epilogue warp owns a 2 KB transposition buffer (32 rows x 64 bytes) and a table of its 32
// rows' output pixel indices, so that global loads / stores of the epilogue are issued row-contiguously (8 rows x 64
// bytes per warp instruction) instead of one 16-byte piece per thread at a row-sized stride.
constexpr int kEpiStageBytes = 2048;
constexpr int kEpiPixBytes = 32 * 8;
// bytes after the pipeline stages: epilogue vectors (bias | g | b | res_bias: 4 x Nc floats), barriers, staging, pixel tables, and
// per epilogue warp 2 x Nc floats for the per-image vectors (timestep shift | attention affine u, c)
// + cluster exchange buffers: [sets][2 (tile parity)][cluster_n][128 rows] float2
__host__ __device__ inline int tc_xchg_bytes(int cluster_n, int sets) { return sets * 2 * cluster_n * 128 * 8; }
__host__ __device__ inline int tc_tail_bytes(int Nc, int cluster_n = 0, int xchg_sets = 0) {
  return 4 * Nc * 4 + 256 + 4 * (kEpiStageBytes + kEpiPixBytes) + 4 * (2 * Nc * 4) + tc_xchg_bytes(cluster_n, xchg_sets);
}
// dynamic smem: [stages][A box | B tiles] (1024-aligned) + tail
__host__ __device__ inline int tc_stage_bytes(int b_off, int vr_max, int Nc) { return b_off + vr_max * Nc * 128; }
__host__ __device__ inline int tc_smem_bytes(int stage_bytes, int stages, int Nc, int cluster_n = 0, int xchg_sets = 0) {
  return 1024 /*alignment slack*/ + stages * stage_bytes + tc_tail_bytes(Nc, cluster_n, xchg_sets);
}

// byte offset of 16-byte chunk `chunk` (0..3) of 64-byte row `row` in a staging buffer (XOR swizzle: conflict-free for
// both the row-per-lane and the 4-lanes-per-row access patterns)
__device__ __forceinline__ uint32_t stg_off(int row, int chunk) {
  return (uint32_t)(row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4));
}
// Every lane holds the 64 bytes w[0..3] of ITS row; row r goes to dst + pix[r] * row_bytes (skipped when pix[r] < 0).
__device__ __forceinline__ void warp_store_rows64(uint8_t* stg, const long long* pixtab, int lane, const uint4 (&w)[4],
                                                  uint8_t* dst, long long row_bytes) {
#pragma unroll
  for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4*>(stg + stg_off(lane, j)) = w[j];
  __syncwarp();
  uint4 v[4];
  long long px[4];
  const int ch = lane & 3;
#pragma unroll
  for (int i = 0; i < 4; ++i) {   // all shared-memory reads first, so that their latencies overlap
    const int row = i * 8 + (lane >> 2);
    v[i] = *reinterpret_cast<const uint4*>(stg + stg_off(row, ch));
    px[i] = pixtab[row];
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (px[i] >= 0) *reinterpret_cast<uint4*>(dst + px[i] * row_bytes + ch * 16) = v[i];
  __syncwarp();
}
// Row-contiguous loads of 32 rows x 64 bytes (4 lanes per row); the data lands in the loading lanes' registers ...
__device__ __forceinline__ void warp_load_rows64(const long long* pixtab, int lane, uint4 (&r)[4], const uint8_t* src,
                                                 long long row_bytes) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = i * 8 + (lane >> 2), ch = lane & 3;
    const long long px = pixtab[row];
    r[i] = px >= 0 ? *reinterpret_cast<const uint4*>(src + px * row_bytes + ch * 16) : make_uint4(0u, 0u, 0u, 0u);
  }
}
// ... and is transposed through the staging buffer so that r[0..3] become the 64 bytes of the lane's OWN row.
__device__ __forceinline__ void warp_rows64_to_own(uint8_t* stg, int lane, uint4 (&r)[4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) *reinterpret_cast<uint4*>(stg + stg_off(i * 8 + (lane >> 2), lane & 3)) = r[i];
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 4; ++j) r[j] = *reinterpret_cast<const uint4*>(stg + stg_off(lane, j));
  __syncwarp();
}
// 8 fp32 values -> 8 fp16 (hi) and the fp16 rounding remainders (lo)
__device__ __forceinline__ void pack_out8(const float (&o)[8], uint4& hi, uint4& lo, bool want_lo) {
  hi.x = pack_half2(o[0], o[1]);
  hi.y = pack_half2(o[2], o[3]);
  hi.z = pack_half2(o[4], o[5]);
  hi.w = pack_half2(o[6], o[7]);
  if (!want_lo) return;
  float2 f;
  f = unpack_half2(hi.x); lo.x = pack_half2(o[0] - f.x, o[1] - f.y);
  f = unpack_half2(hi.y); lo.y = pack_half2(o[2] - f.x, o[3] - f.y);
  f = unpack_half2(hi.z); lo.z = pack_half2(o[4] - f.x, o[5] - f.y);
  f = unpack_half2(hi.w); lo.w = pack_half2(o[6] - f.x, o[7] - f.y);
}

// ------------------------------------------------------------------------------------------------
// Row-wise finish of the K-split ("sliced") form: one warp per output pixel sums the K-split partials (fixed order) and
// applies the same fused epilogues as the in-kernel path (bias | LayerNorm+ReLU+shift | LayerNorm+ReLU+residual).
// Runs inside igemm_tc_kernel (fused finish, TcConvParams::tile_ctr) or as ln_rows_kernel (CDC_FUSE_LNROWS=0).
// ------------------------------------------------------------------------------------------------
struct LnRowsParams {
  const float* raw;
  int k_splits;
  long long split_stride;
  int N;
  long long rows;            // output pixels
  int pix_per_image;
  int epi;                   // EPI_BIAS | EPI_LN_SHIFT | EPI_LN_RES
  const float* bias;
  const float* ln_g;
  const float* ln_b;
  const float* shift;
  int shift_stride;
  const __half* res;
  int res_C0;
  const __half* res2;
  const __half* res_lo;
  const __half* res2_lo;
  __half* out;
  __half* out_lo;
  float2* stats_out;
  int Ntot;                  // == N (name expected by load_res2)
};

// One output pixel `pix` of image `img` by one warp.  The partials were written by OTHER SMs during this launch in the
// fused form: they are read with ld.global.cg (L2), never through L1.
// HOIST: the per-channel vectors, the shift and the residual of the row are loaded BEFORE the K partials arrive (three
// dependent memory round trips fewer; ~48 more registers): used for the small grids of the lowest levels, where the kernel
// is nothing but a chain of L2 latencies.  (Hoisting in the large-grid case costs occupancy: 8 -> 21 us on 8 192 rows.)
template <bool HOIST>
__device__ __forceinline__ void ln_rows_row(const LnRowsParams& p, long long pix, int img, int lane) {
  const int N = p.N, iters = N >> 6;
  const bool ln = p.epi != EPI_BIAS;
  const float* shift = (p.epi == EPI_LN_SHIFT && p.shift) ? p.shift + (size_t)img * p.shift_stride : nullptr;
  const bool has_res = p.res && p.epi != EPI_LN_SHIFT;
  float2 hg[HOIST ? 6 : 1], hb[HOIST ? 6 : 1], hs[HOIST ? 6 : 1], hr[HOIST ? 6 : 1];
  if (HOIST) {
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      const int c = (i * 32 + lane) * 2;
      const bool on = i < iters;
      hg[i] = (on && ln) ? *reinterpret_cast<const float2*>(p.ln_g + c) : make_float2(1.f, 1.f);
      hb[i] = (on && ln) ? *reinterpret_cast<const float2*>(p.ln_b + c) : make_float2(0.f, 0.f);
      hs[i] = (on && shift) ? *reinterpret_cast<const float2*>(shift + c) : make_float2(0.f, 0.f);
      hr[i] = (on && has_res) ? load_res2(p, pix, c) : make_float2(0.f, 0.f);
    }
  }
  float2 v[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const int c = (i * 32 + lane) * 2;
    v[i] = (i < iters && p.bias) ? make_float2(p.bias[c], p.bias[c + 1]) : make_float2(0.f, 0.f);
  }
  // fixed summation order (deterministic); K partials are fetched four at a time so their latencies overlap
  for (int k0 = 0; k0 < p.k_splits; k0 += 4) {
    float2 r[4][6];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const float* src = p.raw + (size_t)(k0 + kk) * p.split_stride + (size_t)pix * N;
#pragma unroll
      for (int i = 0; i < 6; ++i)
        r[kk][i] = (i < iters && k0 + kk < p.k_splits) ? __ldcg(reinterpret_cast<const float2*>(src + (i * 32 + lane) * 2))
                                                       : make_float2(0.f, 0.f);
    }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk)
#pragma unroll
      for (int i = 0; i < 6; ++i)
        if (i < iters) {
          v[i].x += r[kk][i].x;
          v[i].y += r[kk][i].y;
        }
  }
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < 6; ++i)
    if (i < iters) sum += v[i].x + v[i].y;
  float mean = 0.f, rstd = 1.f;
  if (p.epi != EPI_BIAS) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    mean = sum / (float)N;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < 6; ++i)
      if (i < iters) {
        const float d0 = v[i].x - mean, d1 = v[i].y - mean;
        sq += d0 * d0 + d1 * d1;
      }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    rstd = 1.f / sqrtf(sq / (float)N + 1e-5f);
  }
  float osum = 0.f, osq = 0.f;
#pragma unroll
  for (int i = 0; i < 6; ++i)
    if (i < iters) {
      const int c = (i * 32 + lane) * 2;
      float y0 = v[i].x, y1 = v[i].y;
      if (ln) {
        const float2 g2 = HOIST ? hg[i] : make_float2(p.ln_g[c], p.ln_g[c + 1]);
        const float2 b2 = HOIST ? hb[i] : make_float2(p.ln_b[c], p.ln_b[c + 1]);
        y0 = fmaxf((y0 - mean) * rstd * g2.x + b2.x, 0.f);
        y1 = fmaxf((y1 - mean) * rstd * g2.y + b2.y, 0.f);
      }
      if (shift) {
        const float2 s2 = HOIST ? hs[i] : make_float2(shift[c], shift[c + 1]);
        y0 += s2.x;
        y1 += s2.y;
      }
      if (has_res) {
        const float2 rr = HOIST ? hr[i] : load_res2(p, pix, c);
        y0 += rr.x;
        y1 += rr.y;
      }
      const float2 q = unpack_half2(store_out2(p, (size_t)pix * N + c, y0, y1));
      osum += q.x + q.y;
      osq += q.x * q.x + q.y * q.y;
    }
  if (p.stats_out) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      osum += __shfl_xor_sync(0xffffffffu, osum, o);
      osq += __shfl_xor_sync(0xffffffffu, osq, o);
    }
    if (lane == 0) {
      const float m = osum / (float)N;
      const float var = fmaxf(osq / (float)N - m * m, 0.f);
      p.stats_out[pix] = make_float2(m, 1.f / sqrtf(var + 1e-5f));
    }
  }
}

// EPI: fused epilogue (compile-time, prunes the others); OCC: CTAs per SM the register budget is sized for.
// Per-channel epilogue vectors of a C_out == 64 layer passed BY VALUE (kernel parameters live in the constant bank: the
// packed arithmetic reads them through uniform registers instead of shared-memory loads on the latency-bound path).
struct TcVecs64 {
  float bias[64], g[64], b[64], rbias[64];   // conv bias, LayerNorm gain / offset, bias of a fused res_conv
};
struct TcNoVecs {
  int unused;
};

// N64: LayerNorm epilogue specialised for 64-column CTAs (row held in registers, single TMEM pass).
//   1 = epilogue vectors in shared memory (column slices of wider layers), 2 = vectors in the constant bank (C_out == 64).
// RT: the residual is the second TMEM accumulator of a fused res_conv (TcConvParams::res_acc; 64-column CTAs only).
template <int EPI, int OCC, int N64, bool RT = false>
__global__ void __launch_bounds__(kTcThreads, OCC)
igemm_tc_kernel(const __grid_constant__ TcMaps maps, const TcConvParams p,
                const __grid_constant__ typename std::conditional<N64 == 2, TcVecs64, TcNoVecs>::type kc) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const int stage_bytes = tc_stage_bytes(p.b_off, p.vr_max, p.Nc);
  const int vs = p.Nc;                                                     // stride of the epilogue vectors
  float* s_vec = reinterpret_cast<float*>(smem + p.stages * stage_bytes);  // bias | g | b | res_bias: 4 x Nc
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem + p.stages * stage_bytes + 4 * vs * 4);
  uint8_t* s_stage = smem + p.stages * stage_bytes + 4 * vs * 4 + 256;     // 4 x kEpiStageBytes, 4 pixel tables, 4 x 2Nc floats
  // barriers: full[8] empty[8] tmem_full[2] tmem_empty[2]; then the TMEM base address word
  const uint32_t bar_full = smem_u32(s_bar);
  const uint32_t bar_empty = bar_full + 8 * kTcMaxStages;
  const uint32_t bar_tfull = bar_empty + 8 * kTcMaxStages;
  const uint32_t bar_tempty = bar_tfull + 16;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bar + 2 * kTcMaxStages + 4);
  const uint32_t bar_x = bar_full + 8 * 24;   // cluster exchange barriers: [set][parity]
  float2* s_xchg = reinterpret_cast<float2*>(s_stage + 4 * (kEpiStageBytes + kEpiPixBytes) + 4 * (2 * vs * 4));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int N = p.Nc;   // accumulator columns owned by this CTA (== Ntot unless sliced)
  // fused column slice of this CTA (constant: the grid is a multiple of n_slices, units are tile-major / slice-minor)
  const bool fused_slice = EPI != EPI_RAW && p.n_slices > 1;
  const int col0 = fused_slice ? (int)(blockIdx.x % p.n_slices) * N : 0;
  const int Nacc = p.res_acc ? 2 * N : N;   // TMEM columns of one tile buffer
  int tmem_cols = 32;
  while (tmem_cols < p.nbuf * Nacc) tmem_cols <<= 1;

  pdl_launch_dependents();
  if (threadIdx.x == 0) {
    for (int i = 0; i < p.nseg; ++i) {
      tc::prefetch_tmap(&maps.a[i]);
      tc::prefetch_tmap(&maps.b[i]);
    }
    for (int s = 0; s < p.stages; ++s) {
      tc::mbar_init(bar_full + 8 * s, 1);
      tc::mbar_init(bar_empty + 8 * s, 1);
    }
    for (int b = 0; b < p.nbuf; ++b) {
      tc::mbar_init(bar_tfull + 8 * b, 1);
      tc::mbar_init(bar_tempty + 8 * b, 128);
    }
    for (int b = 0; b < 4; ++b) tc::mbar_init(bar_x + 8 * b, 1);
    tc::fence_barrier_init();
  }
  if (warp == 1) {  // TMEM allocation (whole warp), address lands in shared memory
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)),
                 "r"(tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  // epilogue vectors -> smem (all threads)
  for (int i = threadIdx.x; i < (EPI == EPI_RAW ? 0 : N); i += kTcThreads) {
    s_vec[i] = p.bias ? p.bias[col0 + i] : 0.f;
    s_vec[vs + i] = p.ln_g ? p.ln_g[col0 + i] : 1.f;
    s_vec[2 * vs + i] = p.ln_b ? p.ln_b[col0 + i] : 0.f;
    s_vec[3 * vs + i] = p.res_bias ? p.res_bias[col0 + i] : 0.f;
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  if (p.cluster_n > 1) tc::cluster_sync_all();   // every peer's exchange barriers are initialised before anyone signals them
  const uint32_t tmem_base = *s_tmem;
  const bool clk = p.dbg_clk != nullptr && blockIdx.x == 0;
  const long long k_t1 = CDC_CLK();
  pdl_wait();   // everything above (barriers, TMEM, constant vectors) overlapped the previous kernel's tail

  const int tiles_per_phase = p.tiles_x * p.tiles_y * p.tiles_b;
  const int total_tiles = tiles_per_phase * p.phases;
  const int units_per_tile = p.n_slices * p.k_splits;
  const int total_units = total_tiles * units_per_tile;

  if (warp == 0) {
    // =============================== TMA producer ===============================
    // The whole warp walks the loop converged (uniform control flow keeps the index math in uniform registers);
    // one elected lane issues the copies.
    const bool leader = tc::elect_one();
    const uint32_t lead = leader ? 1u : 0u;
    int stage = 0;
    uint32_t phase = 0;
    long long c_wait = 0, c_n = 0;
    const long long c_t0 = CDC_CLK();
    for (int u = blockIdx.x; u < total_units; u += gridDim.x) {
      const int t = fdiv(u, p.fd_upt);
      const int su = u - t * units_per_tile;
      const int slice = fdiv(su, p.fd_ks), ksp = su - slice * p.k_splits;
      const int sc0 = fdiv(ksp * p.total_sc, p.fd_ks), sc1 = fdiv((ksp + 1) * p.total_sc, p.fd_ks);
      const int ph = fdiv(t, p.fd_tpp);
      int r = t - ph * tiles_per_phase;
      const int tb = fdiv(r, p.fd_txy);
      r -= tb * p.tiles_x * p.tiles_y;
      const int ty = fdiv(r, p.fd_tx), tx = r - ty * p.tiles_x;
      const int x0 = tx * p.TW, y0 = ty * p.TH, b0 = tb * p.TB;
      const int dyp = p.phases > 1 ? (ph >> 1) - 1 : 0, dxp = p.phases > 1 ? (ph & 1) - 1 : 0;
      const int wsel = p.phases > 1 ? ph : (p.w_rows_per_image ? b0 : 0);   // weight set: phase | image
      // Only the unit's own K range [sc0, sc1) is visited: a K-split unit used to walk (and skip) every super-chunk before
      // and after its range — ~220 cycles per skipped iteration (divergence bookkeeping), up to 10 k cycles before the
      // first load of the units with the highest K-split index (measured with per-stage timestamps, round 2).
      int seg_base = 0;
      for (int s = 0; s < p.nseg; ++s) {
        const TcSeg sg = p.seg[s];
        const CUtensorMap* mA = &maps.a[s];
        const CUtensorMap* mB = &maps.b[s];
        const uint32_t tx_bytes = (uint32_t)(sg.a_bytes + sg.nw * N * 128);
        const int per_ky = sg.kw * sg.cpt;
        const int nsc_s = (sg.kh / sg.vr) * per_ky;
        const int lo = max(sc0 - seg_base, 0), hi = min(sc1 - seg_base, nsc_s);
        seg_base += nsc_s;
        if (lo >= hi) continue;
        int kyo = 0, kx = 0, cc = 0;
        if (lo > 0) {   // K split: first super-chunk of the range inside this segment
          const int kyi = lo / per_ky, r2 = lo - kyi * per_ky;
          kyo = kyi * sg.vr;
          kx = r2 / sg.cpt;
          cc = r2 - kx * sg.cpt;
        }
        for (int j = lo; j < hi; ++j) {
          const long long w0 = CDC_CLK();
          tc::mbar_wait(bar_empty + 8 * stage, phase ^ 1);
          c_wait += CDC_CLK() - w0;
          ++c_n;
          {
            const uint32_t sA = base + stage * stage_bytes;
            const uint32_t sB = sA + p.b_off;
            const uint32_t full = bar_full + 8 * stage;
            tc::mbar_expect_tx_if(lead, full, tx_bytes);
            // one activation box: TH + vr - 1 tile rows starting at the first vertical tap
            tc::tma_load_4d_if(lead, sA, mA, full, cc * 64, x0 * p.stride + kx + sg.dx0 + dxp,
                               y0 * p.stride + kyo + sg.dy0 + dyp, b0);
            // the weight tiles of the vr vertical taps it feeds: [piece][tap][n_piece rows]
            for (int pc = 0; pc < p.n_split; ++pc)
              tc::tma_load_4d_if(lead, sB + pc * sg.nw * p.n_piece * 128, mB, full, 0,
                                 (kx * sg.cpt + cc) * p.Ntot + slice * N + pc * p.n_piece, kyo,
                                 (sg.wshared || sg.dual) ? 0 : wsel);
          }
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1;
          }
          if (++cc == sg.cpt) {
            cc = 0;
            if (++kx == sg.kw) {
              kx = 0;
              kyo += sg.vr;
            }
          }
        }
      }
    }
    if (clk && leader) {
      p.dbg_clk[0] += CDC_CLK() - c_t0;
      p.dbg_clk[1] += c_wait;
      p.dbg_clk[2] += c_n;
      p.dbg_clk[12] += c_t0 - k_t1;
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    // Whole warp converged, one elected lane issues (the same lane every time: tcgen05.commit tracks the MMAs of
    // the thread that executes it).
    const bool leader = tc::elect_one();
    const uint32_t lead = leader ? 1u : 0u;
    const uint32_t idesc = tc::make_idesc_f16(p.n_piece);
    const uint32_t desc_hi = (uint32_t)(tc::make_desc_sw128(0) >> 32);
    const uint32_t b_step = (uint32_t)(p.n_piece * 128) >> 4;    // one weight tile
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    long long c_wf = 0, c_we = 0;
    const long long c_t0 = CDC_CLK();
    for (int u = blockIdx.x; u < total_units; u += gridDim.x, ++it) {
      const int ksp = u - fdiv(u, p.fd_ks) * p.k_splits;
      const int sc0 = fdiv(ksp * p.total_sc, p.fd_ks), sc1 = fdiv((ksp + 1) * p.total_sc, p.fd_ks);
      const int buf = p.nbuf == 2 ? (it & 1) : 0;
      const uint32_t use = (uint32_t)(p.nbuf == 2 ? (it >> 1) : it);  // how many times this buffer was used before
      const long long w0 = CDC_CLK();
      tc::mbar_wait(bar_tempty + 8 * buf, (use & 1) ^ 1);
      c_we += CDC_CLK() - w0;
      tc::tc_fence_after();
      const uint32_t d_tmem0 = tmem_base + (uint32_t)(buf * Nacc);
      uint32_t acc_started = 0;   // bit a: accumulator a already holds a partial sum of this tile
      int seg_base = 0;
      for (int s = 0; s < p.nseg; ++s) {
        const int vr = p.seg[s].vr;
        const int nw = p.seg[s].nw;
        const uint32_t avs = (uint32_t)p.seg[s].avstep;
        const int sacc = p.seg[s].acc;
        const uint32_t d_tmem = d_tmem0 + (uint32_t)(sacc * N);
        const int nsc = (p.seg[s].kh / vr) * p.seg[s].kw * p.seg[s].cpt;
        const int lo = max(sc0 - seg_base, 0), hi = min(sc1 - seg_base, nsc);   // the unit's K range inside this segment
        seg_base += nsc;
        for (int i = lo; i < hi; ++i) {
          const long long w1 = CDC_CLK();
          tc::mbar_wait(bar_full + 8 * stage, phase);
          c_wf += CDC_CLK() - w1;
          tc::tc_fence_after();
          const uint32_t accumulate = (acc_started >> sacc) & 1u;
          {
            const uint32_t a_lo = (uint32_t)tc::make_desc_sw128(base + stage * stage_bytes);
            const uint32_t b_lo = a_lo + ((uint32_t)p.b_off >> 4);
            if (p.n_split == 1) {
              for (int v = 0; v < nw; ++v) {
                // vertical tap v reads the same box TW pixel rows (a multiple of the swizzle atom) further down;
                // the two weight sets of a dual segment read the same view
                const uint32_t av = a_lo + v * avs, bv = b_lo + v * b_step;
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                  tc::umma_f16_lo_if(lead, d_tmem, av + ks * 2, bv + ks * 2, desc_hi, idesc,
                                     (ks == 0 && v == 0) ? accumulate : 1u);
              }
            } else {
              for (int v = 0; v < nw; ++v)
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                  for (int pc = 0; pc < p.n_split; ++pc)
                    tc::umma_f16_lo_if(lead, d_tmem + pc * p.n_piece, a_lo + v * avs + ks * 2,
                                       b_lo + (pc * nw + v) * b_step + ks * 2, desc_hi, idesc,
                                       (ks == 0 && v == 0) ? accumulate : 1u);
            }
            tc::umma_commit_if(lead, bar_empty + 8 * stage);  // frees this smem stage once the MMAs above retire
          }
          acc_started |= 1u << sacc;
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
      tc::umma_commit_if(lead, bar_tfull + 8 * buf);  // accumulator of this tile complete
    }
    if (clk && leader) {
      p.dbg_clk[3] += CDC_CLK() - c_t0;
      p.dbg_clk[4] += c_wf;
      p.dbg_clk[5] += c_we;
      p.dbg_clk[6] += it;
    }
  } else {
    // =============================== epilogue (4 warps, one GEMM row per thread) ===============================
    const int quad = warp & 3;                 // TMEM lane quadrant this warp may access
    const int row = quad * 32 + lane;          // GEMM row == TMEM lane
    const int lx = row % p.TW, ly = (row / p.TW) % p.TH, lb = row / (p.TW * p.TH);
    const float inv_n = 1.f / (float)p.Ntot;
    uint8_t* stg = s_stage + (warp - 2) * kEpiStageBytes;
    long long* pixtab = reinterpret_cast<long long*>(s_stage + 4 * kEpiStageBytes) + (warp - 2) * 32;
    const long long out_rb = (long long)(EPI == EPI_RAW ? N : p.Ntot) * 2;   // bytes per output pixel row
    const bool has_res = !RT && p.res != nullptr && (EPI != EPI_LN_SHIFT) && (EPI != EPI_RAW);
    const bool has_lo = p.out_lo != nullptr;
    // per-image vectors (timestep shift of the tile's image | attention affine u, c): warp-private copy, fetched with
    // cp.async at the top of the tile so that its latency hides behind the accumulator wait
    float* wvec = reinterpret_cast<float*>(s_stage + 4 * (kEpiStageBytes + kEpiPixBytes)) + (warp - 2) * 2 * vs;
    constexpr bool shift_smem = false;   // measured: slower than reading the L1-resident shift row directly (+15 us at level 0)
    int cached_img = -1;                       // image whose vectors sit in wvec
    int it = 0;
    long long c_wt = 0, c_pre = 0, c_p12 = 0, c_p3 = 0;
    const long long c_t0 = CDC_CLK();
    for (int u = blockIdx.x; u < total_units; u += gridDim.x, ++it) {
      const long long i0 = CDC_CLK();
      const int t = fdiv(u, p.fd_upt);
      const int buf = p.nbuf == 2 ? (it & 1) : 0;
      const uint32_t use = (uint32_t)(p.nbuf == 2 ? (it >> 1) : it);
      const int ph = fdiv(t, p.fd_tpp);
      int r = t - ph * tiles_per_phase;
      const int tb = fdiv(r, p.fd_txy);
      r -= tb * p.tiles_x * p.tiles_y;
      const int ty = fdiv(r, p.fd_tx), tx = r - ty * p.tiles_x;
      const int xx = tx * p.TW + lx, yy = ty * p.TH + ly, bb = tb * p.TB + lb;
      const bool valid = xx < p.W && yy < p.H && lb < p.TB && bb < p.B;
      const int py = p.phases > 1 ? (ph >> 1) : 0, px = p.phases > 1 ? (ph & 1) : 0;
      const long long opix =
          valid ? ((long long)bb * p.out_H + yy * p.out_sy + py) * p.out_W + xx * p.out_sx + px : 0;
      __syncwarp();
      pixtab[lane] = valid ? opix : -1;
      __syncwarp();

      // residual prefetch (hi / lo halves of 32 channels of 32 rows, row-contiguous) — issued ahead of the TMEM waits
      // that would expose it; transposed to one-row-per-lane through the staging buffer right before use
      uint4 rh[4], rl[4];
      auto prefetch_res = [&](int c0) {
        if (!has_res) return;
        const __half* hi;
        const __half* lo;
        long long rb;
        const int ca = col0 + c0;   // absolute output column
        if (ca < p.res_C0) {
          rb = (long long)p.res_C0 * 2;
          hi = p.res + ca;
          lo = p.res_lo ? p.res_lo + ca : nullptr;
        } else {
          rb = (long long)(p.Ntot - p.res_C0) * 2;
          hi = p.res2 + (ca - p.res_C0);
          lo = p.res2_lo ? p.res2_lo + (ca - p.res_C0) : nullptr;
        }
        warp_load_rows64(pixtab, lane, rh, reinterpret_cast<const uint8_t*>(hi), rb);
        if (lo) {
          warp_load_rows64(pixtab, lane, rl, reinterpret_cast<const uint8_t*>(lo), rb);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) rl[j] = make_uint4(0u, 0u, 0u, 0u);
        }
      };
      auto own_res = [&](int c0) {   // uniform per warp
        if (!has_res) return;
        warp_rows64_to_own(stg, lane, rh);
        const bool lo = col0 + c0 < p.res_C0 ? p.res_lo != nullptr : p.res2_lo != nullptr;
        if (lo) warp_rows64_to_own(stg, lane, rl);
      };
      auto add_res = [&](int j, float (&o)[8]) {  // j = 8-channel group inside the prefetched 32
        float2 f;
        f = unpack_half2(rh[j].x); o[0] += f.x; o[1] += f.y;
        f = unpack_half2(rh[j].y); o[2] += f.x; o[3] += f.y;
        f = unpack_half2(rh[j].z); o[4] += f.x; o[5] += f.y;
        f = unpack_half2(rh[j].w); o[6] += f.x; o[7] += f.y;
        f = unpack_half2(rl[j].x); o[0] += f.x; o[1] += f.y;
        f = unpack_half2(rl[j].y); o[2] += f.x; o[3] += f.y;
        f = unpack_half2(rl[j].z); o[4] += f.x; o[5] += f.y;
        f = unpack_half2(rl[j].w); o[6] += f.x; o[7] += f.y;
      };
      prefetch_res(0);

      float2 st = make_float2(0.f, 1.f);
      if (EPI == EPI_AFFINE || shift_smem) {
        if (EPI == EPI_AFFINE && valid) st = p.stats_in[opix];
        const int img = min(tb * p.TB, p.B - 1);  // TB == 1: one image per tile
        if (img != cached_img) {                  // uniform across the warp
          if (EPI == EPI_AFFINE) {
            if (p.aff_parts <= 1) {
              for (int i = lane * 4; i < N; i += 128) {
                cp_async16(smem_u32(wvec + i), p.aff_u + (size_t)img * p.Ntot + col0 + i, 16);
                cp_async16(smem_u32(wvec + vs + i), p.aff_c + (size_t)img * p.Ntot + col0 + i, 16);
              }
            } else {   // per-column-tile partial sums (fused finish of the M_b product): add them in tile order
              for (int i = lane * 4; i < N; i += 128) {
                float4 su = make_float4(0.f, 0.f, 0.f, 0.f), sc = su;
                for (int t2 = 0; t2 < p.aff_parts; ++t2) {
                  const size_t o = ((size_t)img * p.aff_parts + t2) * p.Ntot + col0 + i;
                  const float4 a = *reinterpret_cast<const float4*>(p.aff_u + o);
                  const float4 c4 = *reinterpret_cast<const float4*>(p.aff_c + o);
                  su.x += a.x; su.y += a.y; su.z += a.z; su.w += a.w;
                  sc.x += c4.x; sc.y += c4.y; sc.z += c4.z; sc.w += c4.w;
                }
                *reinterpret_cast<float4*>(wvec + i) = su;
                *reinterpret_cast<float4*>(wvec + vs + i) = sc;
              }
            }
          } else {
            for (int i = lane * 4; i < N; i += 128)
              cp_async16(smem_u32(wvec + i), p.shift + (size_t)img * p.shift_stride + i, 16);
          }
          cp_async_commit();
          cached_img = img;
        }
      }

      const long long w0 = CDC_CLK();
      tc::mbar_wait(bar_tfull + 8 * buf, use & 1);
      const long long e0 = CDC_CLK();
      c_wt += e0 - w0;
      c_pre += w0 - i0;
      long long e1 = e0;
      tc::tc_fence_after();
      if (EPI == EPI_AFFINE || shift_smem) {
        cp_async_wait<0>();
        __syncwarp();
      }
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * Nacc);
      uint32_t v[32];
      if (p.dbg_skip_epi) {
        tc::tmem_ld32(taddr, v);
        tc::tc_fence_before();
        tc::mbar_arrive(bar_tempty + 8 * buf);
        if (v[0] == 0x7fc12345u && valid) p.out[opix] = __float2half(0.f);   // keeps the load alive
        continue;
      }
      uint8_t* const out_b = reinterpret_cast<uint8_t*>(p.out + col0);
      uint8_t* const out_lo_b = reinterpret_cast<uint8_t*>(p.out_lo + col0);

      if (EPI == EPI_RAW) {
        const int su = u - t * units_per_tile;
        const int slice = fdiv(su, p.fd_ks), ksp = su - slice * p.k_splits;
        uint8_t* dst = reinterpret_cast<uint8_t*>(p.raw + (size_t)ksp * p.raw_split_stride + slice * N);
        const long long raw_rb = (long long)p.Ntot * 4;
        for (int c0 = 0; c0 < N; c0 += 32) {
          tc::tmem_ld32(taddr + c0, v);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            uint4 w[4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
              w[j] = make_uint4(v[h * 16 + 4 * j], v[h * 16 + 4 * j + 1], v[h * 16 + 4 * j + 2], v[h * 16 + 4 * j + 3]);
            warp_store_rows64(stg, pixtab, lane, w, dst + (c0 + h * 16) * 4, raw_rb);
          }
        }
      } else if (EPI == EPI_BIAS || EPI == EPI_AFFINE) {
        using tc::f32x2;
        const ulonglong2* vb = reinterpret_cast<const ulonglong2*>(s_vec);            // conv bias
        const ulonglong2* vu = reinterpret_cast<const ulonglong2*>(wvec);        // affine u (per image)
        const ulonglong2* vc = reinterpret_cast<const ulonglong2*>(wvec + vs);   // affine c (per image)
        const f32x2 nmean2 = tc::pk(-st.x, -st.x), rstd2 = tc::pk(st.y, st.y);
        const bool ln_copy = EPI == EPI_BIAS && p.ln_out != nullptr;   // uniform
        for (int c0 = 0; c0 < N && !(ln_copy && p.skip_out); c0 += 32) {
          tc::tmem_ld32(taddr + c0, v);
          uint32_t rv[RT ? 32 : 1];
          if (RT) tc::tmem_ld32(taddr + N + c0, *reinterpret_cast<uint32_t(*)[32]>(&rv[0]));   // residual accumulator
          if (c0 + 32 >= N && !ln_copy) {   // last TMEM read of this tile: hand the accumulator buffer back to the MMA warp
            tc::tc_fence_before();
            tc::mbar_arrive(bar_tempty + 8 * buf);
          }
          own_res(c0);
          uint4 wh[4], wl[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            f32x2 y[4];
            const int q = (c0 >> 2) + 2 * j;
            if (EPI == EPI_AFFINE) {
              const ulonglong2 u0 = vu[q], u1 = vu[q + 1], a0 = vc[q], a1 = vc[q + 1];
              y[0] = tc::fma2(rstd2, tc::fma2(nmean2, u0.x, tc::pku(v[8 * j], v[8 * j + 1])), a0.x);
              y[1] = tc::fma2(rstd2, tc::fma2(nmean2, u0.y, tc::pku(v[8 * j + 2], v[8 * j + 3])), a0.y);
              y[2] = tc::fma2(rstd2, tc::fma2(nmean2, u1.x, tc::pku(v[8 * j + 4], v[8 * j + 5])), a1.x);
              y[3] = tc::fma2(rstd2, tc::fma2(nmean2, u1.y, tc::pku(v[8 * j + 6], v[8 * j + 7])), a1.y);
            } else {
              const ulonglong2 b0 = vb[q], b1 = vb[q + 1];
              y[0] = tc::add2(tc::pku(v[8 * j], v[8 * j + 1]), b0.x);
              y[1] = tc::add2(tc::pku(v[8 * j + 2], v[8 * j + 3]), b0.y);
              y[2] = tc::add2(tc::pku(v[8 * j + 4], v[8 * j + 5]), b1.x);
              y[3] = tc::add2(tc::pku(v[8 * j + 6], v[8 * j + 7]), b1.y);
            }
            float o[8];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float2 f = tc::upk(y[k]);
              o[2 * k] = f.x;
              o[2 * k + 1] = f.y;
            }
            if (has_res) add_res(j, o);
            if (RT) {
#pragma unroll
              for (int k = 0; k < 8; ++k) o[k] += __uint_as_float(rv[RT ? j * 8 + k : 0]);
            }
            pack_out8(o, wh[j], wl[j], has_lo);
          }
          if (c0 + 32 < N) prefetch_res(c0 + 32);
          warp_store_rows64(stg, pixtab, lane, wh, out_b + c0 * 2, out_rb);
          if (has_lo) warp_store_rows64(stg, pixtab, lane, wl, out_lo_b + c0 * 2, out_rb);
        }
        if (ln_copy) {
          // LayerNorm of the 64-column row (conv + bias, fp32) -> fp16 copy for the final convolution
          const ulonglong2* vg = reinterpret_cast<const ulonglong2*>(s_vec + vs);
          const ulonglong2* vo = reinterpret_cast<const ulonglong2*>(s_vec + 2 * vs);
          uint32_t v1[32];
          tc::tmem_ld32_issue(taddr, v);
          tc::tmem_ld32_issue(taddr + 32, v1);
          tc::tmem_wait_ld();
          tc::tc_fence_before();
          tc::mbar_arrive(bar_tempty + 8 * buf);
          f32x2 x[32];
          f32x2 s0 = 0ull, s1 = 0ull, s2 = 0ull, s3 = 0ull;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const ulonglong2 ba = vb[i], bc = vb[8 + i];
            x[2 * i] = tc::add2(tc::pku(v[4 * i], v[4 * i + 1]), ba.x);
            x[2 * i + 1] = tc::add2(tc::pku(v[4 * i + 2], v[4 * i + 3]), ba.y);
            x[16 + 2 * i] = tc::add2(tc::pku(v1[4 * i], v1[4 * i + 1]), bc.x);
            x[16 + 2 * i + 1] = tc::add2(tc::pku(v1[4 * i + 2], v1[4 * i + 3]), bc.y);
            s0 = tc::add2(s0, x[2 * i]);
            s1 = tc::add2(s1, x[2 * i + 1]);
            s2 = tc::add2(s2, x[16 + 2 * i]);
            s3 = tc::add2(s3, x[16 + 2 * i + 1]);
          }
          const float2 fs = tc::upk(tc::add2(tc::add2(s0, s1), tc::add2(s2, s3)));
          const float mean = (fs.x + fs.y) * (1.f / 64.f);
          const f32x2 nm = tc::pk(-mean, -mean);
          f32x2 q0 = 0ull, q1 = 0ull;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            x[2 * i] = tc::add2(x[2 * i], nm);
            x[2 * i + 1] = tc::add2(x[2 * i + 1], nm);
            q0 = tc::fma2(x[2 * i], x[2 * i], q0);
            q1 = tc::fma2(x[2 * i + 1], x[2 * i + 1], q1);
          }
          const float2 fq = tc::upk(tc::add2(q0, q1));
          const float rstd = 1.f / sqrtf((fq.x + fq.y) * (1.f / 64.f) + 1e-5f);
          const f32x2 rstd2 = tc::pk(rstd, rstd);
          uint8_t* const ln_b8 = reinterpret_cast<uint8_t*>(p.ln_out);
#pragma unroll
          for (int g2 = 0; g2 < 2; ++g2) {
            uint4 wn[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int c = g2 * 32 + j * 8;
              const ulonglong2 g0 = vg[c >> 2], g1 = vg[(c >> 2) + 1], b0 = vo[c >> 2], b1 = vo[(c >> 2) + 1];
              const float2 y0 = tc::upk(tc::fma2(x[g2 * 16 + j * 4], tc::mul2(rstd2, g0.x), b0.x));
              const float2 y1 = tc::upk(tc::fma2(x[g2 * 16 + j * 4 + 1], tc::mul2(rstd2, g0.y), b0.y));
              const float2 y2 = tc::upk(tc::fma2(x[g2 * 16 + j * 4 + 2], tc::mul2(rstd2, g1.x), b1.x));
              const float2 y3 = tc::upk(tc::fma2(x[g2 * 16 + j * 4 + 3], tc::mul2(rstd2, g1.y), b1.y));
              wn[j] = make_uint4(pack_half2(y0.x, y0.y), pack_half2(y1.x, y1.y), pack_half2(y2.x, y2.y),
                                 pack_half2(y3.x, y3.y));
            }
            warp_store_rows64(stg, pixtab, lane, wn, ln_b8 + g2 * 64, out_rb);
          }
        }
      } else {
        // ---- channel LayerNorm: exact two-pass statistics from the fp32 accumulator (packed fp32x2 arithmetic) ----
        using tc::f32x2;
        const ulonglong2* vb = reinterpret_cast<const ulonglong2*>(s_vec);            // conv bias
        const ulonglong2* vg = reinterpret_cast<const ulonglong2*>(s_vec + vs);       // LayerNorm gain
        const ulonglong2* vo = reinterpret_cast<const ulonglong2*>(s_vec + 2 * vs);   // LayerNorm offset
        // 4 consecutive entries (two packed pairs) of a per-channel vector: constant bank (N64 == 2) or shared memory
        auto vec4 = [&](const ulonglong2* sm, const float* cst, int q4) -> ulonglong2 {
          if constexpr (N64 == 2) {
            (void)sm;
            return make_ulonglong2(*reinterpret_cast<const unsigned long long*>(cst + 4 * q4),
                                   *reinterpret_cast<const unsigned long long*>(cst + 4 * q4 + 2));
          } else {
            (void)cst;
            return sm[q4];
          }
        };
        const ulonglong2* vrb = reinterpret_cast<const ulonglong2*>(s_vec + 3 * vs);  // bias of a fused res_conv
        const float* c_bias = nullptr;
        const float* c_g = nullptr;
        const float* c_b = nullptr;
        const float* c_rb = nullptr;
        if constexpr (N64 == 2) {
          c_bias = kc.bias;
          c_g = kc.g;
          c_b = kc.b;
          c_rb = kc.rbias;
        }
        constexpr bool res_tmem = RT && N64 != 0 && EPI == EPI_LN_RES;   // residual = second TMEM accumulator
        const ulonglong2* shift2 =
            shift_smem ? reinterpret_cast<const ulonglong2*>(wvec)
                       : (EPI == EPI_LN_SHIFT && p.shift)
                             ? reinterpret_cast<const ulonglong2*>(p.shift + (size_t)(valid ? bb : 0) * p.shift_stride + col0)
                             : nullptr;
        float osum = 0.f, osq = 0.f;
        f32x2 os2 = 0ull, oq2 = 0ull;
        float mean, rstd;
        // finish 8 columns [c, c+8) of the row from their centred values d[0..3] (pairs) -> fp16 hi / lo pieces
        auto finish8 = [&](const f32x2 (&d)[4], int c, int j, uint4& hi, uint4& lo, f32x2 rstd2) {
          const ulonglong2 g0 = vec4(vg, c_g, c >> 2), g1 = vec4(vg, c_g, (c >> 2) + 1);
          const ulonglong2 b0 = vec4(vo, c_b, c >> 2), b1 = vec4(vo, c_b, (c >> 2) + 1);
          f32x2 y[4];
          y[0] = tc::fma2(d[0], tc::mul2(rstd2, g0.x), b0.x);
          y[1] = tc::fma2(d[1], tc::mul2(rstd2, g0.y), b0.y);
          y[2] = tc::fma2(d[2], tc::mul2(rstd2, g1.x), b1.x);
          y[3] = tc::fma2(d[3], tc::mul2(rstd2, g1.y), b1.y);
          float o[8];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float2 f = tc::upk(y[k]);
            o[2 * k] = fmaxf(f.x, 0.f);
            o[2 * k + 1] = fmaxf(f.y, 0.f);
          }
          if (shift2) {
            const ulonglong2 s0 = shift2[c >> 2], s1 = shift2[(c >> 2) + 1];
            float2 f;
            f = tc::upk(tc::add2(tc::pk(o[0], o[1]), s0.x)); o[0] = f.x; o[1] = f.y;
            f = tc::upk(tc::add2(tc::pk(o[2], o[3]), s0.y)); o[2] = f.x; o[3] = f.y;
            f = tc::upk(tc::add2(tc::pk(o[4], o[5]), s1.x)); o[4] = f.x; o[5] = f.y;
            f = tc::upk(tc::add2(tc::pk(o[6], o[7]), s1.y)); o[6] = f.x; o[7] = f.y;
          }
          if (has_res) add_res(j, o);
          if (res_tmem) {   // v[] holds the 32 residual-accumulator columns of the current group
            const ulonglong2 r0 = vec4(vrb, c_rb, c >> 2), r1 = vec4(vrb, c_rb, (c >> 2) + 1);
            const float2 f0 = tc::upk(tc::add2(tc::pku(v[8 * j], v[8 * j + 1]), r0.x));
            const float2 f1 = tc::upk(tc::add2(tc::pku(v[8 * j + 2], v[8 * j + 3]), r0.y));
            const float2 f2 = tc::upk(tc::add2(tc::pku(v[8 * j + 4], v[8 * j + 5]), r1.x));
            const float2 f3 = tc::upk(tc::add2(tc::pku(v[8 * j + 6], v[8 * j + 7]), r1.y));
            o[0] += f0.x; o[1] += f0.y; o[2] += f1.x; o[3] += f1.y;
            o[4] += f2.x; o[5] += f2.y; o[6] += f3.x; o[7] += f3.y;
          }
          pack_out8(o, hi, lo, has_lo);
          if (p.stats_out) {  // statistics of the rounded values the consumer will read (packed accumulators)
            float2 f;
            f32x2 h2;
            f = unpack_half2(hi.x); h2 = tc::pk(f.x, f.y); os2 = tc::add2(os2, h2); oq2 = tc::fma2(h2, h2, oq2);
            f = unpack_half2(hi.y); h2 = tc::pk(f.x, f.y); os2 = tc::add2(os2, h2); oq2 = tc::fma2(h2, h2, oq2);
            f = unpack_half2(hi.z); h2 = tc::pk(f.x, f.y); os2 = tc::add2(os2, h2); oq2 = tc::fma2(h2, h2, oq2);
            f = unpack_half2(hi.w); h2 = tc::pk(f.x, f.y); os2 = tc::add2(os2, h2); oq2 = tc::fma2(h2, h2, oq2);
          }
        };
        if (N64) {
          // the whole row lives in registers: one TMEM pass, and the accumulator buffer is released before the arithmetic
          uint32_t v1[32];
          tc::tmem_ld32_issue(taddr, v);
          tc::tmem_ld32_issue(taddr + 32, v1);
          tc::tmem_wait_ld();
          if (!res_tmem) {   // (with a fused res_conv the buffer is released after the residual columns are read)
            tc::tc_fence_before();
            tc::mbar_arrive(bar_tempty + 8 * buf);
          }
          f32x2 x[32];
          f32x2 s0 = 0ull, s1 = 0ull, s2 = 0ull, s3 = 0ull;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const ulonglong2 ba = vec4(vb, c_bias, i), bc = vec4(vb, c_bias, 8 + i);
            x[2 * i] = tc::add2(tc::pku(v[4 * i], v[4 * i + 1]), ba.x);
            x[2 * i + 1] = tc::add2(tc::pku(v[4 * i + 2], v[4 * i + 3]), ba.y);
            x[16 + 2 * i] = tc::add2(tc::pku(v1[4 * i], v1[4 * i + 1]), bc.x);
            x[16 + 2 * i + 1] = tc::add2(tc::pku(v1[4 * i + 2], v1[4 * i + 3]), bc.y);
            s0 = tc::add2(s0, x[2 * i]);
            s1 = tc::add2(s1, x[2 * i + 1]);
            s2 = tc::add2(s2, x[16 + 2 * i]);
            s3 = tc::add2(s3, x[16 + 2 * i + 1]);
          }
          float lsum;
          {
            const float2 f = tc::upk(tc::add2(tc::add2(s0, s1), tc::add2(s2, s3)));
            lsum = f.x + f.y;
            mean = lsum * (1.f / 64.f);   // mean of this CTA's 64 columns
          }
          const f32x2 nm = tc::pk(-mean, -mean);
          f32x2 q0 = 0ull, q1 = 0ull, q2 = 0ull, q3 = 0ull;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            x[4 * i] = tc::add2(x[4 * i], nm);
            x[4 * i + 1] = tc::add2(x[4 * i + 1], nm);
            x[4 * i + 2] = tc::add2(x[4 * i + 2], nm);
            x[4 * i + 3] = tc::add2(x[4 * i + 3], nm);
            q0 = tc::fma2(x[4 * i], x[4 * i], q0);
            q1 = tc::fma2(x[4 * i + 1], x[4 * i + 1], q1);
            q2 = tc::fma2(x[4 * i + 2], x[4 * i + 2], q2);
            q3 = tc::fma2(x[4 * i + 3], x[4 * i + 3], q3);
          }
          float m2;   // sum of squared deviations from the (slice) mean
          {
            const float2 f = tc::upk(tc::add2(tc::add2(q0, q1), tc::add2(q2, q3)));
            m2 = f.x + f.y;
          }
          if (p.cluster_n > 1) {
            // Row statistics over all column slices: every CTA of the cluster sends (sum, M2) of its 64 columns to every
            // peer; the slices are merged in rank order (Chan's parallel variance: exact, deterministic).
            const int par = it & 1;
            const uint32_t xb = bar_x + 8 * par;
            const uint32_t my_rank = blockIdx.x % p.cluster_n;
            float2* xs = s_xchg + (size_t)par * p.cluster_n * 128;
            if (row == 0) tc::mbar_expect_tx(xb, (uint32_t)p.cluster_n * 128u * 8u);
            const uint32_t slot = smem_u32(xs + my_rank * 128 + row);
            for (int r = 0; r < p.cluster_n; ++r) tc::st_async_f2(tc::mapa(slot, r), lsum, m2, tc::mapa(xb, r));
            tc::mbar_wait(xb, (uint32_t)(it >> 1) & 1u);
            float tot = 0.f;
            for (int r = 0; r < p.cluster_n; ++r) tot += xs[r * 128 + row].x;
            const float gmean = tot * inv_n;
            float gm2 = 0.f;
            for (int r = 0; r < p.cluster_n; ++r) {
              const float2 f = xs[r * 128 + row];
              const float dm = f.x * (1.f / 64.f) - gmean;
              gm2 += f.y + 64.f * dm * dm;
            }
            const float dl = mean - gmean;   // re-centre the registers on the row mean
            const f32x2 dl2 = tc::pk(dl, dl);
#pragma unroll
            for (int i = 0; i < 32; ++i) x[i] = tc::add2(x[i], dl2);
            mean = gmean;
            m2 = gm2;
          }
          rstd = 1.f / sqrtf(m2 * inv_n + 1e-5f);
          e1 = CDC_CLK();
          const f32x2 rstd2 = tc::pk(rstd, rstd);
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            const int c0 = g * 32;
            if (res_tmem) {
              tc::tmem_ld32(taddr + 64 + c0, v);
              if (g == 1) {
                tc::tc_fence_before();
                tc::mbar_arrive(bar_tempty + 8 * buf);
              }
            } else {
              own_res(c0);
            }
            uint4 wh[4], wl[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const f32x2 d[4] = {x[g * 16 + j * 4], x[g * 16 + j * 4 + 1], x[g * 16 + j * 4 + 2], x[g * 16 + j * 4 + 3]};
              finish8(d, c0 + j * 8, j, wh[j], wl[j], rstd2);
            }
            if (g == 0) prefetch_res(32);
            warp_store_rows64(stg, pixtab, lane, wh, out_b + c0 * 2, out_rb);
            if (has_lo) warp_store_rows64(stg, pixtab, lane, wl, out_lo_b + c0 * 2, out_rb);
          }
        } else {
          f32x2 s0 = 0ull, s1 = 0ull;
          for (int c0 = 0; c0 < N; c0 += 32) {
            tc::tmem_ld32(taddr + c0, v);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const ulonglong2 ba = vb[(c0 >> 2) + i];
              s0 = tc::add2(s0, tc::add2(tc::pku(v[4 * i], v[4 * i + 1]), ba.x));
              s1 = tc::add2(s1, tc::add2(tc::pku(v[4 * i + 2], v[4 * i + 3]), ba.y));
            }
          }
          {
            const float2 f = tc::upk(tc::add2(s0, s1));
            mean = (f.x + f.y) * inv_n;
          }
          const f32x2 nm = tc::pk(-mean, -mean);
          f32x2 q0 = 0ull, q1 = 0ull;
          for (int c0 = 0; c0 < N; c0 += 32) {
            tc::tmem_ld32(taddr + c0, v);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const ulonglong2 ba = vb[(c0 >> 2) + i];
              const f32x2 d0 = tc::add2(tc::add2(tc::pku(v[4 * i], v[4 * i + 1]), ba.x), nm);
              const f32x2 d1 = tc::add2(tc::add2(tc::pku(v[4 * i + 2], v[4 * i + 3]), ba.y), nm);
              q0 = tc::fma2(d0, d0, q0);
              q1 = tc::fma2(d1, d1, q1);
            }
          }
          {
            const float2 f = tc::upk(tc::add2(q0, q1));
            rstd = 1.f / sqrtf((f.x + f.y) * inv_n + 1e-5f);
          }
          e1 = CDC_CLK();
          const f32x2 rstd2 = tc::pk(rstd, rstd);
          for (int c0 = 0; c0 < N; c0 += 32) {
            tc::tmem_ld32(taddr + c0, v);
            if (c0 + 32 >= N) {   // last TMEM read of this tile: hand the accumulator buffer back to the MMA warp
              tc::tc_fence_before();
              tc::mbar_arrive(bar_tempty + 8 * buf);
            }
            own_res(c0);
            uint4 wh[4], wl[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const ulonglong2 ba = vb[(c0 >> 2) + 2 * j], bc = vb[(c0 >> 2) + 2 * j + 1];
              f32x2 d[4];
              d[0] = tc::add2(tc::add2(tc::pku(v[8 * j], v[8 * j + 1]), ba.x), nm);
              d[1] = tc::add2(tc::add2(tc::pku(v[8 * j + 2], v[8 * j + 3]), ba.y), nm);
              d[2] = tc::add2(tc::add2(tc::pku(v[8 * j + 4], v[8 * j + 5]), bc.x), nm);
              d[3] = tc::add2(tc::add2(tc::pku(v[8 * j + 6], v[8 * j + 7]), bc.y), nm);
              finish8(d, c0 + j * 8, j, wh[j], wl[j], rstd2);
            }
            if (c0 + 32 < N) prefetch_res(c0 + 32);
            warp_store_rows64(stg, pixtab, lane, wh, out_b + c0 * 2, out_rb);
            if (has_lo) warp_store_rows64(stg, pixtab, lane, wl, out_lo_b + c0 * 2, out_rb);
          }
        }
        if (p.stats_out) {
          const float2 fs = tc::upk(os2), fq = tc::upk(oq2);
          osum = fs.x + fs.y;
          osq = fq.x + fq.y;
        }
        if (N64 && p.cluster_n > 1 && p.xchg_stats) {   // row statistics of the stored output over all column slices
          const int par = it & 1;
          const uint32_t xb = bar_x + 8 * (2 + par);
          const uint32_t my_rank = blockIdx.x % p.cluster_n;
          float2* xs = s_xchg + (size_t)(2 + par) * p.cluster_n * 128;
          if (row == 0) tc::mbar_expect_tx(xb, (uint32_t)p.cluster_n * 128u * 8u);
          const uint32_t slot = smem_u32(xs + my_rank * 128 + row);
          for (int r = 0; r < p.cluster_n; ++r) tc::st_async_f2(tc::mapa(slot, r), osum, osq, tc::mapa(xb, r));
          tc::mbar_wait(xb, (uint32_t)(it >> 1) & 1u);
          osum = 0.f;
          osq = 0.f;
          for (int r = 0; r < p.cluster_n; ++r) {
            const float2 f = xs[r * 128 + row];
            osum += f.x;
            osq += f.y;
          }
        }
        if (p.stats_out && valid && col0 == 0) {
          const float m = osum * inv_n;
          const float var = fmaxf(osq * inv_n - m * m, 0.f);
          p.stats_out[opix] = make_float2(m, 1.f / sqrtf(var + 1e-5f));
        }
      }
      if (EPI == EPI_RAW) {
        tc::tc_fence_before();
        tc::mbar_arrive(bar_tempty + 8 * buf);
        if (p.tile_ctr != nullptr) {
          // ---- fused finish: signal this unit's partial tile, wait for the tile's other units, finish a share of the rows ----
          unsigned int* ctr = p.tile_ctr + 2 * t;
          __threadfence();                                           // this thread's partial stores: visible device-wide
          asm volatile("bar.sync 1, 128;" ::: "memory");             // ... for all 128 epilogue threads
          if (threadIdx.x == 64) {
            asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
            unsigned int seen;
            const long long t0 = clock64();
            do {
              asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(ctr) : "memory");
              if (seen < (unsigned)units_per_tile && clock64() - t0 > 6000000000LL) {
                printf("cdc igemm_tc: tile %d arrival counter stuck at %u of %d (block %d)\n", t, seen, units_per_tile, blockIdx.x);
                __trap();
              }
            } while (seen < (unsigned)units_per_tile);
          }
          asm volatile("bar.sync 1, 128;" ::: "memory");
          __threadfence();
          LnRowsParams q;
          q.raw = p.raw; q.k_splits = p.k_splits; q.split_stride = p.raw_split_stride; q.N = p.Ntot; q.Ntot = p.Ntot;
          q.rows = 0; q.pix_per_image = 0; q.epi = p.fin_epi;
          q.bias = p.bias; q.ln_g = p.ln_g; q.ln_b = p.ln_b; q.shift = p.shift; q.shift_stride = p.shift_stride;
          q.res = p.res; q.res_C0 = p.res_C0; q.res2 = p.res2; q.res_lo = p.res_lo; q.res2_lo = p.res2_lo;
          q.out = p.out; q.out_lo = p.out_lo; q.stats_out = p.stats_out;
          const int su = u - t * units_per_tile;
          const int tile_rows = p.TW * p.TH * p.TB;
          for (int r0 = su * 4 + (warp - 2); r0 < tile_rows; r0 += units_per_tile * 4) {   // warp-uniform
            const int rx = r0 % p.TW, ry = (r0 / p.TW) % p.TH, rb = r0 / (p.TW * p.TH);
            const int x2 = tx * p.TW + rx, y2 = ty * p.TH + ry, b2 = tb * p.TB + rb;
            if (x2 < p.W && y2 < p.H && b2 < p.B) {
              const long long pix2 = ((long long)b2 * p.out_H + y2 * p.out_sy + py) * p.out_W + x2 * p.out_sx + px;
              ln_rows_row<false>(q, pix2, b2, lane);
            }
          }
          asm volatile("bar.sync 1, 128;" ::: "memory");
          if (threadIdx.x == 64) {
            const unsigned int old = atomicAdd(ctr + 1, 1u);
            if (old == (unsigned)units_per_tile - 1u) {   // last unit to leave: nobody waits on this tile any more
              ctr[0] = 0u;
              ctr[1] = 0u;
              __threadfence();
            }
          }
        }
      }
      c_p12 += e1 - e0;
      c_p3 += CDC_CLK() - e1;
    }
    if (clk && threadIdx.x == 64) {
      p.dbg_clk[10] += c_p12;
      p.dbg_clk[11] += c_p3;
      p.dbg_clk[7] += CDC_CLK() - c_t0;
      p.dbg_clk[8] += c_wt;
      p.dbg_clk[9] += c_pre;
    }
  }

  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc::tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols));
  }
}

// Second half of the sliced mode as its own launch (CDC_FUSE_LNROWS=0; the default finishes inside igemm_tc_kernel).
template <bool HOIST>
__global__ void __launch_bounds__(256) ln_rows_kernel(const LnRowsParams p) {
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const long long pix = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (pix >= p.rows) return;
  ln_rows_row<HOIST>(p, pix, (int)(pix / p.pix_per_image), lane);
}

}  // namespace cdc
