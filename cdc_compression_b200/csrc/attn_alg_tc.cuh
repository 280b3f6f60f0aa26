// Per-image C x C algebra of the linear attention on the Blackwell tensor cores (reference: the two einsums and to_out of
// LinearAttention.forward, epsilonparam/modules/network_components.py:135-139, re-associated as in SURVEY.md Appendix E):
//
//   T   = ctx^T (C^-1/2 W_q)        T[e][c]   = sum_d ctxn[d][e] * wq[d][c]
//   M_b = W_out T                   M[o][c]   = sum_e woT[e][o]  * T[e][c]        -> Mg16 = fp16(M * g), row sums (finish)
//
// Both are  D[m][n] = sum_k A[k][m] * B[k][n]  with operands stored K-slowest, i.e. MN-major UMMA operands.  They need
// fp32-grade accuracy (a single fp16 pass costs 1e-4 on the U-Net output), so every operand lives in global memory as an
// fp16 value and its fp16 rounding remainder (x = hi + lo) and the product is accumulated in TMEM as
// lo*hi + hi*lo + hi*hi — the same three passes as gemm3xf16_tn_kernel (attn.cuh), whose mma.sync rate bounded it at
// 15-19 us for C >= 320 (0.27 ms of the step in round 1).
//
// Layout of every operand ("MN-blocked"): element (k, mn) of a [K][MN] matrix at ((mn/64) * K + k) * 64 + mn % 64, so one
// 4-D TMA box {64, 64 k-rows, 2 blocks, 1 image} lands as [block][k][64] = the canonical SWIZZLE_128B MN-major tile
// (128-byte rows of 64 MN elements, 8-row atoms 1024 bytes apart, 64-element MN blocks 8192 bytes apart).
// One CTA = one 128 x 128 tile of D for one image; warp roles as in igemm_tc.cuh (0 TMA, 1 MMA + TMEM, 2-5 epilogue).
#pragma once
#include "igemm_tc.cuh"

namespace cdc {

struct AlgTcParams {
  int C;                 // M = N = K = C (multiple of 64, <= 384)
  int kchunks;           // C / 64
  int n_tiles;           // ceil(C / 128): blockIdx.x = m_tile * n_tiles + n_tile
  int mode;              // 0: store D as hi / lo fp16 in MN-blocked layout with K = D's row (operand B of the next product)
                         // 1: finish — Mg16 = fp16(D * g) in the conv weight layout [C/64][C][64] + per-64-column row sums
  int stages;
  int a_img, b_img;      // 1: the operand is per image (tensor-map dimension 3), 0: shared by all images
  __half* out_hi;        // mode 0: [B][C/64][C][64]
  __half* out_lo;
  const float* g;        // mode 1: LayerNorm gain / offset of the PreNorm, to_out bias
  const float* bln;
  const float* bout;
  __half* Mg16;          // [B][C/64][C][64]
  float* um_part;        // [B][C/64][C]
  float* cm_part;
};

constexpr int kAlgThreads = 192;
constexpr int kAlgStageBytes = 4 * 16384;   // A_hi | A_lo | B_hi | B_lo, each [2 blocks][64 k][64]
__host__ __device__ inline int alg_tc_smem_bytes(int stages) { return 1024 + stages * kAlgStageBytes + 256; }

namespace tc {
// MN-major SWIZZLE_128B operand descriptor: 64-element MN blocks `lbo_bytes` apart, 8-row (K) atoms 1024 bytes apart.
__device__ __forceinline__ uint64_t make_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// kind::f16 instruction descriptor, fp16 A / B both MN-major, fp32 accumulate, M = 128, N = n.
__device__ __forceinline__ uint32_t make_idesc_f16_mn(int n) {
  return (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
}
}  // namespace tc

__global__ void __launch_bounds__(kAlgThreads, 1)
attn_alg_tc_kernel(const __grid_constant__ TcMaps maps, const AlgTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t bars = base + p.stages * kAlgStageBytes;
  const uint32_t bar_full = bars;            // [8]
  const uint32_t bar_empty = bars + 64;      // [8]
  const uint32_t bar_tfull = bars + 128;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + (bars + 136 - base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int img = blockIdx.z;
  const int mt = (int)blockIdx.x / p.n_tiles, nt = (int)blockIdx.x % p.n_tiles;
  constexpr int kTmemCols = 128;

  pdl_launch_dependents();
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      tc::prefetch_tmap(&maps.a[i]);
      tc::prefetch_tmap(&maps.b[i]);
    }
    for (int s = 0; s < p.stages; ++s) {
      tc::mbar_init(bar_full + 8 * s, 1);
      tc::mbar_init(bar_empty + 8 * s, 1);
    }
    tc::mbar_init(bar_tfull, 1);
    tc::fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  pdl_wait();

  if (warp == 0) {
    // =============================== TMA producer ===============================
    const bool leader = tc::elect_one();
    int stage = 0;
    uint32_t phase = 0;
    for (int kc = 0; kc < p.kchunks; ++kc) {
      tc::mbar_wait(bar_empty + 8 * stage, phase ^ 1);
      if (leader) {
        const uint32_t s0 = base + stage * kAlgStageBytes;
        const uint32_t full = bar_full + 8 * stage;
        tc::mbar_expect_tx(full, (uint32_t)kAlgStageBytes);
        tc::tma_load_4d(s0, &maps.a[0], full, 0, kc * 64, mt * 2, p.a_img ? img : 0);
        tc::tma_load_4d(s0 + 16384, &maps.a[1], full, 0, kc * 64, mt * 2, p.a_img ? img : 0);
        tc::tma_load_4d(s0 + 32768, &maps.b[0], full, 0, kc * 64, nt * 2, p.b_img ? img : 0);
        tc::tma_load_4d(s0 + 49152, &maps.b[1], full, 0, kc * 64, nt * 2, p.b_img ? img : 0);
      }
      __syncwarp();
      if (++stage == p.stages) {
        stage = 0;
        phase ^= 1;
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    const bool leader = tc::elect_one();
    const uint32_t idesc = tc::make_idesc_f16_mn(128);
    const uint32_t desc_hi = (uint32_t)(tc::make_desc_mn_sw128(0, 8192) >> 32);
    int stage = 0;
    uint32_t phase = 0;
    for (int kc = 0; kc < p.kchunks; ++kc) {
      tc::mbar_wait(bar_full + 8 * stage, phase);
      tc::tc_fence_after();
      if (leader) {
        const uint32_t s0 = base + stage * kAlgStageBytes;
        const uint32_t ah = (uint32_t)tc::make_desc_mn_sw128(s0, 8192), al = ah + (16384u >> 4);
        const uint32_t bh = ah + (32768u >> 4), bl = ah + (49152u >> 4);
        // small terms first: lo*hi, hi*lo, then hi*hi; one UMMA covers 16 k rows = two 1024-byte atoms
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint32_t o = (uint32_t)ks * (2048u >> 4);
          tc::umma_f16_lo(tmem_base, al + o, bh + o, desc_hi, idesc, (kc > 0 || ks > 0) ? 1u : 0u);
          tc::umma_f16_lo(tmem_base, ah + o, bl + o, desc_hi, idesc, 1u);
          tc::umma_f16_lo(tmem_base, ah + o, bh + o, desc_hi, idesc, 1u);
        }
        tc::umma_commit(bar_empty + 8 * stage);
      }
      __syncwarp();
      if (++stage == p.stages) {
        stage = 0;
        phase ^= 1;
      }
    }
    if (leader) tc::umma_commit(bar_tfull);
    __syncwarp();
  } else {
    // =============================== epilogue: one row of D per thread ===============================
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const int gm = mt * 128 + row;                 // global row (m)
    const bool row_ok = gm < p.C;
    const int blocks = p.C >> 6;
    tc::mbar_wait(bar_tfull, 0);
    tc::tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16);
    uint32_t v[32];
    if (p.mode == 0) {
      for (int c0 = 0; c0 < 128; c0 += 32) {
        const int n0 = nt * 128 + c0;              // uniform
        if (n0 >= p.C) break;
        tc::tmem_ld32(taddr + c0, v);
        if (row_ok) {
          const size_t o = (((size_t)img * blocks + (n0 >> 6)) * p.C + gm) * 64 + (n0 & 63);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float f[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) f[k] = __uint_as_float(v[8 * q + k]);
            uint4 hi, lo;
            pack_out8(f, hi, lo, true);
            *reinterpret_cast<uint4*>(p.out_hi + o + 8 * q) = hi;
            *reinterpret_cast<uint4*>(p.out_lo + o + 8 * q) = lo;
          }
        }
      }
    } else {
      for (int h = 0; h < 2; ++h) {                // 64-column tiles of the conv weight layout
        const int n0 = nt * 128 + h * 64;          // uniform
        if (n0 >= p.C) break;
        const int ct = n0 >> 6;
        float su = 0.f, sc = 0.f;
        __half* mg = p.Mg16 + (((size_t)img * blocks + ct) * p.C + (row_ok ? gm : 0)) * 64;
#pragma unroll
        for (int g2 = 0; g2 < 2; ++g2) {
          tc::tmem_ld32(taddr + h * 64 + g2 * 32, v);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int n = n0 + g2 * 32 + 8 * q;
            const float4 ga = *reinterpret_cast<const float4*>(p.g + n), gb = *reinterpret_cast<const float4*>(p.g + n + 4);
            const float4 ba = *reinterpret_cast<const float4*>(p.bln + n), bb = *reinterpret_cast<const float4*>(p.bln + n + 4);
            const float gg[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
            const float bl[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
            uint32_t hw[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float v0 = __uint_as_float(v[8 * q + 2 * k]), v1 = __uint_as_float(v[8 * q + 2 * k + 1]);
              hw[k] = pack_half2(v0 * gg[2 * k], v1 * gg[2 * k + 1]);
              const float2 f = unpack_half2(hw[k]);
              su += f.x + f.y;
              sc += v0 * bl[2 * k] + v1 * bl[2 * k + 1];
            }
            if (row_ok) *reinterpret_cast<uint4*>(mg + g2 * 32 + 8 * q) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
          }
        }
        if (row_ok) {
          const size_t o = ((size_t)img * blocks + ct) * p.C + gm;
          p.um_part[o] = su;
          p.cm_part[o] = sc + (ct == 0 ? p.bout[gm] : 0.f);
        }
      }
    }
  }

  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc::tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols));
  }
}

}  // namespace cdc
