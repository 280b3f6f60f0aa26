// Linear-attention context on the Blackwell tensor cores (reference: LinearAttention.forward,
// epsilonparam/modules/network_components.py:127-137, with PreNorm's LayerNorm folded as in attn.cuh):
//
//   K,V = rstd*(Wg x - mean*u) + c ;  P = exp(K - m_d) ;  ctx[d,e] = sum_n P[n,d] V[n,e] ;  S_d = sum_n P[n,d]
//
// Same split partials as attn_ctx_kernel (part_ctx / part_m / part_s, merged by attn_combine_kernel), but both GEMMs run
// as tcgen05.mma with TMEM accumulators, and the first one is TRANSPOSED: D1 = Wkv (M = 128 K|V channels) x X^T
// (N = 64 pixels), so that a TMEM lane — hence an epilogue thread — owns one K or V channel across the whole pixel
// chunk.  The softmax over pixels then runs along the thread's own row: running maximum, sum and rescale factor live in
// registers, no cross-thread reduction.  The rows (fp16, pixels contiguous) are written to shared memory as K-major
// SWIZZLE_128B operands of the second GEMM  ctx[d, e] += P[d, 64 px] x V[e, 64 px]^T  (M = 128, N = 64 | 128, K = 64).
//
//   * C == 64 ("stacked"): one 128-row block holds K rows 0..63 and V rows 64..127; GEMM2 uses the same tile as A (rows
//     64.. are ignored garbage) and its rows 64..127 as B.
//   * C >= 128: a CTA owns the K rows of block kb and the V rows of block vb (128 rows each, rows >= C are TMA zero
//     fill); grid.y enumerates the (kb, vb) pairs.
//   * reference maximum m_d: taken from the first tile and raised only when a tile exceeds it by more than 8 (P stays
//     below e^8 = 2981, far inside fp16); raising it rescales the ctx row in TMEM (tcgen05.ld / st) — rare.
//   * warp roles as in igemm_tc.cuh: 0 = TMA producer (+ de-interleaves the per-pixel LayerNorm statistics into shared
//     memory), 1 = MMA issuer (+ TMEM alloc), 2..5 = epilogue (TMEM lane quadrant = warp % 4).
//   * D1 and the P/V tiles are double-buffered: GEMM1 of tile i+1 and GEMM2 of tile i-1 overlap the epilogue of tile i.
#pragma once
#include "igemm_tc.cuh"

namespace cdc {

struct AttnTcParams {
  int C, N;                 // channels, pixels per image (N % 64 == 0)
  int stacked;              // C == 64
  int cpt;                  // C / 64: K chunks of GEMM1
  int kbc;                  // 128-row blocks per K (and per V): ceil(C / 128); grid.y = kbc * kbc (1 when stacked)
  int ntiles;               // 64-pixel tiles per image
  int tiles_per_chunk, nchunks;
  int stages;               // X ring depth
  int pvbufs;               // P/V tile buffers: 2, or 1 when the resident weights leave no room (C = 320)
  const float2* stats;      // [B*N] (mean, rstd) of the x rows
  const float* u;           // [2C] row sums of the folded fp16 weights
  const float* c;           // [2C] W b_ln
  float* part_ctx;          // [B][nchunks][C][C]
  float* part_m;            // [B][nchunks][C]
  float* part_s;            // [B][nchunks][C]
  float* ctxn;              // nchunks == 1 only: write ctx / S straight to [B][C][C] (no combine pass); else nullptr
  __half* ctx16_hi;         // optional (with ctxn): fp16 value + remainder in attn_alg_tc_kernel's MN-blocked operand layout
  __half* ctx16_lo;
};

constexpr int kAttnTcThreads = 192;
constexpr int kAttnStatSlots = 16;
// dynamic smem: W tiles | X ring | P/V tiles (2 buffers) | statistics slots | barriers
__host__ __device__ inline int attn_tc_nblk(int stacked) { return stacked ? 1 : 2; }
__host__ __device__ inline int attn_tc_smem_bytes(int stacked, int cpt, int stages, int pvbufs) {
  const int nblk = attn_tc_nblk(stacked);
  return 1024 + nblk * cpt * 16384 + stages * 8192 + pvbufs * nblk * 16384 + kAttnStatSlots * 512 + 512;
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(kAttnTcThreads, 2)   // <= 168 registers: two CTAs per SM at C = 64
attn_ctx_tc_kernel(const __grid_constant__ TcMaps maps, const AttnTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const int nblk = attn_tc_nblk(p.stacked);
  const uint32_t sW = base;                                   // [blk][cc] 16 KB tiles (128 rows x 64 ch)
  const uint32_t sX = sW + nblk * p.cpt * 16384;              // [stage] 8 KB tiles (64 px x 64 ch)
  const uint32_t sPV = sX + p.stages * 8192;                  // [buf][blk] 16 KB tiles (128 rows x 64 px)
  const uint32_t sStat = sPV + p.pvbufs * nblk * 16384;       // [slot] negmean[64] | rstd[64]
  uint8_t* pPV = smem + (sPV - base);
  float* pStat = reinterpret_cast<float*>(smem + (sStat - base));
  const uint32_t bars = sStat + kAttnStatSlots * 512;
  const uint32_t bar_w = bars;                                // weights landed
  const uint32_t bar_xf = bars + 8;                           // [8] X stage full
  const uint32_t bar_xe = bar_xf + 64;                        // [8] X stage empty
  const uint32_t bar_d1f = bar_xe + 64;                       // [2] GEMM1 accumulator complete
  const uint32_t bar_d1e = bar_d1f + 16;                      // [2] GEMM1 accumulator read by the epilogue
  const uint32_t bar_pvf = bar_d1e + 16;                      // [2] P/V tile written
  const uint32_t bar_pve = bar_pvf + 16;                      // [2] GEMM2 done with the P/V tile
  const uint32_t bar_st = bar_pve + 16;                       // [16] statistics slot written
  const uint32_t bar_fin = bar_st + 8 * kAttnStatSlots;       // all MMAs retired
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + (bar_fin + 8 - base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int chunk = blockIdx.x, b = blockIdx.z;
  const int kb = p.stacked ? 0 : (int)blockIdx.y / p.kbc, vb = p.stacked ? 0 : (int)blockIdx.y % p.kbc;
  const int tile0 = chunk * p.tiles_per_chunk;
  const int ntl = min(p.tiles_per_chunk, p.ntiles - tile0);   // tiles of this CTA (>= 1 by construction of the grid)
  const int n2 = p.stacked ? 64 : 128;                        // GEMM2 N = ctx columns held by this CTA
  const int d1_cols = nblk * 64;
  const int ctx_col = 2 * d1_cols;
  int tmem_cols = 32;
  while (tmem_cols < ctx_col + n2) tmem_cols <<= 1;

  pdl_launch_dependents();
  if (threadIdx.x == 0) {
    tc::prefetch_tmap(&maps.a[0]);
    tc::prefetch_tmap(&maps.b[0]);
    tc::mbar_init(bar_w, 1);
    for (int s = 0; s < p.stages; ++s) {
      tc::mbar_init(bar_xf + 8 * s, 1);
      tc::mbar_init(bar_xe + 8 * s, 1);
    }
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(bar_d1f + 8 * i, 1);
      tc::mbar_init(bar_d1e + 8 * i, 128);
      tc::mbar_init(bar_pvf + 8 * i, 128);
      tc::mbar_init(bar_pve + 8 * i, 1);
    }
    for (int i = 0; i < kAttnStatSlots; ++i) tc::mbar_init(bar_st + 8 * i, 32);
    tc::mbar_init(bar_fin, 1);
    tc::fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)),
                 "r"(tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  pdl_wait();

  if (warp == 0) {
    // =============================== producer ===============================
    const bool leader = tc::elect_one();
    if (leader) {
      tc::mbar_expect_tx(bar_w, (uint32_t)(nblk * p.cpt * 16384));
      for (int blk = 0; blk < nblk; ++blk)
        for (int cc = 0; cc < p.cpt; ++cc) {
          // weight view {64, rows, kv, chunk}: stacked -> rows = 2C (K then V), kv extent 1
          const int row0 = p.stacked ? 0 : (blk == 0 ? kb : vb) * 128;
          const int kv = p.stacked ? 0 : blk;
          tc::tma_load_4d(sW + (blk * p.cpt + cc) * 16384, &maps.b[0], bar_w, 0, row0, kv, cc);
        }
    }
    __syncwarp();
    int stage = 0;
    uint32_t phase = 0;
    for (int i = 0; i < ntl; ++i) {
      const long long pix0 = (long long)b * p.N + (long long)(tile0 + i) * 64;
      // LayerNorm statistics of the tile's 64 pixels: lane l fetches pixels 2l, 2l+1
      const float4 st = *reinterpret_cast<const float4*>(p.stats + pix0 + 2 * lane);
      for (int cc = 0; cc < p.cpt; ++cc) {
        tc::mbar_wait(bar_xe + 8 * stage, phase ^ 1);
        if (leader) {
          tc::mbar_expect_tx(bar_xf + 8 * stage, 8192u);
          tc::tma_load_2d(sX + stage * 8192, &maps.a[0], bar_xf + 8 * stage, cc * 64, (int)pix0);
        }
        __syncwarp();
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1;
        }
      }
      // slot reuse distance (16 tiles) exceeds how far the producer can run ahead of the epilogue (stages + 3 tiles)
      float* slot = pStat + (i % kAttnStatSlots) * 128;
      *reinterpret_cast<float2*>(slot + 2 * lane) = make_float2(-st.x, -st.z);
      *reinterpret_cast<float2*>(slot + 64 + 2 * lane) = make_float2(st.y, st.w);
      // every lane releases its own two stores (count 32): no reliance on __syncwarp ordering other lanes' writes
      tc::mbar_arrive(bar_st + 8 * (i % kAttnStatSlots));
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    const bool leader = tc::elect_one();
    const uint32_t idesc1 = tc::make_idesc_f16(64);
    const uint32_t idesc2 = tc::make_idesc_f16(n2);
    const uint32_t desc_hi = (uint32_t)(tc::make_desc_sw128(0) >> 32);
    tc::mbar_wait(bar_w, 0);
    tc::tc_fence_after();
    int stage = 0;
    uint32_t phase = 0;
    auto gemm2 = [&](int j) {
      const int pb = p.pvbufs == 2 ? (j & 1) : 0;
      const int pu = p.pvbufs == 2 ? (j >> 1) : j;       // how often this buffer was used before
      tc::mbar_wait(bar_pvf + 8 * pb, (uint32_t)pu & 1u);
      tc::tc_fence_after();
      if (leader) {
        const uint32_t a_lo = (uint32_t)tc::make_desc_sw128(sPV + (pb * nblk) * 16384);
        const uint32_t b_lo = p.stacked ? a_lo + (8192u >> 4) : a_lo + (16384u >> 4);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          tc::umma_f16_lo(tmem_base + ctx_col, a_lo + ks * 2, b_lo + ks * 2, desc_hi, idesc2, (j > 0 || ks > 0) ? 1u : 0u);
        tc::umma_commit(bar_pve + 8 * pb);
      }
      __syncwarp();
    };
    for (int i = 0; i < ntl; ++i) {
      const int buf = i & 1;
      tc::mbar_wait(bar_d1e + 8 * buf, ((uint32_t)(i >> 1) & 1u) ^ 1u);
      tc::tc_fence_after();
      for (int cc = 0; cc < p.cpt; ++cc) {
        tc::mbar_wait(bar_xf + 8 * stage, phase);
        tc::tc_fence_after();
        if (leader) {
          const uint32_t x_lo = (uint32_t)tc::make_desc_sw128(sX + stage * 8192);
          for (int blk = 0; blk < nblk; ++blk) {
            const uint32_t w_lo = (uint32_t)tc::make_desc_sw128(sW + (blk * p.cpt + cc) * 16384);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              tc::umma_f16_lo(tmem_base + buf * d1_cols + blk * 64, w_lo + ks * 2, x_lo + ks * 2, desc_hi, idesc1,
                              (cc > 0 || ks > 0) ? 1u : 0u);
          }
          tc::umma_commit(bar_xe + 8 * stage);
        }
        __syncwarp();
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (leader) tc::umma_commit(bar_d1f + 8 * buf);
      __syncwarp();
      if (i > 0) gemm2(i - 1);   // GEMM1 of tile i is in flight before GEMM2 of tile i-1 waits for its P/V tile
    }
    gemm2(ntl - 1);
    if (leader) tc::umma_commit(bar_fin);
    __syncwarp();
  } else {
    // =============================== epilogue ===============================
    using tc::f32x2;
    const int quad = warp & 3;
    const int r = quad * 32 + lane;                       // TMEM lane = row of the 128-row blocks
    const uint32_t tlane = (uint32_t)(quad * 32) << 16;
    // row roles: block 0 = K rows (stacked: rows >= 64 are V rows), block 1 = V rows
    const bool row0_is_k = p.stacked ? (r < 64) : true;   // uniform per warp
    const int gk = p.stacked ? r : kb * 128 + r;          // K channel of my block-0 row (when it is a K row)
    const int gv = p.stacked ? r - 64 : vb * 128 + r;     // V channel of my V row
    const bool k_ok = row0_is_k && gk < p.C, v_ok = (p.stacked ? r >= 64 : true) && gv < p.C;
    const float uk = k_ok ? p.u[gk] : 0.f, ck = k_ok ? p.c[gk] : 0.f;
    const float uv = v_ok ? p.u[p.C + gv] : 0.f, cv = v_ok ? p.c[p.C + gv] : 0.f;
    const f32x2 uk2 = tc::pk(uk, uk), ck2 = tc::pk(ck, ck), uv2 = tc::pk(uv, uv), cv2 = tc::pk(cv, cv);
    const float kLog2e = 1.4426950408889634f;
    float m_ref = -INFINITY, S = 0.f;
    uint32_t v[64];

    for (int i = 0; i < ntl; ++i) {
      const int buf = i & 1;
      const ulonglong2* nm = reinterpret_cast<const ulonglong2*>(pStat + (i % kAttnStatSlots) * 128);   // -mean pairs
      const ulonglong2* rs = nm + 16;                                                                 // rstd pairs
      const int pb = p.pvbufs == 2 ? buf : 0;
      const int pu = p.pvbufs == 2 ? (i >> 1) : i;
      tc::mbar_wait(bar_pve + 8 * pb, ((uint32_t)pu & 1u) ^ 1u);   // the previous GEMM2 on this P/V buffer has retired
      tc::mbar_wait(bar_st + 8 * (i % kAttnStatSlots), (uint32_t)(i / kAttnStatSlots) & 1u);
      tc::mbar_wait(bar_d1f + 8 * buf, (uint32_t)(i >> 1) & 1u);
      tc::tc_fence_after();
      for (int blk = 0; blk < nblk; ++blk) {
        const uint32_t taddr = tmem_base + tlane + (uint32_t)(buf * d1_cols + blk * 64);
        tc::tmem_ld32_issue(taddr, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
        tc::tmem_ld32_issue(taddr + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
        tc::tmem_wait_ld();
        if (blk == nblk - 1) {   // last TMEM read of this tile's GEMM1 accumulator
          tc::tc_fence_before();
          tc::mbar_arrive(bar_d1e + 8 * buf);
        }
        const bool is_k = blk == 0 && row0_is_k;   // uniform per warp
        uint8_t* dst = pPV + (size_t)(pb * nblk + blk) * 16384 + r * 128;
        f32x2 x[32];
        if (is_k) {
          float tmax = -INFINITY;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const ulonglong2 a = nm[j], s2 = rs[j];
            x[2 * j] = tc::fma2(s2.x, tc::fma2(a.x, uk2, tc::pku(v[4 * j], v[4 * j + 1])), ck2);
            x[2 * j + 1] = tc::fma2(s2.y, tc::fma2(a.y, uk2, tc::pku(v[4 * j + 2], v[4 * j + 3])), ck2);
            const float2 f0 = tc::upk(x[2 * j]), f1 = tc::upk(x[2 * j + 1]);
            tmax = fmaxf(tmax, fmaxf(fmaxf(f0.x, f0.y), fmaxf(f1.x, f1.y)));
          }
          // lazily raised reference maximum; raising it rescales what has been accumulated so far
          const float m_new = (tmax > m_ref + 8.f) ? tmax : m_ref;   // first tile: m_ref = -inf -> tmax
          const bool raise = m_new != m_ref && i > 0;
          if (__any_sync(0xffffffffu, raise)) {
            const float alpha = raise ? __expf(m_ref - m_new) : 1.f;
            if (p.pvbufs == 2)   // GEMM2 of tile i-1 retired (with one buffer the wait above already covers it)
              tc::mbar_wait(bar_pve + 8 * ((i - 1) & 1), (uint32_t)((i - 1) >> 1) & 1u);
            tc::tc_fence_after();
            for (int c0 = 0; c0 < n2; c0 += 32) {
              uint32_t t[32];
              tc::tmem_ld32(tmem_base + tlane + (uint32_t)(ctx_col + c0), t);
#pragma unroll
              for (int q = 0; q < 32; ++q) t[q] = __float_as_uint(__uint_as_float(t[q]) * alpha);
              asm volatile(
                  "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
                  "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,"
                  "%28,%29,%30,%31,%32};"
                  ::"r"(tmem_base + tlane + (uint32_t)(ctx_col + c0)), "r"(t[0]), "r"(t[1]), "r"(t[2]), "r"(t[3]),
                  "r"(t[4]), "r"(t[5]), "r"(t[6]), "r"(t[7]), "r"(t[8]), "r"(t[9]), "r"(t[10]), "r"(t[11]), "r"(t[12]),
                  "r"(t[13]), "r"(t[14]), "r"(t[15]), "r"(t[16]), "r"(t[17]), "r"(t[18]), "r"(t[19]), "r"(t[20]),
                  "r"(t[21]), "r"(t[22]), "r"(t[23]), "r"(t[24]), "r"(t[25]), "r"(t[26]), "r"(t[27]), "r"(t[28]),
                  "r"(t[29]), "r"(t[30]), "r"(t[31])
                  : "memory");
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            S *= alpha;
          }
          m_ref = m_new;
          const float mb = -m_ref * kLog2e;
          const f32x2 l2 = tc::pk(kLog2e, kLog2e), mb2 = tc::pk(mb, mb);
#pragma unroll
          for (int j = 0; j < 8; ++j) {   // 8 pixels -> one 16-byte chunk of the row
            uint32_t h[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float2 e = tc::upk(tc::fma2(x[4 * j + q], l2, mb2));
              h[q] = pack_half2(ex2_approx(e.x), ex2_approx(e.y));
              const float2 g = unpack_half2(h[q]);   // sum what the MMA will actually see
              S += g.x + g.y;
            }
            *reinterpret_cast<uint4*>(dst + ((j ^ (r & 7)) << 4)) = make_uint4(h[0], h[1], h[2], h[3]);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            uint32_t h[4];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              const ulonglong2 a = nm[2 * j + q], s2 = rs[2 * j + q];
              const float2 f0 = tc::upk(tc::fma2(s2.x, tc::fma2(a.x, uv2, tc::pku(v[8 * j + 4 * q], v[8 * j + 4 * q + 1])), cv2));
              const float2 f1 =
                  tc::upk(tc::fma2(s2.y, tc::fma2(a.y, uv2, tc::pku(v[8 * j + 4 * q + 2], v[8 * j + 4 * q + 3])), cv2));
              h[2 * q] = pack_half2(f0.x, f0.y);
              h[2 * q + 1] = pack_half2(f1.x, f1.y);
            }
            *reinterpret_cast<uint4*>(dst + ((j ^ (r & 7)) << 4)) = make_uint4(h[0], h[1], h[2], h[3]);
          }
        }
      }
      tc::fence_proxy_async();   // generic-proxy smem writes -> visible to the tensor core (async proxy)
      tc::tc_fence_before();
      tc::mbar_arrive(bar_pvf + 8 * pb);
    }

    // ---- write the split partials: ctx rows of this CTA's K block x V block, reference maximum, sum ----
    tc::mbar_wait(bar_fin, 0);
    tc::tc_fence_after();
    if (row0_is_k) {   // uniform per warp
      const size_t pbase = (size_t)b * p.nchunks + chunk;
      const bool direct = p.ctxn != nullptr;   // single chunk: normalise here, the combine pass is skipped
      float* dst = (direct ? p.ctxn + ((size_t)b * p.C + (k_ok ? gk : 0)) * p.C
                           : p.part_ctx + (pbase * p.C + (k_ok ? gk : 0)) * p.C) + vb * 128;
      const float inv = direct ? 1.f / S : 1.f;
      const int ncol = min(n2, p.C - vb * 128);
      for (int c0 = 0; c0 < ncol; c0 += 32) {
        uint32_t t[32];
        tc::tmem_ld32(tmem_base + tlane + (uint32_t)(ctx_col + c0), t);
        if (k_ok) {
#pragma unroll
          for (int q = 0; q < 8; ++q)
            reinterpret_cast<float4*>(dst + c0)[q] =
                make_float4(__uint_as_float(t[4 * q]) * inv, __uint_as_float(t[4 * q + 1]) * inv,
                            __uint_as_float(t[4 * q + 2]) * inv, __uint_as_float(t[4 * q + 3]) * inv);
          if (direct && p.ctx16_hi) {   // 32 consecutive columns e of row d = gk: half of one 64-element MN block
            const int e0 = vb * 128 + c0;
            const size_t o16 = (((size_t)b * (p.C >> 6) + (e0 >> 6)) * p.C + gk) * 64 + (e0 & 63);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              float f[8];
#pragma unroll
              for (int k = 0; k < 8; ++k) f[k] = __uint_as_float(t[8 * q + k]) * inv;
              uint4 hi, lo;
              pack_out8(f, hi, lo, true);
              *reinterpret_cast<uint4*>(p.ctx16_hi + o16 + 8 * q) = hi;
              *reinterpret_cast<uint4*>(p.ctx16_lo + o16 + 8 * q) = lo;
            }
          }
        }
      }
      if (!direct && k_ok && vb == 0) {
        p.part_m[pbase * p.C + gk] = m_ref;
        p.part_s[pbase * p.C + gk] = S;
      }
    }
  }

  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc::tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols));
  }
}

}  // namespace cdc
