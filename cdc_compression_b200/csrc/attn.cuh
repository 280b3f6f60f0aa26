// Linear attention (heads=1, dim_head=C) with the PreNorm LayerNorm folded into GEMM epilogues.
// Reference: epsilonparam/modules/network_components.py:117-139 (+PreNorm :69-77, Residual :10-16).
//
//   K,V = rstd*(Wg x - mean*u) + c            (Wg = W diag(g_ln) in fp16, u = rowsum(Wg), c = W b_ln)
//   P   = exp(K - max_n K)  (softmax over pixels, per channel d);  S_d = sum_n P
//   ctx[d,e] = sum_n P[n,d] V[n,e] / S_d
//   out = M_b xn + b_out + x,  M_b = W_out ctx^T (C^-1/2 W_q)      
//
// attn_ctx_kernel   : one pass over x per (image, pixel chunk, 64x64 block of ctx): K/V GEMM,
//                     online softmax over pixels, P^T V accumulation; writes split partials.
// attn_combine_kernel, gemm3xf16_tn_kernel, attn_finish_kernel : per-image C x C algebra (fp32-grade: split-fp16 MMAs).
// The final GEMM (out = M_b-folded weights applied to raw x) runs in the generic conv kernel
// with EPI_AFFINE and per-image weights.
#pragma once
#include "common.cuh"

namespace cdc {

struct AttnCtxParams {
  const __half* x;      // [B, N, C] fp16 (NHWC)
  const float2* stats;  // [B*N] (mean, rstd) of x rows
  const __half* Wkv;    // [C/64][2C][64] fp16, LayerNorm gain folded in; rows [0,C) = K, [C,2C) = V
  const float* u;       // [2C] row sums of the folded fp16 weights
  const float* c;       // [2C] W b_ln
  int C, N;
  int tiles_per_chunk;  // 64-pixel tiles walked by one CTA
  int nchunks;
  float* part_ctx;      // [B][nchunks][C][C]
  float* part_m;        // [B][nchunks][C]
  float* part_s;        // [B][nchunks][C]
  float* ctxn;          // nchunks == 1 only: write ctx / S straight to [B][C][C] (no combine pass); else nullptr
  __half* ctx16_hi;     // optional (with ctxn): the same matrix as fp16 value + remainder in the MN-blocked operand layout of
  __half* ctx16_lo;     // attn_alg_tc_kernel: element (d, e) at ((e/64) * C + d) * 64 + e % 64, per image
};

struct AttnCtxSmem {
  static constexpr int kStage = 64 * 128 + 128 * 128;  // x tile + (K|V) weight tile
  static constexpr int kKs = 64 * 65 * 4;
  static constexpr int kPs = 64 * 72 * 2;
  static constexpr int kVs = 64 * 72 * 2;
  static constexpr int kVec = 4 * 64 * 4;  // m_run, alpha, S, spare
  static constexpr int kBytes = 2 * kStage + kKs + kPs + kVs + kVec;
};

__global__ void __launch_bounds__(256) attn_ctx_kernel(const AttnCtxParams p) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t smem_base = smem_u32(smem);
  float* Ks = reinterpret_cast<float*>(smem + 2 * AttnCtxSmem::kStage);
  __half* Ps = reinterpret_cast<__half*>(smem + 2 * AttnCtxSmem::kStage + AttnCtxSmem::kKs);
  __half* Vs = reinterpret_cast<__half*>(smem + 2 * AttnCtxSmem::kStage + AttnCtxSmem::kKs + AttnCtxSmem::kPs);
  float* m_run = reinterpret_cast<float*>(smem + 2 * AttnCtxSmem::kStage + AttnCtxSmem::kKs + AttnCtxSmem::kPs +
                                          AttnCtxSmem::kVs);
  float* alpha = m_run + 128;   // m_run holds two copies (ping-pong per tile)
  float* s_run = m_run + 192;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 2, wn = warp & 3;
  const int C = p.C, N = p.N;
  const int cb = C >> 6;
  const int chunk = blockIdx.x;
  const int dblk = blockIdx.y / cb, eblk = blockIdx.y - dblk * cb;
  const int b = blockIdx.z;
  const int pix_begin = chunk * p.tiles_per_chunk * 64;
  int ntiles = p.tiles_per_chunk;
  {
    const int remaining = (N - pix_begin + 63) / 64;
    if (ntiles > remaining) ntiles = remaining;
  }
  const __half* xb = p.x + (size_t)b * N * C;
  const float2* stb = p.stats + (size_t)b * N;

  if (tid < 64) {
    m_run[tid] = -INFINITY;
    m_run[64 + tid] = -INFINITY;
    s_run[tid] = 0.f;
    alpha[tid] = 0.f;
  }

  float ctx[2][2][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int k = 0; k < 4; ++k) ctx[i][j][k] = 0.f;

  const int total = ntiles * cb;
  const int lchunk = tid & 7;

  auto load = [&](int it, int stage) {
    const int t = it / cb, cc = it - t * cb;
    const uint32_t sX = smem_base + stage * AttnCtxSmem::kStage;
    const uint32_t sW = sX + 64 * 128;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int row = (tid >> 3) + 32 * i;
      const int pix = pix_begin + t * 64 + row;
      const bool ok = pix < N;
      const __half* src = ok ? xb + (size_t)pix * C + cc * 64 + lchunk * 8 : xb;
      cp_async16(sX + swz128(row, lchunk), src, ok ? 16 : 0);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int row = (tid >> 3) + 32 * i;  // 0..127
      const int grow = row < 64 ? dblk * 64 + row : C + eblk * 64 + (row - 64);
      cp_async16(sW + swz128(row, lchunk), p.Wkv + ((size_t)cc * 2 * C + grow) * 64 + lchunk * 8, 16);
    }
  };

  float acc[2][4][4];
  if (total > 0) load(0, 0);
  cp_async_commit();
  __syncthreads();  // m_run / s_run initialised

  for (int it = 0; it < total; ++it) {
    const int t = it / cb, cc = it - t * cb;
    if (it + 1 < total) load(it + 1, (it + 1) & 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    if (cc == 0) {
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int k = 0; k < 4; ++k) acc[i][j][k] = 0.f;
    }
    {
      const uint32_t sX = smem_base + (it & 1) * AttnCtxSmem::kStage;
      const uint32_t sW = sX + 64 * 128;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        uint32_t af[2][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
          ldmatrix_x4(af[mt], sX + swz128(wm * 32 + mt * 16 + (lane & 15), ks * 2 + (lane >> 4)));
#pragma unroll
        for (int np = 0; np < 2; ++np) {
          const int n = wn * 32 + np * 16 + (lane & 7) + ((lane >> 4) << 3);
          uint32_t bf[4];
          ldmatrix_x4(bf, sW + swz128(n, ks * 2 + ((lane >> 3) & 1)));
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) {
            mma_16816(acc[mt][2 * np], af[mt], bf[0], bf[1]);
            mma_16816(acc[mt][2 * np + 1], af[mt], bf[2], bf[3]);
          }
        }
      }
    }
    if (cc == cb - 1) {
      // ---- K/V epilogue for this 64-pixel tile ----
      const int tile_pix = pix_begin + t * 64;
      const int par = t & 1;                       // m_run is ping-ponged per tile (read old / write new)
      float cmx[4][2];                             // per-thread column max over its 4 rows (K warps)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) cmx[nt][0] = cmx[nt][1] = -INFINITY;
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int row = wm * 32 + mt * 16 + (lane >> 2) + h * 8;
          const int pix = tile_pix + row;
          const bool valid = pix < N;
          float2 st = make_float2(0.f, 0.f);
          if (valid) st = stb[pix];
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) {
            const int col = wn * 32 + nt * 8 + (lane & 3) * 2;  // 0..127
            const int gcol = col < 64 ? dblk * 64 + col : C + eblk * 64 + (col - 64);
            float v0 = st.y * (acc[mt][nt][2 * h] - st.x * p.u[gcol]) + p.c[gcol];
            float v1 = st.y * (acc[mt][nt][2 * h + 1] - st.x * p.u[gcol + 1]) + p.c[gcol + 1];
            if (wn < 2) {
              if (!valid) {
                v0 = -INFINITY;
                v1 = -INFINITY;
              }
              acc[mt][nt][2 * h] = v0;
              acc[mt][nt][2 * h + 1] = v1;
              cmx[nt][0] = fmaxf(cmx[nt][0], v0);
              cmx[nt][1] = fmaxf(cmx[nt][1], v1);
            } else {
              if (!valid) {
                v0 = 0.f;
                v1 = 0.f;
              }
              *reinterpret_cast<uint32_t*>(Vs + row * 72 + (col - 64)) = pack_half2(v0, v1);
            }
          }
        }
      if (wn < 2) {  // column max over the warp's 32 rows: lanes sharing (lane & 3) hold the same columns
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
          for (int e2 = 0; e2 < 2; ++e2) {
            float m = cmx[nt][e2];
            m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 4));
            m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 8));
            m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 16));
            if ((lane >> 2) == 0) Ks[wm * 64 + wn * 32 + nt * 8 + (lane & 3) * 2 + e2] = m;   // cmax[wm][col]
          }
      }
      __syncthreads();
      if (wn < 2) {
        float csum[4][2];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          csum[nt][0] = csum[nt][1] = 0.f;
#pragma unroll
          for (int e2 = 0; e2 < 2; ++e2) {
            const int col = wn * 32 + nt * 8 + (lane & 3) * 2 + e2;
            const float m_old = m_run[par * 64 + col];
            const float m_new = fmaxf(m_old, fmaxf(Ks[col], Ks[64 + col]));
            cmx[nt][e2] = m_new;
            if (wm == 0 && (lane >> 2) == 0) {   // one writer per column
              m_run[(par ^ 1) * 64 + col] = m_new;
              alpha[col] = (m_old == -INFINITY) ? 0.f : __expf(m_old - m_new);
            }
          }
        }
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int row = wm * 32 + mt * 16 + (lane >> 2) + h * 8;
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
              const int col = wn * 32 + nt * 8 + (lane & 3) * 2;
              const float p0 = __expf(acc[mt][nt][2 * h] - cmx[nt][0]);       // exp(-inf) = 0 for padded rows
              const float p1 = __expf(acc[mt][nt][2 * h + 1] - cmx[nt][1]);
              const uint32_t hv = pack_half2(p0, p1);
              *reinterpret_cast<uint32_t*>(Ps + row * 72 + col) = hv;
              const float2 q = unpack_half2(hv);   // sum what the MMA will actually see
              csum[nt][0] += q.x;
              csum[nt][1] += q.y;
            }
          }
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
          for (int e2 = 0; e2 < 2; ++e2) {
            float sv = csum[nt][e2];
            sv += __shfl_xor_sync(0xffffffffu, sv, 4);
            sv += __shfl_xor_sync(0xffffffffu, sv, 8);
            sv += __shfl_xor_sync(0xffffffffu, sv, 16);
            if ((lane >> 2) == 0) Ks[128 + wm * 64 + wn * 32 + nt * 8 + (lane & 3) * 2 + e2] = sv;  // csum[wm][col]
          }
      }
      __syncthreads();
      if (tid < 64) s_run[tid] = s_run[tid] * alpha[tid] + Ks[128 + tid] + Ks[192 + tid];
      // ---- ctx[d,e] = ctx*alpha[d] + P^T V ; warps 2(d) x 4(e), warp tile 32 x 16 ----
      {
        const int wd = warp >> 2, we = warp & 3;
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const float a = alpha[wd * 32 + mt * 16 + (lane >> 2) + h * 8];
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
              ctx[mt][nt][2 * h] *= a;
              ctx[mt][nt][2 * h + 1] *= a;
            }
          }
        const uint32_t sP = smem_u32(Ps), sV = smem_u32(Vs);
        const int j = lane >> 3;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          uint32_t bf[4];
          {
            const int prow = ks * 16 + (lane & 7) + 8 * (j & 1);
            const int ecol = we * 16 + 8 * (j >> 1);
            ldmatrix_x4_trans(bf, sV + (prow * 72 + ecol) * 2);
          }
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) {
            uint32_t af[4];
            const int prow = ks * 16 + (lane & 7) + 8 * (j >> 1);
            const int dcol = wd * 32 + mt * 16 + 8 * (j & 1);
            ldmatrix_x4_trans(af, sP + (prow * 72 + dcol) * 2);
            mma_16816(ctx[mt][0], af, bf[0], bf[1]);
            mma_16816(ctx[mt][1], af, bf[2], bf[3]);
          }
        }
      }
    }
    __syncthreads();
  }
  cp_async_wait<0>();

  // ---- write split partials ----
  {
    const int wd = warp >> 2, we = warp & 3;
    const bool direct = p.ctxn != nullptr;   // single chunk: normalise here, the combine pass is skipped
    float* dst = direct ? p.ctxn + (size_t)b * C * C : p.part_ctx + ((size_t)b * p.nchunks + chunk) * C * C;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int dl = wd * 32 + mt * 16 + (lane >> 2) + h * 8;
        const int d = dblk * 64 + dl;
        const float inv = direct ? 1.f / s_run[dl] : 1.f;
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
          const int e = eblk * 64 + we * 16 + nt * 8 + (lane & 3) * 2;
          const float v0 = ctx[mt][nt][2 * h] * inv, v1 = ctx[mt][nt][2 * h + 1] * inv;
          *reinterpret_cast<float2*>(dst + (size_t)d * C + e) = make_float2(v0, v1);
          if (direct && p.ctx16_hi) {
            const size_t o16 = (((size_t)b * cb + (e >> 6)) * C + d) * 64 + (e & 63);
            const uint32_t hv = pack_half2(v0, v1);
            const float2 hf = unpack_half2(hv);
            *reinterpret_cast<uint32_t*>(p.ctx16_hi + o16) = hv;
            *reinterpret_cast<uint32_t*>(p.ctx16_lo + o16) = pack_half2(v0 - hf.x, v1 - hf.y);
          }
        }
      }
    if (!direct && eblk == 0 && tid < 64) {
      const size_t o = ((size_t)b * p.nchunks + chunk) * C + dblk * 64 + tid;
      p.part_m[o] = m_run[(ntiles & 1) * 64 + tid];
      p.part_s[o] = s_run[tid];
    }
  }
}

// ctxn[b][d][e] = sum_c part_ctx[b][c][d][e] * exp(m_c[d]-m[d]) / sum_c part_s[b][c][d]*exp(m_c[d]-m[d])
// One CTA per (d, image).  Warp 0 turns the (<= 256) per-chunk maxima / sums of row d into normalised chunk weights
// (lanes over chunks, shuffle reductions: fixed order, deterministic); every thread then owns columns e and adds the
// weighted chunk partials in chunk order with four loads in flight.
constexpr int kCombineMaxChunks = 256;
__global__ void __launch_bounds__(128) attn_combine_kernel(const float* __restrict__ part_ctx,
                                                           const float* __restrict__ part_m,
                                                           const float* __restrict__ part_s, int C, int nchunks,
                                                           float* __restrict__ ctxn, __half* __restrict__ ctx16_hi,
                                                           __half* __restrict__ ctx16_lo) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float w_s[kCombineMaxChunks];
  const int d = blockIdx.x, b = blockIdx.y;
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    const float* pm = part_m + (size_t)b * nchunks * C + d;
    const float* ps = part_s + (size_t)b * nchunks * C + d;
    // lane l owns chunks l, l + 32, ... (<= kCombineMaxChunks / 32 each); fixed order, shuffle reductions: deterministic
    float mv[kCombineMaxChunks / 32];
    float m = -INFINITY;
#pragma unroll
    for (int j = 0; j < kCombineMaxChunks / 32; ++j) {
      const int c = lane + 32 * j;
      mv[j] = c < nchunks ? pm[(size_t)c * C] : -INFINITY;
      m = fmaxf(m, mv[j]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float S = 0.f;
#pragma unroll
    for (int j = 0; j < kCombineMaxChunks / 32; ++j) {
      const int c = lane + 32 * j;
      mv[j] = c < nchunks ? __expf(mv[j] - m) : 0.f;
      S += c < nchunks ? ps[(size_t)c * C] * mv[j] : 0.f;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) S += __shfl_xor_sync(0xffffffffu, S, o);
    const float inv = 1.f / S;
#pragma unroll
    for (int j = 0; j < kCombineMaxChunks / 32; ++j) w_s[lane + 32 * j] = mv[j] * inv;
  }
  __syncthreads();
  const size_t cstride = (size_t)C * C;
  const float* src = part_ctx + (size_t)b * nchunks * cstride + (size_t)d * C;
  for (int e = threadIdx.x; e < C; e += blockDim.x) {
    float a = 0.f;
    int c = 0;
    for (; c + 4 <= nchunks; c += 4) {
      const float v0 = src[(size_t)c * cstride + e], v1 = src[(size_t)(c + 1) * cstride + e];
      const float v2 = src[(size_t)(c + 2) * cstride + e], v3 = src[(size_t)(c + 3) * cstride + e];
      a = fmaf(v0, w_s[c], a);
      a = fmaf(v1, w_s[c + 1], a);
      a = fmaf(v2, w_s[c + 2], a);
      a = fmaf(v3, w_s[c + 3], a);
    }
    for (; c < nchunks; ++c) a = fmaf(src[(size_t)c * cstride + e], w_s[c], a);
    ctxn[((size_t)b * C + d) * C + e] = a;
    if (ctx16_hi) {   // operand layout of attn_alg_tc_kernel (see AttnCtxParams::ctx16_hi)
      const size_t o16 = (((size_t)b * (C >> 6) + (e >> 6)) * C + d) * 64 + (e & 63);
      const __half hv = __float2half_rn(a);
      ctx16_hi[o16] = hv;
      ctx16_lo[o16] = __float2half_rn(a - __half2float(hv));
    }
  }
}

// Batched GEMM  Cout[b][m][n] = sum_k At[b][k][m] * Bm[b][k][n]   (M % 64 == 0, N % 64 == 0, K % 32 == 0) with fp32
// inputs and outputs on the tensor cores: every operand is split into an fp16 value and its fp16 rounding remainder
// (x = hi + lo, |lo| <= 2^-11 |x|) and the product is accumulated as lo*hi + hi*lo + hi*hi in fp32 (relative error
// ~2^-21, i.e. fp32-grade — a single fp16 pass here costs 1e-4 on the U-Net output).  m16n8k16 fp16 MMAs: half the
// instruction count of a 3xTF32 (m16n8k8) form (measured: 25 -> 19 us at C = 384; the legacy mma.sync rate bounds it).
// These are the per-image C x C products behind M_b = W_out ctx^T (C^-1/2 W_q).
// 64 x 64 tile per CTA, 4 warps of 32 x 32, K slabs of 64 in a 3-stage cp.async ring (102 KB: two CTAs per SM).
struct Gemm3xSmem {
  static constexpr int kLD = 68;      // padded row (floats): the k-pair fragment loads hit 32 distinct banks
  static constexpr int kSlab = 64;    // K per pipeline stage (K % 64 == 0: C is a multiple of 64); 2 x 35 KB in flight per
                                      // CTA — the loop is load-latency bound (round 1: 32-wide slabs, 17 us at C = 384)
  static constexpr int kStages = 3;
  static constexpr int kBytes = 2 * kStages * kSlab * kLD * 4;
};

// (x0, x1) -> packed fp16 pair hi and packed remainders lo
__device__ __forceinline__ void split_f16x2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  hi = pack_half2(x0, x1);
  const float2 h = unpack_half2(hi);
  lo = pack_half2(x0 - h.x, x1 - h.y);
}

// Optional fused "finish" epilogue of the second product (M_b = W_out T): instead of the fp32 matrix the kernel writes
//   Mg16[o][c] = half(M[o][c] * g[c])   in the conv kernel's per-image weight layout [C/64][C][64], and per 64-column
//   tile t the partial row sums  um_part[b][t][o] = sum_c float(Mg16[o][c]),  cm_part[b][t][o] = sum_c M[o][c] b_ln[c]
//   (+ b_out[o] in tile 0); the consumer (EPI_AFFINE epilogue of igemm_tc_kernel) adds the C/64 partials in tile order.
struct GemmFinish {
  const float* g;      // [C] LayerNorm gain; nullptr = plain fp32 output
  const float* bln;    // [C]
  const float* bout;   // [C]
  __half* Mg16;        // [B][C/64][C][64]
  float* um_part;      // [B][C/64][C]
  float* cm_part;
};

__global__ void __launch_bounds__(128) gemm3xf16_tn_kernel(const float* __restrict__ At, const float* __restrict__ Bm,
                                                           float* __restrict__ Cout, int M, int N, int K,
                                                           long long sA, long long sB, long long sC,
                                                           const GemmFinish fin) {
  constexpr int LD = Gemm3xSmem::kLD, KS = Gemm3xSmem::kSlab, ST = Gemm3xSmem::kStages;
  extern __shared__ __align__(16) uint8_t gsm[];
  float (*As)[KS][LD] = reinterpret_cast<float (*)[KS][LD]>(gsm);
  float (*Bs)[KS][LD] = reinterpret_cast<float (*)[KS][LD]>(gsm + ST * KS * LD * 4);
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.z;
  At += (size_t)b * sA;
  Bm += (size_t)b * sB;
  Cout += (size_t)b * sC;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = (warp >> 1) * 32, wn = (warp & 1) * 32;
  const int g = lane >> 2, q = lane & 3;
  float acc[2][4][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[i][j][k] = 0.f;
  auto load = [&](int k0, int buf) {
#pragma unroll
    for (int v = 0; v < KS / 8; ++v) {
      const int idx = tid + v * 128;          // 16-byte piece inside the KS x 64 slab
      const int kk = idx >> 4, c = (idx & 15) * 4;
      cp_async16(smem_u32(&As[buf][kk][c]), At + (size_t)(k0 + kk) * M + m0 + c, 16);
      cp_async16(smem_u32(&Bs[buf][kk][c]), Bm + (size_t)(k0 + kk) * N + n0 + c, 16);
    }
  };
  const int slabs = K / KS;
#pragma unroll
  for (int i = 0; i < ST - 1; ++i) {
    if (i < slabs) load(i * KS, i);
    cp_async_commit();
  }
  for (int sidx = 0; sidx < slabs; ++sidx) {
    const int buf = sidx % ST;
    cp_async_wait<ST - 2>();
    __syncthreads();   // slab sidx landed for everyone; everyone finished reading slab sidx - 1
    if (sidx + ST - 1 < slabs) load((sidx + ST - 1) * KS, (sidx + ST - 1) % ST);
    cp_async_commit();
#pragma unroll
    for (int ks = 0; ks < KS / 16; ++ks) {
      const int k0 = ks * 16 + 2 * q;   // this lane's k pairs (k0, k0+1) and (k0+8, k0+9)
      uint32_t ah[2][4], al[2][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const int m = wm + mt * 16 + g;
        split_f16x2(As[buf][k0][m], As[buf][k0 + 1][m], ah[mt][0], al[mt][0]);
        split_f16x2(As[buf][k0][m + 8], As[buf][k0 + 1][m + 8], ah[mt][1], al[mt][1]);
        split_f16x2(As[buf][k0 + 8][m], As[buf][k0 + 9][m], ah[mt][2], al[mt][2]);
        split_f16x2(As[buf][k0 + 8][m + 8], As[buf][k0 + 9][m + 8], ah[mt][3], al[mt][3]);
      }
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int n = wn + nt * 8 + g;
        uint32_t bh0, bl0, bh1, bl1;
        split_f16x2(Bs[buf][k0][n], Bs[buf][k0 + 1][n], bh0, bl0);
        split_f16x2(Bs[buf][k0 + 8][n], Bs[buf][k0 + 9][n], bh1, bl1);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          mma_16816(acc[mt][nt], al[mt], bh0, bh1);   // small terms first
          mma_16816(acc[mt][nt], ah[mt], bl0, bl1);
          mma_16816(acc[mt][nt], ah[mt], bh0, bh1);
        }
      }
    }
  }
  if (fin.g == nullptr) {
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int m = m0 + wm + mt * 16 + g, n = n0 + wn + nt * 8 + q * 2;
        *reinterpret_cast<float2*>(Cout + (size_t)m * N + n) = make_float2(acc[mt][nt][0], acc[mt][nt][1]);
        *reinterpret_cast<float2*>(Cout + (size_t)(m + 8) * N + n) = make_float2(acc[mt][nt][2], acc[mt][nt][3]);
      }
    return;
  }
  // ---- fused finish (M == N == C, this CTA = rows [m0, +64) x column tile blockIdx.x) ----
  __syncthreads();                                   // the pipeline buffers are free: reuse them for the row sums
  float* red = reinterpret_cast<float*>(gsm);        // [2 (u|c)][2 (wn half)][64 rows]
  const int ct = blockIdx.x;
  __half* mg = fin.Mg16 + (size_t)b * M * N + (size_t)ct * M * 64;
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int ml = wm + mt * 16 + g + h * 8;       // row inside the tile
      float su = 0.f, sc = 0.f;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int nl = wn + nt * 8 + q * 2, n = n0 + nl;
        const float v0 = acc[mt][nt][2 * h], v1 = acc[mt][nt][2 * h + 1];
        const uint32_t hv = pack_half2(v0 * fin.g[n], v1 * fin.g[n + 1]);
        *reinterpret_cast<uint32_t*>(mg + (size_t)(m0 + ml) * 64 + nl) = hv;
        const float2 f = unpack_half2(hv);
        su += f.x + f.y;
        sc += v0 * fin.bln[n] + v1 * fin.bln[n + 1];
      }
      su += __shfl_xor_sync(0xffffffffu, su, 1);
      su += __shfl_xor_sync(0xffffffffu, su, 2);
      sc += __shfl_xor_sync(0xffffffffu, sc, 1);
      sc += __shfl_xor_sync(0xffffffffu, sc, 2);
      if (q == 0) {
        red[(0 * 2 + (warp & 1)) * 64 + ml] = su;
        red[(1 * 2 + (warp & 1)) * 64 + ml] = sc;
      }
    }
  __syncthreads();
  if (tid < 64) {
    const int m = m0 + tid;
    const size_t o = ((size_t)b * (N >> 6) + ct) * M + m;
    fin.um_part[o] = red[tid] + red[64 + tid];
    fin.cm_part[o] = red[128 + tid] + red[192 + tid] + (ct == 0 ? fin.bout[m] : 0.f);
  }
}

// Per (image, output row o): Mg16 = half(M[o][:] * g), um = rowsum(float(Mg16)), cm = M[o][:].b_ln + b_out[o].
// Mg16 is written in the conv kernel's weight layout [C/64][C][64] (per image).
__global__ void attn_finish_kernel(const float* __restrict__ Mf, const float* __restrict__ g,
                                   const float* __restrict__ bln, const float* __restrict__ bout, int C,
                                   __half* __restrict__ Mg16, float* __restrict__ um, float* __restrict__ cm) {
  pdl_launch_dependents();
  pdl_wait();
  const int o = blockIdx.x, b = blockIdx.y;
  const float* row = Mf + ((size_t)b * C + o) * C;
  __half* dst = Mg16 + (size_t)b * C * C;
  float su = 0.f, sc = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float m = row[c];
    const __half hv = __float2half_rn(m * g[c]);
    dst[((size_t)(c >> 6) * C + o) * 64 + (c & 63)] = hv;
    su += __half2float(hv);
    sc += m * bln[c];
  }
  __shared__ float r1[32], r2[32];
  for (int off = 16; off > 0; off >>= 1) {
    su += __shfl_xor_sync(0xffffffffu, su, off);
    sc += __shfl_xor_sync(0xffffffffu, sc, off);
  }
  if ((threadIdx.x & 31) == 0) {
    r1[threadIdx.x >> 5] = su;
    r2[threadIdx.x >> 5] = sc;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, c2 = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) {
      a += r1[i];
      c2 += r2[i];
    }
    um[(size_t)b * C + o] = a;
    cm[(size_t)b * C + o] = c2 + bout[o];
  }
}

}  // namespace cdc
