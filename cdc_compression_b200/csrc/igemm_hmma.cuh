// Implicit-GEMM convolution, baseline mainloop (mma.sync m16n8k16, cp.async 3-stage ring).
//
// One kernel family covers every convolution-shaped op of the denoiser U-Net:
//   Block conv3x3/7x7 (+bias -> channel LayerNorm -> ReLU -> +temb shift | +residual)
//       reference: epsilonparam/modules/network_components.py:83-91, 107-114
//   Downsample conv3x3 s2 (:45-53), Upsample convT4x4 s2 as 4 phase 2x2 convs (:34-42),
//   1x1 res_conv (:105), and the attention output GEMM with the PreNorm LayerNorm folded into an
//   affine epilogue (:69-77, :117-139; SURVEY.md Appendix E re-association).
//
// GEMM view: rows = output pixels (NHWC fp16 activations), columns = output channels,
// K = taps x input channels walked in 64-channel chunks; up to 3 K-segments (sources) so
// torch.cat inputs (unet.py:98,113) are never materialised.  fp16 operands, fp32 accumulation,
// LayerNorm statistics taken from the fp32 accumulators.
#pragma once
#include "common.cuh"

namespace cdc {

constexpr int kMaxSeg = 8;
enum EpiKind { EPI_BIAS = 0, EPI_LN_SHIFT = 1, EPI_LN_RES = 2, EPI_AFFINE = 3 };

struct ConvSeg {
  const __half* src;  // NHWC fp16 [B, Hs, Ws, C]
  int C;              // channels, multiple of 64
  int kh, kw;         // tap grid
  int dy0, dx0;       // input offset of tap (0,0)
  int nchunk;         // kh*kw*(C/64)
  // tcgen05 path only: a ResnetBlock's 1x1 res_conv fused into block2 as extra K segments that accumulate into a SECOND
  // TMEM accumulator (acc = 1); W = this segment's own weight chunks (null: the op's W + preceding chunks)
  int acc;
  const __half* W;
  int wshared;        // with per-image weights (groups > 1): this segment's weights are shared by all images
  // packed network input in window form: src is [B][Hs][Ws + 8][8] fp16 (3 zero pixels left, 5 right) and "channel" k of
  // pixel x is element k of the 64 halves starting at pixel x (k = kx * 8 + c: the horizontal taps x + kx - 3 of the 7 x 7
  // convolution) — an overlapping-window view, 8x smaller than the materialised [B][Hs][Ws][64] tensor
  int win8;
};

struct ConvParams {
  ConvSeg seg[kMaxSeg];
  int nseg;
  int Hs, Ws;            // source spatial size (all segments)
  int Ho, Wo;            // tile-space output size; GEMM rows enumerate (b, oy, ox)
  int stride;            // input pixel = out*stride + tap offset
  int rows_per_group;    // rows per group (groups==1: B*Ho*Wo; else Ho*Wo with group == image)
  int groups;
  long long w_group_stride;  // halfs between per-group weight sets (per-image attention matrices)
  const __half* W;       // [chunk][Ntot][64]
  int Ntot;
  int total_chunks;
  int phases;            // 4 => blockIdx.z is the transposed-conv output phase (py,px)
  long long w_phase_stride;
  __half* out;           // NHWC fp16 [B, out_H, out_W, Ntot]
  __half* out_lo;        // optional compensation tensor: fp16(value - fp16(value)) ("trunk" activations)
  int out_H, out_W, out_sy, out_sx;
  const float* bias;     // [Ntot] or null
  const float* ln_g;     // LayerNorm affine [Ntot]
  const float* ln_b;
  const float* shift;    // [B][shift_stride] additive per-image shift (timestep MLP), or null
  int shift_stride;
  const __half* res;     // residual, same pixel indexing as out, or null: channels [0,res_C0) ...
  int res_C0;            //   ... and channels [res_C0,Ntot) from res2 (identity residual of a torch.cat input)
  const __half* res2;
  const __half* res_lo;  // compensation halves of res / res2 (null when the residual source has none)
  const __half* res2_lo;
  const float2* stats_in;  // EPI_AFFINE: (mean, rstd) per pixel
  const float* aff_u;      // EPI_AFFINE: [groups][Ntot]
  const float* aff_c;
  int aff_group_stride;
  int aff_parts;
  int res_acc;             // tcgen05 path only: residual = second accumulator (+ res_bias) instead of res / res2
  const float* res_bias;
  __half* ln_out;          // tcgen05 path only (see TcConvParams::ln_out)
  int skip_out;           // tcgen05 path only: aff_u / aff_c hold this many partial sums per image (else 0 / 1)
  float2* stats_out;     // optional: LayerNorm stats of the stored output rows
};

// residual (hi + optional lo) for channels col, col+1 of output pixel pix
template <class P>
__device__ __forceinline__ float2 load_res2(const P& p, long long pix, int col) {
  const __half* hi;
  const __half* lo;
  if (col < p.res_C0) {
    const size_t o = (size_t)pix * p.res_C0 + col;
    hi = p.res + o;
    lo = p.res_lo ? p.res_lo + o : nullptr;
  } else {
    const size_t o = (size_t)pix * (p.Ntot - p.res_C0) + (col - p.res_C0);
    hi = p.res2 + o;
    lo = p.res2_lo ? p.res2_lo + o : nullptr;
  }
  float2 r = unpack_half2(*reinterpret_cast<const uint32_t*>(hi));
  if (lo) {
    const float2 l = unpack_half2(*reinterpret_cast<const uint32_t*>(lo));
    r.x += l.x;
    r.y += l.y;
  }
  return r;
}
// store v0,v1 as fp16 (+ the rounding remainder into the compensation tensor); returns the stored hi halves
template <class P>
__device__ __forceinline__ uint32_t store_out2(const P& p, size_t off, float v0, float v1) {
  const uint32_t hv = pack_half2(v0, v1);
  *reinterpret_cast<uint32_t*>(p.out + off) = hv;
  if (p.out_lo) {
    const float2 q = unpack_half2(hv);
    *reinterpret_cast<uint32_t*>(p.out_lo + off) = pack_half2(v0 - q.x, v1 - q.y);
  }
  return hv;
}

template <int BM, int BN>
struct IgemmSmem {
  static constexpr int kStages = 3;
  static constexpr int kABytes = BM * 128;
  static constexpr int kBBytes = BN * 128;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBytes = kStages * kStageBytes;
};

template <int BM, int BN, int EPI>
__global__ void __launch_bounds__(256) igemm_hmma_kernel(const ConvParams p) {
  constexpr int WN = 4;
  constexpr int MT = BM / 2 / 16;
  constexpr int NT = BN / WN / 8;
  static_assert(NT % 2 == 0, "n8 tiles are loaded in pairs");
  constexpr int STAGES = IgemmSmem<BM, BN>::kStages;
  constexpr int A_BYTES = IgemmSmem<BM, BN>::kABytes;
  constexpr int STAGE_BYTES = IgemmSmem<BM, BN>::kStageBytes;
  constexpr int AR = BM / 32;  // A rows per thread per chunk
  constexpr int BR = BN / 32;  // B rows per thread per chunk

  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ float red_a[WN][BM];
  __shared__ float red_b[WN][BM];

  pdl_launch_dependents();
  pdl_wait();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp / WN, wn = warp % WN;
  const uint32_t smem_base = smem_u32(smem);

  const int tiles_per_group = (p.rows_per_group + BM - 1) / BM;
  const int group = blockIdx.x / tiles_per_group;
  const int row0 = (blockIdx.x - group * tiles_per_group) * BM;
  const int n0 = blockIdx.y * BN;
  int py = 0, px = 0;
  const __half* W = p.W + (size_t)group * p.w_group_stride;
  if (p.phases) {
    py = blockIdx.z >> 1;
    px = blockIdx.z & 1;
    W += (size_t)blockIdx.z * p.w_phase_stride;
  }
  const int dyp = p.phases ? py - 1 : 0;
  const int dxp = p.phases ? px - 1 : 0;
  const int HoWo = p.Ho * p.Wo;

  // ---- per-thread A-row bookkeeping (fixed across the K loop) ----
  int a_iy[AR], a_ix[AR], a_pix[AR], a_b[AR];
  const int a_chunk = tid & 7;
#pragma unroll
  for (int i = 0; i < AR; ++i) {
    const int row = (tid >> 3) + 32 * i;
    const int r = row0 + row;
    if (r < p.rows_per_group) {
      const int R = group * p.rows_per_group + r;
      const int b = R / HoWo;
      const int rem = R - b * HoWo;
      const int oy = rem / p.Wo;
      const int ox = rem - oy * p.Wo;
      a_iy[i] = oy * p.stride + dyp;
      a_ix[i] = ox * p.stride + dxp;
      a_pix[i] = b * p.Hs * p.Ws;
      a_b[i] = b;
    } else {
      a_iy[i] = -(1 << 28);
      a_ix[i] = -(1 << 28);
      a_pix[i] = 0;
      a_b[i] = 0;
    }
  }

  auto load_chunk = [&](int q, int stage) {
    int s = 0, qq = q;
    while (s < p.nseg - 1 && qq >= p.seg[s].nchunk) {
      qq -= p.seg[s].nchunk;
      ++s;
    }
    const ConvSeg& sg = p.seg[s];
    const int cpt = sg.C >> 6;
    const int tap = qq / cpt;
    const int cc = qq - tap * cpt;
    const int ty = tap / sg.kw;
    const int tx = tap - ty * sg.kw;
    const int dy = ty + sg.dy0, dx = tx + sg.dx0;
    const uint32_t sA = smem_base + stage * STAGE_BYTES;
    const uint32_t sB = sA + A_BYTES;
#pragma unroll
    for (int i = 0; i < AR; ++i) {
      const int row = (tid >> 3) + 32 * i;
      const int iy = a_iy[i] + dy, ix = a_ix[i] + dx;
      const bool ok = (unsigned)iy < (unsigned)p.Hs && (unsigned)ix < (unsigned)p.Ws;
      const __half* src = sg.src;
      if (ok) {
        if (sg.win8) src += ((size_t)(a_b[i] * p.Hs + iy) * (p.Ws + 8) + ix) * 8 + a_chunk * 8;
        else src += (size_t)(a_pix[i] + iy * p.Ws + ix) * sg.C + cc * 64 + a_chunk * 8;
      }
      cp_async16(sA + swz128(row, a_chunk), src, ok ? 16 : 0);
    }
    const __half* wsrc = W + ((size_t)q * p.Ntot + n0) * 64 + a_chunk * 8;
#pragma unroll
    for (int i = 0; i < BR; ++i) {
      const int n = (tid >> 3) + 32 * i;
      cp_async16(sB + swz128(n, a_chunk), wsrc + (size_t)n * 64, 16);
    }
  };

  float acc[MT][NT][4];
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int j = 0; j < NT; ++j)
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[i][j][k] = 0.f;

  const int nq = p.total_chunks;
#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < nq) load_chunk(s, s);
    cp_async_commit();
  }
  for (int q = 0; q < nq; ++q) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    {
      const int qn = q + STAGES - 1;
      if (qn < nq) load_chunk(qn, qn % STAGES);
      cp_async_commit();
    }
    const uint32_t sA = smem_base + (q % STAGES) * STAGE_BYTES;
    const uint32_t sB = sA + A_BYTES;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      uint32_t af[MT][4];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        const int row = wm * (BM / 2) + mt * 16 + (lane & 15);
        ldmatrix_x4(af[mt], sA + swz128(row, ks * 2 + (lane >> 4)));
      }
#pragma unroll
      for (int np = 0; np < NT / 2; ++np) {
        const int n = wn * (BN / WN) + np * 16 + (lane & 7) + ((lane >> 4) << 3);
        uint32_t bf[4];
        ldmatrix_x4(bf, sB + swz128(n, ks * 2 + ((lane >> 3) & 1)));
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          mma_16816(acc[mt][2 * np], af[mt], bf[0], bf[1]);
          mma_16816(acc[mt][2 * np + 1], af[mt], bf[2], bf[3]);
        }
      }
    }
  }
  cp_async_wait<0>();

  // ---------------------------------- epilogue ----------------------------------
  // fragment -> (row, col): row = wm*BM/2 + mt*16 + (lane>>2) (+8 for regs 2,3); col = wn*BN/4 + nt*8 + (lane&3)*2
  long long out_pix[MT][2];
  int img[MT][2];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int row = wm * (BM / 2) + mt * 16 + (lane >> 2) + h * 8;
      const int r = row0 + row;
      if (r < p.rows_per_group) {
        const int R = group * p.rows_per_group + r;
        const int b = R / HoWo;
        const int rem = R - b * HoWo;
        const int oy = rem / p.Wo;
        const int ox = rem - oy * p.Wo;
        out_pix[mt][h] = ((long long)b * p.out_H + oy * p.out_sy + py) * p.out_W + ox * p.out_sx + px;
        img[mt][h] = b;
      } else {
        out_pix[mt][h] = -1;
        img[mt][h] = 0;
      }
    }
  const int colbase = n0 + wn * (BN / WN) + (lane & 3) * 2;

  if (EPI == EPI_BIAS || EPI == EPI_AFFINE) {
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (out_pix[mt][h] < 0) continue;
        float mean = 0.f, rstd = 1.f;
        if (EPI == EPI_AFFINE) {
          const float2 st = p.stats_in[out_pix[mt][h]];
          mean = st.x;
          rstd = st.y;
        }
        const size_t base = (size_t)out_pix[mt][h] * p.Ntot;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          const int col = colbase + nt * 8;
          float v0 = acc[mt][nt][2 * h], v1 = acc[mt][nt][2 * h + 1];
          if (EPI == EPI_AFFINE) {
            const float* u = p.aff_u + (size_t)group * p.aff_group_stride;
            const float* c = p.aff_c + (size_t)group * p.aff_group_stride;
            v0 = rstd * (v0 - mean * u[col]) + c[col];
            v1 = rstd * (v1 - mean * u[col + 1]) + c[col + 1];
          } else if (p.bias) {
            v0 += p.bias[col];
            v1 += p.bias[col + 1];
          }
          if (p.res) {
            const float2 rr = load_res2(p, out_pix[mt][h], col);
            v0 += rr.x;
            v1 += rr.y;
          }
          store_out2(p, base + col, v0, v1);
        }
      }
    return;
  }

  // ---- LayerNorm epilogues: BN == Ntot, the CTA owns whole pixel rows ----
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    const int col = colbase + nt * 8;
    const float b0 = p.bias[col], b1 = p.bias[col + 1];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      acc[mt][nt][0] += b0;
      acc[mt][nt][1] += b1;
      acc[mt][nt][2] += b0;
      acc[mt][nt][3] += b1;
    }
  }
  float mean[MT][2], rstd[MT][2];
  // pass 1: mean
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float s = 0.f;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) s += acc[mt][nt][2 * h] + acc[mt][nt][2 * h + 1];
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      if ((lane & 3) == 0) red_a[wn][wm * (BM / 2) + mt * 16 + (lane >> 2) + h * 8] = s;
    }
  __syncthreads();
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int row = wm * (BM / 2) + mt * 16 + (lane >> 2) + h * 8;
      mean[mt][h] = (red_a[0][row] + red_a[1][row] + red_a[2][row] + red_a[3][row]) * (1.f / BN);
    }
  // pass 2: biased variance around the mean
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float s = 0.f;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const float d0 = acc[mt][nt][2 * h] - mean[mt][h], d1 = acc[mt][nt][2 * h + 1] - mean[mt][h];
        s += d0 * d0 + d1 * d1;
      }
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      if ((lane & 3) == 0) red_b[wn][wm * (BM / 2) + mt * 16 + (lane >> 2) + h * 8] = s;
    }
  __syncthreads();
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int row = wm * (BM / 2) + mt * 16 + (lane >> 2) + h * 8;
      const float var = (red_b[0][row] + red_b[1][row] + red_b[2][row] + red_b[3][row]) * (1.f / BN);
      rstd[mt][h] = 1.f / sqrtf(var + 1e-5f);
    }
  __syncthreads();  // red_a / red_b are reused below for the output statistics

  float osum[MT][2], osq[MT][2];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      osum[mt][h] = 0.f;
      osq[mt][h] = 0.f;
      const bool valid = out_pix[mt][h] >= 0;
      const size_t base = valid ? (size_t)out_pix[mt][h] * p.Ntot : 0;
      const float* shift = (EPI == EPI_LN_SHIFT && p.shift) ? p.shift + (size_t)img[mt][h] * p.shift_stride : nullptr;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const int col = colbase + nt * 8;
        float v0 = (acc[mt][nt][2 * h] - mean[mt][h]) * rstd[mt][h] * p.ln_g[col] + p.ln_b[col];
        float v1 = (acc[mt][nt][2 * h + 1] - mean[mt][h]) * rstd[mt][h] * p.ln_g[col + 1] + p.ln_b[col + 1];
        v0 = fmaxf(v0, 0.f);
        v1 = fmaxf(v1, 0.f);
        if (EPI == EPI_LN_SHIFT) {
          if (shift) {
            v0 += shift[col];
            v1 += shift[col + 1];
          }
        } else if (valid && p.res) {
          const float2 rr = load_res2(p, out_pix[mt][h], col);
          v0 += rr.x;
          v1 += rr.y;
        }
        const uint32_t hv = valid ? store_out2(p, base + col, v0, v1) : pack_half2(v0, v1);
        if (EPI == EPI_LN_RES) {
          const float2 q = unpack_half2(hv);  // statistics of what the consumer will read
          osum[mt][h] += q.x + q.y;
          osq[mt][h] += q.x * q.x + q.y * q.y;
        }
      }
    }
  if (EPI == EPI_LN_RES && p.stats_out) {
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float s = osum[mt][h], q = osq[mt][h];
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        q += __shfl_xor_sync(0xffffffffu, q, 1);
        q += __shfl_xor_sync(0xffffffffu, q, 2);
        if ((lane & 3) == 0) {
          const int row = wm * (BM / 2) + mt * 16 + (lane >> 2) + h * 8;
          red_a[wn][row] = s;
          red_b[wn][row] = q;
        }
      }
    __syncthreads();
    if (tid < BM) {
      const int r = row0 + tid;
      if (r < p.rows_per_group) {
        const float s = red_a[0][tid] + red_a[1][tid] + red_a[2][tid] + red_a[3][tid];
        const float q = red_b[0][tid] + red_b[1][tid] + red_b[2][tid] + red_b[3][tid];
        const float m = s * (1.f / BN);
        const float var = fmaxf(q * (1.f / BN) - m * m, 0.f);
        // stride-1 LayerNorm blocks: output pixel index == global row index
        p.stats_out[(size_t)group * p.rows_per_group + r] = make_float2(m, 1.f / sqrtf(var + 1e-5f));
      }
    }
  }
}

}  // namespace cdc
