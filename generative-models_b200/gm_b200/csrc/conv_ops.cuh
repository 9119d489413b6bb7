// Building blocks of the DCGAN conv path (BASELINE configs[4]; the reference only recommends DCGAN, README.md:68,96 —
// there is no reference implementation): NHWC bf16 activations as row-major matrices [B*H*W, C], so
// that every convolution / transposed convolution is one tcgen05 GEMM (gemm_umma.cuh) between an im2col / col2im pass:
//   conv   (k4 s2 p1): col = im2col(x) [B*Ho*Wo, 16*Cin];  y = col W^T        (W [Cout, (kh,kw,ci)])
//   convT  (k4 s2 p1): col = x Wm^T    [B*Hi*Wi, 16*Cout]; y = col2im(col)    (Wm [(kh,kw,co), Cin])
// and their gradients are the same two data movements with the roles swapped.  BatchNorm (training mode, batch
// statistics) and the activations are column-statistics + elementwise kernels over the same matrices.  All HBM-bound:
// 16-byte accesses, grids sized in multiples of the SM count, deterministic two-stage reductions.
#pragma once
#include "ptx.cuh"

namespace gm {

// ---------------------------------------------------------------- im2col / col2im (kernel 4, stride 2, pad 1)
// Both are pure HBM-bound data movements (the column matrix is 4x the activation it comes from, up to 537 MB at B = 1024),
// so the kernels are written for memory-level parallelism and cheap index arithmetic (the first versions ran at ~40 % of
// the HBM rate: one 16-byte access in flight per thread behind a chain of 64-bit divisions, profiles/r2i_timeline_dcgan.md):
// 32-bit indices (the host checks the extents), divisions by the grid extents as shifts when they are powers of two, kUnroll
// independent items per thread with all loads issued before the first store.
struct FastDiv {            // x / d and x % d for 32-bit x; shift >= 0 when d is a power of two
  uint32_t d;
  int shift;
};
__host__ __device__ inline FastDiv make_fastdiv(uint32_t d) {
  FastDiv f;
  f.d = d;
  f.shift = -1;
  for (int s = 0; s < 32; ++s)
    if ((1u << s) == d) f.shift = s;
  return f;
}
__device__ __forceinline__ uint32_t fdiv(uint32_t x, const FastDiv f) { return f.shift >= 0 ? (x >> f.shift) : (x / f.d); }
__device__ __forceinline__ void fdivmod(uint32_t x, const FastDiv f, uint32_t& q, uint32_t& r) {
  q = fdiv(x, f);
  r = x - q * f.d;
}
constexpr int kConvUnroll = 4;

// x [B, H, W, C] (row pitch ldx elements per pixel) -> col [(b, ho, wo), (kh, kw, c)] with Ho = H/2, Wo = W/2.
// VEC: C % 8 == 0, one item = (output pixel, tap, 8-channel group) = one 16-byte copy (consecutive items are consecutive
// 16-byte chunks of the column matrix: the dominant stream, the writes, is perfectly linear); else one item = (output pixel,
// tap) with a scalar channel loop (the 3-channel image layer).
template <bool VEC>
__global__ void __launch_bounds__(256) im2col_k4s2_kernel(const __nv_bfloat16* __restrict__ x, int H, int W, int C, int ldx,
                                                          __nv_bfloat16* __restrict__ col, int ldc, uint32_t total, FastDiv dcg,
                                                          FastDiv dwo, FastDiv dho) {
  griddep_sync();
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t t0 = blockIdx.x * blockDim.x + threadIdx.x; t0 < total; t0 += stride * kConvUnroll) {
    const __nv_bfloat16* src[kConvUnroll];
    __nv_bfloat16* dst[kConvUnroll];
    bool ok[kConvUnroll], live[kConvUnroll];
#pragma unroll
    for (int u = 0; u < kConvUnroll; ++u) {
      const uint32_t t = t0 + u * stride;
      live[u] = t < total && t >= t0;                 // t >= t0: no wrap-around past 2^32
      uint32_t r, g = 0;
      if (VEC) fdivmod(t, dcg, r, g); else r = t;
      const uint32_t tap = r & 15u;
      r >>= 4;                                        // output pixel index (b, ho, wo)
      uint32_t q, wo, b, ho;
      fdivmod(r, dwo, q, wo);
      fdivmod(q, dho, b, ho);
      const int kh = int(tap >> 2), kw = int(tap & 3u);
      const int iy = 2 * int(ho) - 1 + kh, ix = 2 * int(wo) - 1 + kw;
      ok[u] = live[u] && iy >= 0 && iy < H && ix >= 0 && ix < W;
      dst[u] = col + size_t(r) * ldc + tap * C + g * 8;
      src[u] = x + ((size_t(b) * H + iy) * W + ix) * ldx + g * 8;
    }
    if (VEC) {
      uint4 v[kConvUnroll];
#pragma unroll
      for (int u = 0; u < kConvUnroll; ++u) {
        v[u] = make_uint4(0, 0, 0, 0);
        if (ok[u]) v[u] = __ldg(reinterpret_cast<const uint4*>(src[u]));
      }
#pragma unroll
      for (int u = 0; u < kConvUnroll; ++u)
        if (live[u]) *reinterpret_cast<uint4*>(dst[u]) = v[u];
    } else {
#pragma unroll
      for (int u = 0; u < kConvUnroll; ++u)
        if (live[u])
          for (int c = 0; c < C; ++c) dst[u][c] = ok[u] ? src[u][c] : __float2bfloat16_rn(0.f);
    }
  }
}

// col [(b, iy, ix), (kh, kw, c)] over an Hi x Wi grid -> y [B, 2Hi, 2Wi, C] (gather form: every output pixel sums the
// <= 4 taps that reach it; deterministic, no atomics).  Fused tail, by `mode`:
//   0: y = sum                       1: y = sigmoid(sum)                  (generator output, src/ns_gan.py:45-46)
//   2: y = sum * lrelu'(aux)         3: y = sum * aux (1 - aux)           (gradients through a LeakyReLU / sigmoid output)
// VEC: one item = (output pixel, 8-channel group): four predicated 16-byte loads (+ one of aux) in flight, 2 items per thread.
enum : int { C2I_NONE = 0, C2I_SIGMOID = 1, C2I_LRELU_GRAD = 2, C2I_SIGMOID_GRAD = 3 };
template <bool VEC>
__global__ void __launch_bounds__(256) col2im_k4s2_kernel(const __nv_bfloat16* __restrict__ col, int ldc, int Hi, int Wi, int C,
                                                          __nv_bfloat16* __restrict__ y, int ldy, int mode,
                                                          const __nv_bfloat16* __restrict__ aux, int ld_aux, float slope, uint32_t total,
                                                          FastDiv dcg, FastDiv dwo, FastDiv dho) {
  griddep_sync();
  constexpr int U = 2;
  const uint32_t stride = gridDim.x * blockDim.x;
  if constexpr (VEC) {
    for (uint32_t t0 = blockIdx.x * blockDim.x + threadIdx.x; t0 < total; t0 += stride * U) {
      uint4 v[U][4], av[U];
      uint32_t pix[U], g[U];
      bool live[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const uint32_t t = t0 + u * stride;
        live[u] = t < total && t >= t0;
        fdivmod(t, dcg, pix[u], g[u]);
        uint32_t q, ox, b, oy;
        fdivmod(pix[u], dwo, q, ox);
        fdivmod(q, dho, b, oy);
        // taps with (oy + 1 - kh) even and in range: kh in {(oy + 1) & 1, ((oy + 1) & 1) + 2}
#pragma unroll
        for (int a = 0; a < 2; ++a) {
          const int kh = int((oy + 1) & 1u) + 2 * a, iy = (int(oy) + 1 - kh) >> 1;
#pragma unroll
          for (int bb = 0; bb < 2; ++bb) {
            const int kw = int((ox + 1) & 1u) + 2 * bb, ix = (int(ox) + 1 - kw) >> 1;
            const bool ok = live[u] && iy >= 0 && iy < Hi && ix >= 0 && ix < Wi;
            uint4 w = make_uint4(0, 0, 0, 0);
            if (ok) w = __ldg(reinterpret_cast<const uint4*>(col + ((size_t(b) * Hi + iy) * Wi + ix) * ldc + (kh * 4 + kw) * C + g[u] * 8));
            v[u][a * 2 + bb] = w;
          }
        }
        av[u] = make_uint4(0, 0, 0, 0);
        if (mode >= C2I_LRELU_GRAD && live[u]) av[u] = __ldg(reinterpret_cast<const uint4*>(aux + size_t(pix[u]) * ld_aux + g[u] * 8));
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (!live[u]) continue;
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {                    // summation order: (kh, kw) ascending
          const uint32_t w[4] = {v[u][k].x, v[u][k].y, v[u][k].z, v[u][k].w};
#pragma unroll
          for (int q = 0; q < 4; ++q) { acc[2 * q] += bf16_lo(w[q]); acc[2 * q + 1] += bf16_hi(w[q]); }
        }
        const uint32_t aw[4] = {av[u].x, av[u].y, av[u].z, av[u].w};
        float o[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          float val = acc[c];
          const float a = (c & 1) ? bf16_hi(aw[c >> 1]) : bf16_lo(aw[c >> 1]);
          if (mode == C2I_SIGMOID) val = 1.f / (1.f + __expf(-val));
          else if (mode == C2I_LRELU_GRAD) val = a > 0.f ? val : slope * val;
          else if (mode == C2I_SIGMOID_GRAD) val = val * a * (1.f - a);
          o[c] = val;
        }
        store_bf16x8(y + size_t(pix[u]) * ldy + g[u] * 8, o, 0);
      }
    }
  } else {
    // C < 8 (the 3-channel image layer): one thread per output pixel, scalar channels
    for (uint32_t pix = blockIdx.x * blockDim.x + threadIdx.x; pix < total; pix += stride) {
      uint32_t q, ox, b, oy;
      fdivmod(pix, dwo, q, ox);
      fdivmod(q, dho, b, oy);
      float acc[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        const int kh = int((oy + 1) & 1u) + 2 * a, iy = (int(oy) + 1 - kh) >> 1;
        if (iy < 0 || iy >= Hi) continue;
#pragma unroll
        for (int bb = 0; bb < 2; ++bb) {
          const int kw = int((ox + 1) & 1u) + 2 * bb, ix = (int(ox) + 1 - kw) >> 1;
          if (ix < 0 || ix >= Wi) continue;
          const __nv_bfloat16* src = col + ((size_t(b) * Hi + iy) * Wi + ix) * ldc + (kh * 4 + kw) * C;
#pragma unroll
          for (int c = 0; c < 8; ++c)
            if (c < C) acc[c] += __bfloat162float(src[c]);
        }
      }
      const __nv_bfloat16* ap = aux + size_t(pix) * ld_aux;
      __nv_bfloat16* dst = y + size_t(pix) * ldy;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        if (c < C) {
          float val = acc[c];
          if (mode == C2I_SIGMOID) val = 1.f / (1.f + __expf(-val));
          else if (mode >= C2I_LRELU_GRAD) {
            const float a = __bfloat162float(ap[c]);
            val = mode == C2I_LRELU_GRAD ? (a > 0.f ? val : slope * val) : val * a * (1.f - a);
          }
          dst[c] = __float2bfloat16_rn(val);
        }
      }
    }
  }
}

// ---------------------------------------------------------------- BatchNorm2d, training mode (batch statistics)
// x [rows, C] bf16 (rows = B*H*W).  Pass 1: per-block partial column sums (sum, sum of squares; fp32 per thread over a
// slab of rows, double across the block) -> part [nblk][2][C].  Pass 2 (a warp per channel): mean, invstd (+ running
// statistics, momentum 0.1, unbiased variance like torch).  Pass 3: y = act(gamma * (x - mean) * invstd + beta).
// act: 0 none, 1 ReLU, 2 LeakyReLU(slope).  Thread mapping of every pass: thread = (8-channel group g, row lane rl), the
// block walks rpi = 256 / groups rows per iteration - no division in the loops, per-channel constants in registers, and
// kBnUnroll rows per thread in flight (HBM-bound: the first versions had one 16-byte load in flight per thread and re-read
// gamma / beta / mean / invstd from global memory for every element, 25-45 % of the HBM rate).
constexpr int kBnThreads = 256;
constexpr int kBnUnroll = 4;
constexpr int kBnStatUnroll = 8;
// The double-precision running sums of a thread live in ITS slots of the block's shared array (the same array the final
// cross-row reduction reads), not in registers: 32 fewer registers per thread = more resident warps with loads in flight.
__global__ void __launch_bounds__(kBnThreads) bn_partial_kernel(const __nv_bfloat16* __restrict__ x, long long rows, int C, int ld,
                                                                 double* __restrict__ part) {
  griddep_sync();
  extern __shared__ double bn_sh[];                  // [rows_per_iter][2][C]
  const int groups = C / 8;
  const int g = threadIdx.x % groups, rl = threadIdx.x / groups, rpi = kBnThreads / groups;
  if (rl < rpi) {
    double* const my1 = bn_sh + (rl * 2 + 0) * C + g * 8;
    double* const my2 = bn_sh + (rl * 2 + 1) * C + g * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) { my1[j] = 0.0; my2[j] = 0.0; }
    float s1[8], s2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { s1[j] = 0.f; s2[j] = 0.f; }
    int since = 0;
    const long long rstep = (long long)gridDim.x * rpi;
    for (long long r0 = (long long)blockIdx.x * rpi + rl; r0 < rows; r0 += rstep * kBnStatUnroll) {
      uint4 v[kBnStatUnroll];
#pragma unroll
      for (int u = 0; u < kBnStatUnroll; ++u) {
        const long long r = r0 + u * rstep;
        v[u] = make_uint4(0, 0, 0, 0);                // zero rows add nothing to either sum
        if (r < rows) v[u] = __ldg(reinterpret_cast<const uint4*>(x + r * ld) + g);
      }
#pragma unroll
      for (int u = 0; u < kBnStatUnroll; ++u) {
        const uint32_t w[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float a = bf16_lo(w[q]), b = bf16_hi(w[q]);
          s1[2 * q] += a; s2[2 * q] = fmaf(a, a, s2[2 * q]);
          s1[2 * q + 1] += b; s2[2 * q + 1] = fmaf(b, b, s2[2 * q + 1]);
        }
      }
      if (++since == 64 / kBnStatUnroll) {   // flush the fp32 running sums into doubles every 64 rows (B*H*W up to 8 M rows)
#pragma unroll
        for (int j = 0; j < 8; ++j) { my1[j] += s1[j]; my2[j] += s2[j]; s1[j] = 0.f; s2[j] = 0.f; }
        since = 0;
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) { my1[j] += s1[j]; my2[j] += s2[j]; }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * C; c += kBnThreads) {
    const int which = c / C, ch = c % C;
    double t = 0.0;
    for (int i = 0; i < rpi; ++i) t += bn_sh[(i * 2 + which) * C + ch];
    part[((long long)blockIdx.x * 2 + which) * C + ch] = t;
  }
}
// stats[0][C] = mean, stats[1][C] = invstd; running[0] mean, running[1] var (nullable).  One WARP per channel: the lanes
// stride over the per-block partials (a thread per channel walked all ~600 partials serially: 67 us for a 2 KB result).
__global__ void bn_finalize_kernel(const double* __restrict__ part, int nblk, int C, double count, float eps,
                                   float* __restrict__ stats, float* __restrict__ running, float momentum) {
  griddep_sync();
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (c >= C) return;
  double s1 = 0.0, s2 = 0.0;
  for (int i0 = lane; i0 < nblk; i0 += 32 * 8) {          // 16 loads in flight per lane (was a serial chain of ~19 round trips)
    double a1[8], a2[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = i0 + 32 * u;
      a1[u] = i < nblk ? part[((long long)i * 2) * C + c] : 0.0;
      a2[u] = i < nblk ? part[((long long)i * 2 + 1) * C + c] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) { s1 += a1[u]; s2 += a2[u]; }
  }
  for (int o = 16; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
  if (lane != 0) return;
  const double mean = s1 / count;
  const double var = fmax(s2 / count - mean * mean, 0.0);       // biased, as torch normalises with
  stats[c] = float(mean);
  stats[C + c] = float(1.0 / sqrt(var + double(eps)));
  if (running) {
    running[c] = (1.f - momentum) * running[c] + momentum * float(mean);
    running[C + c] = (1.f - momentum) * running[C + c] + momentum * float(var * count / fmax(count - 1.0, 1.0));
  }
}
// y = act(x * sc + sh) with sc = gamma invstd, sh = beta - mean sc (two registers per channel)
__global__ void __launch_bounds__(kBnThreads) bn_apply_kernel(const __nv_bfloat16* __restrict__ x, long long rows, int C, int ld,
                                                               const float* __restrict__ stats, const float* __restrict__ gamma,
                                                               const float* __restrict__ beta, int act, float slope,
                                                               __nv_bfloat16* __restrict__ y, int ldy) {
  griddep_sync();
  const int groups = C / 8;
  const int g = threadIdx.x % groups, rl = threadIdx.x / groups, rpi = kBnThreads / groups;
  if (rl >= rpi) return;
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = g * 8 + j;
    sc[j] = gamma[c] * stats[C + c];
    sh[j] = fmaf(-stats[c], sc[j], beta[c]);
  }
  const long long rstep = (long long)gridDim.x * rpi;
  for (long long r0 = (long long)blockIdx.x * rpi + rl; r0 < rows; r0 += rstep * kBnUnroll) {
    uint4 v[kBnUnroll];
#pragma unroll
    for (int u = 0; u < kBnUnroll; ++u) {
      const long long r = r0 + u * rstep;
      if (r < rows) v[u] = __ldg(reinterpret_cast<const uint4*>(x + r * ld) + g);
    }
#pragma unroll
    for (int u = 0; u < kBnUnroll; ++u) {
      const long long r = r0 + u * rstep;
      if (r >= rows) break;
      const uint32_t w[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float xv = (j & 1) ? bf16_hi(w[j >> 1]) : bf16_lo(w[j >> 1]);
        float t = fmaf(xv, sc[j], sh[j]);
        if (act == 1) t = fmaxf(t, 0.f);
        else if (act == 2) t = t > 0.f ? t : slope * t;
        o[j] = t;
      }
      store_bf16x8(y + r * ldy + g * 8, o, 0);
    }
  }
}

// backward: g = dy * act'(gamma xhat + beta); dbeta = sum g; dgamma = sum g xhat;
//           dx = gamma invstd / N * (N g - dbeta - xhat dgamma)            (torch.nn.functional.batch_norm backward)
// Pass 1: partial sums (dbeta, dgamma) with the bn_partial layout; pass 2: finalise into dgb [2][C] fp32; pass 3: dx.
constexpr int kBnBwdUnroll = 4;                       // two operands per row: 8 loads in flight
__global__ void __launch_bounds__(kBnThreads, 2) bn_bwd_partial_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
                                                                        long long rows, int C, int ld, const float* __restrict__ stats,
                                                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                        int act, float slope, double* __restrict__ part) {
  griddep_sync();
  extern __shared__ double bn_sh[];
  const int groups = C / 8;
  const int g = threadIdx.x % groups, rl = threadIdx.x / groups, rpi = kBnThreads / groups;
  if (rl < rpi) {
    float mu[8], is[8], ga[8], be[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { const int c = g * 8 + j; mu[j] = stats[c]; is[j] = stats[C + c]; ga[j] = gamma[c]; be[j] = beta[c]; }
    double* const my1 = bn_sh + (rl * 2 + 0) * C + g * 8;
    double* const my2 = bn_sh + (rl * 2 + 1) * C + g * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) { my1[j] = 0.0; my2[j] = 0.0; }
    float s1[8], s2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { s1[j] = 0.f; s2[j] = 0.f; }
    int since = 0;
    const long long rstep = (long long)gridDim.x * rpi;
    for (long long r0 = (long long)blockIdx.x * rpi + rl; r0 < rows; r0 += rstep * kBnBwdUnroll) {
      uint4 vx[kBnBwdUnroll], vg[kBnBwdUnroll];
#pragma unroll
      for (int u = 0; u < kBnBwdUnroll; ++u) {
        const long long r = r0 + u * rstep;
        vx[u] = make_uint4(0, 0, 0, 0); vg[u] = make_uint4(0, 0, 0, 0);   // dy = 0 rows add nothing
        if (r < rows) {
          vx[u] = __ldg(reinterpret_cast<const uint4*>(x + r * ld) + g);
          vg[u] = __ldg(reinterpret_cast<const uint4*>(dy + r * ld) + g);
        }
      }
#pragma unroll
      for (int u = 0; u < kBnBwdUnroll; ++u) {
        const uint32_t wx[4] = {vx[u].x, vx[u].y, vx[u].z, vx[u].w}, wg[4] = {vg[u].x, vg[u].y, vg[u].z, vg[u].w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float xv = (j & 1) ? bf16_hi(wx[j >> 1]) : bf16_lo(wx[j >> 1]);
          float gg = (j & 1) ? bf16_hi(wg[j >> 1]) : bf16_lo(wg[j >> 1]);
          const float xh = (xv - mu[j]) * is[j];
          const float pre = ga[j] * xh + be[j];
          if (act == 1) gg = pre > 0.f ? gg : 0.f;
          else if (act == 2) gg = pre > 0.f ? gg : slope * gg;
          s1[j] += gg;
          s2[j] = fmaf(gg, xh, s2[j]);
        }
      }
      if (++since == 64 / kBnBwdUnroll) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { my1[j] += s1[j]; my2[j] += s2[j]; s1[j] = 0.f; s2[j] = 0.f; }
        since = 0;
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) { my1[j] += s1[j]; my2[j] += s2[j]; }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * C; c += kBnThreads) {
    const int which = c / C, ch = c % C;
    double t = 0.0;
    for (int i = 0; i < rpi; ++i) t += bn_sh[(i * 2 + which) * C + ch];
    part[((long long)blockIdx.x * 2 + which) * C + ch] = t;
  }
}
// dgb[0][C] = dbeta, dgb[1][C] = dgamma; one warp per output value
__global__ void bn_bwd_finalize_kernel(const double* __restrict__ part, int nblk, int C, float* __restrict__ dgb) {
  griddep_sync();
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (c >= 2 * C) return;
  const int which = c / C, ch = c % C;
  double t = 0.0;
  for (int i0 = lane; i0 < nblk; i0 += 32 * 8) {
    double a[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = i0 + 32 * u;
      a[u] = i < nblk ? part[((long long)i * 2 + which) * C + ch] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) t += a[u];
  }
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  if (lane == 0) dgb[c] = float(t);
}
__global__ void __launch_bounds__(kBnThreads, 2) bn_bwd_apply_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
                                                                   long long rows, int C, int ld, const float* __restrict__ stats,
                                                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                   int act, float slope, const float* __restrict__ dgb, float inv_count,
                                                                   __nv_bfloat16* __restrict__ dx, int lddx) {
  griddep_sync();
  const int groups = C / 8;
  const int g = threadIdx.x % groups, rl = threadIdx.x / groups, rpi = kBnThreads / groups;
  if (rl >= rpi) return;
  float mu[8], is[8], ga[8], be[8], db[8], dg[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = g * 8 + j;
    mu[j] = stats[c]; is[j] = stats[C + c]; ga[j] = gamma[c]; be[j] = beta[c]; db[j] = dgb[c]; dg[j] = dgb[C + c];
  }
  const long long rstep = (long long)gridDim.x * rpi;
  for (long long r0 = (long long)blockIdx.x * rpi + rl; r0 < rows; r0 += rstep * kBnBwdUnroll) {
    uint4 vx[kBnBwdUnroll], vg[kBnBwdUnroll];
#pragma unroll
    for (int u = 0; u < kBnBwdUnroll; ++u) {
      const long long r = r0 + u * rstep;
      if (r < rows) {
        vx[u] = __ldg(reinterpret_cast<const uint4*>(x + r * ld) + g);
        vg[u] = __ldg(reinterpret_cast<const uint4*>(dy + r * ld) + g);
      }
    }
#pragma unroll
    for (int u = 0; u < kBnBwdUnroll; ++u) {
      const long long r = r0 + u * rstep;
      if (r >= rows) break;
      const uint32_t wx[4] = {vx[u].x, vx[u].y, vx[u].z, vx[u].w}, wg[4] = {vg[u].x, vg[u].y, vg[u].z, vg[u].w};
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float xv = (j & 1) ? bf16_hi(wx[j >> 1]) : bf16_lo(wx[j >> 1]);
        float gg = (j & 1) ? bf16_hi(wg[j >> 1]) : bf16_lo(wg[j >> 1]);
        const float xh = (xv - mu[j]) * is[j];
        const float pre = ga[j] * xh + be[j];
        if (act == 1) gg = pre > 0.f ? gg : 0.f;
        else if (act == 2) gg = pre > 0.f ? gg : slope * gg;
        o[j] = ga[j] * is[j] * (gg - inv_count * (db[j] + xh * dg[j]));
      }
      store_bf16x8(dx + r * lddx + g * 8, o, 0);
    }
  }
}

// ---------------------------------------------------------------- small helpers
// fp32 [R, C] -> bf16 copy [R, ld] and (optionally) its transpose [C, ld_t] (GEMM operand forms of a weight matrix)
__global__ void cast_bf16_kernel(const float* __restrict__ src, int R, int C, __nv_bfloat16* __restrict__ dst, int ld,
                                 __nv_bfloat16* __restrict__ dst_t, int ld_t) {
  griddep_sync();
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= (long long)R * C) return;
  const int r = int(i / C), c = int(i % C);
  const __nv_bfloat16 b = __float2bfloat16_rn(src[i]);
  if (dst) dst[(long long)r * ld + c] = b;
  if (dst_t) dst_t[(long long)c * ld_t + r] = b;
}
// out [rows, ld] bf16: column 0 = v[r], the rest 0 (the upstream gradient of a 1-channel output as a GEMM operand)
__global__ void pack_col0_kernel(const float* __restrict__ v, int rows, __nv_bfloat16* __restrict__ out, int ld) {
  griddep_sync();
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= (long long)rows * ld) return;
  out[i] = __float2bfloat16_rn((i % ld) == 0 ? v[i / ld] : 0.f);
}

}  // namespace gm
