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
// x [B, H, W, C] (row pitch ldx elements per pixel) -> col [(b, ho, wo), (kh, kw, c)] with Ho = H/2, Wo = W/2.
// One thread per (output pixel, tap, 8-channel group) when C % 8 == 0, else per (output pixel, tap) with a scalar loop.
__global__ void im2col_k4s2_kernel(const __nv_bfloat16* __restrict__ x, int B, int H, int W, int C, int ldx,
                                   __nv_bfloat16* __restrict__ col, int ldc) {
  griddep_sync();
  const int Ho = H >> 1, Wo = W >> 1;
  const int cg = (C % 8 == 0) ? C / 8 : 1;
  const long long total = (long long)B * Ho * Wo * 16 * cg;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int g = int(t % cg);
    long long r = t / cg;
    const int tap = int(r % 16);
    r /= 16;                                        // output pixel index (b, ho, wo)
    const int wo = int(r % Wo), ho = int((r / Wo) % Ho), b = int(r / ((long long)Wo * Ho));
    const int kh = tap >> 2, kw = tap & 3;
    const int iy = 2 * ho - 1 + kh, ix = 2 * wo - 1 + kw;
    const bool ok = iy >= 0 && iy < H && ix >= 0 && ix < W;
    __nv_bfloat16* dst = col + r * ldc + tap * C;
    const __nv_bfloat16* src = x + (((long long)b * H + iy) * W + ix) * ldx;
    if (C % 8 == 0) {
      uint4 v = make_uint4(0, 0, 0, 0);
      if (ok) v = __ldg(reinterpret_cast<const uint4*>(src) + g);
      reinterpret_cast<uint4*>(dst)[g] = v;
    } else {
      for (int c = 0; c < C; ++c) dst[c] = ok ? src[c] : __float2bfloat16_rn(0.f);
    }
  }
}

// col [(b, iy, ix), (kh, kw, c)] over an Hi x Wi grid -> y [B, 2Hi, 2Wi, C] (gather form: every output pixel sums the
// <= 4 taps that reach it; deterministic, no atomics).  Fused tail, by `mode`:
//   0: y = sum                       1: y = sigmoid(sum)                  (generator output, src/ns_gan.py:45-46)
//   2: y = sum * lrelu'(aux)         3: y = sum * aux (1 - aux)           (gradients through a LeakyReLU / sigmoid output)
enum : int { C2I_NONE = 0, C2I_SIGMOID = 1, C2I_LRELU_GRAD = 2, C2I_SIGMOID_GRAD = 3 };
__global__ void col2im_k4s2_kernel(const __nv_bfloat16* __restrict__ col, int ldc, int B, int Hi, int Wi, int C,
                                   __nv_bfloat16* __restrict__ y, int ldy, int mode, const __nv_bfloat16* __restrict__ aux,
                                   int ld_aux, float slope) {
  griddep_sync();
  const int Ho = 2 * Hi, Wo = 2 * Wi;
  const int cg = (C % 8 == 0) ? C / 8 : 1, cw = (C % 8 == 0) ? 8 : C;
  const long long total = (long long)B * Ho * Wo * cg;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int g = int(t % cg);
    const long long pix = t / cg;
    const int ox = int(pix % Wo), oy = int((pix / Wo) % Ho), b = int(pix / ((long long)Wo * Ho));
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    // taps with (oy + 1 - kh) even and in range: kh in {(oy + 1) & 1, ((oy + 1) & 1) + 2}
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      const int kh = ((oy + 1) & 1) + 2 * a, iy = (oy + 1 - kh) >> 1;
      if (iy < 0 || iy >= Hi) continue;
#pragma unroll
      for (int bb = 0; bb < 2; ++bb) {
        const int kw = ((ox + 1) & 1) + 2 * bb, ix = (ox + 1 - kw) >> 1;
        if (ix < 0 || ix >= Wi) continue;
        const __nv_bfloat16* src = col + (((long long)b * Hi + iy) * Wi + ix) * ldc + (kh * 4 + kw) * C + g * 8;
        if (cw == 8) {
          const uint4 v = __ldg(reinterpret_cast<const uint4*>(src));
          const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int q = 0; q < 4; ++q) { acc[2 * q] += bf16_lo(u[q]); acc[2 * q + 1] += bf16_hi(u[q]); }
        } else {
          for (int c = 0; c < cw; ++c) acc[c] += __bfloat162float(src[c]);
        }
      }
    }
    float av[8];
    if (mode >= C2I_LRELU_GRAD) {
      const __nv_bfloat16* ap = aux + pix * ld_aux + g * 8;
      for (int c = 0; c < cw; ++c) av[c] = __bfloat162float(ap[c]);
    }
    __nv_bfloat16* dst = y + pix * ldy + g * 8;
    float o[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      float v = acc[c];
      if (c < cw) {
        if (mode == C2I_SIGMOID) v = 1.f / (1.f + __expf(-v));
        else if (mode == C2I_LRELU_GRAD) v = av[c] > 0.f ? v : slope * v;
        else if (mode == C2I_SIGMOID_GRAD) v = v * av[c] * (1.f - av[c]);
      }
      o[c] = v;
    }
    if (cw == 8) store_bf16x8(dst, o, 0);
    else for (int c = 0; c < cw; ++c) dst[c] = __float2bfloat16_rn(o[c]);
  }
}

// ---------------------------------------------------------------- BatchNorm2d, training mode (batch statistics)
// x [rows, C] bf16 (rows = B*H*W).  Pass 1: per-block partial column sums (sum, sum of squares; fp32 per thread over a
// slab of rows, double across the block) -> part [nblk][2][C].  Pass 2 (one block per 32 channels... one thread per
// channel): mean, invstd (+ running statistics, momentum 0.1, unbiased variance like torch).  Pass 3: y = act(gamma *
// (x - mean) * invstd + beta).  act: 0 none, 1 ReLU, 2 LeakyReLU(slope).
constexpr int kBnThreads = 256;
__global__ void __launch_bounds__(kBnThreads) bn_partial_kernel(const __nv_bfloat16* __restrict__ x, long long rows, int C, int ld,
                                                                 double* __restrict__ part) {
  griddep_sync();
  extern __shared__ double bn_sh[];                  // [rows_per_iter][2 * C]
  const int groups = C / 8;
  const int g = threadIdx.x % groups, rl = threadIdx.x / groups, rpi = kBnThreads / groups;
  float s1[8], s2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { s1[j] = 0.f; s2[j] = 0.f; }
  double d1[8], d2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { d1[j] = 0.0; d2[j] = 0.0; }
  int since = 0;
  if (rl < rpi) {
    for (long long r = (long long)blockIdx.x * rpi + rl; r < rows; r += (long long)gridDim.x * rpi) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(x + r * ld) + g);
      const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float a = bf16_lo(u[q]), b = bf16_hi(u[q]);
        s1[2 * q] += a; s2[2 * q] = fmaf(a, a, s2[2 * q]);
        s1[2 * q + 1] += b; s2[2 * q + 1] = fmaf(b, b, s2[2 * q + 1]);
      }
      if (++since == 64) {     // flush the fp32 running sums into doubles (long columns: B*H*W up to 8 M rows)
#pragma unroll
        for (int j = 0; j < 8; ++j) { d1[j] += s1[j]; d2[j] += s2[j]; s1[j] = 0.f; s2[j] = 0.f; }
        since = 0;
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) { d1[j] += s1[j]; d2[j] += s2[j]; }
#pragma unroll
    for (int j = 0; j < 8; ++j) { bn_sh[(rl * 2 + 0) * C + g * 8 + j] = d1[j]; bn_sh[(rl * 2 + 1) * C + g * 8 + j] = d2[j]; }
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
  for (int i = lane; i < nblk; i += 32) { s1 += part[((long long)i * 2) * C + c]; s2 += part[((long long)i * 2 + 1) * C + c]; }
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
__global__ void bn_apply_kernel(const __nv_bfloat16* __restrict__ x, long long rows, int C, int ld, const float* __restrict__ stats,
                                const float* __restrict__ gamma, const float* __restrict__ beta, int act, float slope,
                                __nv_bfloat16* __restrict__ y, int ldy) {
  griddep_sync();
  const int groups = C / 8;
  const long long total = rows * groups;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int g = int(t % groups);
    const long long r = t / groups;
    float v[8];
    load_bf16x8(x + r * ld + g * 8, v, 0);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = g * 8 + j;
      float o = gamma[c] * (v[j] - stats[c]) * stats[C + c] + beta[c];
      if (act == 1) o = fmaxf(o, 0.f);
      else if (act == 2) o = o > 0.f ? o : slope * o;
      v[j] = o;
    }
    store_bf16x8(y + r * ldy + g * 8, v, 0);
  }
}

// backward: g = dy * act'(gamma xhat + beta); dbeta = sum g; dgamma = sum g xhat;
//           dx = gamma invstd / N * (N g - dbeta - xhat dgamma)            (torch.nn.functional.batch_norm backward)
// Pass 1: partial sums (dbeta, dgamma) with the bn_partial layout; pass 2: finalise into dgb [2][C] fp32; pass 3: dx.
__global__ void __launch_bounds__(kBnThreads) bn_bwd_partial_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
                                                                     long long rows, int C, int ld, const float* __restrict__ stats,
                                                                     const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                     int act, float slope, double* __restrict__ part) {
  griddep_sync();
  extern __shared__ double bn_sh[];
  const int groups = C / 8;
  const int g = threadIdx.x % groups, rl = threadIdx.x / groups, rpi = kBnThreads / groups;
  float mu[8], is[8], ga[8], be[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { const int c = g * 8 + j; mu[j] = stats[c]; is[j] = stats[C + c]; ga[j] = gamma[c]; be[j] = beta[c]; }
  double d1[8], d2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { d1[j] = 0.0; d2[j] = 0.0; }
  float s1[8], s2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { s1[j] = 0.f; s2[j] = 0.f; }
  int since = 0;
  if (rl < rpi) {
    for (long long r = (long long)blockIdx.x * rpi + rl; r < rows; r += (long long)gridDim.x * rpi) {
      float xv[8], gv[8];
      load_bf16x8(x + r * ld + g * 8, xv, 0);
      load_bf16x8(dy + r * ld + g * 8, gv, 0);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float xh = (xv[j] - mu[j]) * is[j];
        const float pre = ga[j] * xh + be[j];
        float gg = gv[j];
        if (act == 1) gg = pre > 0.f ? gg : 0.f;
        else if (act == 2) gg = pre > 0.f ? gg : slope * gg;
        s1[j] += gg;
        s2[j] = fmaf(gg, xh, s2[j]);
      }
      if (++since == 64) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { d1[j] += s1[j]; d2[j] += s2[j]; s1[j] = 0.f; s2[j] = 0.f; }
        since = 0;
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) { d1[j] += s1[j]; d2[j] += s2[j]; }
#pragma unroll
    for (int j = 0; j < 8; ++j) { bn_sh[(rl * 2 + 0) * C + g * 8 + j] = d1[j]; bn_sh[(rl * 2 + 1) * C + g * 8 + j] = d2[j]; }
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
  for (int i = lane; i < nblk; i += 32) t += part[((long long)i * 2 + which) * C + ch];
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  if (lane == 0) dgb[c] = float(t);
}
__global__ void bn_bwd_apply_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x, long long rows, int C, int ld,
                                    const float* __restrict__ stats, const float* __restrict__ gamma, const float* __restrict__ beta,
                                    int act, float slope, const float* __restrict__ dgb, float inv_count,
                                    __nv_bfloat16* __restrict__ dx, int lddx) {
  griddep_sync();
  const int groups = C / 8;
  const long long total = rows * groups;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int g = int(t % groups);
    const long long r = t / groups;
    float xv[8], gv[8];
    load_bf16x8(x + r * ld + g * 8, xv, 0);
    load_bf16x8(dy + r * ld + g * 8, gv, 0);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = g * 8 + j;
      const float xh = (xv[j] - stats[c]) * stats[C + c];
      const float pre = gamma[c] * xh + beta[c];
      float gg = gv[j];
      if (act == 1) gg = pre > 0.f ? gg : 0.f;
      else if (act == 2) gg = pre > 0.f ? gg : slope * gg;
      gv[j] = gamma[c] * stats[C + c] * (gg - inv_count * (dgb[c] + xh * dgb[C + c]));
    }
    store_bf16x8(dx + r * lddx + g * 8, gv, 0);
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
