// Memory-bound network operators around the convolutions (NHWC fp32): input normalisation,
// BatchNorm (+residual +ReLU) forward/backward, 3x3/2 max-pool, decoder glue
// (nearest-upsample + concat + reflection pad, and its adjoint), activation backward with bias
// reduction, global mean, Adam.  Each replaces the ATen op named in include/fusiondepth_b200.h.
//
// All kernels map threadIdx.x to the channel (fastest) dimension so every warp access is a
// contiguous 128 B line; reductions over pixels accumulate in fp64 so the batch statistics match
// the reference's two-pass CPU result to fp32 rounding.
#include <stdlib.h>

#include "common.cuh"
#include "../../include/fusiondepth_b200.h"

namespace {

__device__ __forceinline__ float act_grad(float y, int act) {
  // derivative expressed through the activation OUTPUT y
  switch (act) {
    case FD_ACT_RELU: return y > 0.f ? 1.f : 0.f;
    case FD_ACT_ELU: return y > 0.f ? 1.f : y + 1.f;          // alpha = 1: exp(x) = y + 1
    case FD_ACT_SIGMOID: return y * (1.f - y);
    case FD_ACT_TANH: return 1.f - y * y;
    default: return 1.f;
  }
}

__global__ void prep_input_kernel(const float* __restrict__ x, float* __restrict__ y, int C, int HW,
                                  float mean, float stdv) {
  // x [B,C,HW] -> y [B,HW,C]; one thread per output pixel, C is tiny (2..6)
  const int b = blockIdx.y;
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= HW) return;
  for (int c = 0; c < C; ++c)
    y[((long)b * HW + p) * C + c] = __fdiv_rn(__fadd_rn(x[((long)b * C + c) * HW + p], -mean), stdv);
}

// ---- 7x7/2 stem as a GEMM: normalised im2col rows ------------------------------------------------
// A[m][k], m = (b,ho,wo), k = (kh*KW + kw)*C + c for k < KH*KW*C, zero up to Kpad (a multiple of 32 so
// the tensor-core 1x1 path takes it).  Values are (x - mean)/std inside the image, 0 in the padding.
// One block per (image, output row): the KH input rows it touches are normalised once into shared
// memory (zero columns on both sides stand in for the padding), then thread (k quad, pixel lane)
// streams float4 rows of A with four shared-memory loads each -- the index arithmetic is hoisted out
// of the pixel loop and every input element is divided by std once instead of once per tap.
__global__ void __launch_bounds__(256) stem_im2col_kernel(const float* __restrict__ x, float* __restrict__ A,
                                                          int C, int H, int W, int Ho, int Wo, int KH, int KW,
                                                          int stride, int pad, int Kpad, float mean, float stdv) {
  extern __shared__ float sm[];                       // [C][KH][Wp]
  const int b = blockIdx.y, ho = blockIdx.x, t = threadIdx.x;
  const int Wp = W + 2 * pad, rowlen = KH * Wp;
  for (int idx = t; idx < C * rowlen; idx += blockDim.x) {
    const int c = idx / rowlen, r = idx - c * rowlen;
    const int kh = r / Wp, w = r - kh * Wp - pad;
    const int h = ho * stride - pad + kh;
    float v = 0.f;
    if (h >= 0 && h < H && w >= 0 && w < W)
      v = __fdiv_rn(__fadd_rn(x[(((long)b * C + c) * H + h) * W + w], -mean), stdv);
    sm[idx] = v;
  }
  __syncthreads();
  const int Kq = Kpad >> 2, G = blockDim.x / Kq;
  const int q = t % Kq, g = t / Kq;
  if (g >= G) return;
  int off[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int k = 4 * q + e;
    off[e] = -1;
    if (k < KH * KW * C) {
      const int c = k % C, tap = k / C;
      const int kh = tap / KW, kw = tap - kh * KW;
      off[e] = (c * KH + kh) * Wp + kw;
    }
  }
  float* row = A + ((long)(b * Ho + ho) * Wo) * Kpad + 4 * q;
  for (int wo = g; wo < Wo; wo += G) {
    const int xo = wo * stride;
    float4 v;
    v.x = off[0] >= 0 ? sm[off[0] + xo] : 0.f;
    v.y = off[1] >= 0 ? sm[off[1] + xo] : 0.f;
    v.z = off[2] >= 0 ? sm[off[2] + xo] : 0.f;
    v.w = off[3] >= 0 ? sm[off[3] + xo] : 0.f;
    *reinterpret_cast<float4*>(row + (long)wo * Kpad) = v;
  }
}

// dst[r][0..cols_dst) = src[r][0..cols_src) zero-padded (cols_dst >= cols_src), or the reverse
// accumulate (unpad: dst[r][c] += src[r][c] for c < cols_dst)
__global__ void pad_rows_kernel(const float* __restrict__ src, float* __restrict__ dst, int rows, int cols_src,
                                int cols_dst, int accumulate) {
  long n = (long)rows * cols_dst;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    int c = (int)(i % cols_dst);
    long r = i / cols_dst;
    float v = c < cols_src ? src[r * cols_src + c] : 0.f;
    if (accumulate) atomicAdd(&dst[i], v); else dst[i] = v;
  }
}

// ---- activation backward + bias gradient ---------------------------------------------------
// grid.x covers channel groups of 32, grid.y pixel chunks; block (32, 8)
__global__ void act_bwd_kernel(const float* __restrict__ y, const float* __restrict__ dy,
                               float* __restrict__ dpre, float* __restrict__ dbias, long M, int C,
                               int act) {
  __shared__ float red[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  float acc = 0.f;
  if (c < C) {
    for (long m = blockIdx.y * 8 + threadIdx.y; m < M; m += (long)gridDim.y * 8) {
      long o = m * C + c;
      float g = dy[o] * act_grad(y[o], act);
      dpre[o] = g;
      acc += g;
    }
  }
  if (dbias == nullptr) return;
  red[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += red[i][threadIdx.x];
    atomicAdd(&dbias[c], s);
  }
}

// ---- BatchNorm -------------------------------------------------------------------------------
// Every thread owns VEC consecutive channels (float4 when C % 4 == 0) and strides over pixels with
// four independent loads in flight; block = (channel groups, pixel lanes), 256 threads.
template <int VEC> struct VecT;
template <> struct VecT<4> { using T = float4; };
template <> struct VecT<1> { using T = float; };
template <int VEC>
__device__ __forceinline__ void ldv(const float* p, float (&v)[VEC]) {
  if (VEC == 4) { float4 t = *reinterpret_cast<const float4*>(p); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[VEC - 1] = t.w; }
  else v[0] = p[0];
}
template <int VEC>
__device__ __forceinline__ void stv(float* p, const float (&v)[VEC]) {
  if (VEC == 4) *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[VEC - 1]);
  else p[0] = v[0];
}

struct BnGeom { int groups, gx, gy; dim3 grid, block; };
constexpr int BN_MAX_PARTS = 296;     // partial rows of the statistic kernels (workspace: (1 + parts) * 2C doubles + counters)

// Last-block fold of the per-block partial sums: block (bx, by) has written its 2*VEC*nx partials to
// part[by][...]; the last block of column bx to arrive (ticket) adds the rows in a fixed order, so the result
// is deterministic and no two blocks ever touch the same address atomically.
template <int VEC>
__device__ __forceinline__ void bn_fold_partials(double* __restrict__ ws, double a[VEC], double b[VEC], int c, int C,
                                                 bool owner, double* red) {
  double* part = ws + 2 * (long)C;
  unsigned* counter = reinterpret_cast<unsigned*>(part + (long)BN_MAX_PARTS * 2 * C);
  __shared__ unsigned s_last;
  if (owner) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      part[(long)blockIdx.y * 2 * C + c + i] = a[i];
      part[(long)blockIdx.y * 2 * C + C + c + i] = b[i];
    }
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0 && threadIdx.y == 0) s_last = atomicAdd(&counter[blockIdx.x], 1u) == gridDim.y - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // the whole last block folds: thread (x, y) takes rows y, y + ny, ... of its channels, then the ny lanes meet
  const int nx = blockDim.x, ny = blockDim.y;
  double sa[VEC], sb[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) sa[i] = sb[i] = 0.0;
  if (c < C) {
    // four partial rows in flight per thread (the row loop is a chain of L2 round trips otherwise: with 96 rows and
    // 16 lanes it cost more than streaming the tensor), summed in row order
#pragma unroll 1
    for (unsigned r0 = threadIdx.y; r0 < gridDim.y; r0 += 4 * ny) {
      double va[4][VEC], vb[4][VEC];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const unsigned r = r0 + j * ny;
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
          va[j][i] = r < gridDim.y ? part[(long)r * 2 * C + c + i] : 0.0;
          vb[j][i] = r < gridDim.y ? part[(long)r * 2 * C + C + c + i] : 0.0;
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < VEC; ++i) { sa[i] += va[j][i]; sb[i] += vb[j][i]; }
    }
  }
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    red[((0 * VEC + i) * ny + threadIdx.y) * nx + threadIdx.x] = sa[i];
    red[((1 * VEC + i) * ny + threadIdx.y) * nx + threadIdx.x] = sb[i];
  }
  __syncthreads();
  if (owner) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      double ta = 0, tb = 0;
      for (int r = 0; r < ny; ++r) {                      // fixed order: deterministic
        ta += red[((0 * VEC + i) * ny + r) * nx + threadIdx.x];
        tb += red[((1 * VEC + i) * ny + r) * nx + threadIdx.x];
      }
      ws[c + i] = ta;
      ws[C + c + i] = tb;
    }
  }
  if (threadIdx.x == 0 && threadIdx.y == 0) counter[blockIdx.x] = 0;      // ready for the next use of this workspace
}

// reduce = true: geometry of the two statistic kernels.  Their blocks end with 2*C fp64 atomics on the
// same C addresses, which serialise in L2: a few dozen blocks stream the (L2-sized) tensor just as
// fast and keep the contention per address small.
inline BnGeom bn_geom(long M, int C, int vec, bool reduce = false) {
  BnGeom g;
  g.groups = C / vec;
  g.gx = g.groups < 32 ? g.groups : 32;          // channel groups per block row
  g.gy = 256 / g.gx;                              // pixel lanes per block
  int bx = (g.groups + g.gx - 1) / g.gx;
  long by = (M + (long)g.gy * 8 - 1) / ((long)g.gy * 8);
  // statistic kernels: every block leaves one partial row (2*C doubles) for the last block of its channel
  // column to fold, so a few blocks per SM stream the tensor at L2 speed without contended atomics
  long cap = reduce ? (148L * 2 + bx - 1) / bx : (148L * 8 + bx - 1) / bx;
  // Partial rows: L2-sized tensors stream just as fast through ~96 blocks (and leave the SMs to the concurrent
  // branches of the step: 96 / 148 / 296 measured within noise, 497-502 images/s); tensors far beyond L2 (the
  // ResNet-50 layers at 1024x320, the stem) are HBM-bound and want two blocks per SM.  FD_BN_PARTS overrides.
  static int parts_env = -2;
  if (parts_env == -2) {
    const char* e = getenv("FD_BN_PARTS");
    parts_env = e ? atoi(e) : -1;
    if (parts_env > BN_MAX_PARTS) parts_env = BN_MAX_PARTS;
  }
  const long parts_cap = parts_env > 0 ? parts_env : (M * C * 4L < (32L << 20) ? 96 : BN_MAX_PARTS);
  if (reduce && cap > parts_cap) cap = parts_cap;
  if (by > cap) by = cap;
  if (by < 1) by = 1;
  g.grid = dim3(bx, (unsigned)by);
  g.block = dim3(g.gx, g.gy);
  return g;
}

// pass 1: per-channel sum / sum of squares (fp32 over short runs, folded into fp64)
template <int VEC>
__global__ void bn_stats_kernel(const float* __restrict__ x, double* __restrict__ ws, long M, int C) {
  extern __shared__ double red[];                  // [2][VEC][blockDim.y][blockDim.x]
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  const int c = q * VEC;
  double s[VEC], ss[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) s[i] = ss[i] = 0.0;
  if (c < C) {
    const long step = (long)gridDim.y * blockDim.y;
    for (long m0 = (long)blockIdx.y * blockDim.y + threadIdx.y; m0 < M; m0 += step * 8) {
      float ps[VEC], pss[VEC];
#pragma unroll
      for (int i = 0; i < VEC; ++i) ps[i] = pss[i] = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        long m = m0 + j * step;
        if (m < M) {
          float v[VEC];
          ldv<VEC>(x + m * C + c, v);
#pragma unroll
          for (int i = 0; i < VEC; ++i) { ps[i] += v[i]; pss[i] = fmaf(v[i], v[i], pss[i]); }
        }
      }
#pragma unroll
      for (int i = 0; i < VEC; ++i) { s[i] += (double)ps[i]; ss[i] += (double)pss[i]; }
    }
  }
  const int nx = blockDim.x, ny = blockDim.y;
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    red[((0 * VEC + i) * ny + threadIdx.y) * nx + threadIdx.x] = s[i];
    red[((1 * VEC + i) * ny + threadIdx.y) * nx + threadIdx.x] = ss[i];
  }
  __syncthreads();
  double fa[VEC], fb[VEC];
  const bool owner = threadIdx.y == 0 && c < C;
  if (owner) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      double a = 0, b = 0;
      for (int r = 0; r < ny; ++r) {
        a += red[((0 * VEC + i) * ny + r) * nx + threadIdx.x];
        b += red[((1 * VEC + i) * ny + r) * nx + threadIdx.x];
      }
      fa[i] = a; fb[i] = b;
    }
  }
  __syncthreads();                                         // `red` is reused by the fold
  bn_fold_partials<VEC>(ws, fa, fb, c, C, owner, red);
}

// Per-channel batch statistics from the fp64 sums (training) or the running buffers (eval); the
// running-statistic update and the saved mean / rstd are written once, by the first block row.
struct BnStat { float mean, rstd; };
__device__ __forceinline__ BnStat bn_channel_stat(const double* __restrict__ ws, float* running_mean,
                                                  float* running_var, int training, float momentum, float eps,
                                                  float stat_weight, float* save_mean, float* save_rstd,
                                                  long M, int C, int c, bool writer) {
  BnStat r;
  if (training) {
    double mean = ws[c] / (double)M;
    double var = ws[C + c] / (double)M - mean * mean;
    if (var < 0) var = 0;
    r.mean = (float)mean;
    r.rstd = (float)(1.0 / sqrt(var + (double)eps));
    if (writer) {
      double unb = M > 1 ? var * (double)M / (double)(M - 1) : var;
      if (stat_weight >= 0.f) {
        // order-free form for calls that run concurrently (other micro-batch / other stream): the
        // caller pre-decays the buffer once per step and every call adds its weighted statistic
        atomicAdd(&running_mean[c], (float)((double)stat_weight * mean));
        atomicAdd(&running_var[c], (float)((double)stat_weight * unb));
      } else {
        running_mean[c] = (float)((1.0 - momentum) * (double)running_mean[c] + (double)momentum * mean);
        running_var[c] = (float)((1.0 - momentum) * (double)running_var[c] + (double)momentum * unb);
      }
    }
  } else {
    r.mean = running_mean[c];
    r.rstd = (float)(1.0 / sqrt((double)running_var[c] + (double)eps));
  }
  if (writer) { save_mean[c] = r.mean; save_rstd[c] = r.rstd; }
  return r;
}

// The normalise + affine expression with its rounding steps spelled out: the backward kernels re-evaluate it to
// rebuild the ReLU mask from x (fd_bn_bwd_xmask), which only works if both sides round identically.
__device__ __forceinline__ float bn_affine(float v, float mu, float rs, float g, float bt) {
  return fmaf(__fmul_rn(__fsub_rn(v, mu), rs), g, bt);
}

// pass 2 (finalize folded in): y = relu?((x - mean) * rstd * gamma + beta [+ residual])
template <int VEC>
__global__ void bn_apply_kernel(const float* __restrict__ x, const float* __restrict__ res,
                                const float* __restrict__ gamma, const float* __restrict__ beta,
                                const double* __restrict__ ws, float* running_mean, float* running_var,
                                int training, float momentum, float eps, float stat_weight, float* save_mean,
                                float* save_rstd, int relu, float* __restrict__ y, long M, int C) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) * VEC;
  if (c >= C) return;
  float mu[VEC], rs[VEC], g[VEC], bt[VEC];
  const bool writer = blockIdx.y == 0 && threadIdx.y == 0;
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    BnStat st = bn_channel_stat(ws, running_mean, running_var, training, momentum, eps, stat_weight, save_mean,
                                save_rstd, M, C, c + i, writer);
    mu[i] = st.mean; rs[i] = st.rstd; g[i] = gamma[c + i]; bt[i] = beta[c + i];
  }
  const long step = (long)gridDim.y * blockDim.y;
  for (long m = (long)blockIdx.y * blockDim.y + threadIdx.y; m < M; m += step) {
    float v[VEC], r[VEC];
    ldv<VEC>(x + m * C + c, v);
    if (res) ldv<VEC>(res + m * C + c, r);
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      float o = bn_affine(v[i], mu[i], rs[i], g[i], bt[i]);
      if (res) o += r[i];
      if (relu) o = fmaxf(o, 0.f);
      v[i] = o;
    }
    stv<VEC>(y + m * C + c, v);
  }
}

// backward pass 1: sum g, sum g*xhat with g = dy * relu_mask.  The mask is y > 0, with y either read back or
// (y == nullptr, residual-free BatchNorm + ReLU) re-evaluated from x with mgamma / mbeta: one tensor read less.
template <int VEC>
__global__ void bn_bwd_reduce_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                     const float* __restrict__ dy, const float* __restrict__ mean,
                                     const float* __restrict__ rstd, int relu,
                                     double* __restrict__ ws, long M, int C,
                                     const float* __restrict__ mgamma, const float* __restrict__ mbeta) {
  extern __shared__ double red[];
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) * VEC;
  double s[VEC], ss[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) s[i] = ss[i] = 0.0;
  if (c < C) {
    float mu[VEC], rs[VEC], mg[VEC], mb[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      mu[i] = mean[c + i]; rs[i] = rstd[c + i];
      mg[i] = (relu && !y) ? mgamma[c + i] : 0.f; mb[i] = (relu && !y) ? mbeta[c + i] : 0.f;
    }
    const long step = (long)gridDim.y * blockDim.y;
    for (long m0 = (long)blockIdx.y * blockDim.y + threadIdx.y; m0 < M; m0 += step * 4) {
      float ps[VEC], pss[VEC];
#pragma unroll
      for (int i = 0; i < VEC; ++i) ps[i] = pss[i] = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        long m = m0 + j * step;
        if (m < M) {
          float xv[VEC], yv[VEC], gv[VEC];
          ldv<VEC>(x + m * C + c, xv);
          ldv<VEC>(dy + m * C + c, gv);
          if (relu && y) ldv<VEC>(y + m * C + c, yv);
#pragma unroll
          for (int i = 0; i < VEC; ++i) {
            float g = gv[i];
            if (relu && !y) yv[i] = bn_affine(xv[i], mu[i], rs[i], mg[i], mb[i]);
            if (relu && !(yv[i] > 0.f)) g = 0.f;
            ps[i] += g;
            pss[i] = fmaf(g, (xv[i] - mu[i]) * rs[i], pss[i]);
          }
        }
      }
#pragma unroll
      for (int i = 0; i < VEC; ++i) { s[i] += (double)ps[i]; ss[i] += (double)pss[i]; }
    }
  }
  const int nx = blockDim.x, ny = blockDim.y;
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    red[((0 * VEC + i) * ny + threadIdx.y) * nx + threadIdx.x] = s[i];
    red[((1 * VEC + i) * ny + threadIdx.y) * nx + threadIdx.x] = ss[i];
  }
  __syncthreads();
  double fa[VEC], fb[VEC];
  const bool owner = threadIdx.y == 0 && c < C;
  if (owner) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      double a = 0, b = 0;
      for (int r = 0; r < ny; ++r) {
        a += red[((0 * VEC + i) * ny + r) * nx + threadIdx.x];
        b += red[((1 * VEC + i) * ny + r) * nx + threadIdx.x];
      }
      fa[i] = a; fb[i] = b;
    }
  }
  __syncthreads();                                         // `red` is reused by the fold
  bn_fold_partials<VEC>(ws, fa, fb, c, C, owner, red);
}

template <int VEC>
__global__ void bn_bwd_apply_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                    const float* __restrict__ dy, const float* __restrict__ gamma,
                                    const float* __restrict__ mean, const float* __restrict__ rstd,
                                    int relu, int training, const double* __restrict__ ws,
                                    float* __restrict__ dx, float* __restrict__ dres,
                                    float* __restrict__ dgamma, float* __restrict__ dbeta, long M,
                                    int C, int accumulate, const float* __restrict__ mbeta) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) * VEC;
  if (c >= C) return;
  float mu[VEC], rs[VEC], gm[VEC], k1[VEC], k2[VEC], mb[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    mu[i] = mean[c + i]; rs[i] = rstd[c + i]; gm[i] = gamma[c + i]; mb[i] = (relu && !y) ? mbeta[c + i] : 0.f;
    const float sg = (float)ws[c + i], sgx = (float)ws[C + c + i];
    k1[i] = training ? sg / (float)M : 0.f;
    k2[i] = training ? sgx / (float)M : 0.f;
    if (blockIdx.y == 0 && threadIdx.y == 0) {
      if (accumulate) {                 // several streams may add into the same parameter gradient
        if (dgamma) atomicAdd(&dgamma[c + i], sgx);
        if (dbeta) atomicAdd(&dbeta[c + i], sg);
      } else {
        if (dgamma) dgamma[c + i] = sgx;
        if (dbeta) dbeta[c + i] = sg;
      }
    }
  }
  const long step = (long)gridDim.y * blockDim.y;
  for (long m = (long)blockIdx.y * blockDim.y + threadIdx.y; m < M; m += step) {
    float xv[VEC], yv[VEC], gv[VEC];
    ldv<VEC>(x + m * C + c, xv);
    ldv<VEC>(dy + m * C + c, gv);
    if (relu && y) ldv<VEC>(y + m * C + c, yv);
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      if (relu && !y) yv[i] = bn_affine(xv[i], mu[i], rs[i], gm[i], mb[i]);
      if (relu && !(yv[i] > 0.f)) gv[i] = 0.f;
    }
    if (dres) stv<VEC>(dres + m * C + c, gv);
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      float xh = (xv[i] - mu[i]) * rs[i];
      xv[i] = gm[i] * rs[i] * (gv[i] - k1[i] - xh * k2[i]);
    }
    stv<VEC>(dx + m * C + c, xv);
  }
}

// ---- small-M BatchNorm: one kernel per direction ----------------------------------------------
// layer4 tensors (M = B*H*W <= 1024 rows) are a few hundred KB: two or three dependent
// launches with fp64 atomics in between cost ~12 us each in launch + DRAM latency.  Here a CTA owns 8
// channels for ALL rows, so the batch statistics never leave the block: pass 1 reduces, pass 2
// re-reads the (L1/L2-resident) columns and writes.  Thread = (channel quad, row lane).
constexpr int BNF_CG = 8;                 // channels per CTA
constexpr int BNF_LANES = 256 / (BNF_CG / 4);

__device__ __forceinline__ void bnf_block_reduce(double (&s)[4], double (&ss)[4], double* red, int cq, int rl) {
  // red: [BNF_LANES][BNF_CG/4][8] doubles; result (all row lanes summed) returned to every thread
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    red[(rl * (BNF_CG / 4) + cq) * 8 + i] = s[i];
    red[(rl * (BNF_CG / 4) + cq) * 8 + 4 + i] = ss[i];
  }
  __syncthreads();
  for (int stride = BNF_LANES / 2; stride > 0; stride >>= 1) {
    if (rl < stride) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        red[(rl * (BNF_CG / 4) + cq) * 8 + i] += red[((rl + stride) * (BNF_CG / 4) + cq) * 8 + i];
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) { s[i] = red[cq * 8 + i]; ss[i] = red[cq * 8 + 4 + i]; }
  __syncthreads();
}

__global__ void __launch_bounds__(256) bn_fwd_fused_kernel(
    const float* __restrict__ x, const float* __restrict__ res, const float* __restrict__ gamma,
    const float* __restrict__ beta, float* running_mean, float* running_var, int training, float momentum,
    float eps, float stat_weight, float* save_mean, float* save_rstd, int relu, float* __restrict__ y, int M,
    int C) {
  extern __shared__ double red[];
  const int cq = threadIdx.x % (BNF_CG / 4), rl = threadIdx.x / (BNF_CG / 4);
  const int c = blockIdx.x * BNF_CG + cq * 4;
  float mu[4], rs[4];
  if (training) {
    double s[4] = {0, 0, 0, 0}, ss[4] = {0, 0, 0, 0};
    for (int m0 = rl; m0 < M; m0 += BNF_LANES * 8) {
      float ps[4] = {0.f, 0.f, 0.f, 0.f}, pss[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int m = m0 + j * BNF_LANES;
        if (m < M) {
          float v[4];
          ldv<4>(x + (long)m * C + c, v);
#pragma unroll
          for (int i = 0; i < 4; ++i) { ps[i] += v[i]; pss[i] = fmaf(v[i], v[i], pss[i]); }
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) { s[i] += (double)ps[i]; ss[i] += (double)pss[i]; }
    }
    bnf_block_reduce(s, ss, red, cq, rl);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const double mean = s[i] / (double)M;
      double var = ss[i] / (double)M - mean * mean;
      if (var < 0) var = 0;
      mu[i] = (float)mean;
      rs[i] = (float)(1.0 / sqrt(var + (double)eps));
      if (rl == 0) {
        const double unb = M > 1 ? var * (double)M / (double)(M - 1) : var;
        if (stat_weight >= 0.f) {
          atomicAdd(&running_mean[c + i], (float)((double)stat_weight * mean));
          atomicAdd(&running_var[c + i], (float)((double)stat_weight * unb));
        } else {
          running_mean[c + i] = (float)((1.0 - momentum) * (double)running_mean[c + i] + (double)momentum * mean);
          running_var[c + i] = (float)((1.0 - momentum) * (double)running_var[c + i] + (double)momentum * unb);
        }
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      mu[i] = running_mean[c + i];
      rs[i] = (float)(1.0 / sqrt((double)running_var[c + i] + (double)eps));
    }
  }
  float g[4], bt[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    g[i] = gamma[c + i]; bt[i] = beta[c + i];
    if (rl == 0) { save_mean[c + i] = mu[i]; save_rstd[c + i] = rs[i]; }
  }
  for (int m = rl; m < M; m += BNF_LANES) {
    float v[4], r[4];
    ldv<4>(x + (long)m * C + c, v);
    if (res) ldv<4>(res + (long)m * C + c, r);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float o = (v[i] - mu[i]) * rs[i] * g[i] + bt[i];
      if (res) o += r[i];
      if (relu) o = fmaxf(o, 0.f);
      v[i] = o;
    }
    stv<4>(y + (long)m * C + c, v);
  }
}

__global__ void __launch_bounds__(256) bn_bwd_fused_kernel(
    const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ dy,
    const float* __restrict__ gamma, const float* __restrict__ mean, const float* __restrict__ rstd, int relu,
    int training, float* __restrict__ dx, float* __restrict__ dres, float* dgamma, float* dbeta, int M, int C,
    int accumulate) {
  extern __shared__ double red[];
  const int cq = threadIdx.x % (BNF_CG / 4), rl = threadIdx.x / (BNF_CG / 4);
  const int c = blockIdx.x * BNF_CG + cq * 4;
  float mu[4], rs[4], gm[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { mu[i] = mean[c + i]; rs[i] = rstd[c + i]; gm[i] = gamma[c + i]; }
  double s[4] = {0, 0, 0, 0}, ss[4] = {0, 0, 0, 0};
  for (int m0 = rl; m0 < M; m0 += BNF_LANES * 4) {
    float ps[4] = {0.f, 0.f, 0.f, 0.f}, pss[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int m = m0 + j * BNF_LANES;
      if (m < M) {
        float xv[4], yv[4], gv[4];
        ldv<4>(x + (long)m * C + c, xv);
        ldv<4>(dy + (long)m * C + c, gv);
        if (relu) ldv<4>(y + (long)m * C + c, yv);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float g = gv[i];
          if (relu && !(yv[i] > 0.f)) g = 0.f;
          ps[i] += g;
          pss[i] = fmaf(g, (xv[i] - mu[i]) * rs[i], pss[i]);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) { s[i] += (double)ps[i]; ss[i] += (double)pss[i]; }
  }
  bnf_block_reduce(s, ss, red, cq, rl);
  float k1[4], k2[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float sg = (float)s[i], sgx = (float)ss[i];
    k1[i] = training ? sg / (float)M : 0.f;
    k2[i] = training ? sgx / (float)M : 0.f;
    if (rl == 0) {
      if (accumulate) {
        if (dgamma) atomicAdd(&dgamma[c + i], sgx);
        if (dbeta) atomicAdd(&dbeta[c + i], sg);
      } else {
        if (dgamma) dgamma[c + i] = sgx;
        if (dbeta) dbeta[c + i] = sg;
      }
    }
  }
  for (int m = rl; m < M; m += BNF_LANES) {
    float xv[4], yv[4], gv[4];
    ldv<4>(x + (long)m * C + c, xv);
    ldv<4>(dy + (long)m * C + c, gv);
    if (relu) ldv<4>(y + (long)m * C + c, yv);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (relu && !(yv[i] > 0.f)) gv[i] = 0.f;
    }
    if (dres) stv<4>(dres + (long)m * C + c, gv);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float xh = (xv[i] - mu[i]) * rs[i];
      xv[i] = gm[i] * rs[i] * (gv[i] - k1[i] - xh * k2[i]);
    }
    stv<4>(dx + (long)m * C + c, xv);
  }
}

static bool bn_fused_ok(long M, int C) {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("FD_BN_FUSED");
    on = (e && e[0] == '1') ? 1 : 0;     // opt-in (FD_BN_FUSED=1): measured no faster than the two-kernel form
  }
  return on && M <= 1024 && C % BNF_CG == 0;
}
constexpr size_t BNF_SMEM = sizeof(double) * BNF_LANES * (BNF_CG / 4) * 8;

// ---- max-pool 3x3 stride 2 pad 1 -------------------------------------------------------------
__global__ void maxpool_fwd_kernel(const float* __restrict__ x, float* __restrict__ y,
                                   unsigned char* __restrict__ idx, int H, int W, int C, int Ho,
                                   int Wo) {
  const int c = blockIdx.x * 32 + threadIdx.x;
  const int b = blockIdx.z;
  if (c >= C) return;
  for (int p = blockIdx.y * blockDim.y + threadIdx.y; p < Ho * Wo; p += gridDim.y * blockDim.y) {
    int ho = p / Wo, wo = p % Wo;
    float best = -INFINITY;
    int bi = 0;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh)
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        int h = 2 * ho - 1 + kh, w = 2 * wo - 1 + kw;
        if (h >= 0 && h < H && w >= 0 && w < W) {
          float v = x[(((long)b * H + h) * W + w) * C + c];
          if (v > best || v != v) { best = v; bi = kh * 3 + kw; }
        }
      }
    long o = (((long)b * Ho + ho) * Wo + wo) * C + c;
    y[o] = best;
    idx[o] = (unsigned char)bi;
  }
}

__global__ void maxpool_bwd_kernel(const float* __restrict__ dy, const unsigned char* __restrict__ idx,
                                   float* __restrict__ dx, int H, int W, int C, int Ho, int Wo) {
  const int c = blockIdx.x * 32 + threadIdx.x;
  const int b = blockIdx.z;
  if (c >= C) return;
  for (int p = blockIdx.y * blockDim.y + threadIdx.y; p < H * W; p += gridDim.y * blockDim.y) {
    int h = p / W, w = p % W;
    float acc = 0.f;
    // windows (ho,wo) containing (h,w): 2ho-1 <= h <= 2ho+1
    for (int ho = h / 2; ho <= (h + 1) / 2; ++ho) {
      if (ho >= Ho) continue;
      int kh = h - (2 * ho - 1);
      for (int wo = w / 2; wo <= (w + 1) / 2; ++wo) {
        if (wo >= Wo) continue;
        int kw = w - (2 * wo - 1);
        long o = (((long)b * Ho + ho) * Wo + wo) * C + c;
        if (idx[o] == kh * 3 + kw) acc += dy[o];
      }
    }
    dx[(((long)b * H + h) * W + w) * C + c] = acc;
  }
}

// One thread per (pixel, 4 channels): float4 / uchar4 traffic, 32-bit indices (C % 4 == 0, < 2^32 elements).
__global__ void maxpool_fwd_v4_kernel(const float* __restrict__ x, float* __restrict__ y,
                                      unsigned char* __restrict__ idx, int H, int W, int C, int Ho, int Wo,
                                      unsigned total) {
  const unsigned Cq = C >> 2;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const unsigned p = i / Cq, c = (i - p * Cq) << 2;
    const unsigned row = p / Wo, wo = p - row * Wo, b = row / Ho, ho = row - b * Ho;
    float best[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    unsigned bi[4] = {0, 0, 0, 0};
#pragma unroll
    for (int kh = 0; kh < 3; ++kh)
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int h = 2 * (int)ho - 1 + kh, w = 2 * (int)wo - 1 + kw;
        if (h >= 0 && h < H && w >= 0 && w < W) {
          const float4 v4 = *reinterpret_cast<const float4*>(x + (((long)b * H + h) * W + w) * C + c);
          const float v[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (v[e] > best[e] || v[e] != v[e]) { best[e] = v[e]; bi[e] = kh * 3 + kw; }
        }
      }
    *reinterpret_cast<float4*>(y + (long)i * 4) = make_float4(best[0], best[1], best[2], best[3]);
    *reinterpret_cast<uchar4*>(idx + (long)i * 4) = make_uchar4(bi[0], bi[1], bi[2], bi[3]);
  }
}

__global__ void maxpool_bwd_v4_kernel(const float* __restrict__ dy, const unsigned char* __restrict__ idx,
                                      float* __restrict__ dx, int H, int W, int C, int Ho, int Wo, unsigned total) {
  const unsigned Cq = C >> 2;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const unsigned p = i / Cq, c = (i - p * Cq) << 2;
    const unsigned row = p / W, w = p - row * W, b = row / H, h = row - b * H;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    // windows (ho,wo) containing (h,w): 2ho-1 <= h <= 2ho+1
    for (int ho = h / 2; ho <= (int)(h + 1) / 2; ++ho) {
      if (ho >= Ho) continue;
      const int kh = (int)h - (2 * ho - 1);
      for (int wo = w / 2; wo <= (int)(w + 1) / 2; ++wo) {
        if (wo >= Wo) continue;
        const int kk = kh * 3 + (int)w - (2 * wo - 1);
        const long o = (((long)b * Ho + ho) * Wo + wo) * C + c;
        const uchar4 id = *reinterpret_cast<const uchar4*>(idx + o);
        const float4 g = *reinterpret_cast<const float4*>(dy + o);
        if (id.x == kk) acc[0] += g.x;
        if (id.y == kk) acc[1] += g.y;
        if (id.z == kk) acc[2] += g.z;
        if (id.w == kk) acc[3] += g.w;
      }
    }
    *reinterpret_cast<float4*>(dx + (long)i * 4) = make_float4(acc[0], acc[1], acc[2], acc[3]);
  }
}

// ---- decoder glue ----------------------------------------------------------------------------
constexpr int MAXSEG = 4;
struct AsmArgs {
  const float* a[MAXSEG];
  const float* b[MAXSEG];
  float* d[MAXSEG];
  float* d2[MAXSEG];   // backward: optional second copy of d (the `b` operand's own gradient tensor)
  int C[MAXSEG], up[MAXSEG], off[MAXSEG];
  int nseg, Ctot, B, H, W, pad;
};

__device__ __forceinline__ int refl(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}

__global__ void assemble_fwd_kernel(AsmArgs a, float* __restrict__ out) {
  const int Hp = a.H + 2 * a.pad, Wp = a.W + 2 * a.pad;
  const long npix = (long)a.B * Hp * Wp;
  for (long p = blockIdx.x; p < npix; p += gridDim.x) {
    int wp = p % Wp, hp = (p / Wp) % Hp, b = p / ((long)Wp * Hp);
    int h = refl(hp - a.pad, a.H), w = refl(wp - a.pad, a.W);
    float* o = out + p * a.Ctot;
    for (int c = threadIdx.x; c < a.Ctot; c += blockDim.x) {
      int s = 0;
#pragma unroll
      for (int i = 1; i < MAXSEG; ++i)
        if (i < a.nseg && c >= a.off[i]) s = i;
      int cs = c - a.off[s];
      int hs = a.up[s] ? h >> 1 : h, ws = a.up[s] ? w >> 1 : w;
      int Hs = a.up[s] ? a.H >> 1 : a.H, Ws = a.up[s] ? a.W >> 1 : a.W;
      long i = (((long)b * Hs + hs) * Ws + ws) * a.C[s] + cs;
      float v = a.a[s][i];
      if (a.b[s]) v += a.b[s][i];
      o[c] = v;
    }
  }
}

// adjoint, gather form: every source element sums the padded positions that read it
__global__ void assemble_bwd_kernel(AsmArgs a, const float* __restrict__ dout, int s) {
  const int Hp = a.H + 2 * a.pad, Wp = a.W + 2 * a.pad;
  const int u = a.up[s] ? 2 : 1;
  const int Hs = a.H / u, Ws = a.W / u, Cs = a.C[s];
  const long npix = (long)a.B * Hs * Ws;
  for (long p = blockIdx.x; p < npix; p += gridDim.x) {
    int ws = p % Ws, hs = (p / Ws) % Hs, b = p / ((long)Ws * Hs);
    for (int c = threadIdx.x; c < Cs; c += blockDim.x) {
      float acc = 0.f;
      for (int dy = 0; dy < u; ++dy) {
        int h = hs * u + dy;
        int hps[3], nh = 0;
        hps[nh++] = h + a.pad;
        if (a.pad) { if (h == 1) hps[nh++] = 0; if (h == a.H - 2) hps[nh++] = Hp - 1; }
        for (int dx = 0; dx < u; ++dx) {
          int w = ws * u + dx;
          int wps[3], nw = 0;
          wps[nw++] = w + a.pad;
          if (a.pad) { if (w == 1) wps[nw++] = 0; if (w == a.W - 2) wps[nw++] = Wp - 1; }
          for (int i = 0; i < nh; ++i)
            for (int j = 0; j < nw; ++j)
              acc += dout[(((long)b * Hp + hps[i]) * Wp + wps[j]) * a.Ctot + a.off[s] + c];
        }
      }
      a.d[s][p * Cs + c] = acc;
      if (a.d2[s]) a.d2[s][p * Cs + c] = acc;
    }
  }
}

// The same with one thread per (pixel, 4-channel group): float4 traffic, 32-bit index arithmetic, no per-thread
// 64-bit divisions.  (The scalar kernels above spend ~100 integer instructions per float moved: 69 us for the
// 96 MB full-resolution tensor against ~17 us of HBM time.)  Needs every segment's channel count % 4 == 0.
__global__ void assemble_fwd_v4_kernel(AsmArgs a, float* __restrict__ out, unsigned total) {
  const unsigned Hp = a.H + 2 * a.pad, Wp = a.W + 2 * a.pad, Cq = a.Ctot >> 2;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const unsigned p = i / Cq, c = (i - p * Cq) << 2;
    const unsigned row = p / Wp, wp = p - row * Wp, b = row / Hp, hp = row - b * Hp;
    const int h = refl((int)hp - a.pad, a.H), w = refl((int)wp - a.pad, a.W);
    int s = 0;
#pragma unroll
    for (int k = 1; k < MAXSEG; ++k)
      if (k < a.nseg && (int)c >= a.off[k]) s = k;
    const int cs = (int)c - a.off[s];
    const int hs = a.up[s] ? h >> 1 : h, ws = a.up[s] ? w >> 1 : w;
    const int Hs = a.up[s] ? a.H >> 1 : a.H, Ws = a.up[s] ? a.W >> 1 : a.W;
    const long src = (((long)b * Hs + hs) * Ws + ws) * a.C[s] + cs;
    float4 v = *reinterpret_cast<const float4*>(a.a[s] + src);
    if (a.b[s]) {
      const float4 y = *reinterpret_cast<const float4*>(a.b[s] + src);
      v = make_float4(v.x + y.x, v.y + y.y, v.z + y.z, v.w + y.w);
    }
    *reinterpret_cast<float4*>(out + (long)i * 4) = v;
  }
}

__global__ void assemble_bwd_v4_kernel(AsmArgs a, const float* __restrict__ dout, int s, unsigned total) {
  const int Hp = a.H + 2 * a.pad, Wp = a.W + 2 * a.pad;
  const int u = a.up[s] ? 2 : 1;
  const unsigned Hs = a.H / u, Ws = a.W / u, Cq = a.C[s] >> 2;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const unsigned p = i / Cq, c = (i - p * Cq) << 2;
    const unsigned row = p / Ws, ws = p - row * Ws, b = row / Hs, hs = row - b * Hs;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int dy = 0; dy < u; ++dy) {
      const int h = (int)hs * u + dy;
      int hps[3], nh = 0;
      hps[nh++] = h + a.pad;
      if (a.pad) { if (h == 1) hps[nh++] = 0; if (h == a.H - 2) hps[nh++] = Hp - 1; }
      for (int dx = 0; dx < u; ++dx) {
        const int w = (int)ws * u + dx;
        int wps[3], nw = 0;
        wps[nw++] = w + a.pad;
        if (a.pad) { if (w == 1) wps[nw++] = 0; if (w == a.W - 2) wps[nw++] = Wp - 1; }
        for (int ii = 0; ii < nh; ++ii)
          for (int jj = 0; jj < nw; ++jj) {
            const float4 v = *reinterpret_cast<const float4*>(
                dout + (((long)b * Hp + hps[ii]) * Wp + wps[jj]) * a.Ctot + a.off[s] + c);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
          }
      }
    }
    *reinterpret_cast<float4*>(a.d[s] + (long)i * 4) = acc;
    if (a.d2[s]) *reinterpret_cast<float4*>(a.d2[s] + (long)i * 4) = acc;
  }
}

__global__ void add_kernel(const float* __restrict__ a, const float* __restrict__ b,
                           float* __restrict__ o, long n) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
    o[i] = a[i] + b[i];
}

__global__ void add_relu_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                float* __restrict__ o, long n4) {
  const float4* a4 = reinterpret_cast<const float4*>(a);
  const float4* b4 = reinterpret_cast<const float4*>(b);
  float4* o4 = reinterpret_cast<float4*>(o);
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
    float4 x = a4[i], y = b4[i];
    o4[i] = make_float4(fmaxf(x.x + y.x, 0.f), fmaxf(x.y + y.y, 0.f), fmaxf(x.z + y.z, 0.f), fmaxf(x.w + y.w, 0.f));
  }
}

__global__ void mean_hw_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int HW, int C,
                                   float scale) {
  const int b = blockIdx.y, c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float acc = 0.f;
  for (int p = 0; p < HW; ++p) acc += x[((long)b * HW + p) * C + c];
  y[b * C + c] = scale * (acc / (float)HW);
}
__global__ void mean_hw_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx, int HW, int C,
                                   float scale) {
  const int b = blockIdx.y;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < (long)HW * C;
       i += (long)gridDim.x * blockDim.x)
    dx[(long)b * HW * C + i] = dy[b * C + (i % C)] * (scale / (float)HW);
}

__global__ void weight_transpose_kernel(const float* __restrict__ w, float* __restrict__ wt, int Cout,
                                        int taps, int Cin) {
  // w [Cout][taps][Cin] -> wt [Cin][taps][Cout]
  long n = (long)Cout * taps * Cin;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    int co = i % Cout;
    int tp = (i / Cout) % taps;
    int ci = i / ((long)Cout * taps);
    wt[i] = w[((long)co * taps + tp) * Cin + ci];
  }
}

// state[0] = step (int bits), state[1] = lr / (1 - beta1^t), state[2] = sqrt(1 - beta2^t);
// kept on the device so a captured CUDA graph advances the step on every replay.  lr < 0: the learning
// rate is the float in state[3] (written by the host between replays: StepLR, trainer.py:131-132, 266).
__global__ void adam_scalars_kernel(int* state, float lr, float beta1, float beta2) {
  int t = state[0] + 1;
  state[0] = t;
  if (lr < 0.f) lr = ((float*)state)[3];
  double bc1 = 1.0 - pow((double)beta1, (double)t);
  double bc2 = 1.0 - pow((double)beta2, (double)t);
  ((float*)state)[1] = (float)((double)lr / bc1);
  ((float*)state)[2] = (float)sqrt(bc2);
}

__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, long n, const int* __restrict__ state, float beta1,
                            float beta2, float eps, float grad_scale) {
  const float step_size = ((const float*)state)[1], bc2_sqrt = ((const float*)state)[2];
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    float gi = g[i] * grad_scale;
    float mi = m[i] * beta1 + gi * (1.f - beta1);
    float vi = v[i] * beta2 + gi * gi * (1.f - beta2);
    m[i] = mi;
    v[i] = vi;
    float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = p[i] - step_size * (mi / denom);
  }
}

inline int pix_chunks(long M) {
  long c = (M + 8 * 64 - 1) / (8 * 64);
  if (c < 1) c = 1;
  if (c > 592) c = 592;          // 4 waves of 148 SMs
  return (int)c;
}

}  // namespace

extern "C" {

int fd_prep_input(const float* x, float* y, int B, int C, int H, int W, float mean, float stdv,
                  void* stream) {
  prep_input_kernel<<<dim3(fd::cdiv((long)H * W, 256), B), 256, 0, (cudaStream_t)stream>>>(
      x, y, C, H * W, mean, stdv);
  FD_CHECK_LAUNCH();
  return 0;
}

int fd_stem_im2col(const float* x_nchw, float* A, int B, int C, int H, int W, int KH, int KW, int stride,
                   int pad, int Kpad, float mean, float stdv, void* stream) {
  FD_REQUIRE(Kpad >= KH * KW * C && Kpad % 4 == 0, "fd_stem_im2col: bad Kpad %d", Kpad);
  int Ho = (H + 2 * pad - KH) / stride + 1, Wo = (W + 2 * pad - KW) / stride + 1;
  const size_t smem = (size_t)C * KH * (W + 2 * pad) * sizeof(float);
  FD_REQUIRE(Kpad <= 1024 && smem <= 220 * 1024, "fd_stem_im2col: Kpad %d / %zu B of row staging not supported",
             Kpad, smem);
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(stem_im2col_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    FD_REQUIRE(e == cudaSuccess, "fd_stem_im2col: cannot reserve shared memory: %s", cudaGetErrorString(e));
    configured = true;
  }
  stem_im2col_kernel<<<dim3(Ho, B), 256, smem, (cudaStream_t)stream>>>(x_nchw, A, C, H, W, Ho, Wo, KH, KW, stride,
                                                                     pad, Kpad, mean, stdv);
  FD_CHECK_LAUNCH();
  return 0;
}

int fd_pad_rows(const float* src, float* dst, int rows, int cols_src, int cols_dst, int accumulate,
                void* stream) {
  long n = (long)rows * cols_dst;
  pad_rows_kernel<<<min(fd::cdiv(n, 256), 148 * 8), 256, 0, (cudaStream_t)stream>>>(src, dst, rows, cols_src,
                                                                                  cols_dst, accumulate);
  FD_CHECK_LAUNCH();
  return 0;
}

int fd_act_bwd(const float* y, const float* dy, float* dpre, float* dbias, long M, int C, int act,
               void* stream) {
  act_bwd_kernel<<<dim3(fd::cdiv(C, 32), pix_chunks(M)), dim3(32, 8), 0, (cudaStream_t)stream>>>(
      y, dy, dpre, dbias, M, C, act);
  FD_CHECK_LAUNCH();
  return 0;
}

size_t fd_bn_workspace_bytes(int C) {
  return sizeof(double) * (size_t)(1 + BN_MAX_PARTS) * 2 * C + sizeof(unsigned) * (size_t)(C + 32);
}

int fd_bn_fwd(const float* x, const float* residual, const float* gamma, const float* beta,
              float* running_mean, float* running_var, int training, float momentum, float eps,
              int relu, float* y, float* save_mean, float* save_rstd, double* ws, long M, int C,
              float stat_weight, int stats_ready, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (!stats_ready && bn_fused_ok(M, C)) {
    bn_fwd_fused_kernel<<<C / BNF_CG, 256, BNF_SMEM, st>>>(x, residual, gamma, beta, running_mean, running_var,
                                                          training, momentum, eps, stat_weight, save_mean,
                                                          save_rstd, relu, y, (int)M, C);
    FD_CHECK_LAUNCH();
    return 0;
  }
  const int vec = (C % 4 == 0) ? 4 : 1;
  BnGeom g = bn_geom(M, C, vec);
  const size_t sm = sizeof(double) * 2 * vec * 256;
  if (training && !stats_ready) {       // stats_ready: ws already holds the sums (fd_conv2d_fwd_tc_stats)
    BnGeom gr = bn_geom(M, C, vec, true);
    cudaMemsetAsync(ws + (size_t)(1 + BN_MAX_PARTS) * 2 * C, 0, sizeof(unsigned) * gr.grid.x, st);
    if (vec == 4) bn_stats_kernel<4><<<gr.grid, gr.block, sm, st>>>(x, ws, M, C);
    else bn_stats_kernel<1><<<gr.grid, gr.block, sm, st>>>(x, ws, M, C);
    FD_CHECK_LAUNCH();
  }
  if (vec == 4)
    bn_apply_kernel<4><<<g.grid, g.block, 0, st>>>(x, residual, gamma, beta, ws, running_mean, running_var,
                                                   training, momentum, eps, stat_weight, save_mean, save_rstd, relu, y,
                                                   M, C);
  else
    bn_apply_kernel<1><<<g.grid, g.block, 0, st>>>(x, residual, gamma, beta, ws, running_mean, running_var,
                                                   training, momentum, eps, stat_weight, save_mean, save_rstd, relu, y,
                                                   M, C);
  FD_CHECK_LAUNCH();
  return 0;
}

static int bn_bwd_launch(const float* x, const float* y, const float* dy, const float* gamma, const float* mbeta,
                         const float* save_mean, const float* save_rstd, int relu, int training, float* dx,
                         float* dresidual, float* dgamma, float* dbeta, double* ws, long M, int C,
                         int accumulate, cudaStream_t st) {
  const int vec = (C % 4 == 0) ? 4 : 1;
  BnGeom g = bn_geom(M, C, vec);
  const size_t sm = sizeof(double) * 2 * vec * 256;
  BnGeom gr = bn_geom(M, C, vec, true);
  cudaMemsetAsync(ws + (size_t)(1 + BN_MAX_PARTS) * 2 * C, 0, sizeof(unsigned) * gr.grid.x, st);
  if (vec == 4) {
    bn_bwd_reduce_kernel<4><<<gr.grid, gr.block, sm, st>>>(x, y, dy, save_mean, save_rstd, relu, ws, M, C, gamma,
                                                           mbeta);
    FD_CHECK_LAUNCH();
    bn_bwd_apply_kernel<4><<<g.grid, g.block, 0, st>>>(x, y, dy, gamma, save_mean, save_rstd, relu, training,
                                                       ws, dx, dresidual, dgamma, dbeta, M, C, accumulate, mbeta);
  } else {
    bn_bwd_reduce_kernel<1><<<gr.grid, gr.block, sm, st>>>(x, y, dy, save_mean, save_rstd, relu, ws, M, C, gamma,
                                                           mbeta);
    FD_CHECK_LAUNCH();
    bn_bwd_apply_kernel<1><<<g.grid, g.block, 0, st>>>(x, y, dy, gamma, save_mean, save_rstd, relu, training,
                                                       ws, dx, dresidual, dgamma, dbeta, M, C, accumulate, mbeta);
  }
  FD_CHECK_LAUNCH();
  return 0;
}

int fd_bn_bwd(const float* x, const float* y, const float* dy, const float* gamma,
              const float* save_mean, const float* save_rstd, int relu, int training, float* dx,
              float* dresidual, float* dgamma, float* dbeta, double* ws, long M, int C,
              int accumulate, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  FD_REQUIRE(y != nullptr || !relu, "fd_bn_bwd: y is needed for the ReLU mask (or use fd_bn_bwd_xmask)");
  if (bn_fused_ok(M, C)) {
    bn_bwd_fused_kernel<<<C / BNF_CG, 256, BNF_SMEM, st>>>(x, y, dy, gamma, save_mean, save_rstd, relu, training,
                                                          dx, dresidual, dgamma, dbeta, (int)M, C, accumulate);
    FD_CHECK_LAUNCH();
    return 0;
  }
  return bn_bwd_launch(x, y, dy, gamma, nullptr, save_mean, save_rstd, relu, training, dx, dresidual, dgamma, dbeta,
                       ws, M, C, accumulate, st);
}

int fd_bn_bwd_xmask_ok(long M, int C) { return bn_fused_ok(M, C) ? 0 : 1; }

int fd_bn_bwd_xmask(const float* x, const float* dy, const float* gamma, const float* beta,
                    const float* save_mean, const float* save_rstd, int training, float* dx, float* dgamma,
                    float* dbeta, double* ws, long M, int C, int accumulate, void* stream) {
  FD_REQUIRE(!bn_fused_ok(M, C), "fd_bn_bwd_xmask: small tensors go through fd_bn_bwd (fd_bn_bwd_xmask_ok)");
  FD_REQUIRE(gamma != nullptr && beta != nullptr, "fd_bn_bwd_xmask: gamma and beta are needed for the mask");
  return bn_bwd_launch(x, nullptr, dy, gamma, beta, save_mean, save_rstd, 1, training, dx, nullptr, dgamma, dbeta,
                       ws, M, C, accumulate, (cudaStream_t)stream);
}

int fd_maxpool3x3s2_fwd(const float* x, float* y, unsigned char* idx, int B, int H, int W, int C,
                        void* stream) {
  int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  if (C % 4 == 0 && (long)B * H * W * C < (1L << 32) && (((uintptr_t)x | (uintptr_t)y | (uintptr_t)idx) & 15) == 0) {
    const unsigned total = (unsigned)((long)B * Ho * Wo * (C / 4));
    const unsigned blocks = (unsigned)min((long)fd::cdiv((long)total, 256), 148L * 16);
    maxpool_fwd_v4_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(x, y, idx, H, W, C, Ho, Wo, total);
    FD_CHECK_LAUNCH();
    return 0;
  }
  dim3 grid(fd::cdiv(C, 32), min(fd::cdiv((long)Ho * Wo, 8), 1024), B);
  maxpool_fwd_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(x, y, idx, H, W, C, Ho, Wo);
  FD_CHECK_LAUNCH();
  return 0;
}
int fd_maxpool3x3s2_bwd(const float* dy, const unsigned char* idx, float* dx, int B, int H, int W,
                        int C, void* stream) {
  int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  if (C % 4 == 0 && (long)B * H * W * C < (1L << 32) && (((uintptr_t)dy | (uintptr_t)dx | (uintptr_t)idx) & 15) == 0) {
    const unsigned total = (unsigned)((long)B * H * W * (C / 4));
    const unsigned blocks = (unsigned)min((long)fd::cdiv((long)total, 256), 148L * 16);
    maxpool_bwd_v4_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(dy, idx, dx, H, W, C, Ho, Wo, total);
    FD_CHECK_LAUNCH();
    return 0;
  }
  dim3 grid(fd::cdiv(C, 32), min(fd::cdiv((long)H * W, 8), 1024), B);
  maxpool_bwd_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(dy, idx, dx, H, W, C, Ho, Wo);
  FD_CHECK_LAUNCH();
  return 0;
}

static int fill_asm(AsmArgs& a, int nseg, const int* C, const int* up, int B, int H, int W, int pad) {
  FD_REQUIRE(nseg >= 1 && nseg <= MAXSEG, "fd_assemble: 1..%d segments supported", MAXSEG);
  FD_REQUIRE(pad == 0 || pad == 1, "fd_assemble: pad must be 0 or 1");
  a.nseg = nseg; a.B = B; a.H = H; a.W = W; a.pad = pad;
  int off = 0;
  for (int i = 0; i < MAXSEG; ++i) {
    a.a[i] = a.b[i] = nullptr; a.d[i] = a.d2[i] = nullptr;
    a.C[i] = i < nseg ? C[i] : 0;
    a.up[i] = i < nseg ? up[i] : 0;
    a.off[i] = off;
    off += a.C[i];
    if (i < nseg && up[i]) FD_REQUIRE(H % 2 == 0 && W % 2 == 0, "fd_assemble: odd size with upsample");
  }
  a.Ctot = off;
  return 0;
}

int fd_assemble_fwd(const fd_segment* segs, int nseg, float* out, int B, int H, int W, int pad,
                    void* stream) {
  AsmArgs a;
  int C[MAXSEG], up[MAXSEG];
  for (int i = 0; i < nseg && i < MAXSEG; ++i) { C[i] = segs[i].C; up[i] = segs[i].up; }
  int rc = fill_asm(a, nseg, C, up, B, H, W, pad);
  if (rc) return rc;
  for (int i = 0; i < nseg; ++i) { a.a[i] = segs[i].a; a.b[i] = segs[i].b; }
  long npix = (long)B * (H + 2 * pad) * (W + 2 * pad);
  bool v4 = npix * a.Ctot < (1L << 32) && (((uintptr_t)out) & 15) == 0;
  for (int i = 0; i < nseg; ++i)
    v4 = v4 && a.C[i] % 4 == 0 && (((uintptr_t)a.a[i] | (uintptr_t)a.b[i]) & 15) == 0;
  if (v4) {
    const unsigned total = (unsigned)(npix * (a.Ctot / 4));
    const unsigned blocks = (unsigned)min((long)fd::cdiv((long)total, 256), 148L * 16);
    assemble_fwd_v4_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(a, out, total);
    FD_CHECK_LAUNCH();
    return 0;
  }
  int threads = a.Ctot >= 256 ? 256 : (a.Ctot >= 128 ? 128 : (a.Ctot >= 64 ? 64 : 32));
  long blocks = npix < 148L * 32 ? npix : 148L * 32;
  assemble_fwd_kernel<<<(int)blocks, threads, 0, (cudaStream_t)stream>>>(a, out);
  FD_CHECK_LAUNCH();
  return 0;
}

int fd_assemble_bwd(const float* dout, float* const* dsegs, const int* C, const int* up, int nseg,
                    int B, int H, int W, int pad, void* stream) {
  return fd_assemble_bwd2(dout, dsegs, nullptr, C, up, nseg, B, H, W, pad, stream);
}

int fd_assemble_bwd2(const float* dout, float* const* dsegs, float* const* dsegs2, const int* C, const int* up,
                     int nseg, int B, int H, int W, int pad, void* stream) {
  AsmArgs a;
  int rc = fill_asm(a, nseg, C, up, B, H, W, pad);
  if (rc) return rc;
  for (int s = 0; s < nseg; ++s) {
    if (!dsegs[s]) {
      FD_REQUIRE(!(dsegs2 && dsegs2[s]), "fd_assemble_bwd2: dsegs2[%d] without dsegs[%d]", s, s);
      continue;
    }
    a.d[s] = dsegs[s];
    a.d2[s] = dsegs2 ? dsegs2[s] : nullptr;
    int u = up[s] ? 2 : 1;
    long npix = (long)B * (H / u) * (W / u);
    bool v4 = npix * C[s] < (1L << 32) && a.Ctot % 4 == 0 &&
              (((uintptr_t)dout | (uintptr_t)dsegs[s] | (uintptr_t)a.d2[s]) & 15) == 0;
    for (int i = 0; i <= s; ++i) v4 = v4 && C[i] % 4 == 0;      // this segment and its channel offset
    if (v4) {
      const unsigned total = (unsigned)(npix * (C[s] / 4));
      const unsigned blocks = (unsigned)min((long)fd::cdiv((long)total, 256), 148L * 16);
      assemble_bwd_v4_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(a, dout, s, total);
      FD_CHECK_LAUNCH();
      continue;
    }
    int threads = C[s] >= 256 ? 256 : (C[s] >= 128 ? 128 : (C[s] >= 64 ? 64 : 32));
    long blocks = npix < 148L * 32 ? npix : 148L * 32;
    assemble_bwd_kernel<<<(int)blocks, threads, 0, (cudaStream_t)stream>>>(a, dout, s);
    FD_CHECK_LAUNCH();
  }
  return 0;
}

int fd_add(const float* a, const float* b, float* out, long n, void* stream) {
  add_kernel<<<min(fd::cdiv(n, 256), 148 * 16), 256, 0, (cudaStream_t)stream>>>(a, b, out, n);
  FD_CHECK_LAUNCH();
  return 0;
}

int fd_add_relu(const float* a, const float* b, float* out, long n, void* stream) {
  FD_REQUIRE(n % 4 == 0 && (((uintptr_t)a | (uintptr_t)b | (uintptr_t)out) & 15) == 0,
             "fd_add_relu: needs 16-byte aligned tensors with a multiple of 4 elements");
  add_relu_kernel<<<min(fd::cdiv(n / 4, 256), 148 * 16), 256, 0, (cudaStream_t)stream>>>(a, b, out, n / 4);
  FD_CHECK_LAUNCH();
  return 0;
}

int fd_mean_hw_fwd(const float* x, float* y, int B, int HW, int C, float scale, void* stream) {
  mean_hw_fwd_kernel<<<dim3(fd::cdiv(C, 64), B), 64, 0, (cudaStream_t)stream>>>(x, y, HW, C, scale);
  FD_CHECK_LAUNCH();
  return 0;
}
int fd_mean_hw_bwd(const float* dy, float* dx, int B, int HW, int C, float scale, void* stream) {
  mean_hw_bwd_kernel<<<dim3(min(fd::cdiv((long)HW * C, 256), 1024), B), 256, 0, (cudaStream_t)stream>>>(
      dy, dx, HW, C, scale);
  FD_CHECK_LAUNCH();
  return 0;
}

int fd_weight_transpose(const float* w, float* wt, int Cout, int taps, int Cin, void* stream) {
  long n = (long)Cout * taps * Cin;
  weight_transpose_kernel<<<min(fd::cdiv(n, 256), 148 * 8), 256, 0, (cudaStream_t)stream>>>(w, wt, Cout,
                                                                                          taps, Cin);
  FD_CHECK_LAUNCH();
  return 0;
}

int fd_adam_step(float* p, const float* g, float* m, float* v, long n, float lr, float beta1,
                 float beta2, float eps, int* state, float grad_scale, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  adam_scalars_kernel<<<1, 1, 0, st>>>(state, lr, beta1, beta2);
  FD_CHECK_LAUNCH();
  adam_kernel<<<min(fd::cdiv(n, 256), 148 * 16), 256, 0, st>>>(p, g, m, v, n, state, beta1, beta2, eps,
                                                              grad_scale);
  FD_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
