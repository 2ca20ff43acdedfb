// Unfused drop-in geometry / loss operators: what the reference's layers.* modules and the F.* calls of the
// UNCHANGED drivers resolve to when the fused fd_photoloss_* path is not patched in.
//
//   fd_upsample_bilinear_{fwd,bwd}   F.interpolate(mode="bilinear", align_corners=False)
//                                    (trainer.py:434-435, 579; refiner.py:325, 335, 680; evaluate_depth.py:207-218)
//   fd_backproject_{fwd,bwd}         layers.BackprojectDepth.forward          (layers.py:133-162)
//   fd_project3d_{fwd,bwd}           layers.Project3D.forward                 (layers.py:204-226)
//   fd_grid_sample_border_{fwd,bwd}  F.grid_sample(padding_mode="border")     (trainer.py:467-470)
//   fd_ssim_{fwd,bwd}                layers.SSIM.forward                      (layers.py:251-281)
//
// All tensors fp32 NCHW contiguous, as the reference's modules produce them.  These are memory-bound
// elementwise / stencil / gather kernels: one thread per output element, coalesced along W.
#include "common.cuh"
#include "../../include/fusiondepth_b200.h"

namespace {

constexpr int NT = 256;

struct Lerp {
  int i0, i1;
  float l0, l1;
};
// area_pixel_compute_source_index(scale = in/out, align_corners=False)
__device__ __forceinline__ Lerp lerp_coords(int dst, int in, int out) {
  Lerp r;
  if (in == out) { r.i0 = r.i1 = dst; r.l0 = 1.f; r.l1 = 0.f; return r; }
  float scale = (float)in / (float)out;
  float f = scale * ((float)dst + 0.5f) - 0.5f;
  if (f < 0.f) f = 0.f;
  r.i0 = min((int)f, in - 1);
  r.i1 = r.i0 + (r.i0 < in - 1 ? 1 : 0);
  r.l1 = f - (float)r.i0;
  r.l0 = 1.f - r.l1;
  return r;
}

__global__ void upsample_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, long planes, int h, int w,
                                    int H, int W) {
  const long n = planes * H * W;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    int X = (int)(i % W), Y = (int)((i / W) % H);
    long p = i / ((long)W * H);
    Lerp ly = lerp_coords(Y, h, H), lx = lerp_coords(X, w, W);
    const float* s = x + p * h * w;
    float t0 = lx.l0 * s[(long)ly.i0 * w + lx.i0] + lx.l1 * s[(long)ly.i0 * w + lx.i1];
    float t1 = lx.l0 * s[(long)ly.i1 * w + lx.i0] + lx.l1 * s[(long)ly.i1 * w + lx.i1];
    y[i] = ly.l0 * t0 + ly.l1 * t1;
  }
}

__global__ void upsample_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx, long planes, int h, int w,
                                    int H, int W) {
  const long n = planes * H * W;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    int X = (int)(i % W), Y = (int)((i / W) % H);
    long p = i / ((long)W * H);
    Lerp ly = lerp_coords(Y, h, H), lx = lerp_coords(X, w, W);
    float* d = dx + p * h * w;
    float g = dy[i];
    atomicAdd(d + (long)ly.i0 * w + lx.i0, g * ly.l0 * lx.l0);
    atomicAdd(d + (long)ly.i0 * w + lx.i1, g * ly.l0 * lx.l1);
    atomicAdd(d + (long)ly.i1 * w + lx.i0, g * ly.l1 * lx.l0);
    atomicAdd(d + (long)ly.i1 * w + lx.i1, g * ly.l1 * lx.l1);
  }
}

// ---- BackprojectDepth: cam = depth * (inv_K[:3,:3] @ [x,y,1]), row 3 = 1 -----------------------------
__device__ __forceinline__ void pixel_ray(const float* __restrict__ ik, int x, int y, float ray[3]) {
  const float fx = (float)x, fy = (float)y;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    float c = ik[k * 4 + 0] * fx;
    c = fmaf(ik[k * 4 + 1], fy, c);
    c = fmaf(ik[k * 4 + 2], 1.0f, c);
    ray[k] = c;
  }
}

__global__ void backproject_fwd_kernel(const float* __restrict__ depth, const float* __restrict__ invK,
                                       float* __restrict__ cam, int B, int H, int W) {
  const long HW = (long)H * W, n = B * HW;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    int b = (int)(i / HW);
    long r = i - b * HW;
    float ray[3];
    pixel_ray(invK + b * 16, (int)(r % W), (int)(r / W), ray);
    float d = depth[i];
    float* o = cam + (long)b * 4 * HW + r;
    o[0] = __fmul_rn(d, ray[0]); o[HW] = __fmul_rn(d, ray[1]); o[2 * HW] = __fmul_rn(d, ray[2]); o[3 * HW] = 1.0f;
  }
}

__global__ void backproject_bwd_kernel(const float* __restrict__ dcam, const float* __restrict__ invK,
                                       float* __restrict__ ddepth, int B, int H, int W) {
  const long HW = (long)H * W, n = B * HW;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    int b = (int)(i / HW);
    long r = i - b * HW;
    float ray[3];
    pixel_ray(invK + b * 16, (int)(r % W), (int)(r / W), ray);
    const float* g = dcam + (long)b * 4 * HW + r;
    ddepth[i] = g[0] * ray[0] + g[HW] * ray[1] + g[2 * HW] * ray[2];
  }
}

// ---- Project3D: P = (K @ T)[:3]; c = P @ pts; uv = c[:2] / (c[2] + eps); normalise by (W-1), (H-1) ----
__device__ __forceinline__ void make_P(const float* __restrict__ K, const float* __restrict__ T, float P[12]) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float acc = K[i * 4 + 0] * T[0 * 4 + j];
      acc = fmaf(K[i * 4 + 1], T[1 * 4 + j], acc);
      acc = fmaf(K[i * 4 + 2], T[2 * 4 + j], acc);
      acc = fmaf(K[i * 4 + 3], T[3 * 4 + j], acc);
      P[i * 4 + j] = acc;
    }
}

__global__ void project_fwd_kernel(const float* __restrict__ pts, const float* __restrict__ K,
                                   const float* __restrict__ T, float* __restrict__ grid, int B, int H, int W,
                                   float eps) {
  __shared__ float P[12];
  const int b = blockIdx.y;
  const long HW = (long)H * W;
  if (threadIdx.x == 0) make_P(K + b * 16, T + b * 16, P);
  __syncthreads();
  const float* p = pts + (long)b * 4 * HW;
  for (long r = blockIdx.x * (long)blockDim.x + threadIdx.x; r < HW; r += (long)gridDim.x * blockDim.x) {
    float X0 = p[r], X1 = p[HW + r], X2 = p[2 * HW + r], X3 = p[3 * HW + r];
    float c[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      float v = P[i * 4 + 0] * X0;
      v = fmaf(P[i * 4 + 1], X1, v);
      v = fmaf(P[i * 4 + 2], X2, v);
      v = fmaf(P[i * 4 + 3], X3, v);
      c[i] = v;
    }
    float z = __fadd_rn(c[2], eps);
    float u = __fdiv_rn(__fdiv_rn(c[0], z), (float)(W - 1));
    float v = __fdiv_rn(__fdiv_rn(c[1], z), (float)(H - 1));
    float2 o = make_float2(__fmul_rn(__fadd_rn(u, -0.5f), 2.f), __fmul_rn(__fadd_rn(v, -0.5f), 2.f));
    reinterpret_cast<float2*>(grid)[(long)b * HW + r] = o;
  }
}

// dpts [B,4,HW]; dPacc [B][12] (+=, zero-initialised): dP[i][j] = sum_pixels dc_i * pts_j
__global__ void project_bwd_kernel(const float* __restrict__ pts, const float* __restrict__ K,
                                   const float* __restrict__ T, const float* __restrict__ dgrid,
                                   float* __restrict__ dpts, float* __restrict__ dPacc, int B, int H, int W,
                                   float eps) {
  __shared__ float P[12];
  __shared__ float red[12 * 32];
  const int b = blockIdx.y;
  const long HW = (long)H * W;
  if (threadIdx.x == 0) make_P(K + b * 16, T + b * 16, P);
  __syncthreads();
  const float* p = pts + (long)b * 4 * HW;
  float acc[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) acc[i] = 0.f;
  for (long r = blockIdx.x * (long)blockDim.x + threadIdx.x; r < HW; r += (long)gridDim.x * blockDim.x) {
    float X[4] = {p[r], p[HW + r], p[2 * HW + r], p[3 * HW + r]};
    float c[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      float v = P[i * 4 + 0] * X[0];
      v = fmaf(P[i * 4 + 1], X[1], v);
      v = fmaf(P[i * 4 + 2], X[2], v);
      v = fmaf(P[i * 4 + 3], X[3], v);
      c[i] = v;
    }
    float z = c[2] + eps, iz = 1.f / z;
    float2 g = reinterpret_cast<const float2*>(dgrid)[(long)b * HW + r];
    float gu = g.x * 2.f / (float)(W - 1), gv = g.y * 2.f / (float)(H - 1);
    float gc[3] = {gu * iz, gv * iz, -(gu * c[0] + gv * c[1]) * iz * iz};
    if (dpts) {
      float* d = dpts + (long)b * 4 * HW + r;
#pragma unroll
      for (int j = 0; j < 4; ++j) d[j * HW] = gc[0] * P[j] + gc[1] * P[4 + j] + gc[2] * P[8 + j];
    }
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i * 4 + j] += gc[i] * X[j];
  }
  fd::block_sum<12>(acc, red);
  if (threadIdx.x == 0 && dPacc) {
#pragma unroll
    for (int i = 0; i < 12; ++i) atomicAdd(dPacc + b * 12 + i, acc[i]);
  }
}

// dT = K[:3,:]^T dP  (P = K T  =>  dT[k][j] = sum_i K[i][k] dP[i][j])
__global__ void project_dT_kernel(const float* __restrict__ K, const float* __restrict__ dPacc, float* __restrict__ dT,
                                  int B) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= B * 16) return;
  int b = t / 16, e = t % 16, k = e / 4, j = e % 4;
  float v = 0.f;
  for (int i = 0; i < 3; ++i) v += K[b * 16 + i * 4 + k] * dPacc[b * 12 + i * 4 + j];
  dT[t] = v;
}

// ---- grid_sample: bilinear, padding_mode="border", align_corners=False --------------------------------
struct Samp {
  int x0, y0;
  float fx, fy, mx, my;
};
__device__ __forceinline__ Samp sample_coords(float gx, float gy, int H, int W) {
  Samp s;
  float ix = __fdiv_rn(__fadd_rn(__fmul_rn(__fadd_rn(gx, 1.f), (float)W), -1.f), 2.f);
  float iy = __fdiv_rn(__fadd_rn(__fmul_rn(__fadd_rn(gy, 1.f), (float)H), -1.f), 2.f);
  float wm = (float)(W - 1), hm = (float)(H - 1);
  // clip_coordinates_set_grad: the gradient is zero where the clip is active
  s.mx = (ix <= 0.f || ix >= wm) ? 0.f : 1.f;
  s.my = (iy <= 0.f || iy >= hm) ? 0.f : 1.f;
  ix = fminf(wm, fmaxf(ix, 0.f));
  iy = fminf(hm, fmaxf(iy, 0.f));
  if (!(ix == ix)) { ix = 0.f; s.mx = 0.f; }
  if (!(iy == iy)) { iy = 0.f; s.my = 0.f; }
  float flx = floorf(ix), fly = floorf(iy);
  s.x0 = (int)flx; s.y0 = (int)fly;
  s.fx = ix - flx; s.fy = iy - fly;
  return s;
}

__global__ void grid_sample_fwd_kernel(const float* __restrict__ img, const float* __restrict__ grid,
                                       float* __restrict__ out, int B, int C, int H, int W, int Ho, int Wo) {
  const long HWo = (long)Ho * Wo, HW = (long)H * W, n = B * HWo;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    int b = (int)(i / HWo);
    long r = i - b * HWo;
    float2 g = reinterpret_cast<const float2*>(grid)[i];
    Samp s = sample_coords(g.x, g.y, H, W);
    int x1 = s.x0 + 1, y1 = s.y0 + 1;
    bool bx1 = x1 <= W - 1, by1 = y1 <= H - 1;
    float wx0 = 1.f - s.fx, wy0 = 1.f - s.fy;
    float nw = wx0 * wy0, ne = s.fx * wy0, sw = wx0 * s.fy, se = s.fx * s.fy;
    for (int c = 0; c < C; ++c) {
      const float* p = img + ((long)b * C + c) * HW;
      float acc = p[(long)s.y0 * W + s.x0] * nw;
      if (bx1) acc += p[(long)s.y0 * W + x1] * ne;
      if (by1) acc += p[(long)y1 * W + s.x0] * sw;
      if (bx1 && by1) acc += p[(long)y1 * W + x1] * se;
      out[((long)b * C + c) * HWo + r] = acc;
    }
  }
}

__global__ void grid_sample_bwd_kernel(const float* __restrict__ img, const float* __restrict__ grid,
                                       const float* __restrict__ dout, float* __restrict__ dgrid,
                                       float* __restrict__ dimg, int B, int C, int H, int W, int Ho, int Wo) {
  const long HWo = (long)Ho * Wo, HW = (long)H * W, n = B * HWo;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    int b = (int)(i / HWo);
    long r = i - b * HWo;
    float2 g = reinterpret_cast<const float2*>(grid)[i];
    Samp s = sample_coords(g.x, g.y, H, W);
    int x1 = s.x0 + 1, y1 = s.y0 + 1;
    bool bx1 = x1 <= W - 1, by1 = y1 <= H - 1;
    float wx0 = 1.f - s.fx, wy0 = 1.f - s.fy;
    float gix = 0.f, giy = 0.f;
    for (int c = 0; c < C; ++c) {
      const float* p = img + ((long)b * C + c) * HW;
      float go = dout[((long)b * C + c) * HWo + r];
      float vnw = p[(long)s.y0 * W + s.x0];
      float vne = bx1 ? p[(long)s.y0 * W + x1] : 0.f;
      float vsw = by1 ? p[(long)y1 * W + s.x0] : 0.f;
      float vse = (bx1 && by1) ? p[(long)y1 * W + x1] : 0.f;
      gix += go * ((vne - vnw) * wy0 + (vse - vsw) * s.fy);
      giy += go * ((vsw - vnw) * wx0 + (vse - vne) * s.fx);
      if (dimg) {
        float* d = dimg + ((long)b * C + c) * HW;
        atomicAdd(d + (long)s.y0 * W + s.x0, go * wx0 * wy0);
        if (bx1) atomicAdd(d + (long)s.y0 * W + x1, go * s.fx * wy0);
        if (by1) atomicAdd(d + (long)y1 * W + s.x0, go * wx0 * s.fy);
        if (bx1 && by1) atomicAdd(d + (long)y1 * W + x1, go * s.fx * s.fy);
      }
    }
    if (dgrid) {
      // ix = ((gx + 1) W - 1) / 2  =>  d ix / d gx = W / 2
      reinterpret_cast<float2*>(dgrid)[i] =
          make_float2(gix * s.mx * 0.5f * (float)W, giy * s.my * 0.5f * (float)H);
    }
  }
}

// ---- SSIM (3x3 mean over reflect-padded images) --------------------------------------------------------
constexpr float C1 = 1e-4f, C2 = 9e-4f;
__device__ __forceinline__ int refl(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}

struct Stats {
  float mx, my, vx, vy, vxy;
};
__device__ __forceinline__ Stats window_stats(const float* __restrict__ x, const float* __restrict__ y, int py, int px,
                                              int H, int W) {
  float sx = 0.f, sy = 0.f, sxx = 0.f, syy = 0.f, sxy = 0.f;
#pragma unroll
  for (int dy = -1; dy <= 1; ++dy) {
    const int yy = refl(py + dy, H);
#pragma unroll
    for (int dx = -1; dx <= 1; ++dx) {
      const int xx = refl(px + dx, W);
      float a = x[(long)yy * W + xx], b = y[(long)yy * W + xx];
      sx = __fadd_rn(sx, a); sy = __fadd_rn(sy, b);
      sxx = __fadd_rn(sxx, __fmul_rn(a, a)); syy = __fadd_rn(syy, __fmul_rn(b, b));
      sxy = __fadd_rn(sxy, __fmul_rn(a, b));
    }
  }
  Stats s;
  s.mx = __fdiv_rn(sx, 9.f); s.my = __fdiv_rn(sy, 9.f);
  s.vx = __fadd_rn(__fdiv_rn(sxx, 9.f), -__fmul_rn(s.mx, s.mx));
  s.vy = __fadd_rn(__fdiv_rn(syy, 9.f), -__fmul_rn(s.my, s.my));
  s.vxy = __fadd_rn(__fdiv_rn(sxy, 9.f), -__fmul_rn(s.mx, s.my));
  return s;
}

__global__ void ssim_fwd_kernel(const float* __restrict__ x, const float* __restrict__ y, float* __restrict__ out,
                                long planes, int H, int W) {
  const long HW = (long)H * W, n = planes * HW;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    long p = i / HW, r = i - p * HW;
    Stats s = window_stats(x + p * HW, y + p * HW, (int)(r / W), (int)(r % W), H, W);
    float nn = __fmul_rn(__fadd_rn(__fmul_rn(__fmul_rn(2.f, s.mx), s.my), C1), __fadd_rn(__fmul_rn(2.f, s.vxy), C2));
    float dd = __fmul_rn(__fadd_rn(__fadd_rn(__fmul_rn(s.mx, s.mx), __fmul_rn(s.my, s.my)), C1),
                         __fadd_rn(__fadd_rn(s.vx, s.vy), C2));
    float v = __fdiv_rn(__fadd_rn(1.f, -__fdiv_rn(nn, dd)), 2.f);
    out[i] = fminf(fmaxf(v, 0.f), 1.f);
  }
}

// pass 1: per-pixel adjoint coefficients, d out_q / d x_p = alpha_q + beta_q x_p + gamma_q y_p for p in q's window
__global__ void ssim_coef_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                 const float* __restrict__ dout, float* __restrict__ coef, long planes, int H, int W) {
  const long HW = (long)H * W, n = planes * HW;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    long p = i / HW, r = i - p * HW;
    Stats s = window_stats(x + p * HW, y + p * HW, (int)(r / W), (int)(r % W), H, W);
    float n1 = 2.f * s.mx * s.my + C1, n2 = 2.f * s.vxy + C2;
    float d1 = s.mx * s.mx + s.my * s.my + C1, d2 = s.vx + s.vy + C2;
    float nn = n1 * n2, dd = d1 * d2;
    float S = (1.f - nn / dd) * 0.5f;
    float al = 0.f, be = 0.f, ga = 0.f;
    if (S >= 0.f && S <= 1.f) {
      float g = dout[i];
      float a0 = (2.f / 9.f) * s.my * (n2 - n1), a1 = (2.f / 9.f) * n1;
      float b0 = (2.f / 9.f) * s.mx * (d2 - d1), b1 = (2.f / 9.f) * d1;
      float inv_d = 1.f / dd, nd2 = nn * inv_d * inv_d;
      al = g * 0.5f * (nd2 * b0 - a0 * inv_d);
      be = g * 0.5f * nd2 * b1;
      ga = -g * 0.5f * a1 * inv_d;
    }
    coef[i] = al; coef[n + i] = be; coef[2 * n + i] = ga;
  }
}

// pass 2: dx_p = sum over padded positions e aliasing p, over q in the 3x3 around e: alpha_q + beta_q x_p + gamma_q y_p
__global__ void ssim_gather_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                   const float* __restrict__ coef, float* __restrict__ dx, long planes, int H, int W) {
  const long HW = (long)H * W, n = planes * HW;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    long p = i / HW, r = i - p * HW;
    int py = (int)(r / W), px = (int)(r % W);
    int eys[3], exs[3], ney = 0, nex = 0;
    eys[ney++] = py; if (py == 1) eys[ney++] = -1; if (py == H - 2) eys[ney++] = H;
    exs[nex++] = px; if (px == 1) exs[nex++] = -1; if (px == W - 2) exs[nex++] = W;
    const float* ca = coef + p * HW;
    float A = 0.f, Bc = 0.f, G = 0.f;
    for (int a = 0; a < ney; ++a)
      for (int b = 0; b < nex; ++b)
        for (int dy = -1; dy <= 1; ++dy)
          for (int dxx = -1; dxx <= 1; ++dxx) {
            int qy = eys[a] + dy, qx = exs[b] + dxx;
            if (qy >= 0 && qy < H && qx >= 0 && qx < W) {
              long o = (long)qy * W + qx;
              A += ca[o]; Bc += ca[n + o]; G += ca[2 * n + o];
            }
          }
    dx[i] = A + Bc * x[i] + G * y[i];
  }
}

// ---- transformation_from_parameters (layers.py:23-97): one thread per batch element --------------------
struct Rod {
  float x, y, z, c, s, C, theta, den;
  float R[9];
};
__device__ __forceinline__ Rod rodrigues(const float* __restrict__ v) {
  Rod r;
  r.theta = sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
  r.den = r.theta + 1e-7f;
  r.x = v[0] / r.den; r.y = v[1] / r.den; r.z = v[2] / r.den;
  r.c = cosf(r.theta); r.s = sinf(r.theta); r.C = 1.f - r.c;
  const float xs = r.x * r.s, ys = r.y * r.s, zs = r.z * r.s;
  const float xC = r.x * r.C, yC = r.y * r.C, zC = r.z * r.C;
  const float xyC = r.x * yC, yzC = r.y * zC, zxC = r.z * xC;
  r.R[0] = r.x * xC + r.c; r.R[1] = xyC - zs;       r.R[2] = zxC + ys;
  r.R[3] = xyC + zs;       r.R[4] = r.y * yC + r.c; r.R[5] = yzC - xs;
  r.R[6] = zxC - ys;       r.R[7] = yzC + xs;       r.R[8] = r.z * zC + r.c;
  return r;
}

__global__ void pose_matrix_fwd_kernel(const float* __restrict__ aa, const float* __restrict__ tr, int invert,
                                       float* __restrict__ M, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  Rod r = rodrigues(aa + 3 * b);
  const float* t = tr + 3 * b;
  float* m = M + 16 * b;
  if (!invert) {                      // M = T(t) @ R
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      m[i * 4 + 0] = r.R[i * 3 + 0]; m[i * 4 + 1] = r.R[i * 3 + 1]; m[i * 4 + 2] = r.R[i * 3 + 2];
      m[i * 4 + 3] = t[i];
    }
  } else {                            // M = R^T @ T(-t)
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      m[i * 4 + 0] = r.R[0 * 3 + i]; m[i * 4 + 1] = r.R[1 * 3 + i]; m[i * 4 + 2] = r.R[2 * 3 + i];
      float acc = r.R[0 * 3 + i] * -t[0];
      acc = fmaf(r.R[1 * 3 + i], -t[1], acc);
      acc = fmaf(r.R[2 * 3 + i], -t[2], acc);
      m[i * 4 + 3] = acc;
    }
  }
  m[12] = 0.f; m[13] = 0.f; m[14] = 0.f; m[15] = 1.f;
}

__global__ void pose_matrix_bwd_kernel(const float* __restrict__ aa, const float* __restrict__ tr, int invert,
                                       const float* __restrict__ dM, float* __restrict__ daa,
                                       float* __restrict__ dtr, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float* v = aa + 3 * b;
  const float* t = tr + 3 * b;
  const float* g = dM + 16 * b;
  Rod r = rodrigues(v);
  float dR[9], dt[3];
  if (!invert) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
      for (int j = 0; j < 3; ++j) dR[i * 3 + j] = g[i * 4 + j];
      dt[i] = g[i * 4 + 3];
    }
  } else {
    // M[:3,:3] = R^T ; M[i][3] = -sum_k R[k][i] t[k]
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      float acc = 0.f;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        dR[k * 3 + i] = g[i * 4 + k] - t[k] * g[i * 4 + 3];
        acc += r.R[k * 3 + i] * g[i * 4 + 3];
      }
      dt[k] = -acc;
    }
  }
  const float x = r.x, y = r.y, z = r.z, s = r.s, C = r.C;
  const float xC = x * C, yC = y * C, zC = z * C;
  float dx = dR[0] * 2.f * xC + dR[1] * yC + dR[2] * zC + dR[3] * yC - dR[5] * s + dR[6] * zC + dR[7] * s;
  float dy = dR[1] * xC + dR[2] * s + dR[3] * xC + dR[4] * 2.f * yC + dR[5] * zC - dR[6] * s + dR[7] * zC;
  float dz = -dR[1] * s + dR[2] * xC + dR[3] * s + dR[5] * yC + dR[6] * xC + dR[7] * yC + dR[8] * 2.f * zC;
  float dC = dR[0] * x * x + (dR[1] + dR[3]) * x * y + (dR[2] + dR[6]) * z * x + dR[4] * y * y +
             (dR[5] + dR[7]) * y * z + dR[8] * z * z;
  float dc = dR[0] + dR[4] + dR[8] - dC;
  float ds = -dR[1] * z + dR[2] * y + dR[3] * z - dR[5] * x - dR[6] * y + dR[7] * x;
  float dtheta = -s * dc + r.c * ds;
  // axis = v / (theta + eps)
  dtheta += -(x * dx + y * dy + z * dz) / r.den;
  const float it = r.theta > 0.f ? 1.f / r.theta : 0.f;
  daa[3 * b + 0] = dx / r.den + dtheta * v[0] * it;
  daa[3 * b + 1] = dy / r.den + dtheta * v[1] * it;
  daa[3 * b + 2] = dz / r.den + dtheta * v[2] * it;
  dtr[3 * b + 0] = dt[0]; dtr[3 * b + 1] = dt[1]; dtr[3 * b + 2] = dt[2];
}

// ---- Cat_xy (layers.py:165-201): x/30, y/2, (z-40)/40 of the back-projected points ----------------------
__global__ void cat_xy_kernel(const float* __restrict__ depth, const float* __restrict__ invK, float* __restrict__ out,
                              int B, int H, int W) {
  const long HW = (long)H * W, n = B * HW;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    int b = (int)(i / HW);
    long r = i - b * HW;
    float ray[3];
    pixel_ray(invK + b * 16, (int)(r % W), (int)(r / W), ray);
    float d = depth[i];
    float* o = out + (long)b * 3 * HW + r;
    o[0] = __fdiv_rn(__fmul_rn(d, ray[0]), 30.0f);
    o[HW] = __fdiv_rn(__fmul_rn(d, ray[1]), 2.0f);
    o[2 * HW] = __fdiv_rn(__fadd_rn(__fmul_rn(d, ray[2]), -40.0f), 40.0f);
  }
}

inline int blocks_for(long n) {
  long b = (n + NT - 1) / NT;
  return (int)(b > 148L * 32 ? 148L * 32 : (b < 1 ? 1 : b));
}

}  // namespace

extern "C" {

int fd_upsample_bilinear_fwd(const float* x, float* y, long planes, int h, int w, int H, int W, void* stream) {
  FD_REQUIRE(planes > 0 && h > 0 && w > 0 && H > 0 && W > 0, "fd_upsample_bilinear_fwd: bad shape");
  upsample_fwd_kernel<<<blocks_for(planes * H * W), NT, 0, (cudaStream_t)stream>>>(x, y, planes, h, w, H, W);
  FD_CHECK_LAUNCH();
  return 0;
}

int fd_upsample_bilinear_bwd(const float* dy, float* dx, long planes, int h, int w, int H, int W, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(dx, 0, sizeof(float) * planes * h * w, st);
  FD_REQUIRE(e == cudaSuccess, "fd_upsample_bilinear_bwd: memset failed: %s", cudaGetErrorString(e));
  upsample_bwd_kernel<<<blocks_for(planes * H * W), NT, 0, st>>>(dy, dx, planes, h, w, H, W);
  FD_CHECK_LAUNCH();
  return 0;
}

int fd_backproject_fwd(const float* depth, const float* inv_K, float* cam, int B, int H, int W, void* stream) {
  backproject_fwd_kernel<<<blocks_for((long)B * H * W), NT, 0, (cudaStream_t)stream>>>(depth, inv_K, cam, B, H, W);
  FD_CHECK_LAUNCH();
  return 0;
}

int fd_backproject_bwd(const float* dcam, const float* inv_K, float* ddepth, int B, int H, int W, void* stream) {
  backproject_bwd_kernel<<<blocks_for((long)B * H * W), NT, 0, (cudaStream_t)stream>>>(dcam, inv_K, ddepth, B, H, W);
  FD_CHECK_LAUNCH();
  return 0;
}

int fd_project3d_fwd(const float* points, const float* K, const float* T, float* grid, int B, int H, int W,
                     float eps, void* stream) {
  FD_REQUIRE(H > 1 && W > 1, "fd_project3d_fwd: H, W must exceed 1");
  dim3 g(min(fd::cdiv((long)H * W, NT), 148 * 4), B);
  project_fwd_kernel<<<g, NT, 0, (cudaStream_t)stream>>>(points, K, T, grid, B, H, W, eps);
  FD_CHECK_LAUNCH();
  return 0;
}

// workspace: B*12 floats
int fd_project3d_bwd(const float* points, const float* K, const float* T, const float* dgrid, float* dpoints,
                     float* dT, int B, int H, int W, float eps, void* workspace, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  float* dP = (float*)workspace;
  cudaError_t e = cudaMemsetAsync(dP, 0, sizeof(float) * B * 12, st);
  FD_REQUIRE(e == cudaSuccess, "fd_project3d_bwd: memset failed: %s", cudaGetErrorString(e));
  dim3 g(min(fd::cdiv((long)H * W, NT), 148 * 2), B);
  project_bwd_kernel<<<g, NT, 0, st>>>(points, K, T, dgrid, dpoints, dT ? dP : nullptr, B, H, W, eps);
  FD_CHECK_LAUNCH();
  if (dT) {
    project_dT_kernel<<<fd::cdiv(B * 16, 64), 64, 0, st>>>(K, dP, dT, B);
    FD_CHECK_LAUNCH();
  }
  return 0;
}

int fd_grid_sample_border_fwd(const float* img, const float* grid, float* out, int B, int C, int H, int W, int Ho,
                              int Wo, void* stream) {
  grid_sample_fwd_kernel<<<blocks_for((long)B * Ho * Wo), NT, 0, (cudaStream_t)stream>>>(img, grid, out, B, C, H, W,
                                                                                          Ho, Wo);
  FD_CHECK_LAUNCH();
  return 0;
}

int fd_grid_sample_border_bwd(const float* img, const float* grid, const float* dout, float* dgrid, float* dimg,
                              int B, int C, int H, int W, int Ho, int Wo, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (dimg) {
    cudaError_t e = cudaMemsetAsync(dimg, 0, sizeof(float) * B * C * H * W, st);
    FD_REQUIRE(e == cudaSuccess, "fd_grid_sample_border_bwd: memset failed: %s", cudaGetErrorString(e));
  }
  grid_sample_bwd_kernel<<<blocks_for((long)B * Ho * Wo), NT, 0, st>>>(img, grid, dout, dgrid, dimg, B, C, H, W, Ho,
                                                                       Wo);
  FD_CHECK_LAUNCH();
  return 0;
}

int fd_pose_matrix_fwd(const float* axisangle, const float* translation, int invert, float* M, int B, void* stream) {
  pose_matrix_fwd_kernel<<<fd::cdiv(B, 32), 32, 0, (cudaStream_t)stream>>>(axisangle, translation, invert, M, B);
  FD_CHECK_LAUNCH();
  return 0;
}

int fd_pose_matrix_bwd(const float* axisangle, const float* translation, int invert, const float* dM,
                       float* d_axisangle, float* d_translation, int B, void* stream) {
  pose_matrix_bwd_kernel<<<fd::cdiv(B, 32), 32, 0, (cudaStream_t)stream>>>(axisangle, translation, invert, dM,
                                                                           d_axisangle, d_translation, B);
  FD_CHECK_LAUNCH();
  return 0;
}

int fd_cat_xy(const float* depth, const float* inv_K, float* out, int B, int H, int W, void* stream) {
  cat_xy_kernel<<<blocks_for((long)B * H * W), NT, 0, (cudaStream_t)stream>>>(depth, inv_K, out, B, H, W);
  FD_CHECK_LAUNCH();
  return 0;
}

int fd_ssim_fwd(const float* x, const float* y, float* out, long planes, int H, int W, void* stream) {
  FD_REQUIRE(H >= 2 && W >= 2, "fd_ssim_fwd: reflection padding needs H, W >= 2");
  ssim_fwd_kernel<<<blocks_for(planes * H * W), NT, 0, (cudaStream_t)stream>>>(x, y, out, planes, H, W);
  FD_CHECK_LAUNCH();
  return 0;
}

// gradient wrt the FIRST argument (SSIM is symmetric: swap x and y for the other one); workspace: 3*planes*H*W floats
int fd_ssim_bwd(const float* x, const float* y, const float* dout, float* dx, long planes, int H, int W,
                void* workspace, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  float* coef = (float*)workspace;
  ssim_coef_kernel<<<blocks_for(planes * H * W), NT, 0, st>>>(x, y, dout, coef, planes, H, W);
  FD_CHECK_LAUNCH();
  ssim_gather_kernel<<<blocks_for(planes * H * W), NT, 0, st>>>(x, y, coef, dx, planes, H, W);
  FD_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
