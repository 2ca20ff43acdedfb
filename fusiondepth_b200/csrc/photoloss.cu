// Fused photometric reprojection loss chain, forward and backward, all four scales in one
// launch each.
//
// Replaces (within 1e-4 relative fp32) the reference's ~1045-op chain
//   trainer.generate_images_pred   (reference trainer.py:425-474)
//   trainer.compute_losses         (trainer.py:490-596)
//   layers.disp_to_depth / BackprojectDepth / Project3D / SSIM / get_smooth_loss
//                                  (layers.py:11-20, 133-162, 204-226, 251-281, 235-248)
//   F.interpolate(bilinear) / F.grid_sample(border)   (trainer.py:434, 467, 579)
// with the reference's default flags (automasking, SSIM, full-res multi-scale, si-loss on
// every scale).  Per-pixel maths: SURVEY.md Appendix A.
//
// Layout: images NCHW fp32 (as the loader delivers them), disparities [B,1,h,w].
// One CTA = one 32x16 pixel tile of one image; the target tile (+halo) is staged in shared
// memory once and reused by the two identity terms and the eight warped SSIM evaluations;
// warped patches are gathered (L1/L2-resident source images) into shared memory so that the
// 3x3 SSIM windows never touch HBM again.  Reductions: warp shuffle -> shared -> one
// per-CTA partial row; a tiny finalize kernel sums the rows in fp64 in a fixed order
// (deterministic, no float atomics).
#include "common.cuh"
#include "../../include/fusiondepth_b200.h"

namespace {

constexpr int TX = 32, TY = 16, NT = 256;
constexpr int F_W = TX + 2, F_H = TY + 2;   // forward: halo 1
constexpr int B_W = TX + 4, B_H = TY + 4;   // backward: halo 2
constexpr float C1 = 1e-4f, C2 = 9e-4f;
constexpr int NPART = 16;                   // per-CTA forward partials: 4 scales x {photo, d, d2, n}
constexpr int NPART_B = 24;                 // per-CTA backward partials: 2 frames x dP[3][4]

struct PLArgs {
  int B, H, W;
  const float* tgt;
  const float* src[2];
  const float* disp[4];
  const float* color[4];
  const float* K;
  const float* invK;
  const float* T[2];
  const float* noise[4];
  const float* beam;
  float min_depth_inv;   // 1/max_depth  (min_disp)
  float disp_range;      // 1/min_depth - 1/max_depth
  float si_thresh, si_var, smooth_w;
  int si_scales;
  float si_pred_mul, si_tgt_mul, si_lo, si_weight;
  // forward outputs
  float* partial;            // [nblk][NPART]
  unsigned char* sel;        // [4][B,H,W] argmin channel (0,1 identity; 2,3 warped)
  float* out_depth[4];       // optional [B,1,H,W]
  float* out_color[4][2];    // optional [B,3,H,W]
  float* out_topt[4];        // optional [B,H,W]
};

__device__ __forceinline__ int reflect_idx(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}

struct Cam {          // per-image matrices, in shared memory
  float ik[9];        // inv_K[:3,:3]
  float P[2][12];     // (K @ T_f)[:3,:]
};

__device__ void load_cam(Cam* cam, const PLArgs& a, int b) {
  int t = threadIdx.x;
  if (t < 9) cam->ik[t] = a.invK[b * 16 + (t / 3) * 4 + (t % 3)];
  if (t >= 32 && t < 32 + 24) {
    int f = (t - 32) / 12, e = (t - 32) % 12, i = e / 4, j = e % 4;
    const float* K = a.K + b * 16;
    const float* T = a.T[f] + b * 16;
    float acc = K[i * 4 + 0] * T[0 * 4 + j];
    acc = fmaf(K[i * 4 + 1], T[1 * 4 + j], acc);
    acc = fmaf(K[i * 4 + 2], T[2 * 4 + j], acc);
    acc = fmaf(K[i * 4 + 3], T[3 * 4 + j], acc);
    cam->P[f][e] = acc;
  }
}

// bilinear upsample of the scale-s disparity at full-res pixel (y,x), align_corners=False
struct Up {
  int y0, y1, x0, x1;
  float ly, lx;
};
__device__ __forceinline__ Up up_coords(int y, int x, int h, int w, int H, int W) {
  Up u;
  float sy = (float)h / (float)H, sx = (float)w / (float)W;
  float fy = sy * ((float)y + 0.5f) - 0.5f, fx = sx * ((float)x + 0.5f) - 0.5f;
  if (fy < 0.f) fy = 0.f;
  if (fx < 0.f) fx = 0.f;
  u.y0 = min((int)fy, h - 1);
  u.x0 = min((int)fx, w - 1);
  u.y1 = u.y0 + (u.y0 < h - 1 ? 1 : 0);
  u.x1 = u.x0 + (u.x0 < w - 1 ? 1 : 0);
  u.ly = fminf(fmaxf(fy - (float)u.y0, 0.f), 1.f);
  u.lx = fminf(fmaxf(fx - (float)u.x0, 0.f), 1.f);
  return u;
}
__device__ __forceinline__ float up_disp(const float* __restrict__ d, int y, int x, int h, int w,
                                         int H, int W) {
  if (h == H && w == W) return d[(long)y * w + x];
  Up u = up_coords(y, x, h, w, H, W);
  float hy = 1.f - u.ly, hx = 1.f - u.lx;
  float t0 = hx * d[(long)u.y0 * w + u.x0] + u.lx * d[(long)u.y0 * w + u.x1];
  float t1 = hx * d[(long)u.y1 * w + u.x0] + u.lx * d[(long)u.y1 * w + u.x1];
  return hy * t0 + u.ly * t1;
}

__device__ __forceinline__ float disp_to_depth(float d, float min_disp, float range) {
  float scaled = __fadd_rn(min_disp, __fmul_rn(range, d));
  return __fdiv_rn(1.0f, scaled);
}

struct Geo {
  float X[3];       // camera-frame point
  float c[3];       // projected homogeneous coords
  float z;          // c2 + eps
  float ix, iy;     // clipped sample coords
  float mx, my;     // d(ix)/d(u) multipliers (0 where the border clip is active)
  int x0, y0;
  float fx, fy;
};

__device__ __forceinline__ Geo project_pixel(const Cam* cam, int f, int y, int x, float depth, int H,
                                             int W) {
  Geo g;
  float fxp = (float)x, fyp = (float)y;
  const float* ik = cam->ik;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    float cm = ik[i * 3 + 0] * fxp;
    cm = fmaf(ik[i * 3 + 1], fyp, cm);
    cm = fmaf(ik[i * 3 + 2], 1.0f, cm);
    g.X[i] = __fmul_rn(depth, cm);
  }
  const float* P = cam->P[f];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    float c = P[i * 4 + 0] * g.X[0];
    c = fmaf(P[i * 4 + 1], g.X[1], c);
    c = fmaf(P[i * 4 + 2], g.X[2], c);
    c = fmaf(P[i * 4 + 3], 1.0f, c);
    g.c[i] = c;
  }
  g.z = __fadd_rn(g.c[2], 1e-7f);
  float u = __fdiv_rn(g.c[0], g.z), v = __fdiv_rn(g.c[1], g.z);
  u = __fdiv_rn(u, (float)(W - 1));
  v = __fdiv_rn(v, (float)(H - 1));
  float gx = __fmul_rn(__fadd_rn(u, -0.5f), 2.f), gy = __fmul_rn(__fadd_rn(v, -0.5f), 2.f);
  float ix = __fdiv_rn(__fadd_rn(__fmul_rn(__fadd_rn(gx, 1.f), (float)W), -1.f), 2.f);
  float iy = __fdiv_rn(__fadd_rn(__fmul_rn(__fadd_rn(gy, 1.f), (float)H), -1.f), 2.f);
  // clip_coordinates_set_grad: borders count as out of bounds for the gradient
  float wm = (float)(W - 1), hm = (float)(H - 1);
  g.mx = (ix <= 0.f || ix >= wm) ? 0.f : 1.f;
  g.my = (iy <= 0.f || iy >= hm) ? 0.f : 1.f;
  ix = fminf(wm, fmaxf(ix, 0.f));
  iy = fminf(hm, fmaxf(iy, 0.f));
  if (!(ix == ix)) { ix = 0.f; g.mx = 0.f; }
  if (!(iy == iy)) { iy = 0.f; g.my = 0.f; }
  g.ix = ix; g.iy = iy;
  float flx = floorf(ix), fly = floorf(iy);
  g.x0 = (int)flx; g.y0 = (int)fly;
  g.fx = ix - flx; g.fy = iy - fly;
  return g;
}

// bilinear border sample of the three channels of `img` (one image, [3,H,W])
__device__ __forceinline__ void sample3(const float* __restrict__ img, const Geo& g, int H, int W,
                                        float out[3]) {
  const long HW = (long)H * W;
  int x0 = g.x0, y0 = g.y0, x1 = x0 + 1, y1 = y0 + 1;
  float wx1 = g.fx, wx0 = 1.f - g.fx, wy1 = g.fy, wy0 = 1.f - g.fy;   // (ix_se-ix) = 1-fx
  bool bx1 = x1 <= W - 1, by1 = y1 <= H - 1;
  float nw = wx0 * wy0, ne = wx1 * wy0, sw = wx0 * wy1, se = wx1 * wy1;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float* p = img + c * HW;
    float acc = p[(long)y0 * W + x0] * nw;
    if (bx1) acc += p[(long)y0 * W + x1] * ne;
    if (by1) acc += p[(long)y1 * W + x0] * sw;
    if (bx1 && by1) acc += p[(long)y1 * W + x1] * se;
    out[c] = acc;
  }
}

// SSIM loss value (before the channel mean) at a tile pixel; xs/ys point at the window centre
__device__ __forceinline__ float ssim_at(const float* __restrict__ xs, const float* __restrict__ ys,
                                         int stride, float mu_y, float sig_y) {
  float sx = 0.f, sxx = 0.f, sxy = 0.f;
#pragma unroll
  for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
    for (int dx = -1; dx <= 1; ++dx) {
      float xv = xs[dy * stride + dx], yv = ys[dy * stride + dx];
      sx = __fadd_rn(sx, xv);
      sxx = __fadd_rn(sxx, __fmul_rn(xv, xv));
      sxy = __fadd_rn(sxy, __fmul_rn(xv, yv));
    }
  float mu_x = __fdiv_rn(sx, 9.f);
  float sig_x = __fadd_rn(__fdiv_rn(sxx, 9.f), -__fmul_rn(mu_x, mu_x));
  float sig_xy = __fadd_rn(__fdiv_rn(sxy, 9.f), -__fmul_rn(mu_x, mu_y));
  float n = __fmul_rn(__fadd_rn(__fmul_rn(__fmul_rn(2.f, mu_x), mu_y), C1),
                      __fadd_rn(__fmul_rn(2.f, sig_xy), C2));
  float d = __fmul_rn(__fadd_rn(__fadd_rn(__fmul_rn(mu_x, mu_x), __fmul_rn(mu_y, mu_y)), C1),
                      __fadd_rn(__fadd_rn(sig_x, sig_y), C2));
  float s = __fdiv_rn(__fadd_rn(1.f, -__fdiv_rn(n, d)), 2.f);
  return fminf(fmaxf(s, 0.f), 1.f);
}

// ---------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT) photoloss_fwd_kernel(PLArgs a) {
  __shared__ float Tg[3][F_H][F_W];
  __shared__ float Xp[3][F_H][F_W];
  __shared__ Cam cam;
  __shared__ float red[NPART * 32];

  const int b = blockIdx.z, x0 = blockIdx.x * TX, y0 = blockIdx.y * TY;
  const int H = a.H, W = a.W;
  const long HW = (long)H * W;
  const int t = threadIdx.x;
  const int px = t & 31, pyb = t >> 5;      // pixels (pyb, px) and (pyb+8, px)

  load_cam(&cam, a, b);
  const float* tgt = a.tgt + (long)b * 3 * HW;
  for (int c = t; c < F_H * F_W; c += NT) {
    int cy = c / F_W, cx = c % F_W;
    int iy = reflect_idx(y0 - 1 + cy, H), ix = reflect_idx(x0 - 1 + cx, W);
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) Tg[ch][cy][cx] = tgt[ch * HW + (long)iy * W + ix];
  }
  __syncthreads();

  // target window statistics, reused by all ten SSIM evaluations
  float mu_t[2][3], sig_t[2][3];
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    int ly = pyb + 8 * k + 1, lx = px + 1;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      float s = 0.f, ss = 0.f;
#pragma unroll
      for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) {
          float v = Tg[ch][ly + dy][lx + dx];
          s = __fadd_rn(s, v);
          ss = __fadd_rn(ss, __fmul_rn(v, v));
        }
      float m = __fdiv_rn(s, 9.f);
      mu_t[k][ch] = m;
      sig_t[k][ch] = __fadd_rn(__fdiv_rn(ss, 9.f), -__fmul_rn(m, m));
    }
  }

  // loss of the tile in Xp against the target, at this thread's pixel k
  auto reproj = [&](int k) -> float {
    int ly = pyb + 8 * k + 1, lx = px + 1;
    float l1 = 0.f, ss = 0.f;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      l1 = __fadd_rn(l1, fabsf(__fadd_rn(Tg[ch][ly][lx], -Xp[ch][ly][lx])));
      ss = __fadd_rn(ss, ssim_at(&Xp[ch][ly][lx], &Tg[ch][ly][lx], F_W, mu_t[k][ch], sig_t[k][ch]));
    }
    return __fadd_rn(__fmul_rn(0.85f, __fdiv_rn(ss, 3.f)), __fmul_rn(0.15f, __fdiv_rn(l1, 3.f)));
  };

  // identity (un-warped source) terms: scale independent
  float ident[2][2];
#pragma unroll
  for (int f = 0; f < 2; ++f) {
    const float* src = a.src[f] + (long)b * 3 * HW;
    for (int c = t; c < F_H * F_W; c += NT) {
      int cy = c / F_W, cx = c % F_W;
      int iy = reflect_idx(y0 - 1 + cy, H), ix = reflect_idx(x0 - 1 + cx, W);
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) Xp[ch][cy][cx] = src[ch * HW + (long)iy * W + ix];
    }
    __syncthreads();
    ident[f][0] = reproj(0);
    ident[f][1] = reproj(1);
    __syncthreads();
  }

  float part[NPART];
#pragma unroll
  for (int i = 0; i < NPART; ++i) part[i] = 0.f;

#pragma unroll 1
  for (int s = 0; s < 4; ++s) {
    const int h = H >> s, w = W >> s;
    const float* disp = a.disp[s] + (long)b * h * w;
    float rp[2][2];
#pragma unroll 1
    for (int f = 0; f < 2; ++f) {
      const float* src = a.src[f] + (long)b * 3 * HW;
      for (int c = t; c < F_H * F_W; c += NT) {
        int cy = c / F_W, cx = c % F_W;
        int ey = y0 - 1 + cy, ex = x0 - 1 + cx;
        int iy = reflect_idx(ey, H), ix = reflect_idx(ex, W);
        float d = up_disp(disp, iy, ix, h, w, H, W);
        float depth = disp_to_depth(d, a.min_depth_inv, a.disp_range);
        Geo g = project_pixel(&cam, f, iy, ix, depth, H, W);
        float col[3];
        sample3(src, g, H, W, col);
        Xp[0][cy][cx] = col[0]; Xp[1][cy][cx] = col[1]; Xp[2][cy][cx] = col[2];
        bool interior = (ey == iy) && (ex == ix) && cy >= 1 && cy <= TY && cx >= 1 && cx <= TX;
        if (interior) {
          long o = (long)iy * W + ix;
          if (a.out_color[s][f]) {
            float* oc = a.out_color[s][f] + (long)b * 3 * HW;
            oc[o] = col[0]; oc[HW + o] = col[1]; oc[2 * HW + o] = col[2];
          }
          if (f == 0 && a.out_depth[s]) a.out_depth[s][(long)b * HW + o] = depth;
        }
      }
      __syncthreads();
      rp[f][0] = reproj(0);
      rp[f][1] = reproj(1);
      __syncthreads();
    }
    const float* nz = a.noise[s] + (long)b * 2 * HW;
    float ps[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      int y = y0 + pyb + 8 * k, x = x0 + px;
      if (y >= H || x >= W) continue;
      long o = (long)y * W + x;
      float c0 = __fadd_rn(ident[0][k], __fmul_rn(nz[o], 1e-5f));
      float c1 = __fadd_rn(ident[1][k], __fmul_rn(nz[HW + o], 1e-5f));
      float m = c0; int idx = 0;
      if (c1 < m) { m = c1; idx = 1; }
      if (rp[0][k] < m) { m = rp[0][k]; idx = 2; }
      if (rp[1][k] < m) { m = rp[1][k]; idx = 3; }
      a.sel[((long)s * a.B + b) * HW + o] = (unsigned char)idx;
      if (a.out_topt[s]) a.out_topt[s][(long)b * HW + o] = m;
      ps[0] += m;
      if ((a.si_scales >> s) & 1) {
        float d = up_disp(disp, y, x, h, w, H, W);
        float D = __fmul_rn(disp_to_depth(d, a.min_depth_inv, a.disp_range), a.si_pred_mul);
        float Bm = __fmul_rn(a.beam[(long)b * HW + o], a.si_tgt_mul);
        bool valid = (Bm > a.si_lo) && (D < 80.f) && (D > a.si_lo) && (fabsf(__fadd_rn(D, -Bm)) < a.si_thresh);
        if (valid) {
          float dl = __fadd_rn(logf(D), -logf(Bm));
          ps[1] += dl;
          ps[2] += dl * dl;
          ps[3] += 1.f;
        }
      }
    }
#pragma unroll
    for (int ss = 0; ss < 4; ++ss)
      if (ss == s) {
#pragma unroll
        for (int i = 0; i < 4; ++i) part[ss * 4 + i] += ps[i];
      }
  }
  fd::block_sum<NPART>(part, red);
  if (t == 0) {
    long blk = ((long)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
#pragma unroll
    for (int i = 0; i < NPART; ++i) a.partial[blk * NPART + i] = part[i];
  }
}

// ---------------------------------------------------------------------------------------
// per-image mean of each disparity scale   (trainer.py:569)
// ---------------------------------------------------------------------------------------
struct SmArgs {
  int B, H, W;
  const float* disp[4];
  const float* color[4];
  float* mean;        // [B*4][SM_CHUNKS] partial sums of disp
  float* sm_partial;  // [B*4][SM_CHUNKS][2]
};
constexpr int SM_CHUNKS = 32;

__global__ void disp_mean_kernel(SmArgs a) {
  // partial sums of disp over SM_CHUNKS blocks per (image, scale); block_mean() finishes them
  __shared__ double red[32];
  const int b = blockIdx.y >> 2, s = blockIdx.y & 3;
  const int n = (a.H >> s) * (a.W >> s);
  const float* d = a.disp[s] + (long)b * n;
  double acc[1] = {0.0};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) acc[0] += (double)d[i];
  fd::block_sum<1>(acc, red);
  if (threadIdx.x == 0) a.mean[(long)blockIdx.y * SM_CHUNKS + blockIdx.x] = (float)acc[0];
}

// mean of disp_s of image b from the partial sums (fixed order => deterministic), for the whole block
__device__ float block_mean(const float* __restrict__ mpart, int bs, int n) {
  __shared__ float mean_sh;
  if (threadIdx.x == 0) {
    double t = 0;
    for (int c = 0; c < SM_CHUNKS; ++c) t += (double)mpart[(long)bs * SM_CHUNKS + c];
    mean_sh = (float)(t / (double)n);
  }
  __syncthreads();
  return mean_sh;
}

// edge-aware smoothness partial sums   (layers.py:235-248 on norm_disp, trainer.py:570-571)
__global__ void smooth_fwd_kernel(SmArgs a) {
  __shared__ double red[64];
  const int b = blockIdx.y >> 2, s = blockIdx.y & 3;
  const int h = a.H >> s, w = a.W >> s, n = h * w;
  const float* d = a.disp[s] + (long)b * n;
  const float* img = a.color[s] + (long)b * 3 * n;
  const float den = __fadd_rn(block_mean(a.mean, blockIdx.y, n), 1e-7f);
  double acc[2] = {0.0, 0.0};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    int y = i / w, x = i % w;
    float n0 = __fdiv_rn(d[i], den);
    if (x + 1 < w) {
      float gd = fabsf(__fadd_rn(n0, -__fdiv_rn(d[i + 1], den)));
      float gi = 0.f;
#pragma unroll
      for (int c = 0; c < 3; ++c) gi = __fadd_rn(gi, fabsf(__fadd_rn(img[c * n + i], -img[c * n + i + 1])));
      acc[0] += (double)__fmul_rn(gd, expf(-__fdiv_rn(gi, 3.f)));
    }
    if (y + 1 < h) {
      float gd = fabsf(__fadd_rn(n0, -__fdiv_rn(d[i + w], den)));
      float gi = 0.f;
#pragma unroll
      for (int c = 0; c < 3; ++c) gi = __fadd_rn(gi, fabsf(__fadd_rn(img[c * n + i], -img[c * n + i + w])));
      acc[1] += (double)__fmul_rn(gd, expf(-__fdiv_rn(gi, 3.f)));
    }
  }
  fd::block_sum<2>(acc, red);
  if (threadIdx.x == 0) {
    a.sm_partial[((long)blockIdx.y * SM_CHUNKS + blockIdx.x) * 2 + 0] = (float)acc[0];
    a.sm_partial[((long)blockIdx.y * SM_CHUNKS + blockIdx.x) * 2 + 1] = (float)acc[1];
  }
}

// losses[0..3] = loss/s, [4..7] = si_loss s, [8] = loss ; stats = saved-for-backward scalars
//   stats[s*4+{0,1,2}] = n_s, mean(delta)_s, sqrt(var term)_s ; stats[16 + (b*4+s)] = L_sm of image b
__global__ void loss_finalize_kernel(const float* __restrict__ partial, int nblk,
                                     const float* __restrict__ sm_partial, int B, int H, int W,
                                     float si_var, float smooth_w, int si_scales, float si_weight,
                                     float* losses, float* stats) {
  __shared__ double red[NPART * 32];
  __shared__ double tot[NPART];
  double acc[NPART];
#pragma unroll
  for (int i = 0; i < NPART; ++i) acc[i] = 0.0;
  for (int r = threadIdx.x; r < nblk; r += blockDim.x)
#pragma unroll
    for (int i = 0; i < NPART; ++i) acc[i] += (double)partial[(long)r * NPART + i];
  fd::block_sum<NPART>(acc, red);
  if (threadIdx.x == 0)
    for (int i = 0; i < NPART; ++i) tot[i] = acc[i];
  __syncthreads();
  // per-(image,scale) smoothness
  for (int p = threadIdx.x; p < B * 4; p += blockDim.x) {
    int s = p & 3;
    int h = H >> s, w = W >> s;
    double sx = 0, sy = 0;
    for (int c = 0; c < SM_CHUNKS; ++c) {
      sx += sm_partial[((long)p * SM_CHUNKS + c) * 2 + 0];
      sy += sm_partial[((long)p * SM_CHUNKS + c) * 2 + 1];
    }
    double nx = (double)B * h * (w - 1), ny = (double)B * (h - 1) * w;
    stats[16 + p] = (float)(sx / nx + sy / ny);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double total = 0;
    for (int s = 0; s < 4; ++s) {
      double photo = tot[s * 4 + 0] / ((double)B * H * W);
      double sm = 0;
      for (int b = 0; b < B; ++b) sm += stats[16 + b * 4 + s];
      double ls = photo + (double)smooth_w * sm / (double)(1 << s);
      losses[s] = (float)ls;
      total += ls;
      if ((si_scales >> s) & 1) {
        double n = tot[s * 4 + 3];
        double m1 = tot[s * 4 + 1] / n, m2 = tot[s * 4 + 2] / n;
        double root = sqrt(m2 - (double)si_var * m1 * m1);
        losses[4 + s] = (float)((double)si_weight * root);
        stats[s * 4 + 0] = (float)n; stats[s * 4 + 1] = (float)m1; stats[s * 4 + 2] = (float)root;
        total += (double)si_weight * root;
      } else {
        losses[4 + s] = 0.f;
      }
    }
    losses[8] = (float)(total / 4.0);
  }
}

// ---------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------
struct PLBwdArgs {
  PLArgs f;
  const float* gout;        // d loss["loss"], device scalar
  const float* stats;       // from loss_finalize_kernel
  float* gd_up[4];          // [B,H,W] gradient wrt the upsampled disparity
  float* partial_dP;        // [nblk][NPART_B]
};

__global__ void __launch_bounds__(NT) photoloss_bwd_kernel(PLBwdArgs ba) {
  const PLArgs& a = ba.f;
  __shared__ float Tg[3][B_H][B_W];
  __shared__ float Xp[3][B_H][B_W];
  __shared__ float Cf[3][3][F_H][F_W];      // [coef alpha,beta,gamma][channel]
  __shared__ unsigned char Sel[F_H][F_W];
  __shared__ Cam cam;
  __shared__ float red[NPART_B * 32];

  const int b = blockIdx.z, x0 = blockIdx.x * TX, y0 = blockIdx.y * TY;
  const int H = a.H, W = a.W;
  const long HW = (long)H * W;
  const int t = threadIdx.x;
  const int px = t & 31, pyb = t >> 5;
  const float g_total = ba.gout[0];
  const float g_photo = g_total * 0.25f / ((float)a.B * (float)H * (float)W);

  load_cam(&cam, a, b);
  const float* tgt = a.tgt + (long)b * 3 * HW;
  for (int c = t; c < B_H * B_W; c += NT) {
    int cy = c / B_W, cx = c % B_W;
    int iy = reflect_idx(y0 - 2 + cy, H), ix = reflect_idx(x0 - 2 + cx, W);
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) Tg[ch][cy][cx] = tgt[ch * HW + (long)iy * W + ix];
  }
  float dP[NPART_B];
#pragma unroll
  for (int i = 0; i < NPART_B; ++i) dP[i] = 0.f;
  __syncthreads();

#pragma unroll 1
  for (int s = 0; s < 4; ++s) {
    const int h = H >> s, w = W >> s;
    const float* disp = a.disp[s] + (long)b * h * w;
    const unsigned char* sel = a.sel + ((long)s * a.B + b) * HW;
    float gd[2] = {0.f, 0.f};
    // selection over tile + halo 1 (cells outside the image select nothing)
    for (int c = t; c < F_H * F_W; c += NT) {
      int cy = c / F_W, cx = c % F_W;
      int ey = y0 - 1 + cy, ex = x0 - 1 + cx;
      Sel[cy][cx] = (ey >= 0 && ey < H && ex >= 0 && ex < W) ? sel[(long)ey * W + ex] : 255;
    }
    __syncthreads();
#pragma unroll 1
    for (int f = 0; f < 2; ++f) {
      int mine = 0;
      for (int c = t; c < F_H * F_W; c += NT) mine |= (Sel[c / F_W][c % F_W] == f + 2);
      if (!__syncthreads_or(mine)) continue;
      const float* src = a.src[f] + (long)b * 3 * HW;
      for (int c = t; c < B_H * B_W; c += NT) {
        int cy = c / B_W, cx = c % B_W;
        int iy = reflect_idx(y0 - 2 + cy, H), ix = reflect_idx(x0 - 2 + cx, W);
        float d = up_disp(disp, iy, ix, h, w, H, W);
        float depth = disp_to_depth(d, a.min_depth_inv, a.disp_range);
        Geo g = project_pixel(&cam, f, iy, ix, depth, H, W);
        float col[3];
        sample3(src, g, H, W, col);
        Xp[0][cy][cx] = col[0]; Xp[1][cy][cx] = col[1]; Xp[2][cy][cx] = col[2];
      }
      __syncthreads();
      // SSIM adjoint coefficients at every selected pixel q of tile + halo 1:
      //   dS_q/dx_p = alpha_q + beta_q x_p + gamma_q y_p   for p in the 3x3 window of q
      const float wq = g_photo * (0.85f / 3.f);
      for (int c = t; c < F_H * F_W; c += NT) {
        int cy = c / F_W, cx = c % F_W;
        bool on = Sel[cy][cx] == f + 2;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
          float al = 0.f, be = 0.f, ga = 0.f;
          if (on) {
            const float* xs = &Xp[ch][cy + 1][cx + 1];
            const float* ys = &Tg[ch][cy + 1][cx + 1];
            float sx = 0.f, sy = 0.f, sxx = 0.f, syy = 0.f, sxy = 0.f;
#pragma unroll
            for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
              for (int dx = -1; dx <= 1; ++dx) {
                float xv = xs[dy * B_W + dx], yv = ys[dy * B_W + dx];
                sx += xv; sy += yv; sxx += xv * xv; syy += yv * yv; sxy += xv * yv;
              }
            float mx = sx / 9.f, my = sy / 9.f;
            float vx = sxx / 9.f - mx * mx, vy = syy / 9.f - my * my, vxy = sxy / 9.f - mx * my;
            float n1 = 2.f * mx * my + C1, n2 = 2.f * vxy + C2;
            float d1 = mx * mx + my * my + C1, d2 = vx + vy + C2;
            float n = n1 * n2, d = d1 * d2;
            float S = (1.f - n / d) * 0.5f;
            if (S >= 0.f && S <= 1.f) {
              float a0 = (2.f / 9.f) * my * (n2 - n1), a1 = (2.f / 9.f) * n1;
              float b0 = (2.f / 9.f) * mx * (d2 - d1), b1 = (2.f / 9.f) * d1;
              float inv_d = 1.f / d, nd2 = n * inv_d * inv_d;
              al = wq * 0.5f * (nd2 * b0 - a0 * inv_d);
              be = wq * 0.5f * nd2 * b1;
              ga = -wq * 0.5f * a1 * inv_d;
            }
          }
          Cf[0][ch][cy][cx] = al; Cf[1][ch][cy][cx] = be; Cf[2][ch][cy][cx] = ga;
        }
      }
      __syncthreads();
      // gradient at the interior pixels
      float dPl[12];
#pragma unroll
      for (int i = 0; i < 12; ++i) dPl[i] = 0.f;
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        int y = y0 + pyb + 8 * k, x = x0 + px;
        if (y >= H || x >= W) continue;
        // extended positions that alias this pixel through the reflection pad
        int eys[3], exs[3], ney = 0, nex = 0;
        eys[ney++] = y; if (y == 1) eys[ney++] = -1; if (y == H - 2) eys[ney++] = H;
        exs[nex++] = x; if (x == 1) exs[nex++] = -1; if (x == W - 2) exs[nex++] = W;
        float G[3];
        bool any = false;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
          float A = 0.f, Bc = 0.f, Gm = 0.f;
          for (int iy = 0; iy < ney; ++iy)
            for (int ix = 0; ix < nex; ++ix) {
              int cy = eys[iy] - (y0 - 1), cx = exs[ix] - (x0 - 1);
#pragma unroll
              for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
                for (int dx = -1; dx <= 1; ++dx) {
                  int qy = cy + dy, qx = cx + dx;
                  if (qy >= 0 && qy < F_H && qx >= 0 && qx < F_W) {
                    A += Cf[0][ch][qy][qx]; Bc += Cf[1][ch][qy][qx]; Gm += Cf[2][ch][qy][qx];
                  }
                }
            }
          int ly = pyb + 8 * k + 2, lx = px + 2;
          float xv = Xp[ch][ly][lx], tv = Tg[ch][ly][lx];
          float g = A + Bc * xv + Gm * tv;
          if (Sel[pyb + 8 * k + 1][px + 1] == f + 2) {
            float df = tv - xv;
            float sg = (df > 0.f) ? 1.f : ((df < 0.f) ? -1.f : 0.f);
            g += g_photo * (0.15f / 3.f) * (-sg);
          }
          G[ch] = g;
          any |= (g != 0.f);
        }
        if (!any) continue;
        float d = up_disp(disp, y, x, h, w, H, W);
        float depth = disp_to_depth(d, a.min_depth_inv, a.disp_range);
        Geo g = project_pixel(&cam, f, y, x, depth, H, W);
        // d(warped)/d(ix,iy)
        int xa = g.x0, ya = g.y0, xb = xa + 1, yb = ya + 1;
        bool bx1 = xb <= W - 1, by1 = yb <= H - 1;
        float gix = 0.f, giy = 0.f;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
          const float* p = src + ch * HW;
          float vnw = p[(long)ya * W + xa];
          float vne = bx1 ? p[(long)ya * W + xb] : 0.f;
          float vsw = by1 ? p[(long)yb * W + xa] : 0.f;
          float vse = (bx1 && by1) ? p[(long)yb * W + xb] : 0.f;
          gix += G[ch] * ((vne - vnw) * (1.f - g.fy) + (vse - vsw) * g.fy);
          giy += G[ch] * ((vsw - vnw) * (1.f - g.fx) + (vse - vne) * g.fx);
        }
        // ix = ((gx+1) W - 1)/2, gx = (u/(W-1) - .5) 2  =>  d ix / d u = W/(W-1)
        float gu = gix * g.mx * ((float)W / (float)(W - 1));
        float gv = giy * g.my * ((float)H / (float)(H - 1));
        float iz = 1.f / g.z;
        float gc0 = gu * iz, gc1 = gv * iz;
        float gc2 = -(gu * g.c[0] + gv * g.c[1]) * iz * iz;
        const float* P = cam.P[f];
        float gdepth = 0.f;
        float gc[3] = {gc0, gc1, gc2};
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          // dc_i/d depth = sum_j P[i][j] cam_j with X_j = depth * cam_j
          float dcd = (P[i * 4 + 0] * g.X[0] + P[i * 4 + 1] * g.X[1] + P[i * 4 + 2] * g.X[2]) / depth;
          gdepth += gc[i] * dcd;
          dPl[i * 4 + 0] += gc[i] * g.X[0];
          dPl[i * 4 + 1] += gc[i] * g.X[1];
          dPl[i * 4 + 2] += gc[i] * g.X[2];
          dPl[i * 4 + 3] += gc[i];
        }
        gd[k] += gdepth * (-a.disp_range * depth * depth);
      }
#pragma unroll
      for (int ff = 0; ff < 2; ++ff)
        if (ff == f) {
#pragma unroll
          for (int i = 0; i < 12; ++i) dP[ff * 12 + i] += dPl[i];
        }
      __syncthreads();
    }
    // si-loss term and store
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      int y = y0 + pyb + 8 * k, x = x0 + px;
      if (y >= H || x >= W) continue;
      long o = (long)y * W + x;
      float g = gd[k];
      if ((a.si_scales >> s) & 1) {
        float d = up_disp(disp, y, x, h, w, H, W);
        float depth = disp_to_depth(d, a.min_depth_inv, a.disp_range);
        float D = __fmul_rn(depth, a.si_pred_mul);
        float Bm = __fmul_rn(a.beam[(long)b * HW + o], a.si_tgt_mul);
        bool valid = (Bm > a.si_lo) && (D < 80.f) && (D > a.si_lo) && (fabsf(__fadd_rn(D, -Bm)) < a.si_thresh);
        if (valid) {
          float n = ba.stats[s * 4 + 0], m1 = ba.stats[s * 4 + 1], root = ba.stats[s * 4 + 2];
          float dl = __fadd_rn(logf(D), -logf(Bm));
          // d(w sqrt(mean d^2 - v mean(d)^2))/d delta_i, then d delta/d depth = 1/depth
          float coef = g_total * 0.25f * a.si_weight * (dl - a.si_var * m1) / (n * root);
          g += coef * (-a.disp_range * depth);
        }
      }
      ba.gd_up[s][(long)b * HW + o] = g;
    }
    __syncthreads();
  }
  fd::block_sum<NPART_B>(dP, red);
  if (t == 0) {
    long blk = ((long)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
#pragma unroll
    for (int i = 0; i < NPART_B; ++i) ba.partial_dP[blk * NPART_B + i] = dP[i];
  }
}

// grad wrt disp_s = bilinear-upsample adjoint of gd_up (gather form) + smoothness gradient
struct DGArgs {
  int B, H, W;
  const float* disp[4];
  const float* color[4];
  const float* gd_up[4];
  const float* mean;     // [B*4][SM_CHUNKS] partial sums of disp
  const float* stats;
  const float* gout;
  float smooth_w;
  float* gdisp[4];
};

__device__ __forceinline__ float sgnf(float v) { return (v > 0.f) ? 1.f : ((v < 0.f) ? -1.f : 0.f); }

__global__ void disp_grad_kernel(DGArgs a) {
  const int b = blockIdx.y >> 2, s = blockIdx.y & 3;
  const int H = a.H, W = a.W, h = H >> s, w = W >> s, n = h * w;
  const int r = 1 << s;
  const float* d = a.disp[s] + (long)b * n;
  const float* img = a.color[s] + (long)b * 3 * n;
  const float* gu = a.gd_up[s] + (long)b * H * W;
  const float den = __fadd_rn(block_mean(a.mean, blockIdx.y, n), 1e-7f);
  const float Lb = a.stats[16 + b * 4 + s];
  const float gs = a.gout[0] * 0.25f * a.smooth_w / (float)r;
  const float inx = 1.f / ((float)a.B * h * (w - 1)), iny = 1.f / ((float)a.B * (h - 1) * w);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    int yl = i / w, xl = i % w;
    float acc = 0.f;
    if (s == 0) {
      acc = gu[i];
    } else {
      int ya = max(0, r * yl - r / 2 - 1), yb = min(H - 1, r * yl + r + r / 2);
      int xa = max(0, r * xl - r / 2 - 1), xb = min(W - 1, r * xl + r + r / 2);
      for (int y = ya; y <= yb; ++y) {
        Up uy = up_coords(y, 0, h, w, H, W);
        float wy = (uy.y0 == yl ? 1.f - uy.ly : 0.f) + (uy.y1 == yl ? uy.ly : 0.f);
        if (wy == 0.f) continue;
        float row = 0.f;
        for (int x = xa; x <= xb; ++x) {
          Up ux = up_coords(0, x, h, w, H, W);
          float wx = (ux.x0 == xl ? 1.f - ux.lx : 0.f) + (ux.x1 == xl ? ux.lx : 0.f);
          if (wx != 0.f) row += wx * gu[(long)y * W + x];
        }
        acc += wy * row;
      }
    }
    // smoothness: d/dn_i of (1/Nx) sum |n_i - n_{i+1}| wx_i + (1/Ny) sum |n_i - n_{i+w}| wy_i
    float n0 = __fdiv_rn(d[i], den);
    float gn = 0.f;
    auto wgt = [&](int i0, int i1) {
      float gi = 0.f;
#pragma unroll
      for (int c = 0; c < 3; ++c) gi = __fadd_rn(gi, fabsf(__fadd_rn(img[c * n + i0], -img[c * n + i1])));
      return expf(-__fdiv_rn(gi, 3.f));
    };
    if (xl + 1 < w) gn += inx * sgnf(__fadd_rn(n0, -__fdiv_rn(d[i + 1], den))) * wgt(i, i + 1);
    if (xl > 0) gn -= inx * sgnf(__fadd_rn(__fdiv_rn(d[i - 1], den), -n0)) * wgt(i - 1, i);
    if (yl + 1 < h) gn += iny * sgnf(__fadd_rn(n0, -__fdiv_rn(d[i + w], den))) * wgt(i, i + w);
    if (yl > 0) gn -= iny * sgnf(__fadd_rn(__fdiv_rn(d[i - w], den), -n0)) * wgt(i - w, i);
    // n = d/(mean+eps): dL/dd_j = (gn_j - L_b/(h w)) / (mean+eps)   (Euler: sum_i n_i gn_i = L_b)
    acc += gs * (gn - Lb / (float)n) / den;
    a.gdisp[s][(long)b * n + i] = acc;
  }
}

// dT_f = K[:3,:]^T dP_f, summed over the tiles of image b
__global__ void pose_grad_kernel(const float* __restrict__ partial_dP, int tiles_per_image,
                                 const float* __restrict__ K, float* gT0, float* gT1) {
  __shared__ double red[NPART_B * 32];
  __shared__ double dP[NPART_B];
  const int b = blockIdx.x;
  double acc[NPART_B];
#pragma unroll
  for (int i = 0; i < NPART_B; ++i) acc[i] = 0.0;
  for (int r = threadIdx.x; r < tiles_per_image; r += blockDim.x)
#pragma unroll
    for (int i = 0; i < NPART_B; ++i)
      acc[i] += (double)partial_dP[((long)b * tiles_per_image + r) * NPART_B + i];
  fd::block_sum<NPART_B>(acc, red);
  if (threadIdx.x == 0)
    for (int i = 0; i < NPART_B; ++i) dP[i] = acc[i];
  __syncthreads();
  if (threadIdx.x < 32) {
    int f = threadIdx.x >> 4, e = threadIdx.x & 15, k = e >> 2, j = e & 3;
    const float* Kb = K + b * 16;
    double v = 0;
    for (int i = 0; i < 3; ++i) v += (double)Kb[i * 4 + k] * dP[f * 12 + i * 4 + j];
    (f == 0 ? gT0 : gT1)[b * 16 + e] = (float)v;
  }
}

int fill_args(PLArgs& a, const fd_photoloss_desc* d) {
  a.B = d->B; a.H = d->H; a.W = d->W;
  a.tgt = d->color[0][0];
  a.src[0] = d->color[1][0];
  a.src[1] = d->color[2][0];
  for (int s = 0; s < 4; ++s) {
    a.disp[s] = d->disp[s];
    a.color[s] = d->color[0][s];
    a.noise[s] = d->noise[s];
    a.out_depth[s] = d->out_depth[s];
    a.out_color[s][0] = d->out_color[s][0];
    a.out_color[s][1] = d->out_color[s][1];
    a.out_topt[s] = d->out_to_optimise[s];
  }
  a.K = d->K; a.invK = d->inv_K;
  a.T[0] = d->T[0]; a.T[1] = d->T[1];
  a.beam = d->beam;
  a.min_depth_inv = (float)(1.0 / (double)d->max_depth);
  a.disp_range = (float)(1.0 / (double)d->min_depth - 1.0 / (double)d->max_depth);
  a.si_thresh = d->si_thresh; a.si_var = d->si_var; a.smooth_w = d->smoothness;
  a.si_scales = d->beam ? d->si_scales : 0;
  a.si_pred_mul = d->si_pred_mul; a.si_tgt_mul = d->si_tgt_mul; a.si_lo = d->si_lo; a.si_weight = d->si_weight;
  a.sel = d->sel;
  return 0;
}

}  // namespace

extern "C" {

// workspace layout (floats): [partial fwd nblk*16][sm_partial B*4*32*2][disp partial sums B*4*32][stats 16+B*4]
//                            [partial_dP nblk*24][gd_up 4*B*H*W]
size_t fd_photoloss_workspace_bytes(int B, int H, int W) {
  long nblk = (long)fd::cdiv(W, TX) * fd::cdiv(H, TY) * B;
  long fl = nblk * NPART + (long)B * 4 * SM_CHUNKS * 2 + (long)B * 4 * SM_CHUNKS + 16 + B * 4 + nblk * NPART_B +
            4L * B * H * W;
  return (size_t)fl * sizeof(float);
}

struct WsView {
  float *partial, *sm_partial, *mean, *stats, *partial_dP, *gd_up;
  long nblk;
};
static WsView ws_view(void* ws, int B, int H, int W) {
  WsView v;
  v.nblk = (long)fd::cdiv(W, TX) * fd::cdiv(H, TY) * B;
  float* p = (float*)ws;
  v.partial = p; p += v.nblk * NPART;
  v.sm_partial = p; p += (long)B * 4 * SM_CHUNKS * 2;
  v.mean = p; p += (long)B * 4 * SM_CHUNKS;
  v.stats = p; p += 16 + B * 4;
  v.partial_dP = p; p += v.nblk * NPART_B;
  v.gd_up = p;
  return v;
}

int fd_photoloss_fwd(const fd_photoloss_desc* d, float* losses, void* workspace, void* stream) {
  FD_REQUIRE(d->B > 0 && d->H >= 32 && d->W >= 32 && d->H % 8 == 0 && d->W % 8 == 0,
             "fd_photoloss_fwd: H,W must be multiples of 8 and >= 32 (got %dx%d)", d->H, d->W);
  FD_REQUIRE(d->sel != nullptr, "fd_photoloss_fwd: sel buffer is required");
  cudaStream_t st = (cudaStream_t)stream;
  PLArgs a;
  fill_args(a, d);
  WsView v = ws_view(workspace, d->B, d->H, d->W);
  a.partial = v.partial;
  SmArgs sa;
  sa.B = d->B; sa.H = d->H; sa.W = d->W;
  for (int s = 0; s < 4; ++s) { sa.disp[s] = a.disp[s]; sa.color[s] = a.color[s]; }
  sa.mean = v.mean; sa.sm_partial = v.sm_partial;
  disp_mean_kernel<<<dim3(SM_CHUNKS, d->B * 4), 256, 0, st>>>(sa);
  FD_CHECK_LAUNCH();
  smooth_fwd_kernel<<<dim3(SM_CHUNKS, d->B * 4), 256, 0, st>>>(sa);
  FD_CHECK_LAUNCH();
  dim3 grid(fd::cdiv(d->W, TX), fd::cdiv(d->H, TY), d->B);
  photoloss_fwd_kernel<<<grid, NT, 0, st>>>(a);
  FD_CHECK_LAUNCH();
  loss_finalize_kernel<<<1, 256, 0, st>>>(v.partial, (int)v.nblk, v.sm_partial, d->B, d->H, d->W,
                                          d->si_var, d->smoothness, a.si_scales, d->si_weight, losses, v.stats);
  FD_CHECK_LAUNCH();
  return 0;
}

int fd_photoloss_bwd(const fd_photoloss_desc* d, const float* grad_loss, float* const grad_disp[4],
                     float* grad_T0, float* grad_T1, void* workspace, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  PLBwdArgs ba;
  fill_args(ba.f, d);
  WsView v = ws_view(workspace, d->B, d->H, d->W);
  ba.gout = grad_loss;
  ba.stats = v.stats;
  long N = (long)d->B * d->H * d->W;
  for (int s = 0; s < 4; ++s) ba.gd_up[s] = v.gd_up + s * N;
  ba.partial_dP = v.partial_dP;
  dim3 grid(fd::cdiv(d->W, TX), fd::cdiv(d->H, TY), d->B);
  photoloss_bwd_kernel<<<grid, NT, 0, st>>>(ba);
  FD_CHECK_LAUNCH();
  DGArgs g;
  g.B = d->B; g.H = d->H; g.W = d->W;
  for (int s = 0; s < 4; ++s) {
    g.disp[s] = d->disp[s]; g.color[s] = d->color[0][s]; g.gd_up[s] = ba.gd_up[s];
    g.gdisp[s] = grad_disp[s];
  }
  g.mean = v.mean; g.stats = v.stats; g.gout = grad_loss; g.smooth_w = d->smoothness;
  disp_grad_kernel<<<dim3(64, d->B * 4), 256, 0, st>>>(g);
  FD_CHECK_LAUNCH();
  int tiles = fd::cdiv(d->W, TX) * fd::cdiv(d->H, TY);
  pose_grad_kernel<<<d->B, 128, 0, st>>>(v.partial_dP, tiles, d->K, grad_T0, grad_T1);
  FD_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
