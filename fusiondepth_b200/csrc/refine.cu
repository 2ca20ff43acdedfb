// Stage-2 (refiner.py) pseudo-3D pack and masked median.
//
// Replaces the per-scale ATen sequence of Refiner.process_batch (reference refiner.py:316-346, default flags
// refine_a0='true', catxy='true') and layers.Cat_xy (layers.py:165-201):
//
//   for scale s:  disp = maxpool2x2^s(disp_0)                          (cumulative, ceil mode)
//                 depth = 1 / (min_disp + range * bilinear(disp -> HxW))
//                 ratio = median(beam[mask] * 100) / median(depth[mask])   mask = beam > 0 inside the crop;
//                                                                          ONE scalar per batch, lower median
//                 depth *= ratio
//                 scaled_disp = (bilinear(1/depth -> hxw) - 0.01) / 9.9
//                 two_cha     = maxpool2x2^s(two_cha)
//                 xyz         = Cat_xy(maxpool2x2^s(depth), inv_K_s) = depth * (inv_K[:3,:3] @ [x,y,1]) -> x/30, y/2, (z-40)/40
//                 out_s       = [scaled_disp | xyz | two_cha]            6 channels
//
// Everything here runs under no_grad in the reference (the stage-1 nets are frozen), so there is no backward.
// Kernels: (1) pooled disparity pyramid, (2) compaction of the masked pixel indices (shared by all five
// medians), (3) five radix selects (one CTA each, depth values recomputed on the fly from the pooled
// disparity), (4) the pack itself, one thread per output pixel, NHWC output for the decoder's gather.
#include "common.cuh"
#include "../../include/fusiondepth_b200.h"

namespace {

__device__ __forceinline__ float bilerp_up(const float* __restrict__ d, int y, int x, int h, int w, int H, int W) {
  // F.interpolate(bilinear, align_corners=False), H/h = W/w = 2^s
  if (h == H && w == W) return d[(long)y * w + x];
  float sy = (float)h / (float)H, sx = (float)w / (float)W;
  float fy = sy * ((float)y + 0.5f) - 0.5f, fx = sx * ((float)x + 0.5f) - 0.5f;
  if (fy < 0.f) fy = 0.f;
  if (fx < 0.f) fx = 0.f;
  int y0 = min((int)fy, h - 1), x0 = min((int)fx, w - 1);
  int y1 = y0 + (y0 < h - 1 ? 1 : 0), x1 = x0 + (x0 < w - 1 ? 1 : 0);
  float ly = fy - (float)y0, lx = fx - (float)x0;
  float hy = 1.f - ly, hx = 1.f - lx;
  float t0 = hx * d[(long)y0 * w + x0] + lx * d[(long)y0 * w + x1];
  float t1 = hx * d[(long)y1 * w + x0] + lx * d[(long)y1 * w + x1];
  return hy * t0 + ly * t1;
}

__device__ __forceinline__ float to_depth(float d, float min_disp, float range) {
  return __fdiv_rn(1.0f, __fadd_rn(min_disp, __fmul_rn(range, d)));
}

// ---- (1) pooled disparity pyramid: P_s[y][x] = max over the 2^s x 2^s block of disp_0 -----------------
__global__ void pool_pyramid_kernel(const float* __restrict__ d0, float* p1, float* p2, float* p3, int B, int H,
                                    int W) {
  // one thread per scale-1 pixel; scales 2 and 3 are folded from registers of neighbouring threads via
  // a second read (the tensors are tiny: three passes of <= B*H*W/4 reads)
  const long n1 = (long)B * (H / 2) * (W / 2);
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n1; i += (long)gridDim.x * blockDim.x) {
    int w1 = W / 2, h1 = H / 2;
    int x = (int)(i % w1), y = (int)((i / w1) % h1), b = (int)(i / ((long)w1 * h1));
    const float* s = d0 + ((long)b * H + 2 * y) * W + 2 * x;
    p1[i] = fmaxf(fmaxf(s[0], s[1]), fmaxf(s[W], s[W + 1]));
  }
  const long n2 = (long)B * (H / 4) * (W / 4);
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n2; i += (long)gridDim.x * blockDim.x) {
    int w2 = W / 4, h2 = H / 4;
    int x = (int)(i % w2), y = (int)((i / w2) % h2), b = (int)(i / ((long)w2 * h2));
    float m = -INFINITY;
    for (int dy = 0; dy < 4; ++dy)
      for (int dx = 0; dx < 4; ++dx) m = fmaxf(m, d0[((long)b * H + 4 * y + dy) * W + 4 * x + dx]);
    p2[i] = m;
  }
  const long n3 = (long)B * (H / 8) * (W / 8);
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n3; i += (long)gridDim.x * blockDim.x) {
    int w3 = W / 8, h3 = H / 8;
    int x = (int)(i % w3), y = (int)((i / w3) % h3), b = (int)(i / ((long)w3 * h3));
    float m = -INFINITY;
    for (int dy = 0; dy < 8; ++dy)
      for (int dx = 0; dx < 8; ++dx) m = fmaxf(m, d0[((long)b * H + 8 * y + dy) * W + 8 * x + dx]);
    p3[i] = m;
  }
}

// ---- (2) indices of the masked pixels (mask_src > 0 inside the crop window) ---------------------------
__global__ void mask_compact_kernel(const float* __restrict__ mask_src, int B, int H, int W, int y0, int y1, int x0,
                                    int x1, float mask_lo, float mask_hi, int* __restrict__ count,
                                    int* __restrict__ idx) {
  const int cw = x1 - x0, ch = y1 - y0;
  const long n = (long)B * ch * cw;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    int x = x0 + (int)(i % cw), y = y0 + (int)((i / cw) % ch), b = (int)(i / ((long)cw * ch));
    long o = ((long)b * H + y) * W + x;
    const float mv = mask_src[o];
    bool on = mv > mask_lo && mv < mask_hi;
    // warp-aggregated append
    unsigned m = __ballot_sync(__activemask(), on);
    if (on) {
      int lane = threadIdx.x & 31;
      int leader = __ffs(m) - 1;
      int base = 0;
      if (lane == leader) base = atomicAdd(count, __popc(m));
      base = __shfl_sync(m, base, leader);
      idx[base + __popc(m & ((1u << lane) - 1u))] = (int)o;
    }
  }
}

// ---- (3) lower median (torch.median semantics: element (n-1)/2 of the sorted values) ------------------
struct MedianArgs {
  const int* count;
  const int* idx;
  int nsrc;
  // source j: kind 0 = x[idx] * scale ; kind 1 = depth(bilinear(pyr -> HxW)) at idx ; kind 2 = 1 / x[idx]
  int avg_even;        // numpy.median semantics: mean of the two middle elements when the count is even
  const float* src[5];
  int kind[5];
  int h[5], w[5];
  float scale[5];
  int H, W;
  float min_disp, range;
  float* out;   // [nsrc]
};

__device__ __forceinline__ float median_value(const MedianArgs& a, int j, int o) {
  if (a.kind[j] == 0) return __fmul_rn(a.src[j][o], a.scale[j]);
  if (a.kind[j] == 2) return __fdiv_rn(1.0f, a.src[j][o]);
  if (a.kind[j] == 3) return fminf(fmaxf(a.src[j][o], a.min_disp), a.range);      // clamp(x, lo, hi)
  const int HW = a.H * a.W;
  int b = o / HW, r = o - b * HW, y = r / a.W, x = r - y * a.W;
  float d = bilerp_up(a.src[j] + (long)b * a.h[j] * a.w[j], y, x, a.h[j], a.w[j], a.H, a.W);
  return to_depth(d, a.min_disp, a.range);
}

__device__ __forceinline__ unsigned order_key(float v) {
  unsigned u = __float_as_uint(v);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_value(unsigned k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

__global__ void __launch_bounds__(1024) radix_median_kernel(MedianArgs a) {
  __shared__ unsigned hist[256];
  __shared__ unsigned s_prefix, s_k;
  const int j = blockIdx.x;
  const int n = *a.count;
  if (n <= 0) {
    if (threadIdx.x == 0) a.out[j] = __int_as_float(0x7fc00000);
    return;
  }
  float result[2];
  const int nsel = (a.avg_even && (n & 1) == 0) ? 2 : 1;
  for (int sel = 0; sel < nsel; ++sel) {
    __syncthreads();
    if (threadIdx.x == 0) { s_prefix = 0u; s_k = (unsigned)(sel == 0 ? (n - 1) / 2 : n / 2); }
    unsigned mask = 0u;
    for (int pass = 3; pass >= 0; --pass) {
      if (threadIdx.x < 256) hist[threadIdx.x] = 0u;
      __syncthreads();
      const unsigned prefix = s_prefix;
      for (int i = threadIdx.x; i < n; i += blockDim.x) {
        unsigned key = order_key(median_value(a, j, a.idx[i]));
        if ((key & mask) == prefix) atomicAdd(&hist[(key >> (8 * pass)) & 255u], 1u);
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        unsigned k = s_k, cum = 0u;
        int bkt = 0;
        for (; bkt < 256; ++bkt) {
          if (cum + hist[bkt] > k) break;
          cum += hist[bkt];
        }
        s_k = k - cum;
        s_prefix = prefix | ((unsigned)bkt << (8 * pass));
      }
      mask |= 0xffu << (8 * pass);
      __syncthreads();
    }
    result[sel] = key_value(s_prefix);
  }
  if (threadIdx.x == 0)
    a.out[j] = nsel == 2 ? __fmul_rn(__fadd_rn(result[0], result[1]), 0.5f) : result[0];
}

// ---- (4) the pack ---------------------------------------------------------------------------------------
struct PackArgs {
  int B, H, W;
  const float* pyr[4];      // pooled disparity, scale s: [B, H>>s, W>>s]
  const float* two_cha;     // [B,2,H,W] NCHW
  const float* inv_K[4];    // [B,4,4]
  const float* med;         // [5]: beam*100 median, depth medians s=0..3
  float min_disp, range;
  float* out[4];            // [B, H>>s, W>>s, 6] NHWC
  float* ratios;            // [4]
};

__global__ void pack_kernel(PackArgs a) {
  const long HW = (long)a.H * a.W;
  const long n0 = a.B * HW, n1 = n0 / 4, n2 = n0 / 16, n3 = n0 / 64;
  const long total = n0 + n1 + n2 + n3;
  if (blockIdx.x == 0 && threadIdx.x < 4) a.ratios[threadIdx.x] = __fdiv_rn(a.med[0], a.med[1 + threadIdx.x]);
  for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
    // coarse scales first: their threads carry the most work
    int s;
    long i;
    if (t < n3) { s = 3; i = t; }
    else if (t < n3 + n2) { s = 2; i = t - n3; }
    else if (t < n3 + n2 + n1) { s = 1; i = t - n3 - n2; }
    else { s = 0; i = t - n3 - n2 - n1; }
    const int r = 1 << s, h = a.H >> s, w = a.W >> s;
    const int x = (int)(i % w), y = (int)((i / w) % h), b = (int)(i / ((long)w * h));
    const float ratio = __fdiv_rn(a.med[0], a.med[1 + s]);
    const float* pyr = a.pyr[s] + (long)b * h * w;
    float dmax = -INFINITY, c2 = 0.f, c4 = 0.f, c5 = 0.f;
    float inv[2][2];
    const int cy = r / 2 - 1, cx = r / 2 - 1;           // the 2x2 centre of the block (s >= 1)
    const float* tc = a.two_cha + (long)b * 2 * HW;
    float t0 = 0.f, t1 = 0.f;
    for (int dy = 0; dy < r; ++dy)
      for (int dx = 0; dx < r; ++dx) {
        const int Y = y * r + dy, X = x * r + dx;
        float D = to_depth(bilerp_up(pyr, Y, X, h, w, a.H, a.W), a.min_disp, a.range);
        D = __fmul_rn(D, ratio);                          // depth *= ratio
        dmax = fmaxf(dmax, D);
        if (s == 0) {
          inv[0][0] = __fdiv_rn(1.0f, D);
        } else if (dy >= cy && dy <= cy + 1 && dx >= cx && dx <= cx + 1) {
          inv[dy - cy][dx - cx] = __fdiv_rn(1.0f, D);
        }
        const float v0 = tc[(long)Y * a.W + X], v1 = tc[HW + (long)Y * a.W + X];
        if (dy == 0 && dx == 0) { t0 = v0; t1 = v1; }
        else { t0 = fmaxf(t0, v0); t1 = fmaxf(t1, v1); }
      }
    float low;
    if (s == 0) {
      low = inv[0][0];
    } else {
      // bilinear downsample, align_corners=False: source = r*dst + r/2 - 0.5, both lambdas exactly 0.5
      float ta = __fadd_rn(__fmul_rn(0.5f, inv[0][0]), __fmul_rn(0.5f, inv[0][1]));
      float tb = __fadd_rn(__fmul_rn(0.5f, inv[1][0]), __fmul_rn(0.5f, inv[1][1]));
      low = __fadd_rn(__fmul_rn(0.5f, ta), __fmul_rn(0.5f, tb));
    }
    const float sdisp = __fdiv_rn(__fadd_rn(low, -0.01f), 9.9f);
    // Cat_xy at scale s: pixel coordinates of the scale-s grid, inv_K of that scale
    const float* ik = a.inv_K[s] + b * 16;
    const float fx = (float)x, fy = (float)y;
    float cam[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      float c = ik[k * 4 + 0] * fx;
      c = fmaf(ik[k * 4 + 1], fy, c);
      c = fmaf(ik[k * 4 + 2], 1.0f, c);
      cam[k] = __fmul_rn(dmax, c);
    }
    c2 = __fdiv_rn(cam[0], 30.0f);
    c4 = __fdiv_rn(cam[1], 2.0f);
    c5 = __fdiv_rn(__fadd_rn(cam[2], -40.0f), 40.0f);
    float* o = a.out[s] + i * 6;
    o[0] = sdisp; o[1] = c2; o[2] = c4; o[3] = c5; o[4] = t0; o[5] = t1;
  }
}

struct Ws {
  int* count;
  float* med;        // [8]
  float* pyr[4];     // pyr[0] aliases disp0
  int* idx;
};

size_t ws_floats(int B, int H, int W, int ch, int cw) {
  long hw = (long)B * H * W;
  return 16 + hw / 4 + hw / 16 + hw / 64 + (long)B * ch * cw;
}

Ws ws_view(void* ws, int B, int H, int W) {
  Ws v;
  float* p = (float*)ws;
  v.count = (int*)p;
  v.med = p + 8;
  p += 16;
  long hw = (long)B * H * W;
  v.pyr[0] = nullptr;
  v.pyr[1] = p; p += hw / 4;
  v.pyr[2] = p; p += hw / 16;
  v.pyr[3] = p; p += hw / 64;
  v.idx = (int*)p;
  return v;
}

}  // namespace

extern "C" {

size_t fd_refine_pack_workspace_bytes(int B, int H, int W) {
  return ws_floats(B, H, W, H, W) * sizeof(float);
}

int fd_refine_pack(const float* disp0, const float* beam, const float* two_cha, const float* const inv_K[4], int B,
                   int H, int W, int crop_y0, int crop_y1, int crop_x0, int crop_x1, float min_depth,
                   float max_depth, float* const out[4], float* ratios, void* workspace, void* stream) {
  FD_REQUIRE(B > 0 && H % 8 == 0 && W % 8 == 0, "fd_refine_pack: H, W must be multiples of 8 (got %dx%d)", H, W);
  crop_y0 = crop_y0 < 0 ? 0 : crop_y0; crop_x0 = crop_x0 < 0 ? 0 : crop_x0;
  crop_y1 = crop_y1 > H ? H : crop_y1; crop_x1 = crop_x1 > W ? W : crop_x1;
  FD_REQUIRE(crop_y1 > crop_y0 && crop_x1 > crop_x0, "fd_refine_pack: empty crop window");
  cudaStream_t st = (cudaStream_t)stream;
  Ws v = ws_view(workspace, B, H, W);
  cudaError_t e = cudaMemsetAsync(v.count, 0, 16 * sizeof(float), st);
  FD_REQUIRE(e == cudaSuccess, "fd_refine_pack: memset failed: %s", cudaGetErrorString(e));
  const long hw = (long)B * H * W;
  pool_pyramid_kernel<<<fd::cdiv(hw / 4, 256), 256, 0, st>>>(disp0, v.pyr[1], v.pyr[2], v.pyr[3], B, H, W);
  FD_CHECK_LAUNCH();
  const long nc = (long)B * (crop_y1 - crop_y0) * (crop_x1 - crop_x0);
  mask_compact_kernel<<<fd::cdiv(nc, 256), 256, 0, st>>>(beam, B, H, W, crop_y0, crop_y1, crop_x0, crop_x1, 0.f,
                                                         INFINITY, v.count, v.idx);
  FD_CHECK_LAUNCH();
  const float min_disp = (float)(1.0 / (double)max_depth);
  const float range = (float)(1.0 / (double)min_depth - 1.0 / (double)max_depth);
  MedianArgs m;
  m.count = v.count; m.idx = v.idx; m.nsrc = 5; m.H = H; m.W = W; m.min_disp = min_disp; m.range = range;
  m.avg_even = 0;
  m.out = v.med;
  m.src[0] = beam; m.kind[0] = 0; m.scale[0] = 100.0f; m.h[0] = H; m.w[0] = W;
  for (int s = 0; s < 4; ++s) {
    m.src[1 + s] = s == 0 ? disp0 : v.pyr[s];
    m.kind[1 + s] = 1; m.scale[1 + s] = 1.f; m.h[1 + s] = H >> s; m.w[1 + s] = W >> s;
  }
  radix_median_kernel<<<5, 1024, 0, st>>>(m);
  FD_CHECK_LAUNCH();
  PackArgs p;
  p.B = B; p.H = H; p.W = W; p.two_cha = two_cha; p.med = v.med; p.min_disp = min_disp; p.range = range;
  p.ratios = ratios;
  for (int s = 0; s < 4; ++s) {
    p.pyr[s] = s == 0 ? disp0 : v.pyr[s];
    p.inv_K[s] = inv_K[s];
    p.out[s] = out[s];
  }
  const long total = hw + hw / 4 + hw / 16 + hw / 64;
  pack_kernel<<<fd::cdiv(total, 128), 128, 0, st>>>(p);
  FD_CHECK_LAUNCH();
  return 0;
}

size_t fd_masked_median_workspace_bytes(int B, int H, int W) { return (16 + (size_t)B * H * W) * sizeof(float); }

int fd_masked_median(const float* x, const float* mask_src, int B, int H, int W, int y0, int y1, int x0, int x1,
                     float scale, float* out, void* workspace, void* stream) {
  y0 = y0 < 0 ? 0 : y0; x0 = x0 < 0 ? 0 : x0;
  y1 = y1 > H ? H : y1; x1 = x1 > W ? W : x1;
  FD_REQUIRE(B > 0 && y1 > y0 && x1 > x0, "fd_masked_median: empty window");
  cudaStream_t st = (cudaStream_t)stream;
  int* count = (int*)workspace;
  int* idx = count + 16;
  cudaError_t e = cudaMemsetAsync(count, 0, 16 * sizeof(int), st);
  FD_REQUIRE(e == cudaSuccess, "fd_masked_median: memset failed: %s", cudaGetErrorString(e));
  const long nc = (long)B * (y1 - y0) * (x1 - x0);
  mask_compact_kernel<<<fd::cdiv(nc, 256), 256, 0, st>>>(mask_src, B, H, W, y0, y1, x0, x1, 0.f, INFINITY, count, idx);
  FD_CHECK_LAUNCH();
  MedianArgs m;
  m.avg_even = 0;
  m.count = count; m.idx = idx; m.nsrc = 1; m.H = H; m.W = W; m.min_disp = 0.f; m.range = 0.f; m.out = out;
  m.src[0] = x; m.kind[0] = 0; m.scale[0] = scale; m.h[0] = H; m.w[0] = W;
  radix_median_kernel<<<1, 1024, 0, st>>>(m);
  FD_CHECK_LAUNCH();
  return 0;
}

// ---- depth error metrics ------------------------------------------------------------------------------
// layers.compute_depth_errors / evaluate_depth.compute_errors (layers.py:284-302, evaluate_depth.py:42-60) behind
// the masking, median scaling and clamping of Trainer.compute_depth_losses (trainer.py:598-630) and of the
// evaluate_depth.py loop (evaluate_depth.py:344-378, 470-478): three launches, nothing leaves the device.
//   mask  = gt > mask_lo && gt < mask_hi inside rows [y0,y1) x cols [x0,x1)
//   pred  = pred_is_disp ? 1 / p : clamp(p, pre_min, pre_max)          (the tensor is already at gt's size)
//   ratio = median(gt[mask]) / median(pred[mask])  (numpy_median: mean of the middle two for even counts;
//           otherwise torch.median's lower median); pred = clamp(pred * ratio, pred_min, pred_max)
//   out[0..6] = abs_rel, sq_rel, rmse, rmse_log, a1, a2, a3 ; out[7] = number of masked pixels ; out[8] = ratio
int fd_depth_errors(const float* gt, const float* pred, int B, int H, int W, int y0, int y1, int x0, int x1,
                    float mask_lo, float mask_hi, int pred_is_disp, float pre_min, float pre_max,
                    int median_scaling, int numpy_median, float pred_min, float pred_max, float* out,
                    void* workspace, void* stream);

namespace {
struct ErrArgs {
  const float* gt;
  const float* pred;
  const int* count;
  const int* idx;
  const float* med;       // [2]: median gt, median pred
  int pred_is_disp, median_scaling;
  float pre_min, pre_max, pred_min, pred_max;
  double* acc;            // [8]
};

__device__ __forceinline__ float err_pred(const ErrArgs& a, int o) {
  const float p = a.pred[o];
  return a.pred_is_disp ? __fdiv_rn(1.0f, p) : fminf(fmaxf(p, a.pre_min), a.pre_max);
}

__global__ void depth_errors_kernel(ErrArgs a) {
  __shared__ double red[7 * 32];
  const int n = *a.count;
  const float ratio = a.median_scaling ? __fdiv_rn(a.med[0], a.med[1]) : 1.0f;
  double s[7] = {0, 0, 0, 0, 0, 0, 0};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int o = a.idx[i];
    const float g = a.gt[o];
    float p = __fmul_rn(err_pred(a, o), ratio);
    p = fminf(fmaxf(p, a.pred_min), a.pred_max);
    const float d = __fadd_rn(g, -p);
    const float th = fmaxf(__fdiv_rn(g, p), __fdiv_rn(p, g));
    const float dl = __fadd_rn(logf(g), -logf(p));
    s[0] += (double)__fdiv_rn(fabsf(d), g);
    s[1] += (double)__fdiv_rn(__fmul_rn(d, d), g);
    s[2] += (double)__fmul_rn(d, d);
    s[3] += (double)__fmul_rn(dl, dl);
    s[4] += th < 1.25f ? 1.0 : 0.0;
    s[5] += th < 1.5625f ? 1.0 : 0.0;
    s[6] += th < 1.953125f ? 1.0 : 0.0;
  }
  fd::block_sum<7>(s, red);
  if (threadIdx.x == 0)
    for (int k = 0; k < 7; ++k) atomicAdd(a.acc + k, s[k]);
}

__global__ void depth_errors_finalize_kernel(const double* __restrict__ acc, const int* __restrict__ count,
                                             const float* __restrict__ med, int median_scaling, float* out) {
  const double n = (double)*count;
  out[0] = (float)(acc[0] / n);
  out[1] = (float)(acc[1] / n);
  out[2] = (float)sqrt(acc[2] / n);
  out[3] = (float)sqrt(acc[3] / n);
  out[4] = (float)(acc[4] / n);
  out[5] = (float)(acc[5] / n);
  out[6] = (float)(acc[6] / n);
  out[7] = (float)n;
  out[8] = median_scaling ? __fdiv_rn(med[0], med[1]) : 1.0f;
}
}  // namespace

size_t fd_depth_errors_workspace_bytes(int B, int H, int W) { return (64 + (size_t)B * H * W) * sizeof(float); }

int fd_depth_errors(const float* gt, const float* pred, int B, int H, int W, int y0, int y1, int x0, int x1,
                    float mask_lo, float mask_hi, int pred_is_disp, float pre_min, float pre_max,
                    int median_scaling, int numpy_median, float pred_min, float pred_max, float* out,
                    void* workspace, void* stream) {
  y0 = y0 < 0 ? 0 : y0; x0 = x0 < 0 ? 0 : x0;
  y1 = y1 > H ? H : y1; x1 = x1 > W ? W : x1;
  FD_REQUIRE(B > 0 && y1 > y0 && x1 > x0, "fd_depth_errors: empty window");
  cudaStream_t st = (cudaStream_t)stream;
  // workspace: [count (16 ints)] [med (16 floats)] [acc (8 doubles = 16 floats... 32 floats reserved)] [idx]
  int* count = (int*)workspace;
  float* med = (float*)workspace + 16;
  double* acc = (double*)((float*)workspace + 32);
  int* idx = (int*)workspace + 64;
  cudaError_t e = cudaMemsetAsync(workspace, 0, 64 * sizeof(float), st);
  FD_REQUIRE(e == cudaSuccess, "fd_depth_errors: memset failed: %s", cudaGetErrorString(e));
  const long nc = (long)B * (y1 - y0) * (x1 - x0);
  mask_compact_kernel<<<fd::cdiv(nc, 256), 256, 0, st>>>(gt, B, H, W, y0, y1, x0, x1, mask_lo, mask_hi, count, idx);
  FD_CHECK_LAUNCH();
  if (median_scaling) {
    MedianArgs m;
    m.count = count; m.idx = idx; m.nsrc = 2; m.H = H; m.W = W; m.min_disp = 0.f; m.range = 0.f; m.out = med;
    m.avg_even = numpy_median;
    m.src[0] = gt; m.kind[0] = 0; m.scale[0] = 1.f; m.h[0] = H; m.w[0] = W;
    m.src[1] = pred; m.kind[1] = pred_is_disp ? 2 : 3; m.scale[1] = 1.f; m.h[1] = H; m.w[1] = W;
    m.min_disp = pre_min; m.range = pre_max;              // kind 3: clamp(x, min_disp, range)
    radix_median_kernel<<<2, 1024, 0, st>>>(m);
    FD_CHECK_LAUNCH();
  }
  ErrArgs a;
  a.gt = gt; a.pred = pred; a.count = count; a.idx = idx; a.med = med; a.pred_is_disp = pred_is_disp;
  a.median_scaling = median_scaling; a.pre_min = pre_min; a.pre_max = pre_max; a.pred_min = pred_min;
  a.pred_max = pred_max; a.acc = acc;
  depth_errors_kernel<<<148, 256, 0, st>>>(a);
  FD_CHECK_LAUNCH();
  depth_errors_finalize_kernel<<<1, 1, 0, st>>>(acc, count, med, median_scaling, out);
  FD_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
