// Sparse-LiDAR kernels: velodyne -> image scatter with the reference's exact duplicate
// semantics, pad + 2x2 ceil max-pool + /100, and the 2-channel (expanded depth, confidence)
// gather stencil.
//
// Replaces (bit-exact): kitti_utils.generate_depth_map  (reference kitti_utils.py:40-102),
//   F.max_pool2d(.,2,ceil_mode=True)->float32->/100      (kitti_dataset.py:105-107,
//                                                         mono_dataset.py:194-198),
//   get_4beam_2channel                                   (gen2channel.py:60-117).
//
// Parallel formulation: per image pixel keep min z (ordered-bits
// 64-bit atomicMin), first and last point index (file order); the value of a pixel is its
// min z, except for the reference's sub2ind quirk where pixels (v,0) and (v-1,W-1) share a
// duplicate key: the pixel holding the earlier first point gets the joint min, the other
// keeps the z of its own last-listed point.
//
// All work is HBM/latency bound integer + fp64 arithmetic; no tensor cores involved.
#include "common.cuh"
#include "../../include/fusiondepth_b200.h"

namespace {

struct PixWs {            // 16 B per image pixel
  unsigned long long minz;  // ordered bits of the smallest z
  unsigned int first;       // smallest point index
  int last;                 // largest point index
};

__device__ __forceinline__ unsigned long long ord_bits(double z) {
  unsigned long long b = (unsigned long long)__double_as_longlong(z);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double ord_inv(unsigned long long k) {
  unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double((long long)b);
}

__device__ __forceinline__ double dot4(const double* P, double x, double y, double z) {
  // fma chain in k order with the homogeneous coordinate forced to 1 (kitti_utils.py:10,64)
  double acc = __dmul_rn(P[0], x);
  acc = __fma_rn(P[1], y, acc);
  acc = __fma_rn(P[2], z, acc);
  acc = __fma_rn(P[3], 1.0, acc);
  return acc;
}

__global__ void lidar_init_kernel(PixWs* ws, long n) {
  long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i < n) {
    PixWs w;
    w.minz = ~0ull;
    w.first = 0xffffffffu;
    w.last = -1;
    ws[i] = w;
  }
}

// one thread per point; frame = blockIdx.y
__global__ void lidar_scatter_kernel(const float4* __restrict__ pts, const int* __restrict__ offsets,
                                     const double* __restrict__ Pall, int W_im, int H_im,
                                     int vel_depth, PixWs* __restrict__ ws) {
  const int f = blockIdx.y;
  const int beg = offsets[f], end = offsets[f + 1];
  const double* P = Pall + 12 * f;
  PixWs* w = ws + (long)f * W_im * H_im;
  for (int i = beg + blockIdx.x * blockDim.x + threadIdx.x; i < end; i += gridDim.x * blockDim.x) {
    float4 p = pts[i];
    double x = p.x, y = p.y, z = p.z;
    if (!(x >= 0)) continue;
    double p0 = dot4(P, x, y, z), p1 = dot4(P + 4, x, y, z), p2 = dot4(P + 8, x, y, z);
    double u = __ddiv_rn(p0, p2), v = __ddiv_rn(p1, p2);
    double zz = vel_depth ? x : p2;
    u = rint(u) - 1.0;
    v = rint(v) - 1.0;
    if (!(u >= 0 && v >= 0 && u < W_im && v < H_im)) continue;
    PixWs* c = w + (long)(int)v * W_im + (int)u;
    atomicMin(&c->minz, ord_bits(zz));
    atomicMin(&c->first, (unsigned int)(i - beg));
    atomicMax(&c->last, i - beg);
  }
}

__device__ __forceinline__ double point_depth(const float4* pts, int idx, const double* P,
                                              int vel_depth) {
  float4 p = pts[idx];
  if (vel_depth) return (double)p.x;
  return dot4(P + 8, (double)p.x, (double)p.y, (double)p.z);
}

// value of image pixel (v,u) after the duplicate fix and the <0 clamp
__device__ double resolve_pixel(const PixWs* w, const float4* pts, const double* P, int W_im,
                                int H_im, int vel_depth, int v, int u) {
  PixWs c = w[(long)v * W_im + u];
  if (c.first == 0xffffffffu) return 0.0;
  double val = ord_inv(c.minz);
  int pv = -1, pu = -1;
  if (W_im > 1) {
    if (u == 0 && v >= 1) { pv = v - 1; pu = W_im - 1; }
    else if (u == W_im - 1 && v + 1 < H_im) { pv = v + 1; pu = 0; }
  }
  if (pv >= 0) {
    PixWs q = w[(long)pv * W_im + pu];
    if (q.first != 0xffffffffu) {
      if (c.first < q.first) {
        double o = ord_inv(q.minz);
        val = o < val ? o : val;
      } else {
        val = point_depth(pts, c.last, P, vel_depth);
      }
    }
  }
  return val < 0 ? 0.0 : val;
}

// one thread per padded-map pixel: writes the fp64 map generate_depth_map returns
__global__ void lidar_resolve_kernel(const PixWs* __restrict__ ws, const float4* __restrict__ pts,
                                     const int* __restrict__ offsets, const double* __restrict__ Pall,
                                     int W_im, int H_im, int vel_depth, int out_h, int out_w,
                                     int row_off, int col_off, double* __restrict__ out) {
  const int f = blockIdx.z;
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  int r = blockIdx.y;
  if (c >= out_w) return;
  int v = r - row_off, u = c - col_off;
  double val = 0.0;
  if (v >= 0 && v < H_im && u >= 0 && u < W_im)
    val = resolve_pixel(ws + (long)f * W_im * H_im, pts + offsets[f], Pall + 12 * f, W_im, H_im,
                        vel_depth, v, u);
  out[((long)f * out_h + r) * out_w + c] = val;
}

// max_pool2d(2, ceil_mode=True) on fp64 -> float32 -> /100.0f
__global__ void lidar_pool_kernel(const double* __restrict__ in, int H, int W, int Ho, int Wo,
                                  float* __restrict__ out) {
  const int f = blockIdx.z;
  int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
  if (j >= Wo) return;
  const double* m = in + (long)f * H * W;
  double best = -INFINITY;
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      int r = 2 * i + a, c = 2 * j + b;
      if (r < H && c < W) {
        double v = m[(long)r * W + c];
        if (v > best || v != v) best = v;
      }
    }
  out[((long)f * Ho + i) * Wo + j] = __fdiv_rn((float)best, 100.0f);
}

// gather form of get_4beam_2channel: sources are non-zero pixels inside the window
__global__ void two_channel_kernel(const float* __restrict__ fb, int H, int W, int r0, int r1,
                                   int c0, int c1, float* __restrict__ out) {
  const int f = blockIdx.z;
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= W) return;
  const float* m = fb + (long)f * H * W;
  auto src = [&](int yy, int xx) -> float {
    if (yy < r0 || yy >= r1 || xx < c0 || xx >= c1) return 0.0f;
    return m[(long)yy * W + xx];
  };
  float e = 0.0f, cf = 0.0f;
  float v = src(y, x);
  if (v != 0.0f) {
    e = v; cf = 1.0f;
  } else {
    // confidence 1/2: (y-1,x), (y+1,x) in raster order of the sources
    float sum = 0.0f; int n = 0;
    float a = src(y - 1, x), b = src(y + 1, x);
    if (a != 0.0f) { sum = a; n = 1; }
    if (b != 0.0f) { sum = n ? __fadd_rn(sum, b) : b; ++n; }
    if (n) {
      e = __fdiv_rn(sum, (float)n); cf = 0.5f;
    } else {
      const int dy[6] = {-2, -1, -1, 1, 1, 2};
      const int dx[6] = {0, -1, 1, -1, 1, 0};
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        float s = src(y + dy[k], x + dx[k]);
        if (s != 0.0f) { sum = n ? __fadd_rn(sum, s) : s; ++n; }
      }
      if (n) { e = __fdiv_rn(sum, (float)n); cf = (float)(1.0 / 3.0); }
    }
  }
  float* o = out + (long)f * 2 * H * W;
  o[(long)y * W + x] = e;
  o[(long)H * W + (long)y * W + x] = cf;
}

}  // namespace

extern "C" {

size_t fd_lidar_workspace_bytes(int n_frames, int W_im, int H_im) {
  return (size_t)n_frames * W_im * H_im * sizeof(PixWs);
}

int fd_lidar_depth_map(const float* points, const int* offsets, int n_frames, int n_points_total,
                       const double* P, int W_im, int H_im, int vel_depth, int shape_h,
                       int shape_w, double* depth_out, int out_h, int out_w, void* workspace,
                       void* stream) {
  FD_REQUIRE(n_frames > 0 && W_im > 0 && H_im > 0, "fd_lidar_depth_map: bad dims");
  cudaStream_t st = (cudaStream_t)stream;
  int row_off = 0, col_off = 0, eh = H_im, ew = W_im;
  if (shape_h > 0) {
    int ypad = shape_h > H_im ? shape_h - H_im : H_im - shape_h;
    int xpad = shape_w - W_im;
    FD_REQUIRE(xpad >= 0, "fd_lidar_depth_map: shape_w < image width");
    int crop = shape_h < H_im ? 2 : 0;
    row_off = ypad - crop;
    col_off = xpad / 2;
    eh = H_im + ypad - crop;
    ew = shape_w;
  }
  FD_REQUIRE(eh == out_h && ew == out_w, "fd_lidar_depth_map: output is %dx%d, expected %dx%d",
             out_h, out_w, eh, ew);
  PixWs* ws = (PixWs*)workspace;
  long npix = (long)n_frames * W_im * H_im;
  lidar_init_kernel<<<fd::cdiv(npix, 256), 256, 0, st>>>(ws, npix);
  FD_CHECK_LAUNCH();
  if (n_points_total > 0) {
    int per = fd::cdiv(fd::cdiv(n_points_total, n_frames) + 1, 256);
    if (per < 1) per = 1;
    if (per > 1024) per = 1024;
    lidar_scatter_kernel<<<dim3(per, n_frames), 256, 0, st>>>((const float4*)points, offsets, P, W_im,
                                                             H_im, vel_depth, ws);
    FD_CHECK_LAUNCH();
  }
  lidar_resolve_kernel<<<dim3(fd::cdiv(out_w, 128), out_h, n_frames), 128, 0, st>>>(
      ws, (const float4*)points, offsets, P, W_im, H_im, vel_depth, out_h, out_w, row_off, col_off,
      depth_out);
  FD_CHECK_LAUNCH();
  return 0;
}

int fd_lidar_pool_scale(const double* depth, int n_frames, int H, int W, float* out, void* stream) {
  int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
  lidar_pool_kernel<<<dim3(fd::cdiv(Wo, 128), Ho, n_frames), 128, 0, (cudaStream_t)stream>>>(
      depth, H, W, Ho, Wo, out);
  FD_CHECK_LAUNCH();
  return 0;
}

int fd_two_channel(const float* fourbeam, int n_frames, int H, int W, int r0, int r1, int c0, int c1,
                   float* out, void* stream) {
  two_channel_kernel<<<dim3(fd::cdiv(W, 128), H, n_frames), 128, 0, (cudaStream_t)stream>>>(
      fourbeam, H, W, r0, r1, c0, c1, out);
  FD_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
