// On-device colour side of the reference's data producer: what
// datasets/mono_dataset.py:85-104, 156-206 does per frame with PIL / torchvision on the host -- horizontal flip,
// the Lanczos ("ANTIALIAS") pyramid where each scale is resized from the previous one, ColorJitter, ToTensor --
// as bit-exact integer / IEEE kernels on uint8 HWC images resident in HBM.  The real trainer is DataLoader-bound
// at B200 step rates (4 PIL workers against a 24 ms step).
//
// The arithmetic is Pillow's (Resample.c, Blend.c, Convert.c) and torchvision's `_functional_pil`, restated in
// oracle/data_oracle.py and pinned there against PIL itself; these kernels are bit-identical to that restatement
// (tests/test_gpu_dataprep.py):
//   resize      separable, horizontal pass first, uint8 intermediate, 22-bit fixed-point Lanczos-3 weights
//               (computed on the host in double, fd_lanczos_coeffs), int32 accumulation from 1 << 21
//   brightness / contrast / saturation   blend(degenerate, image, factor) in float32, truncating;
//               L = (19595 R + 38470 G + 7471 B + 0x8000) >> 16, contrast pivots on the rounded mean of L
//   hue         RGB -> HSV -> h += uint8(factor * 255) -> RGB with Convert.c's float / double mix
//   ToTensor    float(u8) / 255, HWC -> CHW
// No fused multiply-adds anywhere (explicit _rn intrinsics): the host code these replace has none.
#include <math.h>

#include "common.cuh"
#include "../../include/fusiondepth_b200.h"

namespace {

constexpr int PRECISION_BITS = 32 - 8 - 2;

__device__ __forceinline__ unsigned char clip8(int v) { return (unsigned char)(v < 0 ? 0 : (v > 255 ? 255 : v)); }

// in [B][H][Win][3] -> out [B][H][Wout][3]; flip[b] != 0 reads the input row mirrored (= PIL transpose(FLIP_LEFT_RIGHT)
// before the resize, mono_dataset.py get_color)
__global__ void resample_h_kernel(const unsigned char* __restrict__ in, unsigned char* __restrict__ out,
                                  const int* __restrict__ bounds, const int* __restrict__ kk, int ksize, int H,
                                  int Win, int Wout, const unsigned char* __restrict__ flip, long total) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int xx = (int)(i % Wout);
    const long row = i / Wout;                                  // b * H + y
    const int b = (int)(row / H);
    const bool fl = flip && flip[b];
    const int x0 = bounds[2 * xx], n = bounds[2 * xx + 1];
    const int* k = kk + (long)xx * ksize;
    const unsigned char* src = in + row * (long)Win * 3;
    int s0 = 1 << (PRECISION_BITS - 1), s1 = s0, s2 = s0;
    for (int x = 0; x < n; ++x) {
      const int xs = fl ? Win - 1 - (x0 + x) : x0 + x;
      const int c = k[x];
      s0 += src[xs * 3 + 0] * c;
      s1 += src[xs * 3 + 1] * c;
      s2 += src[xs * 3 + 2] * c;
    }
    unsigned char* o = out + i * 3;
    o[0] = clip8(s0 >> PRECISION_BITS);
    o[1] = clip8(s1 >> PRECISION_BITS);
    o[2] = clip8(s2 >> PRECISION_BITS);
  }
}

// in [B][Hin][W][3] -> out [B][Hout][W][3]
__global__ void resample_v_kernel(const unsigned char* __restrict__ in, unsigned char* __restrict__ out,
                                  const int* __restrict__ bounds, const int* __restrict__ kk, int ksize, int Hin,
                                  int Hout, int W, long total) {
  const long rowb = (long)W * 3;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long e = i % rowb;                                    // byte within the row (x * 3 + channel)
    const long r = i / rowb;
    const int yy = (int)(r % Hout);
    const long b = r / Hout;
    const int y0 = bounds[2 * yy], n = bounds[2 * yy + 1];
    const int* k = kk + (long)yy * ksize;
    const unsigned char* src = in + (b * Hin + y0) * rowb + e;
    int s = 1 << (PRECISION_BITS - 1);
    for (int y = 0; y < n; ++y) s += src[y * rowb] * k[y];
    out[i] = clip8(s >> PRECISION_BITS);
  }
}

__device__ __forceinline__ unsigned lum(unsigned r, unsigned g, unsigned b) {
  return (r * 19595u + g * 38470u + b * 7471u + 0x8000u) >> 16;
}

// sums[b] += sum of L over image b (exact integer; the contrast pivot is its rounded mean)
__global__ void lum_sum_kernel(const unsigned char* __restrict__ img, unsigned long long* __restrict__ sums,
                               int npix) {
  const int b = blockIdx.y;
  const unsigned char* p = img + (long)b * npix * 3;
  unsigned long long acc = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += gridDim.x * blockDim.x)
    acc += lum(p[3 * i], p[3 * i + 1], p[3 * i + 2]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0 && acc) atomicAdd(&sums[b], acc);
}

// PIL ImagingBlend on one channel: (UINT8)(in1 + alpha * (in2 - in1)) in float, clipped only outside [0, 1]
__device__ __forceinline__ unsigned char blend1(int d, int v, float a, bool inside) {
  const float t = __fadd_rn((float)d, __fmul_rn(a, (float)(v - d)));
  if (inside) return (unsigned char)(int)t;
  return t <= 0.f ? 0 : (t >= 255.f ? 255 : (unsigned char)(int)t);
}

__device__ __forceinline__ void rgb2hsv(int r, int g, int b, int& uh, int& us, int& uv) {
  const int maxc = max(r, max(g, b)), minc = min(r, min(g, b));
  uv = maxc;
  if (minc == maxc) { uh = 0; us = 0; return; }
  const float cr = (float)(maxc - minc);
  const float s = __fdiv_rn(cr, (float)maxc);
  const float rc = __fdiv_rn((float)(maxc - r), cr), gc = __fdiv_rn((float)(maxc - g), cr),
              bc = __fdiv_rn((float)(maxc - b), cr);
  float h;
  if (r == maxc) h = __fsub_rn(bc, gc);
  else if (g == maxc) h = (float)__dsub_rn(__dadd_rn(2.0, (double)rc), (double)bc);
  else h = (float)__dsub_rn(__dadd_rn(4.0, (double)gc), (double)rc);
  const float hd = (float)fmod(__dadd_rn(__ddiv_rn((double)h, 6.0), 1.0), 1.0);
  const int ih = (int)__dmul_rn((double)hd, 255.0), is = (int)__dmul_rn((double)s, 255.0);
  uh = ih < 0 ? 0 : (ih > 255 ? 255 : ih);
  us = is < 0 ? 0 : (is > 255 ? 255 : is);
}

__device__ __forceinline__ void hsv2rgb(int h, int s, int v, int& r, int& g, int& b) {
  if (s == 0) { r = g = b = v; return; }
  const double hf = __ddiv_rn(__dmul_rn((double)(float)h, 6.0), 255.0);
  const int i = (int)floor(hf);
  const double f = (double)(float)__dsub_rn(hf, (double)(float)i);
  const double fs = (double)(float)__ddiv_rn((double)(float)s, 255.0);
  const double vf = (double)(float)v;
  int p = (int)rint(__dmul_rn(vf, __dsub_rn(1.0, fs)));
  int q = (int)rint(__dmul_rn(vf, __dsub_rn(1.0, __dmul_rn(fs, f))));
  int t = (int)rint(__dmul_rn(vf, __dsub_rn(1.0, __dmul_rn(fs, __dsub_rn(1.0, f)))));
  p = p < 0 ? 0 : (p > 255 ? 255 : p);
  q = q < 0 ? 0 : (q > 255 ? 255 : q);
  t = t < 0 ? 0 : (t > 255 ? 255 : t);
  switch (i % 6) {
    case 0: r = v; g = t; b = p; break;
    case 1: r = q; g = v; b = p; break;
    case 2: r = p; g = v; b = t; break;
    case 3: r = p; g = q; b = v; break;
    case 4: r = t; g = p; b = v; break;
    default: r = v; g = p; b = q; break;
  }
}

// One ColorJitter position for a batch: image b applies op = order[4 b + pos] with factors[4 b + op]
// (0 brightness, 1 contrast, 2 saturation, 3 hue; factor NaN = the op is skipped).  In place.
__global__ void jitter_kernel(unsigned char* __restrict__ img, const int* __restrict__ order,
                              const float* __restrict__ factors, const unsigned long long* __restrict__ sums,
                              int pos, int npix) {
  const int b = blockIdx.y;
  const int op = order[4 * b + pos];
  const float f = factors[4 * b + op];
  if (f != f) return;
  unsigned char* p = img + (long)b * npix * 3;
  if (op == 3) {
    const int shift = (int)f & 255;        // the host passes np.int32(hue * 255).astype(np.uint8), computed in double
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += gridDim.x * blockDim.x) {
      int h, s, v, r, g, bb;
      rgb2hsv(p[3 * i], p[3 * i + 1], p[3 * i + 2], h, s, v);
      h = (h + shift) & 255;
      hsv2rgb(h, s, v, r, g, bb);
      p[3 * i] = (unsigned char)r; p[3 * i + 1] = (unsigned char)g; p[3 * i + 2] = (unsigned char)bb;
    }
    return;
  }
  if (f == 1.0f) return;                                        // Image.blend returns a copy of the image
  const bool inside = f >= 0.f && f <= 1.f;
  const int mean = (int)((double)sums[b] / (double)npix + 0.5);            // int(ImageStat.mean + 0.5)
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += gridDim.x * blockDim.x) {
    const int r = p[3 * i], g = p[3 * i + 1], bl = p[3 * i + 2];
    const int d = op == 0 ? 0 : (op == 1 ? mean : (int)lum(r, g, bl));
    if (f == 0.f) {                                             // ... and of the degenerate image
      p[3 * i] = p[3 * i + 1] = p[3 * i + 2] = (unsigned char)d;
    } else {
      p[3 * i] = blend1(d, r, f, inside);
      p[3 * i + 1] = blend1(d, g, f, inside);
      p[3 * i + 2] = blend1(d, bl, f, inside);
    }
  }
}

// [B][H][W][3] uint8 -> [B][3][H][W] float32 = u8 / 255 (transforms.ToTensor)
__global__ void to_tensor_kernel(const unsigned char* __restrict__ img, float* __restrict__ out, int npix,
                                 long total) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long b = i / npix;
    const int p = (int)(i - b * npix);
    const unsigned char* s = img + i * 3;
    float* o = out + b * 3 * (long)npix + p;
    o[0] = __fdiv_rn((float)s[0], 255.f);
    o[npix] = __fdiv_rn((float)s[1], 255.f);
    o[2 * (long)npix] = __fdiv_rn((float)s[2], 255.f);
  }
}

double lanczos(double x) {
  if (-3.0 <= x && x < 3.0) {
    if (x == 0.0) return 1.0;
    const double a = x * M_PI, b = (x / 3.0) * M_PI;
    return (sin(a) / a) * (sin(b) / b);
  }
  return 0.0;
}

int grid_for(long n) { return (int)(n < 256L * 148 * 8 ? (n + 255) / 256 : 148 * 8); }

}  // namespace

extern "C" {

int fd_lanczos_ksize(int in_size, int out_size) {
  if (in_size <= 0 || out_size <= 0) return 0;
  double fs = (double)in_size / out_size;
  if (fs < 1.0) fs = 1.0;
  return (int)ceil(3.0 * fs) * 2 + 1;
}

int fd_lanczos_coeffs(int in_size, int out_size, int* bounds, int* kk) {
  FD_REQUIRE(in_size > 0 && out_size > 0 && bounds && kk, "fd_lanczos_coeffs: bad arguments");
  // Pillow Resample.c precompute_coeffs (whole-image box) + normalize_coeffs_8bpc, host side, double
  const double scale = (double)in_size / out_size;
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = 3.0 * filterscale;
  const int ksize = (int)ceil(support) * 2 + 1;
  const double ss = 1.0 / filterscale;
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = (xx + 0.5) * scale;
    int xmin = (int)(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    double w[4096];
    FD_REQUIRE(xmax <= 4096, "fd_lanczos_coeffs: window of %d samples", xmax);
    double ww = 0.0;
    for (int x = 0; x < xmax; ++x) {
      w[x] = lanczos((x + xmin - center + 0.5) * ss);
      ww += w[x];
    }
    int* k = kk + (long)xx * ksize;
    for (int x = 0; x < ksize; ++x) {
      if (x < xmax) {
        const double v = ww != 0.0 ? w[x] / ww : w[x];
        k[x] = v < 0 ? (int)(-0.5 + v * (1 << PRECISION_BITS)) : (int)(0.5 + v * (1 << PRECISION_BITS));
      } else {
        k[x] = 0;
      }
    }
    bounds[2 * xx] = xmin;
    bounds[2 * xx + 1] = xmax;
  }
  return 0;
}

int fd_resize_lanczos_u8(const unsigned char* src, unsigned char* tmp, unsigned char* dst, int B, int Hin, int Win,
                         int Hout, int Wout, const int* bounds_w, const int* kk_w, int ksize_w,
                         const int* bounds_h, const int* kk_h, int ksize_h, const unsigned char* flip,
                         void* stream) {
  FD_REQUIRE(B > 0 && Hin > 0 && Win > 0 && Hout > 0 && Wout > 0, "fd_resize_lanczos_u8: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  // horizontal pass over all input rows (Pillow restricts it to the rows the vertical pass reads: the whole
  // image for a whole-image box), then vertical; tmp holds [B][Hin][Wout][3]
  const unsigned char* mid = src;
  if (Wout != Win || flip) {
    const long total = (long)B * Hin * Wout;
    unsigned char* o = Hout != Hin ? tmp : dst;
    resample_h_kernel<<<grid_for(total), 256, 0, st>>>(src, o, bounds_w, kk_w, ksize_w, Hin, Win, Wout, flip, total);
    FD_CHECK_LAUNCH();
    mid = o;
  }
  if (Hout != Hin) {
    const long total = (long)B * Hout * Wout * 3;
    resample_v_kernel<<<grid_for(total), 256, 0, st>>>(mid, dst, bounds_h, kk_h, ksize_h, Hin, Hout, Wout, total);
    FD_CHECK_LAUNCH();
  } else if (mid == src) {
    cudaMemcpyAsync(dst, src, (size_t)B * Hin * Win * 3, cudaMemcpyDeviceToDevice, st);
  }
  return 0;
}

int fd_color_jitter_u8(unsigned char* img, int B, int H, int W, const int* order, const float* factors,
                       unsigned long long* lum_sums, void* stream) {
  FD_REQUIRE(B > 0 && H > 0 && W > 0 && order && factors && lum_sums, "fd_color_jitter_u8: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const int npix = H * W;
  dim3 grid(min((npix + 255) / 256, 148 * 4), B);
  for (int pos = 0; pos < 4; ++pos) {
    cudaMemsetAsync(lum_sums, 0, sizeof(unsigned long long) * B, st);
    lum_sum_kernel<<<grid, 256, 0, st>>>(img, lum_sums, npix);
    FD_CHECK_LAUNCH();
    jitter_kernel<<<grid, 256, 0, st>>>(img, order, factors, lum_sums, pos, npix);
    FD_CHECK_LAUNCH();
  }
  return 0;
}

int fd_image_to_tensor(const unsigned char* img, float* out, int B, int H, int W, void* stream) {
  FD_REQUIRE(B > 0 && H > 0 && W > 0, "fd_image_to_tensor: bad shape");
  const long total = (long)B * H * W;
  to_tensor_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(img, out, H * W, total);
  FD_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
