// Single-output-channel 3x3 convolutions (the four disparity heads, reference
// networks/depth_decoder.py:52-58 "dispconv" = Conv3x3(num_ch_dec[s], 1) + sigmoid, layers.py:115-130):
// forward, data gradient and weight gradient as memory-bound streaming kernels.
//
// An implicit-GEMM tile (128 pixels x 64 channels on the CUDA cores, N >= 16 on the tensor cores)
// wastes 63/64 of its work on N = 1; these kernels read every input element once (forward / weight
// gradient) or write every gradient element once (data gradient):
//   forward   thread = output pixel, weights in shared memory, float4 channel loads
//   dgrad     thread = (input pixel, channel quad), 9 broadcast dY loads, one float4 store
//   wgrad     thread = (channel quad, pixel lane): each input element is loaded once and feeds the
//             nine taps it belongs to; block reduction in shared memory, 9*C atomics per block
// NHWC activations, weights [1][3][3][C], stride 1, zero padding `pad` (0 when the caller has already
// reflection-padded the input through fd_assemble_fwd).
#include "common.cuh"
#include "../../include/fusiondepth_b200.h"

namespace {

constexpr int TAPS = 9;

__device__ __forceinline__ float act1(float v, int act) {
  switch (act) {
    case FD_ACT_RELU: return fmaxf(v, 0.f);
    case FD_ACT_ELU: return v > 0.f ? v : expm1f(v);
    case FD_ACT_SIGMOID: return 1.f / (1.f + expf(-v));
    case FD_ACT_TANH: return tanhf(v);
    default: return v;
  }
}

__global__ void __launch_bounds__(256) cout1_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                        const float* __restrict__ bias, float* __restrict__ y,
                                                        int B, int H, int W, int C, int Ho, int Wo, int pad,
                                                        int act) {
  extern __shared__ float ws[];                       // [9][C]
  for (int i = threadIdx.x; i < TAPS * C; i += blockDim.x) ws[i] = w[i];
  __syncthreads();
  const long M = (long)B * Ho * Wo;
  const float b0 = bias ? bias[0] : 0.f;
  for (long m = (long)blockIdx.x * blockDim.x + threadIdx.x; m < M; m += (long)gridDim.x * blockDim.x) {
    const int wo = (int)(m % Wo);
    const long r = m / Wo;
    const int ho = (int)(r % Ho), b = (int)(r / Ho);
    // four independent partial sums keep the FMA chains short; fixed order => deterministic
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int h = ho - pad + kh;
      if (h < 0 || h >= H) continue;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int wi = wo - pad + kw;
        if (wi < 0 || wi >= W) continue;
        const float4* xp = reinterpret_cast<const float4*>(x + (((long)b * H + h) * W + wi) * C);
        const float4* wp = reinterpret_cast<const float4*>(ws + (kh * 3 + kw) * C);
        for (int c = 0; c < C / 4; ++c) {
          const float4 xv = xp[c], wv = wp[c];
          acc[0] = fmaf(xv.x, wv.x, acc[0]);
          acc[1] = fmaf(xv.y, wv.y, acc[1]);
          acc[2] = fmaf(xv.z, wv.z, acc[2]);
          acc[3] = fmaf(xv.w, wv.w, acc[3]);
        }
      }
    }
    y[m] = act1((acc[0] + acc[1]) + (acc[2] + acc[3]) + b0, act);
  }
}

// dx[b,h,w,c] = sum_{kh,kw} dy[b, h+pad-kh, w+pad-kw] * w[kh][kw][c]
__global__ void __launch_bounds__(256) cout1_dgrad_kernel(const float* __restrict__ dy, const float* __restrict__ w,
                                                          float* __restrict__ dx, int B, int H, int W, int C,
                                                          int Ho, int Wo, int pad) {
  extern __shared__ float ws[];
  for (int i = threadIdx.x; i < TAPS * C; i += blockDim.x) ws[i] = w[i];
  __syncthreads();
  const int C4 = C / 4;
  const long total = (long)B * H * W * C4;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c4 = (int)(i % C4);
    const long p = i / C4;
    const int wi = (int)(p % W);
    const long r = p / W;
    const int h = (int)(r % H), b = (int)(r / H);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int ho = h + pad - kh;
      if (ho < 0 || ho >= Ho) continue;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int wo = wi + pad - kw;
        if (wo < 0 || wo >= Wo) continue;
        const float g = dy[((long)b * Ho + ho) * Wo + wo];
        const float4 wv = reinterpret_cast<const float4*>(ws + (kh * 3 + kw) * C)[c4];
        acc.x = fmaf(g, wv.x, acc.x); acc.y = fmaf(g, wv.y, acc.y);
        acc.z = fmaf(g, wv.z, acc.z); acc.w = fmaf(g, wv.w, acc.w);
      }
    }
    reinterpret_cast<float4*>(dx)[i] = acc;
  }
}

// dw[kh][kw][c] += sum_{b,ho,wo} dy[b,ho,wo] * x[b, ho-pad+kh, wo-pad+kw, c]
// The loop runs over INPUT pixels q = (h, w): x[q] is loaded once and contributes to tap (kh,kw)
// through the output pixel (h+pad-kh, w+pad-kw).
__global__ void __launch_bounds__(256) cout1_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                          float* __restrict__ dw, int B, int H, int W, int C,
                                                          int Ho, int Wo, int pad) {
  extern __shared__ float red[];                      // [lanes][9][C] partials, folded in place
  const int C4 = C / 4;
  const int lanes = blockDim.x / C4;                  // pixel lanes per block (host: C4 divides 256)
  const int c4 = threadIdx.x % C4, pl = threadIdx.x / C4;
  float4 acc[TAPS];
#pragma unroll
  for (int t = 0; t < TAPS; ++t) acc[t] = make_float4(0.f, 0.f, 0.f, 0.f);
  const long P = (long)B * H * W;
  for (long p = (long)blockIdx.x * lanes + pl; p < P; p += (long)gridDim.x * lanes) {
    const int wi = (int)(p % W);
    const long r = p / W;
    const int h = (int)(r % H), b = (int)(r / H);
    const float4 xv = reinterpret_cast<const float4*>(x + p * C)[c4];
    const float* dyb = dy + (long)b * Ho * Wo;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int ho = h + pad - kh;
      if (ho < 0 || ho >= Ho) continue;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int wo = wi + pad - kw;
        if (wo < 0 || wo >= Wo) continue;
        const float g = dyb[(long)ho * Wo + wo];
        float4& a = acc[kh * 3 + kw];
        a.x = fmaf(g, xv.x, a.x); a.y = fmaf(g, xv.y, a.y);
        a.z = fmaf(g, xv.z, a.z); a.w = fmaf(g, xv.w, a.w);
      }
    }
  }
  // fold the pixel lanes: red[pl][t][c]
#pragma unroll
  for (int t = 0; t < TAPS; ++t)
    reinterpret_cast<float4*>(red + ((long)pl * TAPS + t) * C)[c4] = acc[t];
  __syncthreads();
  for (int i = threadIdx.x; i < TAPS * C; i += blockDim.x) {
    float s = 0.f;
    for (int l = 0; l < lanes; ++l) s += red[(long)l * TAPS * C + i];
    atomicAdd(dw + i, s);
  }
}

}  // namespace

extern "C" {

int fd_conv2d_cout1_supported(int Cin, int Cout, int KH, int KW, int stride) {
  // C/4 must divide the 256-thread block of the weight-gradient kernel
  return Cout == 1 && KH == 3 && KW == 3 && stride == 1 && Cin % 4 == 0 && Cin <= 1024 && 256 % (Cin / 4) == 0;
}

int fd_conv2d_cout1_fwd(const float* x, const float* w, const float* bias, float* y, int B, int H, int W,
                        int Cin, int pad, int act, void* stream) {
  FD_REQUIRE(fd_conv2d_cout1_supported(Cin, 1, 3, 3, 1), "fd_conv2d_cout1_fwd: unsupported Cin %d", Cin);
  const int Ho = H + 2 * pad - 2, Wo = W + 2 * pad - 2;
  FD_REQUIRE(Ho > 0 && Wo > 0, "fd_conv2d_cout1_fwd: empty output (%d x %d)", Ho, Wo);
  const long M = (long)B * Ho * Wo;
  cout1_fwd_kernel<<<min(fd::cdiv(M, 256), 148 * 16), 256, TAPS * Cin * sizeof(float), (cudaStream_t)stream>>>(
      x, w, bias, y, B, H, W, Cin, Ho, Wo, pad, act);
  FD_CHECK_LAUNCH();
  return 0;
}

int fd_conv2d_cout1_dgrad(const float* dy, const float* w, float* dx, int B, int H, int W, int Cin, int pad,
                          void* stream) {
  FD_REQUIRE(fd_conv2d_cout1_supported(Cin, 1, 3, 3, 1), "fd_conv2d_cout1_dgrad: unsupported Cin %d", Cin);
  const int Ho = H + 2 * pad - 2, Wo = W + 2 * pad - 2;
  const long total = (long)B * H * W * (Cin / 4);
  cout1_dgrad_kernel<<<min(fd::cdiv(total, 256), 148 * 16), 256, TAPS * Cin * sizeof(float),
                       (cudaStream_t)stream>>>(dy, w, dx, B, H, W, Cin, Ho, Wo, pad);
  FD_CHECK_LAUNCH();
  return 0;
}

int fd_conv2d_cout1_wgrad(const float* x, const float* dy, float* dw, int B, int H, int W, int Cin, int pad,
                          void* stream) {
  FD_REQUIRE(fd_conv2d_cout1_supported(Cin, 1, 3, 3, 1), "fd_conv2d_cout1_wgrad: unsupported Cin %d", Cin);
  const int Ho = H + 2 * pad - 2, Wo = W + 2 * pad - 2;
  const int lanes = 256 / (Cin / 4);
  const size_t smem = (size_t)lanes * TAPS * Cin * sizeof(float);        // = 256 * 9 * 16 B = 36 KB
  const long P = (long)B * H * W;
  int blocks = min(fd::cdiv(P, (long)lanes * 8), 148 * 4);
  if (blocks < 1) blocks = 1;
  cout1_wgrad_kernel<<<blocks, 256, smem, (cudaStream_t)stream>>>(x, dy, dw, B, H, W, Cin, Ho, Wo, pad);
  FD_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
