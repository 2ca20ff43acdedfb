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

// ================================================================================================
// 16-channel 3x3 layers at full / half resolution (DepthDecoder upconv(0,0) 32->16 and upconv(0,1)
// 16->16, reference networks/depth_decoder.py:20-50 with num_ch_dec[0] = 16): direct convolutions on
// the CUDA cores.  N = 16 output channels is one eighth of a tensor-core tile row and the generic
// 128 x 64 implicit-GEMM tile wastes three quarters of its FMAs, while these layers hold the largest
// pixel counts of the network (737k rows).
//
// direct3x3: out[b,h,w,:] = act(bias + sum_{kh,kw,ci} in[b, h+kh-off, w+kw-off, ci] * wk[kh][kw][ci][:])
//   with zeros outside the input.  Forward: wk = W^T, off = pad.  Data gradient: in = dY, out = dX,
//   wk[kh][kw][n][c] = W[n][2-kh][2-kw][c], off = 2 - pad.  Thread = PX consecutive pixels x all CO
//   outputs; the weights sit in shared memory as [tap][ci][co] and every float4 of them (a warp-wide
//   broadcast load) feeds 4*PX FMAs.
// wgrad16: dW[n][kh][kw][c] += sum_p dY[p][n] * X[p + (kh,kw) - pad][c]; thread = (n, channel quad, kh)
//   with a sliding 3-wide register window along the row; one block per slab of rows, atomics out.
// ================================================================================================
namespace {

template <int CI, int CO, int PX, bool DGRAD>
__global__ void __launch_bounds__(128) direct3x3_kernel(const float* __restrict__ in, const float* __restrict__ w,
                                                        const float* __restrict__ bias, float* __restrict__ out,
                                                        int B, int Hi, int Wi, int Ho, int Wo, int off, int act) {
  __shared__ __align__(16) float wk[9 * CI * CO];
  for (int i = threadIdx.x; i < 9 * CI * CO; i += blockDim.x) {
    const int co = i % CO, ci = (i / CO) % CI, tap = i / (CO * CI);
    // forward: W is [CO][9][CI];  data gradient: W is [n = CI][9][c = CO], taps flipped
    wk[i] = DGRAD ? w[((long)ci * 9 + (8 - tap)) * CO + co] : w[((long)co * 9 + tap) * CI + ci];
  }
  __syncthreads();
  const int wt_per_row = (Wo + PX - 1) / PX;
  const long tiles = (long)B * Ho * wt_per_row;
  for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < tiles; t += (long)gridDim.x * blockDim.x) {
    const int wt = (int)(t % wt_per_row);
    const long r = t / wt_per_row;
    const int h = (int)(r % Ho), b = (int)(r / Ho);
    const int w0 = wt * PX;
    float acc[PX][CO];
#pragma unroll
    for (int p = 0; p < PX; ++p)
#pragma unroll
      for (int n = 0; n < CO; ++n) acc[p][n] = 0.f;
#pragma unroll 1
    for (int kh = 0; kh < 3; ++kh) {
      const int hi = h + kh - off;
      if (hi < 0 || hi >= Hi) continue;
      const float* rowp = in + ((long)b * Hi + hi) * Wi * CI;
#pragma unroll 1
      for (int cq = 0; cq < CI / 4; ++cq) {
        float4 xin[PX + 2];
#pragma unroll
        for (int j = 0; j < PX + 2; ++j) {
          const int wi = w0 + j - off;
          xin[j] = (wi >= 0 && wi < Wi) ? *reinterpret_cast<const float4*>(rowp + (long)wi * CI + cq * 4)
                                        : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
#pragma unroll
          for (int ci = 0; ci < 4; ++ci) {
            const float4* wp = reinterpret_cast<const float4*>(wk + ((kh * 3 + kw) * CI + cq * 4 + ci) * CO);
#pragma unroll
            for (int nq = 0; nq < CO / 4; ++nq) {
              const float4 wv = wp[nq];
#pragma unroll
              for (int p = 0; p < PX; ++p) {
                const float4 xv4 = xin[p + kw];
                const float xs = ci == 0 ? xv4.x : (ci == 1 ? xv4.y : (ci == 2 ? xv4.z : xv4.w));
                acc[p][4 * nq + 0] = fmaf(xs, wv.x, acc[p][4 * nq + 0]);
                acc[p][4 * nq + 1] = fmaf(xs, wv.y, acc[p][4 * nq + 1]);
                acc[p][4 * nq + 2] = fmaf(xs, wv.z, acc[p][4 * nq + 2]);
                acc[p][4 * nq + 3] = fmaf(xs, wv.w, acc[p][4 * nq + 3]);
              }
            }
          }
        }
      }
    }
#pragma unroll
    for (int p = 0; p < PX; ++p) {
      const int wo = w0 + p;
      if (wo >= Wo) continue;
      float* op = out + (((long)b * Ho + h) * Wo + wo) * CO;
#pragma unroll
      for (int nq = 0; nq < CO / 4; ++nq) {
        float4 v = make_float4(acc[p][4 * nq], acc[p][4 * nq + 1], acc[p][4 * nq + 2], acc[p][4 * nq + 3]);
        if (!DGRAD) {
          if (bias) { v.x += bias[4 * nq]; v.y += bias[4 * nq + 1]; v.z += bias[4 * nq + 2]; v.w += bias[4 * nq + 3]; }
          v.x = act1(v.x, act); v.y = act1(v.y, act); v.z = act1(v.z, act); v.w = act1(v.w, act);
        }
        reinterpret_cast<float4*>(op)[nq] = v;
      }
    }
  }
}

// thread = (n, channel quad cq, kh); CO = 16 output channels.  Block = 16 * (CI/4) * 3 threads.
template <int CI>
__global__ void __launch_bounds__(16 * (CI / 4) * 3) wgrad16_kernel(const float* __restrict__ x,
                                                                  const float* __restrict__ dy,
                                                                  float* __restrict__ dw, int B, int H, int W,
                                                                  int Ho, int Wo, int pad, int rows_per_block) {
  constexpr int CO = 16;
  const int n = threadIdx.x % CO, cq = (threadIdx.x / CO) % (CI / 4), kh = threadIdx.x / (CO * (CI / 4));
  float4 acc[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  const long rows = (long)B * Ho;
  const long r0 = (long)blockIdx.x * rows_per_block;
  const long r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
  for (long r = r0; r < r1; ++r) {
    const int ho = (int)(r % Ho), b = (int)(r / Ho);
    const int hi = ho - pad + kh;
    if (hi < 0 || hi >= H) continue;                        // uniform per kh group of threads
    const float* xr = x + (((long)b * H + hi) * W) * CI + cq * 4;
    const float* gr = dy + (((long)b * Ho + ho) * Wo) * CO + n;
    // window: x at columns wo-pad, wo-pad+1, wo-pad+2
    float4 x0 = make_float4(0.f, 0.f, 0.f, 0.f), x1 = x0, x2 = x0;
    if (-pad >= 0 && -pad < W) x1 = *reinterpret_cast<const float4*>(xr + (long)(-pad) * CI);
    if (1 - pad >= 0 && 1 - pad < W) x2 = *reinterpret_cast<const float4*>(xr + (long)(1 - pad) * CI);
    auto step = [&](const float4& xn, float g) {
      x0 = x1; x1 = x2; x2 = xn;
      acc[0].x = fmaf(g, x0.x, acc[0].x); acc[0].y = fmaf(g, x0.y, acc[0].y);
      acc[0].z = fmaf(g, x0.z, acc[0].z); acc[0].w = fmaf(g, x0.w, acc[0].w);
      acc[1].x = fmaf(g, x1.x, acc[1].x); acc[1].y = fmaf(g, x1.y, acc[1].y);
      acc[1].z = fmaf(g, x1.z, acc[1].z); acc[1].w = fmaf(g, x1.w, acc[1].w);
      acc[2].x = fmaf(g, x2.x, acc[2].x); acc[2].y = fmaf(g, x2.y, acc[2].y);
      acc[2].z = fmaf(g, x2.z, acc[2].z); acc[2].w = fmaf(g, x2.w, acc[2].w);
    };
    auto load_x = [&](int wi) {
      return (wi >= 0 && wi < W) ? *reinterpret_cast<const float4*>(xr + (long)wi * CI)
                                 : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    int wo = 0;
    for (; wo + 4 <= Wo; wo += 4) {            // four independent load pairs in flight per thread
      float4 xn[4];
      float g[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        xn[u] = load_x(wo + u + 2 - pad);
        g[u] = gr[(long)(wo + u) * CO];
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) step(xn[u], g[u]);
    }
    for (; wo < Wo; ++wo) step(load_x(wo + 2 - pad), gr[(long)wo * CO]);
  }
  // dW layout [n][kh][kw][CI]
#pragma unroll
  for (int kw = 0; kw < 3; ++kw) {
    float* d = dw + (((long)n * 3 + kh) * 3 + kw) * CI + cq * 4;
    atomicAdd(d + 0, acc[kw].x); atomicAdd(d + 1, acc[kw].y);
    atomicAdd(d + 2, acc[kw].z); atomicAdd(d + 3, acc[kw].w);
  }
}

}  // namespace

extern "C" {

int fd_conv2d_c16_supported(int Cin, int Cout, int KH, int KW, int stride) {
  return KH == 3 && KW == 3 && stride == 1 && Cout == 16 && (Cin == 16 || Cin == 32);
}

int fd_conv2d_c16_fwd(const float* x, const float* w, const float* bias, float* y, int B, int H, int W, int Cin,
                      int pad, int act, void* stream) {
  FD_REQUIRE(Cin == 16, "fd_conv2d_c16_fwd: Cin must be 16 (Cin = 32 runs on the tensor cores), got %d", Cin);
  const int Ho = H + 2 * pad - 2, Wo = W + 2 * pad - 2;
  FD_REQUIRE(Ho > 0 && Wo > 0, "fd_conv2d_c16_fwd: empty output");
  const long tiles = (long)B * Ho * ((Wo + 3) / 4);
  direct3x3_kernel<16, 16, 4, false><<<min(fd::cdiv(tiles, 128), 148 * 16), 128, 0, (cudaStream_t)stream>>>(
      x, w, bias, y, B, H, W, Ho, Wo, pad, act);
  FD_CHECK_LAUNCH();
  return 0;
}

int fd_conv2d_c16_dgrad(const float* dy, const float* w, float* dx, int B, int H, int W, int Cin, int pad,
                        void* stream) {
  FD_REQUIRE(Cin == 16 || Cin == 32, "fd_conv2d_c16_dgrad: Cin must be 16 or 32, got %d", Cin);
  const int Ho = H + 2 * pad - 2, Wo = W + 2 * pad - 2;
  cudaStream_t st = (cudaStream_t)stream;
  if (Cin == 16) {
    const long tiles = (long)B * H * ((W + 3) / 4);
    direct3x3_kernel<16, 16, 4, true><<<min(fd::cdiv(tiles, 128), 148 * 16), 128, 0, st>>>(
        dy, w, nullptr, dx, B, Ho, Wo, H, W, 2 - pad, 0);
  } else {
    const long tiles = (long)B * H * ((W + 1) / 2);
    direct3x3_kernel<16, 32, 2, true><<<min(fd::cdiv(tiles, 128), 148 * 16), 128, 0, st>>>(
        dy, w, nullptr, dx, B, Ho, Wo, H, W, 2 - pad, 0);
  }
  FD_CHECK_LAUNCH();
  return 0;
}

int fd_conv2d_c16_wgrad(const float* x, const float* dy, float* dw, int B, int H, int W, int Cin, int pad,
                        void* stream) {
  FD_REQUIRE(Cin == 16 || Cin == 32, "fd_conv2d_c16_wgrad: Cin must be 16 or 32, got %d", Cin);
  const int Ho = H + 2 * pad - 2, Wo = W + 2 * pad - 2;
  const long rows = (long)B * Ho;
  int blocks = (int)(rows < 148 * 4 ? rows : 148 * 4);     // ~4 resident blocks per SM hide the load latency
  const int rpb = (int)((rows + blocks - 1) / blocks);
  blocks = (int)((rows + rpb - 1) / rpb);
  cudaStream_t st = (cudaStream_t)stream;
  if (Cin == 16) wgrad16_kernel<16><<<blocks, 16 * 4 * 3, 0, st>>>(x, dy, dw, B, H, W, Ho, Wo, pad, rpb);
  else wgrad16_kernel<32><<<blocks, 16 * 8 * 3, 0, st>>>(x, dy, dw, B, H, W, Ho, Wo, pad, rpb);
  FD_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
