// fp32 implicit-GEMM convolution on the CUDA cores: forward, data gradient and weight gradient.
//
// This is the exact-fp32 path: it serves every shape the tensor-core path (conv_tc.cu) does not
// take (tiny channel counts such as the 7x7 stem with Cin in {2,3,4,6}, Cout = 1 disparity heads,
// odd refine-decoder channel counts) and is the on-device cross-check for it.
//
// Replaces aten::convolution / convolution_backward under networks/* (SURVEY.md section 2.1).
// GEMM view (NHWC, weights [Cout,KH,KW,Cin]):  M = B*Ho*Wo pixels, N = Cout, K = KH*KW*Cin with
// k = (kh*KW + kw)*Cin + ci.  CTA tile 128(M) x 64(N) x 16(K), 256 threads, 8x4 register
// micro-tile, double-buffered shared memory with register prefetch.
#include "common.cuh"
#include "../../include/fusiondepth_b200.h"

namespace {

constexpr int BM = 128, BN = 64, BK = 16, NTH = 256;
constexpr int AS = BM + 4, BS = BN + 4;

struct ConvArgs {
  const float* x;     // gathered tensor [B,Hg,Wg,Cg]
  const float* w;     // [N][taps*Cg]
  const float* bias;
  float* y;           // [B,Ho,Wo,N]
  int B, Hg, Wg, Cg;  // gathered tensor dims
  int Ho, Wo, N;      // output dims
  int KH, KW, stride, pad, act;
  long M;
  int K;
};

__device__ __forceinline__ float apply_act(float v, int act) {
  switch (act) {
    case FD_ACT_RELU: return fmaxf(v, 0.f);
    case FD_ACT_ELU: return v > 0.f ? v : expm1f(v);
    case FD_ACT_SIGMOID: return 1.f / (1.f + expf(-v));
    case FD_ACT_TANH: return tanhf(v);
    default: return v;
  }
}

// source coordinate of tap (kh,kw) for output pixel (ho,wo).  MODE 0: forward conv;
// MODE 1: data gradient (gathers dy at (h + pad - kh)/stride when divisible).
template <int MODE>
__device__ __forceinline__ bool tap_coord(const ConvArgs& a, int ho, int wo, int kh, int kw, int& h,
                                          int& w) {
  if (MODE == 0) {
    h = ho * a.stride - a.pad + kh;
    w = wo * a.stride - a.pad + kw;
  } else {
    int th = ho + a.pad - kh, tw = wo + a.pad - kw;
    if (th < 0 || tw < 0) return false;
    if (a.stride > 1) {
      if ((th % a.stride) | (tw % a.stride)) return false;
      th /= a.stride; tw /= a.stride;
    }
    h = th; w = tw;
  }
  return h >= 0 && h < a.Hg && w >= 0 && w < a.Wg;
}

template <int MODE, bool VEC>
__global__ void __launch_bounds__(NTH) conv_igemm_kernel(ConvArgs a) {
  __shared__ __align__(16) float As[2][BK][AS];
  __shared__ __align__(16) float Bs[2][BK][BS];
  const int tid = threadIdx.x;
  const long m0 = (long)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int ty = tid >> 4, tx = tid & 15;
  const int HoWo = a.Ho * a.Wo;

  // ---- loader state -------------------------------------------------------------------------
  // VEC: thread owns rows (tid>>2) and (tid>>2)+64, k-quad (tid&3) of the A tile; row tid>>2,
  // k-quad tid&3 of the B tile.   scalar: element e = tid + 256 j -> (row e>>4, k e&15).
  int rb[2], rho[2], rwo[2];
  bool rok[2];
  if (VEC) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      long m = m0 + (tid >> 2) + 64 * j;
      rok[j] = m < a.M;
      long mm = rok[j] ? m : 0;
      rb[j] = (int)(mm / HoWo);
      int r = (int)(mm % HoWo);
      rho[j] = r / a.Wo; rwo[j] = r % a.Wo;
    }
  }
  float4 ra[2], rbq;
  float sa[8], sb[4];

  auto load_tile = [&](int kt) {
    const int k0 = kt * BK;
    if (VEC) {
      const int k = k0 + (tid & 3) * 4;
      const int tap = k / a.Cg, c = k - tap * a.Cg;
      const int kh = tap / a.KW, kw = tap - kh * a.KW;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        int h, w;
        ra[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (rok[j] && k < a.K && tap_coord<MODE>(a, rho[j], rwo[j], kh, kw, h, w))
          ra[j] = *reinterpret_cast<const float4*>(a.x + (((long)rb[j] * a.Hg + h) * a.Wg + w) * a.Cg + c);
      }
      const int n = n0 + (tid >> 2);
      rbq = make_float4(0.f, 0.f, 0.f, 0.f);
      if (n < a.N && k < a.K) rbq = *reinterpret_cast<const float4*>(a.w + (long)n * a.K + k);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        int e = tid + NTH * j;
        int row = e >> 4, k = k0 + (e & 15);
        long m = m0 + row;
        float v = 0.f;
        if (m < a.M && k < a.K) {
          int tap = k / a.Cg, c = k - tap * a.Cg;
          int kh = tap / a.KW, kw = tap - kh * a.KW;
          int b = (int)(m / HoWo), r = (int)(m % HoWo);
          int h, w;
          if (tap_coord<MODE>(a, r / a.Wo, r % a.Wo, kh, kw, h, w))
            v = a.x[(((long)b * a.Hg + h) * a.Wg + w) * a.Cg + c];
        }
        sa[j] = v;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int e = tid + NTH * j;
        int n = n0 + (e >> 4), k = k0 + (e & 15);
        sb[j] = (n < a.N && k < a.K) ? a.w[(long)n * a.K + k] : 0.f;
      }
    }
  };
  auto store_tile = [&](int buf) {
    if (VEC) {
      const int kq = (tid & 3) * 4;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        int row = (tid >> 2) + 64 * j;
        As[buf][kq + 0][row] = ra[j].x; As[buf][kq + 1][row] = ra[j].y;
        As[buf][kq + 2][row] = ra[j].z; As[buf][kq + 3][row] = ra[j].w;
      }
      int n = tid >> 2;
      Bs[buf][kq + 0][n] = rbq.x; Bs[buf][kq + 1][n] = rbq.y;
      Bs[buf][kq + 2][n] = rbq.z; Bs[buf][kq + 3][n] = rbq.w;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) { int e = tid + NTH * j; As[buf][e & 15][e >> 4] = sa[j]; }
#pragma unroll
      for (int j = 0; j < 4; ++j) { int e = tid + NTH * j; Bs[buf][e & 15][e >> 4] = sb[j]; }
    }
  };

  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int nk = (a.K + BK - 1) / BK;
  load_tile(0);
  store_tile(0);
  __syncthreads();
  int cur = 0;
  for (int kt = 0; kt < nk; ++kt) {
    if (kt + 1 < nk) load_tile(kt + 1);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[cur][kk][ty * 8]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[cur][kk][ty * 8 + 4]);
      float4 b = *reinterpret_cast<const float4*>(&Bs[cur][kk][tx * 4]);
      float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (kt + 1 < nk) store_tile(cur ^ 1);
    __syncthreads();
    cur ^= 1;
  }

  // ---- epilogue: bias + activation, NHWC store -------------------------------------------------
  const int n = n0 + tx * 4;
  float bv[4] = {0.f, 0.f, 0.f, 0.f};
  if (a.bias) {
#pragma unroll
    for (int j = 0; j < 4; ++j) if (n + j < a.N) bv[j] = a.bias[n + j];
  }
  const bool vec_store = (a.N % 4 == 0) && (n + 3 < a.N);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    long m = m0 + ty * 8 + i;
    if (m >= a.M) continue;
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j] = apply_act(acc[i][j] + bv[j], a.act);
    float* dst = a.y + m * a.N + n;
    if (vec_store) {
      *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) if (n + j < a.N) dst[j] = o[j];
    }
  }
}

// ---- weight gradient: dW[n][k] += sum_m dY[m][n] * A[m][k] ----------------------------------------
// CTA tile 128 (k) x 64 (n), reduction over a slice of the pixels (blockIdx.z), fp32 atomics out.
struct WgArgs {
  const float* x;   // [B,H,W,Cin]
  const float* dy;  // [B,Ho,Wo,N]
  float* dw;        // [N][K]
  int B, H, W, Cin, Ho, Wo, N, KH, KW, stride, pad;
  long M;
  int K;
  long m_per_split;
};

template <bool VEC>
__global__ void __launch_bounds__(NTH) conv_wgrad_kernel(WgArgs a) {
  __shared__ __align__(16) float As[2][BK][AS];   // [pixel][k]
  __shared__ __align__(16) float Bs[2][BK][BS];   // [pixel][n]
  const int tid = threadIdx.x;
  const int k0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int ty = tid >> 4, tx = tid & 15;
  const int HoWo = a.Ho * a.Wo;
  const long mbeg = (long)blockIdx.z * a.m_per_split;
  const long mend = min(a.M, mbeg + a.m_per_split);
  if (mbeg >= mend) return;

  // VEC: this thread's k quad is fixed: decode its tap once
  const int kq = k0 + (tid & 31) * 4;
  int vkh = 0, vkw = 0, vc = 0;
  if (VEC && kq < a.K) {
    int tap = kq / a.Cin;
    vc = kq - tap * a.Cin;
    vkh = tap / a.KW; vkw = tap - vkh * a.KW;
  }
  float4 ra[2], rbq;
  float sa[8], sb[4];

  auto load_tile = [&](long mt) {
    if (VEC) {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        long m = mt + (tid >> 5) + 8 * j;
        ra[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (m < mend && kq < a.K) {
          int b = (int)(m / HoWo), r = (int)(m % HoWo);
          int h = (r / a.Wo) * a.stride - a.pad + vkh, w = (r % a.Wo) * a.stride - a.pad + vkw;
          if (h >= 0 && h < a.H && w >= 0 && w < a.W)
            ra[j] = *reinterpret_cast<const float4*>(a.x + (((long)b * a.H + h) * a.W + w) * a.Cin + vc);
        }
      }
      long m = mt + (tid >> 4);
      int n = n0 + (tid & 15) * 4;
      rbq = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m < mend && n < a.N) rbq = *reinterpret_cast<const float4*>(a.dy + m * a.N + n);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        int e = tid + NTH * j;
        int kk = e & 127, mm = e >> 7;
        long m = mt + mm;
        int k = k0 + kk;
        float v = 0.f;
        if (m < mend && k < a.K) {
          int tap = k / a.Cin, c = k - tap * a.Cin;
          int kh = tap / a.KW, kw = tap - kh * a.KW;
          int b = (int)(m / HoWo), r = (int)(m % HoWo);
          int h = (r / a.Wo) * a.stride - a.pad + kh, w = (r % a.Wo) * a.stride - a.pad + kw;
          if (h >= 0 && h < a.H && w >= 0 && w < a.W) v = a.x[(((long)b * a.H + h) * a.W + w) * a.Cin + c];
        }
        sa[j] = v;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int e = tid + NTH * j;
        int nn = e & 63, mm = e >> 6;
        long m = mt + mm;
        sb[j] = (m < mend && n0 + nn < a.N) ? a.dy[m * a.N + n0 + nn] : 0.f;
      }
    }
  };
  auto store_tile = [&](int buf) {
    if (VEC) {
#pragma unroll
      for (int j = 0; j < 2; ++j)
        *reinterpret_cast<float4*>(&As[buf][(tid >> 5) + 8 * j][(tid & 31) * 4]) = ra[j];
      *reinterpret_cast<float4*>(&Bs[buf][tid >> 4][(tid & 15) * 4]) = rbq;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) { int e = tid + NTH * j; As[buf][e >> 7][e & 127] = sa[j]; }
#pragma unroll
      for (int j = 0; j < 4; ++j) { int e = tid + NTH * j; Bs[buf][e >> 6][e & 63] = sb[j]; }
    }
  };

  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  load_tile(mbeg);
  store_tile(0);
  __syncthreads();
  int cur = 0;
  for (long mt = mbeg; mt < mend; mt += BK) {
    if (mt + BK < mend) load_tile(mt + BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[cur][kk][ty * 8]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[cur][kk][ty * 8 + 4]);
      float4 b = *reinterpret_cast<const float4*>(&Bs[cur][kk][tx * 4]);
      float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (mt + BK < mend) store_tile(cur ^ 1);
    __syncthreads();
    cur ^= 1;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int k = k0 + ty * 8 + i;
    if (k >= a.K) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n < a.N) atomicAdd(a.dw + (long)n * a.K + k, acc[i][j]);
    }
  }
}

int check_dims(const char* who, int B, int H, int W, int Cin, int Cout, int KH, int KW, int stride,
               int pad) {
  FD_REQUIRE(B > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0 && KH > 0 && KW > 0 && stride > 0 &&
                 pad >= 0 && H + 2 * pad >= KH && W + 2 * pad >= KW,
             "%s: bad dims B=%d H=%d W=%d Cin=%d Cout=%d k=%dx%d s=%d p=%d", who, B, H, W, Cin, Cout,
             KH, KW, stride, pad);
  return 0;
}

}  // namespace

extern "C" {

int fd_conv2d_fwd(const float* x, const float* w, const float* bias, float* y, int B, int H, int W,
                  int Cin, int Cout, int KH, int KW, int stride, int pad, int act, void* stream) {
  int rc = check_dims("fd_conv2d_fwd", B, H, W, Cin, Cout, KH, KW, stride, pad);
  if (rc) return rc;
  ConvArgs a;
  a.x = x; a.w = w; a.bias = bias; a.y = y;
  a.B = B; a.Hg = H; a.Wg = W; a.Cg = Cin;
  a.Ho = (H + 2 * pad - KH) / stride + 1;
  a.Wo = (W + 2 * pad - KW) / stride + 1;
  a.N = Cout; a.KH = KH; a.KW = KW; a.stride = stride; a.pad = pad; a.act = act;
  a.M = (long)B * a.Ho * a.Wo;
  a.K = KH * KW * Cin;
  dim3 grid(fd::cdiv(a.M, BM), fd::cdiv(Cout, BN));
  if (Cin % 16 == 0)
    conv_igemm_kernel<0, true><<<grid, NTH, 0, (cudaStream_t)stream>>>(a);
  else
    conv_igemm_kernel<0, false><<<grid, NTH, 0, (cudaStream_t)stream>>>(a);
  FD_CHECK_LAUNCH();
  return 0;
}

int fd_conv2d_dgrad(const float* dy, const float* wt, float* dx, int B, int H, int W, int Cin,
                    int Cout, int KH, int KW, int stride, int pad, void* stream) {
  int rc = check_dims("fd_conv2d_dgrad", B, H, W, Cin, Cout, KH, KW, stride, pad);
  if (rc) return rc;
  ConvArgs a;
  a.x = dy; a.w = wt; a.bias = nullptr; a.y = dx;
  a.B = B;
  a.Hg = (H + 2 * pad - KH) / stride + 1;
  a.Wg = (W + 2 * pad - KW) / stride + 1;
  a.Cg = Cout;
  a.Ho = H; a.Wo = W; a.N = Cin;
  a.KH = KH; a.KW = KW; a.stride = stride; a.pad = pad; a.act = FD_ACT_NONE;
  a.M = (long)B * H * W;
  a.K = KH * KW * Cout;
  dim3 grid(fd::cdiv(a.M, BM), fd::cdiv(Cin, BN));
  if (Cout % 16 == 0)
    conv_igemm_kernel<1, true><<<grid, NTH, 0, (cudaStream_t)stream>>>(a);
  else
    conv_igemm_kernel<1, false><<<grid, NTH, 0, (cudaStream_t)stream>>>(a);
  FD_CHECK_LAUNCH();
  return 0;
}

int fd_conv2d_wgrad(const float* x, const float* dy, float* dw, int B, int H, int W, int Cin,
                    int Cout, int KH, int KW, int stride, int pad, void* stream) {
  int rc = check_dims("fd_conv2d_wgrad", B, H, W, Cin, Cout, KH, KW, stride, pad);
  if (rc) return rc;
  WgArgs a;
  a.x = x; a.dy = dy; a.dw = dw;
  a.B = B; a.H = H; a.W = W; a.Cin = Cin;
  a.Ho = (H + 2 * pad - KH) / stride + 1;
  a.Wo = (W + 2 * pad - KW) / stride + 1;
  a.N = Cout; a.KH = KH; a.KW = KW; a.stride = stride; a.pad = pad;
  a.M = (long)B * a.Ho * a.Wo;
  a.K = KH * KW * Cin;
  int tiles = fd::cdiv(a.K, BM) * fd::cdiv(Cout, BN);
  long want = (148L * 4 + tiles - 1) / tiles;          // ~4 waves of CTAs
  long max_split = (a.M + 4 * BK - 1) / (4 * BK);      // >= 64 pixels per slice
  long split = want < 1 ? 1 : (want > max_split ? max_split : want);
  if (split < 1) split = 1;
  a.m_per_split = ((a.M + split - 1) / split + BK - 1) / BK * BK;
  split = (a.M + a.m_per_split - 1) / a.m_per_split;
  dim3 grid(fd::cdiv(a.K, BM), fd::cdiv(Cout, BN), (unsigned)split);
  if (Cin % 4 == 0 && Cout % 4 == 0)
    conv_wgrad_kernel<true><<<grid, NTH, 0, (cudaStream_t)stream>>>(a);
  else
    conv_wgrad_kernel<false><<<grid, NTH, 0, (cudaStream_t)stream>>>(a);
  FD_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
