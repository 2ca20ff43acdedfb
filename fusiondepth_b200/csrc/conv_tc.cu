// Implicit-GEMM convolution on the 5th-generation tensor cores (tcgen05 / TMEM), forward and data
// gradient, at fp32 accuracy through error-compensated TF32 splitting ("3xTF32").
//
// Replaces aten::convolution / convolution_backward(input) under networks/* for every layer whose
// gathered channel count is a multiple of 32 and whose output channel count is a multiple of 16
// (all ResNet trunk 3x3/1x1 convs and most decoder convs); conv.cu keeps the rest.
//
// GEMM view (NHWC activations, weights [N][taps*Cg] K-major):
//     D[m][n] = sum_k A[m][k] * B[n][k],   m = output pixel, k = (tap, channel)
// CTA tile 128 (pixels) x BN (channels), K step 32 floats = one 128-byte swizzle row.
//
// Roles (416 threads, one CTA per SM):
//   warps 0-7  loaders: gather the im2col rows of A straight from global memory into a
//              SWIZZLE_128B shared-memory tile with cp.async (zero-fill = padding); one elected
//              thread fetches the two weight tiles (W, W_lo) with TMA (cp.async.bulk.tensor.2d);
//              both complete on the stage's `landed` mbarrier;
//   warps 8-11 splitters: derive the low-order tile A_lo = A - tf32(A) in shared memory, then
//              (after the K loop) run the epilogue: tcgen05.ld the accumulators (one pixel row per
//              thread), bias + activation, vectorised NHWC stores;
//   warp 12    one elected thread issues tcgen05.mma.kind::tf32 (the tensor core reads the top
//              19 bits of each fp32 word, so feeding the raw fp32 tile *is* feeding tf32(A)):
//                  corr  += A_lo*B + A*B_lo          (one TMEM accumulator)
//                  main_i += A*B                      (round-robin over up to 7 TMEM accumulators)
//              The tensor core truncates when it adds into an fp32 accumulator, which biases long
//              K reductions (measured 3e-5 at K = 4608 with one accumulator); keeping the tiny
//              correction terms apart and spreading the main products over several accumulators
//              that the epilogue sums with round-to-nearest adds restores fp32-level accuracy.
// Stage hand-off is mbarrier based (landed: cp.async + TMA tx; full: 128 splitter arrivals;
// empty / accumulator-ready: tcgen05.commit).  W_lo = W - tf32(W) comes from fd_tf32_split.
#include "tc_common.cuh"

namespace {

constexpr int NLOADW = 8, NLOAD = NLOADW * 32;      // loader warps / threads
constexpr int NSPLIT = 128;                          // splitter (+ epilogue) threads: 4 warps
constexpr int NTHREADS = NLOAD + NSPLIT + 32;        // + the MMA warp



template <int BN>
struct Cfg {
  static constexpr int B_TILE = BN * 128;
  static constexpr int STAGE = 2 * A_TILE + 2 * B_TILE;
  static constexpr int STAGES = (BN == 128) ? 3 : 4;
  // TMEM: accumulator 0 = correction terms, 1..NMAIN = main products (round robin)
  static constexpr int TMEM_COLS = BN == 128 ? 512 : (BN == 64 ? 512 : (BN == 32 ? 256 : 128));
  static constexpr int NMAIN = TMEM_COLS / BN - 1 > 7 ? 7 : TMEM_COLS / BN - 1;
  static constexpr int SMEM = STAGES * STAGE + 1024 /*align*/ + 256 /*barriers*/;
};

// MODE 0: forward conv.  MODE 1: data gradient (x = dy, w = transposed weights).
template <int BN, int MODE>
__global__ void __launch_bounds__(NTHREADS, 1)
conv_tc_kernel(TcArgs a, const __grid_constant__ CUtensorMap tm_w, const __grid_constant__ CUtensorMap tm_wlo) {
  using C = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = base + C::STAGES * C::STAGE;
  auto landed_bar = [&](int s) { return bars + 8u * s; };
  auto full_bar = [&](int s) { return bars + 8u * (C::STAGES + s); };
  auto empty_bar = [&](int s) { return bars + 8u * (2 * C::STAGES + s); };
  const uint32_t acc_bar = bars + 8u * (3 * C::STAGES);
  const uint32_t tmem_slot = acc_bar + 8u;
  auto a_raw = [&](int s) { return base + s * C::STAGE; };
  auto a_lo = [&](int s) { return base + s * C::STAGE + A_TILE; };
  auto b_raw = [&](int s) { return base + s * C::STAGE + 2 * A_TILE; };
  auto b_lo = [&](int s) { return base + s * C::STAGE + 2 * A_TILE + C::B_TILE; };

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long m0 = (long)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int nk = a.K / BK;

  if (tid == 0) {
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(landed_bar(s), NLOAD + 1);   // cp.async completions + 1 expect_tx arrive
      mbar_init(full_bar(s), NSPLIT);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(acc_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_w) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_wlo) : "memory");
  }
  if (warp == NLOADW + 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                 "n"(C::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp < NLOADW) {
    // ======================= loaders =======================
    // Everything that depends on the output pixel is hoisted out of the K loop: a base pointer per
    // row and a bit mask of the valid filter rows / columns (bits 0-7: kh, bits 8-15: kw).  Per
    // k-block a row costs a mask test, one pointer add and the cp.async.
    constexpr int ROWS = BM / (NLOAD / 8);          // rows per loader thread
    const int j = tid & 7, rg = tid >> 3;
    const float* rp[ROWS];
    uint32_t vm[ROWS];
    const int HoWo = a.Ho * a.Wo;
#pragma unroll
    for (int i = 0; i < ROWS; ++i) {
      const int m = (int)m0 + rg + (NLOAD / 8) * i;       // M < 2^31 (checked on the host)
      rp[i] = a.x;
      vm[i] = 0;
      if (m < (int)a.M) {
        int b = m / HoWo, r = m - b * HoWo;
        int ho = r / a.Wo, wo = r % a.Wo;
        int hb, wb, hq, wq;
        uint32_t mask = 0;
        if (MODE == 0) {
          hb = ho * a.stride - a.pad; wb = wo * a.stride - a.pad;
          hq = hb; wq = wb;
          for (int t = 0; t < a.KH; ++t) mask |= (uint32_t)(hb + t >= 0 && hb + t < a.Hg) << t;
          for (int t = 0; t < a.KW; ++t) mask |= (uint32_t)(wb + t >= 0 && wb + t < a.Wg) << (8 + t);
        } else {
          hb = ho + a.pad; wb = wo + a.pad;
          hq = hb / a.stride; wq = wb / a.stride;           // hb, wb >= 0
          for (int t = 0; t < a.KH; ++t) {
            int th = hb - t;
            mask |= (uint32_t)(th >= 0 && th % a.stride == 0 && th / a.stride < a.Hg) << t;
          }
          for (int t = 0; t < a.KW; ++t) {
            int tw = wb - t;
            mask |= (uint32_t)(tw >= 0 && tw % a.stride == 0 && tw / a.stride < a.Wg) << (8 + t);
          }
        }
        vm[i] = mask;
        rp[i] = a.x + (((long)b * a.Hg + hq) * a.Wg + wq) * a.Cg + j * 4;
      }
    }
    const uint32_t soff = (uint32_t)rg * 128u + (uint32_t)((j ^ (rg & 7)) << 4);   // (r & 7) == (rg & 7)
    int kh = 0, kw = 0, c0 = 0;
    for (int it = 0; it < nk; ++it) {
      const int s = it % C::STAGES;
      if (it >= C::STAGES) mbar_wait(empty_bar(s), ((it / C::STAGES) - 1) & 1);
      if (warp == 0 && elect_one()) {
        mbar_expect_tx(landed_bar(s), 2 * C::B_TILE);
        tma_load_2d(b_raw(s), &tm_w, it * BK, n0, landed_bar(s));
        tma_load_2d(b_lo(s), &tm_wlo, it * BK, n0, landed_bar(s));
      }
      // element offset of this (tap, channel block) relative to the row base pointer
      const long toff = MODE == 0 ? ((long)kh * a.Wg + kw) * a.Cg + c0
                                  : -((long)(kh / a.stride) * a.Wg + kw / a.stride) * a.Cg + c0;
      const uint32_t dst = a_raw(s) + soff;
#pragma unroll
      for (int i = 0; i < ROWS; ++i) {
        const bool ok = ((vm[i] >> kh) & (vm[i] >> (8 + kw)) & 1u) != 0;
        cp_async16(dst + (uint32_t)i * (NLOAD / 8) * 128u, ok ? rp[i] + toff : a.x, ok ? 16u : 0u);
      }
      cp_async_arrive_noinc(landed_bar(s));
      c0 += BK;
      if (c0 == a.Cg) { c0 = 0; if (++kw == a.KW) { kw = 0; ++kh; } }
    }
  } else if (warp < NLOADW + 4) {
    // ======================= splitters, then epilogue =======================
    const int t = tid - NLOAD;
    for (int it = 0; it < nk; ++it) {
      const int s = it % C::STAGES;
      mbar_wait(landed_bar(s), (it / C::STAGES) & 1);
      const uint32_t ar = a_raw(s), al = a_lo(s);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint32_t so = (uint32_t)(t + 128 * i) * 16u;   // the split is elementwise: any mapping
        float4 v;
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                     : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                     : "r"(ar + so));
        asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(al + so), "f"(lo_part(v.x)),
                     "f"(lo_part(v.y)), "f"(lo_part(v.z)), "f"(lo_part(v.w))
                     : "memory");
      }
      fence_async_proxy();
      mbar_arrive(full_bar(s));
    }
    // ---- epilogue ----
    const int ew = warp - NLOADW;
    mbar_wait(acc_bar, 0);
    tc_fence_after();
    const long m = m0 + ew * 32 + lane;
    const uint32_t trow = tmem_base + ((uint32_t)(ew * 32) << 16);
    const int nmain = nk < C::NMAIN ? nk : C::NMAIN;
#pragma unroll 1
    for (int c = 0; c < BN; c += 16) {
      float acc[16], tmp[16];
      tmem_ld16(trow + (uint32_t)(BN + c), acc);
      for (int q = 1; q < nmain; ++q) {
        tmem_ld16(trow + (uint32_t)((1 + q) * BN + c), tmp);
#pragma unroll
        for (int e = 0; e < 16; ++e) acc[e] += tmp[e];
      }
      tmem_ld16(trow + (uint32_t)c, tmp);
#pragma unroll
      for (int e = 0; e < 16; ++e) acc[e] += tmp[e];
      if (m < a.M && n0 + c < a.N) {
        float o[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          float bv = a.bias ? a.bias[n0 + c + q] : 0.f;
          o[q] = act_fn(acc[q] + bv, a.act);
        }
        float4* dst = reinterpret_cast<float4*>(a.y + m * a.N + n0 + c);
#pragma unroll
        for (int q = 0; q < 4; ++q) dst[q] = make_float4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
      }
    }
  } else {
    // ======================= MMA issuer =======================
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) |
                           ((uint32_t)(BM >> 4) << 24);
    for (int kb = 0; kb < nk; ++kb) {
      const int s = kb % C::STAGES;
      mbar_wait(full_bar(s), (kb / C::STAGES) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t da = make_desc(a_raw(s)), dal = make_desc(a_lo(s));
        const uint64_t db = make_desc(b_raw(s)), dbl = make_desc(b_lo(s));
        const uint32_t d_corr = tmem_base;
        const uint32_t d_main = tmem_base + (uint32_t)((1 + kb % C::NMAIN) * BN);
#pragma unroll
        for (int k = 0; k < BK / 8; ++k) {
          const uint64_t adv = (uint64_t)(k * 32 >> 4);
          umma_tf32(d_corr, dal + adv, db + adv, idesc, (kb | k) != 0);
          umma_tf32(d_corr, da + adv, dbl + adv, idesc, 1);
          umma_tf32(d_main, da + adv, db + adv, idesc, (kb >= C::NMAIN) || (k != 0));
        }
        umma_commit(empty_bar(s));
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(acc_bar);
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == NLOADW + 4) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "n"(C::TMEM_COLS)
                 : "memory");
  }
}


// ================================================================================================
// Weight gradient on the tensor cores:  dW[n][k] += sum_p dY[p][n] * X_im2col[p][k]
//
// GEMM view: D'[k][n] with M' = 128 consecutive k (four 32-channel groups of one or more filter
// taps), N' = BN output channels, reduction over the pixels p.  Both operands are "MN-major" for
// the tensor core (the reduction index p is the strided one): their shared-memory image is the
// [pixel rows x 128 B] block per 32 channels, in the SWIZZLE_128B_BASE32B pattern that MN-major tf32
// operands require.
//   A' = gathered input pixels (cp.async, zero-fill outside the image) -> 4 blocks of [32 px x 128 B]
//   B' = dY rows (TMA 2-D boxes, rows past M zero-filled by the TMA unit) -> BN/32 blocks
// 3xTF32: both low-order tiles are produced by the splitter warps.  The pixel range is split over
// blockIdx.z; partial results are added to dW with coalesced fp32 reductions (red.global.add).
// ================================================================================================
constexpr int BP = 32;                       // pixels (reduction elements) per stage
constexpr int BLK = BP * 128;                // bytes of one [32 px x 32 ch] block

template <int BN>
struct WgCfg {
  static constexpr int NB = BN / 32;                       // dY column blocks
  static constexpr int A_BYTES = 4 * BLK, B_BYTES = NB * BLK;
  static constexpr int STAGE = 2 * A_BYTES + 2 * B_BYTES;
  static constexpr int STAGES = (BN == 128) ? 3 : 4;
  static constexpr int TMEM_COLS = BN == 128 ? 512 : (BN == 64 ? 512 : 256);
  static constexpr int NMAIN = TMEM_COLS / BN - 1 > 7 ? 7 : TMEM_COLS / BN - 1;
  static constexpr int SMEM = STAGES * STAGE + 1024 + 256;
};

// Roles (576 threads): warps 0-7 input-pixel gather (cp.async), 8-15 splitters then epilogue, 16 MMA
// issue, 17 dY tiles by TMA.
constexpr int WG_NSPLITW = 8, WG_NSPLIT = WG_NSPLITW * 32;
constexpr int WG_MMA_WARP = NLOADW + WG_NSPLITW, WG_TMA_WARP = WG_MMA_WARP + 1;
constexpr int WG_NTHREADS = NLOAD + WG_NSPLIT + 64;

template <int BN>
__global__ void __launch_bounds__(WG_NTHREADS, 1)
conv_wgrad_tc_kernel(WgTcArgs a, const __grid_constant__ CUtensorMap tm_dy) {
  using C = WgCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = base + C::STAGES * C::STAGE;
  auto landed_bar = [&](int s) { return bars + 8u * s; };
  auto full_bar = [&](int s) { return bars + 8u * (C::STAGES + s); };
  auto empty_bar = [&](int s) { return bars + 8u * (2 * C::STAGES + s); };
  const uint32_t acc_bar = bars + 8u * (3 * C::STAGES);
  const uint32_t tmem_slot = acc_bar + 8u;
  auto a_raw = [&](int s) { return base + s * C::STAGE; };
  auto a_lo = [&](int s) { return base + s * C::STAGE + C::A_BYTES; };
  auto b_raw = [&](int s) { return base + s * C::STAGE + 2 * C::A_BYTES; };
  auto b_lo = [&](int s) { return base + s * C::STAGE + 2 * C::A_BYTES + C::B_BYTES; };

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int k0 = blockIdx.x * 128, n0 = blockIdx.y * BN;
  const int pbeg = blockIdx.z * a.p_per_split;
  const int pend = min(a.M, pbeg + a.p_per_split);
  const int nst = (pend - pbeg + BP - 1) / BP;

  if (tid == 0) {
    for (int s = 0; s < C::STAGES; ++s) {
      // elected per-warp arrivals (flags bit 0): 32 lanes arriving on one mbarrier serialise
      mbar_init(landed_bar(s), ((a.flags & 1) ? NLOADW : NLOAD) + 1);
      mbar_init(full_bar(s), WG_NSPLITW);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(acc_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_dy) : "memory");
  }
  if (warp == WG_MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                 "n"(C::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp < NLOADW) {
    // ======================= loaders: thread = (pixel row rg, 16-byte chunk j) =======================
    const int j = tid & 7, rg = tid >> 3;            // 32 rows x 8 chunks = 256 threads
    int kh[4], kw[4], cc[4];
    bool kok[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      int k = k0 + 32 * c;
      kok[c] = k < a.K;
      int tap = kok[c] ? k / a.Cin : 0;
      cc[c] = k - tap * a.Cin;
      kh[c] = tap / a.KW; kw[c] = tap - kh[c] * a.KW;
    }
    const uint32_t soff = (uint32_t)rg * 128u + (uint32_t)((((j >> 1) ^ (rg & 3)) << 5) | ((j & 1) << 4));
    const bool groups = (a.flags & 1) != 0;
    // this thread's pixel (pbeg + rg, then +32 per stage) as (b, ho, wo), advanced without divisions;
    // 32-bit element offsets (host: |x| < 2^31 elements)
    int pb = 0, pho = 0, pwo = 0;
    {
      const int p = pbeg + rg;
      const int HoWo = a.Ho * a.Wo;
      pb = p / HoWo;
      const int r = p - pb * HoWo;
      pho = r / a.Wo;
      pwo = r - pho * a.Wo;
    }
    int tapoff[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) tapoff[c] = (kh[c] * a.W + kw[c]) * a.Cin + cc[c] + j * 4;
    for (int it = 0; it < nst; ++it) {
      const int s = it % C::STAGES;
      if (groups && it >= 2) {                   // the group issued two stages ago has landed
        cp_async_wait<1>();
        __syncwarp();
        if (elect_one()) mbar_arrive(landed_bar((it - 2) % C::STAGES));
      }
      if (it >= C::STAGES) mbar_wait(empty_bar(s), ((it / C::STAGES) - 1) & 1);
      const bool pok = pbeg + it * BP + rg < pend;
      const int hb = pho * a.stride - a.pad, wb = pwo * a.stride - a.pad;
      const int ebase = ((pb * a.H + hb) * a.W + wb) * a.Cin;
      const uint32_t dst = a_raw(s) + soff;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int h = hb + kh[c], w = wb + kw[c];
        const bool ok = pok && kok[c] && h >= 0 && h < a.H && w >= 0 && w < a.W;
        cp_async16(dst + c * BLK, a.x + (ok ? ebase + tapoff[c] : 0), ok ? 16u : 0u);
      }
      pwo += BP;
      while (pwo >= a.Wo) { pwo -= a.Wo; if (++pho == a.Ho) { pho = 0; ++pb; } }
      if (groups) cp_async_commit();
      else cp_async_arrive_noinc(landed_bar(s));
    }
    if (groups) {
      cp_async_wait<0>();
      __syncwarp();
      if (lane == 0)
        for (int it = nst > 2 ? nst - 2 : 0; it < nst; ++it) mbar_arrive(landed_bar(it % C::STAGES));
    }
  } else if (warp < WG_MMA_WARP) {
    // ======================= splitters, then epilogue =======================
    const int t = tid - NLOAD;
    constexpr int NV = (C::A_BYTES + C::B_BYTES) / 16 / WG_NSPLIT;  // float4 per thread per stage
    for (int it = 0; it < nst; ++it) {
      const int s = it % C::STAGES;
      mbar_wait(landed_bar(s), (it / C::STAGES) & 1);
      // (dY rows past `pend` meet all-zero A' rows, rows past M are zero-filled by the TMA unit)
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const uint32_t idx = (uint32_t)(t + WG_NSPLIT * i);
        const bool isA = idx < C::A_BYTES / 16;
        const uint32_t so = (isA ? idx : idx - C::A_BYTES / 16) * 16u;
        const uint32_t src = (isA ? a_raw(s) : b_raw(s)) + so, dst = (isA ? a_lo(s) : b_lo(s)) + so;
        float4 v;
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                     : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                     : "r"(src));
        if ((a.dbg == 2 && isA) || (a.dbg == 3 && !isA) || a.dbg == 4) {
          v = make_float4(1.f, 1.f, 1.f, 1.f);
          asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(src), "f"(1.f), "f"(1.f), "f"(1.f), "f"(1.f) : "memory");
        }
        asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(dst), "f"(lo_part(v.x)),
                     "f"(lo_part(v.y)), "f"(lo_part(v.z)), "f"(lo_part(v.w))
                     : "memory");
      }
      fence_async_proxy();
      __syncwarp();
      if (elect_one()) mbar_arrive(full_bar(s));
    }
    // epilogue: warp (q, half) owns k rows 32q..32q+31 and the 16-column chunks half, half+2, ...
    const int ew = warp - NLOADW, q = ew & 3, half = ew >> 2;
    mbar_wait(acc_bar, 0);
    tc_fence_after();
    const int k = k0 + q * 32 + lane;
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    const int nmain = nst < C::NMAIN ? nst : C::NMAIN;
#pragma unroll 1
    for (int c = half * 16; c < BN; c += 32) {
      // accumulators in summation order: main 0..nmain-1, then the correction terms
      float acc[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) acc[e] = 0.f;
      const int nacc = (a.flags & 0x800) ? nmain : nmain + 1;      // single-pass TF32: no correction accumulator
      for (int g = 0; g < nacc; g += 4) {
        uint32_t v[4][16];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int i = g + u;
          if (i < nacc) tmem_ld16_nowait(trow + (uint32_t)((i < nmain ? (1 + i) * BN : 0) + c), v[u]);
        }
        tmem_wait_ld();
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (g + u < nacc) {
#pragma unroll
            for (int e = 0; e < 16; ++e) acc[e] += __uint_as_float(v[u][e]);
          }
        }
      }
      if (a.dbg == 1) {
#pragma unroll
        for (int e = 0; e < 16; ++e) acc[e] = 1.0f;
      }
      if (k < a.K && nst > 0) {
#pragma unroll
        for (int e = 0; e < 16; ++e)
          if (n0 + c + e < a.N) atomicAdd(a.dw + (long)(n0 + c + e) * a.K + k, acc[e]);
      }
    }
  } else if (warp == WG_TMA_WARP) {
    // ======================= dY tiles by TMA =======================
    for (int it = 0; it < nst; ++it) {
      const int s = it % C::STAGES;
      if (it >= C::STAGES) mbar_wait(empty_bar(s), ((it / C::STAGES) - 1) & 1);
      if (elect_one()) {
        const int p0 = pbeg + it * BP;
        mbar_expect_tx(landed_bar(s), C::B_BYTES);
#pragma unroll
        for (int nb = 0; nb < C::NB; ++nb) tma_load_2d(b_raw(s) + nb * BLK, &tm_dy, n0 + 32 * nb, p0, landed_bar(s));
      }
      __syncwarp();
    }
  } else {
    // ======================= MMA issuer =======================
    // converged warp, one elected lane issues (see elect_one in tc_common.cuh)
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) |
                           ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    for (int it = 0; it < nst; ++it) {
      const int s = it % C::STAGES;
      mbar_wait(full_bar(s), (it / C::STAGES) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t da = make_desc_mn(a_raw(s), BLK), dal = make_desc_mn(a_lo(s), BLK);
        const uint64_t db = make_desc_mn(b_raw(s), BLK), dbl = make_desc_mn(b_lo(s), BLK);
        const uint32_t d_corr = tmem_base;
        const uint32_t d_main = tmem_base + (uint32_t)((1 + it % C::NMAIN) * BN);
#pragma unroll
        for (int k = 0; k < BP / 8; ++k) {
          const uint64_t adv = (uint64_t)(k * 1024 >> 4);        // next 8-pixel row group
          if (!(a.flags & 0x800)) {
            umma_tf32(d_corr, dal + adv, db + adv, idesc, (it | k) != 0);
            umma_tf32(d_corr, da + adv, dbl + adv, idesc, 1);
          }
          umma_tf32(d_main, da + adv, db + adv, idesc, (it >= C::NMAIN) || (k != 0));
        }
        umma_commit(empty_bar(s));
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(acc_bar);
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == WG_MMA_WARP) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "n"(C::TMEM_COLS)
                 : "memory");
  }
}

__global__ void tf32_split_kernel(const float* __restrict__ w, float* __restrict__ lo, long n) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    float v = w[i];
    lo[i] = v - __uint_as_float(__float_as_uint(v) & 0xffffe000u);
  }
}

// w [Cout][taps][Cin] -> wt [Cin][taps][Cout] and its low-order part
__global__ void transpose_split_kernel(const float* __restrict__ w, float* __restrict__ wt,
                                       float* __restrict__ wtlo, int Cout, int taps, int Cin) {
  long n = (long)Cout * taps * Cin;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    int co = i % Cout;
    int tp = (i / Cout) % taps;
    int ci = i / ((long)Cout * taps);
    float v = w[((long)co * taps + tp) * Cin + ci];
    wt[i] = v;
    wtlo[i] = v - __uint_as_float(__float_as_uint(v) & 0xffffe000u);
  }
}

// The same for every conv weight of a flat parameter buffer in one launch.  desc[j] = {first output
// element of tensor j in the concatenated index space, offset of the tensor in the flat buffers,
// Cout, taps, Cin}; outputs keep the tensor's offset.
__global__ void transpose_split_batched_kernel(const float* __restrict__ flat, float* __restrict__ flat_t,
                                               float* __restrict__ flat_tlo, const long* __restrict__ desc,
                                               int n, long total) {
  extern __shared__ long starts[];
  for (int j = threadIdx.x; j < n; j += blockDim.x) starts[j] = desc[5 * j];
  __syncthreads();
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    int lo = 0, hi = n - 1;
    while (lo < hi) {                       // last j with starts[j] <= i
      const int mid = (lo + hi + 1) >> 1;
      if (starts[mid] <= i) lo = mid; else hi = mid - 1;
    }
    const long* d = desc + 5 * lo;
    const long off = d[1];
    const int Cout = (int)d[2], taps = (int)d[3], Cin = (int)d[4];
    const int li = (int)(i - d[0]);
    const int co = li % Cout;
    const int tp = (li / Cout) % taps;
    const int ci = li / (Cout * taps);
    const float v = flat[off + ((long)co * taps + tp) * Cin + ci];
    flat_t[off + li] = v;
    flat_tlo[off + li] = v - __uint_as_float(__float_as_uint(v) & 0xffffe000u);
  }
}

template <int BN, int MODE>
int launch_tc(const TcArgs& a, cudaStream_t st) {
  using C = Cfg<BN>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<BN, MODE>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
    if (e != cudaSuccess) {
      fd::set_error("conv_tc: cannot reserve %d B of shared memory: %s", C::SMEM, cudaGetErrorString(e));
      return 1;
    }
    configured = true;
  }
  CUtensorMap tw, twl;
  int rc = make_map_2d(&tw, a.w, a.N, a.K, BN);
  if (rc) return rc;
  rc = make_map_2d(&twl, a.wlo, a.N, a.K, BN);
  if (rc) return rc;
  dim3 grid(fd::cdiv(a.M, BM), fd::cdiv(a.N, BN));
  conv_tc_kernel<BN, MODE><<<grid, NTHREADS, C::SMEM, st>>>(a, tw, twl);
  FD_CHECK_LAUNCH();
  return 0;
}

template <int MODE>
int dispatch_tc(const TcArgs& a, cudaStream_t st) {
  FD_REQUIRE((((uintptr_t)a.w | (uintptr_t)a.wlo | (uintptr_t)a.x) & 15) == 0,
             "conv_tc: operands must be 16-byte aligned");
  FD_REQUIRE(a.M < (1L << 31) && a.KH <= 8 && a.KW <= 8, "conv_tc: problem too large (M=%ld, %dx%d)", a.M,
             a.KH, a.KW);
  // Tile width: the widest BN dividing N.  (Narrower tiles for the few-pixel layers were measured
  // slower once the six trunks run concurrently: every extra N-tile repeats the A gather + split.)
  const int bn = a.N % 128 == 0 ? 128 : (a.N % 64 == 0 ? 64 : (a.N % 32 == 0 ? 32 : 16));
  if (bn == 128) return launch_tc<128, MODE>(a, st);
  if (bn == 64) return launch_tc<64, MODE>(a, st);
  if (bn == 32) return launch_tc<32, MODE>(a, st);
  return launch_tc<16, MODE>(a, st);
}


template <int BN>
int launch_wgrad_tc(const WgTcArgs& a0, const float* dy, cudaStream_t st) {
  using C = WgCfg<BN>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_wgrad_tc_kernel<BN>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
    if (e != cudaSuccess) {
      fd::set_error("conv_wgrad_tc: cannot reserve %d B of shared memory: %s", C::SMEM, cudaGetErrorString(e));
      return 1;
    }
    configured = true;
  }
  WgTcArgs a = a0;
  CUtensorMap tdy;
  int rc = make_map_2d(&tdy, dy, a.M, a.N, BP, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
  if (rc) return rc;
  const int tiles = fd::cdiv(a.K, 128) * (a.N / BN);
  // pixel-range splits: enough CTAs for `waves` waves of 148 SMs.  A CTA pays ~8k cycles of prologue,
  // pipeline fill and atomic epilogue whatever its share, so half a wave of longer CTAs beats two (measured 328 vs 310 img/s):
  // the other streams of the step fill the remaining SMs.
  static int waves_x2 = -1;                 // FD_WGRAD_WAVES (multiples of 0.5, default 0.5), kept as 2x integer
  if (waves_x2 < 0) {
    const char* e = getenv("FD_WGRAD_WAVES");
    waves_x2 = e ? (int)(2.0 * atof(e) + 0.5) : 1;
    if (waves_x2 < 1) waves_x2 = 1;
  }
  int splits = (waves_x2 * 74 + tiles - 1) / tiles;
  const int max_splits = fd::cdiv(a.M, 4 * BP);
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  a.p_per_split = fd::cdiv(fd::cdiv(a.M, splits), BP) * BP;
  splits = fd::cdiv(a.M, a.p_per_split);
  dim3 grid(fd::cdiv(a.K, 128), a.N / BN, splits);
  conv_wgrad_tc_kernel<BN><<<grid, WG_NTHREADS, C::SMEM, st>>>(a, tdy);
  FD_CHECK_LAUNCH();
  return 0;
}

}  // namespace

// FD_CONV_TC=v1 keeps both operands in shared memory (this file); default: A operand in TMEM (conv_tc2.cu)
static int tc2_flags() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("FD_TC2_FLAGS");
    v = e ? atoi(e) : 0;
    // FD_CONV_PRECISION=tf32: single-pass TF32 "fast mode" (the arithmetic the reference gets from cuDNN on a
    // GPU, cudnn.allow_tf32=True); default 3xTF32 = fp32-level accuracy, which the parity tests require.
    const char* p = getenv("FD_CONV_PRECISION");
    if (p && p[0] == 't' && p[1] == 'f') v |= 0x800;
    const char* f = getenv("FD_TC_FENCE");
    if (f && f[0] == '1') v |= 0x1000;
  }
  return v;
}
static bool tc_use_v2() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("FD_CONV_TC");
    v = (e && e[0] == 'v' && e[1] == '1') ? 0 : 1;
  }
  return v == 1;
}

static bool tc_use_v3() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("FD_CONV_TC3");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}
static bool tc_use_v4() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("FD_CONV_TC4");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}
// conv_tc4 (persistent patch variant) for many-tile launches, conv_tc3 (patch variant) where it applies, else conv_tc2
static int tc_v2_or_v3(const TcArgs& a, int mode, cudaStream_t st) {
  if (tc_use_v3() && !(a.flags & ~0x1800)) {
    if (tc_use_v4()) {
      int rc = fd::conv_tc4_dispatch(a, mode, st);
      if (rc >= 0) return rc;
    }
    int rc = fd::conv_tc3_dispatch(a, mode, st);
    if (rc >= 0) return rc;
  }
  return fd::conv_tc2_dispatch(a, mode, st);
}

extern "C" {

int fd_conv2d_tc_supported(int Cin, int Cout) { return (Cin % 32 == 0) && (Cout % 16 == 0); }

int fd_tf32_split(const float* w, float* w_lo, long n, void* stream) {
  tf32_split_kernel<<<min(fd::cdiv(n, 256), 148 * 8), 256, 0, (cudaStream_t)stream>>>(w, w_lo, n);
  FD_CHECK_LAUNCH();
  return 0;
}

int fd_weight_transpose_split(const float* w, float* wt, float* wt_lo, int Cout, int taps, int Cin,
                              void* stream) {
  long n = (long)Cout * taps * Cin;
  transpose_split_kernel<<<min(fd::cdiv(n, 256), 148 * 8), 256, 0, (cudaStream_t)stream>>>(
      w, wt, wt_lo, Cout, taps, Cin);
  FD_CHECK_LAUNCH();
  return 0;
}

int fd_weight_transpose_split_batched(const float* flat, float* flat_t, float* flat_tlo, const long* desc,
                                      int n, long total, void* stream) {
  FD_REQUIRE(n > 0 && n <= 4096, "fd_weight_transpose_split_batched: %d tensors", n);
  transpose_split_batched_kernel<<<min(fd::cdiv(total, 256), 148 * 16), 256, n * sizeof(long),
                                   (cudaStream_t)stream>>>(flat, flat_t, flat_tlo, desc, n, total);
  FD_CHECK_LAUNCH();
  return 0;
}

int fd_conv2d_fwd_tc_stats(const float* x, const float* w, const float* w_lo, const float* bias, float* y,
                           int B, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad,
                           int act, double* stats, void* stream);

int fd_conv2d_fwd_tc(const float* x, const float* w, const float* w_lo, const float* bias, float* y,
                     int B, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad,
                     int act, void* stream) {
  return fd_conv2d_fwd_tc_stats(x, w, w_lo, bias, y, B, H, W, Cin, Cout, KH, KW, stride, pad, act, nullptr,
                                stream);
}

int fd_conv2d_fwd_tc_stats(const float* x, const float* w, const float* w_lo, const float* bias, float* y,
                           int B, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad,
                           int act, double* stats, void* stream) {
  FD_REQUIRE(fd_conv2d_tc_supported(Cin, Cout),
             "fd_conv2d_fwd_tc: needs Cin %% 32 == 0 and Cout %% 16 == 0 (got %d, %d)", Cin, Cout);
  FD_REQUIRE(stats == nullptr || (tc_use_v2() && bias == nullptr && act == FD_ACT_NONE),
             "fd_conv2d_fwd_tc_stats: channel statistics need the conv_tc2 kernels, no bias and no activation");
  TcArgs a{};
  a.trace = nullptr;
  a.stats = stats;
  a.x = x; a.w = w; a.wlo = w_lo; a.bias = bias; a.y = y;
  a.B = B; a.Hg = H; a.Wg = W; a.Cg = Cin;
  a.Ho = (H + 2 * pad - KH) / stride + 1;
  a.Wo = (W + 2 * pad - KW) / stride + 1;
  a.N = Cout; a.KH = KH; a.KW = KW; a.stride = stride; a.pad = pad; a.act = act;
  a.M = (long)B * a.Ho * a.Wo;
  a.K = KH * KW * Cin;
  a.flags = tc2_flags();
  if (tc_use_v2()) return tc_v2_or_v3(a, 0, (cudaStream_t)stream);
  return dispatch_tc<0>(a, (cudaStream_t)stream);
}

int fd_conv2d_dgrad_tc(const float* dy, const float* wt, const float* wt_lo, float* dx, int B, int H,
                       int W, int Cin, int Cout, int KH, int KW, int stride, int pad, void* stream) {
  FD_REQUIRE(fd_conv2d_tc_supported(Cout, Cin),
             "fd_conv2d_dgrad_tc: needs Cout %% 32 == 0 and Cin %% 16 == 0 (got %d, %d)", Cout, Cin);
  TcArgs a{};
  a.trace = nullptr;
  a.stats = nullptr;
  a.x = dy; a.w = wt; a.wlo = wt_lo; a.bias = nullptr; a.y = dx;
  a.B = B;
  a.Hg = (H + 2 * pad - KH) / stride + 1;
  a.Wg = (W + 2 * pad - KW) / stride + 1;
  a.Cg = Cout;
  a.Ho = H; a.Wo = W; a.N = Cin;
  a.KH = KH; a.KW = KW; a.stride = stride; a.pad = pad; a.act = FD_ACT_NONE;
  a.M = (long)B * H * W;
  a.K = KH * KW * Cout;
  a.flags = tc2_flags();
  if (tc_use_v2() && tc_use_v3() && stride == 2 && !(a.flags & ~0x1800)) {
    // the stride-2 data gradient reads 1, 2, 2 or 4 of the 9 taps depending on the output pixel's parity: four
    // stride-1 convolutions over dY instead of one that multiplies 3/4 zeros
    static int use_s2 = -1;
    if (use_s2 < 0) {
      const char* e = getenv("FD_DGRAD_S2");
      use_s2 = (e && e[0] == '0') ? 0 : 1;
    }
    if (use_s2) {
      int rc = fd::conv_tc3_dgrad_s2(a, (cudaStream_t)stream);
      if (rc >= 0) return rc;
    }
  }
  if (tc_use_v2()) return tc_v2_or_v3(a, 1, (cudaStream_t)stream);
  return dispatch_tc<1>(a, (cudaStream_t)stream);
}

int fd_conv2d_wgrad_tc(const float* x, const float* dy, float* dw, int B, int H, int W, int Cin,
                       int Cout, int KH, int KW, int stride, int pad, void* stream) {
  FD_REQUIRE(Cin % 32 == 0 && Cout % 32 == 0,
             "fd_conv2d_wgrad_tc: needs Cin %% 32 == 0 and Cout %% 32 == 0 (got %d, %d)", Cin, Cout);
  WgTcArgs a;
  a.x = x; a.dw = dw;
  a.B = B; a.H = H; a.W = W; a.Cin = Cin;
  a.Ho = (H + 2 * pad - KH) / stride + 1;
  a.Wo = (W + 2 * pad - KW) / stride + 1;
  a.N = Cout; a.KH = KH; a.KW = KW; a.stride = stride; a.pad = pad;
  long M = (long)B * a.Ho * a.Wo;
  FD_REQUIRE(M < (1L << 31), "fd_conv2d_wgrad_tc: too many pixels");
  FD_REQUIRE((long)B * H * W * Cin < (1L << 31), "fd_conv2d_wgrad_tc: input has 2^31 or more elements");
  a.M = (int)M;
  a.K = KH * KW * Cin;
  a.p_per_split = 0;
  a.dbg = getenv("FD_WGRAD_DEBUG") ? atoi(getenv("FD_WGRAD_DEBUG")) : 0;
  a.flags = tc2_flags();
  FD_REQUIRE((((uintptr_t)x | (uintptr_t)dy | (uintptr_t)dw) & 15) == 0, "conv_wgrad_tc: operands must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  // default: Xg operand in tensor memory (conv_wgrad2.cu); FD_WGRAD2=0 keeps both operands in shared memory
  static int use_v2 = -1;
  if (use_v2 < 0) {
    const char* e = getenv("FD_WGRAD2");
    use_v2 = (e && e[0] == '0') ? 0 : 1;
  }
  if (use_v2 && !(a.flags & ~0x1800)) {
    int rc = fd::conv_wgrad2_dispatch(a, dy, st);
    if (rc >= 0) return rc;
  }
  if (Cout % 128 == 0) return launch_wgrad_tc<128>(a, dy, st);
  if (Cout % 64 == 0) return launch_wgrad_tc<64>(a, dy, st);
  return launch_wgrad_tc<32>(a, dy, st);
}

}  // extern "C"
