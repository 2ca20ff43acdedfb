// Weight gradient on tcgen05 with the gathered-input operand in TENSOR MEMORY (TS-mode MMAs).
//
//     dW[n][k] = sum_p dY[p][n] * Xg[p][k],   k = (tap, input channel), p = output pixel, 3xTF32
//
// conv_wgrad_tc_kernel (conv_tc.cu) keeps both operands in shared memory, MN-major.  ncu put its tensor pipe at
// 46 % (BN = 128) / 34 % (BN = 64) of the active cycles, and that IS its shared-memory roofline: per 32-pixel stage
// the 12 MMAs read 12 x (4 KB of A' + BN/32 KB of B'), the splitters read and write both tiles once more and the
// fills write them -- 192 KB (BN = 128) through a 128 B/clk port = 1536 clk against 768 clk of MMAs.
//
// Here the gathered input tile goes where conv_tc2 / conv_tc3 put their A operand: splitter warps read it out of a
// plain [32 pixels][32 channels] staging block (lane = channel, one LDS.32 per pixel: the transpose is free),
// and write Xg and Xg_lo = Xg - tf32(Xg) into a ring of tensor-memory slots (lane = k, column = pixel); the MMAs
// read A from TMEM and only dY / dY_lo from shared memory: 128 KB per stage at BN = 128 (1024 clk), 80 KB at
// BN = 64 (640 clk against 384 clk of MMAs).  For BN <= 64, [dY ; dY_lo] is one N = 2 BN operand feeding a
// [main | corr] accumulator pair (8 MMAs per stage instead of 12).  Two MMA issuer warps take alternate stages
// (tcgen05.mma issue blocks until the pipe accepts it; see conv_tc3.cu).
//
// Roles (736 threads): warps 0-7 gather Xg with cp.async, 8-15 Xg splitters (two groups on alternate stages) then
// the epilogue, 16-19 dY_lo splitters, 20 / 21 MMA issuers, 22 dY tiles by TMA.  The pixel range is split over
// blockIdx.z; partial results are added to dW with coalesced fp32 reductions, as before.
#include "tc_common.cuh"

namespace {

constexpr int W2_BP = 32;                      // pixels (reduction elements) per stage
constexpr int W2_BLK = W2_BP * 128;            // bytes of one [32 px x 32 ch] block
constexpr int W2_LOADW = 8, W2_NLOAD = W2_LOADW * 32;
constexpr int W2_ASPLITW = 8, W2_BSPLITW = 4;
constexpr int W2_ASPLIT0 = W2_LOADW, W2_BSPLIT0 = W2_ASPLIT0 + W2_ASPLITW;      // 8, 16
constexpr int W2_MMA_WARP = W2_BSPLIT0 + W2_BSPLITW;                             // 20 (and 21)
constexpr int W2_TMA_WARP = W2_MMA_WARP + 2;                                     // 22
constexpr int W2_NTHREADS = (W2_TMA_WARP + 1) * 32;                              // 736

template <int BN>
struct W2Cfg {
  static constexpr int NB = BN / 32;                        // dY column blocks
  static constexpr int A_BYTES = 4 * W2_BLK;                // 4 k-blocks of 32 channels x 32 pixels
  static constexpr int B_BYTES = NB * W2_BLK;
  static constexpr int STAGE = A_BYTES + 2 * B_BYTES;       // Xg | dY | dY_lo
  static constexpr int STAGES = BN == 128 ? 4 : (BN == 64 ? 6 : 8);
  static constexpr bool PAIR = BN <= 64;
  static constexpr int ACC0 = 2 * 64;                       // two TMEM slots of [Xg | Xg_lo] x 32 pixels
  static constexpr int NMAIN = BN == 128 ? 2 : 3;
  static constexpr int ACCCOLS = PAIR ? NMAIN * 2 * BN : (1 + NMAIN) * BN;
  static constexpr int TMEM_COLS = ACC0 + ACCCOLS <= 256 ? 256 : 512;
  static_assert(ACC0 + ACCCOLS <= 512, "tensor memory");
  static constexpr int SMEM = STAGES * STAGE + 1024 /*align*/ + 512 /*barriers*/;
};

template <int BN>
__global__ void __launch_bounds__(W2_NTHREADS, 1)
conv_wgrad2_kernel(WgTcArgs a, const __grid_constant__ CUtensorMap tm_dy) {
  using C = W2Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = base + C::STAGES * C::STAGE;
  auto aland_bar = [&](int s) { return bars + 8u * s; };                         // Xg stage gathered (cp.async)
  auto afree_bar = [&](int s) { return bars + 8u * (8 + s); };                   // splitters have read it
  auto bland_bar = [&](int s) { return bars + 8u * (16 + s); };                  // dY tile landed (TMA tx)
  auto bready_bar = [&](int s) { return bars + 8u * (24 + s); };                 // dY_lo written
  auto bfree_bar = [&](int s) { return bars + 8u * (32 + s); };                  // MMAs done with dY / dY_lo
  auto tfull_bar = [&](int t) { return bars + 8u * (40 + t); };                  // Xg / Xg_lo of a stage in TMEM
  auto tfree_bar = [&](int t) { return bars + 8u * (42 + t); };
  auto turn_bar = [&](int i) { return bars + 8u * (44 + i); };
  const uint32_t acc_bar = bars + 8u * 46;
  const uint32_t tmem_slot = bars + 8u * 47;
  auto a_stage = [&](int s) { return base + s * C::STAGE; };
  auto b_raw = [&](int s) { return base + s * C::STAGE + C::A_BYTES; };
  auto b_lo = [&](int s) { return base + s * C::STAGE + C::A_BYTES + C::B_BYTES; };

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int k0 = blockIdx.x * 128, n0 = blockIdx.y * BN;
  const int pbeg = blockIdx.z * a.p_per_split;
  const int pend = min(a.M, pbeg + a.p_per_split);
  const int nst = (pend - pbeg + W2_BP - 1) / W2_BP;

  if (tid == 0) {
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(aland_bar(s), W2_NLOAD);
      mbar_init(afree_bar(s), W2_ASPLITW / 2);
      mbar_init(bland_bar(s), 1);
      mbar_init(bready_bar(s), W2_BSPLITW);
      mbar_init(bfree_bar(s), 1);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(tfull_bar(t), W2_ASPLITW / 2);
      mbar_init(tfree_bar(t), 1);
      mbar_init(turn_bar(t), 1);
    }
    mbar_init(acc_bar, 2);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_dy) : "memory");
  }
  if (warp == W2_MMA_WARP) {
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

  if (warp < W2_LOADW) {
    // ======================= loaders: thread = (pixel row rg, 16-byte chunk j), all four k-blocks =======================
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
    const uint32_t soff = (uint32_t)rg * 128u + (uint32_t)(j << 4);     // plain rows: the splitters read columns
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
      if (it >= C::STAGES) mbar_wait(afree_bar(s), ((it / C::STAGES) - 1) & 1);
      const bool pok = pbeg + it * W2_BP + rg < pend;
      const int hb = pho * a.stride - a.pad, wb = pwo * a.stride - a.pad;
      const int ebase = ((pb * a.H + hb) * a.W + wb) * a.Cin;
      const uint32_t dst = a_stage(s) + soff;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int h = hb + kh[c], w = wb + kw[c];
        const bool ok = pok && kok[c] && h >= 0 && h < a.H && w >= 0 && w < a.W;
        cp_async16(dst + c * W2_BLK, a.x + (ok ? ebase + tapoff[c] : 0), ok ? 16u : 0u);
      }
      pwo += W2_BP;
      while (pwo >= a.Wo) { pwo -= a.Wo; if (++pho == a.Ho) { pho = 0; ++pb; } }
      cp_async_arrive_noinc(aland_bar(s));
    }
  } else if (warp < W2_BSPLIT0) {
    // ======================= Xg splitters: warp (group, q) moves k-block q of the stages it % 2 == group =======================
    const int ew = warp - W2_ASPLIT0, q = ew & 3, group = ew >> 2;
    const uint32_t tcol = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(group * 64);
    uint32_t nuse = 0;
    for (int it = group; it < nst; it += 2, ++nuse) {
      const int s = it % C::STAGES;
      mbar_wait(aland_bar(s), (it / C::STAGES) & 1);
      // lane = channel of this k-block, one word per pixel row: conflict-free, and the [pixel][channel] ->
      // [channel lane][pixel column] transpose costs nothing
      const uint32_t src = a_stage(s) + (uint32_t)q * W2_BLK + (uint32_t)lane * 4u;
      uint32_t hi[32], lo[32];
#pragma unroll
      for (int p = 0; p < 32; ++p)
        asm volatile("ld.shared.b32 %0, [%1];" : "=r"(hi[p]) : "r"(src + (uint32_t)p * 128u));
#pragma unroll
      for (int e = 0; e < 32; e += 2) lo_part2(hi[e], hi[e + 1], lo[e], lo[e + 1]);
      __syncwarp();
      if (elect_one()) mbar_arrive(afree_bar(s));
      if (nuse >= 1) {
        mbar_wait(tfree_bar(group), (nuse - 1u) & 1u);
        tc_fence_after();
      }
      tmem_st16(tcol, *reinterpret_cast<const uint32_t(*)[16]>(&hi[0]));
      tmem_st16(tcol + 16u, *reinterpret_cast<const uint32_t(*)[16]>(&hi[16]));
      if (!(a.flags & 0x800)) {
        tmem_st16(tcol + 32u, *reinterpret_cast<const uint32_t(*)[16]>(&lo[0]));
        tmem_st16(tcol + 48u, *reinterpret_cast<const uint32_t(*)[16]>(&lo[16]));
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (elect_one()) mbar_arrive(tfull_bar(group));
    }
    // ---- epilogue: warp (q, group) owns k rows 32q..32q+31 and the 16-column chunks group, group+2, ... ----
    mbar_wait(acc_bar, 0);
    tc_fence_after();
    const int k = k0 + q * 32 + lane;
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)C::ACC0;
    const int nmain = nst < C::NMAIN ? nst : C::NMAIN;
    const bool single = (a.flags & 0x800) != 0;
    const int nacc = single ? nmain : (C::PAIR ? 2 * nmain : nmain + 1);
#pragma unroll 1
    for (int c = group * 16; c < BN; c += 32) {
      float acc[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) acc[e] = 0.f;
      for (int g = 0; g < nacc; g += 2) {
        uint32_t v[2][16];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int i = g + u;
          if (i < nacc) {
            // PAIR: slot i/2 = [main | corr]; else mains at (1 + i) BN, the correction accumulator (column 0) last
            const int col = C::PAIR ? (single ? 2 * i * BN : i * BN)
                                    : (single ? (1 + i) * BN : (i == nacc - 1 ? 0 : (1 + i) * BN));
            tmem_ld16_nowait(trow + (uint32_t)(col + c), v[u]);
          }
        }
        tmem_wait_ld();
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          if (g + u < nacc) {
#pragma unroll
            for (int e = 0; e < 16; ++e) acc[e] += __uint_as_float(v[u][e]);
          }
        }
      }
      if (k < a.K && nst > 0) {
#pragma unroll
        for (int e = 0; e < 16; ++e)
          if (n0 + c + e < a.N) atomicAdd(a.dw + (long)(n0 + c + e) * a.K + k, acc[e]);
      }
    }
  } else if (warp < W2_MMA_WARP) {
    // ======================= dY_lo = dY - tf32(dY) in shared memory =======================
    const int t = tid - W2_BSPLIT0 * 32;
    constexpr int NV = C::B_BYTES / 16 / (W2_BSPLITW * 32);          // float4 per thread per stage
    for (int it = 0; it < nst; ++it) {
      const int s = it % C::STAGES;
      mbar_wait(bland_bar(s), (it / C::STAGES) & 1);
      if (!(a.flags & 0x800)) {
        // (dY rows past `pend` meet all-zero Xg rows, rows past M are zero-filled by the TMA unit)
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const uint32_t so = (uint32_t)(t + W2_BSPLITW * 32 * i) * 16u;
          float4 v;
          asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                       : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                       : "r"(b_raw(s) + so));
          asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(b_lo(s) + so), "f"(lo_part(v.x)),
                       "f"(lo_part(v.y)), "f"(lo_part(v.z)), "f"(lo_part(v.w))
                       : "memory");
        }
        fence_async_proxy();
      }
      __syncwarp();
      if (elect_one()) mbar_arrive(bready_bar(s));
    }
  } else if (warp == W2_TMA_WARP) {
    // ======================= dY tiles by TMA =======================
    for (int it = 0; it < nst; ++it) {
      const int s = it % C::STAGES;
      if (it >= C::STAGES) mbar_wait(bfree_bar(s), ((it / C::STAGES) - 1) & 1);
      if (elect_one()) {
        const int p0 = pbeg + it * W2_BP;
        mbar_expect_tx(bland_bar(s), C::B_BYTES);
#pragma unroll
        for (int nb = 0; nb < C::NB; ++nb)
          tma_load_2d(b_raw(s) + nb * W2_BLK, &tm_dy, n0 + 32 * nb, p0, bland_bar(s));
      }
      __syncwarp();
    }
  } else {
    // ======================= MMA issuers: issuer `me` takes the stages it % 2 == me (TMEM slot me) =======================
    const int me = warp - W2_MMA_WARP;
    // A from tensor memory (lane = k, column = pixel), B MN-major (pixel-strided) from shared memory
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 16) | ((uint32_t)(BN >> 3) << 17) |
                           ((uint32_t)(128 >> 4) << 24);
    const uint32_t idesc2 = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 16) | ((uint32_t)((2 * BN) >> 3) << 17) |
                            ((uint32_t)(128 >> 4) << 24);                 // N = 2 BN: [dY ; dY_lo]
    const uint32_t acc0 = tmem_base + (uint32_t)C::ACC0;
    const uint32_t ta = tmem_base + (uint32_t)(me * 64), tal = ta + 32u;
    uint32_t nuse = 0;
    for (int it = me; it < nst; it += 2, ++nuse) {
      const int s = it % C::STAGES;
      mbar_wait(bready_bar(s), (it / C::STAGES) & 1);
      mbar_wait(tfull_bar(me), nuse & 1u);
      if (it > 0) mbar_wait(turn_bar(me), (uint32_t)((it - 1) >> 1) & 1u);   // the other issuer has issued stage it-1
      tc_fence_after();
      if (elect_one()) {
        const uint64_t db = make_desc_mn(b_raw(s), W2_BLK), dbl = make_desc_mn(b_lo(s), W2_BLK);
        const int slot = it % C::NMAIN;
        if (a.flags & 0x800) {
          const uint32_t d_main = acc0 + (uint32_t)(C::PAIR ? slot * 2 * BN : (1 + slot) * BN);
#pragma unroll
          for (int k = 0; k < W2_BP / 8; ++k)
            umma_tf32_ts(d_main, ta + 8u * k, db + (uint64_t)(k * 1024 >> 4), idesc, (it >= C::NMAIN) || (k != 0));
        } else if (C::PAIR) {
          const uint32_t d_pair = acc0 + (uint32_t)(slot * 2 * BN);
#pragma unroll
          for (int k = 0; k < W2_BP / 8; ++k) {
            const uint64_t adv = (uint64_t)(k * 1024 >> 4);        // next 8-pixel row group
            umma_tf32_ts(d_pair, ta + 8u * k, db + adv, idesc2, (it >= C::NMAIN) || (k != 0));   // [Xg dY | Xg dY_lo]
            umma_tf32_ts(d_pair + (uint32_t)BN, tal + 8u * k, db + adv, idesc, 1);               // += Xg_lo dY
          }
        } else {
          const uint32_t d_main = acc0 + (uint32_t)((1 + slot) * BN);
#pragma unroll
          for (int k = 0; k < W2_BP / 8; ++k) {
            const uint64_t adv = (uint64_t)(k * 1024 >> 4);
            umma_tf32_ts(acc0, tal + 8u * k, db + adv, idesc, (it | k) != 0);
            umma_tf32_ts(acc0, ta + 8u * k, dbl + adv, idesc, 1);
            umma_tf32_ts(d_main, ta + 8u * k, db + adv, idesc, (it >= C::NMAIN) || (k != 0));
          }
        }
        if (a.flags & 0x1000) tc_fence_before();      // FD_TC_FENCE=1: ordering experiment
        mbar_arrive(turn_bar(me ^ 1));
        umma_commit(bfree_bar(s));
        umma_commit(tfree_bar(me));
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(acc_bar);
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == W2_MMA_WARP) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "n"(C::TMEM_COLS)
                 : "memory");
  }
}

template <int BN>
int launch_wgrad2(const WgTcArgs& a0, const float* dy, cudaStream_t st) {
  using C = W2Cfg<BN>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_wgrad2_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         C::SMEM);
    if (e != cudaSuccess) {
      fd::set_error("conv_wgrad2: cannot reserve %d B of shared memory: %s", C::SMEM, cudaGetErrorString(e));
      return 1;
    }
    configured = true;
  }
  WgTcArgs a = a0;
  CUtensorMap tdy;
  int rc = make_map_2d(&tdy, dy, a.M, a.N, W2_BP, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
  if (rc) return rc;
  const int tiles = fd::cdiv(a.K, 128) * (a.N / BN);
  // pixel-range splits as in conv_tc.cu: half a wave of longer CTAs (FD_WGRAD_WAVES, multiples of 0.5)
  static int waves_x2 = -1;
  if (waves_x2 < 0) {
    const char* e = getenv("FD_WGRAD_WAVES");
    waves_x2 = e ? (int)(2.0 * atof(e) + 0.5) : 1;
    if (waves_x2 < 1) waves_x2 = 1;
  }
  int splits = (waves_x2 * 74 + tiles - 1) / tiles;
  const int max_splits = fd::cdiv(a.M, 4 * W2_BP);
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  a.p_per_split = fd::cdiv(fd::cdiv(a.M, splits), W2_BP) * W2_BP;
  splits = fd::cdiv(a.M, a.p_per_split);
  dim3 grid(fd::cdiv(a.K, 128), a.N / BN, splits);
  conv_wgrad2_kernel<BN><<<grid, W2_NTHREADS, C::SMEM, st>>>(a, tdy);
  FD_CHECK_LAUNCH();
  return 0;
}

}  // namespace

namespace fd {
int conv_wgrad2_dispatch(const WgTcArgs& a, const float* dy, cudaStream_t st) {
  if (a.dbg) return -1;
  if (a.N % 128 == 0) return launch_wgrad2<128>(a, dy, st);
  if (a.N % 64 == 0) return launch_wgrad2<64>(a, dy, st);
  return launch_wgrad2<32>(a, dy, st);
}
}  // namespace fd
