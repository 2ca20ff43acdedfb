// Implicit-GEMM convolution (forward / data gradient) on tcgen05 with the A operand in TENSOR MEMORY.
//
// Same contract and numerics as conv_tc.cu (3xTF32: corr += A_lo*B + A*B_lo, main_i += A*B), but the
// im2col tile never goes back to shared memory after the split:
//
//   warps 0-3   loaders    cp.async zero-fill gather of the [128 px x 32 ch] im2col tile into a
//                          SWIZZLE_128B shared stage; one thread fetches the W / W_lo tiles by TMA;
//   warps 4-11  splitters  thread = (tile row, 16-channel half): 4 conflict-free LDS.128 of its own
//                          row, A_lo = A - tf32(A) in registers, tcgen05.st of A and A_lo into a small
//                          TMEM ring (64 columns per slot); afterwards the epilogue;
//   warp 12     MMA        tcgen05.mma kind::tf32 with A from TMEM, B from shared memory.
//
// Why: with both operands in shared memory an M=128, N=128 tf32 MMA reads 8 KB per 64 cycles -- the
// whole 128 B/clk of shared-memory bandwidth -- while cp.async, TMA and the splitter compete for the
// same port (ncu: tensor pipe ~1/3 busy).  With A in TMEM the MMAs read only B (4 KB per 64 clk), the
// A_lo tile is never written to shared memory, a stage shrinks from 2A+2B to A+2B (one more stage in
// flight at N=128, two more at N=64) and the shared A tile is released as soon as it has been read.
//
// TMEM columns: [0, 64*TST) A ring (slot t: A at 64t, A_lo at 64t+32), then the correction
// accumulator and NMAIN round-robin main accumulators of BN columns each.
#include "tc_common.cuh"

namespace {

constexpr int NLOADW2 = 4, NLOAD2 = NLOADW2 * 32;
constexpr int NSPLITW2 = 8, NSPLIT2 = NSPLITW2 * 32;
constexpr int NTHREADS2 = NLOAD2 + NSPLIT2 + 96;       // + two MMA issuer warps + the weight-tile (TMA) warp
constexpr int MMA_WARP2 = NLOADW2 + NSPLITW2;
constexpr int TMA_WARP2 = MMA_WARP2 + 1;
constexpr int LOOKAHEAD = 2;                     // cp.async groups a loader keeps in flight (< STAGES)

// Timing experiment (FD_TC2_FLAGS bit 7): CTA (0,0) records clock64() at the hand-off points of one
// thread per role into trace[role][k-block][4]; see tests/trace_conv.py.
constexpr int TRACE_KB = 256;
__device__ long long* g_trace = nullptr;
#define FD_TRACE(role, kb, slot)                                                              \
  do {                                                                                         \
    if (tracing && (kb) < TRACE_KB) trace_buf[((role) * TRACE_KB + (kb)) * 4 + (slot)] = clock64(); \
  } while (0)

template <int BN>
struct Cfg2 {
  static constexpr int B_TILE = BN * 128;
  static constexpr int STAGE = A_TILE + 2 * B_TILE;
  static constexpr int STAGES = BN == 128 ? 4 : (BN == 64 ? 6 : 8);
  static constexpr int TST = BN == 128 ? 2 : (BN == 64 ? 3 : 4);      // TMEM A-ring slots
  static constexpr int ACC0 = TST * 64;                                 // first accumulator column
  // PAIR (BN <= 64): W and W_lo sit back to back in the stage, so ONE N = 2*BN MMA computes A*W (main)
  // and A*W_lo (correction) into a [main | corr2] accumulator pair: 8 MMAs per k-block instead of 12
  // (the issuing thread, ~40 cycles per tcgen05.mma, paces the narrow tiles).  A_lo*W keeps its own
  // accumulator.  BN = 128 has no TMEM for pairs and is MMA-execution bound anyway.
  static constexpr bool PAIR = BN <= 64;
  static constexpr int NMAIN_ = PAIR ? (512 - ACC0 - BN) / (2 * BN) : (512 - ACC0) / BN - 1;
  static constexpr int NMAIN = NMAIN_ > 7 ? 7 : NMAIN_;
  static constexpr int SMEM = STAGES * STAGE + 1024 /*align*/ + 512 /*barriers*/ + 1024 /*row table*/ + 1024 /*CTA channel sums*/;
};

template <int BN, int MODE>
__global__ void __launch_bounds__(NTHREADS2, 1)
conv_tc2_kernel(TcArgs a, const __grid_constant__ CUtensorMap tm_w, const __grid_constant__ CUtensorMap tm_wlo,
                const __grid_constant__ CUtensorMap tm_y) {
  using C = Cfg2<BN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = base + C::STAGES * C::STAGE;
  auto landed_bar = [&](int s) { return bars + 8u * s; };                    // smem stage filled
  auto sfree_bar = [&](int s) { return bars + 8u * (C::STAGES + s); };       // smem stage reusable
  auto tfull_bar = [&](int t) { return bars + 8u * (2 * C::STAGES + t); };   // TMEM slot written
  auto tfree_bar = [&](int t) { return bars + 8u * (2 * C::STAGES + C::TST + t); };
  const uint32_t acc_bar = bars + 8u * (2 * C::STAGES + 2 * C::TST);
  const uint32_t tmem_slot = acc_bar + 8u;
  auto turn_bar = [&](int i) { return acc_bar + 16u + 8u * i; };   // issuer i may issue its next k-block
  auto a_smem = [&](int s) { return base + s * C::STAGE; };
  auto b_raw = [&](int s) { return base + s * C::STAGE + A_TILE; };
  auto b_lo = [&](int s) { return base + s * C::STAGE + A_TILE + C::B_TILE; };

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long m0 = (long)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int nk = a.K / BK;
  long long* const trace_buf = (a.flags & 128) ? g_trace : nullptr;     // read once: stamps must stay cheap
  const bool ktrace = trace_buf && blockIdx.x == 0 && blockIdx.y == 0;
  if (ktrace && tid == NLOAD2) trace_buf[3 * TRACE_KB * 4 + 0] = clock64();

  // Row table: tile row r -> (pointer to the first gathered channel of its pixel, valid-tap mask).
  // One row per loader thread, published through shared memory by the prologue barrier -- not eight
  // rows of integer divisions per thread in front of the first load (measured: 6.7k cycles per CTA).
  const uint32_t tab_ptr = bars + 512u, tab_mask = tab_ptr + 4u * BM;
  const uint32_t cta_sums = tab_mask + 4u * BM;            // [2][BN] floats, only with a.stats
  if (a.stats && tid >= BM && tid < BM + 2 * BN)
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(cta_sums + 4u * (uint32_t)(tid - BM)), "f"(0.f) : "memory");
  if (tid < BM) {
    const long m = m0 + tid;
    int eoff = 0;                      // element offset of the pixel's first gathered channel (may be < 0)
    uint32_t mask = 0;
    if (m < a.M) {
      const uint32_t HoWo = (uint32_t)(a.Ho * a.Wo);
      const uint32_t mu = (uint32_t)m;                       // M < 2^31 (checked on the host)
      const uint32_t b = mu / HoWo, r = mu - b * HoWo;
      const int ho = (int)(r / (uint32_t)a.Wo), wo = (int)(r - (uint32_t)ho * (uint32_t)a.Wo);
      int hq, wq;
      if (MODE == 0) {
        const int hb = ho * a.stride - a.pad, wb = wo * a.stride - a.pad;
        hq = hb; wq = wb;
        for (int t = 0; t < a.KH; ++t) mask |= (uint32_t)(hb + t >= 0 && hb + t < a.Hg) << t;
        for (int t = 0; t < a.KW; ++t) mask |= (uint32_t)(wb + t >= 0 && wb + t < a.Wg) << (8 + t);
      } else {
        const int hb = ho + a.pad, wb = wo + a.pad;          // >= 0
        const int sh = a.stride == 1 ? 0 : (a.stride == 2 ? 1 : -1);
        hq = sh >= 0 ? hb >> sh : hb / a.stride;
        wq = sh >= 0 ? wb >> sh : wb / a.stride;
        for (int t = 0; t < a.KH; ++t) {
          const int th = hb - t;
          const bool al = sh >= 0 ? (th & (a.stride - 1)) == 0 : th % a.stride == 0;
          const int tq = sh >= 0 ? th >> sh : th / a.stride;
          mask |= (uint32_t)(th >= 0 && al && tq < a.Hg) << t;
        }
        for (int t = 0; t < a.KW; ++t) {
          const int tw = wb - t;
          const bool al = sh >= 0 ? (tw & (a.stride - 1)) == 0 : tw % a.stride == 0;
          const int tq = sh >= 0 ? tw >> sh : tw / a.stride;
          mask |= (uint32_t)(tw >= 0 && al && tq < a.Wg) << (8 + t);
        }
      }
      eoff = (int)((((long)b * a.Hg + hq) * a.Wg + wq) * a.Cg);      // |x| < 2^31 elements (host check)
    }
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(tab_ptr + 4u * tid), "r"(eoff) : "memory");
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(tab_mask + 4u * tid), "r"(mask) : "memory");
  }

  if (tid == 0) {
    for (int s = 0; s < C::STAGES; ++s) {
      // One elected arrival per warp: 32 lanes arriving on the same mbarrier serialise, and with
      // 256 arrivals per k-block that hand-off latency (not bandwidth) paced the whole pipeline.
      // landed: 1 expect_tx arrive + either one arrival per loader warp (cp.async groups, flags bit 0)
      // or one asynchronous cp.async.mbarrier arrival per loader thread
      mbar_init(landed_bar(s), ((a.flags & 1) ? NLOADW2 : NLOAD2) + 1);
      mbar_init(sfree_bar(s), NSPLITW2 / 2 + 1);   // the 4 splitter warps of this k-block + the MMAs' commit
    }
    for (int t = 0; t < C::TST; ++t) {
      mbar_init(tfull_bar(t), NSPLITW2 / 2);
      mbar_init(tfree_bar(t), 1);
    }
    mbar_init(acc_bar, 2);                       // both MMA issuers commit to it
    mbar_init(turn_bar(0), 1);
    mbar_init(turn_bar(1), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_w) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_wlo) : "memory");
  }
  if (warp == MMA_WARP2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  if (ktrace && tid == NLOAD2) g_trace[3 * TRACE_KB * 4 + 1] = clock64();

  if (warp < NLOADW2) {
    // ======================= loaders =======================
    constexpr int RSTEP = NLOAD2 / 8;               // 16 rows per pass
    constexpr int ROWS = BM / RSTEP;                // 8 rows per thread
    const int j = tid & 7, rg = tid >> 3;
    int ro[ROWS];                                   // 32-bit element offsets: one IMAD.WIDE per copy
    uint32_t vm[ROWS];
#pragma unroll
    for (int i = 0; i < ROWS; ++i) {
      const uint32_t row = (uint32_t)(rg + RSTEP * i);
      asm volatile("ld.shared.u32 %0, [%1];" : "=r"(ro[i]) : "r"(tab_ptr + 4u * row));
      asm volatile("ld.shared.u32 %0, [%1];" : "=r"(vm[i]) : "r"(tab_mask + 4u * row));
      ro[i] += j * 4;
    }
    // rows rg + 16 i share (row & 7) == (rg & 7)
    const uint32_t soff = (uint32_t)rg * 128u + (uint32_t)((j ^ (rg & 7)) << 4);
    int kh = 0, kw = 0, c0 = 0;
    const bool groups = (a.flags & 1) != 0;
    const int look = ((a.flags >> 8) & 3) ? ((a.flags >> 8) & 3) : LOOKAHEAD;   // experiment: flags bits 8-9
    const bool tracing = trace_buf && blockIdx.x == 0 && blockIdx.y == 0 && tid == 0;
    for (int it = 0; it < nk; ++it) {
      const int s = it % C::STAGES;
      FD_TRACE(0, it, 0);
      if (groups && it >= look) {                // the group issued `look` k-blocks ago has landed
        if (look == 1) cp_async_wait<0>();
        else if (look == 2) cp_async_wait<1>();
        else cp_async_wait<2>();
        __syncwarp();
        if (elect_one()) mbar_arrive(landed_bar((it - look) % C::STAGES));
      }
      FD_TRACE(0, it, 1);
      if (it >= C::STAGES) mbar_wait(sfree_bar(s), ((it / C::STAGES) - 1) & 1);
      FD_TRACE(0, it, 2);
      const int toff = MODE == 0 ? (kh * a.Wg + kw) * a.Cg + c0
                                 : -((kh / a.stride) * a.Wg + kw / a.stride) * a.Cg + c0;
      const uint32_t dst = a_smem(s) + soff;
      if (!(a.flags & 8)) {
#pragma unroll
        for (int i = 0; i < ROWS; ++i) {
          const bool ok = ((vm[i] >> kh) & (vm[i] >> (8 + kw)) & 1u) != 0;
          cp_async16(dst + (uint32_t)i * RSTEP * 128u, a.x + (ok ? ro[i] + toff : 0), ok ? 16u : 0u);
        }
      }
      if (groups) cp_async_commit();
      else cp_async_arrive_noinc(landed_bar(s));
      FD_TRACE(0, it, 3);
      c0 += BK;
      if (c0 == a.Cg) { c0 = 0; if (++kw == a.KW) { kw = 0; ++kh; } }
    }
    if (groups) {                                // drain: the last min(nk, LOOKAHEAD) groups
      cp_async_wait<0>();
      __syncwarp();
      if (lane == 0)
        for (int it = nk > look ? nk - look : 0; it < nk; ++it) mbar_arrive(landed_bar(it % C::STAGES));
    }
  } else if (warp < MMA_WARP2) {
    // ======================= splitters, then epilogue =======================
    // Two groups of four warps take alternate k-blocks (thread = one whole tile row of 32 channels):
    // the barrier waits, fences and TMEM-store round trip of a k-block cost ~900 cycles of latency
    // per warp whatever the payload, so each group gets two k-block periods to hide them.
    const int q = warp & 3;                       // TMEM lane quadrant this warp may access
    const int half = (warp - NLOADW2) >> 2;       // splitter group (k-block parity); column half in the epilogue
    const int row = q * 32 + lane;
    const uint32_t row_off = (uint32_t)row * 128u;
    const uint32_t sw = (uint32_t)(row & 7);
    const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16);
    const bool tracing = trace_buf && blockIdx.x == 0 && blockIdx.y == 0 && tid == NLOAD2;
    for (int kb = half; kb < nk; kb += 2) {
      const int s = kb % C::STAGES, t = kb % C::TST;
      FD_TRACE(1, kb, 0);
      mbar_wait(landed_bar(s), (kb / C::STAGES) & 1);
      FD_TRACE(1, kb, 1);
      uint32_t hi[32], lo[32];
      const uint32_t ar = a_smem(s) + row_off;
      const bool skip_split = (a.flags & 4) != 0;      // timing experiment: hand-offs only
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        if (skip_split) { hi[4 * jj] = hi[4 * jj + 1] = hi[4 * jj + 2] = hi[4 * jj + 3] = 0u; continue; }
        asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];"
                     : "=r"(hi[4 * jj]), "=r"(hi[4 * jj + 1]), "=r"(hi[4 * jj + 2]), "=r"(hi[4 * jj + 3])
                     : "r"(ar + (((uint32_t)jj ^ sw) << 4)));
      }
#pragma unroll
      for (int e = 0; e < 32; e += 2) lo_part2(hi[e], hi[e + 1], lo[e], lo[e + 1]);
      // the warp is converged: once the low parts are computed every lane's loads have returned
      __syncwarp();
      if (elect_one()) mbar_arrive(sfree_bar(s));
      if (kb >= C::TST) {
        mbar_wait(tfree_bar(t), ((kb / C::TST) - 1) & 1);
        tc_fence_after();
      }
      FD_TRACE(1, kb, 2);
      const uint32_t tcol = tlane + (uint32_t)(t * 64);
      if (!skip_split) {
        tmem_st16(tcol, *reinterpret_cast<const uint32_t(*)[16]>(&hi[0]));
        tmem_st16(tcol + 16u, *reinterpret_cast<const uint32_t(*)[16]>(&hi[16]));
        tmem_st16(tcol + 32u, *reinterpret_cast<const uint32_t(*)[16]>(&lo[0]));
        tmem_st16(tcol + 48u, *reinterpret_cast<const uint32_t(*)[16]>(&lo[16]));
        tmem_wait_st();
      }
      tc_fence_before();
      __syncwarp();
      if (elect_one()) mbar_arrive(tfull_bar(t));
      FD_TRACE(1, kb, 3);
    }
    // ---- epilogue: warp (q, half) owns rows 32q..32q+31 and the 16-column chunks half, half+2, ... ----
    mbar_wait(acc_bar, 0);
    tc_fence_after();
    if (ktrace && tid == NLOAD2) g_trace[3 * TRACE_KB * 4 + 2] = clock64();
    const long m = m0 + row;
    const uint32_t trow = tlane + (uint32_t)C::ACC0;
    const int nmain = nk < C::NMAIN ? nk : C::NMAIN;
#pragma unroll 1
    for (int c = half * 16; c < BN; c += 32) {
      // accumulators in summation order: main 0..nmain-1, then the (small) correction terms
      float acc[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) acc[e] = 0.f;
      // list of accumulator column bases: PAIR: main_0, corr2_0, main_1, ... then corr1; else main_i, corr.
      // Single-pass TF32 (flags bit 11, FD_CONV_PRECISION=tf32): only the main accumulators exist.
      const bool single = (a.flags & 0x800) != 0;
      const int nacc = single ? nmain : (C::PAIR ? 2 * nmain + 1 : nmain + 1);
      for (int g = 0; g < nacc; g += 4) {
        uint32_t v[4][16];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int i = g + u;
          if (i < nacc) {
            const int col = single ? (C::PAIR ? BN + 2 * i * BN : (1 + i) * BN)
                                   : (i == nacc - 1 ? 0 : (C::PAIR ? BN + i * BN : (1 + i) * BN));
            if (!(a.flags & 0x400)) tmem_ld16_nowait(trow + (uint32_t)(col + c), v[u]);   // 0x400: timing experiment
          }
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
      if (a.stats) {
        // Per-channel sum / sum of squares of this warp's 32 rows (rows past M are exact zeros):
        // transpose-reduce, 16 shuffles per statistic; even lanes end up owning one channel each.
        float s1[16], s2[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) { s1[e] = acc[e]; s2[e] = acc[e] * acc[e]; }
#pragma unroll
        for (int width = 8, off = 16; width >= 1; width >>= 1, off >>= 1) {
          const bool up = (lane & off) != 0;
#pragma unroll
          for (int i = 0; i < width; ++i) {
            const float send1 = up ? s1[i] : s1[i + width], send2 = up ? s2[i] : s2[i + width];
            const float keep1 = up ? s1[i + width] : s1[i], keep2 = up ? s2[i + width] : s2[i];
            s1[i] = keep1 + __shfl_xor_sync(0xffffffffu, send1, off);
            s2[i] = keep2 + __shfl_xor_sync(0xffffffffu, send2, off);
          }
        }
        s1[0] += __shfl_xor_sync(0xffffffffu, s1[0], 1);
        s2[0] += __shfl_xor_sync(0xffffffffu, s2[0], 1);
        if (!(lane & 1)) {
          const int ch = c + ((lane & 16) ? 8 : 0) + ((lane & 8) ? 4 : 0) + ((lane & 4) ? 2 : 0) + ((lane & 2) ? 1 : 0);
          asm volatile("red.shared.add.f32 [%0], %1;" ::"r"(cta_sums + 4u * (uint32_t)ch), "f"(s1[0]) : "memory");
          asm volatile("red.shared.add.f32 [%0], %1;" ::"r"(cta_sums + 4u * (uint32_t)(BN + ch)), "f"(s2[0]) : "memory");
        }
      }
      if (BN >= 32) {
        // Stage the tile in shared memory (the pipeline stages are all consumed by now) and let TMA write it:
        // one row per lane, the direct float4 stores reach global memory as 32 scattered 16-byte pieces per
        // instruction (rows are N*4 bytes apart) -- measured 23 % of the layer-1 kernel.
        float o[16];
        bias_act16(acc, a.bias ? a.bias + n0 + c : nullptr, a.act, o);
        stage_out16(base, row, c, o);
      } else if (m < a.M && n0 + c < a.N) {
        float o[16];
        bias_act16(acc, a.bias ? a.bias + n0 + c : nullptr, a.act, o);
        float4* dst = reinterpret_cast<float4*>(a.y + m * a.N + n0 + c);
#pragma unroll
        for (int e = 0; e < 4; ++e) dst[e] = make_float4(o[4 * e], o[4 * e + 1], o[4 * e + 2], o[4 * e + 3]);
      }
    }
    if (BN >= 32) {
      fence_async_proxy();                                    // generic-proxy writes -> visible to the TMA unit
      asm volatile("bar.sync 2, %0;" ::"n"(NSPLIT2) : "memory");
      if (warp == NLOADW2 && elect_one() && !(a.flags & 0x200)) {
#pragma unroll
        for (int blk = 0; blk < BN / 32; ++blk)
          if (n0 + 32 * blk < a.N) tma_store_2d(&tm_y, base + (uint32_t)blk * (128u * 128u), n0 + 32 * blk, (int)m0);
        tma_store_commit_wait();                              // shared memory is read before the CTA retires
      }
    }
    if (a.stats) {
      // the eight epilogue warps meet, then 2*BN threads fold the CTA's sums into the fp64 totals
      asm volatile("bar.sync 1, %0;" ::"n"(NSPLIT2) : "memory");
      const int i = tid - NLOAD2;
      if (i < 2 * BN) {
        float v;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(cta_sums + 4u * (uint32_t)i));
        const int stat = i / BN, ch = n0 + (i - stat * BN);
        if (ch < a.N) atomicAdd(a.stats + (long)stat * a.N + ch, (double)v);
      }
    }
  } else if (warp == TMA_WARP2) {
    // ======================= weight tiles: W and W_lo by TMA =======================
    for (int it = 0; it < nk; ++it) {
      const int s = it % C::STAGES;
      if (it >= C::STAGES) mbar_wait(sfree_bar(s), ((it / C::STAGES) - 1) & 1);
      if (elect_one()) {
        if (a.flags & 8) {                       // timing experiment: no loads at all
          mbar_arrive(landed_bar(s));
        } else {
          mbar_expect_tx(landed_bar(s), 2 * C::B_TILE);
          tma_load_2d(b_raw(s), &tm_w, it * BK, n0, landed_bar(s));
          tma_load_2d(b_lo(s), &tm_wlo, it * BK, n0, landed_bar(s));
        }
      }
      __syncwarp();
    }
  } else {
    // ======================= MMA issuers =======================
    // The whole warp walks the loop and waits on the barriers (converged); one lane issues.  A lone
    // lane looping while 31 lanes sit at a convergence barrier paid for it in hand-off latency.
    // Two issuer warps take alternate k-blocks (see conv_tc3.cu): tcgen05.mma issue blocks until the pipe
    // accepts it, so one issuer's barrier wait + fence + commits (~250 clk) left the pipe idle every k-block;
    // a `turn` mbarrier keeps the issue order, and with it the summation order, fixed.
    const int me = warp == MMA_WARP2 ? 0 : 1;
    const bool plain = (a.flags & 64) != 0 && (a.flags & 16) != 0;   // experiment: arrive without commit
    const bool tracing = trace_buf && blockIdx.x == 0 && blockIdx.y == 0 && lane == 0;
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) |
                           ((uint32_t)(BM >> 4) << 24);
    const uint32_t idesc2 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)((2 * BN) >> 3) << 17) |
                            ((uint32_t)(BM >> 4) << 24);                    // N = 2*BN: [W ; W_lo]
    const uint32_t d_corr = tmem_base + (uint32_t)C::ACC0;
    for (int kb = me; kb < nk; kb += 2) {
      const int s = kb % C::STAGES, t = kb % C::TST;
      FD_TRACE(2, kb, 0);
      mbar_wait(tfull_bar(t), (kb / C::TST) & 1);
      if (kb > 0) mbar_wait(turn_bar(me), ((kb - 1) >> 1) & 1);   // the other issuer has issued k-block kb-1
      FD_TRACE(2, kb, 1);
      tc_fence_after();
      FD_TRACE(2, kb, 2);
      if (elect_one()) {
        const uint64_t db = make_desc(b_raw(s)), dbl = make_desc(b_lo(s));
        const uint32_t ta = tmem_base + (uint32_t)(t * 64), tal = ta + 32u;
        if (a.flags & 0x800) {
          // single-pass TF32: D += tf32(A) * tf32(W), nothing else (what cuDNN computes with allow_tf32)
          const uint32_t d_main = d_corr + (uint32_t)(C::PAIR ? BN + (kb % C::NMAIN) * 2 * BN : (1 + kb % C::NMAIN) * BN);
#pragma unroll
          for (int k = 0; k < BK / 8; ++k)
            umma_tf32_ts(d_main, ta + 8u * k, db + (uint64_t)(k * 32 >> 4), idesc, (kb >= C::NMAIN) || (k != 0));
        } else if (C::PAIR) {
          const uint32_t d_pair = d_corr + (uint32_t)(BN + (kb % C::NMAIN) * 2 * BN);
#pragma unroll
          for (int k = 0; k < BK / 8; ++k) {
            const uint64_t adv = (uint64_t)(k * 32 >> 4);
            if (!(a.flags & (2 | 16))) umma_tf32_ts(d_corr, tal + 8u * k, db + adv, idesc, (kb | k) != 0);
            if (!(a.flags & 16)) umma_tf32_ts(d_pair, ta + 8u * k, db + adv, idesc2, (kb >= C::NMAIN) || (k != 0));
          }
        } else {
          const uint32_t d_main = d_corr + (uint32_t)((1 + kb % C::NMAIN) * BN);
#pragma unroll
          for (int k = 0; k < BK / 8; ++k) {
            const uint64_t adv = (uint64_t)(k * 32 >> 4);
            if (!(a.flags & (2 | 16))) {
              umma_tf32_ts(d_corr, tal + 8u * k, db + adv, idesc, (kb | k) != 0);
              umma_tf32_ts(d_corr, ta + 8u * k, dbl + adv, idesc, 1);
            }
            if (!(a.flags & 16)) umma_tf32_ts(d_main, ta + 8u * k, db + adv, idesc, (kb >= C::NMAIN) || (k != 0));
          }
        }
        if (a.flags & 0x1000) tc_fence_before();      // FD_TC_FENCE=1: ordering experiment
        mbar_arrive(turn_bar(me ^ 1));
        if (plain) {
          mbar_arrive(sfree_bar(s));
          mbar_arrive(tfree_bar(t));
        } else {
          umma_commit(sfree_bar(s));
          umma_commit(tfree_bar(t));
        }
      }
      __syncwarp();
      FD_TRACE(2, kb, 3);
    }
    if (elect_one()) umma_commit(acc_bar);
    __syncwarp();
  }
  if (ktrace && tid == NLOAD2) g_trace[3 * TRACE_KB * 4 + 3] = clock64();
  tc_fence_before();
  __syncthreads();
  if (ktrace && tid == NLOAD2) g_trace[3 * TRACE_KB * 4 + 4] = clock64();
  if (warp == MMA_WARP2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
  }
}

template <int BN, int MODE>
int launch_tc2(const TcArgs& a, cudaStream_t st) {
  using C = Cfg2<BN>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc2_kernel<BN, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         C::SMEM);
    if (e != cudaSuccess) {
      fd::set_error("conv_tc2: cannot reserve %d B of shared memory: %s", C::SMEM, cudaGetErrorString(e));
      return 1;
    }
    configured = true;
  }
  CUtensorMap tw, twl;
  int rc = make_map_2d(&tw, a.w, a.N, a.K, BN);
  if (rc) return rc;
  rc = make_map_2d(&twl, a.wlo, a.N, a.K, BN);
  if (rc) return rc;
  CUtensorMap ty = tw;
  if (BN >= 32) {
    rc = make_map_2d(&ty, a.y, a.M, a.N, BM);
    if (rc) return rc;
  }
  dim3 grid(fd::cdiv(a.M, BM), fd::cdiv(a.N, BN));
  conv_tc2_kernel<BN, MODE><<<grid, NTHREADS2, C::SMEM, st>>>(a, tw, twl, ty);
  FD_CHECK_LAUNCH();
  return 0;
}

template <int MODE>
int dispatch_tc2(const TcArgs& a, cudaStream_t st) {
  FD_REQUIRE((((uintptr_t)a.w | (uintptr_t)a.wlo | (uintptr_t)a.x | (uintptr_t)a.y) & 15) == 0,
             "conv_tc2: operands must be 16-byte aligned");
  FD_REQUIRE(a.M < (1L << 31) && a.KH <= 8 && a.KW <= 8, "conv_tc2: problem too large (M=%ld, %dx%d)", a.M,
             a.KH, a.KW);
  FD_REQUIRE((long)a.B * a.Hg * a.Wg * a.Cg < (1L << 31), "conv_tc2: gathered tensor has 2^31 or more elements");
  // Tile width: the widest BN dividing N.  Splitting the few-pixel N >= 128 layers (layer3 / layer4,
  // 24-46 CTAs) into 64-wide tiles lowers their latency alone but costs ~1.5x the SM-time (every extra
  // N tile repeats the A gather and split); inside the captured step, where the other streams keep the
  // SMs full, the wide tiles win (368.8 vs 351.7 images/s).  FD_TC2_NARROW=<ctas> restores the split for
  // layers with fewer CTAs than that.
  static int narrow_below = -1;
  if (narrow_below < 0) {
    const char* e = getenv("FD_TC2_NARROW");
    narrow_below = e ? atoi(e) : 0;
  }
  int bn = a.N % 128 == 0 ? 128 : (a.N % 64 == 0 ? 64 : (a.N % 32 == 0 ? 32 : 16));
  if (bn == 128 && fd::cdiv(a.M, BM) * (a.N / 128) < narrow_below) bn = 64;
  if (bn == 128) return launch_tc2<128, MODE>(a, st);
  if (bn == 64) return launch_tc2<64, MODE>(a, st);
  if (bn == 32) return launch_tc2<32, MODE>(a, st);
  return launch_tc2<16, MODE>(a, st);
}

}  // namespace

extern "C" int fd_debug_set_conv_trace(void* device_buffer) {
  // device_buffer: (3 * 256 * 4 + 8) int64 slots, or NULL to switch tracing off
  fd::g_conv_trace_host = (long long*)device_buffer;
  cudaError_t e = cudaMemcpyToSymbol(g_trace, &device_buffer, sizeof(void*));
  return e == cudaSuccess ? 0 : 1;
}

namespace fd {
long long* g_conv_trace_host = nullptr;
int conv_tc2_dispatch(const TcArgs& a, int mode, cudaStream_t st) {
  return mode == 0 ? dispatch_tc2<0>(a, st) : dispatch_tc2<1>(a, st);
}
}  // namespace fd
