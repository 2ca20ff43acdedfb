// Implicit-GEMM convolution (forward / data gradient, stride 1) on tcgen05 -- the "patch" variant of
// conv_tc2.cu.  Same contract and numerics (3xTF32, A operand in tensor memory, accumulator rotation), but
// it moves 2-3x fewer bytes from L2 to the SM:
//
//   * conv_tc2 gathers the [128 px x 32 ch] im2col tile once PER TAP (9 x 16 KB per 32-channel chunk for
//     a 3x3) and fetches W and W_lo for every k-block.  ncu: 210 MB of L2->SM traffic per layer-1 launch
//     against 23.7 MB of tensors -- at ~5 TB/s that IS the launch time: the L2 slices deliver ~6300 B/clk
//     chip-wide (B300_MICROARCH.md, "LTS throughput cap"), i.e. ~43 B/clk per SM when all 148 pull, and a
//     BN=64 k-block needs 32 KB = 770 clk of L2 time against 384 clk of MMAs.
//   * Here the input pixels a tile needs for ALL taps of a 32-channel chunk -- a contiguous raster range of
//     128 + (KH-1)*Wg + KW pixels, since consecutive output pixels read consecutive input pixels at stride 1 --
//     are fetched ONCE by TMA ([pixels x 32 ch] boxes, SWIZZLE_128B, out-of-range rows zero-filled) into a
//     double-buffered "patch"; the splitter warps read tap (kh,kw) of tile row r from patch row
//     r0[r] + kh*Wg + kw.  W_lo = W - tf32(W) is computed in shared memory by four otherwise idle warps
//     instead of being fetched.  Per k-block (BN=64, Wg=160): 6.4 KB of patch + 8 KB of W = 14.4 KB.
//
// Roles (480 threads, one CTA per SM):
//   warps 0-3   weight splitters   W tile (TMA) -> W_lo tile next to it, fence.proxy.async, `wready`
//   warps 4-11  A splitters        two groups on alternate k-blocks: patch row -> registers -> A / A_lo into the
//                                  TMEM ring (tcgen05.st); then the epilogue (shared with conv_tc2's layout)
//   warps 12,15 MMA issue          tcgen05.mma kind::tf32, A from TMEM, B = [W ; W_lo] from shared memory.  Two
//                                  issuers on alternate k-blocks: the role trace (tools/trace_conv3.py) showed
//                                  one issuer spending ~400 clk per k-block in its two mbarrier waits (~90 clk
//                                  each even when complete), fences and commits with the tensor pipe idle,
//                                  because tcgen05.mma issue blocks until the pipe accepts it (862 clk per
//                                  k-block against 384 of MMAs at BN=64, 1200 against 768 at BN=128).  The
//                                  other issuer does its waits meanwhile; a `turn` mbarrier keeps the issue
//                                  order (and with it the summation order) fixed.
//   warp 13     W tiles by TMA     one k-block per stage
//   warp 14     patches by TMA     one 32-channel chunk per buffer, a whole chunk ahead
//
// k-block order: chunk-major (c0 outer, taps inner), W tile column = tap*Cg + c0.
#include <algorithm>

#include "tc_common.cuh"

namespace {

constexpr int T3_WSPLITW = 4, T3_ASPLITW = 8;
constexpr int T3_MMA_WARP = T3_WSPLITW + T3_ASPLITW;        // 12
constexpr int T3_WTMA_WARP = T3_MMA_WARP + 1;               // 13
constexpr int T3_PTMA_WARP = T3_MMA_WARP + 2;               // 14
constexpr int T3_MMA2_WARP = T3_MMA_WARP + 3;               // 15: second MMA issuer (odd k-blocks)
constexpr int T3_NTHREADS = (T3_MMA2_WARP + 1) * 32;        // 512
constexpr int PBOX = 64;                                    // patch rows per TMA box
// Timing experiment (fd_debug_set_conv_trace + tools/trace_conv.py): CTA (0,0) stamps clock64() at the hand-off
// points of one lane per role into trace[role][k-block][4], kernel phases behind them.
constexpr int T3_TRACE_KB = 256;
#define T3_TRACE(role, kb, slot)                                                                   \
  do {                                                                                             \
    if (tracing && (kb) < T3_TRACE_KB) a.trace[((role) * T3_TRACE_KB + (kb)) * 4 + (slot)] = clock64(); \
  } while (0)

template <int BN>
struct Cfg3 {
  static constexpr int B_TILE = BN * 128;
  static constexpr int WSTAGE = 2 * B_TILE;                             // [W ; W_lo]
  // The patch buffers are sized per launch (rows the layer's tiles need, in whole TMA boxes) and the rest of
  // the 227 KB goes to W stages: the trace showed the issuers waiting for W once they no longer waited for each
  // other (stage turn-around = MMAs done -> TMA from L2 -> W_lo split, ~2000 clk, against 3 stages x ~800 clk).
  static constexpr int MAXST = 8;
  static constexpr int MISC = 1024 /*align*/ + 512 /*barriers*/ + 1024 /*row table*/ + 1024 /*CTA channel sums*/;
  static constexpr int SMEM_MAX = 232448;
  static constexpr int TST = BN == 128 ? 2 : (BN == 64 ? 3 : 4);        // TMEM A-ring slots
  static constexpr int ACC0 = TST * 64;
  static constexpr bool PAIR = BN <= 64;
  static constexpr int NMAIN_ = PAIR ? (512 - ACC0 - BN) / (2 * BN) : (512 - ACC0) / BN - 1;
  static constexpr int NMAIN = NMAIN_ > 7 ? 7 : NMAIN_;
};

template <int BN, int MODE>
__global__ void __launch_bounds__(T3_NTHREADS, 1)
conv_tc3_kernel(TcArgs a, const __grid_constant__ CUtensorMap tm_w, const __grid_constant__ CUtensorMap tm_x,
                const __grid_constant__ CUtensorMap tm_y, const int patch_rows, const int nstages) {
  using C = Cfg3<BN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t patch_bytes = (uint32_t)patch_rows * 128u;
  auto patch = [&](int b) { return base + (uint32_t)b * patch_bytes; };
  const uint32_t wbase = base + 2u * patch_bytes;
  auto b_raw = [&](int s) { return wbase + (uint32_t)s * C::WSTAGE; };
  auto b_lo = [&](int s) { return wbase + (uint32_t)s * C::WSTAGE + C::B_TILE; };
  const uint32_t bars = wbase + (uint32_t)nstages * C::WSTAGE;
  auto wland_bar = [&](int s) { return bars + 8u * s; };                        // W tile landed (TMA tx)
  auto wready_bar = [&](int s) { return bars + 8u * (C::MAXST + s); };          // W_lo written
  auto wfree_bar = [&](int s) { return bars + 8u * (2 * C::MAXST + s); };       // MMAs done with the stage
  auto tfull_bar = [&](int t) { return bars + 8u * (3 * C::MAXST + t); };
  auto tfree_bar = [&](int t) { return bars + 8u * (3 * C::MAXST + C::TST + t); };
  auto pfull_bar = [&](int b) { return bars + 8u * (3 * C::MAXST + 2 * C::TST + b); };
  auto pfree_bar = [&](int b) { return bars + 8u * (3 * C::MAXST + 2 * C::TST + 2 + b); };
  const uint32_t acc_bar = bars + 8u * (3 * C::MAXST + 2 * C::TST + 4);
  // W-stage ring position of a role: (stage, phase parity), advanced without divisions
  auto ring_next = [&](int& st, uint32_t& ph) { if (++st == nstages) { st = 0; ph ^= 1u; } };
  const uint32_t tmem_slot = acc_bar + 8u;
  auto turn_bar = [&](int i) { return acc_bar + 48u + 8u * i; };          // issuer i may issue its next k-block
  const uint32_t pinfo = acc_bar + 16u;                   // [0] patch start (pixel index), [1] patch rows needed

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // Tiles never cross an image: tile t of image b covers output pixels [t*128, t*128+128) of that image.  (A
  // tile that crossed into the next image of a pad-0 layer would need the (Hg-Ho)*Wg pixels in between too.)
  const int HoWo = a.Ho * a.Wo;
  const int tpi = (HoWo + BM - 1) / BM;
  const int img = (int)blockIdx.x / tpi, r0 = ((int)blockIdx.x - img * tpi) * BM;
  const long m0 = (long)img * HoWo + r0;
  const int rows_valid = HoWo - r0 < BM ? HoWo - r0 : BM;
  const int n0 = blockIdx.y * BN;
  const int nchunk = a.Cg / BK;
  // MODE 2 (tap table): the class of this CTA (blockIdx.z), shift of its tap t
  TcArgs::TapClass tc = a.cl[0];
  int tsh[4] = {0, 0, 0, 0};
  if (MODE == 2) {
    if (blockIdx.z == 1) tc = a.cl[1];
    if (blockIdx.z == 2) tc = a.cl[2];
    if (blockIdx.z == 3) tc = a.cl[3];
#pragma unroll
    for (int t = 0; t < 4; ++t) tsh[t] = t < tc.ntap ? tc.dh[t] * a.Wg + tc.dw[t] : 0;
  }
  const int taps = MODE == 2 ? tc.ntap : a.KH * a.KW;
  const int KWt = MODE == 2 ? tc.ntap : a.KW;               // taps per kernel row (MODE 2: one row of ntap taps)
  const int nk = taps * nchunk;
  const int span = MODE == 2 ? max(max(tsh[0], tsh[1]), max(tsh[2], tsh[3]))
                             : (a.KH - 1) * a.Wg + (a.KW - 1);        // largest tap shift
  const bool ktrace = a.trace && blockIdx.x == 0 && blockIdx.y == 0;
  if (ktrace && tid == T3_WSPLITW * 32) a.trace[3 * T3_TRACE_KB * 4 + 0] = clock64();

  // Row table: tile row r -> (pixel index of its tap-(0,0) source, valid-tap mask); per-warp min / max of the
  // pixel index in pinfo[0..3] / pinfo[4..7]
  const uint32_t tab_lin = bars + 512u, tab_mask = tab_lin + 4u * BM;
  const uint32_t cta_sums = tab_mask + 4u * BM;            // [2][BN] floats, only with a.stats
  const uint32_t zero_row = bars + 384u;                   // 128 B of zeros: the "row" a padding tap reads
  if (tid >= 256 && tid < 288)
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(zero_row + 4u * (uint32_t)(tid - 256)), "r"(0u) : "memory");
  if (a.stats && tid >= BM && tid < BM + 2 * BN)
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(cta_sums + 4u * (uint32_t)(tid - BM)), "f"(0.f) : "memory");
  if (tid < BM) {
    int lin = 0;
    uint32_t mask = 0;
    const bool valid = tid < rows_valid;
    if (valid) {
      const int r = r0 + tid;
      const int ho = r / a.Wo, wo = r - ho * a.Wo;
      int hq, wq;
      if (MODE == 2) {
        hq = ho; wq = wo;
        mask = 1u;                                          // "kh = 0" is always valid; bit 8 + t = tap t is inside dY
#pragma unroll
        for (int t = 0; t < 4; ++t)
          mask |= (uint32_t)(t < tc.ntap && ho + tc.dh[t] < a.Hg && wo + tc.dw[t] < a.Wg) << (8 + t);
      } else if (MODE == 0) {
        hq = ho - a.pad; wq = wo - a.pad;
        for (int t = 0; t < a.KH; ++t) mask |= (uint32_t)(hq + t >= 0 && hq + t < a.Hg) << t;
        for (int t = 0; t < a.KW; ++t) mask |= (uint32_t)(wq + t >= 0 && wq + t < a.Wg) << (8 + t);
      } else {
        hq = ho + a.pad; wq = wo + a.pad;
        for (int t = 0; t < a.KH; ++t) mask |= (uint32_t)(hq - t >= 0 && hq - t < a.Hg) << t;
        for (int t = 0; t < a.KW; ++t) mask |= (uint32_t)(wq - t >= 0 && wq - t < a.Wg) << (8 + t);
      }
      lin = (img * a.Hg + hq) * a.Wg + wq;
      mask |= 1u << 31;                                   // row exists (even if every tap is padding)
    }
    int lo = valid ? lin : 0x7fffffff, hi = valid ? lin : (int)0x80000000;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
      hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if (lane == 0) {
      asm volatile("st.shared.u32 [%0], %1;" ::"r"(pinfo + 4u * (uint32_t)warp), "r"(lo) : "memory");
      asm volatile("st.shared.u32 [%0], %1;" ::"r"(pinfo + 16u + 4u * (uint32_t)warp), "r"(hi) : "memory");
    }
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(tab_lin + 4u * tid), "r"(lin) : "memory");
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(tab_mask + 4u * tid), "r"(mask) : "memory");
  }

  if (tid == 0) {
    for (int s = 0; s < nstages; ++s) {
      mbar_init(wland_bar(s), 1);
      mbar_init(wready_bar(s), T3_WSPLITW);
      mbar_init(wfree_bar(s), 1);
    }
    for (int t = 0; t < C::TST; ++t) {
      mbar_init(tfull_bar(t), T3_ASPLITW / 2);
      mbar_init(tfree_bar(t), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(pfull_bar(b), 1);
      mbar_init(pfree_bar(b), taps == 1 ? T3_ASPLITW / 2 : T3_ASPLITW);   // 1x1: a chunk belongs to one group
    }
    mbar_init(acc_bar, 2);                                  // both MMA issuers commit to it
    mbar_init(turn_bar(0), 1);
    mbar_init(turn_bar(1), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    // Prologue breakdown (tools/trace_conv3.py): this serial initialisation ends at ~1400-1500 clk, the row table at
    // ~1100, the tensor-memory allocation at ~500.  Spreading the list over two warps (one barrier per thread) brought
    // the __syncthreads from ~1650 to ~1270 clk but made the persistent layer-1 kernel and the step SLOWER
    // (29.05 vs 28.4 us, 500 vs 508 images/s: the extra warps walking the list compete with the first loads), and
    // over all warps it cost 600 clk more than it saved -- measured, reverted.
    if (ktrace) a.trace[3 * T3_TRACE_KB * 4 + 5] = clock64();          // barriers initialised
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_w) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_x) : "memory");
  }
  if (warp == T3_MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    if (ktrace && lane == 0) a.trace[3 * T3_TRACE_KB * 4 + 6] = clock64();   // tensor memory allocated
  }
  if (ktrace && tid == 127) a.trace[3 * T3_TRACE_KB * 4 + 7] = clock64();    // row table written (last table thread)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (ktrace && tid == T3_WSPLITW * 32) a.trace[3 * T3_TRACE_KB * 4 + 1] = clock64();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  int lin_min = 0x7fffffff, lin_max = (int)0x80000000;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int v, w;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(pinfo + 4u * i));
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w) : "r"(pinfo + 16u + 4u * i));
    lin_min = min(lin_min, v);
    lin_max = max(lin_max, w);
  }
  // the patch: pixels [pstart, pstart + prows) of the gathered tensor
  const int pstart = MODE != 1 ? lin_min : lin_min - span;
  const int prows = lin_max - lin_min + span + 1;

  if (warp < T3_WSPLITW) {
    // ======================= weight splitters: W_lo = W - tf32(W) in shared memory =======================
    constexpr int NV = C::B_TILE / 16 / (T3_WSPLITW * 32);          // float4 per thread per stage (>= 1 for BN >= 16)
    int s = 0;
    uint32_t sph = 0;
    const bool tracing = ktrace && tid == 0;
    for (int kb = 0; kb < nk; ++kb, ring_next(s, sph)) {
      mbar_wait(wland_bar(s), sph);
      T3_TRACE(2, kb, 2);
      if (!(a.flags & 0x800)) {
#pragma unroll
        for (int i = 0; i < (NV > 0 ? NV : 1); ++i) {
          const uint32_t idx = (uint32_t)(tid + T3_WSPLITW * 32 * i);
          if (idx < (uint32_t)(C::B_TILE / 16)) {
            float4 v;
            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                         : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                         : "r"(b_raw(s) + idx * 16u));
            asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(b_lo(s) + idx * 16u), "f"(lo_part(v.x)),
                         "f"(lo_part(v.y)), "f"(lo_part(v.z)), "f"(lo_part(v.w))
                         : "memory");
          }
        }
        fence_async_proxy();
      }
      __syncwarp();
      if (elect_one()) mbar_arrive(wready_bar(s));
      T3_TRACE(2, kb, 3);
    }
  } else if (warp < T3_MMA_WARP) {
    // ======================= A splitters, then epilogue =======================
    const int q = warp & 3;                       // TMEM lane quadrant this warp may access
    const int half = (warp - T3_WSPLITW) >> 2;    // splitter group (k-block parity); column half in the epilogue
    const int row = q * 32 + lane;
    const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16);
    int lin;
    uint32_t vm;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(lin) : "r"(tab_lin + 4u * (uint32_t)row));
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(vm) : "r"(tab_mask + 4u * (uint32_t)row));
    // patch row of this tile row's tap (0,0); rows past the image stay inside the buffer and are masked to zero
    const int pr0 = (vm >> 31) ? lin - pstart : (MODE != 1 ? 0 : span);
    int chunk_seen = -1;
    const bool tracing = ktrace && lane == 0 && q == 0 && half == 0;
    const int trole = 0;
    // (chunk, kh, kw) of this group's k-block, stepped two taps at a time: no divisions in the loop (the trace
    // put load + split at ~820 clk per k-block, ~200 instructions of which two runtime divisions, 32 selects
    // for the padding mask and 64 for the low-order parts)
    int c = half / taps, kh, kw;
    {
      const int tap0 = half - c * taps;
      kh = tap0 / KWt;
      kw = tap0 - kh * KWt;
    }
    for (int kb = half; kb < nk; kb += 2) {
      T3_TRACE(trole, kb, 0);
      const int tap = kh * KWt + kw;
      const int t = kb % C::TST, pb = c & 1;
      if (c != chunk_seen) {
        mbar_wait(pfull_bar(pb), (c >> 1) & 1);
        chunk_seen = c;
      }
      const int sh = MODE == 2 ? (kw == 0 ? tsh[0] : (kw == 1 ? tsh[1] : (kw == 2 ? tsh[2] : tsh[3]))) : kh * a.Wg + kw;
      const uint32_t pr = (uint32_t)(MODE != 1 ? pr0 + sh : pr0 - sh);
      const bool ok = ((vm >> kh) & (vm >> (8 + kw)) & 1u) != 0;
      uint32_t hi[32], lo[32];
      // padding taps read a row of zeros instead of being masked element by element
      const uint32_t ar = ok ? patch(pb) + pr * 128u : zero_row;
      const uint32_t sw = pr & 7u;
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];"
                     : "=r"(hi[4 * jj]), "=r"(hi[4 * jj + 1]), "=r"(hi[4 * jj + 2]), "=r"(hi[4 * jj + 3])
                     : "r"(ar + (((uint32_t)jj ^ sw) << 4)));
      }
#pragma unroll
      for (int e = 0; e < 32; e += 2) lo_part2(hi[e], hi[e + 1], lo[e], lo[e + 1]);
      __syncwarp();
      // last k-block of this chunk for this warp: its reads of the patch are complete
      if (tap + 2 >= taps && elect_one()) mbar_arrive(pfree_bar(pb));
#pragma unroll
      for (int i = 0; i < 2; ++i)
        if (++kw == KWt) {
          kw = 0;
          if (MODE == 2 || ++kh == a.KH) { kh = 0; ++c; }
        }
      T3_TRACE(trole, kb, 1);
      if (kb >= C::TST) {
        mbar_wait(tfree_bar(t), ((kb / C::TST) - 1) & 1);
        tc_fence_after();
      }
      T3_TRACE(trole, kb, 2);
      const uint32_t tcol = tlane + (uint32_t)(t * 64);
      tmem_st16(tcol, *reinterpret_cast<const uint32_t(*)[16]>(&hi[0]));
      tmem_st16(tcol + 16u, *reinterpret_cast<const uint32_t(*)[16]>(&hi[16]));
      if (!(a.flags & 0x800)) {
        tmem_st16(tcol + 32u, *reinterpret_cast<const uint32_t(*)[16]>(&lo[0]));
        tmem_st16(tcol + 48u, *reinterpret_cast<const uint32_t(*)[16]>(&lo[16]));
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (elect_one()) mbar_arrive(tfull_bar(t));
      T3_TRACE(trole, kb, 3);
    }
    // ---- epilogue: warp (q, half) owns rows 32q..32q+31 and the 16-column chunks half, half+2, ... ----
    mbar_wait(acc_bar, 0);
    tc_fence_after();
    if (ktrace && tid == T3_WSPLITW * 32) a.trace[3 * T3_TRACE_KB * 4 + 2] = clock64();
    long m = m0 + row;
    const bool row_ok = row < rows_valid;
    if (MODE == 2) {                                         // class pixel (ho, wo) -> (2 ho + oph, 2 wo + opw) of y
      const int r = r0 + row;
      const int ho = r / a.Wo, wo = r - ho * a.Wo;
      m = ((long)img * a.OH + 2 * ho + tc.oph) * a.OW + 2 * wo + tc.opw;
    }
    const uint32_t trow = tlane + (uint32_t)C::ACC0;
    const int nmain = nk < C::NMAIN ? nk : C::NMAIN;
#pragma unroll 1
    for (int c = half * 16; c < BN; c += 32) {
      float acc[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) acc[e] = 0.f;
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
            tmem_ld16_nowait(trow + (uint32_t)(col + c), v[u]);
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
      if (BN >= 32 && MODE != 2) {
        // staged in shared memory (both patch buffers are consumed by now) and written by TMA: see conv_tc2.cu
        float o[16];
        bias_act16(acc, a.bias ? a.bias + n0 + c : nullptr, a.act, o);
        stage_out16(base, row, c, o);
      } else if (row_ok && n0 + c < a.N) {
        float o[16];
        bias_act16(acc, a.bias ? a.bias + n0 + c : nullptr, a.act, o);
        float4* dst = reinterpret_cast<float4*>(a.y + m * a.N + n0 + c);
#pragma unroll
        for (int e = 0; e < 4; ++e) dst[e] = make_float4(o[4 * e], o[4 * e + 1], o[4 * e + 2], o[4 * e + 3]);
      }
    }
    if (BN >= 32 && MODE != 2) {
      fence_async_proxy();
      asm volatile("bar.sync 2, %0;" ::"n"(T3_ASPLITW * 32) : "memory");
      if (warp == T3_WSPLITW && elect_one()) {
        // y viewed as [image][pixel][channel]: rows past the end of the image are clipped by the TMA unit
#pragma unroll
        for (int blk = 0; blk < BN / 32; ++blk)
          tma_store_3d(&tm_y, base + (uint32_t)blk * (128u * 128u), n0 + 32 * blk, r0, img);
        tma_store_commit_wait();
      }
    }
    if (a.stats) {
      asm volatile("bar.sync 1, %0;" ::"n"(T3_ASPLITW * 32) : "memory");
      const int i = tid - T3_WSPLITW * 32;
      if (i < 2 * BN) {
        float v;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(cta_sums + 4u * (uint32_t)i));
        const int stat = i / BN, ch = n0 + (i - stat * BN);
        if (ch < a.N) atomicAdd(a.stats + (long)stat * a.N + ch, (double)v);
      }
    }
    if (ktrace && tid == T3_WSPLITW * 32) a.trace[3 * T3_TRACE_KB * 4 + 3] = clock64();
  } else if (warp == T3_WTMA_WARP) {
    // ======================= W tiles by TMA =======================
    int s = 0;
    uint32_t sph = 0;
    const bool tracing = ktrace && lane == 0;
    for (int kb = 0; kb < nk; ++kb, ring_next(s, sph)) {
      T3_TRACE(2, kb, 0);
      if (kb >= nstages) mbar_wait(wfree_bar(s), sph ^ 1u);
      T3_TRACE(2, kb, 1);
      if (elect_one()) {
        const int c = kb / taps, tap = kb - c * taps;
        const int wtap = MODE == 2 ? (tap == 0 ? tc.wt[0] : (tap == 1 ? tc.wt[1] : (tap == 2 ? tc.wt[2] : tc.wt[3]))) : tap;
        mbar_expect_tx(wland_bar(s), C::B_TILE);
        tma_load_2d(b_raw(s), &tm_w, wtap * a.Cg + c * BK, n0, wland_bar(s));
      }
      __syncwarp();
    }
  } else if (warp == T3_PTMA_WARP) {
    // ======================= patches by TMA: chunk c -> buffer c & 1 =======================
    const int nbox = (prows + PBOX - 1) / PBOX;
    for (int c = 0; c < nchunk; ++c) {
      const int pb = c & 1;
      if (c >= 2) mbar_wait(pfree_bar(pb), ((c >> 1) - 1) & 1);
      if (elect_one()) {
        mbar_expect_tx(pfull_bar(pb), (uint32_t)nbox * PBOX * 128u);
        for (int i = 0; i < nbox; ++i)
          tma_load_2d(patch(pb) + (uint32_t)i * PBOX * 128u, &tm_x, c * BK, pstart + i * PBOX, pfull_bar(pb));
      }
      __syncwarp();
    }
  } else {
    // ======================= MMA issuer =======================
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) |
                           ((uint32_t)(BM >> 4) << 24);
    const uint32_t idesc2 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)((2 * BN) >> 3) << 17) |
                            ((uint32_t)(BM >> 4) << 24);                    // N = 2*BN: [W ; W_lo]
    const uint32_t d_corr = tmem_base + (uint32_t)C::ACC0;
    const bool tracing = ktrace && lane == 0;
    const int me = warp == T3_MMA_WARP ? 0 : 1;             // issuer 0: even k-blocks, issuer 1: odd
    int s = 0;
    uint32_t sph = 0;
    if (me) ring_next(s, sph);
    for (int kb = me; kb < nk; kb += 2, ring_next(s, sph), ring_next(s, sph)) {
      const int t = kb % C::TST;
      T3_TRACE(1, kb, 0);
      mbar_wait(wready_bar(s), sph);
      T3_TRACE(1, kb, 1);
      mbar_wait(tfull_bar(t), (kb / C::TST) & 1);
      if (kb > 0) mbar_wait(turn_bar(me), ((kb - 1) >> 1) & 1);   // the other issuer has issued k-block kb-1
      tc_fence_after();
      T3_TRACE(1, kb, 2);
      if (elect_one()) {
        const uint64_t db = make_desc(b_raw(s)), dbl = make_desc(b_lo(s));
        const uint32_t ta = tmem_base + (uint32_t)(t * 64), tal = ta + 32u;
        if (a.flags & 0x800) {
          const uint32_t d_main = d_corr + (uint32_t)(C::PAIR ? BN + (kb % C::NMAIN) * 2 * BN : (1 + kb % C::NMAIN) * BN);
#pragma unroll
          for (int k = 0; k < BK / 8; ++k)
            umma_tf32_ts(d_main, ta + 8u * k, db + (uint64_t)(k * 32 >> 4), idesc, (kb >= C::NMAIN) || (k != 0));
        } else if (C::PAIR) {
          const uint32_t d_pair = d_corr + (uint32_t)(BN + (kb % C::NMAIN) * 2 * BN);
#pragma unroll
          for (int k = 0; k < BK / 8; ++k) {
            const uint64_t adv = (uint64_t)(k * 32 >> 4);
            umma_tf32_ts(d_corr, tal + 8u * k, db + adv, idesc, (kb | k) != 0);
            umma_tf32_ts(d_pair, ta + 8u * k, db + adv, idesc2, (kb >= C::NMAIN) || (k != 0));
          }
        } else {
          const uint32_t d_main = d_corr + (uint32_t)((1 + kb % C::NMAIN) * BN);
#pragma unroll
          for (int k = 0; k < BK / 8; ++k) {
            const uint64_t adv = (uint64_t)(k * 32 >> 4);
            umma_tf32_ts(d_corr, tal + 8u * k, db + adv, idesc, (kb | k) != 0);
            umma_tf32_ts(d_corr, ta + 8u * k, dbl + adv, idesc, 1);
            umma_tf32_ts(d_main, ta + 8u * k, db + adv, idesc, (kb >= C::NMAIN) || (k != 0));
          }
        }
        if (a.flags & 0x1000) tc_fence_before();      // FD_TC_FENCE=1: ordering experiment
        mbar_arrive(turn_bar(me ^ 1));
        umma_commit(wfree_bar(s));
        umma_commit(tfree_bar(t));
      }
      __syncwarp();
      T3_TRACE(1, kb, 3);
    }
    if (elect_one()) umma_commit(acc_bar);
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (ktrace && tid == T3_WSPLITW * 32) a.trace[3 * T3_TRACE_KB * 4 + 4] = clock64();
  if (warp == T3_MMA_WARP) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
  }
}

// worst-case number of patch rows a 128-pixel tile (inside one image) needs: its own raster span in the gathered
// tensor (every output-row change drifts by |Wg - Wo| pixels) + the tap span, in whole TMA boxes
int patch_rows_needed(const TcArgs& a) {
  const long dW = a.Wg > a.Wo ? a.Wg - a.Wo : a.Wo - a.Wg;
  const long row_changes = (BM - 2) / a.Wo + 1;
  long span1 = (long)(a.KH - 1) * a.Wg + a.KW;             // largest tap shift + 1
  if (a.ncls > 0) {
    span1 = 1;
    for (int i = 0; i < a.ncls; ++i)
      for (int t = 0; t < a.cl[i].ntap; ++t) span1 = std::max(span1, (long)a.cl[i].dh[t] * a.Wg + a.cl[i].dw[t] + 1);
  }
  const long need = (BM - 1) + dW * row_changes + span1;
  const long rows = (need + PBOX - 1) / PBOX * PBOX;
  return rows > 4096 ? -1 : (int)rows;
}
// W stages that fit next to two patch buffers of `prows` rows
template <int BN>
int stages_for(int prows) {
  using C = Cfg3<BN>;
  const int left = C::SMEM_MAX - C::MISC - 2 * prows * 128;
  const int n = left / C::WSTAGE;
  return n > C::MAXST ? C::MAXST : n;
}

template <int BN, int MODE>
int launch_tc3(const TcArgs& a, cudaStream_t st) {
  using C = Cfg3<BN>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc3_kernel<BN, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         C::SMEM_MAX);
    if (e != cudaSuccess) {
      fd::set_error("conv_tc3: cannot reserve %d B of shared memory: %s", C::SMEM_MAX, cudaGetErrorString(e));
      return 1;
    }
    configured = true;
  }
  const int prows = patch_rows_needed(a);
  const int nstages = stages_for<BN>(prows);
  const int smem = 2 * prows * 128 + nstages * C::WSTAGE + C::MISC;
  CUtensorMap tw, tx;
  int rc = make_map_2d(&tw, a.w, a.N, a.K, BN);
  if (rc) return rc;
  rc = make_map_2d(&tx, a.x, (long)a.B * a.Hg * a.Wg, a.Cg, PBOX);
  if (rc) return rc;
  CUtensorMap ty = tw;
  if (BN >= 32 && MODE != 2) {
    rc = make_map_3d(&ty, a.y, a.B, (long)a.Ho * a.Wo, a.N, BM);
    if (rc) return rc;
  }
  dim3 grid(a.B * fd::cdiv((long)a.Ho * a.Wo, BM), fd::cdiv(a.N, BN), MODE == 2 ? a.ncls : 1);
  conv_tc3_kernel<BN, MODE><<<grid, T3_NTHREADS, smem, st>>>(a, tw, tx, ty, prows, nstages);
  FD_CHECK_LAUNCH();
  return 0;
}

template <int BN>
bool patch_fits(const TcArgs& a) {
  const int prows = patch_rows_needed(a);
  // the epilogue stages the output tile (BN * 512 B) over the patch buffers and, past them, the W stages
  return prows > 0 && stages_for<BN>(prows) >= 3;
}

template <int BN>
int launch_tc3_tap(const TcArgs& a, cudaStream_t st) {
  if (!patch_fits<BN>(a)) return -1;
  return launch_tc3<BN, 2>(a, st);
}

}  // namespace

namespace fd {
int conv_tc3_dgrad_s2(const TcArgs& g, cudaStream_t st) {
  // g: x = dY [B,Hg,Wg,Cg = Cout], w = W^T [N = Cin][KH*KW*Cout], y = dX [B,Ho,Wo,N], stride 2
  const int H = g.Ho, W = g.Wo;
  const bool k3 = g.KH == 3 && g.KW == 3 && g.pad == 1, k1 = g.KH == 1 && g.KW == 1 && g.pad == 0;
  if (g.stride != 2 || !(k3 || k1) || (H & 1) || (W & 1) || g.Cg % BK != 0 || g.N % 16 != 0) return -1;
  if (g.Hg != H / 2 || g.Wg != W / 2) return -1;
  if ((long)g.B * g.Hg * g.Wg >= (1L << 31) - 65536 || g.M >= (1L << 31)) return -1;
  if ((((uintptr_t)g.w | (uintptr_t)g.x | (uintptr_t)g.y) & 15) != 0) return -1;
  const int bn = g.N % 128 == 0 ? 128 : (g.N % 64 == 0 ? 64 : (g.N % 32 == 0 ? 32 : 16));
  TcArgs a = g;
  a.trace = nullptr;
  a.ncls = 0;
  bool empty_class = false;
  for (int ph = 0; ph < 2; ++ph)
    for (int pw = 0; pw < 2; ++pw) {
      TcArgs::TapClass c{};
      for (int kh = 0; kh < g.KH; ++kh)
        for (int kw = 0; kw < g.KW; ++kw) {
          const int th = ph + g.pad - kh, tw = pw + g.pad - kw;        // dY row = (h + pad - kh) / 2 for h = 2 i + ph
          if (th < 0 || tw < 0 || (th & 1) || (tw & 1)) continue;
          c.dh[c.ntap] = th / 2; c.dw[c.ntap] = tw / 2; c.wt[c.ntap] = kh * g.KW + kw;
          ++c.ntap;
        }
      if (c.ntap == 0) { empty_class = true; continue; }
      c.oph = ph; c.opw = pw;
      a.cl[a.ncls++] = c;
    }
  // the classes run as blockIdx.z of ONE launch (four launches of 24 ... 90 CTAs each ran one after the other)
  a.KH = 1; a.KW = 4; a.stride = 1; a.pad = 0;
  a.OH = H; a.OW = W;
  a.Ho = H / 2; a.Wo = W / 2;
  a.M = (long)g.B * a.Ho * a.Wo;
  if (empty_class) {                             // 1x1 / 2: only the even-even pixels receive a gradient
    cudaError_t e = cudaMemsetAsync(g.y, 0, sizeof(float) * (size_t)g.B * H * W * g.N, st);
    if (e != cudaSuccess) { fd::set_error("conv_tc3_dgrad_s2: memset failed: %s", cudaGetErrorString(e)); return 1; }
  }
  const int rc = bn == 128 ? launch_tc3_tap<128>(a, st)
                           : (bn == 64 ? launch_tc3_tap<64>(a, st)
                                       : (bn == 32 ? launch_tc3_tap<32>(a, st) : launch_tc3_tap<16>(a, st)));
  return rc;
}

// returns -1 when this variant does not take the problem (the caller falls back to conv_tc2)
int conv_tc3_dispatch(const TcArgs& a0, int mode, cudaStream_t st) {
  TcArgs a = a0;
  a.trace = g_conv_trace_host;
  if (a.stride != 1 || a.Cg % BK != 0 || a.KH > 8 || a.KW > 8 || a.M >= (1L << 31)) return -1;
  if ((long)a.B * a.Hg * a.Wg >= (1L << 31) - 65536) return -1;
  if ((((uintptr_t)a.w | (uintptr_t)a.x | (uintptr_t)a.y) & 15) != 0) return -1;
  const int bn = a.N % 128 == 0 ? 128 : (a.N % 64 == 0 ? 64 : (a.N % 32 == 0 ? 32 : 16));
  if (bn == 128) {
    if (!patch_fits<128>(a)) return -1;
    return mode == 0 ? launch_tc3<128, 0>(a, st) : launch_tc3<128, 1>(a, st);
  }
  if (bn == 64) {
    if (!patch_fits<64>(a)) return -1;
    return mode == 0 ? launch_tc3<64, 0>(a, st) : launch_tc3<64, 1>(a, st);
  }
  if (bn == 32) {
    if (!patch_fits<32>(a)) return -1;
    return mode == 0 ? launch_tc3<32, 0>(a, st) : launch_tc3<32, 1>(a, st);
  }
  if (!patch_fits<16>(a)) return -1;
  return mode == 0 ? launch_tc3<16, 0>(a, st) : launch_tc3<16, 1>(a, st);
}
}  // namespace fd
