// Implicit-GEMM convolution (forward / data gradient, stride 1) on tcgen05 -- the PERSISTENT variant of
// conv_tc3.cu for launches with more tiles than SMs (the 64-channel layers at 1/4 resolution: 360 tiles).
//
// Same contract, numerics and data path as conv_tc3 (3xTF32, input patch by TMA once per 32-channel chunk, A
// operand split into tensor memory, W_lo computed in shared memory, two MMA issuers).  What changes is the
// schedule.  The role trace of conv_tc3 on the layer-1 shape (tools/trace_conv3.py) put a CTA at ~19.4 k cycles
// of which only ~10.2 k are the 18 k-blocks of main loop: ~4.1 k pass before the first MMA (barrier / TMEM
// set-up, ~1.3 k cycles of TMA latency for the first patch and W tile, first split) and ~4.4 k in the epilogue
// (TMEM reads at 64 B/clk, bias / activation / statistics, stores), and 360 one-tile CTAs on 148 SMs run as 3
// waves although the work is 2.43.  Here one CTA per SM walks tiles T = blockIdx.x, + gridDim.x, ...:
//
//   * every pipeline (W stages, TMEM A-ring, patch double buffer, issuer turn) runs on across tile boundaries --
//     the producers are already fetching and splitting tile i+1 while tile i's last MMAs execute;
//   * the epilogue has its own four warps; the accumulators are double-buffered in tensor memory for BN <= 64
//     (2 x [corr | main | corr2]), so tile i drains while tile i+1 accumulates.  BN = 128 has tensor memory for
//     one set only: its issuers wait for the drain (the TMEM reads), not for the stores;
//   * tiles are dealt round-robin, so an SM does 2 or 3 tiles back to back without a launch gap.
//
// With one accumulator per product class for BN <= 64 (no rotation) the longest chain of truncating tensor-core
// accumulations is K = 1152 here; measured error vs fp64 stays below the 1e-5 bound of the parity tests.
//
// Roles (640 threads): warps 0-3 W_lo splitters, 4-11 A splitters (two groups on alternate k-blocks), 12 / 15 MMA
// issuers (even / odd k-blocks), 13 W tiles by TMA, 14 patches by TMA, 16-19 epilogue.
#include "tc_common.cuh"

namespace {

constexpr int T4_WSPLITW = 4, T4_EPIW = 4;
// warp layout for NG splitter groups: [0,4) W_lo splitters, [4, 4+4NG) A splitters, then MMA issuer 0, W TMA,
// patch TMA, MMA issuer 1, four epilogue warps (warp index % 4 = TMEM lane quadrant for splitters and epilogue)
constexpr int PBOX4 = 64;                                   // patch rows per TMA box
// Timing experiment (fd_debug_set_conv_trace + tools/trace_conv3.py): CTA 0 stamps clock64() into
// trace[role][k-block over all its tiles][4] (role 0 = A splitter group 0, 1 = MMA issuers, 2 = epilogue, per tile)
constexpr int T4_TRACE_KB = 256;
#define T4_TRACE(role, kb, slot)                                                                   \
  do {                                                                                             \
    if (tracing && (kb) < T4_TRACE_KB) a.trace[((role) * T4_TRACE_KB + (kb)) * 4 + (slot)] = clock64(); \
  } while (0)

template <int BN>
struct Cfg4 {
  static constexpr int B_TILE = BN * 128;
  static constexpr int WSTAGE = 2 * B_TILE;                             // [W ; W_lo]
  static constexpr int MAXST = 8;
  // Output staging for the TMA stores: SBLK blocks of [128 rows x 32 columns] (SWIZZLE_128B image).  Plain
  // st.global from the epilogue warps (32 rows x 16 B per instruction) halved the k-block rate of the main loop for
  // as long as an epilogue ran -- the splitters' ld.shared queue behind those stores in the SM's one L1 pipe
  // (trace with FD_TC4_EXP=1: no stores, no slow-down) -- the TMA unit reads the tile through the async proxy.
  static constexpr int NBLK = BN / 32;
  static constexpr int SBLK = NBLK < 2 ? NBLK : 2;
  static constexpr int STG = SBLK * 128 * 128;
  static constexpr int MISC = 1024 /*align*/ + STG + 512 /*barriers, zero row*/ + 1024 /*CTA channel sums*/;
  static constexpr int SMEM_MAX = 232448;
  // A splitter groups = TMEM A-ring slots.  Measured: a third group (NG = 3 for BN <= 64, 768 threads at 80
  // registers) leaves the 64-wide kernel at ~570-590 clk per k-block -- it is not splitter-bound but at its
  // shared-memory roofline: per k-block the MMAs read 24 KB of B, TMA writes 6.4 KB of patch + 8 KB of W, the
  // splitters read 16 KB and the W_lo pass reads + writes 16 KB = 70 KB through a 128 B/clk port = 550 clk.
  static constexpr int NG = 2;
  static constexpr int TST = NG;
  static constexpr int ASPLITW = 4 * NG;
  static constexpr int MMA_WARP = T4_WSPLITW + ASPLITW, WTMA_WARP = MMA_WARP + 1, PTMA_WARP = MMA_WARP + 2,
                       MMA2_WARP = MMA_WARP + 3, EPI_WARP0 = MMA_WARP + 4;
  static constexpr int NTHREADS = (EPI_WARP0 + T4_EPIW) * 32;           // 768 / 640
  static constexpr int ACC0 = TST * 64;
  static constexpr bool PAIR = BN <= 64;
  static constexpr int NSETS = BN <= 64 ? 2 : 1;                        // accumulator sets in tensor memory
  static constexpr int NMAIN = BN <= 64 ? 1 : 2;
  // PAIR: [main | corr] per rotation slot -- A * [W ; W_lo] lands in both halves, A_lo * W is added to the corr
  // half (one accumulator less to read back than conv_tc3's separate corr1); else corr + NMAIN mains
  static constexpr int SETCOLS = PAIR ? NMAIN * 2 * BN : (1 + NMAIN) * BN;
  static_assert(ACC0 + NSETS * SETCOLS <= 512, "tensor memory");
};

struct Tile4 {
  int img, r0, n0, rows_valid;
};

// n / d for 0 <= n < 2^23, d > 0, inv = 1.0f / d: the per-tile set-up (tile id -> image / row, row -> (ho, wo)) sat
// on the critical path at every tile boundary with ~1500 clk of dependent integer-division code (role trace)
__device__ __forceinline__ int fast_div(int n, int d, float inv) {
  int q = __float2int_rz(__int2float_rn(n) * inv);
  const int r = n - q * d;
  q += (r >= d) - (r < 0);
  return q;
}
// bits u of [0, n) with lo <= u < hi
__device__ __forceinline__ uint32_t range_bits(int lo, int hi, int n) {
  lo = lo < 0 ? 0 : lo;
  hi = hi > n ? n : hi;
  return hi > lo ? (((1u << hi) - 1u) & ~((1u << lo) - 1u)) : 0u;
}

template <int BN, int MODE>
__global__ void __launch_bounds__(Cfg4<BN>::NTHREADS, 1)
conv_tc4_kernel(TcArgs a, const __grid_constant__ CUtensorMap tm_w, const __grid_constant__ CUtensorMap tm_x,
                const __grid_constant__ CUtensorMap tm_y, const int patch_rows, const int nstages, const int gx, const int total_tiles) {
  using C = Cfg4<BN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t patch_bytes = (uint32_t)patch_rows * 128u;
  auto patch = [&](int b) { return base + (uint32_t)b * patch_bytes; };
  const uint32_t wbase = base + 2u * patch_bytes;
  auto b_raw = [&](int s) { return wbase + (uint32_t)s * C::WSTAGE; };
  auto b_lo = [&](int s) { return wbase + (uint32_t)s * C::WSTAGE + C::B_TILE; };
  const uint32_t stg = wbase + (uint32_t)nstages * C::WSTAGE;      // output staging (epilogue warps)
  const uint32_t bars = stg + (uint32_t)C::STG;
  auto wland_bar = [&](int s) { return bars + 8u * s; };                        // W tile landed (TMA tx)
  auto wready_bar = [&](int s) { return bars + 8u * (C::MAXST + s); };          // W_lo written
  auto wfree_bar = [&](int s) { return bars + 8u * (2 * C::MAXST + s); };       // MMAs done with the stage
  auto tfull_bar = [&](int t) { return bars + 8u * (24 + t); };                 // A / A_lo of a k-block in TMEM
  auto tfree_bar = [&](int t) { return bars + 8u * (28 + t); };
  auto pfull_bar = [&](int b) { return bars + 8u * (32 + b); };                 // patch buffer landed
  auto pfree_bar = [&](int b) { return bars + 8u * (34 + b); };
  auto accfull_bar = [&](int s) { return bars + 8u * (36 + s); };               // a tile's MMAs are complete
  auto accfree_bar = [&](int s) { return bars + 8u * (38 + s); };               // the epilogue has read the set
  auto turn_bar = [&](int i) { return bars + 8u * (40 + i); };                  // issuer i may issue
  const uint32_t tmem_slot = bars + 8u * 42;
  const uint32_t pinfo = bars + 8u * 43;                    // [2] patch start (pixel index) of the chunk in buffer b
  const uint32_t zero_row = bars + 384u;                    // 128 B of zeros: the "row" a padding tap reads
  const uint32_t cta_sums = bars + 512u;                    // [2][BN] floats, only with a.stats
  auto ring_next = [&](int& st, uint32_t& ph) { if (++st == nstages) { st = 0; ph ^= 1u; } };

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool ktrace = a.trace && blockIdx.x == 0;
  if (ktrace && tid == 0) {
    a.trace[3 * T4_TRACE_KB * 4 + 0] = clock64();
    long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    a.trace[3 * T4_TRACE_KB * 4 + 5] = gt;                  // wall-clock ns: launch-to-launch gaps, SM clock rate
  }
  // Tiles never cross an image (see conv_tc3.cu).  Tile id T -> (m tile x, n tile y), x fastest.
  const int HoWo = a.Ho * a.Wo;
  const int tpi = (HoWo + BM - 1) / BM;
  const int taps = a.KH * a.KW;
  const int nchunk = a.Cg / BK;
  const int nk = taps * nchunk;
  const int span = (a.KH - 1) * a.Wg + (a.KW - 1);         // largest tap shift
  // Active splitter groups.  A group must wait on EVERY phase of a patch-buffer barrier it uses (mbarrier parity
  // waits only tell adjacent phases apart), i.e. touch every chunk: true when a chunk spans >= NG k-blocks; a 1x1
  // conv (one k-block per chunk) runs on two groups, where group = chunk parity = buffer.
  const int ng = taps >= C::NG ? C::NG : 2;
  const float inv_gx = 1.0f / (float)gx, inv_tpi = 1.0f / (float)tpi, inv_wo = 1.0f / (float)a.Wo;
  auto tile_of = [&](int T) {
    Tile4 t;
    const int y = fast_div(T, gx, inv_gx), x = T - y * gx;
    t.img = fast_div(x, tpi, inv_tpi);
    t.r0 = (x - t.img * tpi) * BM;
    t.n0 = y * BN;
    t.rows_valid = HoWo - t.r0 < BM ? HoWo - t.r0 : BM;
    return t;
  };
  // tile row -> (pixel index of its tap-(0,0) source, valid-tap mask); bit 31 of the mask = the row exists
  auto row_geom = [&](const Tile4& t, int row, int& lin, uint32_t& mask) {
    lin = 0;
    mask = 0;
    if (row < t.rows_valid) {
      const int r = t.r0 + row;
      const int ho = fast_div(r, a.Wo, inv_wo), wo = r - ho * a.Wo;
      int hq, wq;
      if (MODE == 0) {
        hq = ho - a.pad; wq = wo - a.pad;        // tap u reads hq + u: valid for -hq <= u < Hg - hq
        mask = range_bits(-hq, a.Hg - hq, a.KH) | (range_bits(-wq, a.Wg - wq, a.KW) << 8);
      } else {
        hq = ho + a.pad; wq = wo + a.pad;        // tap u reads hq - u: valid for hq - Hg < u <= hq
        mask = range_bits(hq - a.Hg + 1, hq + 1, a.KH) | (range_bits(wq - a.Wg + 1, wq + 1, a.KW) << 8);
      }
      lin = (t.img * a.Hg + hq) * a.Wg + wq;
      mask |= 1u << 31;
    }
  };

  if (tid >= 256 && tid < 288)
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(zero_row + 4u * (uint32_t)(tid - 256)), "r"(0u) : "memory");
  if (tid == 0) {
    for (int s = 0; s < nstages; ++s) {
      mbar_init(wland_bar(s), 1);
      mbar_init(wready_bar(s), T4_WSPLITW);
      mbar_init(wfree_bar(s), 1);
    }
    for (int t = 0; t < C::TST; ++t) {
      mbar_init(tfull_bar(t), 4);
      mbar_init(tfree_bar(t), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(pfull_bar(b), 1);
      mbar_init(pfree_bar(b), 4 * (taps < ng ? taps : ng));             // the groups that touch a chunk (1x1: one)
      mbar_init(accfull_bar(b), 2);                                        // both MMA issuers commit to it
      mbar_init(accfree_bar(b), T4_EPIW);
      mbar_init(turn_bar(b), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_w) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_x) : "memory");
  }
  if (warp == C::MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp < T4_WSPLITW) {
    // ======================= weight splitters: W_lo = W - tf32(W) in shared memory =======================
    constexpr int NV = C::B_TILE / 16 / (T4_WSPLITW * 32);          // float4 per thread per stage
    int s = 0;
    uint32_t sph = 0;
    for (int T = blockIdx.x; T < total_tiles; T += gridDim.x) {
      for (int kb = 0; kb < nk; ++kb, ring_next(s, sph)) {
        mbar_wait(wland_bar(s), sph);
        if (!(a.flags & 0x800)) {
#pragma unroll
          for (int i = 0; i < (NV > 0 ? NV : 1); ++i) {
            const uint32_t idx = (uint32_t)(tid + T4_WSPLITW * 32 * i);
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
      }
    }
  } else if (warp < C::MMA_WARP) {
    // ======================= A splitters: group `half` takes the k-blocks g with g % NG == half =======================
    // (g counts k-blocks over all of this CTA's tiles; TMEM slot = half)
    const int q = warp & 3;                       // TMEM lane quadrant this warp may access
    const int half = (warp - T4_WSPLITW) >> 2;
    const int row = q * 32 + lane;
    const uint32_t tcol = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * 64);
    int gbase = 0, gcb = 0;                       // k-blocks / chunks of the earlier tiles
    uint32_t nuse = 0;                            // uses of this group's TMEM slot so far
    const bool tracing = ktrace && lane == 0 && q == 0 && half == 0;
    for (int T = blockIdx.x; T < total_tiles; T += gridDim.x, gbase += nk, gcb += nchunk) {
      const Tile4 tl = tile_of(T);
      int lin;
      uint32_t vm;
      row_geom(tl, row, lin, vm);
      if (half >= ng) break;
      const int kb0 = (half + ng - gbase % ng) % ng;                  // first k-block of the tile with g % ng == half
      int c = 0, kh = 0, kw = 0;
      for (int i = 0; i < kb0; ++i)
        if (++kw == a.KW) {
          kw = 0;
          if (++kh == a.KH) { kh = 0; ++c; }
        }
      int chunk_seen = -1, pr0 = 0;
      for (int kb = kb0; kb < nk; kb += ng, ++nuse) {
        T4_TRACE(0, gbase + kb, 0);
        const int tap = kh * a.KW + kw;
        const int gc = gcb + c, pb = gc & 1;
        if (c != chunk_seen) {
          mbar_wait(pfull_bar(pb), (uint32_t)(gc >> 1) & 1u);
          if (chunk_seen < 0) {
            int pstart;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(pstart) : "r"(pinfo + 4u * (uint32_t)pb));
            // rows past the image stay inside the buffer and read the zero row
            pr0 = (vm >> 31) ? lin - pstart : (MODE == 0 ? 0 : span);
          }
          chunk_seen = c;
        }
        const int sh = kh * a.Wg + kw;
        const uint32_t pr = (uint32_t)(MODE == 0 ? pr0 + sh : pr0 - sh);
        const bool ok = ((vm >> kh) & (vm >> (8 + kw)) & 1u) != 0;
        uint32_t hi[32], lo[32];
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
        if (tap + ng >= taps && elect_one()) mbar_arrive(pfree_bar(pb));
        for (int i = 0; i < ng; ++i)
          if (++kw == a.KW) {
            kw = 0;
            if (++kh == a.KH) { kh = 0; ++c; }
          }
        T4_TRACE(0, gbase + kb, 1);
        if (nuse >= 1) {
          mbar_wait(tfree_bar(half), (nuse - 1u) & 1u);
          tc_fence_after();
        }
        T4_TRACE(0, gbase + kb, 2);
        tmem_st16(tcol, *reinterpret_cast<const uint32_t(*)[16]>(&hi[0]));
        tmem_st16(tcol + 16u, *reinterpret_cast<const uint32_t(*)[16]>(&hi[16]));
        if (!(a.flags & 0x800)) {
          tmem_st16(tcol + 32u, *reinterpret_cast<const uint32_t(*)[16]>(&lo[0]));
          tmem_st16(tcol + 48u, *reinterpret_cast<const uint32_t(*)[16]>(&lo[16]));
        }
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (elect_one()) mbar_arrive(tfull_bar(half));
        T4_TRACE(0, gbase + kb, 3);
      }
    }
  } else if (warp == C::WTMA_WARP) {
    // ======================= W tiles by TMA =======================
    int s = 0, g = 0;
    uint32_t sph = 0;
    for (int T = blockIdx.x; T < total_tiles; T += gridDim.x) {
      const Tile4 tl = tile_of(T);
      int c = 0, tap = 0;
      for (int kb = 0; kb < nk; ++kb, ++g, ring_next(s, sph)) {
        if (g >= nstages) mbar_wait(wfree_bar(s), sph ^ 1u);
        if (elect_one()) {
          mbar_expect_tx(wland_bar(s), C::B_TILE);
          tma_load_2d(b_raw(s), &tm_w, tap * a.Cg + c * BK, tl.n0, wland_bar(s));
        }
        __syncwarp();
        if (++tap == taps) { tap = 0; ++c; }
      }
    }
  } else if (warp == C::PTMA_WARP) {
    // ======================= patches by TMA: chunk gc -> buffer gc & 1 =======================
    int gc = 0;
    for (int T = blockIdx.x; T < total_tiles; T += gridDim.x) {
      const Tile4 tl = tile_of(T);
      // pixel range the tile's rows start from (not always row 0 / the last row: a pad-0 data gradient steps
      // back by one pixel at every output-row change)
      int lo = 0x7fffffff, hi = (int)0x80000000;
#pragma unroll
      for (int i = 0; i < BM / 32; ++i) {
        int lin;
        uint32_t vm;
        row_geom(tl, lane + 32 * i, lin, vm);
        if (vm >> 31) { lo = min(lo, lin); hi = max(hi, lin); }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
      }
      const int pstart = MODE == 0 ? lo : lo - span;
      const int prows = hi - lo + span + 1;
      const int nbox = (prows + PBOX4 - 1) / PBOX4;
      for (int c = 0; c < nchunk; ++c, ++gc) {
        const int pb = gc & 1;
        if (gc >= 2) mbar_wait(pfree_bar(pb), (uint32_t)((gc >> 1) - 1) & 1u);
        if (elect_one()) {
          asm volatile("st.shared.u32 [%0], %1;" ::"r"(pinfo + 4u * (uint32_t)pb), "r"(pstart) : "memory");
          mbar_expect_tx(pfull_bar(pb), (uint32_t)nbox * PBOX4 * 128u);
          for (int i = 0; i < nbox; ++i)
            tma_load_2d(patch(pb) + (uint32_t)i * PBOX4 * 128u, &tm_x, c * BK, pstart + i * PBOX4, pfull_bar(pb));
        }
        __syncwarp();
      }
    }
  } else if (warp == C::MMA_WARP || warp == C::MMA2_WARP) {
    // ======================= MMA issuers: issuer `me` takes the k-blocks g with g % 2 == me (TMEM slot g % NG) =======================
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) |
                           ((uint32_t)(BM >> 4) << 24);
    const uint32_t idesc2 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)((2 * BN) >> 3) << 17) |
                            ((uint32_t)(BM >> 4) << 24);                    // N = 2*BN: [W ; W_lo]
    const int me = warp == C::MMA_WARP ? 0 : 1;
    int s = 0, gbase = 0, ti = 0;
    uint32_t sph = 0;
    if (me) ring_next(s, sph);
    const bool tracing = ktrace && lane == 0;
    for (int T = blockIdx.x; T < total_tiles; T += gridDim.x, gbase += nk, ++ti) {
      const int set = ti % C::NSETS;
      const uint32_t d_corr = tmem_base + (uint32_t)(C::ACC0 + set * C::SETCOLS);
      if (ti >= C::NSETS) {                        // the epilogue has read this set's previous tile
        mbar_wait(accfree_bar(set), (uint32_t)(ti / C::NSETS - 1) & 1u);
        tc_fence_after();
      }
      for (int kb = (me - gbase) & 1; kb < nk; kb += 2, ring_next(s, sph), ring_next(s, sph)) {
        const int g = gbase + kb;
        const int slot = g % ng;
        const uint32_t nuse = (uint32_t)(g / ng);
        T4_TRACE(1, g, 0);
        mbar_wait(wready_bar(s), sph);
        T4_TRACE(1, g, 1);
        mbar_wait(tfull_bar(slot), nuse & 1u);
        if (g > 0) mbar_wait(turn_bar(me), (uint32_t)((g - 1) >> 1) & 1u);   // the other issuer has issued g-1
        tc_fence_after();
        T4_TRACE(1, g, 2);
        if (elect_one()) {
          const uint64_t db = make_desc(b_raw(s)), dbl = make_desc(b_lo(s));
          const uint32_t ta = tmem_base + (uint32_t)(slot * 64), tal = ta + 32u;
          if (a.flags & 0x800) {
            const uint32_t d_main = d_corr + (uint32_t)(C::PAIR ? (kb % C::NMAIN) * 2 * BN : (1 + kb % C::NMAIN) * BN);
#pragma unroll
            for (int k = 0; k < BK / 8; ++k)
              umma_tf32_ts(d_main, ta + 8u * k, db + (uint64_t)(k * 32 >> 4), idesc, (kb >= C::NMAIN) || (k != 0));
          } else if (C::PAIR) {
            const uint32_t d_pair = d_corr + (uint32_t)((kb % C::NMAIN) * 2 * BN);
#pragma unroll
            for (int k = 0; k < BK / 8; ++k) {
              const uint64_t adv = (uint64_t)(k * 32 >> 4);
              umma_tf32_ts(d_pair, ta + 8u * k, db + adv, idesc2, (kb >= C::NMAIN) || (k != 0));   // [A W | A W_lo]
              umma_tf32_ts(d_pair + (uint32_t)BN, tal + 8u * k, db + adv, idesc, 1);               // += A_lo W
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
          umma_commit(tfree_bar(slot));
        }
        __syncwarp();
        T4_TRACE(1, g, 3);
      }
      if (elect_one()) umma_commit(accfull_bar(set));
      __syncwarp();
    }
  } else {
    // ======================= epilogue: warp q owns tile rows 32q .. 32q+31, all BN columns =======================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int et = tid - C::EPI_WARP0 * 32;          // 0..127
    const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16);
    const bool single = (a.flags & 0x800) != 0;
    const int nmain = nk < C::NMAIN ? nk : C::NMAIN;
    const int nacc = single ? nmain : (C::PAIR ? 2 * nmain : nmain + 1);
    int ti = 0;
    const bool tracing = ktrace && et == 0;
    for (int T = blockIdx.x; T < total_tiles; T += gridDim.x, ++ti) {
      const Tile4 tl = tile_of(T);
      const int set = ti % C::NSETS;
      T4_TRACE(2, ti, 0);
      const uint32_t trow = tlane + (uint32_t)(C::ACC0 + set * C::SETCOLS);
      if (a.stats) {
        for (int i = et; i < 2 * BN; i += T4_EPIW * 32)
          asm volatile("st.shared.f32 [%0], %1;" ::"r"(cta_sums + 4u * (uint32_t)i), "f"(0.f) : "memory");
        asm volatile("bar.sync 1, %0;" ::"n"(T4_EPIW * 32) : "memory");
      }
      mbar_wait(accfull_bar(set), (uint32_t)(ti / C::NSETS) & 1u);
      tc_fence_after();
      T4_TRACE(2, ti, 1);
#pragma unroll 1
      for (int c = 0; c < BN; c += 16) {
        if ((c & (32 * C::SBLK - 1)) == 0) {
          // the staging blocks are free once the TMA unit has read the previous pass out of them
          if (et == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          asm volatile("bar.sync 2, %0;" ::"n"(T4_EPIW * 32) : "memory");
        }
        // accumulators in summation order: main, corr2 (PAIR) ..., then the A_lo * W correction
        float acc[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) acc[e] = 0.f;
        for (int g = 0; g < nacc; g += 2) {
          uint32_t v[2][16];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int i = g + u;
            if (i < nacc) {
              const int col = C::PAIR ? (single ? 2 * i * BN : i * BN)
                                      : (single ? (1 + i) * BN : (i == nacc - 1 ? 0 : (1 + i) * BN));
              if (!(a.flags & 0x20000)) tmem_ld16_nowait(trow + (uint32_t)(col + c), v[u]);   // timing experiment
              else {
#pragma unroll
                for (int e = 0; e < 16; ++e) v[u][e] = 0u;
              }
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
        if (c + 16 >= BN) {
          // every column of the set is in registers: the issuers may overwrite it
          tc_fence_before();
          __syncwarp();
          if (elect_one()) mbar_arrive(accfree_bar(set));
          T4_TRACE(2, ti, 2);
        }
        if (a.stats) {
          // rows past the tile end hold exact zeros (zero A rows), so they add nothing
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
        {
          float o[16];
          bias_act16(acc, a.bias ? a.bias + tl.n0 + c : nullptr, a.act, o);
          stage_out16(stg, row, c & (32 * C::SBLK - 1), o);
        }
        if (((c + 16) & (32 * C::SBLK - 1)) == 0) {
          // pass complete: y viewed as [image][pixel][channel], rows past the end of the image are clipped by the TMA unit
          fence_async_proxy();
          asm volatile("bar.sync 2, %0;" ::"n"(T4_EPIW * 32) : "memory");
          if (et == 0 && !(a.flags & 0x10000)) {   // 0x10000: timing experiment without the stores
            const int c0 = c + 16 - 32 * C::SBLK;
#pragma unroll
            for (int blk = 0; blk < C::SBLK; ++blk)
              tma_store_3d(&tm_y, stg + (uint32_t)blk * (128u * 128u), tl.n0 + c0 + 32 * blk, tl.r0, tl.img);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }
      }
      if (a.stats) {
        asm volatile("bar.sync 1, %0;" ::"n"(T4_EPIW * 32) : "memory");
        for (int i = et; i < 2 * BN; i += T4_EPIW * 32) {
          float v;
          asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(cta_sums + 4u * (uint32_t)i));
          const int stat = i / BN, ch = tl.n0 + (i - stat * BN);
          if (ch < a.N) atomicAdd(a.stats + (long)stat * a.N + ch, (double)v);
        }
      }
      T4_TRACE(2, ti, 3);
    }
    if (et == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // the last stores have left the CTA
  }
  tc_fence_before();
  __syncthreads();
  if (ktrace && tid == 0) {
    a.trace[3 * T4_TRACE_KB * 4 + 4] = clock64();
    long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    a.trace[3 * T4_TRACE_KB * 4 + 6] = gt;
  }
  if (warp == C::MMA_WARP) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
  }
}

// worst-case number of patch rows a 128-pixel tile (inside one image) needs, in whole TMA boxes (conv_tc3.cu)
int patch_rows_needed4(const TcArgs& a) {
  const long dW = a.Wg > a.Wo ? a.Wg - a.Wo : a.Wo - a.Wg;
  const long row_changes = (BM - 2) / a.Wo + 1;
  const long need = (BM - 1) + dW * row_changes + (long)(a.KH - 1) * a.Wg + a.KW;
  const long rows = (need + PBOX4 - 1) / PBOX4 * PBOX4;
  return rows > 4096 ? -1 : (int)rows;
}
template <int BN>
int stages_for4(int prows) {
  using C = Cfg4<BN>;
  const int left = C::SMEM_MAX - C::MISC - 2 * prows * 128;
  const int n = left / C::WSTAGE;
  return n > C::MAXST ? C::MAXST : n;
}

int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

template <int BN, int MODE>
int launch_tc4(const TcArgs& a, cudaStream_t st) {
  using C = Cfg4<BN>;
  const int prows = patch_rows_needed4(a);
  if (prows < 0) return -1;
  const int nstages = stages_for4<BN>(prows);
  if (nstages < 3) return -1;
  const int gx = a.B * fd::cdiv((long)a.Ho * a.Wo, BM);
  const int total = gx * fd::cdiv(a.N, BN);
  const int sms = sm_count();
  if (total <= sms) return -1;                 // one tile per CTA: nothing to overlap, conv_tc3 does it
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc4_kernel<BN, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         C::SMEM_MAX);
    if (e != cudaSuccess) {
      fd::set_error("conv_tc4: cannot reserve %d B of shared memory: %s", C::SMEM_MAX, cudaGetErrorString(e));
      return 1;
    }
    configured = true;
  }
  CUtensorMap tw, tx;
  int rc = make_map_2d(&tw, a.w, a.N, a.K, BN);
  if (rc) return rc;
  rc = make_map_2d(&tx, a.x, (long)a.B * a.Hg * a.Wg, a.Cg, PBOX4);
  if (rc) return rc;
  CUtensorMap ty;
  rc = make_map_3d(&ty, a.y, a.B, (long)a.Ho * a.Wo, a.N, BM);
  if (rc) return rc;
  const int smem = 2 * prows * 128 + nstages * C::WSTAGE + C::MISC;
  conv_tc4_kernel<BN, MODE><<<sms, C::NTHREADS, smem, st>>>(a, tw, tx, ty, prows, nstages, gx, total);
  FD_CHECK_LAUNCH();
  return 0;
}

}  // namespace

namespace fd {
// returns -1 when this variant does not take the problem (the caller falls back to conv_tc3 / conv_tc2)
int conv_tc4_dispatch(const TcArgs& a0, int mode, cudaStream_t st) {
  TcArgs a = a0;
  a.trace = g_conv_trace_host;
  static int exp_flags = -1;                    // FD_TC4_EXP: timing experiments (1: no output stores, 2: no TMEM reads)
  if (exp_flags < 0) {
    const char* e = getenv("FD_TC4_EXP");
    exp_flags = e ? atoi(e) << 16 : 0;
  }
  a.flags |= exp_flags;
  if (a.stride != 1 || a.Cg % BK != 0 || a.KH > 8 || a.KW > 8 || a.M >= (1L << 23)) return -1;   // fast_div range
  if ((long)a.B * a.Hg * a.Wg >= (1L << 31) - 65536) return -1;
  if ((((uintptr_t)a.w | (uintptr_t)a.x | (uintptr_t)a.y) & 15) != 0) return -1;
  if (a.N % 128 == 0) return mode == 0 ? launch_tc4<128, 0>(a, st) : launch_tc4<128, 1>(a, st);
  if (a.N % 64 == 0) return mode == 0 ? launch_tc4<64, 0>(a, st) : launch_tc4<64, 1>(a, st);
  if (a.N % 32 == 0) return mode == 0 ? launch_tc4<32, 0>(a, st) : launch_tc4<32, 1>(a, st);
  return -1;
}
}  // namespace fd
