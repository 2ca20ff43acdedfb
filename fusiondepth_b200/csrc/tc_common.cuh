// Device / host helpers shared by the tcgen05 convolution kernels (conv_tc.cu, conv_tc2.cu):
// mbarrier, cp.async, TMA, tcgen05 MMA / commit / TMEM load-store wrappers and descriptors.
#pragma once
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "../../include/fusiondepth_b200.h"

namespace fd {
struct TcArgs {
  const float* x;     // gathered tensor [B,Hg,Wg,Cg]
  const float* w;     // [N][K] raw fp32
  const float* wlo;   // [N][K] low-order part
  const float* bias;
  float* y;           // [M][N]
  int B, Hg, Wg, Cg, Ho, Wo, N, KH, KW, stride, pad, act;
  long M;
  int K;
  int flags;          // conv_tc2: bit 0 = loader signals `landed` per warp through cp.async groups
  double* stats;      // conv_tc2 forward, optional: [2][N] per-channel sum / sum of squares of y (+=)
  long long* trace;   // timing experiment (fd_debug_set_conv_trace): clock64() stamps of CTA (0,0), else null
  // conv_tc3 MODE 2 (tap-table mode: the output-parity classes of a stride-2 data gradient, one per blockIdx.z,
  // see conv_tc3.cu): class taps t < ntap read the gathered pixel (ho + dh[t], wo + dw[t]) with weight tap wt[t];
  // output pixel (ho, wo) of the class lands at (2 ho + oph, 2 wo + opw) of the [B, OH, OW, N] tensor y
  struct TapClass { int ntap, dh[4], dw[4], wt[4], oph, opw; };
  TapClass cl[4];
  int ncls, OH, OW;
};
// device buffer registered by fd_debug_set_conv_trace (conv_tc2.cu), or null
extern long long* g_conv_trace_host;
// conv_tc2.cu: A-operand-in-TMEM kernels (mode 0 forward, 1 data gradient)
int conv_tc2_dispatch(const TcArgs& a, int mode, cudaStream_t st);
// conv_tc3.cu: the same with the input patch fetched once per 32-channel chunk by TMA and W_lo computed in
// shared memory (stride-1 layers whose patch fits); returns -1 when it does not take the problem
int conv_tc3_dispatch(const TcArgs& a, int mode, cudaStream_t st);
// stride-2 data gradient as four stride-1 tap-table convolutions over dY, one per output-parity class (conv_tc3.cu);
// a holds the plain dgrad problem (x = dY, w = W^T, y = dX, Ho x Wo = the full input size); -1 = not taken
int conv_tc3_dgrad_s2(const TcArgs& a, cudaStream_t st);
// conv_tc4.cu: persistent variant of conv_tc3 (epilogue overlapped with the next tile) for launches with more
// tiles than SMs; returns -1 when it does not take the problem
int conv_tc4_dispatch(const TcArgs& a, int mode, cudaStream_t st);
struct WgTcArgs {
  const float* x;   // [B,H,W,Cin]
  float* dw;        // [N][K]
  int B, H, W, Cin, Ho, Wo, N, KH, KW, stride, pad;
  int M, K;
  int p_per_split;  // pixels per blockIdx.z (multiple of 32)
  int dbg;
  int flags;        // bit 0: per-warp elected barrier arrivals + cp.async groups; 0x800: single-pass TF32
};
// conv_wgrad2.cu: weight gradient with the gathered-input operand in tensor memory; -1 = not taken
int conv_wgrad2_dispatch(const WgTcArgs& a, const float* dy, cudaStream_t st);
}  // namespace fd

namespace {
using fd::TcArgs;
using fd::WgTcArgs;
constexpr int BM = 128, BK = 32;
constexpr int A_TILE = BM * 128;   // bytes of one [128 rows x 32 floats] tile

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (spin > (1u << 28)) __trap();   // never hang the device on a protocol error
  }
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_async_proxy() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// shared-memory matrix descriptor: K-major, SWIZZLE_128B, 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)1 << 16;                 // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;       // stride byte offset
  d |= (uint64_t)1 << 46;                 // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                 // SWIZZLE_128B
  return d;
}
// MN-major tf32 operands must use the SWIZZLE_128B_BASE32B layout (32-byte swizzle atoms: within a
// 128-byte row the 32-byte chunk index is XORed with row & 3; 4-row groups of 512 B).  Descriptor:
// 32-float column blocks `lbo` bytes apart, 4-row groups 512 B apart.
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr, uint32_t lbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo >> 4) & 0x3fff) << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;                 // SWIZZLE_128B_BASE32B
  return d;
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}

__device__ __forceinline__ float act_fn(float v, int act) {
  switch (act) {
    case FD_ACT_RELU: return fmaxf(v, 0.f);
    case FD_ACT_ELU: return v > 0.f ? v : expm1f(v);
    case FD_ACT_SIGMOID: return 1.f / (1.f + expf(-v));
    case FD_ACT_TANH: return tanhf(v);
    default: return v;
  }
}


// o[e] = act(acc[e] + bias[e]) for one 16-column chunk.  The activation switch sits OUTSIDE the element loop:
// per element (act_fn in an unrolled loop) ptxas emits sixteen jump tables over ~60 KB of inlined expm1f / expf /
// tanhf code, and the epilogue spent ~9 k cycles per CTA in indirect branches and instruction-cache misses
// (23 % of the layer-1 kernel, found with the role trace + SASS).
__device__ __forceinline__ void bias_act16(const float (&acc)[16], const float* __restrict__ bias, int act,
                                           float (&o)[16]) {
#pragma unroll
  for (int e = 0; e < 16; ++e) o[e] = acc[e] + (bias ? bias[e] : 0.f);
  if (act == FD_ACT_NONE) return;
  if (act == FD_ACT_RELU) {
#pragma unroll
    for (int e = 0; e < 16; ++e) o[e] = fmaxf(o[e], 0.f);
  } else if (act == FD_ACT_ELU) {
#pragma unroll 4
    for (int e = 0; e < 16; ++e) o[e] = o[e] > 0.f ? o[e] : expm1f(o[e]);
  } else if (act == FD_ACT_SIGMOID) {
#pragma unroll 4
    for (int e = 0; e < 16; ++e) o[e] = 1.f / (1.f + expf(-o[e]));
  } else {
#pragma unroll 4
    for (int e = 0; e < 16; ++e) o[e] = tanhf(o[e]);
  }
}

__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_arrive_noinc(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int x, int y, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(x), "r"(y)
      : "memory");
}
// TMA tensor store of a staged [rows x 32 floats] SWIZZLE_128B block (bulk async-group completion)
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int x, int y) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(map), "r"(src), "r"(x), "r"(y)
               : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t src, int x, int y, int z) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(map), "r"(src), "r"(x), "r"(y), "r"(z)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit_wait() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
// One output row chunk (16 floats at tile column c) of tile row `row` into the staged output tile: 32-column
// blocks of [128 rows x 128 B], 16-byte chunks XOR-swizzled by (row & 7) -- the SWIZZLE_128B image a TMA store
// expects, and bank-conflict free for "one row per lane" writes.
__device__ __forceinline__ void stage_out16(uint32_t out_smem, int row, int c, const float (&o)[16]) {
  const uint32_t blk = out_smem + (uint32_t)(c >> 5) * (128u * 128u) + (uint32_t)row * 128u;
  const uint32_t j0 = (uint32_t)(c & 31) >> 2, sw = (uint32_t)row & 7u;
#pragma unroll
  for (int e = 0; e < 4; ++e)
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(blk + (((j0 + e) ^ sw) << 4)), "f"(o[4 * e]),
                 "f"(o[4 * e + 1]), "f"(o[4 * e + 2]), "f"(o[4 * e + 3])
                 : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&o)[16]) {
  uint32_t v[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int q = 0; q < 16; ++q) o[q] = __uint_as_float(v[q]);
}
__device__ __forceinline__ float lo_part(float v) {
  return v - __uint_as_float(__float_as_uint(v) & 0xffffe000u);
}

// low-order parts of two values with one packed subtraction (FADD2)
__device__ __forceinline__ void lo_part2(uint32_t v0, uint32_t v1, uint32_t& l0, uint32_t& l1) {
  unsigned long long x, y, d;
  asm("mov.b64 %0, {%1,%2};" : "=l"(x) : "r"(v0), "r"(v1));
  asm("mov.b64 %0, {%1,%2};" : "=l"(y) : "r"(v0 & 0xffffe000u), "r"(v1 & 0xffffe000u));
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(x), "l"(y));
  asm("mov.b64 {%0,%1}, %2;" : "=r"(l0), "=r"(l1) : "l"(d));
}

// One lane of a converged warp.  Unlike `lane == 0`, ptxas knows that code guarded by elect.sync runs
// on exactly one thread, so warp-uniform instructions (UTCHMMA, UTMALDG, UTCBAR, barrier arrives on
// uniform addresses) are issued straight from the uniform datapath instead of through a per-thread
// election loop around every instruction (measured: ~100 cycles per tcgen05.mma issued).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// A operand from tensor memory (lane = tile row, one 32-bit column per K element), B from shared memory
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 16 consecutive 32-bit columns of this thread's TMEM lane (no wait: pair with tmem_wait_ld / _st)
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
        "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- host side: TMA descriptors ------------------------------------------------------------------
// ---- host side: TMA descriptors for the weight tiles --------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// 2-D fp32 tensor [rows][cols] (cols contiguous), box = 32 cols x box_rows, SWIZZLE_128B
int make_map_2d(CUtensorMap* map, const float* ptr, long rows, long cols, int box_rows,
                CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
  EncodeTiledFn enc = encode_tiled();
  FD_REQUIRE(enc != nullptr, "conv_tc: cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)cols * sizeof(float)};
  cuuint32_t box[2] = {32u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)ptr, gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  FD_REQUIRE(r == CUDA_SUCCESS, "conv_tc: cuTensorMapEncodeTiled failed (%d) rows=%ld cols=%ld", (int)r,
             rows, cols);
  return 0;
}

// 3-D fp32 tensor [d2][d1][cols] (cols contiguous), box = 32 cols x box_rows x 1, SWIZZLE_128B
int make_map_3d(CUtensorMap* map, const float* ptr, long d2, long d1, long cols, int box_rows) {
  EncodeTiledFn enc = encode_tiled();
  FD_REQUIRE(enc != nullptr, "conv_tc: cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t gdim[3] = {(cuuint64_t)cols, (cuuint64_t)d1, (cuuint64_t)d2};
  cuuint64_t gstr[2] = {(cuuint64_t)cols * sizeof(float), (cuuint64_t)cols * d1 * sizeof(float)};
  cuuint32_t box[3] = {32u, (cuuint32_t)box_rows, 1u};
  cuuint32_t estr[3] = {1u, 1u, 1u};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)ptr, gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  FD_REQUIRE(r == CUDA_SUCCESS, "conv_tc: cuTensorMapEncodeTiled(3d) failed (%d) %ldx%ldx%ld", (int)r, d2, d1, cols);
  return 0;
}

}  // namespace
