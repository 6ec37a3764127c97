// gemm_tcgen05.cu — sm_100a tensor-core GEMM: TMA -> swizzled smem -> tcgen05.mma (BF16 in,
// FP32 accumulate in TMEM) -> tcgen05.ld -> fused epilogue.
//
//   C[M,N] = epilogue( A[M,K] . W[N,K]^T ),  A and W K-major (row-major, K contiguous) BF16.
//
// One 128 x BN output tile per CTA (the hot path's GEMMs are sub-wave at batch 16, so a persistent
// scheduler would buy nothing).  Warp roles (192 threads):
//   warp 0      TMA producer: cp.async.bulk.tensor.2d loads of a 128x64 A box and a BNx64 W box per
//               k-block into a 4-stage ring, completion on mbarriers (complete_tx::bytes)
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer (UMMA 128 x BN x 16, kind::f16),
//               tcgen05.commit releases ring slots / signals the accumulator
//   warps 2..9  epilogue: tcgen05.ld 32x32b (one accumulator row per thread) -> smem staging tile -> row-
//               contiguous bias / mask / gate / activation / scaled residual, FP32 and/or BF16(-split) stores
// Two CTAs are resident per SM (3 stages, <= 102 registers) so that one CTA's epilogue overlaps the other's
// mainloop; the first ncu capture showed the 4-warp, 1-CTA/SM version spending ~3x longer in the epilogue
// than in the k-loop (profiles/r01_gemm_epilogue.md).
//
// Split mode (UNAV_BF16X2 operands): rows hold hi = bf16(x) and lo = bf16(x - hi) at column ld/2; the
// k-loop runs three segments (Ahi,Whi), (Alo,Whi), (Ahi,Wlo) into the same TMEM accumulator, which
// gives ~FP32 products on the BF16 tensor pipe (measured 5e-6 of range on the logits vs 3e-3 for
// plain BF16 operands; SURVEY.md §7 "hard parts").
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"

namespace unav {

constexpr int TC_BM = 128;
constexpr int TC_BK = 64;          // 64 bf16 = 128 bytes = one SWIZZLE_128B span (BK = 32: 64-byte rows, SWIZZLE_64B)
constexpr int TC_STAGES = 3;          // default ring depth: 3 x 32 KB (BN = 128) -> two CTAs per SM, one's epilogue overlaps the other's k-loop
constexpr int TC_MAX_STAGES = 8;      // deep ring for sub-wave grids: bytes in flight per SM = stages x stage size (latency bound)
constexpr int TC_THREADS = 320;       // TMA warp + MMA warp + 8 epilogue warps
constexpr long long TC_SPIN_LIMIT = 4000000000LL;   // ~2 s of SM clocks, then trap instead of hanging

struct TcGroup {
  // 3-D maps {K, rows, halves}: ONE box {BK, tile rows, 1 | 2} brings the hi (and lo) tiles of an operand, laid out in
  // shared memory as [half][row][k] = the stage's [X_hi | X_lo]: two TMA instructions per stage instead of four.
  // (Measured: a CTA's TMA stream runs at ~33 B/clk plus ~100 clocks per box — 8-row boxes are 4x slower than 128-row
  // ones — and two resident CTAs get twice that, so the byte rate per CTA, not the instruction count, bounds the k-loop.)
  CUtensorMap tmA;
  CUtensorMap tmW;
  EpiParams epi;
};
struct TcParams {
  TcGroup g[UNAV_MAX_GROUPS];
  int M, N, K, op_dtype, act, res_masked, nseg, stages, once, ngroups;
  // implicit k=3 convolution (conv_T > 0): A is the PLAIN operand [items*conv_T, Cin] (K = 3*Cin); a tile is 128 time steps
  // of one item and the k-loop fetches tap t from rows t0 + t - 1 through a 4-D map {Cin, T, items, halves}: rows outside
  // [0, T) are zero-filled by the TMA unit = the per-video zero padding of MaskedConv1D (blocks.py:30-31).  No im2col
  // operand is materialised; the accumulation order (K = tap*Cin + c ascending) is that of the im2col GEMM, bit for bit.
  int conv_T, conv_cin, conv_tiles;
  long long* phase;    // diagnostics (unav_gemm_set_phase_trace): 8 clock64 stamps per CTA, or nullptr
  int phase_cap, fine, no_dry;
};



// ---- PTX wrappers -------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// The suspend-time hint lets the hardware park the waiting warp until the phase completes (or the hint expires)
// instead of re-issuing the poll every ~100 clocks: in the first captures the polling of the 8 epilogue warps and the
// producer was 3/4 of all issued warp-instructions and competed with the co-resident CTA's epilogue for issue slots.
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity), "r"(0x10000u) : "memory");
  return ok;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > TC_SPIN_LIMIT) __trap();   // a lost arrival becomes an error, not a hung GPU
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar,
                                            int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_ld_32x32b_x32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tc_ld_32x32b_x16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
//   [0,14) start address >> 4 | [16,30) LBO >> 4 (unused for swizzled K-major, 1) |
//   [32,46) SBO >> 4 = 1024 B (8 rows x 128 B) | [46,48) version = 1 | [61,64) layout = 2 (SW128)
// BK = 32 (64-byte rows): SWIZZLE_64B, SBO = 8 rows x 64 B = 512 B, layout = 4.
template <int BK = 64>
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  static_assert(BK == 64 || BK == 32, "K-major operand rows are 128 or 64 bytes");
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>((8 * BK * 2) >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(BK == 64 ? 2 : 4) << 61;
  return d;
}

// Instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 (1) @4, a/b format BF16 (1) @7/@10,
// a/b K-major (0) @15/@16, N>>3 @17, M>>4 @24.
__host__ __device__ constexpr uint32_t make_idesc(int m, int n, bool f16 = false) {
  return (1u << 4) | ((f16 ? 0u : 1u) << 7) | ((f16 ? 0u : 1u) << 10) | (static_cast<uint32_t>(n >> 3) << 17) |
         (static_cast<uint32_t>(m >> 4) << 24);      // a/b format: 0 = F16, 1 = BF16
}

template <int BN, int BK = TC_BK>
struct TcSmem {
  static constexpr int A_BYTES = TC_BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int total(int stages, int nparts) { return STAGE_BYTES * nparts * stages + 256 + 1024; }   // + barriers + slack
};

// Phase B of the epilogue for full-width, 16-byte aligned tiles: one instantiation per (activation, operand format) so
// the row loop carries no dtype / activation dispatch (the generic loop below executed ~170 instructions per float4).
// Each lane owns 4 consecutive columns of RPP-strided rows; four rows are in flight per lane.
template <int ACT, bool SPLIT, int PITCH, int RPP, int U, int RES, bool FULL>
__device__ __forceinline__ void epilogue_rows_body(const EpiParams& e, const float* stg_lane, long long m_base, int rows,
                                                   int rsub, int n, int res_masked, const float (&bias)[4],
                                                   const float (&cs)[4]) {
  // The row loop is LATENCY bound, not issue bound (per-CTA phase stamps: ~450 clocks per row with two epilogue warps per
  // scheduler, against ~80 issued instructions): so every address is a hoisted pointer that is bumped per batch (the group's
  // parameters live in constant memory behind a run-time index, and re-deriving `base + m * ld + n` per row put an indexed
  // constant load and a 64-bit multiply chain on every row's critical path), and rows are processed U at a time with all
  // loads (staging tile, mask, scale, gate, residual) issued before the first use.
  // FULL: all 32 staged rows of the warp are inside M — no per-row predicate at all (a predicated row is a reconvergence
  // region per row in SASS, which serialises the U rows that are meant to be in flight together); only the last row tile of a
  // ragged M takes the predicated instantiation.
  static_assert((32 / RPP) % U == 0, "a warp's 32 staged rows must be a whole number of U-row batches per lane");
  // RES: -1 = residual decided at run time, 0 / 1 = compiled out / in (the persistent kernel's lean epilogue)
  const bool has_res = RES < 0 ? e.res != nullptr : RES != 0;
  const bool has_gate = e.gate != nullptr, has_mask = e.rowmask != nullptr,
             has_rs = e.rowscale != nullptr, has_f32 = e.out_f32 != nullptr, has_op = e.out_op != nullptr;
  const bool has_rowq = has_gate || has_mask || has_rs;
  const long long m_first = m_base + rsub;
  const long long op_split = e.ld_op / 2;
  const float* res_p = has_res ? e.res + m_first * e.ldres + n : nullptr;
  const long long res_step = static_cast<long long>(RPP) * e.ldres;
  float* f32_p = has_f32 ? e.out_f32 + m_first * e.ld_f32 + n : nullptr;
  const long long f32_step = static_cast<long long>(RPP) * e.ld_f32;
  uint16_t* op_p = has_op ? reinterpret_cast<uint16_t*>(e.out_op) + m_first * e.ld_op + n : nullptr;
  const long long op_step = static_cast<long long>(RPP) * e.ld_op;
  const uint8_t* mk_p = has_mask ? e.rowmask + m_first : nullptr;
  const float* rs_p = has_rs ? e.rowscale + m_first : nullptr;
  // gate_width % 4 == 0 on this path: one gate group per lane
  const float* gate_p = has_gate ? e.gate + m_first * e.gate_groups + n / e.gate_width : nullptr;
  const long long gate_step = static_cast<long long>(RPP) * e.gate_groups;
#pragma unroll 1
  for (int r0 = rsub; r0 < (FULL ? 32 : rows); r0 += U * RPP) {
    float4 a[U], rr[U];
    float mk[U], mrs[U], gt[U];
    bool ok[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      ok[u] = FULL || (r0 + u * RPP < rows);
      a[u] = *reinterpret_cast<const float4*>(stg_lane + (r0 + u * RPP) * PITCH);      // always inside the warp's 32 staged rows
      mk[u] = 1.f; mrs[u] = 1.f; gt[u] = 1.f;
      rr[u] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (has_res) {
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (ok[u]) rr[u] = *reinterpret_cast<const float4*>(res_p + u * res_step);
    }
    if (has_rowq) {       // warp-uniform: one branch around all row-quantity loads of the batch
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (ok[u]) {
          if (has_mask) { mk[u] = mk_p[u * RPP] ? 1.f : 0.f; mrs[u] = mk[u]; }
          if (has_rs) mrs[u] *= __ldg(rs_p + u * RPP);
          if (has_gate) gt[u] = __ldg(gate_p + u * gate_step);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float v[4] = {a[u].x, a[u].y, a[u].z, a[u].w};
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] = (v[i] + bias[i]) * mrs[u];
      if (has_gate) {
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] *= gt[u];
      }
      if (ACT != UNAV_ACT_NONE) {
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = apply_act_tc(v[i], ACT);
      }
      if (has_res) {
        const float rm = res_masked ? mk[u] : 1.f;
        v[0] = rr[u].x * rm + cs[0] * v[0]; v[1] = rr[u].y * rm + cs[1] * v[1];
        v[2] = rr[u].z * rm + cs[2] * v[2]; v[3] = rr[u].w * rm + cs[3] * v[3];
      }
      a[u] = make_float4(v[0], v[1], v[2], v[3]);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (ok[u]) {
        if (has_f32) *reinterpret_cast<float4*>(f32_p + u * f32_step) = a[u];
        if (has_op) store_op4_16<kHalfF16>(op_p + u * op_step, SPLIT, 0, op_split, a[u]);
      }
    }
    if (has_mask) mk_p += U * RPP;
    if (has_rs) rs_p += U * RPP;
    if (has_gate) gate_p += U * gate_step;
    if (has_res) res_p += U * res_step;
    if (has_f32) f32_p += U * f32_step;
    if (has_op) op_p += U * op_step;
  }
}

template <int ACT, bool SPLIT, int PITCH, int RPP, int U, int RES = -1>
__device__ __forceinline__ void epilogue_rows_fast(const EpiParams& e, const float* stg_lane, long long m_base, int rows,
                                                   int rsub, int n, int res_masked, bool f16, const float (&bias)[4],
                                                   const float (&cs)[4]) {
  (void)f16;
  if (rows >= 32) epilogue_rows_body<ACT, SPLIT, PITCH, RPP, U, RES, true>(e, stg_lane, m_base, rows, rsub, n, res_masked, bias, cs);
  else epilogue_rows_body<ACT, SPLIT, PITCH, RPP, U, RES, false>(e, stg_lane, m_base, rows, rsub, n, res_masked, bias, cs);
}

// 1 KB that the DRY epilogue pass reads and writes (see pp_epilogue_pass / epilogue_tile): contents are never used.  Row-indexed vectors (mask,
// row scale) are read at offsets below 128 + 16, column-indexed ones below 32 floats.
__device__ __align__(256) float g_epi_scratch[256];

// Redirect every pointer of an epilogue to the scratch block, with zero strides (dry pass: same instructions, harmless addresses)
__device__ __forceinline__ void epi_make_dry(EpiParams& e) {
  float* sc = g_epi_scratch;
  if (e.bias) e.bias = sc;
  if (e.colscale) e.colscale = sc;
  if (e.rowmask) e.rowmask = reinterpret_cast<const uint8_t*>(sc);
  if (e.rowscale) e.rowscale = sc;
  if (e.gate) { e.gate = sc; e.gate_groups = 0; }
  if (e.res) { e.res = sc; e.ldres = 0; }
  if (e.out_f32) { e.out_f32 = sc; e.ld_f32 = 0; }
  if (e.out_op) { e.out_op = sc; e.ld_op = 0; }
  e.out_opT = nullptr;
}

// Epilogue of one 128 x BN accumulator tile (TMEM columns [tmem_cols, tmem_cols + BN) of this CTA), executed by the 8
// epilogue warps (TMEM lane quarter = warp % 4, column half = (warp - 2) / 4).
// Phase A: tcgen05.ld (one accumulator row per thread) -> padded FP32 staging tile in the (now idle) pipeline smem;
// phase B: each warp walks its 32 rows x BN/2 columns with lanes across columns, so residual loads and all stores are
// row-contiguous 128-bit accesses.
// U = rows in flight per lane in the row loop: 4 where the register budget allows (the persistent kernel owns its SM), 2 in
// the kernels that keep two CTAs resident per SM (96 registers per thread)
// dry = true: instruction-cache warm-up while the epilogue warps wait for the accumulator (see pp_epilogue_pass).  The staging tile
// of this kernel is the pipeline's shared memory, which the k-loop is still using: the dry pass neither reads tensor memory nor
// writes the staging tile (it only reads it), and all its global accesses go to the scratch block.
template <int BN, int U = 2>
__device__ __forceinline__ void epilogue_tile(const TcParams& p, const TcGroup& g, uint32_t tmem_cols, int m0, int n0,
                                              int warp, int lane, float* stg_base, long long m_limit, bool dry = false) {
  constexpr int PITCH = BN + 4;              // floats; 16-byte groups of consecutive rows fall in distinct banks
  constexpr int HALF = BN / 2;               // columns per warp
  constexpr int LPR = HALF / 4;              // lanes per row in phase B (16 | 8)
  constexpr int RPP = 32 / LPR;              // rows per pass (2 | 4)
  const int q = warp & 3, half = (warp - 2) >> 2;
  const uint32_t tmem_base = tmem_cols;
  float* stg = stg_base + q * 32 * PITCH + half * HALF;
#pragma unroll 1
  for (int c = 0; c < (dry ? 0 : HALF / 32); ++c) {
    uint32_t r[32];
    tc_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + half * HALF + c * 32, r);
    tc_wait_ld();
    float* dst = stg + lane * PITCH + c * 32;
#pragma unroll
    for (int j = 0; j < 32; j += 4)
      *reinterpret_cast<float4*>(dst + j) = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                                        __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
  }
  __syncwarp();
  EpiParams e = g.epi;
  const int cl = lane % LPR, rsub = lane / LPR;
  int n = n0 + half * HALF + cl * 4;
  if (dry) {
    epi_make_dry(e);
    n = cl * 4; n0 = 0; m0 = 0; m_limit = 1 << 20;
  }
  const int nvalid = min(4, p.N - n);        // <= 0: this lane's columns are past N
  const long long op_split = e.ld_op / 2;
  const bool has_res = e.res != nullptr, has_gate = e.gate != nullptr;
  const int act = p.act;
  float bias[4] = {0.f, 0.f, 0.f, 0.f}, cs[4] = {1.f, 1.f, 1.f, 1.f};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (i < nvalid) {
      if (e.bias) bias[i] = __ldg(e.bias + n + i);
      if (e.colscale) cs[i] = __ldg(e.colscale + n + i);
    }
  }
  const bool vec_f32 = nvalid == 4 && e.out_f32 && ((reinterpret_cast<uintptr_t>(e.out_f32 + n) & 15) == 0) && (e.ld_f32 % 4 == 0);
  const bool vec_res = nvalid == 4 && has_res && ((reinterpret_cast<uintptr_t>(e.res + n) & 15) == 0) && (e.ldres % 4 == 0);
  const bool vec_op = nvalid == 4 && e.out_op && (((reinterpret_cast<uintptr_t>(e.out_op) + static_cast<size_t>(n) * 2) & 7) == 0) &&
                      (e.ld_op % 4 == 0) && (op_split % 4 == 0);
  if (e.out_opT) {
    // transposed operand output (V^T for the tensor-core attention): lanes run along the token axis, so each
    // store instruction writes 32 consecutive keys of one channel row; one hoisted row pointer bumped per channel
    const long long mt = static_cast<long long>(m0) + q * 32 + lane;
    const int ncols = e.t_ncols > 0 ? e.t_ncols : p.N;
    const int nc0 = n0 + half * HALF - e.t_col0;           // first transposed row of this warp's column block
    const int jend = min(HALF, p.N - (n0 + half * HALF));  // columns of the block inside N
    if (mt < m_limit && nc0 + HALF > 0 && nc0 < ncols) {
      const long long item = mt / e.t_seg, t = mt - item * e.t_seg;
      const long long spl = e.ld_opT / 2;
      uint16_t* row = reinterpret_cast<uint16_t*>(e.out_opT) + (item * ncols + nc0) * e.ld_opT + t;
      const float* bp = e.bias ? e.bias + n0 + half * HALF : nullptr;
      const float* sp = stg + lane * PITCH;
      const bool split_t = op_is_split(p.op_dtype);
#pragma unroll 4
      for (int j = 0; j < jend; ++j, row += e.ld_opT) {
        if (nc0 + j < 0 || nc0 + j >= ncols) continue;
        float x = sp[j];
        if (bp) x += __ldg(bp + j);
        const uint16_t hi = f2h16(x, kHalfF16);
        row[0] = hi;
        if (split_t) row[spl] = f2h16(x - h162f(hi, kHalfF16), kHalfF16);
      }
    }
  }
  // CTA-uniform test for the specialised row loop: full-width tile and every vector access 16-byte aligned
  const bool fast = (n0 + BN <= p.N) && (!e.out_f32 || (((reinterpret_cast<uintptr_t>(e.out_f32) & 15) == 0) && e.ld_f32 % 4 == 0)) &&
                    (!has_res || (((reinterpret_cast<uintptr_t>(e.res) & 15) == 0) && e.ldres % 4 == 0)) &&
                    (!e.out_op || (((reinterpret_cast<uintptr_t>(e.out_op) & 7) == 0) && e.ld_op % 4 == 0 && op_split % 4 == 0)) &&
                    (!has_gate || e.gate_width % 4 == 0);
  if (fast && (e.out_f32 || e.out_op)) {
    const long long m_base = static_cast<long long>(m0) + q * 32;
    const int rows = static_cast<int>(min(32ll, m_limit - m_base));
    const float* sl = stg + cl * 4;
    const bool split = op_is_split(p.op_dtype);
    constexpr bool f16 = kHalfF16;
#define UNAV_EPI_CASE(A)                                                                                          \
    case A:                                                                                                       \
      if (split) epilogue_rows_fast<A, true, PITCH, RPP, U>(e, sl, m_base, rows, rsub, n, p.res_masked, f16, bias, cs);   \
      else epilogue_rows_fast<A, false, PITCH, RPP, U>(e, sl, m_base, rows, rsub, n, p.res_masked, f16, bias, cs);        \
      break;
    switch (act) {
      UNAV_EPI_CASE(UNAV_ACT_RELU)
      UNAV_EPI_CASE(UNAV_ACT_GELU)
      UNAV_EPI_CASE(UNAV_ACT_SILU)
      default:
        if (split) epilogue_rows_fast<UNAV_ACT_NONE, true, PITCH, RPP, U>(e, sl, m_base, rows, rsub, n, p.res_masked, f16, bias, cs);
        else epilogue_rows_fast<UNAV_ACT_NONE, false, PITCH, RPP, U>(e, sl, m_base, rows, rsub, n, p.res_masked, f16, bias, cs);
    }
#undef UNAV_EPI_CASE
  } else if (nvalid > 0 && (e.out_f32 || e.out_op)) {
#pragma unroll 2
    for (int r = rsub; r < 32; r += RPP) {
      const long long m = static_cast<long long>(m0) + q * 32 + r;
      if (m >= m_limit) break;
      const float4 a4 = *reinterpret_cast<const float4*>(stg + r * PITCH + cl * 4);
      float v[4] = {a4.x, a4.y, a4.z, a4.w};
      const float mk = e.rowmask ? (e.rowmask[m] ? 1.f : 0.f) : 1.f;
      const float rs = e.rowscale ? __ldg(e.rowscale + m) : 1.f;
      const float mrs = mk * rs;
      float rr[4] = {0.f, 0.f, 0.f, 0.f};
      if (has_res) {
        const float* rp = e.res + m * e.ldres + n;
        if (vec_res) {
          const float4 t = *reinterpret_cast<const float4*>(rp);
          rr[0] = t.x; rr[1] = t.y; rr[2] = t.z; rr[3] = t.w;
        } else {
          for (int i = 0; i < nvalid; ++i) rr[i] = rp[i];
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] = (v[i] + bias[i]) * mrs;
      if (has_gate) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (i < nvalid) v[i] *= __ldg(e.gate + m * e.gate_groups + (n + i) / e.gate_width);
      }
      if (act != UNAV_ACT_NONE) {
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = apply_act_tc(v[i], act);
      }
      if (has_res) {
        const float rm = p.res_masked ? mk : 1.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = rr[i] * rm + cs[i] * v[i];
      }
      if (e.out_f32) {
        float* o = e.out_f32 + m * e.ld_f32 + n;
        if (vec_f32) {
          *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
        } else {
          for (int i = 0; i < nvalid; ++i) o[i] = v[i];
        }
      }
      if (e.out_op) {
        char* row = reinterpret_cast<char*>(e.out_op) + static_cast<size_t>(m) * e.ld_op * 2;
        if (vec_op) {
          store_op4(row, p.op_dtype, n, op_split, make_float4(v[0], v[1], v[2], v[3]));
        } else {
          for (int i = 0; i < nvalid; ++i) store_op(row, p.op_dtype, n + i, op_split, v[i]);
        }
      }
    }
  }
}

template <int BN, int BK>
__global__ void __launch_bounds__(TC_THREADS, 2)
gemm_tcgen05_kernel(const __grid_constant__ TcParams p) {
  using Sm = TcSmem<BN, BK>;
  extern __shared__ uint8_t smem_raw[];
  const TcGroup& g = p.g[blockIdx.z];
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;     // SWIZZLE_128B needs 1024 B alignment
  const int NS = p.stages;
  const uint32_t bar_base = base + Sm::STAGE_BYTES * ((p.once && p.nseg > 1) ? 2 : 1) * NS;
  // barriers: full[s] at +8s, empty[s] at +8(S+s), accum at +8(2S); tmem slot at +8(2S+1)
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (TC_MAX_STAGES + s); };
  const uint32_t accum_bar = bar_base + 8u * (2 * TC_MAX_STAGES);
  const uint32_t tmem_slot = bar_base + 8u * (2 * TC_MAX_STAGES + 1);

  const bool conv = p.conv_T > 0;
  const int c_item = conv ? blockIdx.x / p.conv_tiles : 0;
  const int c_t0 = conv ? (blockIdx.x % p.conv_tiles) * TC_BM : 0;
  const int m0 = conv ? c_item * p.conv_T + c_t0 : blockIdx.x * TC_BM;                  // first output row of the tile
  const long long m_limit = conv ? static_cast<long long>(c_item) * p.conv_T + min(p.conv_T, c_t0 + TC_BM) : p.M;
  const int n0 = blockIdx.y * BN;
  const int nkb_c = conv ? (p.conv_cin + BK - 1) / BK : 0;                             // k-blocks per tap
  const int nkb = conv ? 3 * nkb_c : (p.K + BK - 1) / BK;
  const int cta_lin = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
  long long* ph_out = (p.phase && cta_lin < p.phase_cap) ? p.phase + 8ll * cta_lin : nullptr;
  if (ph_out && threadIdx.x == 0) {
    uint32_t smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    ph_out[0] = smid;
    ph_out[1] = clock_stamp();
  }

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&g.tmA);
    tma_prefetch_desc(&g.tmW);
    for (int s = 0; s < NS; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(BN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");
  // PDL: everything above (barrier init, TMEM allocation, tensor-map prefetch) overlapped the previous kernel's tail
  pdl_wait();                 // PDL: the preceding grid has completed; nothing above touched global memory
  pdl_launch_dependents();    // let the next kernel's CTAs start their prologue
  if (ph_out && threadIdx.x == 0) ph_out[2] = clock_stamp();     // setup done

  // Stage layout: plain BF16 operands [A | W]; split operands [A_hi | A_lo | W_hi | W_lo] — every operand part is loaded
  // ONCE per k-block and consumed by the three MMA groups (Ahi.Whi, Alo.Whi, Ahi.Wlo).  (An earlier schedule streamed
  // the three segments through [A | W] stages: 1.5x the L2->SMEM bytes, slower everywhere once BK = 32 stages made the
  // combined layout fit twice per SM; it also accumulated in a different order.)
  const int nparts = (p.once && p.nseg > 1) ? 2 : 1;
  const uint32_t stage_bytes = nparts * Sm::STAGE_BYTES;
  const uint32_t w_off = nparts * Sm::A_BYTES;
  const int iters = nkb;
  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      for (int it = 0; it < iters; ++it) {
        const int s = it % NS;
        const uint32_t ph = (it / NS) & 1;
        mbar_wait(empty_bar(s), ph ^ 1);
        mbar_expect_tx(full_bar(s), stage_bytes);
        const uint32_t sa = base + s * stage_bytes;
        if (conv) {
          const int tap = it / nkb_c, kb = it - tap * nkb_c;
          tma_load_4d(sa, &g.tmA, full_bar(s), kb * BK, c_t0 + tap - 1, c_item, 0);
          tma_load_3d(sa + w_off, &g.tmW, full_bar(s), tap * p.conv_cin + kb * BK, n0, 0);
        } else {
          tma_load_3d(sa, &g.tmA, full_bar(s), it * BK, m0, 0);              // [A_hi | A_lo] (or A alone)
          tma_load_3d(sa + w_off, &g.tmW, full_bar(s), it * BK, n0, 0);      // [W_hi | W_lo]
        }
      }
    }
    __syncwarp();     // lanes 1..31 wait here for the producer lane: the block barrier below must see whole warps
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer =====
      const uint32_t idesc = make_idesc(TC_BM, BN, op_is_f16(p.op_dtype));
      const int nseg_in = p.nseg;
      for (int it = 0; it < iters; ++it) {
        const int s = it % NS;
        const uint32_t ph = (it / NS) & 1;
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        if (ph_out && it == 0) ph_out[3] = clock_stamp();             // first operands landed
        const uint32_t sa = base + s * stage_bytes;
        // Canonical accumulation order of the split mode, the same for every tile shape / schedule so that results
        // do not depend on the grid (batch-size independent outputs): K in steps of 32 columns; inside a step
        // hi.hi, lo.hi, hi.lo; inside a segment the two k=16 MMAs in order.
        if (nseg_in > 1) {
#pragma unroll
          for (int h32 = 0; h32 < BK / 32; ++h32) {
#pragma unroll
            for (int seg = 0; seg < 3; ++seg) {
              if (seg >= nseg_in) break;
              const uint64_t adesc = make_smem_desc<BK>(sa + (seg == 1 ? Sm::A_BYTES : 0));
              const uint64_t bdesc = make_smem_desc<BK>(sa + w_off + (seg == 2 ? Sm::B_BYTES : 0));
#pragma unroll
              for (int k = 0; k < 2; ++k) {
                // advance 16 bf16 = 32 bytes inside the swizzle span: +2 in the (addr >> 4) field
                const uint32_t ko = 2u * (h32 * 2 + k);
                tc_mma_f16(tmem_base, adesc + ko, bdesc + ko, idesc, (it > 0 || h32 > 0 || seg > 0 || k > 0) ? 1u : 0u);
              }
            }
          }
        } else {
          const uint64_t adesc = make_smem_desc<BK>(sa);
          const uint64_t bdesc = make_smem_desc<BK>(sa + w_off);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            tc_mma_f16(tmem_base, adesc + 2u * k, bdesc + 2u * k, idesc, (it > 0 || k > 0) ? 1u : 0u);
        }
        tc_commit(empty_bar(s));        // frees the ring slot once these MMAs have read it
      }
      tc_commit(accum_bar);             // accumulator complete
      if (ph_out) ph_out[4] = clock_stamp();                          // last MMA issued
    }
    __syncwarp();     // same for the MMA issuer's warp
  } else {
    // ===== epilogue: warps 2..9 =====  (iteration 0 = dry pass: the epilogue's code is fetched while the k-loop runs)
    float* stg = reinterpret_cast<float*>(smem_raw + (base - smem_u32(smem_raw)));
#pragma unroll 1
    for (int it = p.no_dry ? 1 : 0; it < 2; ++it) {
      const bool dry = it == 0;
      if (!dry) {
        mbar_wait(accum_bar, 0);
        tc_fence_after();
        if (ph_out && threadIdx.x == 64) ph_out[5] = clock_stamp();       // accumulator ready
      }
      if constexpr (BN == 256) {       // experiment (UNAV_TC_BN=256): two 128-column halves through the 128-wide staging tile
        for (int hf = 0; hf < (dry ? 1 : 2); ++hf) {
          if (n0 + hf * 128 >= p.N) break;
          epilogue_tile<128>(p, g, tmem_base + hf * 128, m0, n0 + hf * 128, warp, lane, stg, m_limit, dry);
          __syncwarp();
        }
      } else {
        epilogue_tile<BN>(p, g, tmem_base, m0, n0, warp, lane, stg, m_limit, dry);
      }
    }
  }
  if (ph_out && threadIdx.x == 64) ph_out[6] = clock_stamp();         // this warp's epilogue done
  tc_fence_before();
  __syncthreads();
  if (ph_out && threadIdx.x == 0) ph_out[7] = clock_stamp();          // all warps done
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(BN) : "memory");
  }
}

// =============================================================================================
// CTA-pair variant (cta_group::2): a cluster of two CTAs computes a 256 x 256 tile.  Rank r owns rows
// [256*pair + 128*r, +128) of A and rows [n0 + 128*r, +128) of W; `tcgen05.mma.cta_group::2` (M = 256, N = 256), issued by
// rank 0 only, reads A from each CTA's own shared memory and each half of B once for both tensor cores, so a CTA
// stages 32 KB per 32-wide k-step for 128 x 256 outputs: half the shared-memory traffic per FLOP of the 128 x 128
// kernel above, which is what bounds its k-loop.  Each CTA keeps its 128 x 256 FP32 accumulator in
// its own TMEM (256 columns, two CTAs per SM still fit) and runs the same row epilogue on two 128-column halves.
// Barrier protocol: both producers signal rank 0's full[s] (rank 0 expects the bytes of both CTAs, rank 1 adds a
// remote arrival; TMA completions are routed to rank 0's barrier), the MMA commit is multicast to empty[s] / accum
// of both CTAs.
// =============================================================================================
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(bar), "r"(cta) : "memory");
}
// TMA load whose completion bytes are credited to the barrier at the same offset in cluster rank 0
__device__ __forceinline__ void tma_load_3d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tc_commit_pair(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(static_cast<uint16_t>(3)) : "memory");
}
__device__ __forceinline__ void tc_mma_f16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

constexpr int P2_BK = 32;
constexpr int P2_PART = 128 * P2_BK * 2;     // one 128-row operand part of a stage: 8 KB

// PN = columns of the pair's tile: 256 (each CTA stages 128 rows of W) or 128 (64 rows of W per CTA: the CTA count of the
// 128 x 128 one-CTA grid with 3/4 of its shared-memory fill and 3/4 of its MMA operand reads).
template <int PN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 2)
gemm_tcgen05_pair_kernel(const __grid_constant__ TcParams p) {
  constexpr int W_PART = (PN / 2) * P2_BK * 2;          // this CTA's half of one W part of a stage
  extern __shared__ uint8_t smem_raw[];
  const TcGroup& g = p.g[blockIdx.z];
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const uint32_t rank = cluster_ctarank();
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int NS = p.stages;
  const int nparts = p.nseg > 1 ? 2 : 1;
  const uint32_t stage_bytes = nparts * (P2_PART + W_PART);           // [A parts | W parts]
  const uint32_t w_off = nparts * P2_PART;
  const uint32_t bar_base = base + stage_bytes * NS;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (TC_MAX_STAGES + s); };
  const uint32_t accum_bar = bar_base + 8u * (2 * TC_MAX_STAGES);
  const uint32_t tmem_slot = bar_base + 8u * (2 * TC_MAX_STAGES + 1);

  const int m0 = (blockIdx.x >> 1) * 256 + static_cast<int>(rank) * 128;      // this CTA's accumulator rows
  const int n0 = blockIdx.y * PN;                                             // the pair's columns
  const int wn0 = n0 + static_cast<int>(rank) * (PN / 2);                     // this CTA's half of W
  const int nkb = (p.K + P2_BK - 1) / P2_BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&g.tmA);
    tma_prefetch_desc(&g.tmW);
    for (int s = 0; s < NS; ++s) { mbar_init(full_bar(s), 2); mbar_init(empty_bar(s), 1); }
    mbar_init(accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(PN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();            // both CTAs' barriers are initialised before any remote arrival / multicast commit
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");
  pdl_wait();
  pdl_launch_dependents();

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer (both CTAs) =====
      for (int it = 0; it < nkb; ++it) {
        const int s = it % NS;
        const uint32_t ph = (it / NS) & 1;
        mbar_wait(empty_bar(s), ph ^ 1);
        if (rank == 0) mbar_expect_tx(full_bar(s), 2u * stage_bytes);       // bytes of both CTAs land on this barrier
        else mbar_arrive_remote(full_bar(s), 0);
        const uint32_t sa = base + s * stage_bytes;
        tma_load_3d_pair(sa, &g.tmA, full_bar(s), it * P2_BK, m0, 0);
        tma_load_3d_pair(sa + w_off, &g.tmW, full_bar(s), it * P2_BK, wn0, 0);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      // ===== MMA issuer (rank 0 only) =====
      const uint32_t idesc = make_idesc(256, PN, kHalfF16);
      for (int it = 0; it < nkb; ++it) {
        const int s = it % NS;
        const uint32_t ph = (it / NS) & 1;
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        const uint32_t sa = base + s * stage_bytes;
        // same canonical order as the 1-CTA kernel: per 32-wide k-step hi.hi, lo.hi, hi.lo
#pragma unroll
        for (int seg = 0; seg < 3; ++seg) {
          if (seg >= p.nseg) break;
          const uint64_t adesc = make_smem_desc<P2_BK>(sa + (seg == 1 ? P2_PART : 0));
          const uint64_t bdesc = make_smem_desc<P2_BK>(sa + w_off + (seg == 2 ? W_PART : 0));
#pragma unroll
          for (int k = 0; k < 2; ++k)
            tc_mma_f16_pair(tmem_base, adesc + 2u * k, bdesc + 2u * k, idesc, (it > 0 || seg > 0 || k > 0) ? 1u : 0u);
        }
        tc_commit_pair(empty_bar(s));
      }
      tc_commit_pair(accum_bar);
    }
    __syncwarp();
  } else {
    // ===== epilogue (both CTAs): the 128-column halves of this CTA's 128 x PN accumulator =====
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    float* stg = reinterpret_cast<float*>(smem_raw + (base - smem_u32(smem_raw)));
    for (int hf = 0; hf < PN / 128; ++hf) {
      if (n0 + hf * 128 >= p.N) break;
      epilogue_tile<128>(p, g, tmem_base + hf * 128, m0, n0 + hf * 128, warp, lane, stg, p.M);
      __syncwarp();
    }
  }
  tc_fence_before();
  cluster_sync_all();            // the peer may still read this CTA's smem / signal its barriers until both are done
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(PN) : "memory");
  }
}

// Lean epilogue pass of the persistent kernel: one 128 x 64 column block of a TMEM accumulator.  The host has already
// established what epilogue_tile checks per call (full-width tiles, 16-byte aligned outputs / residual / bias / column scale,
// gate_width % 4 == 0), so there is no generic fallback and the per-pass set-up is a few dozen instructions (ncu source page of
// the first version: 359 of the 1 053 instructions a warp executed per pass were set-up).  The bias / column-scale loads are
// issued before the tensor-memory read so their latency overlaps it.
// dry = true: instruction-cache warm-up.  The epilogue warps of a persistent CTA idle through the first tile's k-loop (13 - 18 k
// clocks) and then execute the row loop for the first time: the fine trace showed ~10 k clocks for that first 128 x 64 pass
// against 3.2 k for every later one — the loop's ~3 - 5 KB of code arriving line by line from L2.  A dry pass runs the SAME call
// site (hence the same inlined instructions) while the warps would otherwise wait: every pointer of the epilogue is redirected to
// a 1 KB scratch block with zero strides, the tensor-memory read and the transposed store are skipped.
template <int U, int EW>
__device__ __forceinline__ void pp_epilogue_pass(const TcParams& p, const TcGroup& g, uint32_t tmem_cols, int m0, int n0, int warp,
                                                 int lane, float* stg_base, long long m_limit, long long* fine = nullptr,
                                                 bool dry = false) {
  if (fine) fine[0] = clock_stamp();            // diagnostics (UNAV_PP_FINE=1 + unav_set_phase_trace): inside one pass
  // EW epilogue warps share a 64-column pass: TMEM lane quarter = warp % 4 (a warp can only read its own quarter), column
  // block = (warp - 2) / 4 of CW = 32 (EW = 8) or 16 (EW = 16) columns.  Sixteen warps: the pass is latency bound (tensor-memory
  // read -> staging -> residual load -> stores is one dependent chain per warp), so twice the warps is twice the chains in flight.
  constexpr int BN = 64, PITCH = BN + 4, CW = BN / (EW / 4), LPR = CW / 4, RPP = 32 / LPR;
  static_assert(EW == 8 || EW == 16, "8 or 16 epilogue warps");
  const int q = warp & 3, cb = (warp - 2) >> 2;
  EpiParams e = g.epi;
  float* stg = stg_base + q * 32 * PITCH + cb * CW;
  const int cl = lane % LPR, rsub = lane / LPR;
  int n = n0 + cb * CW + cl * 4;
  if (dry) {            // same code, harmless addresses
    epi_make_dry(e);
    n = cl * 4;
    m0 = 0;
    m_limit = 1 << 20;
  }
  float bias[4] = {0.f, 0.f, 0.f, 0.f}, cs[4] = {1.f, 1.f, 1.f, 1.f};
  if (e.bias) {
    const float4 b4 = __ldg(reinterpret_cast<const float4*>(e.bias + n));
    bias[0] = b4.x; bias[1] = b4.y; bias[2] = b4.z; bias[3] = b4.w;
  }
  if (e.colscale) {
    const float4 c4 = __ldg(reinterpret_cast<const float4*>(e.colscale + n));
    cs[0] = c4.x; cs[1] = c4.y; cs[2] = c4.z; cs[3] = c4.w;
  }
  {
    uint32_t r[CW];
    const uint32_t taddr = tmem_cols + (static_cast<uint32_t>(q * 32) << 16) + cb * CW;
    if (!dry) {
      if constexpr (CW == 32) tc_ld_32x32b_x32(taddr, r);
      else tc_ld_32x32b_x16(taddr, r);
      tc_wait_ld();
    } else {
#pragma unroll
      for (int j = 0; j < CW; ++j) r[j] = 0u;
    }
    if (fine) fine[1] = clock_stamp();
    float* dst = stg + lane * PITCH;
#pragma unroll
    for (int j = 0; j < CW; j += 4)
      *reinterpret_cast<float4*>(dst + j) = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                                        __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
  }
  __syncwarp();
  if (fine) fine[2] = clock_stamp();
  if (e.out_opT) {       // transposed operand output (V^T for the tensor-core attention): lanes run along the token axis, so
    // each store instruction writes 32 consecutive keys of one channel row.  Addresses are one hoisted row pointer bumped per
    // channel (the first version re-derived item * ncols * ld per element with 64-bit multiplies).
    const long long mt = static_cast<long long>(m0) + q * 32 + lane;
    const int ncols = e.t_ncols > 0 ? e.t_ncols : p.N;
    const int nc0 = n0 + cb * CW - e.t_col0;               // first transposed row of this warp's column block
    if (mt < m_limit && nc0 + CW > 0 && nc0 < ncols) {
      const long long item = mt / e.t_seg, t = mt - item * e.t_seg;
      const long long spl = e.ld_opT / 2;
      uint16_t* row = reinterpret_cast<uint16_t*>(e.out_opT) + (item * ncols + nc0) * e.ld_opT + t;
      const float* bp = e.bias ? e.bias + n0 + cb * CW : nullptr;
      const float* sp = stg + lane * PITCH;
      const bool split_t = op_is_split(p.op_dtype);
#pragma unroll 4
      for (int j = 0; j < CW; ++j, row += e.ld_opT) {
        if (nc0 + j < 0 || nc0 + j >= ncols) continue;
        float x = sp[j];
        if (bp) x += __ldg(bp + j);
        const uint16_t hi = f2h16(x, kHalfF16);
        row[0] = hi;
        if (split_t) row[spl] = f2h16(x - h162f(hi, kHalfF16), kHalfF16);
      }
    }
  }
  if (!(e.out_f32 || e.out_op)) return;
  const long long m_base = static_cast<long long>(m0) + q * 32;
  const int rows = static_cast<int>(min(32ll, m_limit - m_base));
  const float* sl = stg + cl * 4;
  const bool split = op_is_split(p.op_dtype);
  constexpr bool f16 = kHalfF16;
#define UNAV_PP_CASE(A)                                                                                                        \
  case A:                                                                                                                      \
    if (e.res) {                                                                                                               \
      if (split) epilogue_rows_fast<A, true, PITCH, RPP, U, 1>(e, sl, m_base, rows, rsub, n, p.res_masked, f16, bias, cs);      \
      else epilogue_rows_fast<A, false, PITCH, RPP, U, 1>(e, sl, m_base, rows, rsub, n, p.res_masked, f16, bias, cs);           \
    } else {                                                                                                                   \
      if (split) epilogue_rows_fast<A, true, PITCH, RPP, U, 0>(e, sl, m_base, rows, rsub, n, p.res_masked, f16, bias, cs);      \
      else epilogue_rows_fast<A, false, PITCH, RPP, U, 0>(e, sl, m_base, rows, rsub, n, p.res_masked, f16, bias, cs);           \
    }                                                                                                                          \
    break;
  switch (p.act) {
    UNAV_PP_CASE(UNAV_ACT_RELU)
    UNAV_PP_CASE(UNAV_ACT_GELU)
    UNAV_PP_CASE(UNAV_ACT_SILU)
    UNAV_PP_CASE(UNAV_ACT_NONE)
    default: break;
  }
#undef UNAV_PP_CASE
  if (fine) fine[3] = clock_stamp();
}

// =============================================================================================
// Persistent CTA-pair variant: ONE cluster of two CTAs per SM pair (74 clusters on a B200) walks a static list of 256 x 256
// output tiles.  Same operand staging, MMA shape (cta_group::2, M = 256, N = 256) and accumulation order as the kernel above,
// so the results are bit-identical; what changes is everything AROUND the k-loop, which is what bounds the K = 512 shapes of the
// hot path (a 128 x 256 CTA-tile there spends ~6 us in its k-loop and ~12 us in TMEM allocation, barrier set-up, the first L2
// round trip and the epilogue, and the 1.1 - 1.6 wave grids of one-tile CTAs leave SMs idle in the last wave):
//   * TMEM holds TWO 128 x 256 FP32 accumulators per CTA (all 512 columns): the epilogue warps drain tile i (tcgen05.ld ->
//     private 64-column staging tiles -> fused row epilogue) while the MMA warp already accumulates tile i+1 into the other half;
//   * the TMA ring never drains between tiles (the producer runs ahead into the next tile's k-blocks), so set-up, allocation
//     and the first L2 round trip are paid once per CTA, not once per tile;
//   * 5 x 32 KB stages (160 KB in flight per SM) + a dedicated 34 KB staging tile: the CTA owns the SM (> half of its shared
//     memory and all of its tensor memory), so two cluster kernels of different streams can never co-reside on an SM and wait
//     for each other's tensor memory (the hold-and-wait cycle that CTA pairs with partial allocations can form across streams).
// Barriers: full[s] / empty[s] as above; tfull[a] (accumulator a complete: commit multicast to both CTAs); tempty[a] lives in
// rank 0's shared memory and collects one arrival per CTA once its eight epilogue warps have read accumulator a.
// =============================================================================================
constexpr int PP_STG_BN = 64;                                   // epilogue staging width (columns per pass)
constexpr int PP_STG_BYTES = 128 * (PP_STG_BN + 4) * 4;         // 34 816
constexpr int PP_MAX_SMEM = 227 * 1024;

template <int PN, int EW>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(64 + 32 * EW, 1)
gemm_tcgen05_ppair_kernel(const __grid_constant__ TcParams p) {
  static_assert(PN == 256, "two 128 x PN accumulators must fill the 512 tensor-memory columns");
  constexpr int W_PART = (PN / 2) * P2_BK * 2;
  // 1024-byte aligned by declaration (SWIZZLE_64B / 128B stages): no alignment slack, which is what lets a sixth 32 KB stage fit
  extern __shared__ __align__(1024) uint8_t pp_smem[];
  uint8_t* const smem_raw = pp_smem;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const uint32_t rank = cluster_ctarank();
  const uint32_t base = smem_u32(smem_raw);
  if ((base & 1023u) != 0) __trap();
  const int NS = p.stages;
  const int nparts = p.nseg > 1 ? 2 : 1;
  const uint32_t stage_bytes = nparts * (P2_PART + W_PART);
  const uint32_t w_off = nparts * P2_PART;
  const uint32_t stg_addr = base + stage_bytes * NS;
  const uint32_t bar_base = stg_addr + PP_STG_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (TC_MAX_STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * TC_MAX_STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * TC_MAX_STAGES + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * TC_MAX_STAGES + 4);

  const int pair_id = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  // diagnostics (unav_set_phase_trace): 32 int64 per CTA = 4 slots of 8: slot 0 {smid, start, setup done, exit}, slots 1..3 = this
  // CTA's first three tiles {producer: first load issued, MMA: accumulator free, first operands landed, last MMA issued,
  // epilogue: accumulator ready, first 64-column pass done, all passes done, accumulator released}
  long long* ph_out = (p.phase && 4 * static_cast<int>(blockIdx.x) + 3 < p.phase_cap) ? p.phase + 32ll * blockIdx.x : nullptr;
  if (ph_out && threadIdx.x == 0) {
    uint32_t smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    ph_out[0] = smid;
    ph_out[1] = clock_stamp();
  }
  const int tiles_m = (p.M + 255) / 256, tiles_n = p.N / PN;
  const int tiles_per_group = tiles_m * tiles_n;
  const int ntiles = tiles_per_group * p.ngroups;
  const int nkb = (p.K + P2_BK - 1) / P2_BK;

  if (warp == 0 && lane == 0) {
    for (int g = 0; g < p.ngroups; ++g) { tma_prefetch_desc(&p.g[g].tmA); tma_prefetch_desc(&p.g[g].tmW); }
    for (int s = 0; s < NS; ++s) { mbar_init(full_bar(s), 2); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 2); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(2 * PN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");
  pdl_wait();
  pdl_launch_dependents();
  if (ph_out && threadIdx.x == 0) ph_out[2] = clock_stamp();

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer (both CTAs): k-blocks of this pair's tiles, back to back =====
      uint32_t it = 0, ti = 0;
      for (int tile = pair_id; tile < ntiles; tile += npairs, ++ti) {
        const int g = tile / tiles_per_group, r = tile - g * tiles_per_group;
        const int n_t = r / tiles_m, m_t = r - n_t * tiles_m;
        const int m0 = m_t * 256 + static_cast<int>(rank) * 128;
        const int wn0 = n_t * PN + static_cast<int>(rank) * (PN / 2);
        const TcGroup& grp = p.g[g];
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % NS;
          const uint32_t ph = (it / NS) & 1;
          mbar_wait(empty_bar(s), ph ^ 1);
          if (ph_out && kb == 0 && ti < (p.fine ? 2u : 3u)) ph_out[8 * (ti + 1) + 0] = clock_stamp();
          if (rank == 0) mbar_expect_tx(full_bar(s), 2u * stage_bytes);
          else mbar_arrive_remote(full_bar(s), 0);
          const uint32_t sa = base + s * stage_bytes;
          tma_load_3d_pair(sa, &grp.tmA, full_bar(s), kb * P2_BK, m0, 0);
          tma_load_3d_pair(sa + w_off, &grp.tmW, full_bar(s), kb * P2_BK, wn0, 0);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      // ===== MMA issuer (rank 0 only) =====
      const uint32_t idesc = make_idesc(256, PN, kHalfF16);
      uint32_t it = 0, i = 0;
      for (int tile = pair_id; tile < ntiles; tile += npairs, ++i) {
        const uint32_t a = i & 1, aph = (i >> 1) & 1;
        mbar_wait(tempty_bar(a), aph ^ 1);        // both CTAs' epilogues have drained this accumulator (first use: free)
        tc_fence_after();
        if (ph_out && i < (p.fine ? 2u : 3u)) ph_out[8 * (i + 1) + 1] = clock_stamp();
        const uint32_t acc = tmem_base + a * PN;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % NS;
          const uint32_t ph = (it / NS) & 1;
          mbar_wait(full_bar(s), ph);
          tc_fence_after();
          if (ph_out && kb == 0 && i < (p.fine ? 2u : 3u)) ph_out[8 * (i + 1) + 2] = clock_stamp();
          const uint32_t sa = base + s * stage_bytes;
          // canonical order: per 32-wide k-step hi.hi, lo.hi, hi.lo
#pragma unroll
          for (int seg = 0; seg < 3; ++seg) {
            if (seg >= p.nseg) break;
            const uint64_t adesc = make_smem_desc<P2_BK>(sa + (seg == 1 ? P2_PART : 0));
            const uint64_t bdesc = make_smem_desc<P2_BK>(sa + w_off + (seg == 2 ? W_PART : 0));
#pragma unroll
            for (int k = 0; k < 2; ++k)
              tc_mma_f16_pair(acc, adesc + 2u * k, bdesc + 2u * k, idesc, (kb > 0 || seg > 0 || k > 0) ? 1u : 0u);
          }
          tc_commit_pair(empty_bar(s));
        }
        tc_commit_pair(tfull_bar(a));
        if (ph_out && i < (p.fine ? 2u : 3u)) ph_out[8 * (i + 1) + 3] = clock_stamp();
      }
    }
    __syncwarp();
  } else {
    // ===== epilogue (both CTAs): four 64-column passes over this CTA's 128 x PN accumulator =====
    float* stg = reinterpret_cast<float*>(smem_raw + (stg_addr - smem_u32(smem_raw)));
    uint32_t i = 0;
    bool dry = !p.no_dry && pair_id < ntiles;       // one dry pass first (instruction-cache warm-up, see pp_epilogue_pass)
    for (int tile = pair_id; tile < ntiles;) {
      const int g = tile / tiles_per_group, r = tile - g * tiles_per_group;
      const int n_t = r / tiles_m, m_t = r - n_t * tiles_m;
      const int m0 = m_t * 256 + static_cast<int>(rank) * 128;
      const int n0 = n_t * PN;
      const uint32_t a = i & 1, aph = (i >> 1) & 1;
      if (!dry) {
        mbar_wait(tfull_bar(a), aph);
        tc_fence_after();
      }
      const bool stamp = !dry && ph_out && threadIdx.x == 64 && i < (p.fine ? 2 : 3);
      if (stamp) ph_out[8 * (i + 1) + 4] = clock_stamp();
#pragma unroll 1
      for (int c = 0; c < (dry ? 1 : PN / PP_STG_BN); ++c) {
        // fine trace: the second pass of the CTA's second tile, in the (otherwise unused) epilogue slots of the third tile's row
        long long* fine = nullptr;
        if (!dry && p.fine && ph_out && threadIdx.x == 64) {
          if (i == 0 && c == 0) fine = ph_out + 24;             // the very first pass of the CTA
          else if (i == 1 && c == 1) fine = ph_out + 28;
        }
        pp_epilogue_pass<4, EW>(p, p.g[g], tmem_base + a * PN + c * PP_STG_BN, m0, n0 + c * PP_STG_BN, warp, lane, stg, p.M, fine, dry);
        __syncwarp();
        if (stamp && c == 0) ph_out[8 * (i + 1) + 5] = clock_stamp();
      }
      if (dry) { dry = false; continue; }             // now the same tile for real
      if (stamp) ph_out[8 * (i + 1) + 6] = clock_stamp();
      tc_fence_before();
      asm volatile("bar.sync 1, %0;" ::"n"(32 * EW) : "memory");      // all epilogue warps have read accumulator a
      if (threadIdx.x == 64) mbar_arrive_remote(tempty_bar(a), 0);
      if (stamp) ph_out[8 * (i + 1) + 7] = clock_stamp();
      tile += npairs; ++i;
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (ph_out && threadIdx.x == 0) ph_out[3] = clock_stamp();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(2 * PN) : "memory");
  }
}

// ---- host side -------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* sym = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t err = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres);
  if (err != cudaSuccess || qres != cudaDriverEntryPointSuccess || sym == nullptr) {
    set_error("cuTensorMapEncodeTiled not available: %s", cudaGetErrorString(err));
    return nullptr;
  }
  fn = reinterpret_cast<EncodeTiledFn>(sym);
  return fn;
}

// 3-D K-major 16-bit tensor map over an operand buffer: dims {K, rows, halves} with strides {ld, ld/2} elements (halves =
// 2 for the split formats: hi at column c, lo at column ld/2 + c), box {bk, box_rows, box_halves}, SWIZZLE_128B (bk = 64)
// or SWIZZLE_64B (bk = 32); out-of-bounds elements read as zero (ragged M / N / K tails need no special casing).
static int encode_map(CUtensorMap* map, const void* ptr, long long rows, long long K, long long ld, int box_rows, int bk,
                      bool split, int box_halves) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return UNAV_ERR_DRIVER;
  cuuint64_t dims[3] = {static_cast<cuuint64_t>(K), static_cast<cuuint64_t>(rows), static_cast<cuuint64_t>(split ? 2 : 1)};
  cuuint64_t strides[2] = {static_cast<cuuint64_t>(ld) * 2, static_cast<cuuint64_t>(split ? ld / 2 : ld) * 2};
  cuuint32_t box[3] = {static_cast<cuuint32_t>(bk), static_cast<cuuint32_t>(box_rows), static_cast<cuuint32_t>(box_halves)};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, bk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): ptr=%p rows=%lld K=%lld ld=%lld", (int)r, ptr, rows, K, ld);
    return UNAV_ERR_DRIVER;
  }
  return 0;
}

// 4-D map of a convolution input [segments*T, Cin] (plain operand): dims {Cin, T, segments, halves}; a box is 128 time
// steps of one segment, and rows before / after the segment come back as zeros.
static int encode_conv_map(CUtensorMap* map, const void* ptr, long long segs, long long T, long long cin, long long ld, int bk,
                           bool split, int box_halves) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return UNAV_ERR_DRIVER;
  cuuint64_t dims[4] = {static_cast<cuuint64_t>(cin), static_cast<cuuint64_t>(T), static_cast<cuuint64_t>(segs),
                        static_cast<cuuint64_t>(split ? 2 : 1)};
  cuuint64_t strides[3] = {static_cast<cuuint64_t>(ld) * 2, static_cast<cuuint64_t>(T) * ld * 2,
                           static_cast<cuuint64_t>(split ? ld / 2 : ld) * 2};
  cuuint32_t box[4] = {static_cast<cuuint32_t>(bk), static_cast<cuuint32_t>(TC_BM), 1u, static_cast<cuuint32_t>(box_halves)};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, bk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (conv) failed (%d): ptr=%p segs=%lld T=%lld cin=%lld ld=%lld", (int)r, ptr, segs, T, cin, ld);
    return UNAV_ERR_DRIVER;
  }
  return 0;
}

constexpr int TC_SLOTS = 296;      // 2 resident CTAs x 148 SMs

// Tile / schedule choice.  All choices accumulate every output element in the same order (k ascending; hi.hi, lo.hi,
// hi.lo inside a k-step of 16), so the result does not depend on it — only the time does (scripts/gemm_probe.py):
//   bn     128 unless the 128-wide grid leaves most SMs idle
//   sched  0 = plain BF16 operands, [A|W] stages of BK=64
//          1 = split operands, BK=64: [A_hi|A_lo|W_hi|W_lo] per stage, 64 KB stages (BN=128): the ring only fits one
//              CTA per SM, so it is for grids of at most one CTA per SM
//          2 = split operands, BK=32 (SWIZZLE_64B): 32 KB stages, two CTAs stay resident per SM
struct TcChoice { int bn, sched; };
static TcChoice choose_tile(int M, int N, int K, int ngroups, int nseg) {
  const long long mt = (M + TC_BM - 1) / TC_BM;
  const long long tiles128 = mt * ((N + 127) / 128) * ngroups;
  const long long tiles64 = mt * ((N + 63) / 64) * ngroups;
  // measured with scripts/gemm_probe.py over the 27 shapes of the batch-16 path
  TcChoice c;
  if (nseg == 1) {
    c.bn = (N <= 64 || tiles128 < 240) ? 64 : 128;
    c.sched = 0;
  } else if (N > 64 && tiles128 > 148) {
    c.bn = 128; c.sched = 2;        // multi-CTA-per-SM grids: 32 KB stages keep two CTAs resident
  } else if (N > 64 && tiles128 >= 100) {
    c.bn = 128; c.sched = 1;        // about one CTA per SM: long 64 KB stages, fewer barrier round trips
  } else {
    c.bn = 64; c.sched = 1;         // small grids: more CTAs, 48 KB stages
  }
  (void)tiles64;
  if (const char* env = getenv("UNAV_TC_BN")) {            // experiment knobs (scripts/gemm_probe.py)
    const int v = atoi(env);
    if (v == 64 || v == 128) c.bn = v;
    // 128 x 256 tiles (4/3 of the FLOP per L2 byte of 128 x 128): NOT validated on a GPU yet — written
    // at the end of round 1 without GPU time left; tests/test_gpu_gemm.py::test_tcgen05_bn256_experiment is skipped unless
    // UNAV_TEST_EXPERIMENTAL=1.  Only for full grids of split operands whose N is a multiple of 256.
    if (v == 256 && nseg > 1 && N % 256 == 0 && tiles128 > 148) { c.bn = 256; c.sched = 2; }
  }
  if (const char* env = getenv("UNAV_TC_ONCE")) {
    const int v = atoi(env);
    if (v >= 1 && v <= 2) c.sched = nseg == 1 ? 0 : v;
  }
  return c;
}

template <int BN, int BK>
static int launch_tc(TcParams& p, int ngroups, cudaStream_t stream) {
  static SmemAttr attr = {};
  constexpr int MAX_SMEM = 200 * 1024;
  using Sm = TcSmem<BN, BK>;
  if (int rc = ensure_dyn_smem(gemm_tcgen05_kernel<BN, BK>, attr, MAX_SMEM, "gemm_tcgen05")) return rc;
  dim3 grid(p.conv_T > 0 ? (p.M / p.conv_T) * p.conv_tiles : (p.M + TC_BM - 1) / TC_BM, (p.N + BN - 1) / BN, ngroups);
  // Ring depth: multi-CTA-per-SM grids keep <= ~100 KB per CTA so two CTAs share an SM (one's epilogue overlaps the
  // other's k-loop); grids of at most one CTA per SM take a deeper ring.  The epilogue staging tile (128 x (BN+4)
  // floats) must also fit in the ring's bytes.
  const long long ctas = static_cast<long long>(grid.x) * grid.y * grid.z;
  const int nparts = (p.once && p.nseg > 1) ? 2 : 1;
  const int nkb = (p.K + BK - 1) / BK;
  const int per_stage = Sm::STAGE_BYTES * nparts;
  int small_ring = 192 * 1024;
  if (const char* env = getenv("UNAV_TC_SMALL_RING_KB")) {     // experiment knob: ring bytes of sub-wave grids (co-residency vs depth)
    const int v = atoi(env);
    if (v >= 32 && v <= 192) small_ring = v * 1024;
  }
  int stages = (ctas <= 148 ? small_ring : 100 * 1024) / per_stage;
  if (const char* env = getenv("UNAV_TC_STAGES")) {            // experiment knob (scripts/gemm_probe.py)
    const int v = atoi(env);
    if (v >= 2) stages = v;
  }
  if (stages > TC_MAX_STAGES) stages = TC_MAX_STAGES;
  if (stages > nkb) stages = nkb;
  if (stages < 2) stages = 2;
  while (stages * per_stage < 128 * ((BN > 128 ? 128 : BN) + 4) * 4) ++stages;      // room for the staging tile
  while (Sm::total(stages, nparts) > MAX_SMEM) --stages;
  p.stages = stages;
  launch_pdl(gemm_tcgen05_kernel<BN, BK>, dim3(grid), dim3(TC_THREADS), Sm::total(stages, nparts), stream, p);
  count_launch();
  return finish_launch("gemm_tcgen05");
}

template <int PN>
static int launch_pair(TcParams& p, int ngroups, cudaStream_t stream) {
  static SmemAttr attr = {};
  constexpr int MAX_SMEM = 110 * 1024;
  if (int rc = ensure_dyn_smem(gemm_tcgen05_pair_kernel<PN>, attr, MAX_SMEM, "gemm_tcgen05_pair")) return rc;
  const int nparts = p.nseg > 1 ? 2 : 1;
  const int per_stage = nparts * (P2_PART + (PN / 2) * P2_BK * 2);
  const int nkb = (p.K + P2_BK - 1) / P2_BK;
  int stages = (100 * 1024) / per_stage;
  if (stages > TC_MAX_STAGES) stages = TC_MAX_STAGES;
  if (stages > nkb) stages = nkb;
  if (stages < 2) stages = 2;
  while (stages * per_stage < 128 * (128 + 4) * 4) ++stages;      // room for the epilogue staging tile
  p.stages = stages;
  p.once = p.nseg > 1 ? 1 : 0;
  dim3 grid(2 * ((p.M + 255) / 256), (p.N + PN - 1) / PN, ngroups);
  launch_pdl(gemm_tcgen05_pair_kernel<PN>, dim3(grid), dim3(TC_THREADS), stages * per_stage + 256 + 1024, stream, p);
  count_launch();
  return finish_launch("gemm_tcgen05_pair");
}

// CTA pairs where they pay (scripts/gemm_probe.py, UNAV_TC_PAIR=0/1): grids of at least two 128 x 128 CTAs per SM with a
// long k-loop — [7056,1024,3072] 134 -> 108 us (410 TFLOP/s algorithmic, 1230 executed), 2x[7056,512,1536] 77 -> 65 us.  The
// K = 512 shapes are epilogue-bound and gain nothing; grids of ~1.5 CTAs per SM lose the second resident CTA's overlap.
// Returns the pair tile width: 0 (one-CTA kernels), 256 or 128.
static int use_pair(int M, int N, int K, int ngroups) {
  int v = -1;
  if (const char* env = getenv("UNAV_TC_PAIR")) v = atoi(env);       // experiment knob: 0 never, 1 256-wide whenever possible,
  if (M < 256 || v <= 0) return 0;                                   // 2 also 128-wide pairs for the other full grids; unset = NEVER:
  // round 2: no policy selects a one-tile pair kernel (the persistent pair kernel serves every shape they served, and a pair
  // CTA with a partial tensor-memory allocation next to other streams' kernels is the hazard of DESIGN.md section 11)
  const long long tiles128 = static_cast<long long>((M + 127) / 128) * ((N + 127) / 128) * ngroups;
  if (N % 256 == 0 && (v == 1 || (tiles128 >= 296 && K >= 1024))) return 256;
  if (v == 2 && N % 128 == 0 && tiles128 >= 148) return 128;
  return 0;
}

// Persistent CTA-pair launch: one cluster per SM pair (at most as many as the device can hold at once), 6 x 32 KB stages.
// EW = epilogue warps per CTA: 8 (UNAV_PP_EW=16 selects sixteen for A/B runs: measured SLOWER on every shape of the path, e.g.
// fc1 + GELU 2x[3600,2048,512] 59.4 -> 71.7 us, 1x[7200,1536,512] 47.0 -> 57.3 us — the epilogue competes with the TMA fill and the
// MMA operand reads for shared-memory bandwidth, not for latency-hiding warps).
template <int EW>
static int launch_ppair_ew(TcParams& p, int ngroups, cudaStream_t stream) {
  static SmemAttr attr = {};
  static int max_clusters[kMaxDevices] = {};
  constexpr int THREADS = 64 + 32 * EW;
  if (int rc = ensure_dyn_smem(gemm_tcgen05_ppair_kernel<256, EW>, attr, PP_MAX_SMEM, "gemm_tcgen05_ppair")) return rc;
  const int nparts = p.nseg > 1 ? 2 : 1;
  const int per_stage = nparts * (P2_PART + 128 * P2_BK * 2);
  const int nkb = (p.K + P2_BK - 1) / P2_BK;
  int stages = (PP_MAX_SMEM - PP_STG_BYTES - 256) / per_stage;
  if (const char* env = getenv("UNAV_TC_STAGES")) {
    const int v = atoi(env);
    if (v >= 2 && v < stages) stages = v;
  }
  if (stages > TC_MAX_STAGES) stages = TC_MAX_STAGES;
  const long long tiles = static_cast<long long>((p.M + 255) / 256) * (p.N / 256) * ngroups;
  if (stages > nkb * 2 && nkb * 2 >= 2) stages = nkb * 2;       // never more ring than two tiles' worth of k-blocks
  if (stages < 2) stages = 2;
  p.stages = stages;
  p.once = p.nseg > 1 ? 1 : 0;
  p.ngroups = ngroups;
  const size_t smem = static_cast<size_t>(stages) * per_stage + PP_STG_BYTES + 256;
  int dev = 0;
  cudaGetDevice(&dev);
  int cap = (dev >= 0 && dev < kMaxDevices) ? max_clusters[dev] : 0;
  if (cap == 0) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * 74); cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = PP_MAX_SMEM;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, gemm_tcgen05_ppair_kernel<256, EW>, &cfg) != cudaSuccess || n <= 0) {
      cudaGetLastError();
      n = 74;
    }
    cap = n;
    if (dev >= 0 && dev < kMaxDevices) max_clusters[dev] = cap;
  }
  if (const char* env = getenv("UNAV_TC_PPAIR_CLUSTERS")) {       // experiment knob
    const int v = atoi(env);
    if (v >= 1 && v < cap) cap = v;
  }
  const int npairs = static_cast<int>(tiles < cap ? tiles : cap);
  launch_pdl(gemm_tcgen05_ppair_kernel<256, EW>, dim3(2 * npairs), dim3(THREADS), smem, stream, p);
  count_launch();
  return finish_launch("gemm_tcgen05_ppair");
}
static int launch_ppair(TcParams& p, int ngroups, cudaStream_t stream) {
  const char* env = getenv("UNAV_PP_EW");          // read per call: scripts/gemm_ab.py switches it inside one process
  const int ew = (env && atoi(env) == 16) ? 16 : 8;
  return ew == 8 ? launch_ppair_ew<8>(p, ngroups, stream) : launch_ppair_ew<16>(p, ngroups, stream);
}

// Persistent CTA pairs (UNAV_TC_PPAIR: 0 never, 1 whenever the shape allows, unset = the measured policy below).
// scripts/gemm_ab.py on B200, split operands, cold L2 (profiles/r02_gemm_ab.md): the persistent kernel wins where its tile list
// is at least two rounds of the 74 clusters (1x[7200,1536,512] 59 -> 48 us, 6x[3584,512,512] 57 -> 49, 3x[7168,512,512] 57 -> 48,
// 1x[16384,1280,224] 70 -> 62) or one well-filled round with a long k-loop (2x[3600,512,2048] 54 -> 48, 2x[3584,512,1536] 45 -> 41),
// ties with the one-tile pair kernel on the two head GEMMs (K = 1536 / 3072), and loses where a single partial round leaves
// the un-overlapped epilogue of a 128 x 256 CTA tile exposed (<= 60 tiles with K = 512, everything with <= 30 tiles).
static bool use_ppair(int M, int N, int K, int ngroups) {
  int v = -1;
  if (const char* env = getenv("UNAV_TC_PPAIR")) v = atoi(env);
  if (v == 0 || M < 256 || N % 256 != 0 || K < 64) return false;
  if (v == 1) return true;
  const long long tiles = static_cast<long long>((M + 255) / 256) * (N / 256) * ngroups;
  return tiles >= 148 || (tiles >= 56 && K >= 1024);
}

// which kernel the last tcgen05 GEMM call of this thread used (unav_gemm_last_variant): 0 <64,64>, 1 <128,32>, 2 <128,64>,
// 3 CTA pair 256 wide, 4 <64,32>, 5 CTA pair 128 wide, 6 <256,32> (experiment); -1 before the first call / for the CUDA-core backend
thread_local int g_last_variant = -1;
// cumulative launches per variant (unav_gemm_variant_counts): lets a parity test PROVE which tile variants a configuration
// exercised (the batch-16 path takes <128,32> and the CTA-pair kernels, the batch <= 4 paths never do)
static unsigned long long g_variant_counts[UNAV_GEMM_VARIANTS] = {};

int gemm_tcgen05(const UnavGemmGroup* groups, int ngroups, int M, int N, int K, int op_arg, int act,
                 int res_masked, cudaStream_t stream) {
  const int op_dtype = op_base(op_arg);
  UNAV_REQUIRE(op_is_16bit(op_dtype), "gemm_tcgen05: operands must be BF16 / F16 (got %d)", op_dtype);
  TcParams p;
  p.M = M; p.N = N; p.K = K; p.op_dtype = op_dtype; p.act = act; p.res_masked = res_masked;
  p.nseg = op_passes(op_arg);       // 1 pass on split operands reads the hi halves only
  p.phase = g_phase_buf; p.phase_cap = g_phase_cap;
  p.fine = (g_phase_buf && getenv("UNAV_PP_FINE")) ? 1 : 0;
  p.no_dry = getenv("UNAV_PP_NO_DRY") ? 1 : 0;           // A/B knob: skip the epilogue's instruction-cache warm-up pass
  const int conv_T = groups[0].conv_T;
  for (int i = 0; i < ngroups; ++i) UNAV_REQUIRE(groups[i].conv_T == conv_T, "gemm_tcgen05: groups must share conv_T");
  if (conv_T > 0)
    UNAV_REQUIRE(M % conv_T == 0 && K % 3 == 0 && (K / 3) % 32 == 0, "gemm_tcgen05: implicit conv needs M %% T == 0 and Cin %% 32 == 0");
  p.conv_T = conv_T; p.conv_cin = conv_T > 0 ? K / 3 : 0; p.conv_tiles = conv_T > 0 ? (conv_T + TC_BM - 1) / TC_BM : 0;
  bool ppair = conv_T == 0 && use_ppair(M, N, K, ngroups);
  if (ppair) {      // the persistent kernel's lean epilogue has no unaligned fallback: every vector access must be 16-byte clean
    auto al16 = [](const void* ptr) { return (reinterpret_cast<uintptr_t>(ptr) & 15) == 0; };
    for (int i = 0; i < ngroups && ppair; ++i) {
      const UnavGemmGroup& g = groups[i];
      const long long op_split = g.ld_op / 2;
      ppair = (!g.out_f32 || (al16(g.out_f32) && g.ld_f32 % 4 == 0)) && (!g.res || (al16(g.res) && g.ldres % 4 == 0)) &&
              (!g.out_op || ((reinterpret_cast<uintptr_t>(g.out_op) & 7) == 0 && g.ld_op % 4 == 0 && op_split % 4 == 0)) &&
              (!g.bias || al16(g.bias)) && (!g.colscale || al16(g.colscale)) &&
              (!g.gate || (g.gate_width > 0 && g.gate_width % 4 == 0));
    }
  }
  const int pair = ppair ? 256 : (conv_T == 0 ? use_pair(M, N, K, ngroups) : 0);
  // The dry epilogue pass pays the cold instruction fetch (~10 k clocks) while the epilogue warps wait for the first accumulator;
  // it only helps where that wait is at least as long: the persistent kernel with K >= 512 (13 k-clock k-loop: first pass
  // 9.9 k -> 3.0 k clocks, 1.24 -> 1.21 ms per step over its 21 launches).  In the one-tile kernels it DELAYS the real epilogue
  // (their k-loops last 3 - 8 k clocks: 1x[7168,512,512] 16.3 -> 24.8 us forced on for every K, and still 284 -> 331 us per
  // step for the <128,64> class with K >= 1024 only), so it stays off there (UNAV_TC_DRY=1 enables it for experiments).
  if (ppair ? K < 512 : !getenv("UNAV_TC_DRY")) p.no_dry = 1;
  const TcChoice ch = choose_tile(M, N, K, ngroups, p.nseg);
  const int bn = pair ? pair / 2 : ch.bn, bk = pair ? P2_BK : (ch.sched == 2 ? 32 : 64);
  p.once = ch.sched ? 1 : 0;
  for (int i = 0; i < ngroups; ++i) {
    const UnavGemmGroup& g = groups[i];
    UNAV_REQUIRE((reinterpret_cast<uintptr_t>(g.A) & 15) == 0 && (reinterpret_cast<uintptr_t>(g.W) & 15) == 0,
                 "gemm_tcgen05: A/W must be 16-byte aligned");
    UNAV_REQUIRE(g.lda % 8 == 0 && g.ldw % 8 == 0, "gemm_tcgen05: lda/ldw must be multiples of 8 (got %lld, %lld)", g.lda, g.ldw);
    if (op_is_split(op_dtype))
      UNAV_REQUIRE(g.lda % 16 == 0 && g.ldw % 16 == 0 && g.lda / 2 >= (conv_T > 0 ? K / 3 : K) && g.ldw / 2 >= K,
                   "gemm_tcgen05: split operands need ld %% 16 == 0 and ld/2 >= K");
    int rc;
    const bool split = op_is_split(op_dtype);
    const int halves = p.nseg > 1 ? 2 : 1;             // one pass over split operands fetches the hi halves only
    if (conv_T > 0) {
      if ((rc = encode_conv_map(&p.g[i].tmA, g.A, M / conv_T, conv_T, K / 3, g.lda, bk, split, halves))) return rc;
    } else if ((rc = encode_map(&p.g[i].tmA, g.A, M, K, g.lda, TC_BM, bk, split, halves))) return rc;
    if ((rc = encode_map(&p.g[i].tmW, g.W, N, K, g.ldw, bn, bk, split, halves))) return rc;
    p.g[i].epi = make_epi(g);
  }
  p.ngroups = ngroups;
  g_last_variant = ppair ? 7 : pair == 256 ? 3 : pair == 128 ? 5 : bn == 256 ? 6 : (bn == 128 ? (bk == 32 ? 1 : 2) : (bk == 32 ? 4 : 0));
  __atomic_fetch_add(&g_variant_counts[g_last_variant], 1ull, __ATOMIC_RELAXED);
  if (ppair) return launch_ppair(p, ngroups, stream);
  if (pair == 256) return launch_pair<256>(p, ngroups, stream);
  if (pair == 128) return launch_pair<128>(p, ngroups, stream);
  if (bn == 256) return launch_tc<256, 32>(p, ngroups, stream);
  if (bk == 32) return bn == 64 ? launch_tc<64, 32>(p, ngroups, stream) : launch_tc<128, 32>(p, ngroups, stream);
  return bn == 64 ? launch_tc<64, 64>(p, ngroups, stream) : launch_tc<128, 64>(p, ngroups, stream);
}

}  // namespace unav

extern "C" int unav_gemm_last_variant(void) { return unav::g_last_variant; }
extern "C" int unav_gemm_variant_counts(long long* out, int n) {
  if (!out || n < 0) return UNAV_ERR_BAD_ARG;
  for (int i = 0; i < n; ++i)
    out[i] = i < UNAV_GEMM_VARIANTS ? static_cast<long long>(__atomic_load_n(&unav::g_variant_counts[i], __ATOMIC_RELAXED)) : 0;
  return 0;
}
