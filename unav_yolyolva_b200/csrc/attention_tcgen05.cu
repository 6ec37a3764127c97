// attention_tcgen05.cu — fused masked attention on the sm_100a tensor cores for key lengths up to 256
// (every call of the T=224 configuration; longer sequences use the chunked CUDA-core kernel in attention.cu).
//
// One CTA = 128 queries of one (batch item, head).  192 threads:
//   warp 0      TMA: Q tile + whole K of the head (SWIZZLE_128B boxes) -> smem; after S is done, V^T of the head
//               into the same smem (K is dead by then).  V^T is stored per item, [nb][C][Tk], so that every TMA
//               box starts 16-byte aligned for any Tk (a [C][nb*Tk] layout faulted for Tk % 8 != 0)
//   warp 1      tcgen05.mma issuer: S[128 x Tk] = Q.K^T  (A, B from smem)  into TMEM columns [0, Tk);
//               then O[128 x hs] = P.V  with P read from TMEM (A operand in tensor memory) and V^T from smem
//   warps 2..5  one query row per thread: tcgen05.ld S -> key-validity mask + optional per-query extra key
//               (Alignment's time-aligned cross-modal token) -> row max / exp / row sum in registers ->
//               P packed to BF16 (hi [+ lo]) and written back to TMEM with tcgen05.st -> after the PV MMAs,
//               tcgen05.ld O -> 1/rowsum -> operand-dtype store
// The banded/valid-length mask never exists in memory: it is a per-key byte vector staged in smem and applied
// in-register.  With BF16X2 operands both GEMMs run the 3-pass split (hi.hi + lo.hi + hi.lo), so the result
// matches the FP32 reference to ~1e-5 (tests/test_gpu_attention.py).
#include <cuda.h>

#include "common.cuh"

namespace unav {

// PTX wrappers shared with gemm_tcgen05.cu (kept local: both files are self-contained translation units)
namespace atc {

constexpr long long SPIN_LIMIT = 4000000000LL;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity), "r"(0x10000u) : "memory");      // suspend hint: park instead of polling
  return ok;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > SPIN_LIMIT) __trap();
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] . B[smem]
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem]
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t* r) {
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
__device__ __forceinline__ void st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void st8(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr) {     // K-major, SWIZZLE_128B (see gemm_tcgen05.cu)
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
__device__ __forceinline__ uint32_t idesc_16(int m, int n, bool f16) {      // a/b format: 0 = F16, 1 = BF16
  return (1u << 4) | ((f16 ? 0u : 1u) << 7) | ((f16 ? 0u : 1u) << 10) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}

}  // namespace atc

struct AttnTcGroup {
  CUtensorMap tmQ[2], tmK[2], tmV[2];     // [0] = hi / plain, [1] = lo
  const uint8_t* kmask;
  const float* q32; const float* xk; const float* xv;   // FP32 rows for the optional extra key
  long long ldq32, ldx;
  void* out; long long ldo;
  int x_first;
  const uint8_t* qmask;
};
struct AttnTcParams {
  AttnTcGroup g[4];
  int nb, Tq, Tk, Tkp, nh, hs, op_dtype, nseg, ncols;
  int q_bytes, kv_bytes;      // smem region sizes
  long long* phase; int phase_cap;   // diagnostics (unav_set_phase_trace)
  float scale;
  // Key-chunked mode for Tk > 256 (nchunks > 0; BASELINE.json config 4, T = 2304): blockIdx.z also enumerates 256-key
  // chunks; a CTA attends its 128 queries to ONE chunk and writes the un-normalised partial output (FP32) and the row's
  // (max, sum) instead of the final rows; attention_merge_kernel combines the chunks (and the optional extra key).
  int nchunks, ng;
  float* part_o;              // [nchunks][ng][nb*Tq][nh*hs]
  float* part_ml;             // [nchunks][ng][nb*Tq][nh][2]
};

constexpr int ATC_THREADS = 320;     // TMA warp, MMA warp, 8 softmax / epilogue warps (two per TMEM lane quarter)

// p = exp(s*scale - max) for N (16 | 32) consecutive keys held in r[], packed to BF16 hi (+ lo) words; returns their sum
__device__ __forceinline__ float ex2_ftz(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// sc2 = scale * log2(e), mx2 = max * log2(e): p = 2^(s*sc2 - mx2) is one FFMA + one MUFU per key; packed 2-wide
// conversions (F2FP) — the scalar F2F conversions of an earlier version ran at MUFU rate and dominated the softmax.
template <int N, bool F16>
__device__ __forceinline__ float softmax_chunk_t(const uint32_t* r, uint32_t m, float sc2, float mx2, bool split,
                                                 uint32_t* hi, uint32_t* lo) {
  float l0 = 0.f, l1 = 0.f;
#pragma unroll
  for (int j = 0; j < N; j += 2) {
    const float p0 = ((m >> j) & 1u) ? ex2_ftz(fmaf(__uint_as_float(r[j]), sc2, -mx2)) : 0.f;
    const float p1 = ((m >> (j + 1)) & 1u) ? ex2_ftz(fmaf(__uint_as_float(r[j + 1]), sc2, -mx2)) : 0.f;
    float f0, f1;
    if (F16) {
      const __half2 h2 = __floats2half2_rn(p0, p1);
      hi[j / 2] = *reinterpret_cast<const uint32_t*>(&h2);
      const float2 hf = __half22float2(h2);
      f0 = hf.x; f1 = hf.y;
      if (split) {
        const __half2 l2 = __floats2half2_rn(p0 - f0, p1 - f1);
        lo[j / 2] = *reinterpret_cast<const uint32_t*>(&l2);
      }
    } else {
      const __nv_bfloat162 h2 = __floats2bfloat162_rn(p0, p1);
      hi[j / 2] = *reinterpret_cast<const uint32_t*>(&h2);
      const float2 hf = __bfloat1622float2(h2);
      f0 = hf.x; f1 = hf.y;
      if (split) {
        const __nv_bfloat162 l2 = __floats2bfloat162_rn(p0 - f0, p1 - f1);
        lo[j / 2] = *reinterpret_cast<const uint32_t*>(&l2);
      }
    }
    if (split) { l0 += p0; l1 += p1; }          // hi + lo carries (almost) the full FP32 value
    else { l0 += f0; l1 += f1; }                // normalise by what the MMA will actually sum
  }
  return l0 + l1;
}
template <int N>
__device__ __forceinline__ float softmax_chunk(const uint32_t* r, uint32_t m, float sc, float mx, bool split, bool f16,
                                               uint32_t* hi, uint32_t* lo) {
  (void)f16;
  return softmax_chunk_t<N, kHalfF16>(r, m, sc, mx, split, hi, lo);
}

__global__ void __launch_bounds__(ATC_THREADS, 1)
attention_tcgen05_kernel(const __grid_constant__ AttnTcParams p) {
  using namespace atc;
  extern __shared__ uint8_t smem_raw[];
  const bool chunked = p.nchunks > 0;
  const int zz = chunked ? blockIdx.z % (p.ng * p.nb) : blockIdx.z;
  const int chunk = chunked ? blockIdx.z / (p.ng * p.nb) : 0;
  const int k0 = chunk * 256;                    // first key of this CTA's chunk (0 in the single-tile mode)
  const int gi = zz / p.nb, b = zz % p.nb;
  const AttnTcGroup& g = p.g[gi];
  const int h = blockIdx.y, q0 = blockIdx.x * 128;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t q_smem = base, kv_smem = base + p.q_bytes;
  const uint32_t bar_base = kv_smem + p.kv_bytes;
  const uint32_t bar_qk = bar_base, bar_s = bar_base + 8, bar_v = bar_base + 16, bar_p = bar_base + 24, bar_o = bar_base + 32;
  const uint32_t tmem_slot = bar_base + 40;
  uint32_t* maskw = reinterpret_cast<uint32_t*>(smem_raw + (bar_base - smem_u32(smem_raw)) + 64);   // key validity bits, 9 words
  float* xch = reinterpret_cast<float*>(smem_raw + (bar_base - smem_u32(smem_raw)) + 128);         // [2 max | 2 sum][128 rows]
  const int hs = p.hs, Tkp = p.Tkp;
  const int nparts = p.nseg > 1 ? 2 : 1;
  const int kq = hs / 64;                       // 64-column boxes along the head dim
  const int nkv = (Tkp + 63) / 64;              // 64-key boxes of V^T
  // K is fetched in 64-key boxes into a region of ceil(Tkp / 64) * 64 rows per (part, 64-channel block), so that only the boxes
  // up to the item's last valid key are loaded
  const int q_box = 128 * 128, k_box = ((Tkp + 63) / 64) * 64 * 128, v_box = hs * 128;
  const int cta_lin = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
  long long* ph_out = (p.phase && cta_lin < p.phase_cap) ? p.phase + 8ll * cta_lin : nullptr;
  if (ph_out && threadIdx.x == 0) {
    uint32_t smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    ph_out[0] = smid;
    ph_out[1] = clock_stamp();
  }

  if (warp == 0 && lane == 0) {
    mbar_init(bar_qk, 1); mbar_init(bar_s, 1); mbar_init(bar_v, 1); mbar_init(bar_p, 8); mbar_init(bar_o, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(p.ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  pdl_wait();                 // PDL: the preceding grid has completed; nothing above touched global memory
  pdl_launch_dependents();    // let the next kernel's CTAs start their prologue
  if (warp < 9) {      // key validity as 8 x 32 bits (+ a zero word for the funnel shift of the last chunk)
    const int j = k0 + warp * 32 + lane;
    const bool valid = warp < 8 && j < p.Tk && (!g.kmask || g.kmask[static_cast<long long>(b) * p.Tk + j]);
    const uint32_t w = __ballot_sync(0xffffffffu, valid);
    if (lane == 0) maskw[warp] = w;
  }
  // query validity of this tile (optional): a tile without a valid query does no tensor work at all
  int q_valid = 1;
  if (g.qmask && !chunked) {
    const int qr = q0 + static_cast<int>(threadIdx.x);
    q_valid = threadIdx.x < 128 && qr < p.Tq && g.qmask[static_cast<long long>(b) * p.Tq + qr];
  }
  tc_fence_before();
  const bool any_q = __syncthreads_or(q_valid) != 0;
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");
  const uint32_t tm_s = tmem_base, tm_p = tmem_base + p.ncols / 2, tm_o = tmem_base;
  if (ph_out && threadIdx.x == 0) ph_out[2] = clock_stamp();
  // Keys beyond the last valid one contribute exactly 0 to every row (masked to -inf before the softmax in the reference):
  // they are not loaded, not multiplied and not exponentiated.  Tke = this CTA's key count, a multiple of 16 (>= 16: a fully
  // masked item still runs one masked block and yields the reference's 0/0 rows).
  int kmax = 0;
#pragma unroll
  for (int w = 7; w >= 0; --w)
    if (kmax == 0 && maskw[w] != 0) kmax = w * 32 + 32 - __clz(maskw[w]);
  const int Tke = min(Tkp, max(16, (kmax + 15) & ~15));
  const int nkb_k = (Tke + 63) / 64;            // 64-key boxes of K and of V^T that are needed
  // a chunk with no valid key (beyond the video's length) contributes nothing: record sum = 0 and skip the tensor work
  const bool live = !chunked || kmax != 0;

  if (!any_q) {
    // no valid query in this tile: zero rows (finite, so that the masked projection that follows stays finite)
    if (warp >= 2) {
      const int qd = warp & 3, half = (warp - 2) >> 2;
      const int qi = q0 + qd * 32 + lane;
      if (qi < p.Tq) {
        const size_t es = op_elem_size(p.op_dtype);
        char* orow = reinterpret_cast<char*>(g.out) + (static_cast<size_t>(b) * p.Tq + qi) * g.ldo * es;
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int c = half * (hs / 2); c < (half + 1) * (hs / 2); c += 4) store_op4(orow, p.op_dtype, h * hs + c, g.ldo / 2, z);
      }
    }
  } else if (!live) {
    if (warp >= 2 && ((warp - 2) >> 2) == 0) {
      const int qi = q0 + (warp & 3) * 32 + lane;
      if (qi < p.Tq) {
        const long long rowi = static_cast<long long>(b) * p.Tq + qi;
        float* ml = p.part_ml + (((static_cast<long long>(chunk) * p.ng + gi) * p.nb * p.Tq + rowi) * p.nh + h) * 2;
        ml[0] = -CUDART_INF_F; ml[1] = 0.f;
      }
    }
  } else if (warp == 0) {
    if (lane == 0) {
      // ---- Q tile + K of this (item, head)
      mbar_expect_tx(bar_qk, nparts * kq * (q_box + nkb_k * 64 * 128));
      for (int pt = 0; pt < nparts; ++pt)
        for (int kb = 0; kb < kq; ++kb) {
          tma_load_2d(q_smem + (pt * kq + kb) * q_box, &g.tmQ[pt], bar_qk, h * hs + kb * 64, b * p.Tq + q0);
          for (int i = 0; i < nkb_k; ++i)
            tma_load_2d(kv_smem + (pt * kq + kb) * k_box + i * (64 * 128), &g.tmK[pt], bar_qk, h * hs + kb * 64, b * p.Tk + k0 + i * 64);
        }
      // ---- V^T once the QK^T MMAs have finished reading K
      mbar_wait(bar_s, 0);
      mbar_expect_tx(bar_v, nparts * nkb_k * v_box);
      for (int pt = 0; pt < nparts; ++pt)
        for (int kb = 0; kb < nkb_k; ++kb)
          tma_load_2d(kv_smem + (pt * nkv + kb) * v_box, &g.tmV[pt], bar_v, k0 + kb * 64, b * p.nh * hs + h * hs);
    }
    __syncwarp();     // lanes 1..31 wait for the elected lane: the block barrier at the end must see whole warps
  } else if (warp == 1) {
    if (lane == 0) {
      // ---- S = Q.K^T : segments (Qhi,Khi), (Qlo,Khi), (Qhi,Klo)
      mbar_wait(bar_qk, 0);
      tc_fence_after();
      const uint32_t id_s = idesc_16(128, Tke, op_is_f16(p.op_dtype));
      uint32_t acc = 0;
      for (int seg = 0; seg < p.nseg; ++seg) {
        const int pa = seg == 1 ? 1 : 0, pb = seg == 2 ? 1 : 0;
        for (int ks = 0; ks < hs / 16; ++ks) {
          const uint64_t ad = smem_desc(q_smem + (pa * kq + ks / 4) * q_box) + 2u * (ks % 4);
          const uint64_t bd = smem_desc(kv_smem + (pb * kq + ks / 4) * k_box) + 2u * (ks % 4);
          mma_ss(tm_s, ad, bd, id_s, acc);
          acc = 1;
        }
      }
      tc_commit(bar_s);
      // ---- O = P.V : segments (Phi,Vhi), (Plo,Vhi), (Phi,Vlo); P is the A operand in tensor memory
      mbar_wait(bar_p, 0);
      mbar_wait(bar_v, 0);
      tc_fence_after();
      const uint32_t id_o = idesc_16(128, hs, op_is_f16(p.op_dtype));
      acc = 0;
      for (int seg = 0; seg < p.nseg; ++seg) {
        const int pa = seg == 1 ? 1 : 0, pb = seg == 2 ? 1 : 0;
        for (int ks = 0; ks < Tke / 16; ++ks) {
          const uint32_t a_t = tm_p + pa * (Tkp / 2) + ks * 8;        // 16 bf16 = 8 packed 32-bit columns
          const uint64_t bd = smem_desc(kv_smem + (pb * nkv + ks / 4) * v_box) + 2u * (ks % 4);
          mma_ts(tm_o, a_t, bd, id_o, acc);
          acc = 1;
        }
      }
      tc_commit(bar_o);
    }
    __syncwarp();
  } else {
    // ===== softmax / epilogue warps: two threads per query row (warps w and w+4 share a TMEM lane quarter and split the
    // key columns; row max and row sum are combined through shared memory under a 64-thread named barrier) =====
    const int qd = warp & 3, half = (warp - 2) >> 2;
    const int row = qd * 32 + lane;
    const int qi = q0 + row;
    const bool row_ok = qi < p.Tq;
    const uint32_t lane_addr = static_cast<uint32_t>(qd * 32) << 16;
    const float sc = p.scale;
    const bool split = p.nseg > 1, f16 = op_is_f16(p.op_dtype);
    const int csplit = ((Tke / 16 + 1) / 2) * 16;
    const int c_lo = half ? csplit : 0, c_hi = half ? Tke : csplit;
    // optional extra key: s_x = scale * <q_i, xk_i>
    const bool has_x = !chunked && g.xk != nullptr && row_ok && qi >= g.x_first;     // chunked: the merge kernel adds it
    float s_x = -CUDART_INF_F;
    if (has_x) {
      const float* qr = g.q32 + (static_cast<long long>(b) * p.Tq + qi) * g.ldq32 + h * hs;
      const float* kr = g.xk + (static_cast<long long>(b) * p.Tq + qi) * g.ldx + h * hs;
      float a = 0.f;
      for (int d = 0; d < hs; d += 4) {
        const float4 q4 = *reinterpret_cast<const float4*>(qr + d);
        const float4 k4 = *reinterpret_cast<const float4*>(kr + d);
        a = fmaf(q4.x, k4.x, a); a = fmaf(q4.y, k4.y, a); a = fmaf(q4.z, k4.z, a); a = fmaf(q4.w, k4.w, a);
      }
      s_x = a * sc;
    }
    mbar_wait(bar_s, 0);
    tc_fence_after();
    if (ph_out && threadIdx.x == 64) ph_out[3] = clock_stamp();
    // ---- this thread's S columns (<= 128 of them) are read from TMEM ONCE into registers: all loads in flight, one
    //      wait (the two-pass version re-read S for the exponentials and waited per 16-column chunk)
    uint32_t r[4][32];
    uint32_t mk[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int c = c_lo + 32 * k;
      mk[k] = 0;
      if (c < c_hi) {
        mk[k] = __funnelshift_r(maskw[c >> 5], maskw[(c >> 5) + 1], c & 31);
        if (c + 32 <= c_hi) {
          ld32(tm_s + lane_addr + c, r[k]);
        } else {
          ld16(tm_s + lane_addr + c, r[k]);
          mk[k] &= 0xffffu;
        }
      }
    }
    wait_ld();
    // ---- raw row max over the valid keys (scale > 0, so max(s*scale) = max(s)*scale exactly)
    float mraw = -CUDART_INF_F;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (c_lo + 32 * k < c_hi) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if ((mk[k] >> j) & 1u) mraw = fmaxf(mraw, __uint_as_float(r[k][j]));
      }
    }
    xch[half * 128 + row] = mraw;
    asm volatile("bar.sync %0, 64;" ::"r"(1 + qd) : "memory");
    float mx = fmaxf(s_x, fmaxf(xch[row], xch[128 + row]) * sc);
    if (mx == -CUDART_INF_F) mx = 0.f;         // fully masked row: all probabilities are 0 (0/0 = NaN as in the reference)
    // ---- p = exp(s*scale - max), row sum, P -> TMEM as packed 16-bit (hi, lo)
    const float sc2 = sc * 1.4426950408889634f, mx2 = mx * 1.4426950408889634f;
    float l = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int c = c_lo + 32 * k;
      if (c < c_hi) {
        uint32_t hi[16], lo[16];
        if (c + 32 <= c_hi) {
          l += softmax_chunk<32>(r[k], mk[k], sc2, mx2, split, f16, hi, lo);
          st16(tm_p + lane_addr + c / 2, hi);
          if (split) st16(tm_p + lane_addr + Tkp / 2 + c / 2, lo);
        } else {
          l += softmax_chunk<16>(r[k], mk[k], sc2, mx2, split, f16, hi, lo);
          st8(tm_p + lane_addr + c / 2, hi);
          if (split) st8(tm_p + lane_addr + Tkp / 2 + c / 2, lo);
        }
      }
    }
    wait_st();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_p);
    if (ph_out && threadIdx.x == 64) ph_out[4] = clock_stamp();
    xch[256 + half * 128 + row] = l;
    asm volatile("bar.sync %0, 64;" ::"r"(1 + qd) : "memory");
    l = xch[256 + row] + xch[384 + row];
    float p_x = 0.f;
    if (has_x) { p_x = ex2_ftz(s_x * 1.4426950408889634f - mx2); l += p_x; }
    // ---- epilogue: O / l (+ extra key's value), operand-dtype store; the two threads of a row split the head columns
    mbar_wait(bar_o, 0);
    tc_fence_after();
    if (ph_out && threadIdx.x == 64) ph_out[5] = clock_stamp();
    if (chunked) {
      // un-normalised partial output + (max, sum) of this chunk
      const long long rowi = static_cast<long long>(b) * p.Tq + (row_ok ? qi : 0);
      const long long slab = (static_cast<long long>(chunk) * p.ng + gi) * p.nb * p.Tq + rowi;
      if (row_ok && half == 0) {
        float* ml = p.part_ml + (slab * p.nh + h) * 2;
        ml[0] = l > 0.f ? mx : -CUDART_INF_F; ml[1] = l;
      }
      float* po = p.part_o + slab * (static_cast<long long>(p.nh) * hs) + h * hs;
#pragma unroll 1
      for (int c = half * (hs / 2); c < (half + 1) * (hs / 2); c += 16) {
        uint32_t r[16];
        ld16(tm_o + lane_addr + c, r);
        wait_ld();
        if (!row_ok) continue;
#pragma unroll
        for (int j = 0; j < 16; j += 4)
          *reinterpret_cast<float4*>(po + c + j) = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                                              __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
      }
    } else {
    const float inv = 1.0f / l;
    const size_t es = op_elem_size(p.op_dtype);
    char* orow = reinterpret_cast<char*>(g.out) + (static_cast<size_t>(b) * p.Tq + (row_ok ? qi : 0)) * g.ldo * es;
    const float* xvr = has_x ? g.xv + (static_cast<long long>(b) * p.Tq + qi) * g.ldx + h * hs : nullptr;
#pragma unroll 1
    for (int c = half * (hs / 2); c < (half + 1) * (hs / 2); c += 16) {
      uint32_t r[16];
      ld16(tm_o + lane_addr + c, r);
      wait_ld();
      if (!row_ok) continue;
#pragma unroll
      for (int j = 0; j < 16; j += 4) {
        float4 o = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
        if (xvr) {
          const float4 v4 = *reinterpret_cast<const float4*>(xvr + c + j);
          o.x = fmaf(p_x, v4.x, o.x); o.y = fmaf(p_x, v4.y, o.y); o.z = fmaf(p_x, v4.z, o.z); o.w = fmaf(p_x, v4.w, o.w);
        }
        o.x *= inv; o.y *= inv; o.z *= inv; o.w *= inv;
        store_op4(orow, p.op_dtype, h * hs + c + j, g.ldo / 2, o);
      }
    }
    }
    if (ph_out && threadIdx.x == 64) ph_out[6] = clock_stamp();
  }
  tc_fence_before();
  __syncthreads();
  if (ph_out && threadIdx.x == 0) ph_out[7] = clock_stamp();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.ncols) : "memory");
  }
}

// Combine the key chunks of one query row and head (chunked mode): with M = max_c m_c,
//   out = ( sum_c e^(m_c - M) O_c  +  p_x xv ) / ( sum_c e^(m_c - M) l_c  +  p_x ),   p_x = e^(s_x - M) for the optional extra key
// (Alignment's time-aligned token of the other modality: s_x = scale <q_i, xk_i>, FP32 rows).  One warp per (row, head).
struct AttnMergeParams {
  const float* part_o; const float* part_ml;
  const float* q32[4]; const float* xk[4]; const float* xv[4];
  long long ldq32[4], ldx[4];
  void* out[4]; long long ldo[4];
  int x_first[4];
  int nchunks, ng, nb, Tq, nh, hs, op_dtype;
  float scale;
};

__global__ void __launch_bounds__(256)
attention_merge_kernel(const __grid_constant__ AttnMergeParams p) {
  pdl_wait();
  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long rows = static_cast<long long>(p.nb) * p.Tq;
  const long long rowi = static_cast<long long>(blockIdx.x) * 8 + warp;
  if (rowi >= rows) return;
  const int h = blockIdx.y, gi = blockIdx.z;
  const int hs = p.hs, C = p.nh * hs;
  const int qi = static_cast<int>(rowi % p.Tq);
  const int per = hs / 32;                                 // 2 | 4 consecutive head columns per lane
  float s_x = -CUDART_INF_F;
  const bool has_x = p.xk[gi] != nullptr && qi >= p.x_first[gi];
  if (has_x) {
    const float* qr = p.q32[gi] + rowi * p.ldq32[gi] + h * hs;
    const float* kr = p.xk[gi] + rowi * p.ldx[gi] + h * hs;
    float a = 0.f;
    for (int d = lane; d < hs; d += 32) a = fmaf(qr[d], kr[d], a);
    s_x = warp_sum(a) * p.scale;
  }
  float M = s_x;
  for (int c = 0; c < p.nchunks; ++c)
    M = fmaxf(M, p.part_ml[((((static_cast<long long>(c) * p.ng + gi) * rows + rowi) * p.nh) + h) * 2]);
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  float den = 0.f;
  if (M > -CUDART_INF_F) {
    for (int c = 0; c < p.nchunks; ++c) {
      const long long slab = (static_cast<long long>(c) * p.ng + gi) * rows + rowi;
      const float* ml = p.part_ml + (slab * p.nh + h) * 2;
      const float l = ml[1];
      if (l > 0.f) {
        const float w = __expf(ml[0] - M);
        den = fmaf(w, l, den);
        const float* po = p.part_o + slab * C + h * hs + lane * per;
        for (int i = 0; i < per; ++i) acc[i] = fmaf(w, po[i], acc[i]);
      }
    }
    if (has_x) {
      const float px = __expf(s_x - M);
      den += px;
      const float* xr = p.xv[gi] + rowi * p.ldx[gi] + h * hs + lane * per;
      for (int i = 0; i < per; ++i) acc[i] = fmaf(px, xr[i], acc[i]);
    }
  }
  const float inv = 1.0f / den;                            // fully masked row: 0 / 0 = NaN, as in the reference's softmax
  const size_t es = op_elem_size(p.op_dtype);
  char* orow = reinterpret_cast<char*>(p.out[gi]) + static_cast<size_t>(rowi) * p.ldo[gi] * es;
  for (int i = 0; i < per; ++i) store_op(orow, p.op_dtype, h * hs + lane * per + i, p.ldo[gi] / 2, acc[i] * inv);
}

// ---- host ---------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* sym = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t err = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres);
  if (err != cudaSuccess || qres != cudaDriverEntryPointSuccess || !sym) {
    set_error("cuTensorMapEncodeTiled not available: %s", cudaGetErrorString(err));
    return nullptr;
  }
  fn = reinterpret_cast<EncodeTiledFn>(sym);
  return fn;
}

static int encode2d(CUtensorMap* map, const void* ptr, long long rows, long long cols, long long ld, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return UNAV_ERR_DRIVER;
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {64u, static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("attention_tc: cuTensorMapEncodeTiled failed (%d): rows=%lld cols=%lld ld=%lld box_rows=%d", (int)r, rows, cols, ld, box_rows);
    return UNAV_ERR_DRIVER;
  }
  return 0;
}


// =================================================================================================================
// MaxSigmoid gate on the tensor cores: gate[r, h] = sigmoid( max_n <x[r, h*hc:], G[item*nwords + n, h*hc:]> / sqrt(hc) + bias[h] )
// One CTA = 128 rows of one (item, head): S[128 x 512] = X_h . G_h^T as two N=256 tcgen05.mma chains into the whole
// TMEM (512 columns), then one row per thread reduces its 512 scores with tcgen05.ld + fmax.  (multimodal_backbones.py:170-191)
// =================================================================================================================
struct MaxsigTcParams {
  CUtensorMap tmX[2], tmG[2];
  const float* head_bias;
  float* gate;
  int nb, T, nwords, H, hc, nseg, x_col0, g_col0, f16;
};

__global__ void __launch_bounds__(192, 1)
maxsig_tcgen05_kernel(const __grid_constant__ MaxsigTcParams p) {
  using namespace atc;
  extern __shared__ uint8_t smem_raw[];
  const int b = blockIdx.z, h = blockIdx.y, t0 = blockIdx.x * 128;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int nparts = p.nseg > 1 ? 2 : 1;
  const int nhalf = p.nwords / 256;                         // 256-word boxes (nwords = 512 -> 2)
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t x_smem = base, g_smem = base + nparts * 16384;
  const uint32_t g_box = 256 * 128;
  const uint32_t bar_base = g_smem + nparts * nhalf * g_box;
  const uint32_t bar_ld = bar_base, bar_s = bar_base + 8, tmem_slot = bar_base + 16;
  if (warp == 0 && lane == 0) {
    mbar_init(bar_ld, 1); mbar_init(bar_s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");
  pdl_wait();                 // PDL: the preceding grid has completed; nothing above touched global memory
  pdl_launch_dependents();    // let the next kernel's CTAs start their prologue
  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(bar_ld, nparts * (16384 + nhalf * g_box));
      for (int pt = 0; pt < nparts; ++pt) {
        // 64-column boxes starting at this head's first channel; only the first hc columns are consumed
        tma_load_2d(x_smem + pt * 16384, &p.tmX[pt], bar_ld, p.x_col0 + h * p.hc, b * p.T + t0);
        for (int hf = 0; hf < nhalf; ++hf)
          tma_load_2d(g_smem + (pt * nhalf + hf) * g_box, &p.tmG[pt], bar_ld, p.g_col0 + h * p.hc, b * p.nwords + hf * 256);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      mbar_wait(bar_ld, 0);
      tc_fence_after();
      const uint32_t id = idesc_16(128, 256, p.f16 != 0);
      for (int hf = 0; hf < nhalf; ++hf) {
        uint32_t acc = 0;
        for (int seg = 0; seg < p.nseg; ++seg) {
          const int pa = seg == 1 ? 1 : 0, pb = seg == 2 ? 1 : 0;
          for (int ks = 0; ks < p.hc / 16; ++ks) {
            mma_ss(tmem_base + hf * 256, smem_desc(x_smem + pa * 16384) + 2u * ks,
                   smem_desc(g_smem + (pb * nhalf + hf) * g_box) + 2u * ks, id, acc);
            acc = 1;
          }
        }
      }
      tc_commit(bar_s);
    }
    __syncwarp();
  } else {
    const int qd = warp & 3;
    const int t = t0 + qd * 32 + lane;
    mbar_wait(bar_s, 0);
    tc_fence_after();
    // 64 scores per round trip to tensor memory (two x32 loads in flight, four independent max chains); a warp whose 32 rows
    // all lie beyond T (short pyramid levels: T = 7 .. 56 of a 128-row tile) skips the read.  The first version waited on a
    // 16-column load 32 times in a row: a ~4 k-clock dependent chain per CTA whatever T was.
    float mx = -CUDART_INF_F;
    if (t0 + qd * 32 < p.T) {
      float m0 = -CUDART_INF_F, m1 = -CUDART_INF_F, m2 = -CUDART_INF_F, m3 = -CUDART_INF_F;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(qd * 32) << 16);
#pragma unroll 1
      for (int c = 0; c < p.nwords; c += 64) {
        uint32_t ra[32], rb[32];
        ld32(taddr + c, ra);
        ld32(taddr + c + 32, rb);
        wait_ld();
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          m0 = fmaxf(m0, __uint_as_float(ra[j])); m1 = fmaxf(m1, __uint_as_float(ra[j + 1]));
          m2 = fmaxf(m2, __uint_as_float(rb[j])); m3 = fmaxf(m3, __uint_as_float(rb[j + 1]));
        }
      }
      mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
    }
    if (t < p.T)
      p.gate[(static_cast<long long>(b) * p.T + t) * p.H + h] = sigmoidf_(mx * (1.0f / sqrtf(static_cast<float>(p.hc))) + p.head_bias[h]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

}  // namespace unav

static size_t attn_long_ws_bytes(int ngroups, int nb, int Tq, int Tk, int nh, int hs) {
  const size_t nchunks = static_cast<size_t>((Tk + 255) / 256);
  const size_t rows = static_cast<size_t>(ngroups) * nb * Tq;
  return nchunks * rows * (static_cast<size_t>(nh) * hs + 2 * nh) * sizeof(float) + 256;
}

extern "C" size_t unav_attention_tc_workspace_bytes(int ngroups, int nb, int Tq, int Tk, int nh, int hs) {
  return Tk > 256 ? attn_long_ws_bytes(ngroups, nb, Tq, Tk, nh, hs) : 0;
}

static int attention_tc_impl(const UnavAttnTcGroup* groups, int ngroups, int nb, int Tq, int Tk, int nh, int hs,
                             float scale, int op_arg, void* workspace, size_t ws_bytes, void* stream) {
  using namespace unav;
  const int op_dtype = op_base(op_arg);
  UNAV_REQUIRE(groups && ngroups >= 1 && ngroups <= 4, "attention_tc: bad group count %d", ngroups);
  UNAV_REQUIRE(op_is_16bit(op_dtype), "attention_tc: operands must be BF16 / F16");
  UNAV_REQUIRE_OP(op_dtype, "attention_tc");
  UNAV_REQUIRE(hs == 64 || hs == 128, "attention_tc: head size %d not in {64,128}", hs);
  UNAV_REQUIRE(nb > 0 && Tq > 0 && Tk > 0 && nh > 0, "attention_tc: bad shape");
  const bool chunked = Tk > 256;
  UNAV_REQUIRE(!chunked || (workspace && ws_bytes >= attn_long_ws_bytes(ngroups, nb, Tq, Tk, nh, hs)),
               "attention_tc: Tk = %d > 256 needs unav_attention_tc_long with a workspace of unav_attention_tc_workspace_bytes()", Tk);
  AttnTcParams p;
  p.nb = nb; p.Tq = Tq; p.Tk = Tk; p.nh = nh; p.hs = hs; p.op_dtype = op_dtype; p.scale = scale;
  p.Tkp = chunked ? 256 : (Tk + 15) / 16 * 16;
  if (p.Tkp < 64) p.Tkp = 64;          // keys beyond Tk are masked; keeps every MMA / TMA box at least 64 wide
  p.nseg = op_passes(op_arg) == 2 ? 3 : op_passes(op_arg);     // 1 = hi halves only, else the full split
  p.ncols = p.Tkp > 128 ? 512 : 256;
  p.phase = g_phase_buf; p.phase_cap = g_phase_cap;
  p.nchunks = chunked ? (Tk + 255) / 256 : 0;
  p.ng = ngroups;
  p.part_o = nullptr; p.part_ml = nullptr;
  const long long C = static_cast<long long>(nh) * hs;
  if (chunked) {
    p.part_o = reinterpret_cast<float*>(workspace);
    p.part_ml = p.part_o + static_cast<size_t>(p.nchunks) * ngroups * nb * Tq * C;
  }
  const int nparts = p.nseg > 1 ? 2 : 1;
  const int kq = hs / 64, nkv = (p.Tkp + 63) / 64;
  p.q_bytes = nparts * kq * 128 * 128;
  const int kbytes = nparts * kq * nkv * 64 * 128, vbytes = nparts * nkv * hs * 128;      // K: whole 64-key boxes
  p.kv_bytes = ((kbytes > vbytes ? kbytes : vbytes) + 1023) / 1024 * 1024;
  for (int i = 0; i < ngroups; ++i) {
    const UnavAttnTcGroup& s = groups[i];
    UNAV_REQUIRE(s.q && s.k && s.vt && s.out, "attention_tc: null pointer in group %d", i);
    UNAV_REQUIRE(!s.xk || (s.xv && s.q32 && Tq == Tk), "attention_tc: extra key needs q32, xv and Tq == Tk");
    UNAV_REQUIRE(s.ldq % 8 == 0 && s.ldk % 8 == 0 && s.ldvt % 8 == 0, "attention_tc: leading dimensions must be multiples of 8");
    AttnTcGroup& d = p.g[i];
    int rc;
    const long long mq = static_cast<long long>(nb) * Tq, mk = static_cast<long long>(nb) * Tk;
    for (int pt = 0; pt < nparts; ++pt) {
      const __nv_bfloat16* q = reinterpret_cast<const __nv_bfloat16*>(s.q) + (pt ? s.ldq / 2 : 0);
      const __nv_bfloat16* k = reinterpret_cast<const __nv_bfloat16*>(s.k) + (pt ? s.ldk / 2 : 0);
      const __nv_bfloat16* v = reinterpret_cast<const __nv_bfloat16*>(s.vt) + (pt ? s.ldvt / 2 : 0);
      if ((rc = encode2d(&d.tmQ[pt], q, mq, C, s.ldq, 128))) return rc;
      if ((rc = encode2d(&d.tmK[pt], k, mk, C, s.ldk, 64))) return rc;
      if ((rc = encode2d(&d.tmV[pt], v, static_cast<long long>(nb) * C, Tk, s.ldvt, hs))) return rc;
    }
    if (nparts == 1) { d.tmQ[1] = d.tmQ[0]; d.tmK[1] = d.tmK[0]; d.tmV[1] = d.tmV[0]; }
    d.kmask = s.kmask; d.q32 = s.q32; d.xk = s.xk; d.xv = s.xv; d.ldq32 = s.ldq32; d.ldx = s.ldx;
    d.out = s.out; d.ldo = s.ldo; d.x_first = s.x_first; d.qmask = s.qmask;
  }
  const int smem = p.q_bytes + p.kv_bytes + 128 + 4 * 128 * 4 + 1024;     // + barriers, mask words, max/sum exchange, slack
  static SmemAttr attr = {};
  if (int rc = ensure_dyn_smem(attention_tcgen05_kernel, attr, smem, "attention_tc")) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  dim3 grid((Tq + 127) / 128, nh, nb * ngroups * (chunked ? p.nchunks : 1));
  launch_pdl(attention_tcgen05_kernel, dim3(grid), dim3(ATC_THREADS), smem, st, p);
  count_launch();
  int rc = finish_launch("attention_tc");
  if (rc || !chunked) return rc;
  AttnMergeParams m;
  m.part_o = p.part_o; m.part_ml = p.part_ml;
  for (int i = 0; i < 4; ++i) {
    const bool on = i < ngroups;
    m.q32[i] = on ? groups[i].q32 : nullptr; m.xk[i] = on ? groups[i].xk : nullptr; m.xv[i] = on ? groups[i].xv : nullptr;
    m.ldq32[i] = on ? groups[i].ldq32 : 0; m.ldx[i] = on ? groups[i].ldx : 0;
    m.out[i] = on ? groups[i].out : nullptr; m.ldo[i] = on ? groups[i].ldo : 0; m.x_first[i] = on ? groups[i].x_first : 0;
  }
  m.nchunks = p.nchunks; m.ng = ngroups; m.nb = nb; m.Tq = Tq; m.nh = nh; m.hs = hs; m.op_dtype = op_dtype; m.scale = scale;
  dim3 mg(static_cast<unsigned>((static_cast<long long>(nb) * Tq + 7) / 8), nh, ngroups);
  launch_pdl(attention_merge_kernel, dim3(mg), dim3(256), 0, st, m);
  count_launch();
  return finish_launch("attention_merge");
}

extern "C" int unav_attention_tc(const UnavAttnTcGroup* groups, int ngroups, int nb, int Tq, int Tk, int nh, int hs,
                                 float scale, int op_arg, void* stream) {
  return attention_tc_impl(groups, ngroups, nb, Tq, Tk, nh, hs, scale, op_arg, nullptr, 0, stream);
}

extern "C" int unav_attention_tc_long(const UnavAttnTcGroup* groups, int ngroups, int nb, int Tq, int Tk, int nh, int hs,
                                      float scale, int op_arg, void* workspace, size_t workspace_bytes, void* stream) {
  return attention_tc_impl(groups, ngroups, nb, Tq, Tk, nh, hs, scale, op_arg, workspace, workspace_bytes, stream);
}


extern "C" int unav_maxsig_gate_tc(const void* x, long long ldx, int x_col0, const void* G, long long ldg, int g_col0,
                                   const float* head_bias, float* gate, int nb, int T, int nwords, int H, int hc,
                                   int op_arg, void* stream) {
  using namespace unav;
  const int op_dtype = op_base(op_arg);
  UNAV_REQUIRE(x && G && head_bias && gate, "maxsig_gate_tc: null pointer");
  UNAV_REQUIRE(op_is_16bit(op_dtype), "maxsig_gate_tc: operands must be BF16 / F16");
  UNAV_REQUIRE_OP(op_dtype, "maxsig_gate_tc");
  UNAV_REQUIRE((hc == 32 || hc == 64) && nwords == 512, "maxsig_gate_tc: needs hc in {32,64} and 512 guide words (got %d, %d)", hc, nwords);
  UNAV_REQUIRE(ldx % 8 == 0 && ldg % 8 == 0 && x_col0 % 8 == 0 && g_col0 % 8 == 0, "maxsig_gate_tc: unaligned operand views");
  MaxsigTcParams p;
  p.head_bias = head_bias; p.gate = gate; p.nb = nb; p.T = T; p.nwords = nwords; p.H = H; p.hc = hc;
  p.nseg = op_passes(op_arg) == 2 ? 3 : op_passes(op_arg);
  p.f16 = kHalfF16 ? 1 : 0;
  p.x_col0 = x_col0; p.g_col0 = g_col0;
  const int nparts = p.nseg > 1 ? 2 : 1;
  // logical widths of the two operand matrices (columns beyond are read as zero by TMA)
  const long long xw = (op_is_split(op_dtype) ? ldx / 2 : ldx), gw = (op_is_split(op_dtype) ? ldg / 2 : ldg);
  for (int pt = 0; pt < nparts; ++pt) {
    const __nv_bfloat16* xp = reinterpret_cast<const __nv_bfloat16*>(x) + (pt ? ldx / 2 : 0);
    const __nv_bfloat16* gp = reinterpret_cast<const __nv_bfloat16*>(G) + (pt ? ldg / 2 : 0);
    int rc;
    if ((rc = encode2d(&p.tmX[pt], xp, static_cast<long long>(nb) * T, xw, ldx, 128))) return rc;
    if ((rc = encode2d(&p.tmG[pt], gp, static_cast<long long>(nb) * nwords, gw, ldg, 256))) return rc;
  }
  if (nparts == 1) { p.tmX[1] = p.tmX[0]; p.tmG[1] = p.tmG[0]; }
  const int smem = nparts * 16384 + nparts * (nwords / 256) * 256 * 128 + 64 + 1024;
  static SmemAttr attr = {};
  if (int rc = ensure_dyn_smem(maxsig_tcgen05_kernel, attr, smem, "maxsig_gate_tc")) return rc;
  dim3 grid((T + 127) / 128, H, nb);
  launch_pdl(maxsig_tcgen05_kernel, dim3(grid), dim3(192), smem, reinterpret_cast<cudaStream_t>(stream), p);
  count_launch();
  return finish_launch("maxsig_gate_tc");
}
