// common.cuh — shared device/host helpers for the UnAV B200 hot-path kernels.
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/unav_b200.h"

// Operand formats.  UNAV_F32 / UNAV_BF16 are public; BF16X2 is the split format used for the
// FP32-accurate tensor-core mode: a row holds hi = bf16(x) in its first half and
// lo = bf16(x - hi) at column offset ld/2, so A.W^T ~ Ahi.Whi + Alo.Whi + Ahi.Wlo (3 MMA passes).

namespace unav {

// ---- error plumbing ---------------------------------------------------------------------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);
int finish_launch(const char* what);  // cudaGetLastError -> code, records message

#define UNAV_REQUIRE(cond, ...)            \
  do {                                     \
    if (!(cond)) {                         \
      ::unav::set_error(__VA_ARGS__);      \
      return UNAV_ERR_BAD_ARG;             \
    }                                      \
  } while (0)

// ---- programmatic dependent launch (PDL) -------------------------------------------------
// Every kernel of the hot path is launched with cudaLaunchAttributeProgrammaticStreamSerialization, so the NEXT
// kernel's CTAs may be scheduled (launch latency, smem carve-up, barrier init, TMEM allocation, tensor-map prefetch)
// while this one is still running.  Protocol: no global-memory access before pdl_wait(); pdl_wait() blocks until the
// preceding grid has completed and flushed; pdl_launch_dependents() right after it lets the successor start its own
// prologue.  Opt-in with UNAV_PDL=1 (without the launch attribute the device-side calls are no-ops).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
extern long long* g_phase_buf;
extern int g_phase_cap;
bool pdl_enabled();

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a PER-DEVICE setting: remember the largest size set per (kernel,
// device) and raise it under a lock, so a second GPU in the process (nn.DataParallel with two ids, tests on cuda:1) and
// concurrent host threads (DataParallel replicas) each get the attribute before their first launch.
constexpr int kMaxDevices = 64;
struct SmemAttr { size_t set[kMaxDevices]; };
void smem_attr_lock();
void smem_attr_unlock();
template <typename KernelT>
inline int ensure_dyn_smem(KernelT kernel, SmemAttr& st, size_t bytes, const char* what) {
  if (bytes <= 48 * 1024) return 0;                    // the default limit needs no opt-in
  int dev = 0;
  cudaGetDevice(&dev);
  const bool tracked = dev >= 0 && dev < kMaxDevices;
  smem_attr_lock();
  int rc = 0;
  if (!tracked || st.set[dev] < bytes) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes));
    if (e != cudaSuccess) {
      set_error("%s: cudaFuncSetAttribute(dynamic smem %zu) on device %d: %s", what, bytes, dev, cudaGetErrorString(e));
      rc = static_cast<int>(e);
    } else if (tracked) {
      st.set[dev] = bytes;
    }
  }
  smem_attr_unlock();
  return rc;
}

__device__ __forceinline__ long long clock_stamp() {   // "memory": keep the read where it is written
  long long t;
  asm volatile("mov.u64 %0, %%clock64;" : "=l"(t) :: "memory");
  return t;
}

// ---- small device helpers ---------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ float gelu_erf(float x) {
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}
// Single-instruction special functions (MUFU.RCP / MUFU.EX2, flush-to-zero): no range check, no slow-path call.  The IEEE
// variants (__frcp_rn, division, expf with its denormal rescale) each compile to a reconvergence region (BSSY / BRA / CALL /
// BSYNC) per ELEMENT, which the compiler cannot interleave across elements — the ncu source page of the fc1 + GELU launch
// (profiles/r02_ncu_full_ppair.csv) showed ~60 thread instructions per output element executed as one serial chain.
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// GELU for the tensor-core GEMM epilogues: erf by Abramowitz & Stegun 7.1.26 (|error| <= 1.5e-7): one MUFU.RCP, one MUFU.EX2
// and a degree-5 polynomial, branch-free (~16 instructions that interleave freely across the elements a lane holds, against
// ~40 plus a divergent branch for erff).  GELU max abs error vs float64 over [-8, 8]: < 1e-6 (torch's own FP32 CPU GELU, which
// the reference computes, is at 1.2e-6).  The epilogue of the FFN fc1 GEMMs (N = 2048) is issue bound on this function.
__device__ __forceinline__ float gelu_fast(float x) {
  const float z = x * 0.70710678118654752440f;
  const float a = fabsf(z);
  const float t = rcp_approx(fmaf(0.3275911f, a, 1.0f));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  poly *= t;
  const float e = ex2_approx(a * a * -1.4426950408889634f);
  const float erf_abs = fmaf(-poly, e, 1.0f);
  return 0.5f * x * (1.0f + copysignf(erf_abs, z));
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float silu_(float x) { return x / (1.0f + expf(-x)); }
// SiLU of the tensor-core GEMM epilogues, branch-free: x * rcp(1 + 2^(-x log2 e)); x -> -inf gives -0 like x / (1 + inf)
__device__ __forceinline__ float silu_fast(float x) { return x * rcp_approx(1.0f + ex2_approx(x * -1.4426950408889634f)); }

__device__ __forceinline__ float apply_act(float v, int act) {
  switch (act) {
    case UNAV_ACT_RELU: return fmaxf(v, 0.0f);
    case UNAV_ACT_GELU: return gelu_erf(v);
    case UNAV_ACT_SILU: return silu_(v);
    default: return v;
  }
}

// activation of the tcgen05 GEMM epilogues (every tile variant uses this one, so results do not depend on the variant)
__device__ __forceinline__ float apply_act_tc(float v, int act) {
  switch (act) {
    case UNAV_ACT_RELU: return fmaxf(v, 0.0f);
    case UNAV_ACT_GELU: return gelu_fast(v);
    case UNAV_ACT_SILU: return silu_fast(v);
    default: return v;
  }
}

// The 16-bit half type is a BUILD-time choice: libunav_b200.so is compiled for BF16 halves, libunav_b200_f16.so (same
// sources, -DUNAV_HALF_F16) for FP16 halves, and each accepts only its own op dtypes.  (A runtime switch in
// store_op4 / the GEMM epilogue cost 3.7 % of the whole step for a format that is chosen once per model.)
#ifdef UNAV_HALF_F16
constexpr bool kHalfF16 = true;
#else
constexpr bool kHalfF16 = false;
#endif
__host__ __device__ __forceinline__ bool op_is_split(int op) { return op == UNAV_BF16X2 || op == UNAV_F16X2; }
__host__ __device__ __forceinline__ constexpr bool op_is_f16(int) { return kHalfF16; }
__host__ __device__ __forceinline__ bool op_build_ok(int op) {
  return op == UNAV_F32 || (kHalfF16 ? (op == UNAV_F16 || op == UNAV_F16X2) : (op == UNAV_BF16 || op == UNAV_BF16X2));
}
#define UNAV_REQUIRE_OP(op, what)                                                                                 \
  UNAV_REQUIRE(unav::op_build_ok((op) & 0xff), "%s: op_dtype %d does not belong to this build (%s halves)", what, \
               (op) & 0xff, unav::kHalfF16 ? "FP16" : "BF16")
__host__ __device__ __forceinline__ bool op_is_16bit(int op) { return op >= UNAV_BF16 && op <= UNAV_F16X2; }
// op_dtype arguments may carry a pass count in bits 8..9 (UNAV_PASSES)
__host__ __device__ __forceinline__ int op_base(int op_arg) { return op_arg & 0xff; }
__host__ __device__ __forceinline__ int op_passes(int op_arg) {
  const int n = (op_arg >> 8) & 3;
  return !op_is_split(op_arg & 0xff) ? 1 : (n == 0 ? 3 : n);
}

// round-to-nearest 16-bit halves as raw bits (FP16 saturates to the largest finite value instead of overflowing to inf)
__device__ __forceinline__ uint16_t f2h16(float v, bool f16) {
  if (f16) {
    uint16_t h;
    asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(h) : "f"(v));
    return h;
  }
  return __bfloat16_as_ushort(__float2bfloat16_rn(v));
}
__device__ __forceinline__ float h162f(uint16_t h, bool f16) {
  return f16 ? __half2float(__ushort_as_half(h)) : __bfloat162float(__ushort_as_bfloat16(h));
}
__device__ __forceinline__ uint32_t pack2_h16(float a, float b, bool f16) {
  return static_cast<uint32_t>(f2h16(a, f16)) | (static_cast<uint32_t>(f2h16(b, f16)) << 16);
}

// Store one value into an operand buffer row (`p` points at the row start, `c` is the column).
// split_off = ld/2 for the split formats.
__device__ __forceinline__ void store_op(void* row, int op_dtype, long long c, long long split_off,
                                         float v) {
  if (op_dtype == UNAV_F32) {
    reinterpret_cast<float*>(row)[c] = v;
  } else {
    const bool f16 = op_is_f16(op_dtype);
    const uint16_t hi = f2h16(v, f16);
    reinterpret_cast<uint16_t*>(row)[c] = hi;
    if (op_is_split(op_dtype)) reinterpret_cast<uint16_t*>(row)[split_off + c] = f2h16(v - h162f(hi, f16), f16);
  }
}

// 4 consecutive columns (c % 4 == 0, buffers 16-byte aligned, ld % 8 == 0).
template <bool F16>
__device__ __forceinline__ void store_op4_16(uint16_t* r, bool split, long long c, long long split_off, float4 v) {
  uint2 pk, pl;
  if (F16) {
    const __half2 h01 = __floats2half2_rn(v.x, v.y), h23 = __floats2half2_rn(v.z, v.w);
    pk.x = *reinterpret_cast<const uint32_t*>(&h01);
    pk.y = *reinterpret_cast<const uint32_t*>(&h23);
    if (split) {
      const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
      const __half2 l01 = __floats2half2_rn(v.x - f01.x, v.y - f01.y), l23 = __floats2half2_rn(v.z - f23.x, v.w - f23.y);
      pl.x = *reinterpret_cast<const uint32_t*>(&l01);
      pl.y = *reinterpret_cast<const uint32_t*>(&l23);
    }
  } else {
    const __nv_bfloat162 h01 = __floats2bfloat162_rn(v.x, v.y), h23 = __floats2bfloat162_rn(v.z, v.w);
    pk.x = *reinterpret_cast<const uint32_t*>(&h01);
    pk.y = *reinterpret_cast<const uint32_t*>(&h23);
    if (split) {
      const float2 f01 = __bfloat1622float2(h01), f23 = __bfloat1622float2(h23);
      const __nv_bfloat162 l01 = __floats2bfloat162_rn(v.x - f01.x, v.y - f01.y), l23 = __floats2bfloat162_rn(v.z - f23.x, v.w - f23.y);
      pl.x = *reinterpret_cast<const uint32_t*>(&l01);
      pl.y = *reinterpret_cast<const uint32_t*>(&l23);
    }
  }
  *reinterpret_cast<uint2*>(r + c) = pk;
  if (split) *reinterpret_cast<uint2*>(r + split_off + c) = pl;
}

__device__ __forceinline__ void store_op4(void* row, int op_dtype, long long c, long long split_off,
                                          float4 v) {
  if (op_dtype == UNAV_F32) {
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(row) + c) = v;
  } else {
    store_op4_16<kHalfF16>(reinterpret_cast<uint16_t*>(row), op_is_split(op_dtype), c, split_off, v);
  }
}

__host__ __device__ __forceinline__ size_t op_elem_size(int op_dtype) {
  return op_dtype == UNAV_F32 ? 4 : 2;
}

// Load an operand element as float (SIMT GEMM path). For the split formats returns hi + lo.
__device__ __forceinline__ float load_op(const void* row, int op_dtype, long long c,
                                         long long split_off) {
  if (op_dtype == UNAV_F32) return reinterpret_cast<const float*>(row)[c];
  const bool f16 = op_is_f16(op_dtype);
  const uint16_t* r = reinterpret_cast<const uint16_t*>(row);
  float v = h162f(r[c], f16);
  if (op_is_split(op_dtype)) v += h162f(r[split_off + c], f16);
  return v;
}

// ---- shared GEMM epilogue ---------------------------------------------------------------
struct EpiParams {
  const float* bias;
  const uint8_t* rowmask;
  const float* rowscale;
  const float* gate;
  const float* res;
  const float* colscale;
  float* out_f32;
  void* out_op;
  void* out_opT;
  long long ldres, ld_f32, ld_op, ld_opT;
  int gate_groups, gate_width;
  int t_seg, t_col0, t_ncols;
};

__device__ __forceinline__ float epilogue_value(const EpiParams& e, int act, int res_masked,
                                                long long m, int n, float acc) {
  float v = acc;
  if (e.bias) v += __ldg(e.bias + n);
  float mk = 1.0f;
  if (e.rowmask) {
    mk = e.rowmask[m] ? 1.0f : 0.0f;
    v *= mk;
  }
  if (e.rowscale) v *= __ldg(e.rowscale + m);
  if (e.gate) v *= __ldg(e.gate + m * e.gate_groups + n / e.gate_width);
  v = apply_act(v, act);
  if (e.res) {
    float r = e.res[m * e.ldres + n];
    if (res_masked) r *= mk;
    float cs = e.colscale ? __ldg(e.colscale + n) : 1.0f;
    v = r + cs * v;
  }
  return v;
}

static inline EpiParams make_epi(const UnavGemmGroup& g) {
  EpiParams e;
  e.bias = g.bias; e.rowmask = g.rowmask; e.rowscale = g.rowscale; e.gate = g.gate;
  e.res = g.res; e.colscale = g.colscale; e.out_f32 = g.out_f32; e.out_op = g.out_op;
  e.ldres = g.ldres; e.ld_f32 = g.ld_f32; e.ld_op = g.ld_op;
  e.gate_groups = g.gate_groups > 0 ? g.gate_groups : 1;
  e.gate_width = g.gate_width > 0 ? g.gate_width : 1;
  e.out_opT = g.out_opT; e.ld_opT = g.ld_opT; e.t_seg = g.t_seg > 0 ? g.t_seg : 1;
  e.t_col0 = g.t_col0; e.t_ncols = g.t_ncols;
  return e;
}

// tcgen05 backend entry (gemm_tcgen05.cu)
int gemm_tcgen05(const UnavGemmGroup* groups, int ngroups, int M, int N, int K, int op_dtype,
                 int act, int res_masked, cudaStream_t stream);

}  // namespace unav
