// rowops.cu — the bandwidth-bound row kernels of the hot path (token-major [rows, C] FP32 activations):
// LayerNorm (+pre-add, activation, position table, im2col scatter), depthwise-conv + mask + LayerNorm,
// row gather / up-sample / im2col copies, transposes, Alignment token embedding, pyramid masks,
// avg-pool + match projection, and the MaxSigmoid gate.  All 128-bit vectorised, one warp per row.
#include "common.cuh"

namespace unav {

constexpr int MAXV = 8;   // float4 chunks per lane: C <= 32 * 4 * MAXV = 1024

// =============================================================================================
// LayerNorm rows
// =============================================================================================
struct LnParams {
  UnavLnGroup g[UNAV_MAX_GROUPS];
  int M, C, act, op_dtype;
  float eps;
};

// NV = float4 chunks per lane.  EXACT: C == 128 * NV, no column predicates (the widths the hot path uses: 256, 512, 1024);
// otherwise the generic MAXV-chunk loop with `c < C` tests.  ACT >= 0: compile-time activation; -1: p.act at run time.
// The arithmetic (summation order, rounding) is the same in every instantiation, so the results are bit-identical.
// History (profiles/r01f_ln_rows_ncu_full.csv): the single generic kernel executed 836 warp instructions per 512-wide row
// (8 predicated chunks, run-time activation / output dispatch per chunk) at 82 registers = 5 CTAs of 128 threads per SM,
// 22 us for [7168, 512]: issue- and latency-bound at 1.2 TB/s.
template <int NV, int ACT, bool EXACT>
__device__ __forceinline__ void ln_rows_body(const LnParams& p) {
  const UnavLnGroup& g = p.g[blockIdx.y];
  const int lane = threadIdx.x & 31;
  const long long r = static_cast<long long>(blockIdx.x) * 4 + (threadIdx.x >> 5);
  if (r >= p.M) return;
  const int C = EXACT ? NV * 128 : p.C;
  const int act = ACT >= 0 ? ACT : p.act;
  long long src = r;
  if (g.x_seg_rows > 0) src = (r / g.x_seg_rows) * g.x_seg_stride + (r % g.x_seg_rows) + g.x_row_off;
  const float* xr = g.x + src * g.ldx;
  const float* ar = g.add ? g.add + r * g.ldadd : nullptr;
  float4 v[NV];
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int c = (j * 32 + lane) * 4;
    if (EXACT || c < C) {
      float4 t = *reinterpret_cast<const float4*>(xr + c);
      if (ar) {
        const float4 a = *reinterpret_cast<const float4*>(ar + c);
        t.x += a.x; t.y += a.y; t.z += a.z; t.w += a.w;
      }
      v[j] = t;
      s += (t.x + t.y) + (t.z + t.w);
    }
  }
  const float mean = warp_sum(s) / C;
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int c = (j * 32 + lane) * 4;
    if (EXACT || c < C) {
      v[j].x -= mean; v[j].y -= mean; v[j].z -= mean; v[j].w -= mean;
      q += (v[j].x * v[j].x + v[j].y * v[j].y) + (v[j].z * v[j].z + v[j].w * v[j].w);
    }
  }
  const float rstd = 1.0f / sqrtf(warp_sum(q) / C + p.eps);
  const float mk = g.rowmask ? (g.rowmask[r] ? 1.f : 0.f) : 1.f;
  const float* pr = g.post ? g.post + static_cast<long long>(r % g.post_rows) * C : nullptr;
  const uint8_t edge = g.edge ? g.edge[r] : 0;
  const size_t es = op_elem_size(p.op_dtype);
  float* f32_row = g.out_f32 ? g.out_f32 + r * g.ld_f32 : nullptr;
  char* op_row = g.out_op ? reinterpret_cast<char*>(g.out_op) + static_cast<size_t>(r) * g.ld_op * es : nullptr;
  char* ic_row = g.out_im2col ? reinterpret_cast<char*>(g.out_im2col) + static_cast<size_t>(r) * g.ld_im2col * es : nullptr;
  const long long ic_rowbytes = g.ld_im2col * static_cast<long long>(es);
  const long long sp_op = g.ld_op / 2, sp_ic = g.ld_im2col / 2;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int c = (j * 32 + lane) * 4;
    if (EXACT || c < C) {
      const float4 w = *reinterpret_cast<const float4*>(g.w + c);
      const float4 b = *reinterpret_cast<const float4*>(g.b + c);
      float4 y;
      y.x = apply_act(v[j].x * rstd * w.x + b.x, act);
      y.y = apply_act(v[j].y * rstd * w.y + b.y, act);
      y.z = apply_act(v[j].z * rstd * w.z + b.z, act);
      y.w = apply_act(v[j].w * rstd * w.w + b.w, act);
      if (pr) {
        const float4 t = *reinterpret_cast<const float4*>(pr + c);
        y.x += t.x * mk; y.y += t.y * mk; y.z += t.z * mk; y.w += t.w * mk;
      }
      if (f32_row) *reinterpret_cast<float4*>(f32_row + c) = y;
      if (op_row) store_op4(op_row, p.op_dtype, c, sp_op, y);
      if (ic_row) {
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        store_op4(ic_row, p.op_dtype, C + c, sp_ic, y);
        if (edge & 1) store_op4(ic_row, p.op_dtype, c, sp_ic, z);
        else store_op4(ic_row - ic_rowbytes, p.op_dtype, 2 * C + c, sp_ic, y);
        if (edge & 2) store_op4(ic_row, p.op_dtype, 2 * C + c, sp_ic, z);
        else store_op4(ic_row + ic_rowbytes, p.op_dtype, c, sp_ic, y);
      }
    }
  }
}

__global__ void __launch_bounds__(128)
ln_rows_kernel(const __grid_constant__ LnParams p) {
  pdl_wait();                 // PDL: the preceding grid has completed; nothing above touched global memory
  pdl_launch_dependents();    // let the next kernel's CTAs start their prologue
  ln_rows_body<MAXV, -1, false>(p);
}

template <int NV, int ACT>
__global__ void __launch_bounds__(128, NV == 8 ? 6 : 8)       // <= 64 registers (NV <= 4) / 80: 8 / 6 CTAs per SM instead of 5
ln_rows_exact_kernel(const __grid_constant__ LnParams p) {
  pdl_wait();
  pdl_launch_dependents();
  ln_rows_body<NV, ACT, true>(p);
}

// =============================================================================================
// depthwise conv (k=3) * mask -> LayerNorm, with optional pre-LayerNorms of the input rows
// =============================================================================================
struct DwLnParams {
  UnavDwLnGroup g[UNAV_MAX_GROUPS];
  int nseg, seg_len_in, seg_len_out, stride, C, n_pre, n_out, op_dtype;
  float eps;
};

constexpr int DWV = 4;        // float4 chunks per lane: C <= 512

// One block = R consecutive output rows of one segment (R = 8 for C <= 256, 4 for C <= 512), one WARP per unit of
// work so that no warp runs a second dependent pass: warp w stages input row w (phase 1: global load, pre-LayerNorm
// statistics, row -> smem) and then produces ONE (output row, output) pair (phase 2: depthwise taps over the staged
// rows with the pre-LN affine applied on the fly, mask, LayerNorm, vectorised FP32 / operand stores).  The weights of
// phase 2 (taps, pre-LN and LN affine) are requested before phase 1's reductions, so their L2 latency overlaps it.
// History: the 16-row / 8-warp version gave every warp ~3 rows of phase 1 and 6 pairs of phase 2 back to back; every
// launch took ~17 us whatever its grid size (profiles/r01b_launch_summary.md) — a latency chain, not bandwidth.
template <int NV>
__global__ void __launch_bounds__(NV == 2 ? 768 : 384)
dwconv_ln_kernel(const __grid_constant__ DwLnParams p) {
  pdl_wait();                 // PDL: the preceding grid has completed; nothing above touched global memory
  pdl_launch_dependents();    // let the next kernel's CTAs start their prologue
  extern __shared__ __align__(16) float dw_smem[];
  constexpr int R = NV == 2 ? 8 : 4;
  const UnavDwLnGroup& g = p.g[blockIdx.y];
  const int C = p.C;
  const int tiles_per_seg = (p.seg_len_out + R - 1) / R;
  const int seg = blockIdx.x / tiles_per_seg, t0 = (blockIdx.x % tiles_per_seg) * R;
  const int n_in = R * p.stride + 2;                       // staged input rows: stride*t0 - 1 ...
  float* xs = dw_smem;                                     // [n_in][C]
  float* st_mean = dw_smem + n_in * C;                     // [n_in]
  float* st_rstd = st_mean + n_in;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ---- phase 1 loads (input row `warp`)
  float4 v[NV];
  bool ok = false;
  if (warp < n_in) {
    const int ti = p.stride * t0 - 1 + warp;
    ok = ti >= 0 && ti < p.seg_len_in;
    const float* xr = g.x + (static_cast<long long>(seg) * p.seg_len_in + (ok ? ti : 0)) * g.ldx;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int c = (j * 32 + lane) * 4;
      v[j] = (ok && c < C) ? *reinterpret_cast<const float4*>(xr + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  // ---- phase 2 weights (pair `warp` = (row ro, output o)), requested now, consumed after the barrier
  const int ro = warp / p.n_out, o = warp - ro * p.n_out;
  const int t = t0 + ro;
  const bool task = ro < R && t < p.seg_len_out;
  const UnavDwLnOut& od = g.out[task ? o : 0];
  const bool has_pre = task && od.src >= 0;
  const bool has_ln = od.ln_w != nullptr;   // NULL: plain masked depthwise conv (MaskedConv1D with groups = C)
  // NV == 2 (C <= 256) keeps them in registers; NV == 4 (C <= 512) would need 112 registers for them (154 in total, one
  // CTA per SM, 12 waves on the [7168,512] launches), so there the CTA stages all its weights in shared memory instead.
  constexpr bool WSM = NV == 4;
  constexpr int NR = WSM ? 1 : NV;
  float4 d0[NR], d1[NR], d2[NR], lw[NR], lb[NR], nw[NR], nb[NR];
  float* wsm = st_rstd + n_in;                             // [n_out][dw 3C | ln_w C | ln_b C] then [2][pre_w C | pre_b C]
  float mk = 1.f;
  long long r = 0;
  if (WSM) {
    const int C4 = C / 4;
    for (int oo = 0; oo < p.n_out; ++oo) {
      const UnavDwLnOut& q = g.out[oo];
      float4* dst = reinterpret_cast<float4*>(wsm + oo * 5 * C);
      for (int i = threadIdx.x; i < 3 * C4; i += blockDim.x) dst[i] = reinterpret_cast<const float4*>(q.dw)[i];
      if (q.ln_w)
        for (int i = threadIdx.x; i < C4; i += blockDim.x) {
          dst[3 * C4 + i] = reinterpret_cast<const float4*>(q.ln_w)[i];
          dst[4 * C4 + i] = reinterpret_cast<const float4*>(q.ln_b)[i];
        }
    }
    for (int sidx = 0; sidx < p.n_pre; ++sidx) {
      float4* dst = reinterpret_cast<float4*>(wsm + p.n_out * 5 * C + sidx * 2 * C);
      for (int i = threadIdx.x; i < C4; i += blockDim.x) {
        dst[i] = reinterpret_cast<const float4*>(g.pre_w[sidx])[i];
        dst[C4 + i] = reinterpret_cast<const float4*>(g.pre_b[sidx])[i];
      }
    }
  }
  if (task) {
    r = static_cast<long long>(seg) * p.seg_len_out + t;
    mk = g.mask_out ? (g.mask_out[r] ? 1.f : 0.f) : 1.f;
  }
  if (task && !WSM) {
    const float* pw = has_pre ? g.pre_w[od.src] : nullptr;
    const float* pb = has_pre ? g.pre_b[od.src] : nullptr;
#pragma unroll
    for (int j = 0; j < NR; ++j) {
      const int c = (j * 32 + lane) * 4;
      if (c < C) {
        d0[j] = *reinterpret_cast<const float4*>(od.dw + c * 3);       // taps of 4 channels: 12 floats
        d1[j] = *reinterpret_cast<const float4*>(od.dw + c * 3 + 4);
        d2[j] = *reinterpret_cast<const float4*>(od.dw + c * 3 + 8);
        if (has_pre) { lw[j] = *reinterpret_cast<const float4*>(pw + c); lb[j] = *reinterpret_cast<const float4*>(pb + c); }
        if (has_ln) { nw[j] = *reinterpret_cast<const float4*>(od.ln_w + c); nb[j] = *reinterpret_cast<const float4*>(od.ln_b + c); }
      }
    }
  }
  // ---- phase 1 finish: row -> smem, statistics
  if (warp < n_in) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int c = (j * 32 + lane) * 4;
      if (c < C) {
        *reinterpret_cast<float4*>(xs + warp * C + c) = v[j];
        s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
      }
    }
    if (p.n_pre > 0) {
      const float mean = warp_sum(s) / C;
      float q = 0.f;
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const int c = (j * 32 + lane) * 4;
        if (c < C) {
          const float a = v[j].x - mean, b = v[j].y - mean, cc = v[j].z - mean, d = v[j].w - mean;
          q += (a * a + b * b) + (cc * cc + d * d);
        }
      }
      const float rstd = 1.0f / sqrtf(warp_sum(q) / C + p.eps);
      if (lane == 0) { st_mean[warp] = mean; st_rstd[warp] = ok ? rstd : 0.f; }   // rstd 0 + skipped affine = zero padding
    }
  }
  __syncthreads();
  if (!task) return;

  // ---- phase 2
  const size_t es = op_elem_size(p.op_dtype);
  const int i0 = ro * p.stride;                          // staged row of tap 0
  bool okt[3];
  float tm[3], tr[3];
#pragma unroll
  for (int tap = 0; tap < 3; ++tap) {
    const int ti = p.stride * t + tap - 1;
    okt[tap] = ti >= 0 && ti < p.seg_len_in;
    tm[tap] = has_pre ? st_mean[i0 + tap] : 0.f;
    tr[tap] = has_pre ? st_rstd[i0 + tap] : 1.f;
  }
  float4 z[NV];
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int c = (j * 32 + lane) * 4;
    z[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < C) {
      const int jr = WSM ? 0 : j;
      if (WSM) {
        const float* wo = wsm + o * 5 * C;
        d0[0] = *reinterpret_cast<const float4*>(wo + c * 3);
        d1[0] = *reinterpret_cast<const float4*>(wo + c * 3 + 4);
        d2[0] = *reinterpret_cast<const float4*>(wo + c * 3 + 8);
        if (has_pre) {
          const float* wp = wsm + p.n_out * 5 * C + od.src * 2 * C;
          lw[0] = *reinterpret_cast<const float4*>(wp + c);
          lb[0] = *reinterpret_cast<const float4*>(wp + C + c);
        }
      }
      const float wt[4][3] = {{d0[jr].x, d0[jr].y, d0[jr].z}, {d0[jr].w, d1[jr].x, d1[jr].y}, {d1[jr].z, d1[jr].w, d2[jr].x},
                              {d2[jr].y, d2[jr].z, d2[jr].w}};
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int tap = 0; tap < 3; ++tap) {
        if (!okt[tap]) continue;
        float4 u = *reinterpret_cast<const float4*>(xs + (i0 + tap) * C + c);
        if (has_pre) {
          u.x = (u.x - tm[tap]) * tr[tap] * lw[jr].x + lb[jr].x; u.y = (u.y - tm[tap]) * tr[tap] * lw[jr].y + lb[jr].y;
          u.z = (u.z - tm[tap]) * tr[tap] * lw[jr].z + lb[jr].z; u.w = (u.w - tm[tap]) * tr[tap] * lw[jr].w + lb[jr].w;
        }
        acc[0] = fmaf(wt[0][tap], u.x, acc[0]); acc[1] = fmaf(wt[1][tap], u.y, acc[1]);
        acc[2] = fmaf(wt[2][tap], u.z, acc[2]); acc[3] = fmaf(wt[3][tap], u.w, acc[3]);
      }
      z[j] = make_float4(acc[0] * mk, acc[1] * mk, acc[2] * mk, acc[3] * mk);
      s += (z[j].x + z[j].y) + (z[j].z + z[j].w);
    }
  }
  const float mu = warp_sum(s) / C;
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int c = (j * 32 + lane) * 4;
    if (c < C) {
      z[j].x -= mu; z[j].y -= mu; z[j].z -= mu; z[j].w -= mu;
      q += (z[j].x * z[j].x + z[j].y * z[j].y) + (z[j].z * z[j].z + z[j].w * z[j].w);
    }
  }
  const float rs = 1.0f / sqrtf(warp_sum(q) / C + p.eps);
  char* op_row = od.out_op ? reinterpret_cast<char*>(od.out_op) + static_cast<size_t>(r) * od.ld_op * es : nullptr;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int c = (j * 32 + lane) * 4;
    if (c < C) {
      float4 y;
      if (has_ln) {
        const int jr = WSM ? 0 : j;
        if (WSM) {
          nw[0] = *reinterpret_cast<const float4*>(wsm + o * 5 * C + 3 * C + c);
          nb[0] = *reinterpret_cast<const float4*>(wsm + o * 5 * C + 4 * C + c);
        }
        y.x = z[j].x * rs * nw[jr].x + nb[jr].x; y.y = z[j].y * rs * nw[jr].y + nb[jr].y;
        y.z = z[j].z * rs * nw[jr].z + nb[jr].z; y.w = z[j].w * rs * nw[jr].w + nb[jr].w;
      } else {
        y = make_float4(z[j].x + mu, z[j].y + mu, z[j].z + mu, z[j].w + mu);
      }
      if (od.out_f32) *reinterpret_cast<float4*>(od.out_f32 + r * od.ld_f32 + c) = y;
      if (op_row) store_op4(op_row, p.op_dtype, c, od.ld_op / 2, y);
    }
  }
}

// =============================================================================================
// dwconv_ln, streaming version (the default): same math, same per-lane channel mapping and operation order as
// dwconv_ln_kernel above (bit-identical outputs, tests/test_gpu_rowops.py), different schedule.  The tiled kernel stages a
// CTA's weights (up to 39 KB: taps + LN affine of three outputs + two pre-LN affines) for FOUR output rows and sends every
// input row through shared memory; on the [7168, 512] x 3 stem launches that is 70 MB of weight traffic for 59 MB of data and
// 57 % issue utilisation on 154 registers.  Here a CTA stages the weights ONCE and each of its warps walks a strip of
// consecutive output rows of one segment with the three input rows of the current window in REGISTERS (normalised once per
// input row when a pre-LayerNorm is present: n = (x - mean) * rstd; the per-source affine is one FMA per tap), the next
// input row's loads issued before the current row is computed.
// =============================================================================================
// One (output row, output) pair of the streaming kernel: taps over the three window rows held in registers, mask, LayerNorm,
// stores.  PRE / EDGE are compile-time so that the interior rows (all but the first and last row of a segment) carry no per-tap
// test at all: the ncu source page of the first version showed 30 reconvergence regions (BSSY / BSYNC) and 52 branches per
// pair, most of them the `skip this tap` test repeated per float4 chunk.  Without a pre-LayerNorm a window row outside the
// segment is all zeros and fmaf(w, 0, acc) == acc, so the test is not needed there either (same bits as skipping the tap).
template <int NV, bool PRE, bool EDGE>
__device__ __forceinline__ void dws_pair(const float4 (&w0)[NV], const float4 (&w1)[NV], const float4 (&w2)[NV], const float* wo,
                                         const float* wp, bool ok0, bool ok2, float mk, float eps, const UnavDwLnOut& od,
                                         long long r, int lane, int op_dtype) {
  constexpr int C = NV * 128;
  const bool has_ln = od.ln_w != nullptr;
  float4 z[NV];
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int c = (j * 32 + lane) * 4;
    const float4 d0 = *reinterpret_cast<const float4*>(wo + c * 3);
    const float4 d1 = *reinterpret_cast<const float4*>(wo + c * 3 + 4);
    const float4 d2 = *reinterpret_cast<const float4*>(wo + c * 3 + 8);
    float4 lw = make_float4(1.f, 1.f, 1.f, 1.f), lb = make_float4(0.f, 0.f, 0.f, 0.f);
    if (PRE) { lw = *reinterpret_cast<const float4*>(wp + c); lb = *reinterpret_cast<const float4*>(wp + C + c); }
    const float wt[4][3] = {{d0.x, d0.y, d0.z}, {d0.w, d1.x, d1.y}, {d1.z, d1.w, d2.x}, {d2.y, d2.z, d2.w}};
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int tap = 0; tap < 3; ++tap) {
      if (PRE && EDGE && ((tap == 0 && !ok0) || (tap == 2 && !ok2))) continue;      // warp-uniform, edge rows only
      float4 u = tap == 0 ? w0[j] : (tap == 1 ? w1[j] : w2[j]);
      if (PRE) {
        u.x = u.x * lw.x + lb.x; u.y = u.y * lw.y + lb.y; u.z = u.z * lw.z + lb.z; u.w = u.w * lw.w + lb.w;
      }
      acc[0] = fmaf(wt[0][tap], u.x, acc[0]); acc[1] = fmaf(wt[1][tap], u.y, acc[1]);
      acc[2] = fmaf(wt[2][tap], u.z, acc[2]); acc[3] = fmaf(wt[3][tap], u.w, acc[3]);
    }
    z[j] = make_float4(acc[0] * mk, acc[1] * mk, acc[2] * mk, acc[3] * mk);
    sum += (z[j].x + z[j].y) + (z[j].z + z[j].w);
  }
  const float mu = warp_sum(sum) / C;
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    z[j].x -= mu; z[j].y -= mu; z[j].z -= mu; z[j].w -= mu;
    q += (z[j].x * z[j].x + z[j].y * z[j].y) + (z[j].z * z[j].z + z[j].w * z[j].w);
  }
  const float rs = 1.0f / sqrtf(warp_sum(q) / C + eps);
  const size_t es = op_elem_size(op_dtype);
  char* op_row = od.out_op ? reinterpret_cast<char*>(od.out_op) + static_cast<size_t>(r) * od.ld_op * es : nullptr;
  float* f32_row = od.out_f32 ? od.out_f32 + r * od.ld_f32 : nullptr;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int c = (j * 32 + lane) * 4;
    float4 y;
    if (has_ln) {
      const float4 nw = *reinterpret_cast<const float4*>(wo + 3 * C + c);
      const float4 nb = *reinterpret_cast<const float4*>(wo + 4 * C + c);
      y.x = z[j].x * rs * nw.x + nb.x; y.y = z[j].y * rs * nw.y + nb.y;
      y.z = z[j].z * rs * nw.z + nb.z; y.w = z[j].w * rs * nw.w + nb.w;
    } else {
      y = make_float4(z[j].x + mu, z[j].y + mu, z[j].z + mu, z[j].w + mu);
    }
    if (f32_row) *reinterpret_cast<float4*>(f32_row + c) = y;
    if (op_row) store_op4(op_row, op_dtype, c, od.ld_op / 2, y);
  }
}

template <int NV, int STRIDE>
__global__ void __launch_bounds__(256, 2)
dwconv_ln_stream_kernel(const __grid_constant__ DwLnParams p, int strip) {
  pdl_wait();
  pdl_launch_dependents();
  extern __shared__ __align__(16) float dws_smem[];     // [n_out][dw 3C | ln_w C | ln_b C] then [n_pre][pre_w C | pre_b C]
  const UnavDwLnGroup& g = p.g[blockIdx.y];
  constexpr int C = NV * 128;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  {
    constexpr int C4 = C / 4;
    for (int oo = 0; oo < p.n_out; ++oo) {
      const UnavDwLnOut& q = g.out[oo];
      float4* dst = reinterpret_cast<float4*>(dws_smem + oo * 5 * C);
      for (int i = threadIdx.x; i < 3 * C4; i += blockDim.x) dst[i] = reinterpret_cast<const float4*>(q.dw)[i];
      if (q.ln_w)
        for (int i = threadIdx.x; i < C4; i += blockDim.x) {
          dst[3 * C4 + i] = reinterpret_cast<const float4*>(q.ln_w)[i];
          dst[4 * C4 + i] = reinterpret_cast<const float4*>(q.ln_b)[i];
        }
    }
    for (int sidx = 0; sidx < p.n_pre; ++sidx) {
      float4* dst = reinterpret_cast<float4*>(dws_smem + p.n_out * 5 * C + sidx * 2 * C);
      for (int i = threadIdx.x; i < C4; i += blockDim.x) {
        dst[i] = reinterpret_cast<const float4*>(g.pre_w[sidx])[i];
        dst[C4 + i] = reinterpret_cast<const float4*>(g.pre_b[sidx])[i];
      }
    }
  }
  __syncthreads();
  const bool pre = p.n_pre > 0;
  const int strips_per_seg = (p.seg_len_out + strip - 1) / strip;
  const int total = p.nseg * strips_per_seg;

  // load input row ti of segment seg (zeros outside the segment); with a pre-LayerNorm the row comes back normalised
  auto load_row = [&](int seg, int ti, float4 (&v)[NV]) {
    const bool ok = ti >= 0 && ti < p.seg_len_in;
    const float* xr = g.x + (static_cast<long long>(seg) * p.seg_len_in + (ok ? ti : 0)) * g.ldx;
#pragma unroll
    for (int j = 0; j < NV; ++j)
      v[j] = ok ? *reinterpret_cast<const float4*>(xr + (j * 32 + lane) * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
  };
  auto normalise = [&](float4 (&v)[NV]) {           // statistics exactly as the tiled kernel computes them
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) sum += (v[j].x + v[j].y) + (v[j].z + v[j].w);
    const float mean = warp_sum(sum) / C;
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const float a = v[j].x - mean, b = v[j].y - mean, cc = v[j].z - mean, d = v[j].w - mean;
      q += (a * a + b * b) + (cc * cc + d * d);
    }
    const float rstd = 1.0f / sqrtf(warp_sum(q) / C + p.eps);
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      v[j].x = (v[j].x - mean) * rstd; v[j].y = (v[j].y - mean) * rstd;
      v[j].z = (v[j].z - mean) * rstd; v[j].w = (v[j].w - mean) * rstd;
    }
  };

  // A masked output row (mask_out == 0) is exactly its LayerNorm bias: z = conv * 0 = 0, so mean = 0, variance = 0 and
  // y = 0 * rstd * w + b = b (plain zero without a LayerNorm), whatever the input holds — the reference computes the same
  // thing the long way (blocks.py:56-61 then :99-103).  Such rows are written without loading or convolving anything; with
  // the padded batches of the path (valid length 60 .. 187 of 224 frames) that is ~45 % of the rows.
  auto masked_row = [&](long long r) {
    const size_t es = op_elem_size(p.op_dtype);
    for (int o = 0; o < p.n_out; ++o) {
      const UnavDwLnOut& od = g.out[o];
      const float* wo = dws_smem + o * 5 * C;
      char* op_row = od.out_op ? reinterpret_cast<char*>(od.out_op) + static_cast<size_t>(r) * od.ld_op * es : nullptr;
      float* f32_row = od.out_f32 ? od.out_f32 + r * od.ld_f32 : nullptr;
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const int c = (j * 32 + lane) * 4;
        const float4 y = od.ln_w ? *reinterpret_cast<const float4*>(wo + 4 * C + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        if (f32_row) *reinterpret_cast<float4*>(f32_row + c) = y;
        if (op_row) store_op4(op_row, p.op_dtype, c, od.ld_op / 2, y);
      }
    }
  };

  for (int sidx = blockIdx.x * nwarps + warp; sidx < total; sidx += gridDim.x * nwarps) {
    const int seg = sidx / strips_per_seg;
    const int t0 = (sidx - seg * strips_per_seg) * strip;
    const int t1 = min(t0 + strip, p.seg_len_out);
    const long long r0 = static_cast<long long>(seg) * p.seg_len_out + t0;
    // validity bits of the strip's rows (strip <= 32): rows after the last valid one need no input at all
    const bool lane_valid = lane < t1 - t0 && (!g.mask_out || g.mask_out[r0 + lane] != 0);
    const uint32_t vbits = __ballot_sync(0xffffffffu, lane_valid);
    const int t_last = vbits ? t0 + 31 - __clz(vbits) : t0 - 1;          // last valid row of the strip
    for (int t = t_last + 1; t < t1; ++t) masked_row(r0 + (t - t0));
    if (t_last < t0) continue;
    float4 w0[NV], w1[NV], w2[NV], nx[NV], nx2[NV];     // window rows STRIDE*t - 1, STRIDE*t, STRIDE*t + 1 and the prefetch
    load_row(seg, STRIDE * t0 - 1, w0);
    load_row(seg, STRIDE * t0, w1);
    load_row(seg, STRIDE * t0 + 1, w2);
    if (pre) { normalise(w0); normalise(w1); normalise(w2); }
    for (int t = t0; t <= t_last; ++t) {
      const bool more = t < t_last;
      if (more) {                                     // next window's new rows: in flight while this row is computed
        if (STRIDE == 1) load_row(seg, t + 2, nx);
        else { load_row(seg, 2 * t + 2, nx); load_row(seg, 2 * t + 3, nx2); }
      }
      const long long r = r0 + (t - t0);
      if ((vbits >> (t - t0)) & 1u) {
        const bool ok0 = STRIDE * t - 1 >= 0, ok2 = STRIDE * t + 1 < p.seg_len_in;      // tap 1 is always inside the segment
        for (int o = 0; o < p.n_out; ++o) {
          const UnavDwLnOut& od = g.out[o];
          const float* wo = dws_smem + o * 5 * C;
          if (od.src >= 0) {
            const float* wp = dws_smem + p.n_out * 5 * C + od.src * 2 * C;
            if (ok0 && ok2) dws_pair<NV, true, false>(w0, w1, w2, wo, wp, true, true, 1.f, p.eps, od, r, lane, p.op_dtype);
            else dws_pair<NV, true, true>(w0, w1, w2, wo, wp, ok0, ok2, 1.f, p.eps, od, r, lane, p.op_dtype);
          } else {
            dws_pair<NV, false, false>(w0, w1, w2, wo, nullptr, true, true, 1.f, p.eps, od, r, lane, p.op_dtype);
          }
        }
      } else {
        masked_row(r);                                // a masked row between valid ones (not a prefix mask): same shortcut
      }
      if (more) {                                     // slide the window
        if (pre) { normalise(nx); if (STRIDE == 2) normalise(nx2); }
#pragma unroll
        for (int j = 0; j < NV; ++j) {
          if (STRIDE == 1) { w0[j] = w1[j]; w1[j] = w2[j]; w2[j] = nx[j]; }
          else { w0[j] = w2[j]; w1[j] = nx[j]; w2[j] = nx2[j]; }
        }
      }
    }
  }
}

// =============================================================================================
// row copy jobs (gather / nearest up-sample / im2col / concat)
// =============================================================================================
struct CopyParams {
  UnavCopyJob j[UNAV_MAX_COPY_JOBS];
  int op_dtype;
};

// IdxT: index arithmetic type.  Every job of the hot path has far fewer than 2^31 float4 elements, and the three 64-bit
// divide / modulo pairs per element were most of the kernel's instructions; 64-bit indices remain for larger jobs.
template <typename IdxT>
__device__ __forceinline__ void rowcopy_body(const UnavCopyJob& jb, int op_dtype, long long total_ll) {
  const IdxT cv = jb.C / 4;
  const IdxT per_row = static_cast<IdxT>(jb.ntaps) * cv;
  const IdxT total = static_cast<IdxT>(total_ll);
  const IdxT seg_out = static_cast<IdxT>(jb.seg_len_out);
  const size_t es = op_elem_size(op_dtype);
  const long long sp = jb.ld_dst / 2;
  for (IdxT i = static_cast<IdxT>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<IdxT>(gridDim.x) * blockDim.x) {
    const IdxT row = i / per_row;
    const int rem = static_cast<int>(i - row * per_row);
    const int tap = rem / static_cast<int>(cv), c = (rem - tap * static_cast<int>(cv)) * 4;
    const IdxT segi = row / seg_out;
    const int seg = static_cast<int>(segi), t = static_cast<int>(row - segi * seg_out);
    const int ti = (t * jb.num) / jb.den + tap - jb.ntaps / 2;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ti >= 0 && ti < jb.seg_len_in)
      v = *reinterpret_cast<const float4*>(jb.src + (static_cast<long long>(seg) * jb.seg_len_in + ti) * jb.ld_src + c);
    const long long drow = static_cast<long long>(seg) * jb.dst_seg_stride + jb.dst_row_off + t;
    char* dr = reinterpret_cast<char*>(jb.dst) + static_cast<size_t>(drow) * jb.ld_dst * es;
    store_op4(dr, op_dtype, static_cast<long long>(tap) * jb.tap_stride + c, sp, v);
  }
}

__global__ void __launch_bounds__(256)
rowcopy_kernel(const __grid_constant__ CopyParams p) {
  pdl_wait();                 // PDL: the preceding grid has completed; nothing above touched global memory
  pdl_launch_dependents();    // let the next kernel's CTAs start their prologue
  const UnavCopyJob& jb = p.j[blockIdx.y];
  const long long total = static_cast<long long>(jb.nseg) * jb.seg_len_out * jb.ntaps * (jb.C / 4);
  if (total + static_cast<long long>(gridDim.x) * blockDim.x < (1LL << 31)) rowcopy_body<unsigned int>(jb, p.op_dtype, total);
  else rowcopy_body<long long>(jb, p.op_dtype, total);
}

// k = 3 im2col of same-length segments as a SCATTER (every job of the launch has ntaps == 3, num == den, C / 4 a power of two
// and a 16-bit operand type): one thread per SOURCE float4 — loaded once, converted to its hi / lo halves once, stored to the
// three (output row, tap) positions that read it (row t+1 tap 0, row t tap 1, row t-1 tap 2), zeros for the taps that fall
// outside the segment.  The gather form above loads and converts every source element three times and pays three integer
// divisions per element (ncu: 68 % issue utilisation, math-pipe throttle, 2.1 TB/s of writes on the head's 87 MB operand).
// grid: (ceil(max seg_len * C/4 / 256), max nseg, njobs)
__global__ void __launch_bounds__(256)
rowcopy_im2col_kernel(const __grid_constant__ CopyParams p) {
  pdl_wait();
  pdl_launch_dependents();
  const UnavCopyJob& jb = p.j[blockIdx.z];
  const int seg = blockIdx.y;
  if (seg >= jb.nseg) return;
  const int cv = jb.C / 4;
  const int shift = 31 - __clz(cv);
  const int i = blockIdx.x * 256 + threadIdx.x;
  const int t = i >> shift, c = (i & (cv - 1)) * 4;
  const int T = jb.seg_len_in;
  if (t >= T) return;
  const float4 v = *reinterpret_cast<const float4*>(jb.src + (static_cast<long long>(seg) * T + t) * jb.ld_src + c);
  const bool split = op_is_split(p.op_dtype);
  const long long sp = jb.ld_dst / 2;
  uint2 pk, pl = make_uint2(0u, 0u);
  if (kHalfF16) {
    const __half2 h01 = __floats2half2_rn(v.x, v.y), h23 = __floats2half2_rn(v.z, v.w);
    pk.x = *reinterpret_cast<const uint32_t*>(&h01); pk.y = *reinterpret_cast<const uint32_t*>(&h23);
    if (split) {
      const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
      const __half2 l01 = __floats2half2_rn(v.x - f01.x, v.y - f01.y), l23 = __floats2half2_rn(v.z - f23.x, v.w - f23.y);
      pl.x = *reinterpret_cast<const uint32_t*>(&l01); pl.y = *reinterpret_cast<const uint32_t*>(&l23);
    }
  } else {
    const __nv_bfloat162 h01 = __floats2bfloat162_rn(v.x, v.y), h23 = __floats2bfloat162_rn(v.z, v.w);
    pk.x = *reinterpret_cast<const uint32_t*>(&h01); pk.y = *reinterpret_cast<const uint32_t*>(&h23);
    if (split) {
      const float2 f01 = __bfloat1622float2(h01), f23 = __bfloat1622float2(h23);
      const __nv_bfloat162 l01 = __floats2bfloat162_rn(v.x - f01.x, v.y - f01.y), l23 = __floats2bfloat162_rn(v.z - f23.x, v.w - f23.y);
      pl.x = *reinterpret_cast<const uint32_t*>(&l01); pl.y = *reinterpret_cast<const uint32_t*>(&l23);
    }
  }
  uint16_t* d0 = reinterpret_cast<uint16_t*>(jb.dst) + (static_cast<long long>(seg) * jb.dst_seg_stride + jb.dst_row_off + t) * jb.ld_dst + c;
  const uint2 zero = make_uint2(0u, 0u);
  auto put = [&](uint16_t* q, uint2 hi, uint2 lo) {
    *reinterpret_cast<uint2*>(q) = hi;
    if (split) *reinterpret_cast<uint2*>(q + sp) = lo;
  };
  put(d0 + jb.tap_stride, pk, pl);                                         // row t, tap 1
  if (t + 1 < T) put(d0 + jb.ld_dst, pk, pl);                              // row t + 1, tap 0
  else put(d0 + 2 * jb.tap_stride, zero, zero);                            // last row: its tap 2 reads past the segment
  if (t > 0) put(d0 - jb.ld_dst + 2 * jb.tap_stride, pk, pl);              // row t - 1, tap 2
  else put(d0, zero, zero);                                                // first row: its tap 0 reads before the segment
}

// =============================================================================================
// batched transpose + cast: in [nb, R, Cc] -> out [nb, Cc, R]
// =============================================================================================
__global__ void __launch_bounds__(256)
transpose_cast_kernel(const float* __restrict__ in, long long ld_in, void* out, long long ld_out, int R, int Cc,
                      int op_dtype) {
  pdl_wait();                 // PDL: the preceding grid has completed; nothing above touched global memory
  pdl_launch_dependents();    // let the next kernel's CTAs start their prologue
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  const float* ib = in + static_cast<long long>(b) * R * ld_in;
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    const int r = r0 + ty + i, c = c0 + tx;
    tile[ty + i][tx] = (r < R && c < Cc) ? ib[static_cast<long long>(r) * ld_in + c] : 0.f;
  }
  __syncthreads();
  const size_t es = op_elem_size(op_dtype);
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    const int c = c0 + ty + i, r = r0 + tx;
    if (c < Cc && r < R) {
      char* orow = reinterpret_cast<char*>(out) + (static_cast<size_t>(b) * Cc + c) * ld_out * es;
      store_op(orow, op_dtype, r, ld_out / 2, tile[tx][ty + i]);
    }
  }
}

// =============================================================================================
// weight packing at load time: [rows, K] FP32 -> operand rows of ld_dst elements (hi | lo halves at ld_dst / 2 for the
// split formats), padding columns zeroed — the whole destination row is written, so the buffer needs no prior memset
// =============================================================================================
__global__ void __launch_bounds__(256)
pack_operand_kernel(const float* __restrict__ src, long long ld_src, void* dst, long long ld_dst, long long rows, int K,
                    int op_dtype) {
  const long long width = op_is_split(op_dtype) ? ld_dst / 2 : ld_dst;     // padded logical width of a row
  const size_t es = op_elem_size(op_dtype);
  const long long total = rows * width;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / width, c = i - r * width;
    const float v = c < K ? __ldg(src + r * ld_src + c) : 0.f;
    store_op(reinterpret_cast<char*>(dst) + static_cast<size_t>(r) * ld_dst * es, op_dtype, c, ld_dst / 2, v);
  }
}

// =============================================================================================
// Alignment token embedding
// =============================================================================================
__global__ void __launch_bounds__(128)
align_embed_kernel(const float* __restrict__ x0, const float* cls_v, const float* cls_a, const float* pos_v,
                   const float* pos_a, const float* type_v, const float* type_a, float* tokens, int nb, int T,
                   int C) {
  pdl_wait();                 // PDL: the preceding grid has completed; nothing above touched global memory
  pdl_launch_dependents();    // let the next kernel's CTAs start their prologue
  // grid: (T+1, nb, 2)
  const int n = blockIdx.x, b = blockIdx.y, m = blockIdx.z;
  const float* cls = m ? cls_a : cls_v;
  const float* pos = (m ? pos_a : pos_v) + static_cast<long long>(n) * C;
  const float* typ = m ? type_a : type_v;
  const float* src = n == 0 ? cls : x0 + ((static_cast<long long>(m) * nb + b) * T + (n - 1)) * C;
  float* dst = tokens + ((static_cast<long long>(m) * nb + b) * (T + 1) + n) * C;
  for (int c = threadIdx.x * 4; c < C; c += blockDim.x * 4) {
    const float4 a = *reinterpret_cast<const float4*>(src + c);
    const float4 p = *reinterpret_cast<const float4*>(pos + c);
    const float4 t = *reinterpret_cast<const float4*>(typ + c);
    float4 y;
    y.x = (a.x + p.x) + t.x; y.y = (a.y + p.y) + t.y; y.z = (a.z + p.z) + t.z; y.w = (a.w + p.w) + t.w;
    *reinterpret_cast<float4*>(dst + c) = y;
  }
}

// =============================================================================================
// pyramid masks
// =============================================================================================
__global__ void build_masks_kernel(const uint8_t* __restrict__ mask, uint8_t* out_true, uint8_t* out_up,
                                   uint8_t* out_cls, uint8_t* out_heads, int nb, int nb_src, int T, int L) {
  pdl_wait();                 // PDL: the preceding grid has completed; nothing above touched global memory
  pdl_launch_dependents();    // let the next kernel's CTAs start their prologue
  long long off = 0, off_up = 0;
  int lvl_off = 0;
  const int Ttot = 2 * T - (T >> (L - 1));
  const long long i0 = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (int l = 0; l < L; ++l) {
    const int Tl = T >> l;
    const long long n = static_cast<long long>(nb) * Tl;
    for (long long i = i0; i < n; i += stride) {
      const int b = static_cast<int>(i / Tl) % nb_src, t = static_cast<int>(i % Tl);
      const uint8_t mv = mask[static_cast<long long>(b) * T + (static_cast<long long>(t) << l)] != 0;
      out_true[off + i] = mv;
      if (out_heads && i < static_cast<long long>(nb_src) * Tl)   // video-major copy for the heads / decode
        out_heads[static_cast<long long>(b) * Ttot + lvl_off + t] = mv;
      if (l + 1 < L)   // mask of level l+1 repeated twice, laid out at level l's resolution
        out_up[off_up + i] = mask[static_cast<long long>(b) * T + (static_cast<long long>(t >> 1) << (l + 1))] != 0;
    }
    off += n;
    lvl_off += Tl;
    if (l + 1 < L) off_up += n;
  }
  if (out_cls) {   // [nb_src, T+1]: a always-valid CLS slot in front of the frame mask (multimodal_backbones.py:1159)
    const long long n = static_cast<long long>(nb_src) * (T + 1);
    for (long long i = i0; i < n; i += stride) {
      const int b = static_cast<int>(i / (T + 1)), t = static_cast<int>(i % (T + 1));
      out_cls[i] = t == 0 ? 1 : (mask[static_cast<long long>(b) * T + t - 1] != 0);
    }
  }
}

// =============================================================================================
// device-side collate: ragged per-video features -> padded batch + validity mask
// =============================================================================================
// ragged = the videos' [C, len_b] row-major feature blocks back to back (offsets[b] = first float of video b).
__global__ void __launch_bounds__(256)
collate_pad_kernel(const float* __restrict__ ragged, const long long* __restrict__ offsets, const int* __restrict__ lens,
                   float* __restrict__ out, uint8_t* __restrict__ mask, int C, int T, float pad) {
  pdl_wait();
  pdl_launch_dependents();
  const int b = blockIdx.y;
  const int len = min(lens[b], T);
  const float* src = ragged + offsets[b];
  float* dst = out + static_cast<long long>(b) * C * T;
  const long long n = static_cast<long long>(C) * T;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i / T), t = static_cast<int>(i % T);
    dst[i] = t < len ? src[static_cast<long long>(c) * lens[b] + t] : pad;
  }
  if (mask && blockIdx.x == 0)
    for (int t = threadIdx.x; t < T; t += blockDim.x) mask[static_cast<long long>(b) * T + t] = t < len ? 1 : 0;
}

// =============================================================================================
// adaptive avg-pool (P bins) of three levels + match projection
// =============================================================================================
// One block = 32 channels of one item, one warp per (level, bin): 3P warps pool concurrently (the one-thread-per-channel
// version walked all T0+T1+T2 frames serially: 95 us for 1.4 MB), then the same warps share the Tq output rows.
// Summation order per bin and per output is unchanged (sequential in t, then j), so results are bit-identical.
__global__ void __launch_bounds__(512)
pool_match_kernel(const float* __restrict__ u0, const float* __restrict__ u1, const float* __restrict__ u2, int T0,
                  int T1, int T2, long long ldu, const float* __restrict__ Wm, const float* __restrict__ bm,
                  float* q, long long ldq, int C, int Tq, int P) {
  pdl_wait();                 // PDL: the preceding grid has completed; nothing above touched global memory
  pdl_launch_dependents();    // let the next kernel's CTAs start their prologue
  extern __shared__ float sm[];          // pooled[3P][32] then Wm[Tq][3P], bm[Tq]
  const int J = 3 * P;
  float* pooled = sm;
  float* ws = sm + J * 32;
  float* bs = ws + Tq * J;
  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  for (int i = threadIdx.x; i < Tq * J; i += blockDim.x) ws[i] = Wm[i];
  for (int i = threadIdx.x; i < Tq; i += blockDim.x) bs[i] = bm[i];
  for (int j = warp; j < J; j += nwarps) {
    const int l = j / P, pb = j % P;
    const float* u = l == 0 ? u0 : (l == 1 ? u1 : u2);
    const int T = l == 0 ? T0 : (l == 1 ? T1 : T2);
    const int s = (pb * T) / P;                 // floor(pb*T/P)
    const int e = ((pb + 1) * T + P - 1) / P;   // ceil((pb+1)*T/P)
    float acc = 0.f;
    if (c < C) {
#pragma unroll 8
      for (int t = s; t < e; ++t) acc += u[(static_cast<long long>(b) * T + t) * ldu + c];
    }
    pooled[j * 32 + lane] = acc / static_cast<float>(e - s);
  }
  __syncthreads();
  if (c >= C) return;
  for (int t = warp; t < Tq; t += nwarps) {
    float acc = 0.f;
    for (int j = 0; j < J; ++j) acc = fmaf(ws[t * J + j], pooled[j * 32 + lane], acc);
    q[(static_cast<long long>(b) * Tq + t) * ldq + c] = acc + bs[t];
  }
}

// =============================================================================================
// MaxSigmoid gate: per (batch, head) [T x hc] . [hc x nwords] -> row max -> sigmoid
// =============================================================================================
constexpr int MS_TR = 32, MS_TN = 64;

__global__ void __launch_bounds__(128)
maxsig_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ G, long long ldg,
              const float* __restrict__ head_bias, float* gate, int T, int nwords, int H, int hc) {
  pdl_wait();                 // PDL: the preceding grid has completed; nothing above touched global memory
  pdl_launch_dependents();    // let the next kernel's CTAs start their prologue
  __shared__ float Xs[64][MS_TR + 4];   // [k][row]
  __shared__ float Gs[64][MS_TN + 4];   // [k][n]
  const int b = blockIdx.z, h = blockIdx.y, t0 = blockIdx.x * MS_TR;
  const int tid = threadIdx.x;
  const int tr = tid / 16, tc = tid % 16;       // rows 4*tr.., words 4*tc..
  for (int i = tid; i < MS_TR * hc; i += 128) {
    const int row = i / hc, k = i % hc;
    const int t = t0 + row;
    Xs[k][row] = t < T ? x[(static_cast<long long>(b) * T + t) * ldx + h * hc + k] : 0.f;
  }
  float mx[4] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};
  for (int n0 = 0; n0 < nwords; n0 += MS_TN) {
    __syncthreads();
    for (int i = tid; i < MS_TN * hc; i += 128) {
      const int n = i / hc, k = i % hc;
      Gs[k][n] = (n0 + n < nwords) ? G[(static_cast<long long>(b) * nwords + n0 + n) * ldg + h * hc + k] : 0.f;
    }
    __syncthreads();
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int k = 0; k < hc; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&Xs[k][tr * 4]);
      const float4 g4 = *reinterpret_cast<const float4*>(&Gs[k][tc * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float gv[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], gv[j], acc[i][j]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (n0 + tc * 4 + j < nwords) mx[i] = fmaxf(mx[i], acc[i][j]);
  }
  // reduce over the 16 threads (tc) sharing a row group: they are 16 consecutive lanes
#pragma unroll
  for (int i = 0; i < 4; ++i) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) mx[i] = fmaxf(mx[i], __shfl_xor_sync(0xffffffffu, mx[i], o));
  }
  if (tc == 0) {
    const float inv = 1.0f / sqrtf(static_cast<float>(hc));
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int t = t0 + tr * 4 + i;
      if (t < T) gate[(static_cast<long long>(b) * T + t) * H + h] = sigmoidf_(mx[i] * inv + head_bias[h]);
    }
  }
}

}  // namespace unav

// =============================================================================================
// C ABI
// =============================================================================================
using namespace unav;

extern "C" int unav_layernorm_rows(const UnavLnGroup* groups, int ngroups, int M, int C, float eps, int act,
                                   int op_dtype, void* stream) {
  UNAV_REQUIRE(groups && ngroups >= 1 && ngroups <= UNAV_MAX_GROUPS, "layernorm_rows: bad group count");
  UNAV_REQUIRE_OP(op_dtype, "layernorm_rows");
  UNAV_REQUIRE(M > 0 && C > 0 && C % 4 == 0 && C <= 128 * MAXV, "layernorm_rows: unsupported C=%d", C);
  LnParams p;
  for (int i = 0; i < ngroups; ++i) {
    p.g[i] = groups[i];
    UNAV_REQUIRE(groups[i].x && groups[i].w && groups[i].b, "layernorm_rows: null input");
    UNAV_REQUIRE(!groups[i].out_im2col || groups[i].edge, "layernorm_rows: im2col output needs edge flags");
    UNAV_REQUIRE(!groups[i].post || groups[i].post_rows > 0, "layernorm_rows: post table needs post_rows");
  }
  p.M = M; p.C = C; p.act = act; p.op_dtype = op_dtype; p.eps = eps;
  dim3 grid((M + 3) / 4, ngroups);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const bool generic_only = getenv("UNAV_LN_GENERIC") != nullptr;               // A/B knob (tests compare the bits)
#define UNAV_LN_CASE(NV, ACT)                                                              \
  if (C == 128 * NV && act == ACT) {                                                       \
    launch_pdl(ln_rows_exact_kernel<NV, ACT>, dim3(grid), dim3(128), 0, st, p);            \
    count_launch();                                                                        \
    return finish_launch("layernorm_rows");                                                \
  }
  if (!generic_only) {
    UNAV_LN_CASE(2, UNAV_ACT_NONE) UNAV_LN_CASE(2, UNAV_ACT_RELU) UNAV_LN_CASE(2, UNAV_ACT_GELU) UNAV_LN_CASE(2, UNAV_ACT_SILU)
    UNAV_LN_CASE(4, UNAV_ACT_NONE) UNAV_LN_CASE(4, UNAV_ACT_RELU) UNAV_LN_CASE(4, UNAV_ACT_GELU) UNAV_LN_CASE(4, UNAV_ACT_SILU)
    UNAV_LN_CASE(8, UNAV_ACT_NONE) UNAV_LN_CASE(8, UNAV_ACT_RELU) UNAV_LN_CASE(8, UNAV_ACT_GELU) UNAV_LN_CASE(8, UNAV_ACT_SILU)
  }
#undef UNAV_LN_CASE
  launch_pdl(ln_rows_kernel, dim3(grid), dim3(128), 0, st, p);
  count_launch();
  return finish_launch("layernorm_rows");
}

extern "C" int unav_dwconv_ln(const UnavDwLnGroup* groups, int ngroups, int nseg, int seg_len_in, int stride,
                              int C, int n_pre, int n_out, float eps, int op_dtype, void* stream) {
  UNAV_REQUIRE(groups && ngroups >= 1 && ngroups <= UNAV_MAX_GROUPS, "dwconv_ln: bad group count");
  UNAV_REQUIRE_OP(op_dtype, "dwconv_ln");
  UNAV_REQUIRE(C % 4 == 0 && C <= 128 * DWV, "dwconv_ln: unsupported C=%d", C);
  UNAV_REQUIRE((stride == 1 || stride == 2) && seg_len_in % stride == 0, "dwconv_ln: bad stride/length");
  UNAV_REQUIRE(n_pre >= 0 && n_pre <= 2 && n_out >= 1 && n_out <= 3, "dwconv_ln: bad n_pre/n_out");
  DwLnParams p;
  for (int i = 0; i < ngroups; ++i) p.g[i] = groups[i];
  p.nseg = nseg; p.seg_len_in = seg_len_in; p.seg_len_out = seg_len_in / stride; p.stride = stride; p.C = C;
  p.n_pre = n_pre; p.n_out = n_out; p.op_dtype = op_dtype; p.eps = eps;
  // Streaming kernel where it wins (scripts/dwconv_probe.py, B200, cold L2, round 2 after the per-tap branches were compiled
  // out): every launch over >= 3584 input rows — the stem's 2 x [3584, 512] x 3 with pre-LayerNorms 55.3 -> 44.9 us, the CSP
  // blocks' [7168, 256] x 3 30.6 -> 26.6, [7168, 512] stride 2 22.5 -> 18.4, [3584, 512] stride 2 14.3 -> 13.1, a tie at
  // [3584, 256] x 3 — and the tiled kernel on the short pyramid levels (<= 1792 rows: 13.2 vs 15.2 us at T = 56, 9 vs 14 below),
  // where one CTA per eight rows spreads a handful of rows over more SMs.  UNAV_DWCONV_STREAM=1 forces it (tests),
  // UNAV_DWCONV_TILED=1 disables it.  Both kernels are latency / issue bound (a warp owns a strip of 4 - 16 rows and a CTA
  // first stages 15 - 39 KB of weights for 32 - 128 rows of data), not bandwidth bound: 1.1 - 1.6 TB/s of algorithmic bytes.
  const char* force_stream = getenv("UNAV_DWCONV_STREAM");
  const bool want_stream = force_stream ? force_stream[0] == '1'
                                        : (static_cast<long long>(nseg) * seg_len_in * ngroups >= 3584);
  if ((C == 256 || C == 512) && want_stream && !getenv("UNAV_DWCONV_TILED")) {
    bool aligned = true;
    for (int i = 0; i < ngroups; ++i) aligned = aligned && groups[i].ldx % 4 == 0 && (reinterpret_cast<uintptr_t>(groups[i].x) & 15) == 0;
    if (aligned) {
      const size_t smem = static_cast<size_t>(5 * n_out + 2 * n_pre) * C * sizeof(float);
      const long long rows = static_cast<long long>(nseg) * p.seg_len_out;
      // strip length: enough warps in flight to cover the memory latency (about 12 warps per SM) before strips get longer
      int strip = static_cast<int>((rows * ngroups + 148 * 12 - 1) / (148 * 12));
      strip = strip < 2 ? 2 : (strip > 16 ? 16 : strip);
      if (const char* env = getenv("UNAV_DWCONV_STRIP")) { const int v = atoi(env); if (v >= 1 && v <= 32) strip = v; }
      const int nw = 8;
      const long long strips = static_cast<long long>(nseg) * ((p.seg_len_out + strip - 1) / strip);
      long long blocks = (strips + nw - 1) / nw;
      long long cap_blocks = 148 * 4;
      if (const char* env = getenv("UNAV_DWCONV_BLOCKS")) { const int v = atoi(env); if (v >= 1) cap_blocks = v; }
      if (blocks > cap_blocks) blocks = cap_blocks;
      cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
      dim3 grid(static_cast<unsigned>(blocks), ngroups);
#define UNAV_DWS_CASE(NVV, STR)                                                                         \
      if (C == 128 * NVV && stride == STR) {                                                            \
        static SmemAttr attr = {};                                                                      \
        if (int rc = ensure_dyn_smem(dwconv_ln_stream_kernel<NVV, STR>, attr, smem, "dwconv_ln")) return rc; \
        launch_pdl(dwconv_ln_stream_kernel<NVV, STR>, dim3(grid), dim3(32 * nw), smem, st, p, strip);   \
        count_launch();                                                                                 \
        return finish_launch("dwconv_ln");                                                              \
      }
      UNAV_DWS_CASE(2, 1) UNAV_DWS_CASE(2, 2) UNAV_DWS_CASE(4, 1) UNAV_DWS_CASE(4, 2)
#undef UNAV_DWS_CASE
    }
  }
  const bool narrow = C <= 256;
  const int R = narrow ? 8 : 4;
  const int tiles_per_seg = (p.seg_len_out + R - 1) / R;
  const int n_in = R * stride + 2;
  size_t smem = (static_cast<size_t>(n_in) * C + 2 * n_in) * sizeof(float);
  if (!narrow) smem += static_cast<size_t>(5 * n_out + 2 * n_pre) * C * sizeof(float);      // the CTA's weights
  const int nwarps = n_in > R * n_out ? n_in : R * n_out;     // one warp per staged row and per (row, output) pair
  static SmemAttr attr = {};
  if (!narrow)
    if (int rc = ensure_dyn_smem(dwconv_ln_kernel<4>, attr, smem, "dwconv_ln")) return rc;
  dim3 grid(static_cast<unsigned>(nseg * tiles_per_seg), ngroups);
  if (narrow)
    launch_pdl(dwconv_ln_kernel<2>, dim3(grid), dim3(32 * nwarps), smem, reinterpret_cast<cudaStream_t>(stream), p);
  else
    launch_pdl(dwconv_ln_kernel<4>, dim3(grid), dim3(32 * nwarps), smem, reinterpret_cast<cudaStream_t>(stream), p);
  count_launch();
  return finish_launch("dwconv_ln");
}

extern "C" int unav_rowcopy(const UnavCopyJob* jobs, int njobs, int op_dtype, void* stream) {
  UNAV_REQUIRE(jobs && njobs >= 1 && njobs <= UNAV_MAX_COPY_JOBS, "rowcopy: bad job count %d", njobs);
  UNAV_REQUIRE_OP(op_dtype, "rowcopy");
  CopyParams p;
  long long maxtotal = 0;
  for (int i = 0; i < njobs; ++i) {
    p.j[i] = jobs[i];
    UNAV_REQUIRE(jobs[i].C % 4 == 0 && jobs[i].den > 0 && jobs[i].ntaps >= 1, "rowcopy: bad job %d", i);
    const long long tot = static_cast<long long>(jobs[i].nseg) * jobs[i].seg_len_out * jobs[i].ntaps * (jobs[i].C / 4);
    maxtotal = tot > maxtotal ? tot : maxtotal;
  }
  p.op_dtype = op_dtype;
  {   // scatter form for k = 3 im2col jobs of same-length segments (see rowcopy_im2col_kernel)
    bool scatter = op_is_16bit(op_dtype) && !getenv("UNAV_ROWCOPY_GATHER");
    int max_seg = 0, max_work = 0;
    for (int i = 0; i < njobs && scatter; ++i) {
      const UnavCopyJob& j = jobs[i];
      const int cv = j.C / 4;
      scatter = j.ntaps == 3 && j.num == j.den && j.seg_len_in == j.seg_len_out && cv > 0 && (cv & (cv - 1)) == 0 &&
                j.ld_src % 4 == 0 && j.ld_dst % 8 == 0 && j.tap_stride % 4 == 0 && (reinterpret_cast<uintptr_t>(j.src) & 15) == 0 &&
                (reinterpret_cast<uintptr_t>(j.dst) & 7) == 0 && static_cast<long long>(j.seg_len_in) * cv < (1LL << 30);
      max_seg = j.nseg > max_seg ? j.nseg : max_seg;
      max_work = j.seg_len_in * cv > max_work ? j.seg_len_in * cv : max_work;
    }
    if (scatter && max_seg > 0 && max_seg <= 65535 && max_work > 0) {
      dim3 grid(static_cast<unsigned>((max_work + 255) / 256), static_cast<unsigned>(max_seg), njobs);
      launch_pdl(rowcopy_im2col_kernel, dim3(grid), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), p);
      count_launch();
      return finish_launch("rowcopy");
    }
  }
  long long blocks = (maxtotal + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  dim3 grid(static_cast<unsigned>(blocks), njobs);
  launch_pdl(rowcopy_kernel, dim3(grid), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), p);
  count_launch();
  return finish_launch("rowcopy");
}

extern "C" int unav_transpose_cast(const float* in, long long ld_in, void* out, long long ld_out, int nb, int R,
                                   int Cc, int op_dtype, void* stream) {
  UNAV_REQUIRE(in && out && nb > 0 && R > 0 && Cc > 0, "transpose_cast: bad arguments");
  UNAV_REQUIRE_OP(op_dtype, "transpose_cast");
  dim3 grid((Cc + 31) / 32, (R + 31) / 32, nb);
  launch_pdl(transpose_cast_kernel, dim3(grid), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), in, ld_in, out, ld_out, R, Cc, op_dtype);
  count_launch();
  return finish_launch("transpose_cast");
}

extern "C" int unav_pack_operand(const float* src, long long ld_src, void* dst, long long ld_dst, long long rows, int K,
                                 int op_dtype, void* stream) {
  UNAV_REQUIRE(src && dst && rows > 0 && K > 0 && ld_src >= K, "pack_operand: bad arguments");
  UNAV_REQUIRE_OP(op_dtype, "pack_operand");
  UNAV_REQUIRE(op_is_split(op_dtype) ? (ld_dst % 2 == 0 && ld_dst / 2 >= K) : ld_dst >= K, "pack_operand: ld_dst %lld too small for K = %d", ld_dst, K);
  const long long total = rows * (op_is_split(op_dtype) ? ld_dst / 2 : ld_dst);
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  launch_pdl(pack_operand_kernel, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), src, ld_src, dst,
             ld_dst, rows, K, op_dtype);
  count_launch();
  return finish_launch("pack_operand");
}

extern "C" int unav_align_embed(const float* x0, const float* cls_v, const float* cls_a, const float* pos_v,
                                const float* pos_a, const float* type_v, const float* type_a, float* tokens,
                                int nb, int T, int C, void* stream) {
  UNAV_REQUIRE(x0 && tokens && C % 4 == 0, "align_embed: bad arguments");
  dim3 grid(T + 1, nb, 2);
  launch_pdl(align_embed_kernel, dim3(grid), dim3(128), 0, reinterpret_cast<cudaStream_t>(stream), x0, cls_v, cls_a, pos_v, pos_a, type_v,
                                                                             type_a, tokens, nb, T, C);
  count_launch();
  return finish_launch("align_embed");
}

extern "C" int unav_build_masks(const uint8_t* mask, uint8_t* out_true, uint8_t* out_up, uint8_t* out_cls,
                                uint8_t* out_heads, int nb, int nb_src, int T, int L, void* stream) {
  UNAV_REQUIRE(mask && out_true && (out_up || L == 1) && L >= 1 && (T % (1 << (L - 1))) == 0 && nb_src >= 1 &&
                   nb % nb_src == 0, "build_masks: bad arguments");
  int blocks = (nb * T + 255) / 256;
  launch_pdl(build_masks_kernel, dim3(blocks), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), mask, out_true, out_up, out_cls,
                                                                               out_heads, nb, nb_src, T, L);
  count_launch();
  return finish_launch("build_masks");
}

extern "C" int unav_pool_match(const float* u0, const float* u1, const float* u2, int T0, int T1, int T2,
                               long long ldu, const float* Wm, const float* bm, float* q, long long ldq, int nb,
                               int C, int Tq, int P, void* stream) {
  UNAV_REQUIRE(u0 && u1 && u2 && Wm && bm && q, "pool_match: null pointer");
  UNAV_REQUIRE(P >= 1 && 3 * P <= 16, "pool_match: %d bins per level not supported", P);
  const size_t smem = (static_cast<size_t>(3 * P) * 32 + static_cast<size_t>(Tq) * 3 * P + Tq) * sizeof(float);
  static SmemAttr attr = {};
  if (int rc = ensure_dyn_smem(pool_match_kernel, attr, smem, "pool_match")) return rc;
  dim3 grid((C + 31) / 32, nb);
  launch_pdl(pool_match_kernel, dim3(grid), dim3(32 * 3 * P), smem, reinterpret_cast<cudaStream_t>(stream), u0, u1, u2, T0, T1, T2, ldu, Wm, bm, q,
                                                                              ldq, C, Tq, P);
  count_launch();
  return finish_launch("pool_match");
}

extern "C" int unav_maxsig_gate(const float* x, long long ldx, const float* G, long long ldg, const float* head_bias,
                                float* gate, int nb, int T, int nwords, int H, int hc, void* stream) {
  UNAV_REQUIRE(x && G && head_bias && gate, "maxsig_gate: null pointer");
  UNAV_REQUIRE(hc >= 1 && hc <= 64, "maxsig_gate: head channels %d > 64", hc);
  dim3 grid((T + MS_TR - 1) / MS_TR, H, nb);
  launch_pdl(maxsig_kernel, dim3(grid), dim3(128), 0, reinterpret_cast<cudaStream_t>(stream), x, ldx, G, ldg, head_bias, gate, T, nwords, H, hc);
  count_launch();
  return finish_launch("maxsig_gate");
}

extern "C" int unav_collate_pad(const float* ragged, const long long* offsets, const int* lens, float* out,
                                uint8_t* mask, int B, int C, int T, float pad, void* stream) {
  UNAV_REQUIRE(ragged && offsets && lens && out && B > 0 && C > 0 && T > 0, "collate_pad: bad arguments");
  const long long n = static_cast<long long>(C) * T;
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  dim3 grid(static_cast<unsigned>(blocks), B);
  launch_pdl(collate_pad_kernel, dim3(grid), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), ragged, offsets, lens, out,
             mask, C, T, pad);
  count_launch();
  return finish_launch("collate_pad");
}
