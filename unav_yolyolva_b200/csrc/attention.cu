// attention.cu — fused masked attention core (FP32 CUDA-core version, K/V chunks staged in shared memory,
// online softmax, key-validity mask and optional per-query extra key applied in-kernel).
//
// One CTA = (32 queries, one head, one batch item).  Per 64-key chunk:
//   phase 1  S = scale * Q.K^T   (4x4 register tiles, Q and K held transposed in smem), mask -> -inf
//   phase 2  online softmax update of (row max, row sum), P written transposed to smem
//   phase 3  O = alpha*O + P.V    (4 x HS/16 register tiles)
// The Alignment transformer's fused mask (multimodal_backbones.py:1173-1183) is "own-modality valid keys
// + the time-aligned token of the other modality": the latter is the optional extra key xk/xv, handled as
// a final one-key chunk, so the [B,450,450] mask tensor of the reference is never materialised.
#include "common.cuh"

namespace unav {

struct AttnParams {
  UnavAttnGroup g[UNAV_MAX_GROUPS];
  int nb, Tq, Tk, nh, op_dtype;
  float scale;
};

constexpr int AT_Q = 32, AT_K = 64;

template <int HS>
struct AttnSmem {
  float Qs[HS][AT_Q + 4];     // [d][q]
  float Ks[HS][AT_K + 4];     // [d][k]
  float Vs[AT_K][HS + 4];     // [k][d]
  float Pt[AT_K][AT_Q + 4];   // [k][q]
  float row_m[AT_Q], row_l[AT_Q], row_a[AT_Q], px[AT_Q];
};

template <int HS>
__global__ void __launch_bounds__(128)
attention_kernel(const __grid_constant__ AttnParams p) {
  pdl_wait();                 // PDL: the preceding grid has completed; nothing above touched global memory
  pdl_launch_dependents();    // let the next kernel's CTAs start their prologue
  extern __shared__ __align__(16) uint8_t smem_raw[];
  AttnSmem<HS>& sm = *reinterpret_cast<AttnSmem<HS>*>(smem_raw);
  constexpr int DPT = HS / 16;
  const int gi = blockIdx.z / p.nb, b = blockIdx.z % p.nb;
  const UnavAttnGroup& g = p.g[gi];
  const int h = blockIdx.y, q0 = blockIdx.x * AT_Q;
  const int tid = threadIdx.x;
  const int tq = tid / 16, t16 = tid % 16;

  // load Q tile transposed
  for (int i = tid; i < AT_Q * HS; i += 128) {
    const int q = i / HS, d = i % HS;
    const int qi = q0 + q;
    sm.Qs[d][q] = qi < p.Tq ? g.q[(static_cast<long long>(b) * p.Tq + qi) * g.ldq + h * HS + d] : 0.f;
  }
  if (tid < AT_Q) { sm.row_m[tid] = -CUDART_INF_F; sm.row_l[tid] = 0.f; }
  float acc[4][DPT];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < DPT; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < p.Tk; k0 += AT_K) {
    __syncthreads();   // previous chunk fully consumed (also covers the Q load on the first pass)
    for (int i = tid; i < AT_K * HS; i += 128) {
      const int k = i / HS, d = i % HS;
      const int kj = k0 + k;
      float kv = 0.f, vv = 0.f;
      if (kj < p.Tk) {
        const long long row = static_cast<long long>(b) * p.Tk + kj;
        kv = g.k[row * g.ldk + h * HS + d];
        vv = g.v[row * g.ldv + h * HS + d];
      }
      sm.Ks[d][k] = kv;
      sm.Vs[k][d] = vv;
    }
    __syncthreads();
    // ---- phase 1: 4 queries x 4 keys per thread
    {
      float s[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
#pragma unroll 8
      for (int d = 0; d < HS; ++d) {
        const float4 a = *reinterpret_cast<const float4*>(&sm.Qs[d][tq * 4]);
        const float4 k4 = *reinterpret_cast<const float4*>(&sm.Ks[d][t16 * 4]);
        const float av[4] = {a.x, a.y, a.z, a.w};
        const float kv[4] = {k4.x, k4.y, k4.z, k4.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) s[i][j] = fmaf(av[i], kv[j], s[i][j]);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int kj = k0 + t16 * 4 + j;
        const bool valid = kj < p.Tk && (!g.kmask || g.kmask[static_cast<long long>(b) * p.Tk + kj]);
#pragma unroll
        for (int i = 0; i < 4; ++i) sm.Pt[t16 * 4 + j][tq * 4 + i] = valid ? s[i][j] * p.scale : -CUDART_INF_F;
      }
    }
    __syncthreads();
    // ---- phase 2: 4 threads per query row
    {
      const int q = tid / 4, sub = tid % 4;
      float mx = -CUDART_INF_F;
      for (int k = sub; k < AT_K; k += 4) mx = fmaxf(mx, sm.Pt[k][q]);
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      const float m_old = sm.row_m[q];
      const float m_new = fmaxf(m_old, mx);
      float sum = 0.f;
      for (int k = sub; k < AT_K; k += 4) {
        const float sv = sm.Pt[k][q];
        const float pv = (m_new == -CUDART_INF_F) ? 0.f : expf(sv - m_new);
        sm.Pt[k][q] = pv;
        sum += pv;
      }
      sum += __shfl_xor_sync(0xffffffffu, sum, 1);
      sum += __shfl_xor_sync(0xffffffffu, sum, 2);
      if (sub == 0) {
        const float alpha = (m_old == -CUDART_INF_F) ? 0.f : expf(m_old - m_new);
        sm.row_a[q] = alpha;
        sm.row_l[q] = sm.row_l[q] * alpha + sum;
        sm.row_m[q] = m_new;
      }
    }
    __syncthreads();
    // ---- phase 3: 4 queries x DPT dims per thread
    {
      const float4 al = *reinterpret_cast<const float4*>(&sm.row_a[tq * 4]);
      const float alv[4] = {al.x, al.y, al.z, al.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < DPT; ++j) acc[i][j] *= alv[i];
      const int kmax = min(AT_K, p.Tk - k0);
      for (int k = 0; k < kmax; ++k) {
        const float4 p4 = *reinterpret_cast<const float4*>(&sm.Pt[k][tq * 4]);
        const float pv[4] = {p4.x, p4.y, p4.z, p4.w};
        float vv[DPT];
#pragma unroll
        for (int j = 0; j < DPT; j += 4) {
          const float4 v4 = *reinterpret_cast<const float4*>(&sm.Vs[k][t16 * DPT + j]);
          vv[j] = v4.x; vv[j + 1] = v4.y; vv[j + 2] = v4.z; vv[j + 3] = v4.w;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < DPT; ++j) acc[i][j] = fmaf(pv[i], vv[j], acc[i][j]);
      }
    }
  }
  // ---- optional extra key (one per query)
  if (g.xk) {
    __syncthreads();
    if (tid < AT_Q) {
      const int qi = q0 + tid;
      float pxv = 0.f, alpha = 1.f;
      if (qi < p.Tq && qi >= g.x_first) {
        const float* kr = g.xk + (static_cast<long long>(b) * p.Tq + qi) * g.ldx + h * HS;
        float s = 0.f;
        for (int d = 0; d < HS; ++d) s = fmaf(sm.Qs[d][tid], kr[d], s);
        s *= p.scale;
        const float m_old = sm.row_m[tid];
        const float m_new = fmaxf(m_old, s);
        alpha = (m_old == -CUDART_INF_F) ? 0.f : expf(m_old - m_new);
        pxv = expf(s - m_new);
        sm.row_l[tid] = sm.row_l[tid] * alpha + pxv;
        sm.row_m[tid] = m_new;
      }
      sm.row_a[tid] = alpha;
      sm.px[tid] = pxv;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int q = tq * 4 + i, qi = q0 + q;
      const float alpha = sm.row_a[q], pxv = sm.px[q];
      if (pxv != 0.f || alpha != 1.f) {
        const float* vr = g.xv + (static_cast<long long>(b) * p.Tq + min(qi, p.Tq - 1)) * g.ldx + h * HS + t16 * DPT;
#pragma unroll
        for (int j = 0; j < DPT; ++j) acc[i][j] = acc[i][j] * alpha + pxv * vr[j];
      }
    }
  }
  __syncthreads();
  // ---- normalise and store
  const size_t es = op_elem_size(p.op_dtype);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int q = tq * 4 + i, qi = q0 + q;
    if (qi >= p.Tq) continue;
    const float inv = 1.0f / sm.row_l[q];
    char* orow = reinterpret_cast<char*>(g.out) + (static_cast<size_t>(b) * p.Tq + qi) * g.ldo * es;
#pragma unroll
    for (int j = 0; j < DPT; j += 4)
      store_op4(orow, p.op_dtype, h * HS + t16 * DPT + j, g.ldo / 2,
                make_float4(acc[i][j] * inv, acc[i][j + 1] * inv, acc[i][j + 2] * inv, acc[i][j + 3] * inv));
  }
}

template <int HS>
static int launch_attention(const AttnParams& p, int ngroups, cudaStream_t stream) {
  static SmemAttr attr = {};
  const int smem = static_cast<int>(sizeof(AttnSmem<HS>));
  if (int rc = ensure_dyn_smem(attention_kernel<HS>, attr, smem, "attention")) return rc;
  dim3 grid((p.Tq + AT_Q - 1) / AT_Q, p.nh, p.nb * ngroups);
  launch_pdl(attention_kernel<HS>, dim3(grid), dim3(128), smem, stream, p);
  count_launch();
  return finish_launch("attention");
}

}  // namespace unav

extern "C" int unav_attention(const UnavAttnGroup* groups, int ngroups, int nb, int Tq, int Tk, int nh, int hs,
                              float scale, int op_dtype, void* stream) {
  using namespace unav;
  UNAV_REQUIRE(groups && ngroups >= 1 && ngroups <= UNAV_MAX_GROUPS, "attention: bad group count");
  UNAV_REQUIRE(hs == 64 || hs == 128, "attention: head size %d not in {64,128}", hs);
  UNAV_REQUIRE_OP(op_dtype, "attention");
  UNAV_REQUIRE(nb > 0 && Tq > 0 && Tk > 0 && nh > 0, "attention: bad shape");
  AttnParams p;
  for (int i = 0; i < ngroups; ++i) {
    p.g[i] = groups[i];
    UNAV_REQUIRE(groups[i].q && groups[i].k && groups[i].v && groups[i].out, "attention: null pointer");
    UNAV_REQUIRE(!groups[i].xk || (groups[i].xv && Tq == Tk), "attention: extra key needs xv and Tq == Tk");
  }
  p.nb = nb; p.Tq = Tq; p.Tk = Tk; p.nh = nh; p.op_dtype = op_dtype; p.scale = scale;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  return hs == 64 ? launch_attention<64>(p, ngroups, s) : launch_attention<128>(p, ngroups, s);
}
