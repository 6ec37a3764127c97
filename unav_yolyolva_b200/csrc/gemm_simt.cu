// gemm_simt.cu — FP32 CUDA-core GEMM with the shared fused epilogue.
//
// Role: (1) the exact-FP32 mode of the hot path (the <=1e-5 parity gate of SURVEY.md §8d needs FP32
// products; the tcgen05 path reaches it only through the 3-pass split), (2) the on-device reference
// the tcgen05 kernel is validated against, (3) the fallback for shapes the tensor-core kernel does
// not take (K not a multiple of 8, unaligned views).  C[M,N] = epi(A[M,K] . W[N,K]^T).
#include "common.cuh"

namespace unav {

struct SimtGroup {
  const void* A; const void* W;
  long long lda, ldw;
  EpiParams epi;
};
struct SimtParams {
  SimtGroup g[UNAV_MAX_GROUPS];
  int M, N, K, op_dtype, act, res_masked;
};

constexpr int SBM = 64, SBN = 64, SBK = 16;

__global__ void __launch_bounds__(256)
gemm_simt_kernel(const __grid_constant__ SimtParams p) {
  pdl_wait();                 // PDL: the preceding grid has completed; nothing above touched global memory
  pdl_launch_dependents();    // let the next kernel's CTAs start their prologue
  const SimtGroup& g = p.g[blockIdx.z];
  __shared__ float As[SBK][SBM + 4];
  __shared__ float Ws[SBK][SBN + 4];
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;          // 16x16 threads, 4x4 outputs each
  const long long m0 = (long long)blockIdx.x * SBM;
  const int n0 = blockIdx.y * SBN;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const long long a_split = g.lda / 2, w_split = g.ldw / 2;
  const size_t es = op_elem_size(p.op_dtype);
  // each thread loads 4 elements of A and 4 of W per k-tile: row = tid / 4, k = (tid % 4) * 4 + i
  const int lr = tid / 4, lk = (tid % 4) * 4;
  for (int k0 = 0; k0 < p.K; k0 += SBK) {
    {
      long long m = m0 + lr;
      const char* row = reinterpret_cast<const char*>(g.A) + (size_t)m * g.lda * es;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        int k = k0 + lk + i;
        As[lk + i][lr] = (m < p.M && k < p.K) ? load_op(row, p.op_dtype, k, a_split) : 0.f;
      }
      int n = n0 + lr;
      const char* wrow = reinterpret_cast<const char*>(g.W) + (size_t)n * g.ldw * es;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        int k = k0 + lk + i;
        Ws[lk + i][lr] = (n < p.N && k < p.K) ? load_op(wrow, p.op_dtype, k, w_split) : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < SBK; ++kk) {
      float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      float4 w = *reinterpret_cast<const float4*>(&Ws[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
    }
    __syncthreads();
  }
  const long long op_split = g.epi.ld_op / 2;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    long long m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n >= p.N) continue;
      float v = epilogue_value(g.epi, p.act, p.res_masked, m, n, acc[i][j]);
      if (g.epi.out_f32) g.epi.out_f32[m * g.epi.ld_f32 + n] = v;
      if (g.epi.out_op) {
        char* row = reinterpret_cast<char*>(g.epi.out_op) + (size_t)m * g.epi.ld_op * es;
        store_op(row, p.op_dtype, n, op_split, v);
      }
      if (g.epi.out_opT) {
        const int ncols = g.epi.t_ncols > 0 ? g.epi.t_ncols : p.N;
        const int nn = n - g.epi.t_col0;
        if (nn >= 0 && nn < ncols) {
          const long long item = m / g.epi.t_seg, t = m % g.epi.t_seg;
          char* row = reinterpret_cast<char*>(g.epi.out_opT) + (size_t)(item * ncols + nn) * g.epi.ld_opT * es;
          store_op(row, p.op_dtype, t, g.epi.ld_opT / 2, v);
        }
      }
    }
  }
}

static int gemm_simt(const UnavGemmGroup* groups, int ngroups, int M, int N, int K, int op_dtype,
                     int act, int res_masked, cudaStream_t stream) {
  SimtParams p;
  for (int i = 0; i < ngroups; ++i) {
    p.g[i].A = groups[i].A; p.g[i].W = groups[i].W;
    p.g[i].lda = groups[i].lda; p.g[i].ldw = groups[i].ldw;
    p.g[i].epi = make_epi(groups[i]);
  }
  p.M = M; p.N = N; p.K = K; p.op_dtype = op_dtype; p.act = act; p.res_masked = res_masked;
  dim3 grid((M + SBM - 1) / SBM, (N + SBN - 1) / SBN, ngroups);
  launch_pdl(gemm_simt_kernel, dim3(grid), dim3(256), 0, stream, p);
  count_launch();
  return finish_launch("gemm_simt");
}

}  // namespace unav

extern "C" int unav_gemm(const UnavGemmGroup* groups, int ngroups, int M, int N, int K,
                         int op_dtype, int act, int res_masked, int backend, void* stream) {
  using namespace unav;
  UNAV_REQUIRE(groups && ngroups >= 1 && ngroups <= UNAV_MAX_GROUPS, "unav_gemm: bad group count %d", ngroups);
  UNAV_REQUIRE(M > 0 && N > 0 && K > 0, "unav_gemm: bad shape %d %d %d", M, N, K);
  UNAV_REQUIRE(op_base(op_dtype) >= UNAV_F32 && op_base(op_dtype) <= UNAV_F16X2, "unav_gemm: bad op_dtype %d", op_dtype);
  UNAV_REQUIRE_OP(op_dtype, "unav_gemm");
  for (int i = 0; i < ngroups; ++i) {
    UNAV_REQUIRE(groups[i].A && groups[i].W, "unav_gemm: null operand in group %d", i);
    UNAV_REQUIRE(groups[i].out_f32 || groups[i].out_op || groups[i].out_opT, "unav_gemm: group %d has no output", i);
    UNAV_REQUIRE(groups[i].conv_T == 0 || backend == UNAV_GEMM_TCGEN05, "unav_gemm: implicit convolution (conv_T) needs the tcgen05 backend");
  }
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (backend == UNAV_GEMM_TCGEN05)
    return gemm_tcgen05(groups, ngroups, M, N, K, op_dtype, act, res_masked, s);
  return gemm_simt(groups, ngroups, M, N, K, op_base(op_dtype), act, res_masked, s);   // CUDA cores: always hi + lo
}
