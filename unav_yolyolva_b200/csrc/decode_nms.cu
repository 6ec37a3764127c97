// decode_nms.cu — candidate decoding and per-class temporal soft-NMS, device resident.
//
// Replaces the reference's per-video / per-level Python decode loop
// (libs/modeling/multimodal_meta_archs.py:745-817), the per-class Python loop + CPU extension of
// libs/utils/nms.py:103-190 / libs/utils/csrc/nms_cpu.cpp:67-160 and the seconds conversion
// (multimodal_meta_archs.py:852-856).  Integer / compare work is bit-exact w.r.t. the reference:
//   * top-k selection is a radix select on the score bits (ties: lower flat index),
//   * IoU / decay arithmetic uses explicit round-to-nearest intrinsics (no FMA contraction),
//   * the gaussian weight uses glibc's expf algorithm (exp2f_data table, double arithmetic), which
//     reproduces libm bit-for-bit on [-1/sigma, 0].
#include <stdlib.h>

#include "common.cuh"

namespace unav {

// =============================================================================================
// decode
// =============================================================================================
constexpr int DEC_THREADS = 512;

struct DecodeParams {
  const float* logits; const float* offsets; const uint8_t* masks; const float* points;
  float* cand_segs; float* cand_scores; int32_t* cand_labels;
  int level_off[9];     // rows
  int cap_off[9];       // candidate slots
  int B, L, ncls, class_aware, topk, cap, Ttot, use_smem;
  float thresh, dur_thresh;
};

__device__ __forceinline__ float decode_prob(const DecodeParams& p, int b, int row0, int i) {
  const int row = row0 + i / p.ncls;
  const long long r = static_cast<long long>(b) * p.Ttot + row;
  const float m = p.masks[r] ? 1.f : 0.f;
  return sigmoidf_(p.logits[r * p.ncls + (i % p.ncls)]) * m;
}

// block-wide exclusive scan of a 0/1 flag in index order; returns this thread's offset, total in *total
__device__ __forceinline__ int block_scan_flag(bool flag, int* warp_tot, int* total) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const unsigned bal = __ballot_sync(0xffffffffu, flag);
  const int within = __popc(bal & ((1u << lane) - 1u));
  if (lane == 0) warp_tot[w] = __popc(bal);
  __syncthreads();
  int base = 0, tot = 0;
  const int nw = blockDim.x >> 5;
  for (int i = 0; i < nw; ++i) {
    const int c = warp_tot[i];
    if (i < w) base += c;
    tot += c;
  }
  __syncthreads();
  *total = tot;
  return base + within;
}

// block-wide exclusive scan of a small per-thread count in thread order; returns this thread's offset, total in *total
__device__ __forceinline__ int block_scan_cnt(int v, int* warp_tot, int* total) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) warp_tot[w] = incl;
  __syncthreads();
  int base = 0, tot = 0;
  const int nw = blockDim.x >> 5;
  for (int i = 0; i < nw; ++i) {
    const int c = warp_tot[i];
    if (i < w) base += c;
    tot += c;
  }
  __syncthreads();
  *total = tot;
  return base + incl - v;
}

constexpr int DEC_E = 4;      // consecutive candidates per thread and emit round (one block scan per DEC_THREADS*DEC_E)

__global__ void __launch_bounds__(DEC_THREADS)
decode_kernel(const __grid_constant__ DecodeParams p) {
  pdl_wait();                 // PDL: the preceding grid has completed; nothing above touched global memory
  pdl_launch_dependents();    // let the next kernel's CTAs start their prologue
  extern __shared__ __align__(16) float dec_p[];      // [n] probabilities of this (video, level), computed once
  __shared__ int hist[256];
  __shared__ int warp_tot[DEC_THREADS / 32];
  __shared__ unsigned s_prefix;
  __shared__ int s_need, s_count;
  const int l = blockIdx.x, b = blockIdx.y;
  const int row0 = p.level_off[l];
  const int n = (p.level_off[l + 1] - row0) * p.ncls;
  const int slot0 = p.cap_off[l], nslots = p.cap_off[l + 1] - p.cap_off[l];
  const int tid = threadIdx.x;

  // ---- count candidates above the threshold
  if (tid == 0) s_count = 0;
  __syncthreads();
  int cnt = 0;
  for (int i = tid; i < n; i += DEC_THREADS) {
    const float pr = decode_prob(p, b, row0, i);
    if (p.use_smem) dec_p[i] = pr;
    cnt += pr > p.thresh;
  }
  cnt = __reduce_add_sync(0xffffffffu, cnt);
  if ((tid & 31) == 0 && cnt) atomicAdd(&s_count, cnt);
  __syncthreads();
  const int count = s_count;

  // ---- radix select of the topk-th largest score (only when more than topk pass the threshold)
  unsigned vstar = 0;       // key of the k-th largest; everything strictly above is taken
  int need_eq = 0x7fffffff; // how many elements with key == vstar are taken (index order)
  if (count > p.topk) {
    if (tid == 0) { s_prefix = 0; s_need = p.topk; }
    for (int pass = 0; pass < 4; ++pass) {
      for (int i = tid; i < 256; i += DEC_THREADS) hist[i] = 0;
      __syncthreads();
      const unsigned prefix = s_prefix;
      const int shift = 24 - 8 * pass;
      for (int i = tid; i < n; i += DEC_THREADS) {
        const float pr = p.use_smem ? dec_p[i] : decode_prob(p, b, row0, i);
        if (pr > p.thresh) {
          const unsigned key = __float_as_uint(pr);
          if (pass == 0 || (key >> (shift + 8)) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1);
        }
      }
      __syncthreads();
      if (tid == 0) {
        int need = s_need, bin = 255;
        for (; bin > 0; --bin) {
          if (hist[bin] >= need) break;
          need -= hist[bin];
        }
        s_need = need;
        s_prefix = (prefix << 8) | static_cast<unsigned>(bin);
      }
      __syncthreads();
    }
    vstar = s_prefix;
    need_eq = s_need;
  }

  // ---- emit in flat-index order: DEC_E consecutive candidates per thread, two block scans per round
  int eq_seen = 0, out_pos = 0;
  for (int i0 = 0; i0 < n; i0 += DEC_THREADS * DEC_E) {
    const int ib = i0 + tid * DEC_E;
    float pr[DEC_E];
    bool above[DEC_E], eq[DEC_E];
    int eq_cnt = 0;
#pragma unroll
    for (int e = 0; e < DEC_E; ++e) {
      const int i = ib + e;
      pr[e] = 0.f; above[e] = false; eq[e] = false;
      if (i < n) {
        pr[e] = p.use_smem ? dec_p[i] : decode_prob(p, b, row0, i);
        if (pr[e] > p.thresh) {
          const unsigned key = __float_as_uint(pr[e]);
          above[e] = (count <= p.topk) || key > vstar;
          eq[e] = (count > p.topk) && key == vstar;
        }
      }
      eq_cnt += eq[e];
    }
    int eq_tot;
    int eq_rank = eq_seen + block_scan_cnt(eq_cnt, warp_tot, &eq_tot);
    eq_seen += eq_tot;
    float left[DEC_E], right[DEC_E];
    bool keep[DEC_E];
    int keep_cnt = 0;
#pragma unroll
    for (int e = 0; e < DEC_E; ++e) {
      const bool sel = above[e] || (eq[e] && eq_rank < need_eq);
      eq_rank += eq[e];
      keep[e] = false;
      if (sel) {
        const int i = ib + e;
        const int row = row0 + i / p.ncls;
        const int cls = i % p.ncls;
        const long long r = static_cast<long long>(b) * p.Ttot + row;
        const float* off = p.class_aware ? p.offsets + (r * p.ncls + cls) * 2 : p.offsets + r * 2;
        const float t = p.points[row * 4 + 0], st = p.points[row * 4 + 3];
        left[e] = __fsub_rn(t, __fmul_rn(off[0], st));
        right[e] = __fadd_rn(t, __fmul_rn(off[1], st));
        keep[e] = __fsub_rn(right[e], left[e]) > p.dur_thresh;
      }
      keep_cnt += keep[e];
    }
    int keep_tot;
    int pos = out_pos + block_scan_cnt(keep_cnt, warp_tot, &keep_tot);
#pragma unroll
    for (int e = 0; e < DEC_E; ++e) {
      if (keep[e]) {
        const long long s = static_cast<long long>(b) * p.cap + slot0 + pos;
        p.cand_segs[s * 2 + 0] = left[e];
        p.cand_segs[s * 2 + 1] = right[e];
        p.cand_scores[s] = pr[e];
        p.cand_labels[s] = (ib + e) % p.ncls;
        ++pos;
      }
    }
    out_pos += keep_tot;
  }
  // ---- mark unused slots of this (video, level) region
  for (int s = out_pos + tid; s < nslots; s += DEC_THREADS) {
    const long long g = static_cast<long long>(b) * p.cap + slot0 + s;
    p.cand_labels[g] = -1;
    p.cand_scores[g] = 0.f;
    p.cand_segs[g * 2] = 0.f;
    p.cand_segs[g * 2 + 1] = 0.f;
  }
}

// =============================================================================================
// soft-NMS
// =============================================================================================
__constant__ unsigned long long kExp2fTab[32] = {
    0x3ff0000000000000ULL, 0x3fefd9b0d3158574ULL, 0x3fefb5586cf9890fULL, 0x3fef9301d0125b51ULL,
    0x3fef72b83c7d517bULL, 0x3fef54873168b9aaULL, 0x3fef387a6e756238ULL, 0x3fef1e9df51fdee1ULL,
    0x3fef06fe0a31b715ULL, 0x3feef1a7373aa9cbULL, 0x3feedea64c123422ULL, 0x3feece086061892dULL,
    0x3feebfdad5362a27ULL, 0x3feeb42b569d4f82ULL, 0x3feeab07dd485429ULL, 0x3feea47eb03a5585ULL,
    0x3feea09e667f3bcdULL, 0x3fee9f75e8ec5f74ULL, 0x3feea11473eb0187ULL, 0x3feea589994cce13ULL,
    0x3feeace5422aa0dbULL, 0x3feeb737b0cdc5e5ULL, 0x3feec49182a3f090ULL, 0x3feed503b23e255dULL,
    0x3feee89f995ad3adULL, 0x3feeff76f2fb5e47ULL, 0x3fef199bdd85529cULL, 0x3fef3720dcef9069ULL,
    0x3fef5818dcfba487ULL, 0x3fef7c97337b9b5fULL, 0x3fefa4afa2a490daULL, 0x3fefd0765b6e4540ULL};

// glibc expf (sysdeps/ieee754/flt-32/e_expf.c, EXP2F_TABLE_BITS = 5) for |x| small enough that no
// overflow/underflow path is taken (here x in [-1/sigma, 0]).  Double arithmetic, explicit rounding.
__device__ __forceinline__ float expf_glibc(float x, const unsigned long long* tab) {
  const double InvLn2N = 0x1.71547652b82fep+0 * 32.0;
  const double Shift = 0x1.8p+52;
  const double C0 = 0x1.c6af84b912394p-5 / 32.0 / 32.0 / 32.0;
  const double C1 = 0x1.ebfce50fac4f3p-3 / 32.0 / 32.0;
  const double C2 = 0x1.62e42ff0c52d6p-1 / 32.0;
  const double xd = static_cast<double>(x);
  const double z = __dmul_rn(InvLn2N, xd);
  double kd = __dadd_rn(z, Shift);
  const unsigned long long ki = static_cast<unsigned long long>(__double_as_longlong(kd));
  kd = __dsub_rn(kd, Shift);
  const double r = __dsub_rn(z, kd);
  unsigned long long t = tab[ki & 31ULL];
  t += ki << (52 - 5);
  const double s = __longlong_as_double(static_cast<long long>(t));
  const double zz = __dadd_rn(__dmul_rn(C0, r), C1);
  const double r2 = __dmul_rn(r, r);
  double y = __dadd_rn(__dmul_rn(C2, r), 1.0);
  y = __dadd_rn(__dmul_rn(zz, r2), y);
  y = __dmul_rn(y, s);
  return static_cast<float>(y);
}

struct NmsParams {
  const float* cand_segs; const float* cand_scores; const int32_t* cand_labels;
  float* ws_dets;      // [B, ncls, max_seg, 3]
  int32_t* ws_counts;  // [B, ncls]
  int B, cap, ncls, method, max_seg, maxn;
  float iou_thr, sigma, min_score;
};

// One block per (class, video).  Candidates of the class are gathered in slot order into shared memory;
// each round takes the best live one (ties: lowest slot), emits it and decays the rest.
__global__ void softnms_kernel(const __grid_constant__ NmsParams p) {
  pdl_wait();                 // PDL: the preceding grid has completed; nothing above touched global memory
  pdl_launch_dependents();    // let the next kernel's CTAs start their prologue
  extern __shared__ __align__(16) uint8_t nms_smem[];
  float* x1 = reinterpret_cast<float*>(nms_smem);
  float* x2 = x1 + p.maxn;
  float* ar = x2 + p.maxn;
  float* sc = ar + p.maxn;
  __shared__ int s_n;
  __shared__ float red_s[8];
  __shared__ int red_i[8];
  __shared__ int s_win;
  __shared__ unsigned long long s_tab[32];     // exp2f table in smem: lanes index it divergently
  const int c = blockIdx.x, b = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  if (tid < 32) s_tab[tid] = kExp2fTab[tid];

  // ---- gather (warp 0, ordered)
  if (warp == 0) {
    int n = 0;
    const long long base = static_cast<long long>(b) * p.cap;
    for (int s0 = 0; s0 < p.cap; s0 += 128) {
      int lab[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {            // four independent label loads in flight per lane
        const int s = s0 + u * 32 + lane;
        lab[u] = s < p.cap ? p.cand_labels[base + s] : -1;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int s = s0 + u * 32 + lane;
        bool hit = lab[u] == c;
        float sv = 0.f;
        if (hit) {
          sv = p.cand_scores[base + s];
          if (p.method == 3 && !(sv > p.min_score)) hit = false;   // NMSop pre-filter (nms.py:15-19)
        }
        const unsigned bal = __ballot_sync(0xffffffffu, hit);
        const int pos = n + __popc(bal & ((1u << lane) - 1u));
        if (hit && pos < p.maxn) {
          const float a = p.cand_segs[(base + s) * 2], e = p.cand_segs[(base + s) * 2 + 1];
          x1[pos] = a; x2[pos] = e;
          ar[pos] = __fadd_rn(__fsub_rn(e, a), 1e-6f);
          sc[pos] = sv;
        }
        n += __popc(bal);
      }
    }
    if (lane == 0) s_n = n < p.maxn ? n : p.maxn;
  }
  __syncthreads();
  const int n = s_n;
  float* dets = p.ws_dets + (static_cast<long long>(b) * p.ncls + c) * p.max_seg * 3;
  int emitted = 0;
  const int rounds = n < p.max_seg ? n : p.max_seg;
  for (int r = 0; r < rounds; ++r) {
    // ---- argmax over live candidates
    float bs = -CUDART_INF_F;     // dead candidates carry -inf
    int bi = 0x7fffffff;
    for (int j = tid; j < n; j += blockDim.x) {
      const float s = sc[j];
      if (s > bs) { bs = s; bi = j; }       // strict > keeps the lowest index within a thread's stride order
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float os = __shfl_xor_sync(0xffffffffu, bs, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (os > bs || (os == bs && oi < bi)) { bs = os; bi = oi; }
    }
    if (nwarps > 1) {
      if (lane == 0) { red_s[warp] = bs; red_i[warp] = bi; }
      __syncthreads();
      if (tid == 0) {
        for (int w = 1; w < nwarps; ++w)
          if (red_s[w] > bs || (red_s[w] == bs && red_i[w] < bi)) { bs = red_s[w]; bi = red_i[w]; }
        s_win = bi;
      }
      __syncthreads();
      bi = s_win;
    }
    if (bi == 0x7fffffff) break;
    const float ix1 = x1[bi], ix2 = x2[bi], ia = ar[bi], is = sc[bi];
    if (nwarps > 1) __syncthreads();   // everyone has read the winner before it is retired
    else __syncwarp();
    if (tid == 0) {
      dets[r * 3 + 0] = ix1; dets[r * 3 + 1] = ix2; dets[r * 3 + 2] = is;
      sc[bi] = -CUDART_INF_F;
    }
    emitted = r + 1;
    // ---- decay the others
    for (int j = tid; j < n; j += blockDim.x) {
      if (j == bi) continue;
      float s = sc[j];
      if (s == -CUDART_INF_F) continue;
      const float xx1 = fmaxf(ix1, x1[j]);
      const float xx2 = fminf(ix2, x2[j]);
      const float inter = fmaxf(0.f, __fsub_rn(xx2, xx1));
      const float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(ia, ar[j]), inter));
      float w = 1.f;
      if (p.method == 0 || p.method == 3) { if (ovr >= p.iou_thr) w = 0.f; }
      else if (p.method == 1) { if (ovr >= p.iou_thr) w = __fsub_rn(1.f, ovr); }
      else if (inter > 0.f) w = expf_glibc(__fdiv_rn(-__fmul_rn(ovr, ovr), p.sigma), s_tab);   // exp(-0) == 1 exactly
      s = __fmul_rn(s, w);
      if (p.method == 3) { if (w == 0.f) s = -CUDART_INF_F; }   // hard NMS: suppressed, scores otherwise untouched
      else if (s < p.min_score) s = -CUDART_INF_F;
      sc[j] = s;
    }
    if (nwarps > 1) __syncthreads();
    else __syncwarp();
  }
  if (tid == 0) p.ws_counts[b * p.ncls + c] = emitted;
}

// =============================================================================================
// per-video merge of the (sorted) per-class lists + seconds conversion
// =============================================================================================
struct MergeParams {
  const float* ws_dets; const int32_t* ws_counts; const float* vid_meta;
  float* out_segs; float* out_scores; int64_t* out_labels; int32_t* out_counts;
  int ncls, max_seg;
};

__global__ void __launch_bounds__(256)
merge_kernel(const __grid_constant__ MergeParams p) {
  pdl_wait();                 // PDL: the preceding grid has completed; nothing above touched global memory
  pdl_launch_dependents();    // let the next kernel's CTAs start their prologue
  __shared__ float red_s[8];
  __shared__ int red_c[8];
  __shared__ int s_win;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  // thread t owns classes t, t + blockDim, ...  (ncls <= 4 * blockDim supported through the loop below)
  constexpr int OWN = 4;
  int head[OWN], cnt[OWN];
  float cur[OWN];                                  // score at the head of each owned class list
#pragma unroll
  for (int o = 0; o < OWN; ++o) {
    const int c = tid + o * blockDim.x;
    head[o] = 0;
    cnt[o] = c < p.ncls ? p.ws_counts[b * p.ncls + c] : 0;
    cur[o] = cnt[o] > 0 ? p.ws_dets[((static_cast<long long>(b) * p.ncls + c) * p.max_seg) * 3 + 2] : 0.f;
  }
  float stride = 1.f, half = 0.f, fps = 1.f, dur = 0.f;
  if (p.vid_meta) {
    stride = p.vid_meta[b * 4 + 0];
    half = 0.5f * p.vid_meta[b * 4 + 1];
    fps = p.vid_meta[b * 4 + 2];
    dur = p.vid_meta[b * 4 + 3];
  }
  int produced = 0;
  for (int r = 0; r < p.max_seg; ++r) {
    float bs = -CUDART_INF_F;
    int bc = 0x7fffffff;
#pragma unroll
    for (int o = 0; o < OWN; ++o) {
      const int c = tid + o * blockDim.x;
      if (head[o] < cnt[o]) {
        const float s = cur[o];
        if (s > bs || (s == bs && c < bc)) { bs = s; bc = c; }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float os = __shfl_xor_sync(0xffffffffu, bs, o);
      const int oc = __shfl_xor_sync(0xffffffffu, bc, o);
      if (os > bs || (os == bs && oc < bc)) { bs = os; bc = oc; }
    }
    if (lane == 0) { red_s[warp] = bs; red_c[warp] = bc; }
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < nwarps; ++w)
        if (red_s[w] > bs || (red_s[w] == bs && red_c[w] < bc)) { bs = red_s[w]; bc = red_c[w]; }
      s_win = bc;
    }
    __syncthreads();
    const int win = s_win;
    if (win == 0x7fffffff) break;
    const int o = win / blockDim.x;
    if (static_cast<int>(win % blockDim.x) == tid) {
      int hd = 0;
#pragma unroll
      for (int k = 0; k < OWN; ++k) if (k == o) { hd = head[k]; head[k]++; }
      const float* d = p.ws_dets + ((static_cast<long long>(b) * p.ncls + win) * p.max_seg + hd) * 3;
#pragma unroll
      for (int k = 0; k < OWN; ++k)
        if (k == o && head[k] < cnt[k]) cur[k] = d[3 + 2];            // next entry of this class list
      float s0 = d[0], s1 = d[1];
      if (p.vid_meta) {
        // (segs * stride + 0.5 * nframes) / fps, then clamp to [0, duration] the way the reference does
        s0 = __fdiv_rn(__fadd_rn(__fmul_rn(s0, stride), half), fps);
        s1 = __fdiv_rn(__fadd_rn(__fmul_rn(s1, stride), half), fps);
        if (s0 <= 0.f) s0 = __fmul_rn(s0, 0.f);
        if (s1 <= 0.f) s1 = __fmul_rn(s1, 0.f);
        if (s0 >= dur) s0 = __fadd_rn(__fmul_rn(s0, 0.f), dur);
        if (s1 >= dur) s1 = __fadd_rn(__fmul_rn(s1, 0.f), dur);
      }
      const long long orow = static_cast<long long>(b) * p.max_seg + r;
      p.out_segs[orow * 2] = s0;
      p.out_segs[orow * 2 + 1] = s1;
      p.out_scores[orow] = d[2];
      p.out_labels[orow] = win;
    }
    produced = r + 1;
    __syncthreads();
  }
  for (int r = produced + tid; r < p.max_seg; r += blockDim.x) {
    const long long orow = static_cast<long long>(b) * p.max_seg + r;
    p.out_segs[orow * 2] = 0.f; p.out_segs[orow * 2 + 1] = 0.f;
    p.out_scores[orow] = 0.f; p.out_labels[orow] = 0;
  }
  if (tid == 0) p.out_counts[b] = produced;
}


// =============================================================================================
// lazy per-video soft-NMS: exact, but only the rounds the global top-K needs
// =============================================================================================
// The reference runs every class to completion (<= max_seg rounds each) and then keeps the global top max_seg.
// A class's emissions are a deterministic non-increasing sequence that does not depend on other classes, and the
// global top-K is the K-way merge of those sequences, so it is enough to always advance the class whose NEXT
// emission is the largest: max_seg decay rounds per video instead of ncls * max_seg.  One block per video; all
// candidates of the video live in shared memory grouped by class.
struct LazyNmsParams {
  const float* cand_segs; const float* cand_scores; const int32_t* cand_labels; const float* vid_meta;
  float* out_segs; float* out_scores; int64_t* out_labels; int32_t* out_counts;
  int cap, ncls, method, max_seg, big;
  float iou_thr, sigma, min_score;
};

__global__ void __launch_bounds__(256)
softnms_lazy_kernel(const __grid_constant__ LazyNmsParams p) {
  pdl_wait();                 // PDL: the preceding grid has completed; nothing above touched global memory
  pdl_launch_dependents();    // let the next kernel's CTAs start their prologue
  extern __shared__ __align__(16) uint8_t lz_smem[];
  float* x1 = reinterpret_cast<float*>(lz_smem);
  float* x2 = x1 + p.cap;
  float* sc = x2 + p.cap;
  int* sl = reinterpret_cast<int*>(sc + p.cap);
  int* cls_off = sl + p.cap;                 // [ncls + 1]
  int* cursor = cls_off + p.ncls + 1;        // [ncls]
  float* head_s = reinterpret_cast<float*>(cursor + p.ncls);   // [ncls] score of the class's next emission
  int* head_i = reinterpret_cast<int*>(head_s + p.ncls);       // [ncls] its index in the smem arrays
  __shared__ unsigned long long s_tab[32];
  __shared__ float red_s[8];
  __shared__ int red_a[8], red_b[8];
  __shared__ int s_win;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const long long base = static_cast<long long>(b) * p.cap;
  if (tid < 32) s_tab[tid] = kExp2fTab[tid];
  for (int c = tid; c < p.ncls; c += blockDim.x) cursor[c] = 0;
  __syncthreads();
  // ---- counting sort by class (slot id kept for the "first in input order" tie-break)
  for (int s0 = tid; s0 < p.cap; s0 += blockDim.x) {
    const int lb = p.cand_labels[base + s0];
    if (lb >= 0 && lb < p.ncls && !(p.method == 3 && !(p.cand_scores[base + s0] > p.min_score))) atomicAdd(&cursor[lb], 1);
  }
  __syncthreads();
  if (tid == 0) {
    int acc = 0;
    for (int c = 0; c < p.ncls; ++c) { cls_off[c] = acc; acc += cursor[c]; cursor[c] = 0; }
    cls_off[p.ncls] = acc;
  }
  __syncthreads();
  for (int s0 = tid; s0 < p.cap; s0 += blockDim.x) {
    const int lb = p.cand_labels[base + s0];
    if (lb >= 0 && lb < p.ncls) {
      const float sv = p.cand_scores[base + s0];
      if (p.method == 3 && !(sv > p.min_score)) continue;
      const int pos = cls_off[lb] + atomicAdd(&cursor[lb], 1);
      x1[pos] = p.cand_segs[(base + s0) * 2];
      x2[pos] = p.cand_segs[(base + s0) * 2 + 1];
      sc[pos] = sv;
      sl[pos] = s0;
    }
  }
  __syncthreads();
  // ---- first emission of every class: highest score, ties -> lowest slot
  for (int c = warp; c < p.ncls; c += nwarps) {
    float bs = -CUDART_INF_F; int bslot = 0x7fffffff, bi = -1;
    for (int j = cls_off[c] + lane; j < cls_off[c + 1]; j += 32) {
      const float v = sc[j];
      if (v > bs || (v == bs && sl[j] < bslot)) { bs = v; bslot = sl[j]; bi = j; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float os = __shfl_xor_sync(0xffffffffu, bs, o);
      const int osl = __shfl_xor_sync(0xffffffffu, bslot, o), oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (os > bs || (os == bs && osl < bslot)) { bs = os; bslot = osl; bi = oi; }
    }
    if (lane == 0) { head_s[c] = bs; head_i[c] = bi; }
  }
  __syncthreads();
  float stride = 1.f, half = 0.f, fps = 1.f, dur = 0.f;
  if (p.vid_meta) {
    stride = p.vid_meta[b * 4 + 0]; half = 0.5f * p.vid_meta[b * 4 + 1];
    fps = p.vid_meta[b * 4 + 2]; dur = p.vid_meta[b * 4 + 3];
  }
  // ---- emission rounds.  The <= max_seg rounds are one dependent chain (each decays one class and finds its next emission),
  // so what matters is the latency of a round.  Round 1 ran every round on the whole block: four block barriers and two
  // block reductions, ~3.7 us per round, 370 us for 100 rounds.  Now WARP 0 drives the chain alone — class selection and, for
  // classes of at most LZ_BIG candidates (the normal case: ~10 100 slots over 100 classes), the decay too, with warp shuffles
  // and no block barrier; the other warps sleep on a named hardware barrier and are woken only for a round whose class is
  // bigger (the stress configuration: thousands of candidates in one class).  Same arithmetic, same tie-breaks (max score,
  // ties -> lowest input slot), hence the same bits as before.
  const int LZ_BIG = p.big;
  __shared__ int s_cw, s_iw, s_done, s_produced;
  // one thread's share of the decay of class cw against the emitted candidate iw: returns its best next emission
  auto decay = [&](int cw, int iw, int first, int step, float& ns, int& nslot, int& ni) {
    const float ix1 = x1[iw], ix2 = x2[iw];
    const float ia = __fadd_rn(__fsub_rn(ix2, ix1), 1e-6f);
    ns = -CUDART_INF_F; nslot = 0x7fffffff; ni = -1;
    for (int j = cls_off[cw] + first; j < cls_off[cw + 1]; j += step) {
      if (j == iw) continue;
      float v = sc[j];
      if (v == -CUDART_INF_F) continue;
      const float jx1 = x1[j], jx2 = x2[j];
      const float inter = fmaxf(0.f, __fsub_rn(fminf(ix2, jx2), fmaxf(ix1, jx1)));
      float w = 1.f;
      // a candidate that does not overlap the emitted segment keeps its score (overlap 0: below any positive threshold, and the
      // gaussian branch of the reference is taken for inter > 0 only): skip the two IEEE divisions and the FP64 exponential
      if (inter > 0.f || !(p.iou_thr > 0.f)) {
        const float ja = __fadd_rn(__fsub_rn(jx2, jx1), 1e-6f);
        const float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(ia, ja), inter));
        if (p.method == 0 || p.method == 3) { if (ovr >= p.iou_thr) w = 0.f; }
        else if (p.method == 1) { if (ovr >= p.iou_thr) w = __fsub_rn(1.f, ovr); }
        else if (inter > 0.f) w = expf_glibc(__fdiv_rn(-__fmul_rn(ovr, ovr), p.sigma), s_tab);
        v = __fmul_rn(v, w);
      }
      if (p.method == 3) { if (w == 0.f) v = -CUDART_INF_F; }
      else if (v < p.min_score) v = -CUDART_INF_F;
      sc[j] = v;
      if (v != -CUDART_INF_F && (v > ns || (v == ns && sl[j] < nslot))) { ns = v; nslot = sl[j]; ni = j; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float os = __shfl_xor_sync(0xffffffffu, ns, o);
      const int osl = __shfl_xor_sync(0xffffffffu, nslot, o), oi = __shfl_xor_sync(0xffffffffu, ni, o);
      if (oi >= 0 && (ni < 0 || os > ns || (os == ns && osl < nslot))) { ns = os; nslot = osl; ni = oi; }
    }
  };
  if (tid == 0) { s_done = 0; s_produced = 0; }
  __syncthreads();
  if (warp == 0) {
    int produced = 0;
    for (int r = 0; r < p.max_seg; ++r) {
      // ---- which class emits next: largest head score, ties -> lower class
      float bs = -CUDART_INF_F; int bc = 0x7fffffff;
      for (int c = lane; c < p.ncls; c += 32) {
        const float v = head_s[c];
        if (head_i[c] >= 0 && (bc == 0x7fffffff || v > bs || (v == bs && c < bc))) { bs = v; bc = c; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float os = __shfl_xor_sync(0xffffffffu, bs, o);
        const int oc = __shfl_xor_sync(0xffffffffu, bc, o);
        if (oc != 0x7fffffff && (bc == 0x7fffffff || os > bs || (os == bs && oc < bc))) { bs = os; bc = oc; }
      }
      const int cw = bc;
      if (cw == 0x7fffffff) break;
      const int iw = head_i[cw];
      if (lane == 0) {
        float s0 = x1[iw], s1 = x2[iw];
        if (p.vid_meta) {
          s0 = __fdiv_rn(__fadd_rn(__fmul_rn(s0, stride), half), fps);
          s1 = __fdiv_rn(__fadd_rn(__fmul_rn(s1, stride), half), fps);
          if (s0 <= 0.f) s0 = __fmul_rn(s0, 0.f);
          if (s1 <= 0.f) s1 = __fmul_rn(s1, 0.f);
          if (s0 >= dur) s0 = __fadd_rn(__fmul_rn(s0, 0.f), dur);
          if (s1 >= dur) s1 = __fadd_rn(__fmul_rn(s1, 0.f), dur);
        }
        const long long orow = static_cast<long long>(b) * p.max_seg + r;
        p.out_segs[orow * 2] = s0; p.out_segs[orow * 2 + 1] = s1;
        p.out_scores[orow] = sc[iw]; p.out_labels[orow] = cw;
      }
      produced = r + 1;
      // ---- one decay round of that class + its next emission
      float ns; int nslot, ni;
      if (cls_off[cw + 1] - cls_off[cw] <= LZ_BIG) {
        decay(cw, iw, lane, 32, ns, nslot, ni);
      } else {
        if (lane == 0) { s_cw = cw; s_iw = iw; }
        __syncwarp();
        asm volatile("bar.sync 1, 256;" ::: "memory");          // wake the helper warps
        decay(cw, iw, tid, 256, ns, nslot, ni);
        if (lane == 0) { red_s[0] = ns; red_a[0] = nslot; red_b[0] = ni; }
        asm volatile("bar.sync 2, 256;" ::: "memory");          // their partial results are in red_*
        for (int w = 1; w < nwarps; ++w)
          if (red_b[w] >= 0 && (ni < 0 || red_s[w] > ns || (red_s[w] == ns && red_a[w] < nslot))) { ns = red_s[w]; nslot = red_a[w]; ni = red_b[w]; }
      }
      if (lane == 0) {
        sc[iw] = -CUDART_INF_F;
        head_s[cw] = ns; head_i[cw] = ni;
      }
      __syncwarp();
    }
    if (lane == 0) { s_done = 1; s_produced = produced; }
    __syncwarp();
    asm volatile("bar.sync 1, 256;" ::: "memory");              // release the helpers for good
  } else {
    for (;;) {
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (s_done) break;
      float ns; int nslot, ni;
      decay(s_cw, s_iw, tid, 256, ns, nslot, ni);
      if (lane == 0) { red_s[warp] = ns; red_a[warp] = nslot; red_b[warp] = ni; }
      asm volatile("bar.sync 2, 256;" ::: "memory");
    }
  }
  __syncthreads();
  const int produced = s_produced;
  for (int r = produced + tid; r < p.max_seg; r += blockDim.x) {
    const long long orow = static_cast<long long>(b) * p.max_seg + r;
    p.out_segs[orow * 2] = 0.f; p.out_segs[orow * 2 + 1] = 0.f; p.out_scores[orow] = 0.f; p.out_labels[orow] = 0;
  }
  if (tid == 0) p.out_counts[b] = produced;
}

}  // namespace unav

using namespace unav;

extern "C" int unav_decode(const float* logits, const float* offsets, const uint8_t* masks, const float* points,
                           const int* level_off, int B, int L, int ncls, int class_aware, float pre_nms_thresh,
                           int pre_nms_topk, float duration_thresh, float* cand_segs, float* cand_scores,
                           int32_t* cand_labels, int cap, void* stream) {
  UNAV_REQUIRE(logits && offsets && masks && points && level_off && cand_segs && cand_scores && cand_labels,
               "decode: null pointer");
  UNAV_REQUIRE(L >= 1 && L <= 8 && B >= 1 && ncls >= 1 && pre_nms_topk >= 1 && pre_nms_thresh >= 0.f,
               "decode: bad arguments");
  DecodeParams p;
  p.logits = logits; p.offsets = offsets; p.masks = masks; p.points = points;
  p.cand_segs = cand_segs; p.cand_scores = cand_scores; p.cand_labels = cand_labels;
  int slots = 0;
  for (int l = 0; l <= L; ++l) p.level_off[l] = level_off[l];
  for (int l = 0; l < L; ++l) {
    p.cap_off[l] = slots;
    const long long n = static_cast<long long>(level_off[l + 1] - level_off[l]) * ncls;
    slots += static_cast<int>(n < pre_nms_topk ? n : pre_nms_topk);
  }
  p.cap_off[L] = slots;
  UNAV_REQUIRE(cap >= slots, "decode: cap %d < required %d candidate slots", cap, slots);
  p.B = B; p.L = L; p.ncls = ncls; p.class_aware = class_aware; p.topk = pre_nms_topk; p.cap = cap;
  p.Ttot = level_off[L]; p.thresh = pre_nms_thresh; p.dur_thresh = duration_thresh;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (cap > slots) {   // slots past the last level region are never written by the kernel
    for (int b = 0; b < B; ++b) {
      cudaError_t e = cudaMemsetAsync(cand_labels + static_cast<long long>(b) * cap + slots, 0xff,
                                      sizeof(int32_t) * (cap - slots), s);
      if (e != cudaSuccess) { set_error("decode: memset: %s", cudaGetErrorString(e)); return (int)e; }
    }
  }
  int max_n = 0;
  for (int l = 0; l < L; ++l) {
    const int n = (level_off[l + 1] - level_off[l]) * ncls;
    max_n = n > max_n ? n : max_n;
  }
  size_t smem = static_cast<size_t>(max_n) * sizeof(float);
  p.use_smem = smem <= 200 * 1024;        // long sequences (T = 2304: 921 KB at level 0) recompute the sigmoid per pass
  if (!p.use_smem) smem = 0;
  static unav::SmemAttr attr = {};
  if (int rc = unav::ensure_dyn_smem(decode_kernel, attr, smem, "decode")) return rc;
  dim3 grid(L, B);
  launch_pdl(decode_kernel, dim3(grid), dim3(DEC_THREADS), smem, s, p);
  count_launch();
  return finish_launch("decode");
}

extern "C" size_t unav_softnms_workspace_bytes(int B, int ncls, int max_seg_num) {
  const size_t dets = static_cast<size_t>(B) * ncls * max_seg_num * 3 * sizeof(float);
  const size_t cnts = static_cast<size_t>(B) * ncls * sizeof(int32_t);
  return ((dets + 255) / 256) * 256 + ((cnts + 255) / 256) * 256;
}

extern "C" int unav_softnms_batched(const float* cand_segs, const float* cand_scores, const int32_t* cand_labels,
                                    int B, int cap, int ncls, float iou_threshold, float sigma, float min_score,
                                    int method, int max_seg_num, int max_per_class, const float* vid_meta, float* out_segs,
                                    float* out_scores, int64_t* out_labels, int32_t* out_counts, void* workspace,
                                    size_t workspace_bytes, void* stream) {
  UNAV_REQUIRE(cand_segs && cand_scores && cand_labels && out_segs && out_scores && out_labels && out_counts,
               "softnms: null pointer");
  UNAV_REQUIRE(B >= 1 && cap >= 1 && ncls >= 1 && ncls <= 1024 && max_seg_num >= 1, "softnms: bad arguments");
  UNAV_REQUIRE(method >= 0 && method <= 3, "softnms: method %d not in 0..3", method);
  UNAV_REQUIRE(workspace && workspace_bytes >= unav_softnms_workspace_bytes(B, ncls, max_seg_num),
               "softnms: workspace too small");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  {
    // lazy per-video kernel when a whole video's candidates fit in shared memory (always true on the model path)
    const size_t lz = static_cast<size_t>(cap) * 16 + (static_cast<size_t>(ncls) * 4 + 1) * 4 + 64;
    if (lz <= 200 * 1024 && !getenv("UNAV_NMS_PER_CLASS")) {
      static unav::SmemAttr lz_attr = {};
      if (int rc = unav::ensure_dyn_smem(softnms_lazy_kernel, lz_attr, lz, "softnms_lazy")) return rc;
      LazyNmsParams q;
      q.cand_segs = cand_segs; q.cand_scores = cand_scores; q.cand_labels = cand_labels; q.vid_meta = vid_meta;
      q.out_segs = out_segs; q.out_scores = out_scores; q.out_labels = out_labels; q.out_counts = out_counts;
      q.cap = cap; q.ncls = ncls; q.method = method; q.max_seg = max_seg_num;
      q.iou_thr = iou_threshold; q.sigma = sigma; q.min_score = min_score;
      q.big = 64;                                   // classes above this size are decayed by the whole block (see the kernel)
      if (const char* env = getenv("UNAV_NMS_WARP_MAX")) { const int v = atoi(env); if (v >= 0) q.big = v; }
      launch_pdl(softnms_lazy_kernel, dim3(B), dim3(256), lz, s, q);
      count_launch();
      return finish_launch("softnms_lazy");
    }
  }
  NmsParams p;
  p.cand_segs = cand_segs; p.cand_scores = cand_scores; p.cand_labels = cand_labels;
  p.ws_dets = reinterpret_cast<float*>(workspace);
  const size_t dets = static_cast<size_t>(B) * ncls * max_seg_num * 3 * sizeof(float);
  p.ws_counts = reinterpret_cast<int32_t*>(reinterpret_cast<char*>(workspace) + ((dets + 255) / 256) * 256);
  p.B = B; p.cap = cap; p.ncls = ncls; p.method = method; p.max_seg = max_seg_num;
  p.iou_thr = iou_threshold; p.sigma = sigma; p.min_score = min_score;
  // a class holds at most max_per_class candidates (caller's structural bound; 0 = cap); small classes run
  // as a single warp, big ones as a full block
  p.maxn = (max_per_class > 0 && max_per_class < cap) ? max_per_class : cap;
  const size_t smem = static_cast<size_t>(p.maxn) * 4 * sizeof(float);
  UNAV_REQUIRE(smem <= 200 * 1024, "softnms: %d candidates per class exceed the shared-memory budget", p.maxn);
  static unav::SmemAttr attr = {};
  if (int rc = unav::ensure_dyn_smem(softnms_kernel, attr, smem, "softnms")) return rc;
  const int threads = p.maxn <= 1024 ? 32 : 256;
  dim3 grid(ncls, B);
  launch_pdl(softnms_kernel, dim3(grid), dim3(threads), smem, s, p);
  count_launch();
  int rc = finish_launch("softnms");
  if (rc) return rc;
  MergeParams m;
  m.ws_dets = p.ws_dets; m.ws_counts = p.ws_counts; m.vid_meta = vid_meta;
  m.out_segs = out_segs; m.out_scores = out_scores; m.out_labels = out_labels; m.out_counts = out_counts;
  m.ncls = ncls; m.max_seg = max_seg_num;
  launch_pdl(merge_kernel, dim3(B), dim3(256), 0, s, m);
  count_launch();
  return finish_launch("softnms_merge");
}
