/*
 * unav_b200.h — C ABI of the B200-native UnAV inference hot path.
 *
 * This header is the drop-in boundary: a plain-C interface (device pointers, sizes, a
 * cudaStream_t passed as void*) with no torch types.  Python host code binds it with
 * ctypes (unav_yolyolva_b200/_cabi.py); see INTEGRATION.md for the reference-side stubs.
 *
 * What each entry point replaces in the reference (paths relative to /root/reference):
 *   unav_softnms_batched   libs/utils/csrc/nms_cpu.cpp:67-160 (softnms_1d_cpu), :172-182
 *                          (pybind11 module nms_1d_cpu) and the per-class Python loop +
 *                          final sort of libs/utils/nms.py:103-190 (batched_nms), plus the
 *                          seconds conversion of libs/modeling/multimodal_meta_archs.py:852-856
 *   unav_decode            libs/modeling/multimodal_meta_archs.py:745-817
 *                          (PtTransformer.inference_single_video)
 *   unav_gemm              every dense nn.Linear / nn.Conv1d(k=1) / im2col'd k=3 conv on the
 *                          path (libs/modeling/blocks.py:36-61, :187-196, :296-302;
 *                          multimodal_backbones.py:867-870, :929-932, :989-996, :1014-1034;
 *                          multimodal_meta_archs.py:132-151, :214-242)
 *   unav_layernorm_rows    libs/modeling/blocks.py:91-103 (channel LayerNorm) and the
 *                          nn.LayerNorm(512) uses of multimodal_backbones.py:946-952,1011-1034
 *   unav_dwconv_ln         depthwise MaskedConv1D + LayerNorm pairs: blocks.py:205-211
 *                          (q/k/v convs of MaskedMHCA) and multimodal_backbones.py:44-48
 *                          (Downsample_pyramid_levels)
 *   unav_attention         blocks.py:218-240 (MaskedMHCA core) and
 *                          multimodal_backbones.py:899-918 (MultiHeadAttention core with the
 *                          fused mask of :1173-1183)
 *   unav_attention_tc      same as unav_attention, tcgen05/TMEM version for key lengths <= 256
 *   unav_maxsig_gate       multimodal_backbones.py:170-191 (MaxSigmoidAttnBlock weight)
 *   unav_pool_match        multimodal_backbones.py:591-598 (avg-pool(4) x3 + match_projection)
 *   unav_rowcopy           torch.cat / nn.Upsample(nearest) / im2col glue:
 *                          multimodal_backbones.py:565-576, :609-610; meta_archs.py:469
 *   unav_transpose_cast    the [B,C,T] <-> [B,T,C] transposes (multimodal_backbones.py:1145-1146,
 *                          :1200-1201, :170)
 *   unav_align_embed       multimodal_backbones.py:1157-1166 (CLS + pos + type embedding)
 *   unav_pack_operand      load-time conversion of nn.Linear / nn.Conv1d FP32 weights into GEMM operand rows
 *   unav_map_match         libs/utils/metrics.py:340-398 (greedy tIoU matching of compute_average_precision_detection)
 *   unav_collate_pad       libs/datasets/data_utils.py:178-205 (padding + mask of collate_fcn, on the device)
 *   unav_build_masks       blocks.py:45-51 (mask[::s]) and multimodal_backbones.py:568-570
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the comment says "host";
 *   - activations are token-major: a [nb, T, C] tensor is a row-major matrix of nb*T rows
 *     ("rows" = time steps of consecutive videos) with leading dimension ld >= C in elements;
 *   - "operand dtype" (op_dtype) is UNAV_F32 or UNAV_BF16 and names the element type of GEMM
 *     operands (A, W, and out_op buffers); accumulation, residual stream, LayerNorm, softmax
 *     and all detections are FP32;
 *   - calls are asynchronous on `stream`, re-entrant per stream, own no memory (the caller
 *     provides every buffer), and return 0 or a non-zero code: >0 = cudaError_t,
 *     <0 = UNAV_ERR_*.  unav_last_error() returns a host string describing the last failure
 *     of the calling thread.
 */
#ifndef UNAV_B200_H_
#define UNAV_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* UNAV_BF16X2 / UNAV_F16X2: split pair, hi = round16(x) at column c, lo = round16(x - hi) at column ld/2 + c; the
 * tensor-core kernels run 3 MMA passes over it (hi.hi + lo.hi + hi.lo).  F16 halves carry 11 significand bits each
 * (hi + lo ~ 22 bits: FP32-grade products, measured 1e-6 of range on the logits), BF16 halves 8 each but FP32 range.
 * Wherever an `op_dtype` argument is taken, bits 8..9 may carry a pass count for split operands:
 * UNAV_PASSES(1) = hi.hi only (the lo halves are not even loaded), UNAV_PASSES(2) = hi.hi + lo.hi, 0 / 3 = all. */
enum { UNAV_F32 = 0, UNAV_BF16 = 1, UNAV_BF16X2 = 2, UNAV_F16 = 3, UNAV_F16X2 = 4 };
#define UNAV_PASSES(n) ((n) << 8)
enum { UNAV_ACT_NONE = 0, UNAV_ACT_RELU = 1, UNAV_ACT_GELU = 2, UNAV_ACT_SILU = 3 };
enum { UNAV_GEMM_SIMT = 0, UNAV_GEMM_TCGEN05 = 1 };
enum {
  UNAV_ERR_BAD_ARG = -1,
  UNAV_ERR_UNSUPPORTED = -2,
  UNAV_ERR_NO_DEVICE = -3,
  UNAV_ERR_DRIVER = -4
};
#define UNAV_MAX_GROUPS 8
#define UNAV_MAX_COPY_JOBS 16

/* ---- library / device ---------------------------------------------------------------- */
const char* unav_version(void);
const char* unav_last_error(void);
/* 0 if device `dev` is an sm_100 part usable by this library. */
int unav_check_device(int dev);
/* number of kernel launches issued by this library in the calling process so far */
long long unav_launch_count(void);

/* Which kernel the calling thread's last UNAV_GEMM_TCGEN05 call launched (the tile is chosen inside the library):
 * 0 gemm_tcgen05_kernel<64,64>, 1 <128,32>, 2 <128,64>, 3 gemm_tcgen05_pair_kernel, 4 <64,32>; -1 if none yet.
 * bench.py uses it to attribute its per-launch CUDA-event times to the kernels ncu lists. */
int unav_gemm_last_variant(void);
/* Cumulative number of UNAV_GEMM_TCGEN05 launches per variant id (same numbering; 5 = 128-wide CTA pair, 6 = <256,32>,
 * 7 = persistent CTA-pair kernel) since the library was loaded: out[0..n).  Parity tests use the difference across a run to
 * prove which tile variants a configuration exercised. */
#define UNAV_GEMM_VARIANTS 8
int unav_gemm_variant_counts(long long* out, int n);

/* Diagnostics for the tcgen05 kernels (scripts/gemm_phases.py): while a buffer is set, every CTA with linear index
 * below capacity_ctas writes 8 int64: GEMM {smid, clock64 at: start, setup done, first operands landed, last MMA
 * issued, accumulator ready, first epilogue warp done, all warps done}; attention {smid, start, setup done (TMEM
 * allocated), S ready, P written, O ready, epilogue done, all warps done}.  NULL switches it off (the default). */
int unav_set_phase_trace(long long* device_buf, int capacity_ctas);

/* Programmatic dependent launch for the launches that follow (every kernel of the library waits with griddepcontrol.wait
 * before it touches global memory): 1 = on, 0 = off, -1 = follow the UNAV_PDL environment variable (default off).  The launch
 * attribute is captured into CUDA graphs, so the engine turns it on around the capture of its batch <= 2 latency plans only:
 * it shortens the dependent chain of a single small batch (2.83 -> 2.64 ms at batch 1) but costs throughput with several
 * batches in flight (early-launched CTAs hold SM resources while they wait). */
int unav_set_pdl(int mode);

/* ---- GEMM: C[M,N] = epilogue(A[M,K] . W[N,K]^T) ---------------------------------------- */
/* Epilogue, per element (m, n), in this order:
 *   v = acc + bias[n]; v *= rowmask[m]; v *= rowscale[m]; v *= gate[m*gate_groups + n/gate_width];
 *   v = act(v); if res: v = res[m,n] * (res_masked ? rowmask[m] : 1) + colscale[n] * v
 * (each factor skipped when its pointer is NULL; colscale NULL = 1).  v is stored to
 * out_f32[m*ld_f32 + n] and/or, converted, to out_op[m*ld_op + n]. */
typedef struct UnavGemmGroup {
  const void* A;   long long lda;     /* [M,K] operand dtype */
  const void* W;   long long ldw;     /* [N,K] operand dtype */
  const float* bias;                  /* [N] */
  const uint8_t* rowmask;             /* [M] */
  const float* rowscale;              /* [M] */
  const float* gate;                  /* [M, gate_groups] */
  const float* res; long long ldres;  /* [M, >=N] */
  const float* colscale;              /* [N] */
  float* out_f32;  long long ld_f32;
  void* out_op;    long long ld_op;
  int gate_groups; int gate_width;
  /* optional TRANSPOSED operand-dtype output of the column window [t_col0, t_col0 + t_ncols) (t_ncols = 0: all N):
   *   out_opT[((m / t_seg) * ncols + (n - t_col0)) * ld_opT + (m % t_seg)] = acc + bias[n]
   * i.e. per item of t_seg rows a [ncols, t_seg] matrix — the V^T layout unav_attention_tc consumes. */
  void* out_opT;   long long ld_opT;
  int t_seg; int t_col0; int t_ncols;
  /* conv_T > 0 (UNAV_GEMM_TCGEN05 only): implicit k=3 convolution with per-segment zero padding (MaskedConv1D, stride 1,
   * blocks.py:30-31).  A is then the PLAIN operand [M, K/3] of the convolution's input (M = segments * conv_T rows) and
   * W the [N, K] weight with K = 3 * Cin laid out tap-major (column tap*Cin + c); row m = (s, t) of the product uses
   * A rows (s, t-1), (s, t), (s, t+1), zero outside the segment.  Same result, bit for bit, as a GEMM over the
   * materialised im2col operand. */
  int conv_T;
} UnavGemmGroup;

/* groups: host array of ngroups (<= UNAV_MAX_GROUPS) problems of identical M, N, K.
 * backend UNAV_GEMM_TCGEN05 requires a 16-bit op_dtype (BF16 / F16, plain or split), 16-byte aligned A/W and
 * lda, ldw multiples of 8 elements. */
int unav_gemm(const UnavGemmGroup* groups, int ngroups, int M, int N, int K,
              int op_dtype, int act, int res_masked, int backend, void* stream);


/* ---- row LayerNorm (+pre-add, +activation, +position table, +im2col scatter) ----------- */
/* For output row r (0 <= r < M):
 *   src = x_seg_rows ? (r / x_seg_rows) * x_seg_stride + r % x_seg_rows + x_row_off : r
 *   u = x[src, :] (+ add[r, :]);  y = (u - mean) / sqrt(var + eps) * w + b;  y = act(y)
 *   if post: y += post[r % post_rows, :] * (rowmask ? rowmask[r] : 1)
 * y goes to out_f32 (ld_f32), out_op (ld_op) and/or, as a k=3 im2col scatter, to out_im2col:
 *   out_im2col[r, C + c] = y[c];  out_im2col[r+1, c] = y[c] unless edge[r]&2;
 *   out_im2col[r-1, 2C + c] = y[c] unless edge[r]&1; rows with edge bit0 (first of a video
 *   segment) zero their own tap-0 block, rows with bit1 (last) zero their tap-2 block. */
typedef struct UnavLnGroup {
  const float* x;   long long ldx;
  const float* add; long long ldadd;
  const float* w;   const float* b;          /* [C] */
  const float* post;                         /* [post_rows, C] or NULL */
  const uint8_t* rowmask;                    /* [M] or NULL */
  const uint8_t* edge;                       /* [M], needed with out_im2col */
  float* out_f32;     long long ld_f32;
  void* out_op;       long long ld_op;
  void* out_im2col;   long long ld_im2col;
  int x_seg_rows, x_seg_stride, x_row_off, post_rows;
} UnavLnGroup;

int unav_layernorm_rows(const UnavLnGroup* groups, int ngroups, int M, int C, float eps,
                        int act, int op_dtype, void* stream);

/* ---- depthwise k=3 conv (stride 1|2, zero pad 1) * mask -> LayerNorm, up to 3 outputs --- */
/* Input x: nseg segments of seg_len_in rows.  Optional pre-LayerNorms (n_pre <= 2) are applied
 * to the input rows first (TransformerBlock ln11/ln12, blocks.py:314).  Output j (< n_out):
 *   z[t, c] = sum_tap dw_j[c, tap] * pre_{src_j}(x)[stride*t + tap - 1, c]   (0 outside the segment)
 *   z *= mask_out[seg*seg_len_out + t];  y = LN_j(z) (y = z when ln_w is NULL) -> out_f32_j / out_op_j */
typedef struct UnavDwLnOut {
  const float* dw;                           /* [C,3] */
  const float* ln_w; const float* ln_b;      /* [C] */
  float* out_f32; long long ld_f32;
  void* out_op;   long long ld_op;
  int src;                                   /* -1 raw x, 0|1 pre-LN index */
  int pad_;
} UnavDwLnOut;

typedef struct UnavDwLnGroup {
  const float* x; long long ldx;
  const uint8_t* mask_out;                   /* [nseg*seg_len_out] */
  const float* pre_w[2]; const float* pre_b[2];
  UnavDwLnOut out[3];
} UnavDwLnGroup;

int unav_dwconv_ln(const UnavDwLnGroup* groups, int ngroups, int nseg, int seg_len_in,
                   int stride, int C, int n_pre, int n_out, float eps, int op_dtype,
                   void* stream);

/* ---- fused masked attention core --------------------------------------------------------- */
/* For batch b < nb, head h < nh, query i < Tq:
 *   s_j = scale * <q[b,i,h,:], k[b,j,h,:]>  over keys j < Tk with kmask[b*Tk+j] != 0,
 *   plus, when xk != NULL and i >= x_first, one extra key/value row xk/xv[b*Tq + i] (always
 *   allowed: the time-aligned token of the other modality, multimodal_backbones.py:1179-1183);
 *   p = softmax(s);  out[b,i,h,:] = sum_j p_j * v[b,j,h,:].
 * q/k/v rows are FP32 with head h at columns [h*hs, (h+1)*hs); out has operand dtype. */
typedef struct UnavAttnGroup {
  const float* q; long long ldq;
  const float* k; long long ldk;
  const float* v; long long ldv;
  const uint8_t* kmask;
  const float* xk; const float* xv; long long ldx;
  void* out; long long ldo;
  int x_first; int pad_;
} UnavAttnGroup;

int unav_attention(const UnavAttnGroup* groups, int ngroups, int nb, int Tq, int Tk,
                   int nh, int hs, float scale, int op_dtype, void* stream);

/* ---- fused attention on the tensor cores (tcgen05 / TMEM), key length <= 256 ------------------------ */
/* Same math as unav_attention.  q, k: operand-dtype (16-bit, plain or split) token-major rows [nb*T, ld] with
 * head h at columns [h*hs, (h+1)*hs); vt: the values TRANSPOSED per batch item, operand dtype [nb*nh*hs, ldvt >= Tk]
 * (row = item*C + channel, column = key), e.g. unav_transpose_cast(v, nb, R = Tk, Cc = C).  The optional extra key uses FP32 rows q32 / xk / xv
 * (token-major, ld = ldq32 / ldx).  out has operand dtype.  S = Q.K^T and O = P.V run as tcgen05.mma with the
 * accumulators and P in tensor memory; BF16X2 operands use the 3-pass split for FP32-level accuracy.
 * Masked work is not executed: keys beyond an item's last valid key (kmask) are neither loaded nor multiplied nor
 * exponentiated — they contribute exactly 0 in the reference's softmax (blocks.py:233-236) — so the cost of an item follows
 * its valid length, not the padded one. */
typedef struct UnavAttnTcGroup {
  const void* q;  long long ldq;
  const void* k;  long long ldk;
  const void* vt; long long ldvt;
  const uint8_t* kmask;
  const float* q32; long long ldq32;
  const float* xk; const float* xv; long long ldx;
  void* out; long long ldo;
  int x_first; int pad_;
  /* optional [nb, Tq] query validity (NULL = all valid): a 128-query tile whose queries are all invalid is skipped and its
   * output rows are written as zeros (the reference computes them and multiplies them by the mask afterwards,
   * blocks.py:243).  Ignored for Tk > 256. */
  const uint8_t* qmask;
} UnavAttnTcGroup;

int unav_attention_tc(const UnavAttnTcGroup* groups, int ngroups, int nb, int Tq, int Tk,
                      int nh, int hs, float scale, int op_dtype, void* stream);

/* Same contract for ANY key length (BASELINE.json config 4: max_seq_len = 2304; blocks.py:218-240 with T = 2304,
 * multimodal_backbones.py:845-924 with 2305 tokens): for Tk > 256 the keys are processed in 256-key chunks by separate CTAs
 * of the same tcgen05 kernel (S and P in tensor memory, chunks beyond a video's valid length skipped), each writing an
 * un-normalised FP32 partial output and the row's (max, sum) into `workspace`; a second kernel merges the chunks and the
 * optional extra key.  workspace: at least unav_attention_tc_workspace_bytes(...) bytes (0 for Tk <= 256, where the call is
 * identical to unav_attention_tc). */
size_t unav_attention_tc_workspace_bytes(int ngroups, int nb, int Tq, int Tk, int nh, int hs);
int unav_attention_tc_long(const UnavAttnTcGroup* groups, int ngroups, int nb, int Tq, int Tk,
                           int nh, int hs, float scale, int op_dtype, void* workspace,
                           size_t workspace_bytes, void* stream);

/* ---- MaxSigmoid gate ----------------------------------------------------------------------- */
/* gate[r, h] = sigmoid( max_{n<nwords} <x[r, h*hc:(h+1)*hc], G[(r/T)*nwords + n, h*hc:...]> / sqrt(hc)
 *                       + head_bias[h] ),   r < nb*T, h < H.  x, G FP32. */
int unav_maxsig_gate(const float* x, long long ldx, const float* G, long long ldg,
                     const float* head_bias, float* gate, int nb, int T, int nwords, int H,
                     int hc, void* stream);

/* Tensor-core version: x and G as operand-dtype rows (column windows x_col0 / g_col0 of wider buffers with leading
 * dimensions ldx / ldg); S = X_h . G_h^T runs as tcgen05.mma into all 512 TMEM columns, the row max is taken from
 * tensor memory.  Needs nwords == 512 and hc in {32, 64}. */
int unav_maxsig_gate_tc(const void* x, long long ldx, int x_col0, const void* G, long long ldg, int g_col0,
                        const float* head_bias, float* gate, int nb, int T, int nwords, int H, int hc,
                        int op_dtype, void* stream);

/* ---- AdaptiveAvgPool1d(P) of 3 levels + Conv1d(3P -> Tq, k=1) along the pooled axis ------- */
/* q[b*Tq + t, c] = bm[t] + sum_{l<3, p<P} Wm[t, l*P + p] * mean_{s in bin p of level l} u_l[b*T_l + s, c] */
int unav_pool_match(const float* u0, const float* u1, const float* u2, int T0, int T1, int T2,
                    long long ldu, const float* Wm, const float* bm, float* q, long long ldq,
                    int nb, int C, int Tq, int P, void* stream);

/* ---- row gather / nearest up-sample / k=3 im2col / concat into operand buffers ------------ */
/* dst[(seg*dst_seg_stride + dst_row_off + t), tap*tap_stride + c] =
 *     src[seg*seg_len_in + (t*num)/den + tap - ntaps/2, c]   (0 if outside [0, seg_len_in))
 * for seg < nseg, t < seg_len_out, tap < ntaps, c < C.  dst has operand dtype (or FP32 when
 * dst_f32 != 0). */
typedef struct UnavCopyJob {
  const float* src; long long ld_src;
  void* dst;        long long ld_dst;
  int nseg, seg_len_in, seg_len_out, dst_seg_stride, dst_row_off;
  int num, den, ntaps, tap_stride, C;
} UnavCopyJob;

int unav_rowcopy(const UnavCopyJob* jobs, int njobs, int op_dtype, void* stream);

/* in [nb, R, Cc] FP32 (row stride ld_in) -> out [nb, Cc, R] operand dtype (row stride ld_out) */
int unav_transpose_cast(const float* in, long long ld_in, void* out, long long ld_out,
                        int nb, int R, int Cc, int op_dtype, void* stream);

/* Weight packing at model-load time (replaces the reference's implicit use of nn.Conv1d / nn.Linear FP32 weights,
 * libs/modeling/blocks.py:30-31, multimodal_backbones.py:989-1034): src [rows, K] FP32 (row stride ld_src) -> dst operand rows
 * of ld_dst elements in op_dtype (split formats: hi = half(x) at column c, lo = half(x - hi) at column ld_dst/2 + c); columns
 * K..width-1 are written as zeros, so dst needs no prior memset. */
int unav_pack_operand(const float* src, long long ld_src, void* dst, long long ld_dst, long long rows, int K,
                      int op_dtype, void* stream);

/* tokens[m][b][0,:] = cls_m + pos_m[0] + type_m; tokens[m][b][1+t,:] = x0[m][b][t,:] + pos_m[1+t] + type_m
 * for m in {0: video, 1: audio}; x0 is [2, nb, T, C], tokens [2, nb, T+1, C]. */
int unav_align_embed(const float* x0, const float* cls_v, const float* cls_a,
                     const float* pos_v, const float* pos_a, const float* type_v,
                     const float* type_a, float* tokens, int nb, int T, int C, void* stream);

/* Device-side collate (what libs/datasets/data_utils.py:178-205 does on the host with per-video copy_ loops):
 * ragged = the B videos' [C, lens[b]] row-major feature blocks back to back, offsets[b] = index of the first float of
 * video b.  out[b, c, t] = t < min(lens[b], T) ? ragged[offsets[b] + c*lens[b] + t] : pad  ([B, C, T], the model's
 * input layout); mask (optional) [B, T] = t < lens[b]  (data_utils.py:201: arange(max_len) < feats_lens). */
int unav_collate_pad(const float* ragged, const long long* offsets, const int* lens, float* out,
                     uint8_t* mask, int B, int C, int T, float pad, void* stream);

/* Pyramid masks from the level-0 mask [nb_src, T] replicated to nb = k*nb_src batch items (item b uses
 * mask[b % nb_src]):  out_true[l][b][t] = mask[b][t << l]  (blocks.py:45-51),
 * out_up[l][b][t] = out_true[l+1][b][t >> 1]  (multimodal_backbones.py:568-570), both concatenated
 * over levels (level l starts at row offset nb * sum_{j<l} (T >> j)); out_up has L-1 levels.
 * out_cls (optional) [nb_src, T+1] = [1, mask] (the CLS-extended mask, multimodal_backbones.py:1159).
 * out_heads (optional) [nb_src, sum_l (T >> l)]: the true masks of all levels, video-major (the row
 * order of the heads / decode, multimodal_meta_archs.py:478-493). */
int unav_build_masks(const uint8_t* mask, uint8_t* out_true, uint8_t* out_up, uint8_t* out_cls,
                     uint8_t* out_heads, int nb, int nb_src, int T, int L, void* stream);

/* ---- decode ---------------------------------------------------------------------------------- */
/* logits [B, Ttot, ncls], offsets [B, Ttot, ncls, 2] (class_aware) or [B, Ttot, 2], masks [B, Ttot],
 * points [Ttot, 4] = (t, reg_lo, reg_hi, stride); level l covers rows [level_off[l], level_off[l+1])
 * (host array, L+1 entries).  Per (video, level): p = sigmoid(logit)*mask; keep p > pre_nms_thresh;
 * keep the pre_nms_topk largest (ties: lower flat index); segment = (t - off0*stride, t + off1*stride);
 * keep right-left > duration_thresh.  Candidates of video b land in slots
 * [b*cap + cap_off[l], ...) in flat-index order; unused slots get label -1.
 * cap must be >= sum_l min(pre_nms_topk, rows_l*ncls). */
int unav_decode(const float* logits, const float* offsets, const uint8_t* masks,
                const float* points, const int* level_off, int B, int L, int ncls,
                int class_aware, float pre_nms_thresh, int pre_nms_topk, float duration_thresh,
                float* cand_segs, float* cand_scores, int32_t* cand_labels, int cap,
                void* stream);

/* ---- per-class temporal soft-NMS + per-video merge ------------------------------------------ */
/* Candidates: cand_segs [B, cap, 2], cand_scores [B, cap], cand_labels [B, cap] (int32, -1 = empty).
 * method: 0 hard (weight 0 if IoU >= iou_threshold), 1 linear, 2 gaussian exp(-IoU^2/sigma)
 * (nms_cpu.cpp:126-141); 3 = the NMSop path (nms.py:8-35 -> nms_cpu.cpp:19-58): pre-filter
 * score > min_score, greedy suppression at IoU >= iou_threshold, scores unchanged.
 * max_per_class: upper bound on candidates of one class in one video (0 = cap); the model path
 * passes sum_l T_l because every (point, class) yields at most one candidate.  Per class: repeatedly take the highest-scoring live candidate (ties:
 * lowest slot), emit it, decay the others, drop those below min_score; at most max_seg_num per
 * class.  Then the per-class lists are merged by score (ties: lower class, earlier emission) and
 * the best max_seg_num kept.  If vid_meta != NULL ([B,4] = feat_stride, feat_num_frames, fps,
 * duration) segments are converted to seconds and clamped (multimodal_meta_archs.py:852-856).
 * Outputs: out_segs [B, max_seg_num, 2], out_scores [B, max_seg_num], out_labels [B, max_seg_num]
 * (int64), out_counts [B] (int32); rows beyond the count are zero.
 * workspace: at least unav_softnms_workspace_bytes(B, ncls, max_seg_num) bytes. */
size_t unav_softnms_workspace_bytes(int B, int ncls, int max_seg_num);
int unav_softnms_batched(const float* cand_segs, const float* cand_scores,
                         const int32_t* cand_labels, int B, int cap, int ncls,
                         float iou_threshold, float sigma, float min_score, int method,
                         int max_seg_num, int max_per_class, const float* vid_meta,
                         float* out_segs,
                         float* out_scores, int64_t* out_labels, int32_t* out_counts,
                         void* workspace, size_t workspace_bytes, void* stream);

/* ---- detection mAP: greedy matching of ranked detections to ground truth ------------------------------------ */
/* libs/utils/metrics.py:340-398 for every class, video and tIoU threshold at once.  Detections and ground truth are grouped
 * by (class, video): group g owns detections det_ptr[g] .. det_ptr[g+1] (det_seg rows [start, end] in FP64, in the class's
 * descending-score order) and ground-truth rows gt_ptr[2g] .. gt_ptr[2g+1] of gt_seg (original order; may be empty).
 * tp[t*ndet + i] = 1 if detection i is a true positive at tious[t], else 0 (it is then a false positive).
 * lock: workspace of nt*ngt bytes. */
int unav_map_match(const double* det_seg, const double* gt_seg, const int* det_ptr, const int* gt_ptr, int ngroups,
                   const double* tious, int nt, int ndet, int ngt, uint8_t* tp, uint8_t* lock, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* UNAV_B200_H_ */
