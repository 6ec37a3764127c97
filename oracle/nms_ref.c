/*
 * nms_ref.c — CPU restatement of the reference's temporal (soft-)NMS.  TEST INFRASTRUCTURE ONLY: used by
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg as the checker / timed baseline; the
 * product path never links or calls it.
 *
 * Restates, in plain C:
 *   softnms_ref     /root/reference/libs/utils/csrc/nms_cpu.cpp:67-160  (softnms_1d_cpu: the swap-based loop,
 *                   methods 0 hard / 1 linear / 2 gaussian, in-place shrink when score < min_score)
 *   hardnms_ref     /root/reference/libs/utils/csrc/nms_cpu.cpp:19-58   (nms_1d_cpu) + the score pre-filter and
 *                   cap of NMSop.forward, /root/reference/libs/utils/nms.py:8-35
 *   batched_nms_ref /root/reference/libs/utils/nms.py:103-190 (per-class loop in ascending class id, cap to
 *                   max_seg_num per class, concatenate, global sort by score, keep max_seg_num), with the
 *                   reference's unspecified tie order pinned to the canonical one of SURVEY.md §8a
 *                   (score descending, then position in the concatenated array ascending)
 *   to_seconds_ref  /root/reference/libs/modeling/multimodal_meta_archs.py:852-856
 *
 * Parity pin: checked against the compiled reference extension (oracle/_ref/nms_1d_cpu.so) and against
 * tests/golden/nms_cases.npz in tests/test_oracle_golden.py.
 * Float semantics: build with -O2 -ffp-contract=off so that no FMA contraction changes the rounding of
 * (area_i + area_j) - inter or -(ovr*ovr)/sigma; exp is libm's expf, as in the reference build.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* Returns the number of surviving candidates n_out; dets[i*3..] = (x1, x2, score) and inds[i] = original
 * index, in selection order, for i < n_out. segs [n,2], scores [n]. */
int softnms_ref(const float* segs, const float* scores, int n, float iou_threshold, float sigma,
                float min_score, int method, float* dets, int64_t* inds) {
  if (n == 0) return 0;
  float* x1 = (float*)malloc(sizeof(float) * n);
  float* x2 = (float*)malloc(sizeof(float) * n);
  float* sc = (float*)malloc(sizeof(float) * n);
  float* ar = (float*)malloc(sizeof(float) * n);
  for (int i = 0; i < n; ++i) {
    x1[i] = segs[2 * i]; x2[i] = segs[2 * i + 1]; sc[i] = scores[i];
    ar[i] = (x2[i] - x1[i]) + 1e-6f;
    inds[i] = i;
  }
  int nsegs = n;
  for (int i = 0; i < nsegs; ++i) {
    float max_score = sc[i];
    int max_pos = i;
    for (int pos = i + 1; pos < nsegs; ++pos)
      if (max_score < sc[pos]) { max_score = sc[pos]; max_pos = pos; }
    float ix1 = dets[i * 3 + 0] = x1[max_pos];
    float ix2 = dets[i * 3 + 1] = x2[max_pos];
    float iscore = dets[i * 3 + 2] = sc[max_pos];
    float iarea = ar[max_pos];
    int64_t iind = inds[max_pos];
    x1[max_pos] = x1[i]; x2[max_pos] = x2[i]; sc[max_pos] = sc[i]; ar[max_pos] = ar[i]; inds[max_pos] = inds[i];
    x1[i] = ix1; x2[i] = ix2; sc[i] = iscore; ar[i] = iarea; inds[i] = iind;
    int pos = i + 1;
    while (pos < nsegs) {
      float xx1 = ix1 > x1[pos] ? ix1 : x1[pos];
      float xx2 = ix2 < x2[pos] ? ix2 : x2[pos];
      float inter = xx2 - xx1;
      if (inter < 0.f) inter = 0.f;
      float ovr = inter / (iarea + ar[pos] - inter);
      float weight = 1.f;
      if (method == 0) { if (ovr >= iou_threshold) weight = 0.f; }
      else if (method == 1) { if (ovr >= iou_threshold) weight = 1.f - ovr; }
      else if (method == 2) { weight = expf(-(ovr * ovr) / sigma); }
      sc[pos] *= weight;
      if (sc[pos] < min_score) {
        x1[pos] = x1[nsegs - 1]; x2[pos] = x2[nsegs - 1]; sc[pos] = sc[nsegs - 1];
        ar[pos] = ar[nsegs - 1]; inds[pos] = inds[nsegs - 1];
        nsegs -= 1;
        pos -= 1;
      }
      pos += 1;
    }
  }
  free(x1); free(x2); free(sc); free(ar);
  return nsegs;
}

typedef struct { float s; int i; } SortItem;
static int cmp_desc(const void* a, const void* b) {
  const SortItem* x = (const SortItem*)a; const SortItem* y = (const SortItem*)b;
  if (x->s > y->s) return -1;
  if (x->s < y->s) return 1;
  return x->i - y->i;        /* canonical tie-break: earlier position first */
}

/* NMSop.forward: filter score > min_score, greedy suppression in descending score order, cap max_num.
 * Returns count; keep[] = original indices in descending score order. */
int hardnms_ref(const float* segs, const float* scores, int n, float iou_threshold, float min_score,
                int max_num, int64_t* keep) {
  SortItem* it = (SortItem*)malloc(sizeof(SortItem) * (n > 0 ? n : 1));
  int m = 0;
  for (int i = 0; i < n; ++i)
    if (!(min_score > 0) || scores[i] > min_score) { it[m].s = scores[i]; it[m].i = i; ++m; }
  qsort(it, m, sizeof(SortItem), cmp_desc);
  char* sel = (char*)malloc(m > 0 ? m : 1);
  memset(sel, 1, m > 0 ? m : 1);
  int cnt = 0;
  for (int a = 0; a < m; ++a) {
    if (!sel[a]) continue;
    int i = it[a].i;
    float ix1 = segs[2 * i], ix2 = segs[2 * i + 1];
    float iarea = (ix2 - ix1) + 1e-6f;
    for (int b = a + 1; b < m; ++b) {
      if (!sel[b]) continue;
      int j = it[b].i;
      float xx1 = ix1 > segs[2 * j] ? ix1 : segs[2 * j];
      float xx2 = ix2 < segs[2 * j + 1] ? ix2 : segs[2 * j + 1];
      float inter = xx2 - xx1;
      if (inter < 0.f) inter = 0.f;
      float jarea = (segs[2 * j + 1] - segs[2 * j]) + 1e-6f;
      float ovr = inter / (iarea + jarea - inter);
      if (ovr >= iou_threshold) sel[b] = 0;
    }
  }
  for (int a = 0; a < m; ++a)
    if (sel[a] && (max_num <= 0 || cnt < max_num)) keep[cnt++] = it[a].i;
  free(it); free(sel);
  return cnt;
}

/* batched_nms (multiclass=True).  labels int64 [n]; outputs sized max_seg_num; returns count.
 * out_src (optional) receives the original candidate index of each kept detection. */
int batched_nms_ref(const float* segs, const float* scores, const int64_t* labels, int n, float iou_threshold,
                    float min_score, int max_seg_num, int use_soft, float sigma, float* out_segs,
                    float* out_scores, int64_t* out_labels, int64_t* out_src) {
  if (n == 0) return 0;
  int64_t maxc = 0;
  for (int i = 0; i < n; ++i) if (labels[i] > maxc) maxc = labels[i];
  float* csegs = (float*)malloc(sizeof(float) * 2 * n);
  float* csc = (float*)malloc(sizeof(float) * n);
  int64_t* cidx = (int64_t*)malloc(sizeof(int64_t) * n);
  float* dets = (float*)malloc(sizeof(float) * 3 * n);
  int64_t* inds = (int64_t*)malloc(sizeof(int64_t) * n);
  /* concatenated per-class results */
  float* asegs = (float*)malloc(sizeof(float) * 2 * n);
  SortItem* all = (SortItem*)malloc(sizeof(SortItem) * n);
  int64_t* alab = (int64_t*)malloc(sizeof(int64_t) * n);
  int64_t* asrc = (int64_t*)malloc(sizeof(int64_t) * n);
  int tot = 0;
  for (int64_t c = 0; c <= maxc; ++c) {          /* torch.unique -> ascending class ids */
    int m = 0;
    for (int i = 0; i < n; ++i)
      if (labels[i] == c) { csegs[2 * m] = segs[2 * i]; csegs[2 * m + 1] = segs[2 * i + 1]; csc[m] = scores[i]; cidx[m] = i; ++m; }
    if (m == 0) continue;
    if (use_soft) {
      int k = softnms_ref(csegs, csc, m, iou_threshold, sigma, min_score, 2, dets, inds);
      if (max_seg_num > 0 && k > max_seg_num) k = max_seg_num;
      for (int i = 0; i < k; ++i) {
        asegs[2 * tot] = dets[3 * i]; asegs[2 * tot + 1] = dets[3 * i + 1];
        all[tot].s = dets[3 * i + 2]; all[tot].i = tot; alab[tot] = c; asrc[tot] = cidx[inds[i]]; ++tot;
      }
    } else {
      int k = hardnms_ref(csegs, csc, m, iou_threshold, min_score, max_seg_num, inds);
      for (int i = 0; i < k; ++i) {
        asegs[2 * tot] = csegs[2 * inds[i]]; asegs[2 * tot + 1] = csegs[2 * inds[i] + 1];
        all[tot].s = csc[inds[i]]; all[tot].i = tot; alab[tot] = c; asrc[tot] = cidx[inds[i]]; ++tot;
      }
    }
  }
  qsort(all, tot, sizeof(SortItem), cmp_desc);
  int k = tot < max_seg_num ? tot : max_seg_num;
  for (int i = 0; i < k; ++i) {
    int j = all[i].i;
    out_segs[2 * i] = asegs[2 * j]; out_segs[2 * i + 1] = asegs[2 * j + 1];
    out_scores[i] = all[i].s; out_labels[i] = alab[j];
    if (out_src) out_src[i] = asrc[j];
  }
  free(csegs); free(csc); free(cidx); free(dets); free(inds); free(asegs); free(all); free(alab); free(asrc);
  return k;
}

/* (segs * stride + 0.5 * nframes) / fps, clamp to [0, duration] the way the reference writes it. */
void to_seconds_ref(float* segs, int n2, float stride, float nframes, float fps, float duration) {
  float half = 0.5f * nframes;
  for (int i = 0; i < n2; ++i) {
    float v = (segs[i] * stride + half) / fps;
    if (v <= 0.0f) v *= 0.0f;
    if (v >= duration) v = v * 0.0f + duration;
    segs[i] = v;
  }
}
