// metrics.cu — detection-mAP matching on the device.
//
// The reference evaluator (libs/utils/metrics.py:306-407) walks every class's detections in descending-score order with
// pandas `iterrows`, and per detection and tIoU threshold greedily takes the not-yet-matched ground-truth segment of the
// same video with the highest tIoU above the threshold.  Matches only interact inside one (class, video) group and one
// threshold, so the whole evaluation is ngroups x nt independent short chains: one thread each.
#include "common.cuh"

namespace unav {

// tIoU in FP64 with the reference's operation order (metrics.py:430-437); no FMA contraction possible (no products).
__device__ __forceinline__ double seg_iou(double t0, double t1, double g0, double g1) {
  const double tt1 = fmax(t0, g0), tt2 = fmin(t1, g1);
  const double inter = fmax(__dsub_rn(tt2, tt1), 0.0);
  const double uni = __dsub_rn(__dadd_rn(__dsub_rn(g1, g0), __dsub_rn(t1, t0)), inter);
  return __ddiv_rn(inter, uni);
}

__global__ void __launch_bounds__(128)
map_match_kernel(const double* __restrict__ det_seg, const double* __restrict__ gt_seg, const int* __restrict__ det_ptr,
                 const int* __restrict__ gt_ptr, int ngroups, const double* __restrict__ tious, int nt, int ndet, int ngt,
                 uint8_t* __restrict__ tp, uint8_t* __restrict__ lock) {
  const long long id = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (id >= static_cast<long long>(ngroups) * nt) return;
  const int g = static_cast<int>(id / nt), t = static_cast<int>(id % nt);
  const double thr = tious[t];
  const int d0 = det_ptr[g], d1 = det_ptr[g + 1], g0 = gt_ptr[2 * g], g1 = gt_ptr[2 * g + 1];
  uint8_t* lk = lock + static_cast<long long>(t) * ngt;
  uint8_t* out = tp + static_cast<long long>(t) * ndet;
  for (int j = g0; j < g1; ++j) lk[j] = 0;
  for (int i = d0; i < d1; ++i) {
    const double t0 = det_seg[2 * i], t1 = det_seg[2 * i + 1];
    // the reference scans the video's ground truth in descending tIoU (ties: later row first, NaN first) and stops at
    // the first one below the threshold; equivalent: best not-yet-matched row among those not below the threshold
    int best = -1;
    double best_key = 0.0;
    for (int j = g0; j < g1; ++j) {
      const double iou = seg_iou(t0, t1, gt_seg[2 * j], gt_seg[2 * j + 1]);
      if (iou < thr) continue;                               // NaN (zero-length union) is not "< thr", as in numpy
      if (lk[j]) continue;
      const double key = (iou != iou) ? CUDART_INF : iou;
      if (best < 0 || key >= best_key) { best = j; best_key = key; }
    }
    if (best >= 0) lk[best] = 1;
    out[i] = best >= 0 ? 1 : 0;
  }
}

}  // namespace unav

extern "C" int unav_map_match(const double* det_seg, const double* gt_seg, const int* det_ptr, const int* gt_ptr, int ngroups,
                              const double* tious, int nt, int ndet, int ngt, uint8_t* tp, uint8_t* lock, void* stream) {
  using namespace unav;
  UNAV_REQUIRE(det_seg && det_ptr && gt_ptr && tious && tp && ngroups > 0 && nt > 0 && ndet > 0, "map_match: bad arguments");
  UNAV_REQUIRE(ngt == 0 || (gt_seg && lock), "map_match: ground truth without buffers");
  const long long n = static_cast<long long>(ngroups) * nt;
  launch_pdl(map_match_kernel, dim3(static_cast<unsigned>((n + 127) / 128)), dim3(128), 0, reinterpret_cast<cudaStream_t>(stream),
             det_seg, gt_seg, det_ptr, gt_ptr, ngroups, tious, nt, ndet, ngt, tp, lock);
  count_launch();
  return finish_launch("map_match");
}
