// FaceBoxes post-processing on the GPU: prior boxes, box decode, score filter, top-k ordering.
// Replaces FaceBoxes/FaceBoxes.py:98-120 (+ utils/prior_box.py:12-48, utils/box_utils.py:177-195); the greedy NMS that
// follows (:122-127) is nms_mask_kernel / nms_scan_kernel in kernels_render.cuh.
//
// Prior boxes are a closed form of the prior index (no table in memory): the reference builds them in Python doubles
// and rounds to float32 once (`torch.Tensor(anchors)`), which is what prior_of() does.  The decode is float32 torch
// arithmetic, one rounding per operation; only exp() has no bit pattern to match (torch's CPU kernel is a Sleef
// vector exp), so boxes agree to ~1e-7 relative and the ordering / NMS index work is exact given equal scores.
#pragma once
#include "common.cuh"
#include "render_math.h"

namespace syn {

// cfg of FaceBoxes/utils/config.py: min_sizes [[32, 64, 128], [256], [512]], steps [32, 64, 128], variance [0.1, 0.2]
__host__ __device__ inline int fb_cells(int size, int step) { return (size + step - 1) / step; }      // ceil(size / step), prior_box.py:19
__host__ __device__ inline int faceboxes_num_priors(int h, int w) {
  return 21 * fb_cells(h, 32) * fb_cells(w, 32) + fb_cells(h, 64) * fb_cells(w, 64) + fb_cells(h, 128) * fb_cells(w, 128);
}

// prior `idx` -> (cx, cy, s_kx, s_ky), the order of prior_box.py:23-43: level, row i, column j, then 16 + 4 + 1 anchors
__device__ inline void prior_of(int idx, int h, int w, float* p) {
  const int n0 = 21 * fb_cells(h, 32) * fb_cells(w, 32), n1 = fb_cells(h, 64) * fb_cells(w, 64);
  double cx, cy, ms, step;
  if (idx < n0) {
    const int cell = idx / 21, a = idx - cell * 21, cols = fb_cells(w, 32);
    const int i = cell / cols, j = cell - i * cols;
    step = 32.0;
    if (a < 16) { ms = 32.0; cx = j + 0.25 * (a & 3); cy = i + 0.25 * (a >> 2); }
    else if (a < 20) { ms = 64.0; cx = j + 0.5 * ((a - 16) & 1); cy = i + 0.5 * ((a - 16) >> 1); }
    else { ms = 128.0; cx = j + 0.5; cy = i + 0.5; }
  } else if (idx < n0 + n1) {
    const int cell = idx - n0, cols = fb_cells(w, 64);
    step = 64.0; ms = 256.0; cx = cell % cols + 0.5; cy = cell / cols + 0.5;
  } else {
    const int cell = idx - n0 - n1, cols = fb_cells(w, 128);
    step = 128.0; ms = 512.0; cx = cell % cols + 0.5; cy = cell / cols + 0.5;
  }
  p[0] = (float)(cx * step / (double)w);
  p[1] = (float)(cy * step / (double)h);
  p[2] = (float)(ms / (double)w);
  p[3] = (float)(ms / (double)h);
}

// cand[0] = number of priors whose face score conf[:, 1] exceeds the threshold (FaceBoxes.py:112), cand[1..] = their indices
__global__ void faceboxes_select_kernel(const float* __restrict__ conf, int np, float thresh, int32_t* __restrict__ cand) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < np && conf[2 * i + 1] > thresh) cand[1 + atomicAdd(cand, 1)] = i;
}

// Rank of every candidate in descending score order (ties: the higher prior index first = a stable ascending argsort
// read backwards, :117), the first top_k decoded and written in that order as rows [x1 y1 x2 y2 score] (:121).
__global__ void faceboxes_rank_decode_kernel(const float* __restrict__ loc, const float* __restrict__ conf, int h, int w,
                                             float box_scale_w, float box_scale_h, float scale, int top_k,
                                             const int32_t* __restrict__ cand, float* __restrict__ dets, int32_t* __restrict__ n_dets) {
  const int n = cand[0];
  if (blockIdx.x == 0 && threadIdx.x == 0) *n_dets = min(n, top_k);
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < n; c += gridDim.x * blockDim.x) {
    const int i = cand[1 + c];
    const float s = conf[2 * i + 1];
    int rank = 0;
    for (int q = 0; q < n; ++q) {
      const int j = cand[1 + q];
      const float sj = conf[2 * j + 1];
      rank += (sj > s) || (sj == s && j > i);
    }
    if (rank >= top_k) continue;
    float p[4];
    prior_of(i, h, w, p);
    const float* l = loc + 4 * (size_t)i;
    using namespace rmath;
    const float cx = add(p[0], mul(mul(l[0], 0.1f), p[2])), cy = add(p[1], mul(mul(l[1], 0.1f), p[3]));       // box_utils.py:191
    const float bw = mul(p[2], expf(mul(l[2], 0.2f))), bh = mul(p[3], expf(mul(l[3], 0.2f)));                  // :192
    const float x1 = sub(cx, dvd(bw, 2.0f)), y1 = sub(cy, dvd(bh, 2.0f));                                       // :193
    const float x2 = add(bw, x1), y2 = add(bh, y1);                                                             // :194
    float* o = dets + 5 * (size_t)rank;
    o[0] = dvd(mul(x1, box_scale_w), scale);                                                                    // FaceBoxes.py:104
    o[1] = dvd(mul(y1, box_scale_h), scale);
    o[2] = dvd(mul(x2, box_scale_w), scale);
    o[3] = dvd(mul(y2, box_scale_h), scale);
    o[4] = s;
  }
}

}  // namespace syn
