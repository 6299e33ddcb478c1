// ms_row.cuh -- per-row multi-similarity weighting shared by the tuple-mode and flat-mode kernels.
//
// Follows /root/reference/model/losses.py:26-58 (wms_loss) and :95-120 (ms_loss) for one anchor row i:
// given the clamped similarities s_ij and the soft masks wp_ij / wn_ij it produces the row's loss term and
// d loss_row / d s_ij (before the 1/B mean and before the relu gate).
#pragma once
#include "common.cuh"

namespace scl {

// GPS-distance soft masks, model/losses.py:11-19.  The caller subtracts the identity from wp (:22).
__device__ __forceinline__ void wms_masks(float d, float d_alpha, float d_beta, int wfunction, float& wp, float& wn) {
  if (wfunction == SCL_WF_LIN) {
    float r = d / d_beta;
    wp = d < d_beta ? 1.0f - r : 0.0f;
    wn = d < d_beta ? r : 1.0f;
  } else if (wfunction == SCL_WF_TANH) {
    // correctly rounded float32 tanh (float64 evaluation, one rounding): tanhf's last-ulp freedom near saturation would
    // flip the `mask_pos > 0` test of losses.py:50 for every pair beyond ~8.7 * d_beta
    float t = float(tanh(double(d / d_beta)));
    wp = 1.0f - t;
    wn = t;
  } else {
    wp = 1.0f / (1.0f + expf(d_alpha * (d - d_beta)));
    wn = 1.0f / (1.0f + expf(d_alpha * (d_beta - d)));
  }
}

struct MsRowStats {
  float maxv;  // max_j neg_ij            (losses.py:32)
  float tmp;   // max_j pos_ij            (losses.py:33)
  float minv;  // min_j (s_ij-tmp)*wp_ij + tmp   (losses.py:34)
};

// Element-level pieces once the row statistics are known.
//   keptp/keptn : survived mining and mask > 0 (losses.py:36-37 then the >0 tests at :40-41/:50,:53)
//   ep/en       : exp terms of the 'ms' sum (0 when not kept)
__device__ __forceinline__ void ms_elem(float s, float wp, float wn, const MsRowStats& st, const scl_ms_params& p,
                                        bool& keptp, bool& keptn, float& ep, float& en) {
  float pos = s * wp, neg = s * wn;
  float wpm = wp, wnm = wn;
  if (p.ms_mining) {
    wpm = (pos < st.maxv + p.eps) ? wp : 0.0f;
    wnm = (neg > st.minv - p.eps) ? wn : 0.0f;
  }
  keptp = wpm > 0.0f;
  keptn = wnm > 0.0f;
  if (p.sumfunction == SCL_SUM_MS) {
    ep = keptp ? expf(-p.alpha * (pos - p.lamb)) : 0.0f;
    en = keptn ? expf(p.beta * (neg - p.lamb)) : 0.0f;
  } else {
    ep = keptp ? pos : 0.0f;
    en = keptn ? neg : 0.0f;
  }
}

// Row loss and the factors of d loss_row / d s_ij:
//   'ms'   : loss = log(1+A)/alpha + log(1+B)/beta ;  dL/ds = -wp*ep/(1+A) + wn*en/(1+B)
//   'plain': loss = B - A                          ;  dL/ds = -wp*[keptp] + wn*[keptn]
__device__ __forceinline__ float ms_row_loss(float A, float B, const scl_ms_params& p) {
  if (p.sumfunction == SCL_SUM_MS) return logf(1.0f + A) / p.alpha + logf(1.0f + B) / p.beta;
  return B - A;
}
__device__ __forceinline__ float ms_elem_grad(float wp, float wn, bool keptp, bool keptn, float ep, float en, float A,
                                              float B, const scl_ms_params& p) {
  if (p.sumfunction == SCL_SUM_MS) return -wp * __fdividef(ep, 1.0f + A) + wn * __fdividef(en, 1.0f + B);   // gradient only: 2 ulp
  return (keptn ? wn : 0.0f) - (keptp ? wp : 0.0f);
}

}  // namespace scl
