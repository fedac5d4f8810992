// fit_head.cuh -- the plain fit and its backward pass fed by the NOC HEAD OUTPUT instead of a materialised NOC crop
// (the ROI-align resize of Detection/tracker/postprocess.py:141-147 fused into the fit's loaders).
// Part of libposefit_b200.so: included by posefit_kernels.cu.
//
// The reference resizes the 3 x 28 x 28 sigmoid output of the NOC head (nocs_head.py:232-235) to every instance's box
// with roi_align and hands the h x w x 3 patch to run_pose.  Materialising that patch costs 12 B/pixel to write and 12
// B/pixel to read again -- more than everything else the fit reads (depth 4 + mask 1) -- and the same again for its
// gradient on the way back.  Here the 9.4 KB head output of an object is staged in shared memory and every pixel's NOC
// value is sampled from it on the fly, with the arithmetic of K-resample (aux_kernels.cuh: sample_head, shared by both,
// so the sampled values are bit-identical to a materialised crop's); the backward kernel scatters the NOC gradient
// straight into a shared-memory copy of the head gradient.  HBM traffic per 64x64 object: forward 20 KB + 9.4 KB
// instead of 69.6 KB, backward 20 KB + 9.4 KB read and 9.4 KB written instead of 69.6 KB read and 49 KB written --
// and no K-resample / K-resample-backward launches (another 2 x 49 KB each way).  The price is arithmetic: ~50
// instructions per pixel for the taps, so these kernels are issue-bound, not HBM-bound.
//
//   fit_head_moments_kernel   one 256-thread CTA per object at a time (persistent grid): head -> shared memory, the
//                             depth / mask crop streamed once with 128-bit loads, 16 double sums + count per thread,
//                             block reduction -> ONE partial-moment record per object; fit_solve_kernel does the rest.
//   fit_head_backward_kernel  same staging; per pixel the NOC value (for x~), the gradient of fit_backward_kernel
//                             (bwd_point) and its adjoint through the bilinear taps: shared-memory atomics into the
//                             head gradient, written out once per object.  grad_depth as in fit_backward_kernel.
#pragma once

#include "aux_kernels.cuh"
#include "fit_backward.cuh"
#include "fit_moments.cuh"

namespace posefit {

constexpr int kHeadThreads = 256;

struct HeadParams {
  const float* head;          // [B][3][Hh][Wh]
  const int32_t* roi_hw;      // [B][2] size (h_i, w_i) the head output is resized to (the instance's box)
  const float* depth;         // [B][H][W]
  const uint8_t* mask;        // [B][H][W]
  const uint8_t* inlier_mask; // backward only, may be NULL
  const int32_t* bbox;
  const double* kinv;
  int kinv_per_object;
  int B, Hh, Wh, H, W, P;
  double* ws;                 // forward: [B][17] moment records
  const BwdCoef* coef;        // backward: per-object adjoint coefficients (fit_backward_coef_kernel)
  float* grad_head;           // backward: [B][3][Hh][Wh]
  float* grad_depth;          // backward: [B][H][W] or NULL
  int vec_ok;
};

// geometry of the resize of one object (roi_align over the whole map, aligned = True, sampling_ratio = -1)
struct HeadGeom {
  int oh, ow, grid_h, grid_w;
  float bin_h, bin_w, count;
};

__device__ __forceinline__ HeadGeom head_geom(const int32_t* roi_hw, int obj, int Hh, int Wh) {
  HeadGeom g;
  g.oh = roi_hw[2 * obj];
  g.ow = roi_hw[2 * obj + 1];
  const float roi_h = (float)Hh, roi_w = (float)Wh;
  g.bin_h = g.oh > 0 ? roi_h / (float)g.oh : 0.0f;
  g.bin_w = g.ow > 0 ? roi_w / (float)g.ow : 0.0f;
  g.grid_h = g.oh > 0 ? (int)ceilf(roi_h / (float)g.oh) : 1;
  g.grid_w = g.ow > 0 ? (int)ceilf(roi_w / (float)g.ow) : 1;
  g.count = (float)max(g.grid_h * g.grid_w, 1);
  return g;
}

template <bool BACKWARD>
__global__ void __launch_bounds__(kHeadThreads, 2) fit_head_kernel(const HeadParams p) {
  extern __shared__ __align__(16) float hsm[];                 // head [3][Hh][Wh] | (backward) head gradient | tables | red
  const int hw = p.Hh * p.Wh;
  float* smap = hsm;
  float* sgrad = hsm + 3 * hw;                                 // backward only
  double* rxc = reinterpret_cast<double*>(hsm + (BACKWARD ? 6 : 3) * hw + ((BACKWARD ? 6 : 3) * hw & 1));
  double* ryr = rxc + p.W;
  double* red = ryr + p.H;                                     // [8][24] + [24]
  double* mom = red + (kHeadThreads / 32) * 24;
  __shared__ BwdCoef coef_s;
#if __CUDA_ARCH__ >= 900
  asm volatile("griddepcontrol.launch_dependents;");
  asm volatile("griddepcontrol.wait;" ::: "memory");            // head / ctx / coefficients come from kernels before
#endif
  const int tid = threadIdx.x;
  const int P = p.P;
  for (int obj = blockIdx.x; obj < p.B; obj += gridDim.x) {
    // ---- stage the head output (and, backward, clear the gradient copy and fetch the coefficients) ----------------
    const float* head = p.head + (size_t)obj * 3 * hw;
    for (int i = tid; i < 3 * hw; i += kHeadThreads) {
      smap[i] = head[i];
      if (BACKWARD) sgrad[i] = 0.0f;
    }
    if (BACKWARD && tid < 36) reinterpret_cast<int*>(&coef_s)[tid] = reinterpret_cast<const int*>(p.coef + obj)[tid];
    const double* K = p.kinv + (p.kinv_per_object ? 9 * (size_t)obj : 0);
    ObjGeom g;
    g.k = K;
    g.k0 = K[0]; g.k2 = K[2]; g.k4 = K[4]; g.k5 = K[5];
    g.x0 = p.bbox[2 * (size_t)obj];
    g.y0 = p.bbox[2 * (size_t)obj + 1];
    g.simple = (K[1] == 0.0 && K[3] == 0.0 && K[6] == 0.0 && K[7] == 0.0 && K[8] == 1.0);
    if (!BACKWARD) {
      for (int i = tid; i < p.W; i += kHeadThreads) rxc[i] = g.k0 * (double)(g.x0 + i) + g.k2;
      for (int i = tid; i < p.H; i += kHeadThreads) ryr[i] = g.k4 * (double)(g.y0 + i) + g.k5;
    }
    const HeadGeom hg = head_geom(p.roi_hw, obj, p.Hh, p.Wh);
    __syncthreads();
    const size_t ob = (size_t)obj * P;
    LaneSums acc;
    acc.clear();
    const bool live = !BACKWARD || coef_s.live != 0;
    // ---- the crop, 4 consecutive pixels per thread and iteration (scalar tail for ragged shapes) --------------------
    const int step = p.vec_ok ? 4 : 1;
    for (int i = tid * step; i < P; i += kHeadThreads * step) {
      float zz[4] = {0.f, 0.f, 0.f, 0.f};
      uint32_t m4 = 0u, im4 = 0x01010101u;
      if (p.vec_ok) {
        const float4 z4 = __ldcs(reinterpret_cast<const float4*>(p.depth + ob + i));
        zz[0] = z4.x; zz[1] = z4.y; zz[2] = z4.z; zz[3] = z4.w;
        m4 = __ldcs(reinterpret_cast<const uint32_t*>(p.mask + ob + i));
        if (BACKWARD && p.inlier_mask) im4 = __ldcs(reinterpret_cast<const uint32_t*>(p.inlier_mask + ob + i));
      } else {
        zz[0] = p.depth[ob + i];
        m4 = p.mask[ob + i];
        if (BACKWARD && p.inlier_mask) im4 = p.inlier_mask[ob + i];
      }
      const int row = i / p.W, col0 = i - row * p.W;
      float gz[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (j >= step) break;
        const int col = col0 + j;
        const bool ok = (m4 & (0xffu << (8 * j))) != 0u && zz[j] > 0.0f &&       // pose_estimation.py:23-25
                        (!BACKWARD || (im4 & (0xffu << (8 * j))) != 0u) && live;
        if (!ok) continue;
        // the NOC value the reference's roi_align would put at (row, col): zero outside the instance's box (padding)
        float noc[3] = {0.f, 0.f, 0.f};
        const bool inside = row < hg.oh && col < hg.ow;
        if (inside) sample_head<false>(smap, nullptr, hw, p.Hh, p.Wh, row, col, hg.bin_h, hg.bin_w, hg.grid_h, hg.grid_w, hg.count, noc);
        if (!BACKWARD) {
          double y0, y1, y2;
          backproject_px(g, rxc, ryr, row, col, (double)zz[j], y0, y1, y2);
          acc.add((double)noc[0], (double)noc[1], (double)noc[2], y0, y1, -y2);
          ++acc.cnt;
        } else {
          float g0, g1, g2;
          bwd_point(coef_s, noc[0], noc[1], noc[2], zz[j], true, row, col, g0, g1, g2, gz[j]);
          if (inside) {
            float gn[3] = {g0, g1, g2};
            sample_head<true>(nullptr, sgrad, hw, p.Hh, p.Wh, row, col, hg.bin_h, hg.bin_w, hg.grid_h, hg.grid_w, hg.count, gn);
          }
        }
      }
      if (BACKWARD && p.grad_depth) {
        if (p.vec_ok) __stcs(reinterpret_cast<float4*>(p.grad_depth + ob + i), make_float4(gz[0], gz[1], gz[2], gz[3]));
        else p.grad_depth[ob + i] = gz[0];
      }
    }
    if (!BACKWARD) {
      // one partial-moment record per object (fit_solve_kernel merges "1 part")
      double out[kAccPlain];
      acc.finish(0.5, true, out);                                  // warp totals, valid in every lane
      const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
      for (int i = 0; i < kAccPlain; ++i)
        if (lane == i) red[warp * 24 + i] = out[i];
      __syncthreads();
      if (tid < kAccPlain) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < kHeadThreads / 32; ++w) s += red[w * 24 + tid];
        p.ws[(size_t)obj * kAccPlain + tid] = s;
      }
      (void)mom;
    } else {
      __syncthreads();
      float* gh = p.grad_head + (size_t)obj * 3 * hw;
      for (int i = tid; i < 3 * hw; i += kHeadThreads) gh[i] = sgrad[i];
    }
    __syncthreads();                                               // smap / sgrad / red are free again
  }
}

}  // namespace posefit
