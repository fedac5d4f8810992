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
//   fit_head_kernel<false>    one 256-thread CTA per object at a time (persistent grid): head -> shared memory, per-row /
//                             per-column tap tables, the depth / mask crop streamed once with 128-bit loads, 16 double
//                             sums + count per thread, block reduction -> ONE partial-moment record per object;
//                             fit_solve_kernel does the rest.
//   fit_head_kernel<true>     same staging; per pixel the NOC value (for x~), the gradient of fit_backward_kernel
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

// Software pipeline of one CTA over its objects: the head output of object k+1 is requested with cp.async (16-byte
// copies, no register staging) into the other half of a double buffer while object k is processed, and an object's
// depth / mask lines are prefetched into L2 before the barrier that waits for its head.  The pixel loop is rolled
// (one 4-pixel group per trip): unrolled, its body -- taps, conversions, 20 double sums per pixel -- was 6.7 k
// instructions and ran out of the instruction cache.

template <bool BACKWARD>
__global__ void __launch_bounds__(kHeadThreads, 2) fit_head_kernel(const HeadParams p) {
  extern __shared__ __align__(16) float hsm[];         // head[2][3 hw] | (backward) head gradient | tap tables | ray tables | red
  const int hw = p.Hh * p.Wh;
  const int hw3 = (3 * hw + 3) & ~3;                            // 16-byte granules
  float* sgrad = hsm + 2 * hw3;                                 // backward only
  TapEntry* rows = reinterpret_cast<TapEntry*>(hsm + (BACKWARD ? 3 : 2) * hw3);
  TapEntry* cols = rows + (p.H > 2 * p.Hh ? p.H : 2 * p.Hh);
  double* rxc = reinterpret_cast<double*>(cols + (p.W > 2 * p.Wh ? p.W : 2 * p.Wh));
  double* ryr = rxc + p.W;
  double* red = ryr + p.H;                                      // [8][24]
  __shared__ __align__(16) BwdCoef coef_s[2];
#if __CUDA_ARCH__ >= 900
  asm volatile("griddepcontrol.launch_dependents;");
  asm volatile("griddepcontrol.wait;" ::: "memory");            // head / ctx / coefficients come from kernels before
#endif
  const int tid = threadIdx.x;
  const int P = p.P;
  const bool head16 = (hw * 3) % 4 == 0 && (reinterpret_cast<uintptr_t>(p.head) & 15u) == 0;
  auto request = [&](int obj, int buf) {                        // head (and coefficients) of `obj` -> buffer `buf`
    if (obj < p.B) {
      const float* head = p.head + (size_t)obj * 3 * hw;
      float* dst = hsm + buf * hw3;
      if (head16) {
        for (int i = tid; i < 3 * hw / 4; i += kHeadThreads) cp_async_16(dst + 4 * i, head + 4 * i);
      } else {
        for (int i = tid; i < 3 * hw; i += kHeadThreads) cp_async_4(dst + i, head + i);
      }
      if (BACKWARD && tid < 9)
        cp_async_16(reinterpret_cast<unsigned char*>(&coef_s[buf]) + 16 * tid,
                    reinterpret_cast<const unsigned char*>(p.coef + obj) + 16 * tid);
    }
    cp_async_commit();
  };
  request((int)blockIdx.x, 0);
  int k = 0;
  for (int obj = blockIdx.x; obj < p.B; obj += gridDim.x, ++k) {
    const int buf = k & 1;
    float* smap = hsm + buf * hw3;
    request(obj + (int)gridDim.x, buf ^ 1);                     // next object's head, behind this object's work
    // ---- this object's crop: pulled into L2 now (no registers held), read in the pixel loop below ---------------------
    const size_t ob = (size_t)obj * P;
    const int step = p.vec_ok ? 4 : 1;
    if (p.vec_ok) {
      for (int i = tid * 32; i < P; i += kHeadThreads * 32) {     // one 128-byte line of depth per prefetch
        asm volatile("prefetch.global.L2 [%0];" ::"l"(p.depth + ob + i));
        if ((tid & 3) == 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.mask + ob + i));
      }
    }
    const double* K = p.kinv + (p.kinv_per_object ? 9 * (size_t)obj : 0);
    ObjGeom g;
    g.k = K;
    g.k0 = K[0]; g.k2 = K[2]; g.k4 = K[4]; g.k5 = K[5];
    g.x0 = p.bbox[2 * (size_t)obj];
    g.y0 = p.bbox[2 * (size_t)obj + 1];
    g.simple = (K[1] == 0.0 && K[3] == 0.0 && K[6] == 0.0 && K[7] == 0.0 && K[8] == 1.0);
    const HeadGeom hg = head_geom(p.roi_hw, obj, p.Hh, p.Wh);
    if (!BACKWARD) {
      for (int i = tid; i < p.W; i += kHeadThreads) rxc[i] = g.k0 * (double)(g.x0 + i) + g.k2;
      for (int i = tid; i < p.H; i += kHeadThreads) ryr[i] = g.k4 * (double)(g.y0 + i) + g.k5;
    } else {
      for (int i = tid; i < 3 * hw; i += kHeadThreads) sgrad[i] = 0.0f;
    }
    build_tap_tables(rows, cols, min(hg.oh, p.H), min(hg.ow, p.W), hg.bin_h, hg.bin_w, hg.grid_h, hg.grid_w, p.Hh, p.Wh, tid,
                     kHeadThreads);
    cp_async_wait_group<1>();                                   // this object's head (requested one object ago) has landed
    __syncthreads();
    LaneSums acc;
    acc.clear();
    const BwdCoef& cf = coef_s[buf];
    const bool live = !BACKWARD || cf.live != 0;
    const bool one_tap = hg.grid_h == 1 && hg.grid_w == 1;       // box >= map on both axes: one bilinear tap per pixel
    // ---- pixels: 4 consecutive ones per thread and iteration (1 for ragged shapes) ------------------------------------
    {
#pragma unroll 1
      for (int i = tid * step; i < P; i += kHeadThreads * step) {
        float4 zq;
        uint32_t mq, imq = 0x01010101u;
        if (p.vec_ok) {
          zq = __ldcs(reinterpret_cast<const float4*>(p.depth + ob + i));
          mq = __ldcs(reinterpret_cast<const uint32_t*>(p.mask + ob + i));
          if (BACKWARD && p.inlier_mask) imq = __ldcs(reinterpret_cast<const uint32_t*>(p.inlier_mask + ob + i));
        } else {
          zq = make_float4(p.depth[ob + i], 0.f, 0.f, 0.f);
          mq = p.mask[ob + i];
          if (BACKWARD && p.inlier_mask) imq = p.inlier_mask[ob + i];
        }
        const float zz[4] = {zq.x, zq.y, zq.z, zq.w};
        const int row = i / p.W, col0 = i - row * p.W;
        float gz[4] = {0.f, 0.f, 0.f, 0.f};
        TapEntry ry = {0, 0, 0.f, 0.f};
        if (one_tap && row < hg.oh) ry = rows[row];                  // the 4 pixels of a group share their row
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int col = col0 + j;
          const bool ok = j < step && (mq & (0xffu << (8 * j))) != 0u && zz[j] > 0.0f &&       // pose_estimation.py:23-25
                          (!BACKWARD || (imq & (0xffu << (8 * j))) != 0u) && live;
          if (!ok) continue;
          // the NOC value the reference's roi_align would put at (row, col): zero outside the instance's box (padding)
          float noc[3] = {0.f, 0.f, 0.f};
          const bool inside = row < hg.oh && col < hg.ow;
          if (inside) {
            if (one_tap) sample_head_1tap<false>(smap, nullptr, hw, ry, cols[col], noc);
            else sample_head<false>(smap, nullptr, hw, rows, cols, row, col, hg.grid_h, hg.grid_w, hg.count, noc);
          }
          if (!BACKWARD) {
            double y0, y1, y2;
            backproject_px(g, rxc, ryr, row, col, (double)zz[j], y0, y1, y2);
            acc.add((double)noc[0], (double)noc[1], (double)noc[2], y0, y1, -y2);
            ++acc.cnt;
          } else {
            float g0, g1, g2;
            bwd_point(cf, noc[0], noc[1], noc[2], zz[j], true, row, col, g0, g1, g2, gz[j]);
            if (inside) {
              float gn[3] = {g0, g1, g2};
              if (one_tap) sample_head_1tap<true>(nullptr, sgrad, hw, ry, cols[col], gn);
              else sample_head<true>(nullptr, sgrad, hw, rows, cols, row, col, hg.grid_h, hg.grid_w, hg.count, gn);
            }
          }
        }
        if (BACKWARD && p.grad_depth) {
          if (p.vec_ok) __stcs(reinterpret_cast<float4*>(p.grad_depth + ob + i), make_float4(gz[0], gz[1], gz[2], gz[3]));
          else p.grad_depth[ob + i] = gz[0];
        }
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
        double sum = 0.0;
#pragma unroll
        for (int w = 0; w < kHeadThreads / 32; ++w) sum += red[w * 24 + tid];
        p.ws[(size_t)obj * kAccPlain + tid] = sum;
      }
    } else {
      __syncthreads();
      float* gh = p.grad_head + (size_t)obj * 3 * hw;
      for (int i = tid; i < 3 * hw; i += kHeadThreads) gh[i] = sgrad[i];
    }
    __syncthreads();                                               // tables / sgrad / red / this head buffer are free again
  }
}

}  // namespace posefit
