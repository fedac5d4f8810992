// fit_backward.cuh -- K-backward: adjoint of the fit (no reference counterpart, postprocess.py:151 detaches)
// Part of libposefit_b200.so: included by posefit_kernels.cu (one translation unit, so every kernel sees the
// same inlined helpers and the build stays a single nvcc call).  See include/posefit.h for the C ABI.
#pragma once

#include "posefit_common.cuh"

namespace posefit {

// ---------------------------------------------------------------------------------------------
// K-backward
// ---------------------------------------------------------------------------------------------
struct BwdCoef;

struct BwdParams {
  const float* noc;
  const float* depth;
  const uint8_t* mask;
  const uint8_t* inlier_mask;
  const int32_t* bbox;
  const double* kinv;
  const double* ctx;
  const int32_t* status;
  const float* g_scale;
  const float* g_R;
  const float* g_t;
  float* grad_noc;
  float* grad_depth;
  BwdCoef* coef;          // [B] workspace
  int kinv_per_object;
  int B, H, W, P;
  int chunk_px, chunks_per_obj;
  int vec_ok;
  int early_dep;
  int opc;                // K-backward-coef: objects per CTA (posefit_common.cuh: solve_object)
  unsigned int* dyn_counter;   // long launches: ticket counter of K-backward's units (zeroed by K-backward-coef), else NULL
};

struct BwdCoef {          // per-object coefficients, already scaled by 1/n
  float GC[9];            // row-major G_C
  float gvar2;            // 2 * g_var
  float gmux[3], gmuy[3];
  float mux[3], muy[3];
  float k[9];
  int x0, y0;
  int live, simple;
  int pad;                // sizeof == 144 == 9 x 16 bytes (fetched with cp.async)
};

__device__ __forceinline__ void bwd_point(const BwdCoef& c, float n0, float n1, float n2, float z, bool w, int row, int col,
                                          float& g0, float& g1, float& g2, float& gz) {
  g0 = g1 = g2 = gz = 0.0f;
  if (!w) return;
  const float u = (float)(c.x0 + col), v = (float)(c.y0 + row);
  float rx, ry, rz;
  if (c.simple) {
    rx = fmaf(c.k[0], u, c.k[2]);
    ry = fmaf(c.k[4], v, c.k[5]);
    rz = 1.0f;
  } else {
    const float Z = c.k[6] * u + c.k[7] * v + c.k[8];
    rx = (c.k[0] * u + c.k[1] * v + c.k[2]) / Z;
    ry = (c.k[3] * u + c.k[4] * v + c.k[5]) / Z;
    rz = 1.0f;
  }
  const float yt0 = rx * z - c.muy[0], yt1 = -(ry * z) - c.muy[1], yt2 = -(rz * z) - c.muy[2];
  const float xt0 = (n0 - 0.5f) - c.mux[0], xt1 = (n1 - 0.5f) - c.mux[1], xt2 = (n2 - 0.5f) - c.mux[2];
  // dL/dx = GC^T y~ + 2 gvar x~ + gmux
  g0 = c.GC[0] * yt0 + c.GC[3] * yt1 + c.GC[6] * yt2 + c.gvar2 * xt0 + c.gmux[0];
  g1 = c.GC[1] * yt0 + c.GC[4] * yt1 + c.GC[7] * yt2 + c.gvar2 * xt1 + c.gmux[1];
  g2 = c.GC[2] * yt0 + c.GC[5] * yt1 + c.GC[8] * yt2 + c.gvar2 * xt2 + c.gmux[2];
  // dL/dy = GC x~ + gmuy ; y = (rx z, -ry z, -z)
  const float h0 = c.GC[0] * xt0 + c.GC[1] * xt1 + c.GC[2] * xt2 + c.gmuy[0];
  const float h1 = c.GC[3] * xt0 + c.GC[4] * xt1 + c.GC[5] * xt2 + c.gmuy[1];
  const float h2 = c.GC[6] * xt0 + c.GC[7] * xt1 + c.GC[8] * xt2 + c.gmuy[2];
  gz = rx * h0 - ry * h1 - rz * h2;
}

// Per-object coefficients of the adjoint (fp64, ~300 dependent instructions) from the saved
// context and the upstream gradients.
__device__ __forceinline__ void bwd_coefficients(const BwdParams& p, int obj, BwdCoef& coef) {
  const double* cx = p.ctx + (size_t)obj * POSEFIT_CTX_DOUBLES;
  Fit f;
#pragma unroll
  for (int i = 0; i < 9; ++i) f.R[i] = cx[i];
#pragma unroll
  for (int i = 0; i < 6; ++i) { f.Linv[i] = cx[9 + i]; f.H[i] = cx[15 + i]; }
  f.s = cx[21];
  f.var = cx[22];
  f.n = cx[23];
#pragma unroll
  for (int i = 0; i < 3; ++i) { f.mux[i] = cx[24 + i]; f.muy[i] = cx[27 + i]; }
  const bool live = (p.status[obj] == PF_OK) && (f.n > 0.0);
  double gR[9], gt[3];
  const double gs = p.g_scale ? (double)p.g_scale[obj] : 0.0;
#pragma unroll
  for (int i = 0; i < 9; ++i) gR[i] = p.g_R ? (double)p.g_R[(size_t)obj * 9 + i] : 0.0;
#pragma unroll
  for (int i = 0; i < 3; ++i) gt[i] = p.g_t ? (double)p.g_t[(size_t)obj * 3 + i] : 0.0;
  FitAdjoint a;
  fit_adjoint(f, gs, gR, gt, a);
  const double rn = live ? 1.0 / f.n : 0.0;
#pragma unroll
  for (int i = 0; i < 9; ++i) coef.GC[i] = (float)(a.GC[i] * rn);
  coef.gvar2 = (float)(2.0 * a.gvar * rn);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    coef.gmux[i] = (float)(a.gmux[i] * rn);
    coef.gmuy[i] = (float)(a.gmuy[i] * rn);
    coef.mux[i] = (float)f.mux[i];
    coef.muy[i] = (float)f.muy[i];
  }
  const double* K = p.kinv + (p.kinv_per_object ? 9 * (size_t)obj : 0);
#pragma unroll
  for (int i = 0; i < 9; ++i) coef.k[i] = (float)K[i];
  coef.simple = (K[1] == 0.0 && K[3] == 0.0 && K[6] == 0.0 && K[7] == 0.0 && K[8] == 1.0);
  coef.x0 = p.bbox[2 * obj];
  coef.y0 = p.bbox[2 * obj + 1];
  coef.live = live ? 1 : 0;
}

// One thread per object: adjoint coefficients -> coef[B] (144 B each) in the workspace.
__global__ void __launch_bounds__(128) fit_backward_coef_kernel(const BwdParams p) {
#if __CUDA_ARCH__ >= 900
  PF_TRACE_BEGIN(2);
  if (p.early_dep & 4) asm volatile("griddepcontrol.launch_dependents;");   // K-backward's CTAs queue up behind us
  asm volatile("griddepcontrol.wait;" ::: "memory");          // ctx / status come from the forward kernels
  if (!(p.early_dep & 4)) asm volatile("griddepcontrol.launch_dependents;");
#endif
  PF_TRACE_BEGIN(10);
  if (p.dyn_counter != nullptr && blockIdx.x == 0 && threadIdx.x == 0) *p.dyn_counter = 0u;   // K-backward's tickets
  const int o = solve_object(p.opc, p.B);
  if (o >= p.B) return;
  BwdCoef c;
  bwd_coefficients(p, o, c);
  p.coef[o] = c;
  PF_TRACE_END(2);
}

struct BwdLoad {
  float4 a0, a1, a2, zz;
  uchar4 mm, im;
};

// Streaming pass: (object, chunk) units; the 144-byte coefficient record of the NEXT unit is fetched with cp.async
// while the current one streams, so no fp64 and no global-load latency sit between units.  Two iterations of loads
// are issued before the first is consumed.  MINB = 3: 80 registers, nothing spilled, 24 warps per SM -- measured
// 6.19 TB/s on the config-5 shard against 5.88 TB/s for MINB = 4 (64 registers, 36 bytes of spills, 32 warps).
// (Computing the record inside this kernel instead of fit_backward_coef_kernel -- one thread per CTA, behind the
// streaming of the previous unit -- was built and measured on BASELINE config 4: 64.7 us against 64.9 us, i.e. the
// separate kernel already hides behind the programmatic dependent launch; not kept.)
template <int NT, int MINB, bool DYN = false>
__global__ void __launch_bounds__(NT, MINB) fit_backward_kernel(const BwdParams p) {
  __shared__ __align__(16) BwdCoef coefs[2];
  static_assert(sizeof(BwdCoef) == 144, "coefficient record is 9 x 16 bytes");
  PF_TRACE_BEGIN(3);
  PF_TRACE_END(6);                                              // (latest first instruction)
#if __CUDA_ARCH__ >= 900
  if (p.early_dep & 8) asm volatile("griddepcontrol.launch_dependents;");   // the next call's first kernel may queue up
  asm volatile("griddepcontrol.wait;" ::: "memory");            // coefficients written by fit_backward_coef_kernel
  if (!(p.early_dep & 8)) asm volatile("griddepcontrol.launch_dependents;");
#endif
  PF_TRACE_BEGIN(11);
  const int tid = threadIdx.x;
  const int n_units = p.B * p.chunks_per_obj;
  auto provide = [&](int unit, int buf) {
    if (tid < 9 && unit < n_units)
      cp_async_16(reinterpret_cast<unsigned char*>(&coefs[buf]) + 16 * tid,
                  reinterpret_cast<const unsigned char*>(p.coef + unit / p.chunks_per_obj) + 16 * tid);
    cp_async_commit();
  };
  auto load = [&](size_t ob, const float* n0p, int i, BwdLoad& d) {
    d.a0 = __ldcs(reinterpret_cast<const float4*>(n0p + i));
    d.a1 = __ldcs(reinterpret_cast<const float4*>(n0p + p.P + i));
    d.a2 = __ldcs(reinterpret_cast<const float4*>(n0p + 2 * (size_t)p.P + i));
    d.zz = __ldcs(reinterpret_cast<const float4*>(p.depth + ob + i));
    d.mm = __ldcs(reinterpret_cast<const uchar4*>(p.mask + ob + i));
    d.im = make_uchar4(1, 1, 1, 1);
    if (p.inlier_mask) d.im = __ldcs(reinterpret_cast<const uchar4*>(p.inlier_mask + ob + i));
  };
  // Which unit comes next: blockIdx + k * gridDim (short launches), or -- long launches, where SMs that stream more slowly
  // than others would otherwise set the kernel's time -- the first two of them and then tickets from p.dyn_counter.
  // Thread 0 draws the ticket of iteration k + 2 at the top of iteration k and publishes it at the bottom (the value is
  // back by then); the barrier at the top of iteration k + 1 makes it visible, one iteration before the unit is needed
  // (its coefficient record is requested a unit ahead).
  __shared__ int tickets[2];
  constexpr bool dyn = DYN;                          // (its own instantiation: the short-launch kernel keeps its registers)
  provide((int)blockIdx.x, 0);
  int k = 0;
  int nxt = (int)blockIdx.x + (int)gridDim.x;
  for (int unit = blockIdx.x; unit < n_units; unit = nxt, ++k) {
    const int obj = unit / p.chunks_per_obj;
    const int ch = unit - obj * p.chunks_per_obj;
    const int px0 = ch * p.chunk_px;
    const int px1 = min(px0 + p.chunk_px, p.P);
    PF_CHECK(obj >= 0 && obj < p.B && px0 < p.P);
    const size_t ob = (size_t)obj * p.P;
    const float* n0p = p.noc + ob * 3;
    float* g0p = p.grad_noc + ob * 3;
    float* g1p = g0p + p.P;
    float* g2p = g1p + p.P;
    cp_async_wait_all();
    __syncthreads();                                   // this unit's record is visible; the other buffer is free
    if (dyn) { if (k > 0) nxt = tickets[(k - 1) & 1]; }
    else nxt = unit + (int)gridDim.x;
    provide(nxt, (k + 1) & 1);
    unsigned int drawn = 0u;
    if (dyn && tid == 0) drawn = atomicAdd(p.dyn_counter, 1u);
    const BwdCoef c = coefs[k & 1];
    if (p.vec_ok) {
      auto emit = [&](int ii, const BwdLoad& d) {
        float4 go0, go1, go2, gz;
        const int row = ii / p.W, col = ii - row * p.W;
        bwd_point(c, d.a0.x, d.a1.x, d.a2.x, d.zz.x, d.mm.x && d.im.x && d.zz.x > 0.0f, row, col + 0, go0.x, go1.x, go2.x, gz.x);
        bwd_point(c, d.a0.y, d.a1.y, d.a2.y, d.zz.y, d.mm.y && d.im.y && d.zz.y > 0.0f, row, col + 1, go0.y, go1.y, go2.y, gz.y);
        bwd_point(c, d.a0.z, d.a1.z, d.a2.z, d.zz.z, d.mm.z && d.im.z && d.zz.z > 0.0f, row, col + 2, go0.z, go1.z, go2.z, gz.z);
        bwd_point(c, d.a0.w, d.a1.w, d.a2.w, d.zz.w, d.mm.w && d.im.w && d.zz.w > 0.0f, row, col + 3, go0.w, go1.w, go2.w, gz.w);
        __stcs(reinterpret_cast<float4*>(g0p + ii), go0);
        __stcs(reinterpret_cast<float4*>(g1p + ii), go1);
        __stcs(reinterpret_cast<float4*>(g2p + ii), go2);
        if (p.grad_depth) __stcs(reinterpret_cast<float4*>(p.grad_depth + ob + ii), gz);
      };
      if (c.live) {
        BwdLoad d0, d1;
        int i = px0 + 4 * tid;
        for (; i + 4 * NT < px1; i += 8 * NT) {          // two iterations in flight
          load(ob, n0p, i, d0);
          load(ob, n0p, i + 4 * NT, d1);
          emit(i, d0);
          emit(i + 4 * NT, d1);
        }
        if (i < px1) {
          load(ob, n0p, i, d0);
          emit(i, d0);
        }
      } else {
        const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int ii = px0 + 4 * tid; ii < px1; ii += 4 * NT) {
          __stcs(reinterpret_cast<float4*>(g0p + ii), zero);
          __stcs(reinterpret_cast<float4*>(g1p + ii), zero);
          __stcs(reinterpret_cast<float4*>(g2p + ii), zero);
          if (p.grad_depth) __stcs(reinterpret_cast<float4*>(p.grad_depth + ob + ii), zero);
        }
      }
    } else {
      const float* n1p = n0p + p.P;
      const float* n2p = n1p + p.P;
      for (int ii = px0 + tid; ii < px1; ii += NT) {
        float o0 = 0, o1 = 0, o2 = 0, oz = 0;
        if (c.live) {
          const float z = p.depth[ob + ii];
          bool w = p.mask[ob + ii] != 0 && z > 0.0f;
          if (p.inlier_mask) w = w && p.inlier_mask[ob + ii] != 0;
          const int row = ii / p.W, col = ii - row * p.W;
          bwd_point(c, n0p[ii], n1p[ii], n2p[ii], z, w, row, col, o0, o1, o2, oz);
        }
        g0p[ii] = o0;
        g1p[ii] = o1;
        g2p[ii] = o2;
        if (p.grad_depth) p.grad_depth[ob + ii] = oz;
      }
    }
    if (dyn && tid == 0) tickets[k & 1] = 2 * (int)gridDim.x + (int)drawn;   // the unit of iteration k + 2
  }
  PF_TRACE_END(3);
  PF_TRACE_BEGIN(7);                                            // (earliest warp done)
}

}  // namespace posefit
