// posefit_common.cuh -- PTX helpers (mbarrier, TMA bulk copy, cp.async), launch parameters, tile loaders/views, block reduction
// Part of libposefit_b200.so: included by posefit_kernels.cu (one translation unit, so every kernel sees the
// same inlined helpers and the build stays a single nvcc call).  See include/posefit.h for the C ABI.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "posefit.h"
#include "posefit_math.h"

// Debug build only (make EXTRA=-DPF_BOUNDS OUT=...): every computed index of the three hot kernels -- shared-memory
// rings, tables, bitmap, select list, queues, per-object global offsets -- is range-checked; a violation traps (the
// launch fails with an error) instead of corrupting memory.  compute-sanitizer is closed on this pool, so this
// build, run by tests/test_gpu_guards.py, is the memory-safety evidence for reads and shared memory (the guard-zone
// test only sees stray global WRITES).
#ifdef PF_BOUNDS
#define PF_CHECK(cond) do { if (!(cond)) __trap(); } while (0)
#else
#define PF_CHECK(cond) do { } while (0)
#endif

namespace posefit {

constexpr int kAccPlain = 17;     // n, sx3, sy3, syx9, sxx
constexpr int kAccRansac = 23;    // n, sx3, sy3, syx9, sxx6 (xx,xy,xz,yy,yz,zz), syy

// ---------------------------------------------------------------------------------------------
// PTX helpers: mbarrier + 1-D TMA bulk copy (SASS: UBLKCP, SYNCS)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a lost copy traps (kernel error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---------------------------------------------------------------------------------------------
// launch parameters
// ---------------------------------------------------------------------------------------------
struct FwdParams {
  const float* noc;
  const float* depth;
  const uint8_t* mask;
  const int32_t* bbox;
  const double* kinv;
  const int32_t* sample_idx;
  const double* src_pts;        // points mode: [B][3][P] float64 source (already centred NOC)
  const double* dst_pts;        // points mode: [B][3][P] float64 target
  double* pose;
  double* ctx;
  int32_t* status;
  int32_t* n_valid;
  uint8_t* inlier_mask;
  int32_t* winner;
  float* out_scale;             // optional float32 copies of the pose, written by the solve kernels
  float* out_rot;
  float* out_trans;
  uint8_t* valid_mask;          // optional (plain fit): 1 where the pixel took part in the fit
  int32_t* redo_flag;           // RANSAC: set by fit_ransac_crop_kernel when fit_ransac_kernel has to do the batch (skewed K)
  double ratio_adapt;
  double pass_override, stop_override;   // > 0: use instead of the data-derived PassT / StopT (getRANSACInliers' arguments)
  int kinv_per_object;
  int B, H, W, P;
  int n_hyp, n_samp, ref_compat;
  int idx_bits;                 // sample_idx holds uniform 32-bit values (device-side draws), mapped to floor(u N / 2^32)
  int no_fast;                  // debugging: force the generic per-pixel passes of the RANSAC kernel
  int no_idx_preload;           // debugging: per-sample index loads in the RANSAC gather loop
  int no_early_issue;           // debugging: request every crop at the top of its own iteration
  int no_screen;                // debugging: SCREEN kernel, but every hypothesis is fitted in double (= v1 arithmetic)
  int global_tile;              // RANSAC kernel: crop too large for shared memory, passes read global memory
  int tile_px, tiles_per_obj;   // a tile = tile_px consecutive pixels (whole rows in crop mode)
  int n_stages, tma_ok;
  int early_dep;                // bit k: kernel k of the chain signals its dependents before its own wait
  int prewarm;                  // K-solve kernels: run a warm-up pass before griddepcontrol.wait (small grids)
  unsigned int* dyn_counter;    // K-moments, DYN instantiation: ticket counter (workspace, zeroed before the launch)
  int dyn_base;                 //   first ticket = number of warps in the grid (warp gw starts with object gw)
  int opc;                      // K-solve kernels: objects per CTA (threads beyond it idle; see solve_object)
  int n_words;                  // ceil(P / 32)
  uint32_t w_magic;             // ceil(2^32 / W): px / W == __umulhi(px, w_magic) for px, W < 65536
  // plain path (K-moments / K-solve)
  double* ws;                   // [B][max_parts][17] partial moments
  long long total_chunks;
  int chunks_per_obj, chunks_per_warp, max_parts, vec_ok;
  uint32_t warp_smem_bytes;     // per-warp shared memory: cp.async ring + ray tables
  // shared-memory carve-up (bytes from the dynamic smem base)
  uint32_t off_ftab;            // RANSAC SCREEN variant: float ray tables (rx[W], ry[H]), 16-byte aligned
  uint32_t off_geom, off_tables, off_red, off_bits, off_prefix, off_stats, off_res, off_tf, off_stages;
  uint32_t stage_bytes, st_depth, st_mask, st_idx;   // offsets inside one stage
};

// One-thread-per-object kernels (K-solve, K-solve-ransac, K-backward-coef): the object of this thread.  Every thread
// reads and writes its own records element by element -- 32 different lines per warp instruction, which one SM's load /
// store unit takes one at a time -- so small batches are spread over all SMs, `opc` objects per CTA (the first opc
// threads; the others idle), instead of filling a few CTAs.  Returns n_objects for an idle thread.
__device__ __forceinline__ int solve_object(int opc, int n_objects) {
  const int o = (int)blockIdx.x * opc + (int)threadIdx.x;
  return ((int)threadIdx.x < opc && o < n_objects) ? o : n_objects;
}

// first object of the calling warp and how many of its lanes hold one (0 .. 32)
__device__ __forceinline__ int solve_warp_rows(int opc, int n_objects, long long& row0) {
  const int w0 = (int)threadIdx.x & ~31;
  row0 = (long long)blockIdx.x * opc + w0;
  long long n = (long long)(opc - w0);
  if (n > n_objects - row0) n = n_objects - row0;
  return n < 0 ? 0 : (n > 32 ? 32 : (int)n);
}

// Per-object geometry (K^-1 and the crop origin) lives in shared memory, double buffered: the
// record of object j+1 is fetched with cp.async (LDGSTS, no register staging) while object j is
// being processed.
struct GeomSmem {
  double k[9];
  int xy0[2];
};

struct ObjGeom {
  const double* k;   // -> shared
  double k0, k2, k4, k5;
  int x0, y0;
  bool simple;
};

__device__ __forceinline__ void cp_async_8(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_4(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// Sample index of a RANSAC hypothesis (pose_utils.py:73) from the caller's tensor: an index into the compacted
// correspondences, clamped into [0, N) -- or, with idx_bits, a uniform 32-bit value u mapped to floor(u N / 2^32)
// (draws made on the device before N is known: include/posefit.h, POSEFIT_SAMPLES_ARE_BITS).
__device__ __forceinline__ int sample_index(int raw, int N, int idx_bits) {
  return idx_bits ? (int)__umulhi((uint32_t)raw, (uint32_t)N) : max(0, min(raw, N - 1));
}

// threads 0..10 request the geometry record of `obj` into `dst`
__device__ __forceinline__ void fetch_geom(const FwdParams& p, int obj, GeomSmem* dst, int tid) {
  if (tid < 9) cp_async_8(&dst->k[tid], p.kinv + (p.kinv_per_object ? 9 * (size_t)obj : 0) + tid);
  else if (tid < 11) cp_async_4(&dst->xy0[tid - 9], p.bbox + 2 * (size_t)obj + (tid - 9));
  cp_async_commit();
}

// after cp_async_wait_all by the fetching threads + __syncthreads
__device__ __forceinline__ void read_geom(const GeomSmem* src, ObjGeom& g) {
  g.k = src->k;
  g.k0 = src->k[0]; g.k2 = src->k[2]; g.k4 = src->k[4]; g.k5 = src->k[5];
  g.x0 = src->xy0[0];
  g.y0 = src->xy0[1];
  g.simple = (src->k[1] == 0.0 && src->k[3] == 0.0 && src->k[6] == 0.0 && src->k[7] == 0.0 && src->k[8] == 1.0);
}

// Camera-space point of frame pixel (x0+col, y0+row) at depth zd, pose_estimation.py:34-41:
// K^-1 [u v 1]^T scaled to depth, y and z negated.  `simple` = pinhole K without skew, where
// the third ray component is exactly 1 and the per-column / per-row ray tables are used.
__device__ __forceinline__ void backproject_px(const ObjGeom& g, const double* rxc, const double* ryr, int row, int col,
                                               double zd, double& y0, double& y1, double& y2) {
  if (g.simple) {
    y0 = rxc[col] * zd;
    y1 = -(ryr[row] * zd);
    y2 = -zd;
  } else {
    const double u = (double)(g.x0 + col), v = (double)(g.y0 + row);
    const double X = g.k[0] * u + g.k[1] * v + g.k[2];
    const double Y = g.k[3] * u + g.k[4] * v + g.k[5];
    const double Z = g.k[6] * u + g.k[7] * v + g.k[8];
    y0 = X * zd / Z;
    y1 = -(Y * zd / Z);
    y2 = -(Z * zd / Z);
  }
}

__device__ __forceinline__ void build_ray_tables(const FwdParams& p, const ObjGeom& g, double* rxc, double* ryr, int tid,
                                                 int nt) {
  for (int i = tid; i < p.W; i += nt) rxc[i] = g.k0 * (double)(g.x0 + i) + g.k2;
  for (int i = tid; i < p.H; i += nt) ryr[i] = g.k4 * (double)(g.y0 + i) + g.k5;
}

// Issue the copies of one tile (pixels [i0, i0+npx) of object obj; idx too when with_idx).
// Called by ONE thread.  Crop-mode stage layout: noc plane c at c*npx floats, depth at st_depth,
// mask at st_mask, sample indices at st_idx.  Points mode: src plane c at c*npx doubles, dst
// planes at st_depth, mask at st_mask.
template <bool POINTS>
__device__ __forceinline__ void issue_tile(const FwdParams& p, unsigned char* stage, uint64_t* bar, int obj, int i0,
                                           int npx_i, bool with_idx) {
  const uint32_t npx = (uint32_t)npx_i;
  const size_t base = (size_t)obj * p.P + (size_t)i0;
  const uint32_t idx_bytes = with_idx ? (uint32_t)p.n_hyp * p.n_samp * 4u : 0u;
  fence_proxy_async();
  mbar_expect_tx(bar, npx * (POINTS ? 49u : 17u) + idx_bytes);
  if (POINTS) {
    if (npx_i == p.P) {
      bulk_g2s(stage, p.src_pts + (size_t)obj * 3 * p.P, npx * 24u, bar);
      bulk_g2s(stage + p.st_depth, p.dst_pts + (size_t)obj * 3 * p.P, npx * 24u, bar);
    } else {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        bulk_g2s(stage + (size_t)c * npx * 8, p.src_pts + ((size_t)obj * 3 + c) * p.P + i0, npx * 8u, bar);
        bulk_g2s(stage + p.st_depth + (size_t)c * npx * 8, p.dst_pts + ((size_t)obj * 3 + c) * p.P + i0, npx * 8u, bar);
      }
    }
  } else {
    if (npx_i == p.P) {
      bulk_g2s(stage, p.noc + (size_t)obj * 3 * p.P, npx * 12u, bar);
    } else {
#pragma unroll
      for (int c = 0; c < 3; ++c)
        bulk_g2s(stage + (size_t)c * npx * 4, p.noc + ((size_t)obj * 3 + c) * p.P + i0, npx * 4u, bar);
    }
    bulk_g2s(stage + p.st_depth, p.depth + base, npx * 4u, bar);
  }
  bulk_g2s(stage + p.st_mask, p.mask + base, npx, bar);
  if (with_idx) bulk_g2s(stage + p.st_idx, p.sample_idx + (size_t)obj * p.n_hyp * p.n_samp, idx_bytes, bar);
}

// Fallback loader for shapes/pointers the bulk copy cannot take (16-byte rules): all threads copy.
template <bool POINTS>
__device__ __forceinline__ void load_tile_generic(const FwdParams& p, unsigned char* stage, int obj, int i0, int npx,
                                                  bool with_idx, int tid, int nt) {
  const size_t base = (size_t)obj * p.P + (size_t)i0;
  uint8_t* smsk = stage + p.st_mask;
  if (POINTS) {
    double* ssrc = reinterpret_cast<double*>(stage);
    double* sdst = reinterpret_cast<double*>(stage + p.st_depth);
    for (int i = tid; i < npx; i += nt) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        ssrc[c * npx + i] = p.src_pts[((size_t)obj * 3 + c) * p.P + i0 + i];
        sdst[c * npx + i] = p.dst_pts[((size_t)obj * 3 + c) * p.P + i0 + i];
      }
      smsk[i] = p.mask[base + i];
    }
  } else {
    float* snoc = reinterpret_cast<float*>(stage);
    float* sdep = reinterpret_cast<float*>(stage + p.st_depth);
    for (int i = tid; i < npx; i += nt) {
#pragma unroll
      for (int c = 0; c < 3; ++c) snoc[c * npx + i] = p.noc[((size_t)obj * 3 + c) * p.P + i0 + i];
      sdep[i] = p.depth[base + i];
      smsk[i] = p.mask[base + i];
    }
  }
  if (with_idx) {
    int32_t* sidx = reinterpret_cast<int32_t*>(stage + p.st_idx);
    const int n = p.n_hyp * p.n_samp;
    for (int i = tid; i < n; i += nt) sidx[i] = p.sample_idx[(size_t)obj * n + i];
  }
}

// Uniform view of the correspondences held in one stage.
//   crop mode  : x = noc - 0.5 (pose_estimation.py:323), y = back-projected depth (:34-41),
//                valid = mask & depth > 0 (:23-25)
//   points mode: x, y given explicitly (the [4,N] arrays of pose_utils.py), valid = mask
template <bool POINTS>
struct TileView {
  const float* noc;
  const float* dep;
  const double* src;
  const double* dst;
  const uint8_t* msk;
  int npx;
  __device__ __forceinline__ TileView(const FwdParams& p, const unsigned char* stage, int npx_) : npx(npx_) {
    noc = reinterpret_cast<const float*>(stage);
    dep = reinterpret_cast<const float*>(stage + p.st_depth);
    src = reinterpret_cast<const double*>(stage);
    dst = reinterpret_cast<const double*>(stage + p.st_depth);
    msk = stage + p.st_mask;
  }
  // Large-crop mode of the RANSAC kernel: the same view straight over the object's arrays in global
  // memory (the crop does not fit in shared memory; the passes re-read it through L2).
  __device__ __forceinline__ TileView(const FwdParams& p, int obj) : npx(p.P) {
    noc = p.noc + (size_t)obj * 3 * p.P;
    dep = p.depth + (size_t)obj * p.P;
    src = p.src_pts + (size_t)obj * 3 * p.P;
    dst = p.dst_pts + (size_t)obj * 3 * p.P;
    msk = p.mask + (size_t)obj * p.P;
  }
  __device__ __forceinline__ bool valid(int i, float& z) const {
    if (POINTS) { z = 1.0f; return msk[i] != 0; }
    z = dep[i];
    return msk[i] != 0 && z > 0.0f;
  }
  __device__ __forceinline__ void xy(int i, float z, const ObjGeom& g, const double* rxc, const double* ryr, int row,
                                     int col, double& x0, double& x1, double& x2, double& y0, double& y1,
                                     double& y2) const {
    if (POINTS) {
      x0 = src[i]; x1 = src[npx + i]; x2 = src[2 * npx + i];
      y0 = dst[i]; y1 = dst[npx + i]; y2 = dst[2 * npx + i];
    } else {
      x0 = (double)noc[i] - 0.5;
      x1 = (double)noc[npx + i] - 0.5;
      x2 = (double)noc[2 * npx + i] - 0.5;
      backproject_px(g, rxc, ryr, row, col, (double)z, y0, y1, y2);
    }
  }
};

// Sum v[0..N) over the block (N <= 24).  red: [nwarps][24] doubles.  Result in out[0..N) (shared),
// valid after the NEXT __syncthreads of the caller.
// Warp stage: instead of a 5-step butterfly per value (5 N shuffles) the two lanes of a pair SPLIT the
// remaining values between them at offsets 16, 8 and 4 (24 -> 12 -> 6 -> 3 values per lane), and only
// the last 3 values go through the two remaining butterfly steps: 27 fp64 shuffles instead of 5 N.
// Afterwards lane L (L % 4 == 0) holds the warp totals of values 12 b4 + 6 b3 + 3 b2 + {0,1,2}.
template <int HALF>
__device__ __forceinline__ void split_step(double (&w)[24], bool up, int offset) {
#pragma unroll
  for (int i = 0; i < HALF; ++i) {
    const double keep = up ? w[HALF + i] : w[i];
    const double send = up ? w[i] : w[HALF + i];
    w[i] = keep + __shfl_xor_sync(0xffffffffu, send, offset);
  }
}

template <int N, int NT>
__device__ __forceinline__ void block_reduce(double (&v)[N], double* red, double* out, int tid) {
  static_assert(N <= 24, "block_reduce: at most 24 values");
  const int lane = tid & 31, warp = tid >> 5;
  double w[24];
#pragma unroll
  for (int i = 0; i < 24; ++i) w[i] = i < N ? v[i] : 0.0;
  split_step<12>(w, (lane & 16) != 0, 16);
  split_step<6>(w, (lane & 8) != 0, 8);
  split_step<3>(w, (lane & 4) != 0, 4);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    w[i] += __shfl_xor_sync(0xffffffffu, w[i], 2);
    w[i] += __shfl_xor_sync(0xffffffffu, w[i], 1);
  }
  if ((lane & 3) == 0) {
    double* dst = red + warp * 24 + ((lane >> 4) & 1) * 12 + ((lane >> 3) & 1) * 6 + ((lane >> 2) & 1) * 3;
    dst[0] = w[0]; dst[1] = w[1]; dst[2] = w[2];
  }
  __syncthreads();
  if (tid < N) {
    double s = 0.0;
#pragma unroll
    for (int wi = 0; wi < NT / 32; ++wi) s += red[wi * 24 + tid];
    out[tid] = s;
  }
}

// Debug build only (make trace): wall-clock span of every kernel of the plain path -- earliest start and latest end over
// its warps (%globaltimer, ns) -- read back through posefit_debug_trace (tools/trace_step.py).
// ids: 0 moments, 1 solve, 2 backward coefficients, 3 backward; 8 + id = the same kernel after its griddepcontrol.wait;
// 12 / 13 = solve kernel: moments merged / solved; 4 / 6 = latest first instruction of moments / backward (slot [2k+1]),
// 5 / 7 = their earliest finished warp (slot [2k]).
#ifdef PF_TRACE
__device__ unsigned long long g_trace[32];
__device__ unsigned long long g_trace_warp[4096];          // moments kernel: when each warp (CTA * 16 + warp) finished
__device__ __forceinline__ unsigned long long pf_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define PF_TRACE_BEGIN(k) do { if ((threadIdx.x & 31) == 0) atomicMin(&g_trace[2 * (k)], pf_now()); } while (0)
#define PF_TRACE_END(k) do { if ((threadIdx.x & 31) == 0) atomicMax(&g_trace[2 * (k) + 1], pf_now()); } while (0)
#define PF_TRACE_BOTH(k) do { PF_TRACE_BEGIN(k); PF_TRACE_END(k); } while (0)   /* earliest and latest warp at a point */
#define PF_TRACE_WARP_END() do { if ((threadIdx.x & 31) == 0) g_trace_warp[(blockIdx.x * 16 + (threadIdx.x >> 5)) & 4095] = pf_now(); } while (0)
#else
#define PF_TRACE_WARP_END() do { } while (0)
#define PF_TRACE_BEGIN(k) do { } while (0)
#define PF_TRACE_END(k) do { } while (0)
#define PF_TRACE_BOTH(k) do { } while (0)
#endif

// ---- records of the one-thread-per-object kernels cross global memory through a per-warp tile ---------------------
// A thread that reads or writes its own object's record element by element touches 32 different lines per warp
// instruction, and an SM's load / store unit takes those one at a time (tools/trace_step.py: with 28 objects per warp
// the solve kernel of BASELINE config 2 spent more time in its loads and stores than in the 3x3 solve).  So the warp
// moves the records of the <= 32 consecutive objects it owns with consecutive addresses, through rows of a shared-memory
// tile (row = object, odd stride); every thread works on its own row.
constexpr int kTileStride = 33;
constexpr int kTileDoubles = 32 * kTileStride;            // per warp

template <typename T, int ROW, int STRIDE>
__device__ __forceinline__ void warp_tile_store(T* g, long long row0, int n_rows, const T* tile) {
  const int lane = threadIdx.x & 31;
  T* dst = g + row0 * ROW;
  for (int i = lane; i < n_rows * ROW; i += 32) dst[i] = tile[(i / ROW) * STRIDE + (i % ROW)];
}

// Write the outputs of the warp's objects row0 .. row0 + n_rows - 1 (include/posefit.h: pose[16], ctx[32], status,
// n_valid, optional float32 copies); lane j holds object row0 + j.  Called by ALL 32 lanes (lanes >= n_rows hold
// anything and write nothing); `tile` = this warp's kTileDoubles of shared memory.  ctx30: RANSAC iteration count.
__device__ __forceinline__ void write_pose(const FwdParams& p, long long row0, int n_rows, double* tile, const Fit& f,
                                           int status, double n_fit, double ratio, double pass_t, double n_valid,
                                           double ctx30 = 0.0) {
  const int lane = threadIdx.x & 31;
  double* row = tile + lane * kTileStride;
  __syncwarp();
  row[0] = f.s;
#pragma unroll
  for (int i = 0; i < 9; ++i) row[1 + i] = f.R[i];
#pragma unroll
  for (int i = 0; i < 3; ++i) row[10 + i] = f.t[i];
  row[13] = (status == PF_OK) ? n_fit : 0.0;
  row[14] = ratio;
  row[15] = pass_t;
  __syncwarp();
  warp_tile_store<double, POSEFIT_POSE_DOUBLES, kTileStride>(p.pose, row0, n_rows, tile);
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 9; ++i) row[i] = f.R[i];
#pragma unroll
  for (int i = 0; i < 6; ++i) { row[9 + i] = f.Linv[i]; row[15 + i] = f.H[i]; }
  row[21] = f.s;
  row[22] = f.var;
  row[23] = (status == PF_OK) ? n_fit : 0.0;
#pragma unroll
  for (int i = 0; i < 3; ++i) { row[24 + i] = f.mux[i]; row[27 + i] = f.muy[i]; }
  row[30] = ctx30;
  row[31] = 0.0;
  __syncwarp();
  warp_tile_store<double, POSEFIT_CTX_DOUBLES, kTileStride>(p.ctx, row0, n_rows, tile);
  if (lane < n_rows) {                                      // one word per object: consecutive as they are
    p.status[row0 + lane] = status;
    p.n_valid[row0 + lane] = (int)n_valid;
    if (p.out_scale != nullptr) p.out_scale[row0 + lane] = (float)f.s;
  }
  if (p.out_rot != nullptr || p.out_trans != nullptr) {
    float* tf = reinterpret_cast<float*>(tile);
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 9; ++i) tf[lane * 13 + i] = (float)f.R[i];
#pragma unroll
    for (int i = 0; i < 3; ++i) tf[lane * 13 + 9 + i] = (float)f.t[i];
    __syncwarp();
    if (p.out_rot != nullptr) warp_tile_store<float, 9, 13>(p.out_rot, row0, n_rows, tf);
    if (p.out_trans != nullptr) warp_tile_store<float, 3, 13>(p.out_trans, row0, n_rows, tf + 9);
  }
  __syncwarp();
}

// The same outputs written by the object's own thread, element by element: for the RANSAC solve kernel, whose CTAs must
// fit beside three crop-kernel CTAs that leave no shared memory for a tile (a tile there cost C3 6 us: the solve CTAs
// no longer warmed up beside their producer).
__device__ __forceinline__ void write_pose_direct(const FwdParams& p, int obj, const Fit& f, int status, double n_fit,
                                                  double ratio, double pass_t, double n_valid, double ctx30) {
  double* po = p.pose + (size_t)obj * POSEFIT_POSE_DOUBLES;
  po[0] = f.s;
#pragma unroll
  for (int i = 0; i < 9; ++i) po[1 + i] = f.R[i];
#pragma unroll
  for (int i = 0; i < 3; ++i) po[10 + i] = f.t[i];
  po[13] = (status == PF_OK) ? n_fit : 0.0;
  po[14] = ratio;
  po[15] = pass_t;
  double* cx = p.ctx + (size_t)obj * POSEFIT_CTX_DOUBLES;
#pragma unroll
  for (int i = 0; i < 9; ++i) cx[i] = f.R[i];
#pragma unroll
  for (int i = 0; i < 6; ++i) { cx[9 + i] = f.Linv[i]; cx[15 + i] = f.H[i]; }
  cx[21] = f.s;
  cx[22] = f.var;
  cx[23] = (status == PF_OK) ? n_fit : 0.0;
#pragma unroll
  for (int i = 0; i < 3; ++i) { cx[24 + i] = f.mux[i]; cx[27 + i] = f.muy[i]; }
  cx[30] = ctx30;
  cx[31] = 0.0;
  p.status[obj] = status;
  p.n_valid[obj] = (int)n_valid;
  if (p.out_scale != nullptr) p.out_scale[obj] = (float)f.s;
  if (p.out_rot != nullptr) {
#pragma unroll
    for (int i = 0; i < 9; ++i) p.out_rot[(size_t)obj * 9 + i] = (float)f.R[i];
  }
  if (p.out_trans != nullptr) {
#pragma unroll
    for (int i = 0; i < 3; ++i) p.out_trans[(size_t)obj * 3 + i] = (float)f.t[i];
  }
}

}  // namespace posefit
