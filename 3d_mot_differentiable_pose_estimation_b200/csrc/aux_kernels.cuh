// aux_kernels.cuh -- kernels behind the drop-ins and the optional pre/post stages (compact, evaluate, transform, epilogue, clip, SOR, resample, gather, tracker edges)
// Part of libposefit_b200.so: included by posefit_kernels.cu (one translation unit, so every kernel sees the
// same inlined helpers and the build stays a single nvcc call).  See include/posefit.h for the C ABI.
#pragma once

#include "posefit_common.cuh"

namespace posefit {

// ---------------------------------------------------------------------------------------------
// Utility kernels behind the same-named Python drop-ins (not on the throughput path)
// ---------------------------------------------------------------------------------------------
struct CompactParams {
  const float* noc;        // may be NULL (backproject only)
  const float* depth;
  const uint8_t* mask;
  const int32_t* bbox;
  const double* kinv;
  double* src;             // [B][P][3] noc - 0.5 (may be NULL)
  double* dst;             // [B][P][3] camera-space points
  int32_t* rows;           // [B][P] frame row of every kept pixel
  int32_t* cols;           // [B][P]
  int32_t* count;          // [B]
  int kinv_per_object, B, H, W, P;
};

// Stable row-major compaction of one crop per CTA: np.where order (pose_estimation.py:27), points
// as backproject builds them (:34-41), NOC gather of run_pose (:323).
__global__ void __launch_bounds__(1024) compact_kernel(const CompactParams p) {
  __shared__ int warp_count[32];
  __shared__ int warp_base[33];
  const int obj = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int per_warp = ((p.P + 31) / 32 + 31) / 32 * 32;       // pixels per warp, multiple of 32
  const int begin = warp * per_warp, end = min(begin + per_warp, p.P);
  const size_t ob = (size_t)obj * p.P;
  int cnt = 0;
  for (int i = begin + lane; i < begin + per_warp; i += 32) {
    const bool v = i < end && p.mask[ob + i] != 0 && p.depth[ob + i] > 0.0f;
    cnt += __popc(__ballot_sync(0xffffffffu, v));
  }
  if (lane == 0) warp_count[warp] = cnt;
  __syncthreads();
  if (warp == 0) {
    const int c = warp_count[lane];
    int incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    warp_base[lane] = incl - c;
    if (lane == 31) { warp_base[32] = incl; p.count[obj] = incl; }
  }
  __syncthreads();
  const double* K = p.kinv + (p.kinv_per_object ? 9 * (size_t)obj : 0);
  const int x0 = p.bbox[2 * obj], y0 = p.bbox[2 * obj + 1];
  int base = warp_base[warp];
  for (int i = begin + lane; i < begin + per_warp; i += 32) {
    float z = 0.0f;
    const bool v = i < end && p.mask[ob + i] != 0 && (z = p.depth[ob + i]) > 0.0f;
    const unsigned b = __ballot_sync(0xffffffffu, v);
    if (v) {
      const int k = base + __popc(b & ((1u << lane) - 1u));
      const int row = i / p.W, col = i - row * p.W;
      const double u = (double)(x0 + col), vv = (double)(y0 + row), zd = (double)z;
      const double X = K[0] * u + K[1] * vv + K[2];
      const double Y = K[3] * u + K[4] * vv + K[5];
      const double Z = K[6] * u + K[7] * vv + K[8];
      double* d = p.dst + (ob + k) * 3;
      d[0] = X * zd / Z;
      d[1] = -(Y * zd / Z);
      d[2] = -(Z * zd / Z);
      if (p.src != nullptr && p.noc != nullptr) {
        double* sp = p.src + (ob + k) * 3;
        sp[0] = (double)p.noc[ob * 3 + i] - 0.5;
        sp[1] = (double)p.noc[ob * 3 + p.P + i] - 0.5;
        sp[2] = (double)p.noc[ob * 3 + 2 * (size_t)p.P + i] - 0.5;
      }
      p.rows[ob + k] = y0 + row;
      p.cols[ob + k] = x0 + col;
    }
    base += __popc(b);
  }
}

// evaluateModel (pose_utils.py:5-14) for one explicit 4x4 transform per object.
// stats[b] = {Residual, n_inliers, point-0-is-inlier, n_points}
__global__ void __launch_bounds__(256) evaluate_kernel(const double* tf, const double* src, const double* dst,
                                                       const uint8_t* mask, const double* pass_t, int pass_per_object,
                                                       int N, double* stats, uint8_t* inlier_mask) {
  __shared__ double red[8 * 3];
  const int obj = blockIdx.x, tid = threadIdx.x;
  const double* T = tf + (size_t)obj * 16;
  const double pt = pass_t[pass_per_object ? obj : 0];
  const size_t ob = (size_t)obj * N;
  double acc[3] = {0.0, 0.0, 0.0};                             // sum r^2, inliers, points
  int first_seen = 0x7fffffff, first_inl = 0;
  for (int i = tid; i < N; i += 256) {
    uint8_t flag = 0;
    if (mask[ob + i] != 0) {
      const double x0 = src[ob * 3 + i], x1 = src[ob * 3 + N + i], x2 = src[ob * 3 + 2 * (size_t)N + i];
      const double e0 = dst[ob * 3 + i] - (T[0] * x0 + T[1] * x1 + T[2] * x2 + T[3]);
      const double e1 = dst[ob * 3 + N + i] - (T[4] * x0 + T[5] * x1 + T[6] * x2 + T[7]);
      const double e2 = dst[ob * 3 + 2 * (size_t)N + i] - (T[8] * x0 + T[9] * x1 + T[10] * x2 + T[11]);
      const double r2 = e0 * e0 + e1 * e1 + e2 * e2;
      acc[0] += r2;
      acc[2] += 1.0;
      if (sqrt(r2) < pt) { acc[1] += 1.0; flag = 1; }          // :8-10
      if (i < first_seen) { first_seen = i; first_inl = flag; }
    }
    inlier_mask[ob + i] = flag;
  }
  // smallest selected index over the block decides the "index 0" quirk (:11)
  const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    double x = acc[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if (lane == 0) red[warp * 3 + k] = x;
  }
  __shared__ int first_idx[8];
  __shared__ int first_val[8];
  {
    // per-warp (index, flag) of the smallest selected index
    int idx = first_seen, val = first_inl;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
      const int ov = __shfl_xor_sync(0xffffffffu, val, o);
      if (oi < idx) { idx = oi; val = ov; }
    }
    if (lane == 0) { first_idx[warp] = idx; first_val[warp] = val; }
  }
  __syncthreads();
  if (tid == 0) {
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    int idx = 0x7fffffff, val = 0;
    for (int w = 0; w < 8; ++w) {
      s0 += red[w * 3]; s1 += red[w * 3 + 1]; s2 += red[w * 3 + 2];
      if (first_idx[w] < idx) { idx = first_idx[w]; val = first_val[w]; }
    }
    double* st = stats + (size_t)obj * 4;
    st[0] = sqrt(s0);                                           // :9
    st[1] = s1;
    st[2] = (double)val;
    st[3] = s2;
  }
}

// out = A * p + t for interleaved [N][3] points; M = [A | t] row-major 3x4 per object.
// transform_pc (pose_estimation.py:45-57) and cam2world (:59-70).
__global__ void __launch_bounds__(256) transform_kernel(const double* M, int m_per_object, const double* pts, double* out,
                                                        long long n_per_object, long long total) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const double* m = M + (m_per_object ? (i / n_per_object) * 12 : 0);
    const double x = pts[3 * i], y = pts[3 * i + 1], z = pts[3 * i + 2];
    out[3 * i] = m[0] * x + m[1] * y + m[2] * z + m[3];
    out[3 * i + 1] = m[4] * x + m[5] * y + m[6] * z + m[7];
    out[3 * i + 2] = m[8] * x + m[9] * y + m[10] * z + m[11];
  }
}

// ---------------------------------------------------------------------------------------------
// K-epilogue: the tail of run_pose (pose_estimation.py:367-412) for a whole batch, on the GPU:
// object->world chaining with the camera pose, scale, XYZ Euler angles of the unscaled rotation
// (postprocess.py:158-160) and the world-space axis-aligned box of the object's depth points in
// the reference's sort_bbox corner order (:72-93, :373-380).  One CTA per object streams depth +
// mask (5 B/px) for the box.
// ---------------------------------------------------------------------------------------------
struct EpiParams {
  const float* depth;
  const uint8_t* mask;
  const int32_t* bbox;
  const double* kinv;
  const double* pose;        // [B][16]
  const int32_t* status;     // [B]
  const double* campose;     // [n][16] row-major 4x4 (NULL = identity: run_pose_office)
  const int32_t* cam_index;  // [B] row of campose per object (NULL = object index, or 0 if one pose)
  double* out;               // [B][40]: global_rot(9, scale embedded) | trans(3) | scale | euler(3) | box(8x3)
  int kinv_per_object, n_campose, B, H, W, P;
};

__global__ void __launch_bounds__(128) pose_epilogue_kernel(const EpiParams p) {
  __shared__ double red[4][6];
  const int obj = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double C[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};           // camera-to-world [R | t]
  if (p.campose != nullptr) {
    const int ci = p.cam_index ? p.cam_index[obj] : (p.n_campose == 1 ? 0 : obj);
#pragma unroll
    for (int i = 0; i < 12; ++i) C[i] = p.campose[(size_t)ci * 16 + i];
  }
  const double* K = p.kinv + (p.kinv_per_object ? 9 * (size_t)obj : 0);
  const int x0 = p.bbox[2 * obj], y0 = p.bbox[2 * obj + 1];
  const size_t ob = (size_t)obj * p.P;
  double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
  for (int i = tid; i < p.P; i += 128) {
    const float z = p.depth[ob + i];
    if (p.mask[ob + i] != 0 && z > 0.0f) {
      const int row = i / p.W, col = i - row * p.W;
      const double u = (double)(x0 + col), v = (double)(y0 + row), zd = (double)z;
      const double X = K[0] * u + K[1] * v + K[2], Y = K[3] * u + K[4] * v + K[5], Z = K[6] * u + K[7] * v + K[8];
      const double c0 = X * zd / Z, c1 = -(Y * zd / Z), c2 = -(Z * zd / Z);     // backproject, :34-41
#pragma unroll
      for (int a = 0; a < 3; ++a) {                                               // cam2world, :59-70
        const double w = C[4 * a] * c0 + C[4 * a + 1] * c1 + C[4 * a + 2] * c2 + C[4 * a + 3];
        lo[a] = fmin(lo[a], w);
        hi[a] = fmax(hi[a], w);
      }
    }
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo[a] = fmin(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
      hi[a] = fmax(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
    }
    if (lane == 0) { red[warp][a] = lo[a]; red[warp][3 + a] = hi[a]; }
  }
  __syncthreads();
  if (tid != 0) return;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    lo[a] = fmin(fmin(red[0][a], red[1][a]), fmin(red[2][a], red[3][a]));
    hi[a] = fmax(fmax(red[0][3 + a], red[1][3 + a]), fmax(red[2][3 + a], red[3][3 + a]));
  }
  const double* po = p.pose + (size_t)obj * POSEFIT_POSE_DOUBLES;
  double* out = p.out + (size_t)obj * 40;
  const double s = po[0];
  // global = campose @ [diag(S) Rotation^T | t] = campose @ [s R | t]   (:401-407)
  double G[9], Ru[9], gt[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      Ru[3 * i + j] = C[4 * i] * po[1 + j] + C[4 * i + 1] * po[4 + j] + C[4 * i + 2] * po[7 + j];
      G[3 * i + j] = s * Ru[3 * i + j];
    }
    gt[i] = C[4 * i] * po[10] + C[4 * i + 1] * po[11] + C[4 * i + 2] * po[12] + C[4 * i + 3];
  }
  // unscaled rotation = global_rot / column norms (get_scale, inference_utils.py:20-23)
  double M[9];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const double nrm = sqrt(G[j] * G[j] + G[3 + j] * G[3 + j] + G[6 + j] * G[6 + j]);
#pragma unroll
    for (int i = 0; i < 3; ++i) M[3 * i + j] = nrm > 0.0 ? G[3 * i + j] / nrm : Ru[3 * i + j];
  }
  // XYZ Euler angles as mathutils.Matrix.to_euler() picks them: two candidates, the one with the
  // smaller |x|+|y|+|z| wins (Blender mat3_normalized_to_eul2); computed here in double
  const double cy = hypot(M[0], M[3]);
  double e1[3], e2[3];
  if (cy > 16.0 * 1.1920929e-07) {
    e1[0] = atan2(M[7], M[8]);   e1[1] = atan2(-M[6], cy);  e1[2] = atan2(M[3], M[0]);
    e2[0] = atan2(-M[7], -M[8]); e2[1] = atan2(-M[6], -cy); e2[2] = atan2(-M[3], -M[0]);
  } else {
    e1[0] = atan2(-M[5], M[4]); e1[1] = atan2(-M[6], cy); e1[2] = 0.0;
    e2[0] = e1[0]; e2[1] = e1[1]; e2[2] = e1[2];
  }
  const bool second = fabs(e1[0]) + fabs(e1[1]) + fabs(e1[2]) > fabs(e2[0]) + fabs(e2[1]) + fabs(e2[2]);
  const bool okp = p.status[obj] == PF_OK;
#pragma unroll
  for (int i = 0; i < 9; ++i) out[i] = G[i];
#pragma unroll
  for (int i = 0; i < 3; ++i) { out[9 + i] = gt[i]; out[13 + i] = second ? e2[i] : e1[i]; }
  out[12] = s;
  // corners in the order sort_bbox (:72-93) gives an axis-aligned box:
  // (H,H,H) (H,H,L) (L,H,L) (L,H,H) (H,L,H) (H,L,L) (L,L,L) (L,L,H)
  const int cx[8] = {1, 1, 0, 0, 1, 1, 0, 0}, cyy[8] = {1, 1, 1, 1, 0, 0, 0, 0}, cz[8] = {1, 0, 0, 1, 1, 0, 0, 1};
  const bool has = okp && hi[0] >= lo[0];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    out[16 + 3 * c] = has ? (cx[c] ? hi[0] : lo[0]) : 0.0;
    out[17 + 3 * c] = has ? (cyy[c] ? hi[1] : lo[1]) : 0.0;
    out[18 + 3 * c] = has ? (cz[c] ? hi[2] : lo[2]) : 0.0;
  }
}

// ---------------------------------------------------------------------------------------------
// K-clip: the GT-box pre-filter of run_pose (clean_depth, pose_estimation.py:107-134, applied at
// :293-299): keep the correspondences whose WORLD-space depth point lies strictly inside the
// axis-aligned extent of the object's 8x3 GT box, but only if more than `min_keep` (20) survive;
// expressed as a new validity mask so the fit kernels need no other change.
// ---------------------------------------------------------------------------------------------
struct ClipParams {
  const float* depth;
  const uint8_t* mask;
  const int32_t* bbox;
  const double* kinv;
  const double* campose;     // [n][16]
  const int32_t* cam_index;  // [B] or NULL
  const double* gt_box;      // [B][8][3]
  uint8_t* out_mask;         // [B][H][W]
  int32_t* kept;             // [B] number of correspondences that survive (optional)
  int kinv_per_object, n_campose, B, H, W, P, min_keep;
};

__global__ void __launch_bounds__(256) clip_mask_kernel(const ClipParams p) {
  __shared__ int warp_cnt[8];
  __shared__ int total;
  const int obj = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ci = p.cam_index ? p.cam_index[obj] : (p.n_campose == 1 ? 0 : obj);
  double C[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) C[i] = p.campose[(size_t)ci * 16 + i];
  const double* gb = p.gt_box + (size_t)obj * 24;
  double lo[3], hi[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    lo[a] = gb[a];
    hi[a] = gb[a];
#pragma unroll
    for (int c = 1; c < 8; ++c) { lo[a] = fmin(lo[a], gb[3 * c + a]); hi[a] = fmax(hi[a], gb[3 * c + a]); }
  }
  const double* K = p.kinv + (p.kinv_per_object ? 9 * (size_t)obj : 0);
  const int x0 = p.bbox[2 * obj], y0 = p.bbox[2 * obj + 1];
  const size_t ob = (size_t)obj * p.P;
  auto inside = [&](int i, bool& valid) {
    const float z = p.depth[ob + i];
    valid = p.mask[ob + i] != 0 && z > 0.0f;
    if (!valid) return false;
    const int row = i / p.W, col = i - row * p.W;
    const double u = (double)(x0 + col), v = (double)(y0 + row), zd = (double)z;
    const double X = K[0] * u + K[1] * v + K[2], Y = K[3] * u + K[4] * v + K[5], Z = K[6] * u + K[7] * v + K[8];
    const double c0 = X * zd / Z, c1 = -(Y * zd / Z), c2 = -(Z * zd / Z);
    bool in = true;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const double w = C[4 * a] * c0 + C[4 * a + 1] * c1 + C[4 * a + 2] * c2 + C[4 * a + 3];
      in = in && (w > lo[a]) && (w < hi[a]);                  // strict, :127-128
    }
    return in;
  };
  int cnt = 0;
  for (int i = tid; i < p.P; i += 256) {
    bool valid;
    cnt += inside(i, valid) ? 1 : 0;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if (lane == 0) warp_cnt[warp] = cnt;
  __syncthreads();
  if (tid == 0) {
    int t = 0;
    for (int w = 0; w < 8; ++w) t += warp_cnt[w];
    total = t;
  }
  __syncthreads();
  const bool use_clip = total > p.min_keep;                    // "if len(new_idxs) > 20", :295
  int kept = 0;
  for (int i = tid; i < p.P; i += 256) {
    bool valid;
    const bool in = inside(i, valid);
    const bool keep = use_clip ? in : valid;
    p.out_mask[ob + i] = keep ? 1 : 0;
    kept += keep ? 1 : 0;
  }
  if (p.kept != nullptr) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) kept += __shfl_xor_sync(0xffffffffu, kept, o);
    if (lane == 0) atomicAdd(&p.kept[obj], kept);
  }
}

// ---------------------------------------------------------------------------------------------
// K-sor: statistical outlier removal as a mask filter -- the two Open3D
// `remove_statistical_outlier(nb_neighbors=20, std_ratio=2)` passes of run_pose
// (pose_estimation.py:311-318 on the depth cloud, :341-349 on the NOC cloud; only when the cloud has
// more than 100 points).  Semantics restated from Open3D's PointCloud::RemoveStatisticalOutliers
// (open3d==0.10.0.0 is not vendored: UNPINNED): avg_i = mean distance to the 20 nearest neighbours
// (the query itself included), threshold = mean(avg) + std_ratio * std(avg, ddof=1), keep
// 0 < avg_i < threshold.  Exact brute-force kNN: one CTA per object, candidates tiled through shared
// memory in fp32 (centred), the 20 selected distances recomputed in fp64.
// ---------------------------------------------------------------------------------------------
struct SorParams {
  const float* noc;
  const float* depth;
  const uint8_t* mask;
  const int32_t* bbox;
  const double* kinv;
  uint8_t* out_mask;
  double* ws_pts;     // [B][P][3] compacted points
  int32_t* ws_px;     // [B][P]    their pixel index
  double* ws_avg;     // [B][P]
  double std_ratio;
  int kinv_per_object, source, min_points, B, H, W, P;
};

constexpr int kSorK = 20;
constexpr int kSorTile = 2048;
constexpr int kSorThreads = 256;

__global__ void __launch_bounds__(kSorThreads) sor_mask_kernel(const SorParams p) {
  __shared__ float tile[kSorTile * 3];
  __shared__ int warp_cnt[kSorThreads / 32];
  __shared__ int warp_base[kSorThreads / 32 + 1];
  __shared__ double red[kSorThreads / 32][4];
  __shared__ double stat[4];
  const int obj = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const size_t ob = (size_t)obj * p.P;
  double* pts = p.ws_pts + ob * 3;
  int32_t* pxs = p.ws_px + ob;
  double* avg = p.ws_avg + ob;
  const double* K = p.kinv + (p.kinv_per_object ? 9 * (size_t)obj : 0);
  const int x0 = p.bbox[2 * obj], y0 = p.bbox[2 * obj + 1];

  // ---- 1. stable compaction of the selected points (same scheme as compact_kernel) -------------
  const int per_warp = ((p.P + kSorThreads / 32 - 1) / (kSorThreads / 32) + 31) / 32 * 32;
  const int begin = warp * per_warp, end = min(begin + per_warp, p.P);
  int cnt = 0;
  for (int i = begin + lane; i < begin + per_warp; i += 32) {
    const bool v = i < end && p.mask[ob + i] != 0 && p.depth[ob + i] > 0.0f;
    cnt += __popc(__ballot_sync(0xffffffffu, v));
  }
  if (lane == 0) warp_cnt[warp] = cnt;
  __syncthreads();
  if (tid == 0) {
    int run = 0;
    for (int w = 0; w < kSorThreads / 32; ++w) { warp_base[w] = run; run += warp_cnt[w]; }
    warp_base[kSorThreads / 32] = run;
  }
  __syncthreads();
  const int N = warp_base[kSorThreads / 32];
  if (N <= p.min_points) {                                      // "if depth_pts.shape[0] > 100", :311 / :341
    for (int i = tid; i < p.P; i += kSorThreads)
      p.out_mask[ob + i] = (p.mask[ob + i] != 0 && p.depth[ob + i] > 0.0f) ? 1 : 0;
    return;
  }
  double csum[3] = {0.0, 0.0, 0.0};
  {
    int base = warp_base[warp];
    for (int i = begin + lane; i < begin + per_warp; i += 32) {
      float z = 0.0f;
      const bool v = i < end && p.mask[ob + i] != 0 && (z = p.depth[ob + i]) > 0.0f;
      const unsigned b = __ballot_sync(0xffffffffu, v);
      if (v) {
        const int k = base + __popc(b & ((1u << lane) - 1u));
        double q[3];
        if (p.source == 0) {
          const int row = i / p.W, col = i - row * p.W;
          const double u = (double)(x0 + col), vv = (double)(y0 + row), zd = (double)z;
          const double X = K[0] * u + K[1] * vv + K[2], Y = K[3] * u + K[4] * vv + K[5], Z = K[6] * u + K[7] * vv + K[8];
          q[0] = X * zd / Z; q[1] = -(Y * zd / Z); q[2] = -(Z * zd / Z);
        } else {
          q[0] = (double)p.noc[ob * 3 + i] - 0.5;
          q[1] = (double)p.noc[ob * 3 + p.P + i] - 0.5;
          q[2] = (double)p.noc[ob * 3 + 2 * (size_t)p.P + i] - 0.5;
        }
        pts[3 * k] = q[0]; pts[3 * k + 1] = q[1]; pts[3 * k + 2] = q[2];
        pxs[k] = i;
        csum[0] += q[0]; csum[1] += q[1]; csum[2] += q[2];
      }
      base += __popc(b);
    }
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) csum[a] += __shfl_xor_sync(0xffffffffu, csum[a], o);
    if (lane == 0) red[warp][a] = csum[a];
  }
  __syncthreads();                                              // also publishes pts / pxs to the block
  if (tid == 0) {
    for (int a = 0; a < 3; ++a) {
      double t = 0.0;
      for (int w = 0; w < kSorThreads / 32; ++w) t += red[w][a];
      stat[a] = t / N;
    }
  }
  __syncthreads();
  const double cen[3] = {stat[0], stat[1], stat[2]};

  // ---- 2. exact 20-NN of every point (queries strided over the block, candidates tiled) ---------
  const int n_rounds = (N + kSorThreads - 1) / kSorThreads;
  double lsum = 0.0, lsq = 0.0;
  for (int r = 0; r < n_rounds; ++r) {
    const int qi = r * kSorThreads + tid;
    const bool live = qi < N;
    float qx = 0.f, qy = 0.f, qz = 0.f;
    if (live) {
      qx = (float)(pts[3 * qi] - cen[0]); qy = (float)(pts[3 * qi + 1] - cen[1]); qz = (float)(pts[3 * qi + 2] - cen[2]);
    }
    float bd[kSorK];
    int bi[kSorK];
#pragma unroll
    for (int s2 = 0; s2 < kSorK; ++s2) { bd[s2] = 3.0e38f; bi[s2] = -1; }
    float dmax = 3.0e38f;
    int imax = 0;
    for (int t0 = 0; t0 < N; t0 += kSorTile) {
      const int tn = min(kSorTile, N - t0);
      __syncthreads();
      for (int j = tid; j < tn; j += kSorThreads) {
        tile[3 * j] = (float)(pts[3 * (t0 + j)] - cen[0]);
        tile[3 * j + 1] = (float)(pts[3 * (t0 + j) + 1] - cen[1]);
        tile[3 * j + 2] = (float)(pts[3 * (t0 + j) + 2] - cen[2]);
      }
      __syncthreads();
      if (live) {
        for (int j = 0; j < tn; ++j) {
          const float dx = tile[3 * j] - qx, dy = tile[3 * j + 1] - qy, dz = tile[3 * j + 2] - qz;
          const float d2 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
          if (d2 < dmax) {
#pragma unroll
            for (int s2 = 0; s2 < kSorK; ++s2)
              if (s2 == imax) { bd[s2] = d2; bi[s2] = t0 + j; }
            dmax = bd[0];
            imax = 0;
#pragma unroll
            for (int s2 = 1; s2 < kSorK; ++s2)
              if (bd[s2] > dmax) { dmax = bd[s2]; imax = s2; }
          }
        }
      }
    }
    if (live) {
      const double ax = pts[3 * qi], ay = pts[3 * qi + 1], az = pts[3 * qi + 2];
      double sum = 0.0;
      int got = 0;
#pragma unroll
      for (int s2 = 0; s2 < kSorK; ++s2)
        if (bi[s2] >= 0) {
          const double dx = pts[3 * bi[s2]] - ax, dy = pts[3 * bi[s2] + 1] - ay, dz = pts[3 * bi[s2] + 2] - az;
          sum += sqrt(dx * dx + dy * dy + dz * dz);
          ++got;
        }
      const double a = got > 0 ? sum / got : -1.0;
      avg[qi] = a;
      if (a > 0.0) lsum += a;
    }
  }
  // ---- 3. threshold = mean + ratio * std (Bessel), over the points with a neighbourhood --------
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
  if (lane == 0) red[warp][0] = lsum;
  __syncthreads();
  if (tid == 0) {
    double t = 0.0;
    for (int w = 0; w < kSorThreads / 32; ++w) t += red[w][0];
    stat[3] = t / N;                                            // every point has >= 1 neighbour (itself)
  }
  __syncthreads();
  const double mean = stat[3];
  for (int qi = tid; qi < N; qi += kSorThreads) {
    const double a = avg[qi];
    if (a > 0.0) lsq += (a - mean) * (a - mean);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) lsq += __shfl_xor_sync(0xffffffffu, lsq, o);
  if (lane == 0) red[warp][1] = lsq;
  __syncthreads();
  if (tid == 0) {
    double t = 0.0;
    for (int w = 0; w < kSorThreads / 32; ++w) t += red[w][1];
    stat[2] = mean + p.std_ratio * sqrt(t / (double)(N - 1));
  }
  __syncthreads();
  const double thr = stat[2];
  for (int i = tid; i < p.P; i += kSorThreads) p.out_mask[ob + i] = 0;
  __syncthreads();
  for (int qi = tid; qi < N; qi += kSorThreads) {
    const double a = avg[qi];
    if (a > 0.0 && a < thr) p.out_mask[ob + pxs[qi]] = 1;
  }
}

// ---------------------------------------------------------------------------------------------
// Batched front end of the per-instance loop of postprocess_dets
// (Detection/tracker/postprocess.py:131-152):
//  * K-resample: the ROI-align resize of the NOC head output (3 x 28 x 28, nocs_head.py:232-235) to
//    each instance's integer box size (postprocess.py:141-147: roi_align(noc[None], [0,0,28,28],
//    output_size=(h_i, w_i), aligned=True); detectron2's roi_align is torchvision.ops.roi_align,
//    sampling_ratio = -1 -> ceil(roi/out) samples per bin), written zero-padded into the common
//    [B,3,H,W] crop layout -- one launch instead of one roi_align call per instance;
//  * K-resample-backward: its adjoint (gradient w.r.t. the head output);
//  * K-gather: depth / mask windows of every instance cut out of the frame tensors
//    (pose_estimation.py:260-262, :290).
// ---------------------------------------------------------------------------------------------
struct ResampleParams {
  const float* head;        // [B][3][Hh][Wh]
  const int32_t* roi_hw;    // [B][2] output size (h_i, w_i) of every instance
  float* crop;              // [B][3][H][W]   (forward: written; backward: gradient, read)
  float* grad_head;         // [B][3][Hh][Wh] (backward only, must be zeroed by the caller)
  int B, Hh, Wh, H, W;
};

// torchvision roi_align bilinear tap (cpu/roi_align_common.h pre_calc_for_bilinear_interpolate),
// float arithmetic with the same operation order; no FMA contraction.
struct BilinearTap {
  int pos1, pos2, pos3, pos4;
  float w1, w2, w3, w4;
};

__device__ __forceinline__ BilinearTap bilinear_tap(float y, float x, int height, int width) {
  BilinearTap t;
  if (y < -1.0f || y > (float)height || x < -1.0f || x > (float)width) {
    t.pos1 = t.pos2 = t.pos3 = t.pos4 = 0;
    t.w1 = t.w2 = t.w3 = t.w4 = 0.0f;
    return t;
  }
  if (y <= 0.0f) y = 0.0f;
  if (x <= 0.0f) x = 0.0f;
  int y_low = (int)y, x_low = (int)x, y_high, x_high;
  if (y_low >= height - 1) { y_high = y_low = height - 1; y = (float)y_low; } else { y_high = y_low + 1; }
  if (x_low >= width - 1) { x_high = x_low = width - 1; x = (float)x_low; } else { x_high = x_low + 1; }
  const float ly = __fsub_rn(y, (float)y_low), lx = __fsub_rn(x, (float)x_low);
  const float hy = __fsub_rn(1.0f, ly), hx = __fsub_rn(1.0f, lx);
  t.w1 = __fmul_rn(hy, hx); t.w2 = __fmul_rn(hy, lx); t.w3 = __fmul_rn(ly, hx); t.w4 = __fmul_rn(ly, lx);
  t.pos1 = y_low * width + x_low;  t.pos2 = y_low * width + x_high;
  t.pos3 = y_high * width + x_low; t.pos4 = y_high * width + x_high;
  return t;
}

// Per-axis tap tables of one object's resize.  The sample coordinate of output row ph, tap iy is
//   yy = (roi_start + ph * bin_h) + ((iy + 0.5) * bin_h) / grid_h          (torchvision roi_align, aligned = True)
// and depends on nothing else, so its bilinear decomposition (pre_calc_for_bilinear_interpolate: clamp, low / high
// cell, weights l and h = 1 - l) is computed ONCE per (row, tap) and per (column, tap) instead of once per pixel,
// channel and tap -- same float operations in the same order, only hoisted.  Entry = { low, high, l, h } (16 bytes);
// rows store low / high already multiplied by the map width.  n entries = size * grid.
struct TapEntry {
  int lo, hi;
  float l, h;
};

__device__ __forceinline__ TapEntry tap_entry(int p, int it, float bin, int grid, int extent, int stride) {
  const float roi_start = -0.5f;                              // ROI = the whole map, aligned: offset 0.5
  float y = __fadd_rn(__fadd_rn(roi_start, __fmul_rn((float)p, bin)), __fdiv_rn(__fmul_rn((float)it + 0.5f, bin), (float)grid));
  TapEntry e;
  if (y < -1.0f || y > (float)extent) {                      // outside the map: the tap contributes nothing
    e.lo = e.hi = 0;
    e.l = e.h = 0.0f;
    return e;
  }
  if (y <= 0.0f) y = 0.0f;
  int lo = (int)y, hi;
  if (lo >= extent - 1) { hi = lo = extent - 1; y = (float)lo; } else { hi = lo + 1; }
  e.l = __fsub_rn(y, (float)lo);
  e.h = __fsub_rn(1.0f, e.l);
  e.lo = lo * stride;
  e.hi = hi * stride;
  return e;
}

// rows[oh * grid_h], cols[ow * grid_w]; all threads of the CTA take part (caller synchronises afterwards)
__device__ __forceinline__ void build_tap_tables(TapEntry* rows, TapEntry* cols, int oh, int ow, float bin_h, float bin_w,
                                                 int grid_h, int grid_w, int Hh, int Wh, int tid, int nt) {
  for (int e = tid; e < oh * grid_h; e += nt) rows[e] = tap_entry(e / grid_h, e - (e / grid_h) * grid_h, bin_h, grid_h, Hh, Wh);
  for (int e = tid; e < ow * grid_w; e += nt) cols[e] = tap_entry(e / grid_w, e - (e / grid_w) * grid_w, bin_w, grid_w, Wh, 1);
}

// One output pixel (ph, pw) of the resize: forward, v[c] = mean over the bin's grid_h x grid_w bilinear taps of channel
// c of `smap`; backward, the adjoint: v[c] holds the pixel's gradient and is scattered into `sgrad` (shared-memory
// atomics).  K-resample and the head-fed fit kernels (fit_head.cuh) share this function, so a NOC value sampled on the
// fly is bit-identical to the one a materialised crop would hold.  (A tap outside the map has l = h = 0: torchvision
// zeroes all four weights when EITHER coordinate is outside, which the products below reproduce.)
template <bool BACKWARD>
__device__ __forceinline__ void sample_head(const float* smap, float* sgrad, int hw, const TapEntry* rows, const TapEntry* cols,
                                            int ph, int pw, int grid_h, int grid_w, float count, float (&v)[3]) {
  float acc[3] = {0.0f, 0.0f, 0.0f}, g[3] = {0.0f, 0.0f, 0.0f};
  if (BACKWARD) {
#pragma unroll
    for (int c = 0; c < 3; ++c) g[c] = count != 1.0f ? v[c] / count : v[c];
  }
#pragma unroll 1
  for (int iy = 0; iy < grid_h; ++iy) {
    const TapEntry ry = rows[ph * grid_h + iy];
#pragma unroll 1
    for (int ix = 0; ix < grid_w; ++ix) {
      const TapEntry cx = cols[pw * grid_w + ix];
      const float w1 = __fmul_rn(ry.h, cx.h), w2 = __fmul_rn(ry.h, cx.l), w3 = __fmul_rn(ry.l, cx.h), w4 = __fmul_rn(ry.l, cx.l);
      const int pos1 = ry.lo + cx.lo, pos2 = ry.lo + cx.hi, pos3 = ry.hi + cx.lo, pos4 = ry.hi + cx.hi;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        if (!BACKWARD) {
          const float* m = smap + c * hw;
          const float val = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(w1, m[pos1]), __fmul_rn(w2, m[pos2])),
                                                __fmul_rn(w3, m[pos3])), __fmul_rn(w4, m[pos4]));
          acc[c] = __fadd_rn(acc[c], val);
        } else {
          float* m = sgrad + c * hw;
          atomicAdd(m + pos1, g[c] * w1);
          atomicAdd(m + pos2, g[c] * w2);
          atomicAdd(m + pos3, g[c] * w3);
          atomicAdd(m + pos4, g[c] * w4);
        }
      }
    }
  }
  if (!BACKWARD) {
#pragma unroll
    for (int c = 0; c < 3; ++c) v[c] = count != 1.0f ? acc[c] / count : acc[c];
  }
}

// The common case -- the box is at least as large as the map, ONE tap per pixel (grid 1 x 1, count 1) -- without the tap
// loops: same operations in the same order as sample_head for that case.  `ry` is the pixel's row entry (shared by the
// pixels of a row, loaded once by the caller).
template <bool BACKWARD>
__device__ __forceinline__ void sample_head_1tap(const float* smap, float* sgrad, int hw, const TapEntry& ry, const TapEntry& cx,
                                                 float (&v)[3]) {
  const float w1 = __fmul_rn(ry.h, cx.h), w2 = __fmul_rn(ry.h, cx.l), w3 = __fmul_rn(ry.l, cx.h), w4 = __fmul_rn(ry.l, cx.l);
  const int pos1 = ry.lo + cx.lo, pos2 = ry.lo + cx.hi, pos3 = ry.hi + cx.lo, pos4 = ry.hi + cx.hi;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    if (!BACKWARD) {
      const float* m = smap + c * hw;
      // (0 + val: sample_head starts its accumulator at zero; adding it is exact)
      v[c] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(w1, m[pos1]), __fmul_rn(w2, m[pos2])), __fmul_rn(w3, m[pos3])),
                       __fmul_rn(w4, m[pos4]));
    } else {
      float* m = sgrad + c * hw;
      atomicAdd(m + pos1, v[c] * w1);
      atomicAdd(m + pos2, v[c] * w2);
      atomicAdd(m + pos3, v[c] * w3);
      atomicAdd(m + pos4, v[c] * w4);
    }
  }
}


// ---- instance masks on the wire: one BIT per pixel ------------------------------------------------------------------
// The detector's instance masks are boolean (Detection/tracker/postprocess.py:134-139 thresholds them); shipping them
// to the device as bits instead of bytes takes 3.6 KB off the 29.9 KB an object costs on PCIe.  bits: little-endian
// within a byte (numpy.packbits(..., bitorder='little')): pixel i = bit (i & 7) of byte (i >> 3); mask: 0 / 1 bytes.
__global__ void __launch_bounds__(256) unpack_mask_kernel(const uint8_t* __restrict__ bits, uint8_t* __restrict__ mask,
                                                          long long n_pixels, int wide) {
  const long long n_bytes = (n_pixels + 7) >> 3;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_bytes; i += stride) {
    const unsigned long long b = (unsigned long long)__ldg(bits + i);
    // byte k of x keeps bit k of b; + 0x7f carries any non-zero byte into its bit 7
    unsigned long long x = (b * 0x0101010101010101ULL) & 0x8040201008040201ULL;
    x = ((x + 0x7f7f7f7f7f7f7f7fULL) >> 7) & 0x0101010101010101ULL;
    if (wide && 8 * i + 8 <= n_pixels) {
      __stcs(reinterpret_cast<unsigned long long*>(mask + 8 * i), x);
    } else {
      for (int k = 0; k < 8 && 8 * i + k < n_pixels; ++k) mask[8 * i + k] = (uint8_t)((x >> (8 * k)) & 1ULL);
    }
  }
}

// bytes of shared memory the two tap tables need for a height x width canvas and an Hh x Wh map
static inline size_t tap_table_bytes(int hh, int wh, int height, int width) {
  const int nr = height > 2 * hh ? height : 2 * hh, nc = width > 2 * wh ? width : 2 * wh;
  return (size_t)(nr + nc) * sizeof(TapEntry);
}

template <bool BACKWARD>
__global__ void __launch_bounds__(256) resample_noc_kernel(const ResampleParams p) {
  extern __shared__ __align__(16) float smap[];               // [3][Hh][Wh] head map (fwd) / gradient (bwd) | tap tables
  const int obj = blockIdx.x, tid = threadIdx.x;
  const int hw = p.Hh * p.Wh;
  TapEntry* rows = reinterpret_cast<TapEntry*>(smap + ((3 * hw + 3) & ~3));
  TapEntry* cols = rows + (p.H > 2 * p.Hh ? p.H : 2 * p.Hh);
  const float* head = p.head + (size_t)obj * 3 * hw;
  if (!BACKWARD) {
    for (int i = tid; i < 3 * hw; i += 256) smap[i] = head[i];
  } else {
    for (int i = tid; i < 3 * hw; i += 256) smap[i] = 0.0f;
  }
  const int oh = p.roi_hw[2 * obj], ow = p.roi_hw[2 * obj + 1];
  const int P = p.H * p.W;
  float* crop = p.crop + (size_t)obj * 3 * P;
  // ROI = the whole map: x1 = y1 = 0, x2 = Wh, y2 = Hh, spatial_scale 1, aligned -> offset 0.5 (tap_entry)
  const float roi_h = (float)p.Hh, roi_w = (float)p.Wh;       // (Hh - 0.5) - (-0.5)
  const float bin_h = oh > 0 ? roi_h / (float)oh : 0.0f, bin_w = ow > 0 ? roi_w / (float)ow : 0.0f;
  const int grid_h = oh > 0 ? (int)ceilf(roi_h / (float)oh) : 1, grid_w = ow > 0 ? (int)ceilf(roi_w / (float)ow) : 1;
  const float count = (float)max(grid_h * grid_w, 1);
  build_tap_tables(rows, cols, min(oh, p.H), min(ow, p.W), bin_h, bin_w, grid_h, grid_w, p.Hh, p.Wh, tid, 256);
  __syncthreads();
  for (int i = tid; i < P; i += 256) {
    const int ph = i / p.W, pw = i - ph * p.W;
    const bool inside = ph < oh && pw < ow;
    float v[3] = {0.0f, 0.0f, 0.0f};
    if (BACKWARD && inside) {
#pragma unroll
      for (int c = 0; c < 3; ++c) v[c] = crop[c * P + i];
    }
    if (inside) sample_head<BACKWARD>(smap, smap, hw, rows, cols, ph, pw, grid_h, grid_w, count, v);
    if (!BACKWARD) {
#pragma unroll
      for (int c = 0; c < 3; ++c) crop[c * P + i] = inside ? v[c] : 0.0f;
    }
  }
  if (BACKWARD) {
    __syncthreads();
    float* gh = p.grad_head + (size_t)obj * 3 * hw;
    for (int i = tid; i < 3 * hw; i += 256) gh[i] = smap[i];
  }
}

struct GatherParams {
  const float* depth_frames;   // [F][FH][FW]
  const uint8_t* mask_frames;  // [B][FH][FW] full-frame instance masks
  const int32_t* frame_of;     // [B] frame index of every instance (NULL: all in frame 0)
  const int32_t* bbox_xyxy;    // [B][4] integer box (x0, y0, x1, y1), exclusive upper corner
  float* depth;                // [B][H][W]
  uint8_t* mask;               // [B][H][W]
  int32_t* bbox_xy0;           // [B][2]
  int32_t* roi_hw;             // [B][2] (h_i, w_i) clipped to (H, W) and to the frame
  int B, FH, FW, H, W;
};

__global__ void __launch_bounds__(256) gather_crops_kernel(const GatherParams p) {
  const int obj = blockIdx.x, tid = threadIdx.x;
  const int f = p.frame_of ? p.frame_of[obj] : 0;
  int x0 = p.bbox_xyxy[4 * obj], y0 = p.bbox_xyxy[4 * obj + 1], x1 = p.bbox_xyxy[4 * obj + 2], y1 = p.bbox_xyxy[4 * obj + 3];
  x0 = max(0, min(x0, p.FW)); x1 = max(x0, min(x1, p.FW));
  y0 = max(0, min(y0, p.FH)); y1 = max(y0, min(y1, p.FH));
  // A box larger than the canvas cannot be represented (cropping depth / mask while the NOC patch is resized to the
  // truncated size would pair the wrong pixels): such an instance is emitted EMPTY (no valid pixel -> status 1, the
  // reference's "no correspondences" answer) instead of silently wrong.  Callers that hold the boxes on the host
  // are told up front (frontend.run_pose_batched raises).
  const bool fits = (y1 - y0) <= p.H && (x1 - x0) <= p.W;
  const int h = fits ? y1 - y0 : 0, w = fits ? x1 - x0 : 0;
  if (tid == 0) {
    p.bbox_xy0[2 * obj] = x0; p.bbox_xy0[2 * obj + 1] = y0;
    p.roi_hw[2 * obj] = h;    p.roi_hw[2 * obj + 1] = w;
  }
  const float* df = p.depth_frames + (size_t)f * p.FH * p.FW;
  const uint8_t* mf = p.mask_frames + (size_t)obj * p.FH * p.FW;
  const int P = p.H * p.W;
  for (int i = tid; i < P; i += 256) {
    const int r = i / p.W, c = i - r * p.W;
    const bool in = r < h && c < w;
    const size_t src = (size_t)(y0 + r) * p.FW + (x0 + c);
    p.depth[(size_t)obj * P + i] = in ? df[src] : 0.0f;              // pose_estimation.py:260-262
    p.mask[(size_t)obj * P + i] = (in && mf[src] != 0) ? 1 : 0;      // :290
  }
}

// ---------------------------------------------------------------------------------------------
// Tracker graph edges straight from the pose tensors:
// GraphDataset.get_edge_data / get_edge_data_office (Tracking/datasets/graph_dataset.py:30-199,
// :232-330).  Nodes of a sequence are its detections in frame order; for every frame t and every
// frame in its window (t+1 .. t+max_frame_dist, < min(max_seq_len, F), :60-65) every pair
// (n in t, m in frame) is a candidate edge, in that nesting order (:67-118).  With per-node ground
// truth ids (what check_pair returns for the node, -1 = None) a pair is kept only when both ends
// are matched (:93-97, :145-146) and its target is id_n == id_m (:141-144).  Edge features (:166-177,
// :187-199): translation difference, Euler-angle difference, log scale ratio, frame distance --
// computed in float64 like the reference's tensors and rounded once to float32.
// Three tiny kernels: per-sequence counting + ranks, a scan over sequences, the pair writer.  The
// output order is exactly the reference's loop order, so edge_index can be compared element-wise.
// ---------------------------------------------------------------------------------------------
struct EdgeParams {
  const double* trans;         // [N][3]
  const double* rot;           // [N][3] XYZ Euler angles
  const double* scale;         // [N][scale_dim]
  const int32_t* frame_start;  // [S*F + 1] node offset of every (sequence, frame)
  const int32_t* node_id;      // [N] ground-truth id of the node, < 0 = unmatched; NULL = keep all
  int S, F, D, max_len, scale_dim;
  // workspace
  int32_t* rank;               // [N] rank of the node among the matched nodes of its frame, -1 = unmatched
  int32_t* mt;                 // [S*F] matched nodes per frame
  int32_t* block_off;          // [S][F*D + 1] exclusive edge offsets of the (t, d) blocks inside the sequence
  long long* seq_off;          // [S + 1] exclusive edge offsets of the sequences
  int32_t* seq_fp;             // [S] false positives (:95-96, :133-136)
  // outputs
  long long max_edges;         // row stride of edge_index
  long long* edge_index;       // [2][max_edges], node indices LOCAL to the sequence
  float* edge_attr;            // [max_edges][7 + scale_dim]
  float* targets;              // [max_edges] (may be NULL)
  int8_t* consecutive;         // [max_edges] (may be NULL)
  int32_t* edge_seq;           // [max_edges] sequence of every edge (may be NULL)
  long long* totals;           // [2] = { number of directed edges, false positives }
};

__global__ void __launch_bounds__(128) edge_count_kernel(const EdgeParams p) {
  const int s = blockIdx.x;
  const int32_t* fs = p.frame_start + (size_t)s * p.F;
  int32_t* mt = p.mt + (size_t)s * p.F;
  for (int f = threadIdx.x; f < p.F; f += blockDim.x) {
    int c = 0;
    for (int n = fs[f]; n < fs[f + 1]; ++n) {
      const bool ok = p.node_id == nullptr || p.node_id[n] >= 0;
      p.rank[n] = ok ? c : -1;
      c += ok ? 1 : 0;
    }
    mt[f] = c;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int32_t* bo = p.block_off + (size_t)s * (p.F * p.D + 1);
    int run = 0, fp = 0;
    for (int t = 0; t + 1 < p.F; ++t) {
      const int n_t = fs[t + 1] - fs[t];
      bool first = true;
      for (int d = 1; d <= p.D; ++d) {
        const int frame = t + d;
        bo[t * p.D + d - 1] = run;
        if (frame >= p.max_len) continue;                       // graph_dataset.py:60-65
        if (first) fp += n_t - mt[t];                           // :95-96 (j == 0)
        first = false;
        run += mt[t] * mt[frame];
        // :133-136 -- last frame pair: unmatched detections of the window frame are counted when the
        // LAST detection of frame t is itself matched (the loop reaches them only then)
        if (t == p.F - 2 && n_t > 0 && p.rank[fs[t + 1] - 1] >= 0) fp += (fs[frame + 1] - fs[frame]) - mt[frame];
      }
    }
    for (int i = (p.F - 1) * p.D; i <= p.F * p.D; ++i) bo[i] = run;
    p.seq_off[s] = run;                                        // turned into an exclusive scan by edge_scan_kernel
    p.seq_fp[s] = fp;
  }
}

__global__ void __launch_bounds__(32) edge_scan_kernel(const EdgeParams p) {
  const int lane = threadIdx.x;
  const int per = (p.S + 31) / 32;
  long long local = 0, fp = 0;
  for (int i = lane * per; i < min(p.S, (lane + 1) * per); ++i) { local += p.seq_off[i]; fp += p.seq_fp[i]; }
  long long incl = local;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const long long t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) fp += __shfl_xor_sync(0xffffffffu, fp, o);
  long long run = incl - local;
  for (int i = lane * per; i < min(p.S, (lane + 1) * per); ++i) {
    const long long c = p.seq_off[i];
    p.seq_off[i] = run;
    run += c;
  }
  if (lane == 31) { p.seq_off[p.S] = incl; p.totals[0] = incl; }
  if (lane == 0) p.totals[1] = fp;
}

__global__ void __launch_bounds__(128) edge_write_kernel(const EdgeParams p) {
  const int s = blockIdx.y;
  const int t = blockIdx.x / p.D, d = blockIdx.x % p.D + 1;
  const int frame = t + d;
  if (frame >= p.max_len) return;
  const int32_t* fs = p.frame_start + (size_t)s * p.F;
  const int n0 = fs[t], n_t = fs[t + 1] - n0, m0 = fs[frame], n_f = fs[frame + 1] - m0;
  const int mtf = p.mt[(size_t)s * p.F + frame];
  const long long base = p.seq_off[s] + p.block_off[(size_t)s * (p.F * p.D + 1) + t * p.D + d - 1];
  const int A = 7 + p.scale_dim;
  const int node0 = fs[0];
  for (int i = threadIdx.x; i < n_t * n_f; i += blockDim.x) {
    const int n = n0 + i / n_f, m = m0 + i % n_f;
    const int rn = p.rank[n], rm = p.rank[m];
    if (rn < 0 || rm < 0) continue;                             // :93-97, :145-146
    const long long e = base + (long long)rn * mtf + rm;
    if (e >= p.max_edges) continue;
    p.edge_index[e] = n - node0;                                // :164
    p.edge_index[p.max_edges + e] = m - node0;
    float* a = p.edge_attr + e * A;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      a[k] = (float)(p.trans[3 * (size_t)m + k] - p.trans[3 * (size_t)n + k]);         // :169-170
      a[3 + k] = (float)(p.rot[3 * (size_t)m + k] - p.rot[3 * (size_t)n + k]);         // :171-172
    }
    for (int k = 0; k < p.scale_dim; ++k)                                              // :166-168
      a[6 + k] = (float)log(p.scale[(size_t)m * p.scale_dim + k] / p.scale[(size_t)n * p.scale_dim + k]);
    a[6 + p.scale_dim] = (float)(frame - t);                                           // :173-175
    if (p.targets) p.targets[e] = (p.node_id != nullptr && p.node_id[n] == p.node_id[m]) ? 1.0f : 0.0f;   // :141-144
    if (p.consecutive) p.consecutive[e] = (frame == t + 1) ? 1 : 0;                    // :149-162
    if (p.edge_seq) p.edge_seq[e] = s;
  }
}

}  // namespace posefit
