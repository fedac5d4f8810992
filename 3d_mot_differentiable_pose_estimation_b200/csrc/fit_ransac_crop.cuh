// fit_ransac_crop.cuh -- K-ransac-crop: getRANSACInliers / estimateSimilarityTransform (pose_utils.py:63-117) for the
// common case -- fp32 crops, pinhole intrinsics shared by the batch, W % 4 == 0, the crop resident in shared memory.
// Part of libposefit_b200.so: included by posefit_kernels.cu.  fit_ransac_kernel (fit_ransac.cuh) stays the general
// kernel (points mode, per-object / skewed intrinsics, ragged shapes, crops beyond shared memory, > 1024 hypotheses).
//
// Why a second kernel.  ncu of fit_ransac_kernel on BASELINE config 3 (profiles/r01_n_*, r02_a_*): 34-36 k
// warp-instructions per object, 51 % issue-active with three 4-warp CTAs per SM, 40 % of all warp time inside the
// one-hypothesis-per-thread double fits and the barriers around them, a code footprint of 90-170 KB walked once per
// object.  This kernel keeps the data flow (crop in shared memory by 1-D TMA bulk copies, bitmap + select list,
// closed-form residuals, winner-only inlier pass) and changes what that profile blamed:
//   * hypotheses are SCREENED in float: rolled gather loop, Markley's closed-form rotation (posefit_math.h:
//     screen_fit32), residual of the float transform in double, an interval that contains the double fit's
//     residual.  Only hypotheses whose interval reaches below the smallest upper end (or below StopT) can win;
//   * normally that leaves ONE candidate.  It is then the winner whatever its exact residual is, and the inlier pass
//     runs on its FLOAT transform at once; pixels whose residual lies within the guard band of PassT go to a small
//     queue.  Only if that queue is not empty (a few % of objects) is the candidate fitted in double -- by its whole
//     warp: lane j gathers sample j, the moments are folded across the lanes, one lane polishes the float rotation
//     with the double Newton steps -- and the queued pixels decided exactly as the reference decides them.  With two
//     or more candidates the double fits run before the pass (same routine) and pick the winner (pose_utils.py:76-81);
//   * 18 double sums per pixel instead of 23 (residual_sq_iso), the outliers of the inlier pass -- whose moments are
//     accumulated in double -- are queued per warp and drained densely instead of a divergent loop per pixel group;
//   * centring, thresholds and bitmap prefix run in different warps at the same time; loops are rolled: the code a
//     warp walks through per object is ~1/3 of fit_ransac_kernel's.
// Results: winner, inlier mask and inlier moments are those of the all-double evaluation whenever the intervals hold
// (tests/test_math_host.py checks them on eight data regimes; tests/test_gpu_parity.py and tools/fuzz_parity.py
// compare this kernel with fit_ransac_kernel and with the oracle bit for bit on the masks).
#pragma once

#include "fit_ransac.cuh"

namespace posefit {

constexpr int kCropAcc = 20;        // n, sum a (3), sum (y0, y1, z) (3), sum (y0,y1,z) a^T (9), sum |a|^2, sum |y|^2, sum |x|, sum |y|
constexpr int kCropThreads = 128;   // default CTA size (POSEFIT_RANSAC_THREADS: 128 / 160 / 192 / 256)
constexpr int kBandCap = 256;       // guard-band queue entries per object (overflow: the slow path decides in place)

struct CropShared {                 // lives at off_stats
  double n, mux[3], muy[3], Syy, Syx[9], trSxx, x_rms;   // centred global statistics (closed-form residual)
  double pass_t, pass2, stop2;
  double wtf[12];                   // winner's scoring transform in double: A (9), t (3)
  double hi_min[8];                 // per-warp minimum of the residual intervals' upper ends
  float w32[12];                    // winner's float transform (single-candidate path)
  float start[8][10];               // per-warp start rotation of the double polish (+ flag)
  float pass2_f, band_rel;
  int n_valid, first_px, winner, first_is_inlier;
  int n_band;                       // entries in the guard-band queue
  int stopped;                      // the winner ended the search early (pose_utils.py:80-81); 2 = its interval straddles StopT
  uint16_t spx[8][32];              // per warp: pixel of every sample of the hypothesis being fitted in double
};

// ---- pass 1: validity bitmap + 18 raw sums (a = noc, z instead of y2 = -z, see LaneSums) + the two norm sums ---------
__device__ __forceinline__ float sqrt_approx(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

template <int NT>
__device__ __forceinline__ void crop_pass1(const FwdParams& p, const unsigned char* stage, const double* rxc,
                                           const double* ryr, const float* cwf, const float* rwf, uint32_t* bits, int tid,
                                           double (&raw)[kCropAcc]) {
  const int P = p.P, lane = tid & 31;
  const float* snoc = reinterpret_cast<const float*>(stage);
  const float* sdep = reinterpret_cast<const float*>(stage + p.st_depth);
  const unsigned char* smsk = stage + p.st_mask;
  double sa[3] = {0, 0, 0}, sy[3] = {0, 0, 0}, sya[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, saa = 0.0, syy = 0.0;
  float sum_nx = 0.0f, sum_ny = 0.0f;
  int cnt = 0;
  const int n_iter = (P + 4 * NT - 1) / (4 * NT);
  const int drow = (4 * NT) / p.W, dcol = (4 * NT) - drow * p.W;
  int nrow = (4 * tid) / p.W, ncol = 4 * tid - nrow * p.W;
#pragma unroll 1
  for (int k = 0; k < n_iter; ++k) {
    const int i4 = (k * NT + tid) * 4;
    const bool act = i4 < P;
    uint32_t m4 = 0u;
    float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f), a4 = z4, b4 = z4, c4 = z4;
    int row = 0, col = 0;
    if (act) {
      m4 = *reinterpret_cast<const uint32_t*>(smsk + i4);
      z4 = *reinterpret_cast<const float4*>(sdep + i4);
      a4 = *reinterpret_cast<const float4*>(snoc + i4);
      b4 = *reinterpret_cast<const float4*>(snoc + P + i4);
      c4 = *reinterpret_cast<const float4*>(snoc + 2 * P + i4);
      row = nrow;
      col = ncol;
    }
    nrow += drow;
    ncol += dcol;
    if (ncol >= p.W) { ncol -= p.W; ++nrow; }
    const float zz[4] = {z4.x, z4.y, z4.z, z4.w};
    const float n0[4] = {a4.x, a4.y, a4.z, a4.w}, n1[4] = {b4.x, b4.y, b4.z, b4.w}, n2[4] = {c4.x, c4.y, c4.z, c4.w};
    uint32_t nib = 0;
    bool ok[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      ok[j] = (m4 & (0xffu << (8 * j))) != 0u && zz[j] > 0.0f;  // pose_estimation.py:23-25
      nib |= (ok[j] ? 1u : 0u) << j;
    }
    // word (i4 / 32) of the bitmap = nibbles of 8 consecutive lanes
    uint32_t v = nib << (4 * (lane & 7));
    v |= __shfl_xor_sync(0xffffffffu, v, 1);
    v |= __shfl_xor_sync(0xffffffffu, v, 2);
    v |= __shfl_xor_sync(0xffffffffu, v, 4);
    PF_CHECK(!act || (i4 >> 5) < p.n_words);
    PF_CHECK(row >= 0 && row < p.H && col >= 0 && col + 3 < p.W);
    if ((lane & 7) == 0 && act) bits[i4 >> 5] = v;
    cnt += __popc(nib);
    const double nry = -ryr[row];
    const double2 rxa = *reinterpret_cast<const double2*>(rxc + col);
    const double2 rxb = *reinterpret_cast<const double2*>(rxc + col + 2);
    const double rx[4] = {rxa.x, rxa.y, rxb.x, rxb.y};
    const float4 cw4 = *reinterpret_cast<const float4*>(cwf + col);   // 1 + rx^2 per column, ry^2 per row
    const float cw[4] = {cw4.x, cw4.y, cw4.z, cw4.w};
    const float rwv = rwf[row];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t okm = ok[j] ? 0xffffffffu : 0u;
      const float zf = and_bits(zz[j], okm), f0 = and_bits(n0[j], okm), f1 = and_bits(n1[j], okm), f2 = and_bits(n2[j], okm);
      const double zd = (double)zf, a0 = (double)f0, a1 = (double)f1, a2 = (double)f2;
      const double y0 = rx[j] * zd, y1 = nry * zd;            // y = (rx z, -ry z, -z), :34-41
      sa[0] += a0; sa[1] += a1; sa[2] += a2;
      sy[0] += y0; sy[1] += y1; sy[2] += zd;
      sya[0] = fma(y0, a0, sya[0]); sya[1] = fma(y0, a1, sya[1]); sya[2] = fma(y0, a2, sya[2]);
      sya[3] = fma(y1, a0, sya[3]); sya[4] = fma(y1, a1, sya[4]); sya[5] = fma(y1, a2, sya[5]);
      sya[6] = fma(zd, a0, sya[6]); sya[7] = fma(zd, a1, sya[7]); sya[8] = fma(zd, a2, sya[8]);
      saa = fma(a0, a0, fma(a1, a1, fma(a2, a2, saa)));
      const double yy = fma(y0, y0, fma(y1, y1, zd * zd));
      syy += yy;
      // mean norms for PassT (pose_utils.py:91-92), float: |y| = z sqrt(1 + rx^2 + ry^2) from the float ray tables,
      // |x| from the float NOC; one MUFU.SQRT each (1 ulp; the two means enter PassT as a ratio)
      const float x0f = f0 - 0.5f, x1f = f1 - 0.5f, x2f = f2 - 0.5f;
      sum_ny = fmaf(zf, sqrt_approx(cw[j] + rwv), sum_ny);          // invalid pixels: zf == 0
      sum_nx += and_bits(sqrt_approx(fmaf(x0f, x0f, fmaf(x1f, x1f, x2f * x2f))), okm);
    }
  }
  raw[0] = (double)cnt;
#pragma unroll
  for (int i = 0; i < 3; ++i) { raw[1 + i] = sa[i]; raw[4 + i] = sy[i]; }
#pragma unroll
  for (int i = 0; i < 9; ++i) raw[7 + i] = sya[i];
  raw[16] = saa;
  raw[17] = syy;
  raw[18] = (double)sum_nx;
  raw[19] = (double)sum_ny;
}

// lane j: pixel of sample j of the hypothesis whose indices start at gidx_h (pose_utils.py:73), into spx[j]
__device__ __forceinline__ void crop_sample_pixels(const uint16_t* klist, const uint32_t* bits, const int32_t* gidx_h,
                                                   int n_samp, int N, int idx_bits, uint16_t* spx) {
  const int lane = threadIdx.x & 31;
  if (lane < n_samp) {
    const int k = sample_index(__ldg(gidx_h + lane), N, idx_bits);
    PF_CHECK(k >= 0 && k < N);
    const int px = select_px_list(klist, bits, k);
    PF_CHECK(px >= 0 && px < 65536);
    spx[lane] = (uint16_t)px;
  }
  __syncwarp();
}

// ---- the double fit of ONE hypothesis by a whole warp ----------------------------------------------------------------
// lane j gathers sample j (n_samp <= 32; spx[j] = its pixel, resolved through the select list by crop_sample_pixels
// BEFORE the inlier pass reuses that list's memory), the shifted moments are folded across the lanes, lane 0 polishes the start
// rotation (sh->start[warp], flag in [9]) with the double Newton steps of solve_rotation, or runs the full solve when
// there is none.  Returns the closed-form residual^2 (NaN: not a usable fit) in every lane; A, t go to out12 (shared).
// Kept out of line: two call sites (candidates before the inlier pass, guard-band decisions after it), both rare paths.
__device__ __noinline__ double crop_fit_hypothesis(const float* snoc, const float* sdep, const double* rxc,
                                                   const double* ryr, const uint16_t* spx, const CropShared* sh,
                                                   double* scratch, double* out12, int n_samp, int P, int W,
                                                   uint32_t w_magic, int ref_compat) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  auto point = [&](int j, double (&xs)[3], double (&ys)[3]) {
    const int px = (int)spx[j];
    PF_CHECK(px < P);
    const int row = (int)__umulhi((uint32_t)px, w_magic), col = px - row * W;
    const double zd = (double)sdep[px];
    xs[0] = (double)snoc[px] - 0.5; xs[1] = (double)snoc[P + px] - 0.5; xs[2] = (double)snoc[2 * P + px] - 0.5;   // :323
    ys[0] = rxc[col] * zd; ys[1] = -(ryr[row] * zd); ys[2] = -zd;                                                 // :34-41
  };
  double ox[3], oy[3];
  point(0, ox, oy);                                                   // the origin: every lane reads sample 0 (broadcast)
  double xs[3] = {0, 0, 0}, ys[3] = {0, 0, 0};
  if (lane < n_samp) {
    point(lane, xs, ys);
#pragma unroll
    for (int i = 0; i < 3; ++i) { xs[i] -= ox[i]; ys[i] -= oy[i]; }
  }
  // The moments are accumulated in the reference's sample order by EVERY lane (broadcast shuffles), exactly the
  // sequence of operations fit_ransac_kernel's one-thread-per-hypothesis loop runs: bit-identical sums, so a hypothesis
  // fitted without a start rotation gets bit-identical results in the two kernels.
  Moments mo;
  mo.n = (double)n_samp;
#pragma unroll
  for (int i = 0; i < 3; ++i) { mo.sx[i] = 0.0; mo.sy[i] = 0.0; }
#pragma unroll
  for (int i = 0; i < 9; ++i) mo.syx[i] = 0.0;
  mo.sxx = 0.0;
#pragma unroll 1
  for (int j = 0; j < n_samp; ++j) {
    double x[3], y[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) { x[i] = __shfl_sync(0xffffffffu, xs[i], j); y[i] = __shfl_sync(0xffffffffu, ys[i], j); }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      mo.sx[i] += x[i];
      mo.sy[i] += y[i];
      mo.sxx = fma(x[i], x[i], mo.sxx);
#pragma unroll
      for (int jj = 0; jj < 3; ++jj) mo.syx[3 * i + jj] = fma(y[i], x[jj], mo.syx[3 * i + jj]);
    }
  }
  (void)scratch;
  double r2 = 0.0;
  if (lane == 0) {
    float st[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) st[i] = sh->start[warp][i];
    Fit f;
    fit_from_moments<false>(mo, f, ox, oy, st, sh->start[warp][9] != 0.0f);   // pose_utils.py:74
    double A[9];
    scoring_transform(f, ref_compat != 0, A);                               // :57-59 (F3)
    r2 = residual_sq_iso(sh->n, sh->mux, sh->muy, sh->Syy, sh->Syx, sh->trSxx, A, f.t, f.s);   // :7-9 in closed form
    if (f.status != PF_OK) r2 = __longlong_as_double(0x7ff8000000000000LL);
#pragma unroll
    for (int i = 0; i < 9; ++i) out12[i] = A[i];
#pragma unroll
    for (int i = 0; i < 3; ++i) out12[9 + i] = f.t[i];
  }
  __syncwarp();
  return __shfl_sync(0xffffffffu, r2, 0);
}

// ---- winner's inlier pass ----------------------------------------------------------------------------------------
// Float screen of r^2 against PassT^2 with the constant source shift folded into the translation and float ray
// tables.  EXACT = true: the transform is the double fit's and a pixel inside the guard band is decided in double on the
// spot.  EXACT = false: the transform is the screen's float fit; a pixel inside the (wider, sh->band_rel) band is
// neither inlier nor outlier yet -- its index goes to the band queue and the caller decides it after the double fit.
// Outliers (their moments are accumulated in double) go to per-warp queues of 4-pixel groups, drained densely.
// out_raw[18] = { n_out, sum a(3), sum(y0,y1,z)(3), sum (y0,y1,z) a^T (9), sum |a|^2, n_inliers } (thread partials).
template <int NT, bool EXACT>
__device__ __forceinline__ void crop_pass2(const FwdParams& p, const unsigned char* stage, const double* rxc,
                                           const double* ryr, const float* rxf, const float* ryf, const uint32_t* bits,
                                           CropShared* sh, int win, uint8_t* om, uint16_t* queue, uint16_t* band_q,
                                           int tid, LaneSums& acc, int& n_inl) {
  const int P = p.P, lane = tid & 31, warp = tid >> 5;
  const float* snoc = reinterpret_cast<const float*>(stage);
  const float* sdep = reinterpret_cast<const float*>(stage + p.st_depth);
  float Af[9], tf[3];
  if (EXACT) {
#pragma unroll
    for (int i = 0; i < 9; ++i) Af[i] = win >= 0 ? (float)sh->wtf[i] : 0.0f;
#pragma unroll
    for (int i = 0; i < 3; ++i)                                     // x = noc - 1/2 folded into the translation
      tf[i] = win >= 0 ? (float)(sh->wtf[9 + i] - 0.5 * (sh->wtf[3 * i] + sh->wtf[3 * i + 1] + sh->wtf[3 * i + 2])) : 0.0f;
  } else {
#pragma unroll
    for (int i = 0; i < 9; ++i) Af[i] = sh->w32[i];
#pragma unroll
    for (int i = 0; i < 3; ++i) tf[i] = sh->w32[9 + i] - 0.5f * (Af[3 * i] + Af[3 * i + 1] + Af[3 * i + 2]);
  }
  const float pass2_f = sh->pass2_f;
  const float band_w = (EXACT ? 2e-3f : sh->band_rel) * pass2_f;
  const int first_px = sh->first_px;
  const int qcap = ((P / 2 - kBandCap) / (NT / 32)) & ~31;      // entries per warp: the mask plane minus the band queue
  uint16_t* q = queue + warp * qcap;
  int qn = 0;                                                   // warp-uniform fill
  const uint32_t lt = (1u << lane) - 1u;
  const int drow = (4 * NT) / p.W, dcol = (4 * NT) - drow * p.W;
  int nrow = (4 * tid) / p.W, ncol = 4 * tid - nrow * p.W;
  const int n_iter = (P + 4 * NT - 1) / (4 * NT);
#pragma unroll 1
  for (int it = 0; it <= n_iter; ++it) {
    if (it == n_iter || qn + 32 > qcap) {
      // drain: one queued group per lane and round; entry = group index (12 bits) | outlier bits << 12
      __syncwarp();
      for (int e0 = 0; e0 < qn; e0 += 32) {
        const int e = e0 + lane;
        uint32_t ent = e < qn ? (uint32_t)q[e] : 0u;
        const int px0 = (int)(ent & 0xfffu) * 4;
        uint32_t todo = ent >> 12;
        PF_CHECK(px0 + 3 < P || todo == 0u);
        const int row = (int)__umulhi((uint32_t)px0, p.w_magic), col = px0 - row * p.W;
        while (__any_sync(0xffffffffu, todo != 0u)) {
          if (todo != 0u) {
            const int j = __ffs(todo) - 1;
            todo &= todo - 1u;
            const int px = px0 + j;
            const double zd = (double)sdep[px];
            acc.add((double)snoc[px], (double)snoc[P + px], (double)snoc[2 * P + px], rxc[col + j] * zd, -(ryr[row] * zd), zd);
            ++acc.cnt;
          }
        }
      }
      qn = 0;
      __syncwarp();
      if (it == n_iter) break;
    }
    const int i4 = (it * NT + tid) * 4;
    const bool act = i4 < P;
    uint32_t okb = 0u;
    float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f), a4 = z4, b4 = z4, c4 = z4, r4 = z4;
    float ryv = 0.f;
    if (act) {
      okb = (bits[i4 >> 5] >> (i4 & 31)) & 15u;                  // 4 validity bits of this group
      z4 = *reinterpret_cast<const float4*>(sdep + i4);
      a4 = *reinterpret_cast<const float4*>(snoc + i4);
      b4 = *reinterpret_cast<const float4*>(snoc + P + i4);
      c4 = *reinterpret_cast<const float4*>(snoc + 2 * P + i4);
      r4 = *reinterpret_cast<const float4*>(rxf + ncol);
      ryv = ryf[nrow];
    }
    const int row = nrow, col = ncol;
    nrow += drow;
    ncol += dcol;
    if (ncol >= p.W) { ncol -= p.W; ++nrow; }
    const float zz[4] = {z4.x, z4.y, z4.z, z4.w};
    const float n0[4] = {a4.x, a4.y, a4.z, a4.w}, n1[4] = {b4.x, b4.y, b4.z, b4.w}, n2[4] = {c4.x, c4.y, c4.z, c4.w};
    const float rxv[4] = {r4.x, r4.y, r4.z, r4.w};
    uint32_t inb = okb, band = 0u;
    if (win >= 0) {
      inb = 0u;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float z = zz[j];
        const float d0 = rxv[j] * z - (Af[0] * n0[j] + Af[1] * n1[j] + Af[2] * n2[j] + tf[0]);
        const float d1 = -(ryv * z) - (Af[3] * n0[j] + Af[4] * n1[j] + Af[5] * n2[j] + tf[1]);
        const float d2 = -z - (Af[6] * n0[j] + Af[7] * n1[j] + Af[8] * n2[j] + tf[2]);
        const float r2f = d0 * d0 + d1 * d1 + d2 * d2;
        inb |= (r2f < pass2_f ? 1u : 0u) << j;
        band |= (!(fabsf(r2f - pass2_f) > band_w) ? 1u : 0u) << j;
      }
      inb &= okb;
      band &= okb;
      if (band != 0u) {                                          // guard band (rare)
        if (EXACT) {                                             // decide in double, pose_utils.py:7-10
          const double ryd = ryr[row], pass2 = sh->pass2;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if ((band >> j) & 1u) {
              const double xd0 = (double)n0[j] - 0.5, xd1 = (double)n1[j] - 0.5, xd2 = (double)n2[j] - 0.5, zd = (double)zz[j];
              const double e0 = rxc[col + j] * zd - (sh->wtf[0] * xd0 + sh->wtf[1] * xd1 + sh->wtf[2] * xd2 + sh->wtf[9]);
              const double e1 = -(ryd * zd) - (sh->wtf[3] * xd0 + sh->wtf[4] * xd1 + sh->wtf[5] * xd2 + sh->wtf[10]);
              const double e2 = -zd - (sh->wtf[6] * xd0 + sh->wtf[7] * xd1 + sh->wtf[8] * xd2 + sh->wtf[11]);
              const bool in64 = (e0 * e0 + e1 * e1 + e2 * e2) < pass2;
              inb = (inb & ~(1u << j)) | ((in64 ? 1u : 0u) << j);
            }
          }
          band = 0u;
        } else {                                                 // undecided for now: neither inlier nor outlier
          inb &= ~band;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if ((band >> j) & 1u) {
              const int slot = atomicAdd(&sh->n_band, 1);
              if (slot < kBandCap) band_q[slot] = (uint16_t)(i4 + j);
            }
          }
        }
      }
    }
    n_inl += __popc(inb);
    const uint32_t fo = (uint32_t)(first_px - i4);
    if (fo < 4u && ((inb >> fo) & 1u) != 0u) sh->first_is_inlier = 1;
    // spread the 4 bits into 4 bytes: bit j -> byte j
    if (act) *reinterpret_cast<uint32_t*>(om + i4) = (inb * 0x00204081u) & 0x01010101u;
    // groups with valid pixels that are NOT inliers go to the warp's queue
    const uint32_t pending = okb & ~inb & ~band;
    const uint32_t b = __ballot_sync(0xffffffffu, pending != 0u);
    PF_CHECK(qn + 32 <= qcap && (i4 >> 2) < 4096);
    if (pending != 0u) q[qn + __popc(b & lt)] = (uint16_t)((uint32_t)(i4 >> 2) | (pending << 12));
    qn += __popc(b);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
template <int NT>
__global__ void __launch_bounds__(NT, 3) fit_ransac_crop_kernel(const FwdParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  GeomSmem* geo = reinterpret_cast<GeomSmem*>(smem + p.off_geom);       // [2]
  double* rxc = reinterpret_cast<double*>(smem + p.off_tables);
  double* ryr = rxc + p.W;
  float* rxf = reinterpret_cast<float*>(smem + p.off_ftab);
  float* ryf = rxf + p.W;
  float* cwf = ryf + ((p.H + 3) & ~3);                                 // 1 + rx^2 per column (16-byte aligned), ry^2 per row
  float* rwf = cwf + p.W;
  double* red = reinterpret_cast<double*>(smem + p.off_red);           // [NT/32][24] | mom[24] | tot[24] | cur | kept
  double* mom = red + (NT / 32) * 24;
  double* tot = mom + 24;                                              // raw totals of pass 1, kept for the record
  double* cur = tot + 24;                                              // [NT/32][12] transform of a warp's current candidate
  double* kept = cur + (NT / 32) * 12;                                 // [NT/32][12] transform of a warp's leading candidate
  uint32_t* bits = reinterpret_cast<uint32_t*>(smem + p.off_bits);
  uint32_t* prefix = reinterpret_cast<uint32_t*>(smem + p.off_prefix);
  CropShared* sh = reinterpret_cast<CropShared*>(smem + p.off_stats);
  // [n_hyp] lower ends of the residual intervals, rounded DOWN to float; they take the bitmap prefix's place (dead once
  // the select list is built): with them, three CTAs of up to 192 threads still fit the SM's shared memory
  float* slo = reinterpret_cast<float*>(smem + p.off_prefix);
  unsigned char* stage = smem + p.off_stages;
  uint16_t* klist = reinterpret_cast<uint16_t*>(stage + p.st_mask);    // select list, later the outlier queues
  uint16_t* band_q = klist + (p.P / 2 - kBandCap);                     // last kBandCap entries of the mask plane

#if __CUDA_ARCH__ >= 900
  asm volatile("griddepcontrol.launch_dependents;");
#endif
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int G = gridDim.x;
  const int n_obj = (p.B - (int)blockIdx.x + G - 1) / G;
  const int P = p.P;
  const float* snoc = reinterpret_cast<const float*>(stage);
  const float* sdep = reinterpret_cast<const float*>(stage + p.st_depth);

  if (tid == 0) {
    mbar_init(&full[0], 1);
    fence_mbar_init();
  }
  __syncthreads();

  ObjGeom g = {};
  if (n_obj > 0) fetch_geom(p, (int)blockIdx.x, &geo[0], tid);
#ifdef PF_RANSAC_TIMING
  long long t_phase = clock64();
#endif
  for (int it = 0; it < n_obj; ++it) {
    const int obj = (int)blockIdx.x + it * G;
    cp_async_wait_all();
    __syncthreads();
    read_geom(&geo[it & 1], g);
    if (!g.simple) {
      // skewed / projective K^-1 (shared by the batch, so every CTA sees it at its first object, before any copy is
      // in flight): this kernel does nothing; fit_ransac_kernel, launched next, sees the flag and does the work
      if (tid == 0) *p.redo_flag = 1;
      return;
    }
    if (it == 0 && blockIdx.x == 0 && tid == 0) *p.redo_flag = 0;
    if (tid == 0 && it == 0) issue_tile<false>(p, stage, &full[0], obj, 0, P, false);
    if (it + 1 < n_obj) fetch_geom(p, obj + G, &geo[(it + 1) & 1], tid);
    {
      // this thread's sample indices are needed only after pass 1: pull their lines into L1 now
      const int32_t* gi = p.sample_idx + (size_t)obj * p.n_hyp * p.n_samp;
      for (int h = tid; h < p.n_hyp; h += NT) {
        asm volatile("prefetch.global.L1 [%0];" ::"l"(gi + h * p.n_samp));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(gi + h * p.n_samp + p.n_samp - 1));
      }
    }
    for (int i = tid; i < p.W; i += NT) {
      const double r = g.k0 * (double)(g.x0 + i) + g.k2;
      rxc[i] = r;
      rxf[i] = (float)r;
      cwf[i] = (float)fma(r, r, 1.0);
    }
    for (int i = tid; i < p.H; i += NT) {
      const double r = g.k4 * (double)(g.y0 + i) + g.k5;
      ryr[i] = r;
      ryf[i] = (float)r;
      rwf[i] = (float)(r * r);
    }
    __syncthreads();
    mbar_wait(&full[0], (uint32_t)(it & 1));
    PF_PHASE(0);                                                  // geometry + wait for the crop

    const int32_t* gidx = p.sample_idx + (size_t)obj * p.n_hyp * p.n_samp;
    // ---- pass 1 ----------------------------------------------------------------------------------------------------
    {
      double acc[kCropAcc];
      crop_pass1<NT>(p, stage, rxc, ryr, cwf, rwf, bits, tid, acc);
      PF_PHASE(1);
      block_reduce<kCropAcc, NT>(acc, red, mom, tid);
    }
    __syncthreads();
    PF_PHASE(2);

    // ---- global statistics (warp 0, one output per lane), bitmap prefix (warp 1), thresholds (last warp) -----------
    if (warp == 0) {
      // mom holds RAW totals (a = noc, z): with h = 1/2, exactly in double:
      //   sum x_j = Sa_j - h n,  sum y_i x_j = +-(Sya_ij - h Sy_i),  sum |x|^2 = Saa - (Sa_0 + Sa_1 + Sa_2) + 3/4 n
      if (lane < kCropAcc) tot[lane] = mom[lane];
      const double h = 0.5, n = mom[0];
      const double rn = n > 0.0 ? 1.0 / n : 0.0;
      auto sx = [&](int j) { return mom[1 + j] - h * n; };
      auto sy = [&](int i) { return i == 2 ? -mom[6] : mom[4 + i]; };
      if (lane < 9) {
        const int i = lane / 3, j = lane - 3 * i;
        const double vv = mom[7 + 3 * i + j] - h * mom[4 + i];
        sh->Syx[lane] = (i == 2 ? -vv : vv) - n * (sy(i) * rn) * (sx(j) * rn);
      } else if (lane == 9) {
        const double sx2 = mom[16] - (mom[1] + mom[2] + mom[3]) + 0.75 * n;          // sum |x|^2 about the origin
        const double m0 = sx(0) * rn, m1 = sx(1) * rn, m2 = sx(2) * rn;
        sh->trSxx = sx2 - n * (m0 * m0 + m1 * m1 + m2 * m2);
        sh->x_rms = sqrt(fmax(sx2, 0.0) * rn);
      } else if (lane == 10) {
        const double m0 = sy(0) * rn, m1 = sy(1) * rn, m2 = sy(2) * rn;
        sh->Syy = mom[17] - n * (m0 * m0 + m1 * m1 + m2 * m2);
      } else if (lane < 14) {
        sh->mux[lane - 11] = sx(lane - 11) * rn;
      } else if (lane < 17) {
        sh->muy[lane - 14] = sy(lane - 14) * rn;
      } else if (lane == 17) {
        sh->n_valid = (int)n;
        sh->n = n;
        sh->winner = -1;
        sh->first_is_inlier = 0;
        sh->n_band = 0;
        sh->stopped = 0;
      }
    } else if (warp == NT / 32 - 1) {
      if (lane == 0) {
        const double n = mom[0];
        const double rn = n > 0.0 ? 1.0 / n : 0.0;
        const double s_norm = mom[18] * rn, t_norm = mom[19] * rn;            // pose_utils.py:91-92
        const double ts = t_norm / s_norm, st = s_norm / t_norm;              // :93-94
        const double pass_t = (st > ts ? st : ts) * p.ratio_adapt;            // :95
        const double stop_t = pass_t / 100.0;                                 // :96
        sh->pass_t = pass_t;
        sh->pass2 = pass_t * pass_t;
        sh->pass2_f = (float)(pass_t * pass_t);
        sh->stop2 = stop_t * stop_t;
      }
    }
    if (warp == 1 % (NT / 32)) {
      const int per = (p.n_words + 31) / 32;
      const int w0 = lane * per;
      uint32_t local = 0;
      for (int j = 0; j < per; ++j)
        if (w0 + j < p.n_words) local += __popc(bits[w0 + j]);
      uint32_t incl = local;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      uint32_t run = incl - local;
      int first = 0x7fffffff;
      for (int j = 0; j < per; ++j)
        if (w0 + j < p.n_words) {
          const uint32_t w = bits[w0 + j];
          prefix[w0 + j] = run;
          run += __popc(w);
          if (w != 0u && first == 0x7fffffff) first = (w0 + j) * 32 + __ffs(w) - 1;   // compacted point 0
        }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) first = min(first, __shfl_xor_sync(0xffffffffu, first, o));
      if (lane == 0) sh->first_px = (first == 0x7fffffff) ? -1 : first;
    }
    __syncthreads();
    PF_PHASE(3);

    // ---- select list: pixel of every even-ranked valid point, into the (now dead) mask plane ----------------------------
    for (int w = tid; w < p.n_words; w += NT) {
      const uint32_t v = bits[w], r0 = prefix[w];
      uint32_t pp = v;                                         // inclusive prefix parity of the word
      pp ^= pp << 1; pp ^= pp << 2; pp ^= pp << 4; pp ^= pp << 8; pp ^= pp << 16;
      uint32_t e = v & ((r0 & 1u) ? ~pp : pp);                 // set bits whose GLOBAL rank is even
      uint32_t qi = (r0 + 1u) >> 1;
      const int base = w * 32;
      while (e != 0u) {
        PF_CHECK(qi < (uint32_t)(P / 2));
        klist[qi++] = (uint16_t)(base + __ffs(e) - 1);
        e &= e - 1u;
      }
    }
    __syncthreads();
    PF_PHASE(4);

    const int N = sh->n_valid;
    const double stop2 = sh->stop2;
    // ---- S1: float screen (posefit_math.h: screen_fit32), one hypothesis per thread and round ----------------------------
    double hi_min = __longlong_as_double(0x7ff0000000000000LL);
    float R32[9] = {1.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 1.f}, T32[12], rho32 = 0.f, s32 = 1.f;
    bool have32 = false;                                         // R32 / T32 belong to hypothesis `tid` (n_hyp <= NT)
    float hi32 = __int_as_float(0x7f800000);                     // upper end of its residual interval, rounded up
#pragma unroll
    for (int i = 0; i < 12; ++i) T32[i] = 0.f;
    if (N > 0) {
      const double n_all = sh->n, x_rms = sh->x_rms, tr_sxx = sh->trSxx;
#pragma unroll 1
      for (int h = tid; h < p.n_hyp; h += NT) {
        float ox[3] = {0.f, 0.f, 0.f}, oy[3] = {0.f, 0.f, 0.f}, sx[3] = {0.f, 0.f, 0.f}, sy[3] = {0.f, 0.f, 0.f};
        float syx[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, sxx = 0.f, syy = 0.f;
        const int32_t* gi = gidx + h * p.n_samp;
        int k_next = __ldg(gi);
#pragma unroll 1
        for (int j = 0; j < p.n_samp; ++j) {
          const int k = sample_index(k_next, N, p.idx_bits);              // pose_utils.py:73
          if (j + 1 < p.n_samp) k_next = __ldg(gi + j + 1);
          PF_CHECK(k >= 0 && k < N);
          const int px = select_px_list(klist, bits, k);
          const int row = (int)__umulhi((uint32_t)px, p.w_magic), col = px - row * p.W;
          PF_CHECK(px >= 0 && px < P && row < p.H && col < p.W);
          const float z = sdep[px];
          const float xs[3] = {snoc[px] - 0.5f, snoc[P + px] - 0.5f, snoc[2 * P + px] - 0.5f};   // pose_estimation.py:323
          const float ys[3] = {rxf[col] * z, -(ryf[row] * z), -z};                                // :34-41
          if (j == 0) {
#pragma unroll
            for (int i = 0; i < 3; ++i) { ox[i] = xs[i]; oy[i] = ys[i]; }
          }
          float x[3], y[3];
#pragma unroll
          for (int i = 0; i < 3; ++i) { x[i] = xs[i] - ox[i]; y[i] = ys[i] - oy[i]; }
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            sx[i] += x[i];
            sy[i] += y[i];
            sxx = fmaf(x[i], x[i], sxx);
            syy = fmaf(y[i], y[i], syy);
#pragma unroll
            for (int jj = 0; jj < 3; ++jj) syx[3 * i + jj] = fmaf(y[i], x[jj], syx[3 * i + jj]);
          }
        }
        PF_PHASE(5);                                              // float gathers
        ScreenFit sf;
        screen_fit32(p.n_samp, sx, sy, syx, sxx, syy, ox, oy, p.ref_compat != 0, sf);
        double A[9], t[3];
#pragma unroll
        for (int i = 0; i < 9; ++i) A[i] = (double)sf.A[i];
#pragma unroll
        for (int i = 0; i < 3; ++i) t[i] = (double)sf.t[i];
        const double r2 = residual_sq_iso(n_all, sh->mux, sh->muy, sh->Syy, sh->Syx, tr_sxx, A, t, (double)sf.s);
        const double e = screen_interval(sf, r2, n_all, x_rms, tr_sxx);
        slo[h] = __double2float_rd(r2 - e);                       // -inf: the float fit is not usable
        const double hi = r2 + e;
        if (hi < hi_min) hi_min = hi;
        if (p.n_hyp <= NT) {
          have32 = sf.rho < 1e30f;
          hi32 = __double2float_ru(hi);
          rho32 = sf.rho;
          s32 = sf.s;
#pragma unroll
          for (int i = 0; i < 9; ++i) { R32[i] = sf.R[i]; T32[i] = sf.A[i]; }
#pragma unroll
          for (int i = 0; i < 3; ++i) T32[9 + i] = sf.t[i];
        }
        PF_PHASE(6);                                              // float fit + residual interval
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) hi_min = fmin(hi_min, __shfl_xor_sync(0xffffffffu, hi_min, o));
    if (lane == 0) sh->hi_min[warp] = hi_min;
    __syncthreads();
    double U = sh->hi_min[0];
#pragma unroll
    for (int w = 1; w < NT / 32; ++w) U = fmin(U, sh->hi_min[w]);
    // candidates: hypotheses that can win (interval reaches below every upper end) or stop (below StopT); a NaN stays in
    int my_cands = 0;
    if (N > 0) {
      for (int h = tid; h < p.n_hyp; h += NT) {
        const double lo = (double)slo[h];
        if (!(lo > U && lo >= stop2)) ++my_cands;
      }
    }
    const int n_cand = __syncthreads_count(my_cands != 0);        // (threads, = candidates when n_hyp <= NT)
    PF_PHASE(7);                                                  // wait for the other warps + candidate count

    int win = -1;
    const bool single = (n_cand == 1) && (p.n_hyp <= NT) && !p.no_screen;
    bool float_pass = false;
    if (single) {
      // One candidate: the winner, whatever its exact residual (if it is below StopT it stops the search AT itself).
      // With a usable float fit the inlier pass starts on it right away; the double fit only happens if a pixel lands
      // in the guard band.  band_rel covers float(T) against the double fit: |d r| <= s rho (|x| <= 0.87) + |dt|.
      if (my_cands != 0) {
        sh->winner = tid;
        // did it stop the search (residual < StopT)?  Known from the interval unless that straddles StopT (then the
        // double fit below decides)
        sh->stopped = ((double)hi32 < stop2) ? 1 : (((double)slo[tid] >= stop2) ? 0 : 2);
        const float pass_t = (float)sh->pass_t;
        const float amx = fabsf(T32[9]) + fabsf(T32[10]) + fabsf(T32[11]);
        const float dr = have32 ? (fabsf(s32) * rho32 * 1.8f + 8.f * kScreenEps * (amx + 1.f)) : 1e30f;
        // relative half-width in r^2: 2 dr / PassT, floor 1e-4 (float arithmetic of the screen: ~1e-6), doubled
        const float band = 2.0f * fmaxf(1e-4f, 2.0f * dr / pass_t);
        sh->band_rel = band;
#pragma unroll
        for (int i = 0; i < 12; ++i) sh->w32[i] = T32[i];
#pragma unroll
        for (int i = 0; i < 9; ++i) sh->start[0][i] = R32[i];   // the polish, if it happens, is warp 0's
        sh->start[0][9] = have32 ? 1.0f : 0.0f;
      }
      __syncthreads();
      win = sh->winner;
      float_pass = sh->band_rel < 0.05f && sh->stopped != 2;      // an unusable / sloppy float fit: take the exact path
      // the winner's sample pixels, resolved now: the inlier pass reuses the select list's memory for its queues
      if (warp == 0) {
        crop_sample_pixels(klist, bits, gidx + win * p.n_samp, p.n_samp, N, p.idx_bits, sh->spx[0]);
        if (!float_pass) {
          const double r2w = crop_fit_hypothesis(snoc, sdep, rxc, ryr, sh->spx[0], sh, red, sh->wtf, p.n_samp, P, p.W,
                                                 p.w_magic, p.ref_compat);
          if (lane == 0 && sh->stopped == 2) sh->stopped = (r2w < stop2) ? 1 : 0;
        }
      }
      __syncthreads();                                            // (the select list is free from here on)
    } else if (n_cand > 0) {
      // Two or more candidates (or several hypotheses per thread): double fits first, then the reference's selection
      // (pose_utils.py:68-81): first h below StopT, else the first minimum.  Warps take candidates in turn.
      double best = 1e20;                                // (1e10)^2, :68
      int best_h = 0x7fffffff, stop_h = 0x7fffffff;
#pragma unroll 1
      for (int base = 0; base < p.n_hyp; base += NT) {
        const int h = base + tid;
        bool cand = false;
        if (h < p.n_hyp) {
          const double lo = (double)slo[h];
          cand = !(lo > U && lo >= stop2) || p.no_screen;
        }
        uint32_t cm = __ballot_sync(0xffffffffu, cand);
#pragma unroll 1
        while (cm != 0u) {
          const int src = __ffs(cm) - 1;
          cm &= cm - 1u;
          const int hc = base + (warp << 5) + src;
          if (lane == src) {
#pragma unroll
            for (int i = 0; i < 9; ++i) sh->start[warp][i] = R32[i];
            sh->start[warp][9] = (have32 && p.n_hyp <= NT) ? 1.0f : 0.0f;
          }
          __syncwarp();
          crop_sample_pixels(klist, bits, gidx + hc * p.n_samp, p.n_samp, N, p.idx_bits, sh->spx[warp]);
          const double r2 = crop_fit_hypothesis(snoc, sdep, rxc, ryr, sh->spx[warp], sh, red + warp * 24, cur + warp * 12,
                                                p.n_samp, P, p.W, p.w_magic, p.ref_compat);
          // keep the transform of this warp's leading hypothesis
          bool keep = false;
          if (r2 < stop2 && stop_h == 0x7fffffff) { stop_h = hc; keep = true; }
          if (r2 < best) { best = r2; best_h = hc; keep = keep || stop_h == 0x7fffffff; }
          if (keep && lane < 12) kept[warp * 12 + lane] = cur[warp * 12 + lane];
          __syncwarp();
        }
      }
      // every lane of a warp holds the same (best, best_h, stop_h); combine the warps
      if (lane == 0) {
        red[warp * 24] = best;
        reinterpret_cast<int2*>(red + warp * 24 + 1)[0] = make_int2(best_h, stop_h);
      }
      __syncthreads();
      best = 1e20; best_h = 0x7fffffff; stop_h = 0x7fffffff;
      int best_w = -1, stop_w = -1;
#pragma unroll
      for (int w = 0; w < NT / 32; ++w) {
        const double ob = red[w * 24];
        const int2 hh = reinterpret_cast<const int2*>(red + w * 24 + 1)[0];
        if (ob < best || (ob == best && hh.x < best_h)) { best = ob; best_h = hh.x; best_w = w; }
        if (hh.y < stop_h) { stop_h = hh.y; stop_w = w; }
      }
      win = (stop_h != 0x7fffffff) ? stop_h : (best_h != 0x7fffffff ? best_h : -1);
      if (tid == 0) sh->stopped = (stop_h != 0x7fffffff) ? 1 : 0;
      const int win_w = (stop_h != 0x7fffffff) ? stop_w : best_w;
      if (win >= 0 && tid < 12) sh->wtf[tid] = kept[win_w * 12 + tid];
      __syncthreads();
    }
    PF_PHASE(8);                                                  // candidates / winner broadcast

    // ---- pass 2: inlier mask of the winner + moments of the outliers ---------------------------------------------------
    {
      LaneSums acc;
      acc.clear();
      int n_inl = 0;
      uint8_t* om = p.inlier_mask + (size_t)obj * P;
      if (float_pass) {
        crop_pass2<NT, false>(p, stage, rxc, ryr, rxf, ryf, bits, sh, win, om, klist, band_q, tid, acc, n_inl);
        // guard-band pixels: fit the winner in double (warp 0), then decide them as the reference does (:7-10)
        if (__syncthreads_or(sh->n_band != 0)) {
          if (warp == 0)
            crop_fit_hypothesis(snoc, sdep, rxc, ryr, sh->spx[0], sh, red, sh->wtf, p.n_samp, P, p.W, p.w_magic, p.ref_compat);
          __syncthreads();
          const int nb = min(sh->n_band, kBandCap);
          const bool overflow = sh->n_band > kBandCap;              // absurdly many: decide EVERY pixel again, exactly
          const double pass2 = sh->pass2;
          if (!overflow) {
            for (int e = tid; e < nb; e += NT) {
              const int px = (int)band_q[e];
              const int row = (int)__umulhi((uint32_t)px, p.w_magic), col = px - row * p.W;
              const double a0 = (double)snoc[px], a1 = (double)snoc[P + px], a2 = (double)snoc[2 * P + px], zd = (double)sdep[px];
              const double xd0 = a0 - 0.5, xd1 = a1 - 0.5, xd2 = a2 - 0.5;
              const double y0 = rxc[col] * zd, y1 = -(ryr[row] * zd);
              const double e0 = y0 - (sh->wtf[0] * xd0 + sh->wtf[1] * xd1 + sh->wtf[2] * xd2 + sh->wtf[9]);
              const double e1 = y1 - (sh->wtf[3] * xd0 + sh->wtf[4] * xd1 + sh->wtf[5] * xd2 + sh->wtf[10]);
              const double e2 = -zd - (sh->wtf[6] * xd0 + sh->wtf[7] * xd1 + sh->wtf[8] * xd2 + sh->wtf[11]);
              if ((e0 * e0 + e1 * e1 + e2 * e2) < pass2) {
                om[px] = 1;
                ++n_inl;
                if (px == sh->first_px) sh->first_is_inlier = 1;
              } else {
                acc.add(a0, a1, a2, y0, y1, zd);
                ++acc.cnt;
              }
            }
          } else {
            __syncthreads();
            acc.clear();
            n_inl = 0;
            if (tid == 0) sh->first_is_inlier = 0;
            __syncthreads();
            crop_pass2<NT, true>(p, stage, rxc, ryr, rxf, ryf, bits, sh, win, om, klist, band_q, tid, acc, n_inl);
          }
        }
      } else {
        crop_pass2<NT, true>(p, stage, rxc, ryr, rxf, ryf, bits, sh, win, om, klist, band_q, tid, acc, n_inl);
      }
      PF_PHASE(9);
      double outl[kAccPlain + 1];
      outl[0] = (double)acc.cnt;
#pragma unroll
      for (int i = 0; i < 3; ++i) { outl[1 + i] = acc.sa[i]; outl[4 + i] = acc.sy[i]; }
#pragma unroll
      for (int i = 0; i < 9; ++i) outl[7 + i] = acc.sya[i];
      outl[16] = acc.saa;
      outl[17] = (double)n_inl;
      block_reduce<kAccPlain + 1, NT>(outl, red, mom, tid);        // mom[0..16] raw OUTLIER sums, mom[17] = #inliers
      // every thread is past the barrier inside the reduction, i.e. done with the crop: request the next one now
      if (tid == 0 && it + 1 < n_obj) issue_tile<false>(p, stage, &full[0], obj + G, 0, P, false);
    }
    __syncthreads();
    {
      // inliers = all valid - outliers (raw sums), then centre the source (x = noc - 0.5) and flip z (y2 = -z)
      double* rec = p.ws + (size_t)obj * kRansacRecord;
      const double h = 0.5;
      auto r = [&](int i) { return tot[i] - mom[i]; };              // i < 16: same layout in both; 16: sum |a|^2
      const double n_inl = r(0);
      if (tid < kAccPlain) {
        double v;
        if (tid == 0) v = n_inl;
        else if (tid < 4) v = r(tid) - h * n_inl;
        else if (tid < 6) v = r(tid);
        else if (tid == 6) v = -r(6);
        else if (tid < 13) v = r(tid) - h * r(tid < 10 ? 4 : 5);
        else if (tid < 16) v = -(r(tid) - h * r(6));
        else v = r(16) - 2.0 * h * (r(1) + r(2) + r(3)) + 3.0 * h * h * n_inl;
        rec[tid] = v;
      }
      if (tid == 32 % NT) {
        // the reference counts non-zero INDEX values: compacted point 0 is never counted (F5)
        rec[17] = (double)N;
        rec[18] = n_inl - ((p.ref_compat != 0 && sh->first_is_inlier) ? 1.0 : 0.0);
        rec[19] = sh->pass_t;
        rec[20] = (double)win;
        rec[21] = (win >= 0) ? 1.0 : 0.0;
        // iterations the reference's loop runs (= 10 np.random draws each): up to and including the one that stops it
        rec[22] = (N > 0) ? (double)((sh->stopped == 1 && win >= 0) ? win + 1 : p.n_hyp) : 0.0;
      }
    }
    __syncthreads();                                              // stage, mom and sh are free again
    PF_PHASE(10);
  }
}

}  // namespace posefit
