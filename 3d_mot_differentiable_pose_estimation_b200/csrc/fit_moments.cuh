// fit_moments.cuh -- K-moments + K-solve: the plain Umeyama fit (pose_utils.py:16-61 over pose_estimation.py:16-43, :323)
// Part of libposefit_b200.so: included by posefit_kernels.cu (one translation unit, so every kernel sees the
// same inlined helpers and the build stays a single nvcc call).  See include/posefit.h for the C ABI.
#pragma once

#include "posefit_common.cuh"

namespace posefit {

// ---------------------------------------------------------------------------------------------
// K-moments + K-solve: plain fit (BASELINE configs 1, 2, 4-forward, 5-forward)
//
// v1 of this path staged row bands through a CTA-wide TMA ring and reduced per object across the
// whole CTA; ncu (profiles/r01_a_*) showed 78 % of its instructions in per-tile / per-object
// overhead (16 warps each running the full shuffle reduction, barriers, tile bookkeeping) for
// 8 pixels of work per thread.  v2 gives every WARP its own contiguous range of 128-pixel chunks:
// no block barrier, one shuffle reduction per (warp, object), 128-bit streaming loads that are
// requested one chunk ahead, partial moments to a small workspace, and a second tiny kernel
// (programmatic dependent launch) that merges the parts and does the 3x3 solves.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void accumulate_plain(double* acc, double x0, double x1, double x2, double y0, double y1,
                                                 double y2) {
  acc[0] += 1.0;
  acc[1] += x0; acc[2] += x1; acc[3] += x2;
  acc[4] += y0; acc[5] += y1; acc[6] += y2;
  acc[7] = fma(y0, x0, acc[7]);   acc[8] = fma(y0, x1, acc[8]);   acc[9] = fma(y0, x2, acc[9]);
  acc[10] = fma(y1, x0, acc[10]); acc[11] = fma(y1, x1, acc[11]); acc[12] = fma(y1, x2, acc[12]);
  acc[13] = fma(y2, x0, acc[13]); acc[14] = fma(y2, x1, acc[14]); acc[15] = fma(y2, x2, acc[15]);
  acc[16] = fma(x0, x0, fma(x1, x1, fma(x2, x2, acc[16])));
}

constexpr int kChunkPx = 128;     // pixels per warp iteration: 4 consecutive pixels per lane
constexpr int kChunkBytes = 2176; // one warp's chunk in shared memory: 4 float4 planes + uchar4 per lane

__device__ __forceinline__ void cp_async_16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
// variants taking the 32-bit shared-window address directly (no generic -> shared conversion per copy)
__device__ __forceinline__ void cp_async_16_s(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_4_s(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
// x & m on the raw bits, opaque to the optimiser: masking the FLOAT (one instruction) instead of the converted
// double (the compiler otherwise moves the select behind the conversion: two FSELs per value)
__device__ __forceinline__ float and_bits(float x, uint32_t m) {
  uint32_t r;
  asm("and.b32 %0, %1, %2;" : "=r"(r) : "r"(__float_as_uint(x)), "r"(m));
  return __uint_as_float(r);
}
template <int N>
__device__ __forceinline__ void cp_async_wait_group() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// Request this lane's 4 pixels into its own slots of `stage` (LDGSTS: no register staging,
// completion tracked per thread by cp.async groups).  Every lane reads back only what it requested
// itself, so the per-warp ring needs no barrier at all.  n0 / dz / mk point at this lane's first
// pixel in the NOC plane 0, the depth crop and the mask crop.
// VEC: 0 = ragged shapes / unaligned pointers, 1 = 128-bit copies (P % 4 == 0, aligned bases), 2 = as 1 and
// every chunk is full (P % 128 == 0: no end-of-object tests).  stage_s = shared-window address of `stage`.
template <int VEC>
__device__ __forceinline__ void request_chunk(unsigned char* stage, uint32_t stage_s, const float* n0, const float* dz,
                                              const uint8_t* mk, int P, int px, int lane) {
  if (VEC != 2 && px >= P) return;
  unsigned char* s = stage + lane * 16;
  if (VEC != 0) {
    const uint32_t s32 = stage_s + lane * 16;
    cp_async_16_s(s32, n0);
    cp_async_16_s(s32 + 512, n0 + P);
    cp_async_16_s(s32 + 1024, n0 + 2 * (size_t)P);
    cp_async_16_s(s32 + 1536, dz);
    cp_async_4_s(stage_s + 2048 + lane * 4, mk);
  } else {
    // ragged shapes / unaligned pointers: 4-byte copies for the floats, plain byte loads for the mask
    unsigned char mm[4] = {0, 0, 0, 0};
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (px + j < P) {
        cp_async_4(s + 4 * j, n0 + j);
        cp_async_4(s + 512 + 4 * j, n0 + P + j);
        cp_async_4(s + 1024 + 4 * j, n0 + 2 * (size_t)P + j);
        cp_async_4(s + 1536 + 4 * j, dz + j);
        mm[j] = mk[j];
      }
    *reinterpret_cast<uchar4*>(stage + 2048 + lane * 4) = make_uchar4(mm[0], mm[1], mm[2], mm[3]);
  }
}

// Per-lane accumulators of the plain path.  To keep the inner loop at 20 fp64 operations per pixel
// they hold sums of a = noc (NOT noc - 0.5) and, in crop mode, of z instead of y2 = -z; invalid
// pixels are folded in as zeros (branch-free), the count is an integer.  `finish` turns the
// warp-reduced sums into the Moments layout (n, sum x, sum y, sum y x^T, sum |x|^2) exactly:
//   x = a - h  =>  sum x = Sa - h n,  sum y_i x_j = Sya_ij - h Sy_i,  sum|x|^2 = Saa - 2h sum Sa + 3 h^2 n.
struct LaneSums {
  double sa[3], sy[3], sya[9], saa;
  int cnt;
  __device__ __forceinline__ void clear() {
#pragma unroll
    for (int i = 0; i < 3; ++i) { sa[i] = 0.0; sy[i] = 0.0; }
#pragma unroll
    for (int i = 0; i < 9; ++i) sya[i] = 0.0;
    saa = 0.0;
    cnt = 0;
  }
  __device__ __forceinline__ void add(double a0, double a1, double a2, double y0, double y1, double y2) {
    sa[0] += a0; sa[1] += a1; sa[2] += a2;
    sy[0] += y0; sy[1] += y1; sy[2] += y2;
    sya[0] = fma(y0, a0, sya[0]); sya[1] = fma(y0, a1, sya[1]); sya[2] = fma(y0, a2, sya[2]);
    sya[3] = fma(y1, a0, sya[3]); sya[4] = fma(y1, a1, sya[4]); sya[5] = fma(y1, a2, sya[5]);
    sya[6] = fma(y2, a0, sya[6]); sya[7] = fma(y2, a1, sya[7]); sya[8] = fma(y2, a2, sya[8]);
    saa = fma(a0, a0, fma(a1, a1, fma(a2, a2, saa)));
  }
  // warp reduction + conversion; the result is valid in every lane.  h = 0.5 in crop mode
  // (pose_estimation.py:323), neg2: the third target component was accumulated as +z (:41).
  __device__ __forceinline__ void finish(double h, bool neg2, double* out /*[17]*/) {
    double v[16];
#pragma unroll
    for (int i = 0; i < 3; ++i) { v[i] = sa[i]; v[3 + i] = sy[i]; }
#pragma unroll
    for (int i = 0; i < 9; ++i) v[6 + i] = sya[i];
    v[15] = saa;
    int c = cnt;
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) c += __shfl_xor_sync(0xffffffffu, c, s);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      double x = v[i];
#pragma unroll
      for (int s = 16; s > 0; s >>= 1) x += __shfl_xor_sync(0xffffffffu, x, s);
      v[i] = x;
    }
    const double n = (double)c;
    const double sg = neg2 ? -1.0 : 1.0;
    out[0] = n;
#pragma unroll
    for (int j = 0; j < 3; ++j) out[1 + j] = v[j] - h * n;
    out[4] = v[3];
    out[5] = v[4];
    out[6] = sg * v[5];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      out[7 + j] = v[6 + j] - h * v[3];
      out[10 + j] = v[9 + j] - h * v[4];
      out[13 + j] = sg * (v[12 + j] - h * v[5]);
    }
    out[16] = v[15] - 2.0 * h * (v[0] + v[1] + v[2]) + 3.0 * h * h * n;
  }
};

// PAIR (only with VEC == 2, an even number of chunks per object and per warp): two chunks per loop iteration --
// one set of request / object / ring bookkeeping per 256 pixels, the second chunk addressed by immediate
// offsets.  The per-chunk overhead (a third of the loop's instructions) is what makes this kernel issue-limited
// when the board's power cap lowers the SM clock.
// DYN (long batches of full chunks, at least DEPTH chunks per object): a warp takes WHOLE objects -- object gw first, the
// following ones from a ticket counter in the workspace (zeroed by the host before the launch) -- instead of a fixed
// range of the chunk stream.  SMs do not stream at the same rate (tools/trace_step.py: on the 125 000-object shard the
// last warps of some TPCs finish 60 us after most others), and with equal ranges the slowest SM sets the kernel's time.
// The request stream draws the next ticket when it enters an object, a whole object ahead of its use; the consumer
// follows it (it is less than one object behind).  Every object is summed by ONE warp in a fixed lane / chunk order,
// whichever warp that is: results do not depend on the assignment.
template <bool POINTS, int DEPTH, int VEC, bool PAIR = false, bool DYN = false>
__global__ void __launch_bounds__(512, 1) fit_moments_kernel(const FwdParams p) {
  static_assert(!PAIR || (!POINTS && VEC == 2 && DEPTH % 2 == 0 && DEPTH >= 4), "PAIR: full aligned chunks, even ring");
  static_assert(!DYN || (!POINTS && VEC == 2 && !PAIR), "DYN: full aligned chunks, one chunk per iteration");
  extern __shared__ __align__(128) unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned char* ring = smem + (size_t)warp * p.warp_smem_bytes;               // DEPTH stages of kChunkBytes
  double* rxc = reinterpret_cast<double*>(ring + DEPTH * kChunkBytes);         // this warp's ray tables
  double* ryr = rxc + p.W;
#if __CUDA_ARCH__ >= 900
  // Dependents first: the CTAs of K-solve become resident (where registers and shared memory
  // allow) while this kernel still runs, and block in their own griddepcontrol.wait until this grid
  // has completed.  Every kernel of the chain touches global memory only after its own wait, so
  // completion order (every RAW / WAR dependence between consecutive kernels) is unchanged.
  PF_TRACE_BEGIN(0);
  PF_TRACE_END(4);                                            // (latest first instruction: the launch stagger)
  if (p.early_dep & 1) asm volatile("griddepcontrol.launch_dependents;");
  asm volatile("griddepcontrol.wait;" ::: "memory");          // inputs may come from the previous kernel in the stream
  if (!(p.early_dep & 1)) asm volatile("griddepcontrol.launch_dependents;");
  PF_TRACE_BEGIN(8);
#endif
  const long long gw = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
  const long long c_begin = DYN ? gw * p.chunks_per_obj : gw * p.chunks_per_warp;
  if (c_begin >= p.total_chunks) return;
  const int n_chunks = (int)min((long long)p.chunks_per_warp, p.total_chunks - c_begin);

  const int cpo = p.chunks_per_obj;
  int obj = (int)(c_begin / cpo);
  int ch = (int)(c_begin - (long long)obj * cpo);
  int ticket = 0;                                             // DYN, lane 0: the object after the one being requested
  if (DYN && lane == 0) ticket = p.dyn_base + (int)atomicAdd(p.dyn_counter, 1u);
  LaneSums acc;
  acc.clear();
  ObjGeom g = {};
  int cur_obj = -1;
  int row = 0, col = 0;                                       // of this lane's first pixel in the chunk
  const int drow = kChunkPx / p.W, dcol = kChunkPx % p.W;
  const bool row_fast = (p.W % 4 == 0);                       // a lane's 4 pixels never straddle a row
  bool fast_px = false;                                       // row_fast and a pinhole K (set per object)

  auto write_part = [&](int o) {
    double out[kAccPlain];
    acc.finish(POINTS ? 0.0 : 0.5, !POINTS, out);
    if (lane == 0) {
      const long long first = ((long long)o * cpo) / p.chunks_per_warp;     // first warp that touches object o
      double* w = p.ws + ((size_t)o * p.max_parts + (DYN ? (size_t)0 : (size_t)(gw - first))) * kAccPlain;
#pragma unroll
      for (int i = 0; i < kAccPlain; ++i) w[i] = out[i];
    }
  };

  // request stream: runs DEPTH-1 chunks ahead of the consumer; running pointers, no per-chunk
  // address arithmetic beyond three increments
  const int P = p.P;
  const uint32_t ring_s = smem_u32(ring);
  int q_left = n_chunks, q_obj = obj, q_ch = ch, q_slot = 0;
  int q_px = ch * kChunkPx + 4 * lane;
  const float* q_n0 = p.noc + (size_t)obj * 3 * P + q_px;
  const float* q_dz = p.depth + (size_t)obj * P + q_px;
  const uint8_t* q_mk = p.mask + (size_t)obj * P + q_px;
  auto request_next = [&]() {
    if (DYN ? q_obj < p.B : q_left > 0) {
      request_chunk<VEC>(ring + q_slot * kChunkBytes, ring_s + q_slot * kChunkBytes, q_n0, q_dz, q_mk, P, q_px, lane);
      --q_left;
      if (++q_slot == DEPTH) q_slot = 0;
      if (++q_ch == cpo) {
        q_ch = 0;
        if (DYN) {
          q_obj = __shfl_sync(0xffffffffu, ticket, 0);        // drawn when the stream entered the object it now leaves
          if (lane == 0 && q_obj < p.B) ticket = p.dyn_base + (int)atomicAdd(p.dyn_counter, 1u);
          if (q_obj > p.B) q_obj = p.B;                       // (pointer arithmetic below stays inside one object past the end)
        } else {
          ++q_obj;
        }
        q_px = 4 * lane;
        q_n0 = p.noc + (size_t)q_obj * 3 * P + q_px;
        q_dz = p.depth + (size_t)q_obj * P + q_px;
        q_mk = p.mask + (size_t)q_obj * P + q_px;
      } else {
        q_px += kChunkPx;
        q_n0 += kChunkPx;
        q_dz += kChunkPx;
        q_mk += kChunkPx;
      }
    }
    cp_async_commit();
  };
  int slot = 0;
  int px0 = ch * kChunkPx + 4 * lane;
  if constexpr (PAIR) {
    auto request_pair = [&]() {
      if (q_left > 0) {
        const uint32_t sa = ring_s + q_slot * kChunkBytes + lane * 16;           // slot A; slot B follows it
        cp_async_16_s(sa, q_n0);
        cp_async_16_s(sa + kChunkBytes, q_n0 + kChunkPx);
        cp_async_16_s(sa + 512, q_n0 + P);
        cp_async_16_s(sa + kChunkBytes + 512, q_n0 + P + kChunkPx);
        cp_async_16_s(sa + 1024, q_n0 + 2 * (size_t)P);
        cp_async_16_s(sa + kChunkBytes + 1024, q_n0 + 2 * (size_t)P + kChunkPx);
        cp_async_16_s(sa + 1536, q_dz);
        cp_async_16_s(sa + kChunkBytes + 1536, q_dz + kChunkPx);
        const uint32_t ma = ring_s + q_slot * kChunkBytes + 2048 + lane * 4;
        cp_async_4_s(ma, q_mk);
        cp_async_4_s(ma + kChunkBytes, q_mk + kChunkPx);
        q_left -= 2;
        q_slot += 2;
        if (q_slot == DEPTH) q_slot = 0;
        q_ch += 2;
        if (q_ch == cpo) {
          q_ch = 0;
          ++q_obj;
          q_n0 = p.noc + (size_t)q_obj * 3 * P + 4 * lane;
          q_dz = p.depth + (size_t)q_obj * P + 4 * lane;
          q_mk = p.mask + (size_t)q_obj * P + 4 * lane;
        } else {
          q_n0 += 2 * kChunkPx;
          q_dz += 2 * kChunkPx;
          q_mk += 2 * kChunkPx;
        }
      }
      cp_async_commit();
    };
#pragma unroll
    for (int i = 0; i < DEPTH / 2 - 1; ++i) request_pair();
    for (int it = 0; it < n_chunks; it += 2) {
      request_pair();                                           // refills the two stages consumed last iteration
      if (obj != cur_obj) {
        if (cur_obj >= 0) write_part(cur_obj);
        cur_obj = obj;
        acc.clear();
        const double* K = p.kinv + (p.kinv_per_object ? 9 * (size_t)obj : 0);
        g.k = K;
        g.k0 = K[0]; g.k2 = K[2]; g.k4 = K[4]; g.k5 = K[5];
        g.x0 = p.bbox[2 * (size_t)obj];
        g.y0 = p.bbox[2 * (size_t)obj + 1];
        g.simple = (K[1] == 0.0 && K[3] == 0.0 && K[6] == 0.0 && K[7] == 0.0 && K[8] == 1.0);
        fast_px = row_fast && g.simple;
        __syncwarp();
        build_ray_tables(p, g, rxc, ryr, lane, 32);
        __syncwarp();
        row = px0 / p.W;
        col = px0 - row * p.W;
      }
      cp_async_wait_group<DEPTH / 2 - 1>();                     // this lane's copies of both chunks have landed
      PF_CHECK(slot >= 0 && slot + 1 < DEPTH && obj < p.B && px0 + kChunkPx + 3 < P);
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const unsigned char* st = ring + (slot + u) * kChunkBytes + lane * 16;
        const uint32_t m4 = *reinterpret_cast<const uint32_t*>(ring + (slot + u) * kChunkBytes + 2048 + lane * 4);
        const float4 z4 = *reinterpret_cast<const float4*>(st + 1536);
        const float zz[4] = {z4.x, z4.y, z4.z, z4.w};
        bool ok[4];
        bool any = false;
#pragma unroll
        for (int j = 0; j < 4; ++j) {                           // pose_estimation.py:23-25
          ok[j] = (m4 & (0xffu << (8 * j))) != 0u && zz[j] > 0.0f;
          any = any || ok[j];
        }
        if (p.valid_mask != nullptr)
          *reinterpret_cast<uint32_t*>(p.valid_mask + (size_t)obj * P + px0 + u * kChunkPx) =
              (ok[0] ? 1u : 0u) | (ok[1] ? 0x100u : 0u) | (ok[2] ? 0x10000u : 0u) | (ok[3] ? 0x1000000u : 0u);
        if (__any_sync(0xffffffffu, any)) {
          const float4 a4 = *reinterpret_cast<const float4*>(st);
          const float4 b4 = *reinterpret_cast<const float4*>(st + 512);
          const float4 c4 = *reinterpret_cast<const float4*>(st + 1024);
          const float n0[4] = {a4.x, a4.y, a4.z, a4.w};
          const float n1[4] = {b4.x, b4.y, b4.z, b4.w};
          const float n2[4] = {c4.x, c4.y, c4.z, c4.w};
          if (fast_px) {
            const double nry = -ryr[row];
            const double2 rxa = *reinterpret_cast<const double2*>(rxc + col);
            const double2 rxb = *reinterpret_cast<const double2*>(rxc + col + 2);
            const double rx[4] = {rxa.x, rxa.y, rxb.x, rxb.y};
            uint32_t okm[4];                                    // all ones / zero per pixel
#pragma unroll
            for (int j = 0; j < 4; ++j) okm[j] = ok[j] ? 0xffffffffu : 0u;
            acc.cnt -= (int)(okm[0] + okm[1]) + (int)(okm[2] + okm[3]);
#pragma unroll
            for (int j = 0; j < 4; ++j) {                       // branch-free: invalid pixels contribute zeros
              const double zd = (double)and_bits(zz[j], okm[j]);
              const double a0 = (double)and_bits(n0[j], okm[j]);
              const double a1 = (double)and_bits(n1[j], okm[j]);
              const double a2 = (double)and_bits(n2[j], okm[j]);
              acc.add(a0, a1, a2, rx[j] * zd, nry * zd, zd);    // y = (rx z, -ry z, [-]z), :34-41
            }
          } else {
            int r = row, cc = col;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (ok[j]) {
                double y0, y1, y2;
                backproject_px(g, rxc, ryr, r, cc, (double)zz[j], y0, y1, y2);
                acc.add((double)n0[j], (double)n1[j], (double)n2[j], y0, y1, -y2);
                ++acc.cnt;
              }
              if (++cc >= p.W) { cc = 0; ++r; }
            }
          }
        }
        row += drow;
        col += dcol;
        if (col >= p.W) { col -= p.W; ++row; }
      }
      px0 += 2 * kChunkPx;
      ch += 2;
      if (ch == cpo) { ch = 0; ++obj; px0 = 4 * lane; }
      slot += 2;
      if (slot == DEPTH) slot = 0;
    }
    write_part(cur_obj);
    PF_TRACE_END(0);
    PF_TRACE_BEGIN(5);                                        // (earliest warp done)
    PF_TRACE_WARP_END();
    return;
  }
  if (!POINTS) {
#pragma unroll
    for (int i = 0; i < DEPTH - 1; ++i) request_next();
  }

  for (int it = 0; DYN ? obj < p.B : it < n_chunks; ++it) {
    if (!POINTS) request_next();                              // refill the stage consumed last iteration

    if (obj != cur_obj) {
      if (cur_obj >= 0) write_part(cur_obj);
      cur_obj = obj;
      acc.clear();
      if (!POINTS) {
        const double* K = p.kinv + (p.kinv_per_object ? 9 * (size_t)obj : 0);
        g.k = K;                                              // general-K path reads K from global (L1-resident)
        g.k0 = K[0]; g.k2 = K[2]; g.k4 = K[4]; g.k5 = K[5];
        g.x0 = p.bbox[2 * (size_t)obj];
        g.y0 = p.bbox[2 * (size_t)obj + 1];
        g.simple = (K[1] == 0.0 && K[3] == 0.0 && K[6] == 0.0 && K[7] == 0.0 && K[8] == 1.0);
        fast_px = row_fast && g.simple;
        __syncwarp();
        build_ray_tables(p, g, rxc, ryr, lane, 32);
        __syncwarp();
        row = px0 / p.W;
        col = px0 - row * p.W;
      }
    }

    if (POINTS) {
      const size_t ob = (size_t)obj * p.P;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int px = px0 + j;
        if (px < p.P && p.mask[ob + px] != 0) {
          const double* s = p.src_pts + ob * 3 + px;
          const double* t = p.dst_pts + ob * 3 + px;
          acc.add(s[0], s[p.P], s[2 * (size_t)p.P], t[0], t[p.P], t[2 * (size_t)p.P]);
          ++acc.cnt;
        }
      }
    } else {
      cp_async_wait_group<DEPTH - 1>();                       // this lane's copies of this chunk have landed
      PF_CHECK(slot >= 0 && slot < DEPTH && obj < p.B && (VEC == 2 ? px0 + 3 < P : true));
      const unsigned char* st = ring + slot * kChunkBytes + lane * 16;
      uint32_t m4 = 0u;                                       // 4 mask bytes
      float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
      const bool in_obj = (VEC == 2) || px0 < P;              // VEC == 2: every chunk is full
      if (in_obj) {
        m4 = *reinterpret_cast<const uint32_t*>(ring + slot * kChunkBytes + 2048 + lane * 4);
        z4 = *reinterpret_cast<const float4*>(st + 1536);
      }
      const float zz[4] = {z4.x, z4.y, z4.z, z4.w};
      bool ok[4];
      bool any = false;
#pragma unroll
      for (int j = 0; j < 4; ++j) {                           // pose_estimation.py:23-25
        ok[j] = (m4 & (0xffu << (8 * j))) != 0u && zz[j] > 0.0f && (VEC != 0 || px0 + j < P);
        any = any || ok[j];
      }
      if (p.valid_mask != nullptr && in_obj) {
        uint8_t* vm = p.valid_mask + (size_t)obj * P + px0;
        if (VEC != 0) {
          *reinterpret_cast<uint32_t*>(vm) =
              (ok[0] ? 1u : 0u) | (ok[1] ? 0x100u : 0u) | (ok[2] ? 0x10000u : 0u) | (ok[3] ? 0x1000000u : 0u);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (px0 + j < P) vm[j] = ok[j] ? 1 : 0;
        }
      }
      if (__any_sync(0xffffffffu, any)) {
        const float4 a4 = *reinterpret_cast<const float4*>(st);
        const float4 b4 = *reinterpret_cast<const float4*>(st + 512);
        const float4 c4 = *reinterpret_cast<const float4*>(st + 1024);
        const float n0[4] = {a4.x, a4.y, a4.z, a4.w};
        const float n1[4] = {b4.x, b4.y, b4.z, b4.w};
        const float n2[4] = {c4.x, c4.y, c4.z, c4.w};
        if (fast_px) {
          // lanes past the end of the object (last, partial chunk) carry zeros: keep their table reads
          // inside the tables (with a narrow crop `row` would run far past H, out of the CTA's shared memory)
          const int trow = in_obj ? row : 0, tcol = in_obj ? col : 0;
          PF_CHECK(trow < p.H && tcol + 3 < p.W);
          const double nry = -ryr[trow];
          const double2 rxa = *reinterpret_cast<const double2*>(rxc + tcol);
          const double2 rxb = *reinterpret_cast<const double2*>(rxc + tcol + 2);
          const double rx[4] = {rxa.x, rxa.y, rxb.x, rxb.y};
          uint32_t okm[4];                                    // all ones / zero per pixel
#pragma unroll
          for (int j = 0; j < 4; ++j) okm[j] = ok[j] ? 0xffffffffu : 0u;
          acc.cnt -= (int)(okm[0] + okm[1]) + (int)(okm[2] + okm[3]);   // each valid pixel contributes -(-1)
#pragma unroll
          for (int j = 0; j < 4; ++j) {                       // branch-free: invalid pixels contribute zeros
            const double zd = (double)and_bits(zz[j], okm[j]);
            const double a0 = (double)and_bits(n0[j], okm[j]);
            const double a1 = (double)and_bits(n1[j], okm[j]);
            const double a2 = (double)and_bits(n2[j], okm[j]);
            acc.add(a0, a1, a2, rx[j] * zd, nry * zd, zd);    // y = (rx z, -ry z, [-]z), :34-41
          }
        } else {
          int r = row, cc = col;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (ok[j]) {
              double y0, y1, y2;
              backproject_px(g, rxc, ryr, r, cc, (double)zz[j], y0, y1, y2);
              acc.add((double)n0[j], (double)n1[j], (double)n2[j], y0, y1, -y2);
              ++acc.cnt;
            }
            if (++cc >= p.W) { cc = 0; ++r; }
          }
        }
      }
      row += drow;
      col += dcol;
      if (col >= p.W) { col -= p.W; ++row; }
    }
    px0 += kChunkPx;
    if (++ch == cpo) {
      ch = 0;
      px0 = 4 * lane;
      if (DYN) obj = q_obj;                                   // the request stream is inside this warp's next object
      else ++obj;
    }
    if (++slot == DEPTH) slot = 0;
  }
  write_part(cur_obj);
  PF_TRACE_END(0);
  PF_TRACE_BEGIN(5);
  PF_TRACE_WARP_END();
}

// One thread per object: merge the partial moments and solve (pose_utils.py:16-61).
//
// This kernel is pure latency: ~3 k dependent instructions executed once per thread.  What keeps it short:
//  * short batches put ceil(B / SMs) objects in a CTA (solve_object): the element-wise loads of the partial records
//    touch one line per lane, which an SM's load / store unit takes one at a time, so every SM takes a share;
//  * the partial records of an object are read four at a time (68 independent loads in flight) instead of one record
//    per round trip, in the same summation order;
//  * the pose / ctx records leave through a per-warp shared-memory tile, with consecutive addresses (write_pose);
//  * optional warm-up pass (p.prewarm, POSEFIT_PREWARM=1; the default of round 1, off now): the CTAs are scheduled while
//    K-moments still runs (programmatic dependent launch) and walk through the same solve code on synthetic moments
//    BEFORE griddepcontrol.wait, so that the real pass runs out of a warm instruction cache.
__global__ void __launch_bounds__(128) fit_solve_kernel(const FwdParams p) {
#if __CUDA_ARCH__ >= 900
  if (p.early_dep & 2) asm volatile("griddepcontrol.launch_dependents;");
#endif
  PF_TRACE_BEGIN(1);
  extern __shared__ __align__(16) double solve_tiles[];              // [warps of the CTA][kTileDoubles]
  const int o = solve_object(p.opc, p.B);
  double* tile = solve_tiles + (threadIdx.x >> 5) * kTileDoubles;
  long long row0;
  const int n_rows = solve_warp_rows(p.opc, p.B, row0);
#pragma unroll 1
  for (int pass = p.prewarm ? 0 : 1; pass < 2; ++pass) {
    double s[kAccPlain];
    if (pass == 0) {
      // a generic well-conditioned cloud: every branch of the solve is the one real data takes
#pragma unroll
      for (int i = 0; i < kAccPlain; ++i) s[i] = 0.25 * (double)(i + 1 + (threadIdx.x & 3));
      s[0] = 16.0; s[7] = 9.0; s[11] = 7.0; s[15] = 5.0; s[16] = 40.0;
    } else {
#if __CUDA_ARCH__ >= 900
      asm volatile("griddepcontrol.wait;" ::: "memory");        // K-moments has completed and flushed
      if (!(p.early_dep & 2)) asm volatile("griddepcontrol.launch_dependents;");
#endif
      PF_TRACE_BOTH(9);
      if (n_rows == 0) return;                                  // (whole warps only: the copies below are the warp's)
      // merge the partial records: four at a time (68 independent loads in flight), in order
#pragma unroll
      for (int i = 0; i < kAccPlain; ++i) s[i] = 0.0;
      if (o < p.B) {                                            // (an idle lane of a live warp stays an empty object)
        const long long c0 = (long long)o * p.chunks_per_obj;
        const long long w0 = c0 / p.chunks_per_warp, w1 = (c0 + p.chunks_per_obj - 1) / p.chunks_per_warp;
        const int n_parts = (int)(w1 - w0) + 1;
        const double* base = p.ws + (size_t)o * p.max_parts * kAccPlain;
#pragma unroll 1
        for (int k = 0; k < n_parts; k += 4) {
          double v[4][kAccPlain];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const bool on = k + u < n_parts;
            const double* part = base + (size_t)(on ? k + u : k) * kAccPlain;
#pragma unroll
            for (int i = 0; i < kAccPlain; ++i) v[u][i] = part[i];
            if (!on) {
#pragma unroll
              for (int i = 0; i < kAccPlain; ++i) v[u][i] = 0.0;
            }
          }
#pragma unroll
          for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int i = 0; i < kAccPlain; ++i) s[i] += v[u][i];
        }
      }
    }
    Moments mo;
    mo.n = s[0];
#pragma unroll
    for (int i = 0; i < 3; ++i) { mo.sx[i] = s[1 + i]; mo.sy[i] = s[4 + i]; }
#pragma unroll
    for (int i = 0; i < 9; ++i) mo.syx[i] = s[7 + i];
    mo.sxx = s[16];
#ifdef PF_TRACE
    if (pass == 1 && mo.sxx != -1.2345e300) PF_TRACE_BOTH(12);   // the merged moments are in registers
#endif
    Fit f;
    fit_from_moments<true>(mo, f);
#ifdef PF_TRACE
    if (pass == 1 && f.s != -1.2345e300) PF_TRACE_BOTH(13);      // solved, nothing written yet
#endif
    const int status = (mo.n > 0.0) ? f.status : PF_EMPTY;      // pose_estimation.py:361-362
    if (pass == 1) {
      write_pose(p, row0, n_rows, tile, f, status, mo.n, 1.0, 0.0, mo.n);
      PF_TRACE_END(1);
    } else if (f.s == -1.2345e300 && p.pose != nullptr && o < p.B) {
      p.pose[(size_t)o * POSEFIT_POSE_DOUBLES] = f.R[0] + f.t[0] + f.Linv[0] + f.H[0];   // never true: keeps the warm-up pass alive
    }
  }
}

}  // namespace posefit
