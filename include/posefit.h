/* posefit.h -- C ABI of the B200 pose solver (libposefit_b200.so).
 *
 * Drop-in boundary for the per-object 7-DoF pose fit of
 * DomiSchmauser/3D_MOT_Differentiable_Pose_Estimation.  The reference has no FFI for this
 * path -- it is a set of Python functions called once per detected instance --
 * so each entry point below cites the reference function(s) whose work it replaces; the
 * Python side that mirrors the reference's own signatures lives in
 * 3d_mot_differentiable_pose_estimation_b200/{pose_utils,pose_estimation,function}.py and the
 * binding a maintainer would add to the reference is shown in INTEGRATION.md.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no exceptions.  Every pointer is a DEVICE pointer
 *     owned by the caller (inputs, outputs and workspace); the library allocates nothing and
 *     keeps no state except a per-device attribute cache.
 *   - All calls are asynchronous on `stream` (a cudaStream_t passed as void*), never
 *     synchronise, never read device data on the host, and are CUDA-graph capturable.
 *   - Return value: 0 = ok, negative = invalid argument (POSEFIT_E_*), positive = cudaError_t.
 *   - Layout ("crop layout"): object b owns NOC planes noc[b][3][H][W] (float32, values in
 *     [0,1]; the reference's HxWx3 patch before its permute, Detection/tracker/postprocess.py:145-147),
 *     depth[b][H][W] (float32 metres, the bbox window of the frame's depth map,
 *     PoseEst/pose_estimation.py:260-262), mask[b][H][W] (uint8, the bbox window of the
 *     instance mask, :290) and bbox_xy0[b] = (x0, y0), the window's top-left pixel in the frame
 *     (int32), so pixel (i, j) of the crop is frame pixel (u, v) = (x0 + j, y0 + i).
 *   - kinv: inverse intrinsics, float64 row-major 3x3, either one matrix shared by all objects
 *     (kinv_per_object = 0) or one per object (= 1)  (np.linalg.inv(intrinsics), pose_estimation.py:22).
 *   - pose[b][16] (float64): [0] scale s, [1..9] TRUE rotation R row-major (the reference's
 *     `Rotation` is R^T, PoseEst/pose_utils.py:44), [10..12] translation t, [13] number of
 *     correspondences in the final fit, [14] inlier ratio as the reference counts it
 *     (pose_utils.py:10-12, only meaningful for the RANSAC entry), [15] pass threshold PassT.
 *   - ctx[b][32] (float64): state saved for posefit_backward (R, (tr(H)I-H)^-1, H, s, var, n,
 *     mu_x, mu_y); opaque to callers except ctx[b][30], written by the RANSAC entries: the number of
 *     iterations getRANSACInliers' loop would have run (pose_utils.py:72-82: up to and including the
 *     one whose residual falls below StopT, n_hyp when none does, 0 without correspondences) -- one
 *     np.random.randint(N, size=10) each (:73), which is what a host needs to leave the global
 *     random stream where the reference leaves it.
 *   - status[b] (int32): 0 ok; 1 no valid correspondence (run_pose returns 6xNone,
 *     pose_estimation.py:361-362); 2 inlier ratio < 0.1 (4xNone, pose_utils.py:105-107);
 *     3 NaN covariance (RuntimeError, pose_utils.py:32-36).  For status != 0 the pose is
 *     (1, I, 0) and all gradients are zero.
 */
#ifndef POSEFIT_H_
#define POSEFIT_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define POSEFIT_ABI_VERSION 1
#define POSEFIT_POSE_DOUBLES 16
#define POSEFIT_CTX_DOUBLES 32

#define POSEFIT_E_NULL (-1)       /* a required pointer is NULL */
#define POSEFIT_E_SHAPE (-2)      /* non-positive or unsupported size */
#define POSEFIT_E_WORKSPACE (-3)  /* workspace too small */
#define POSEFIT_E_SMEM (-4)       /* crop does not fit the shared-memory staging of the RANSAC path */

/* Flag bits of the RANSAC entries' `ref_compat` argument. */
#define POSEFIT_REF_COMPAT 1          /* reproduce the reference's scoring quirks (see posefit_forward_ransac) */
#define POSEFIT_SAMPLES_ARE_BITS 2    /* sample_idx holds uniform 32-bit values u (drawn on the device, before the number of
                                         correspondences N is known); hypothesis sample = floor(u * N / 2^32) */

/* ABI version of the loaded library. */
int posefit_version(void);

/* Human-readable text for a return code of this library (static storage). */
const char* posefit_error_string(int code);

/* Bytes of device scratch the forward entries need for these sizes (n_hyp = 0 for the
 * plain fit).  May be 0.  The scratch needs no initialisation and carries nothing from call to call
 * (partial moments, per-object records and -- long batches -- a ticket counter that the entry itself
 * zeroes); two calls in flight at the same time need two workspaces. */
size_t posefit_workspace_bytes(int n_objects, int height, int width, int n_hyp, int n_samp);

/* Plain fit on all valid correspondences of every object.
 * Replaces, per object: the padding + backproject + NOC gather of run_pose
 * (PoseEst/pose_estimation.py:256-290, :323 and backproject :16-43) followed by
 * estimateSimilarityUmeyama (PoseEst/pose_utils.py:16-61).  n_valid[b] receives the number of
 * correspondences (mask != 0 and depth > 0). */
int posefit_forward(const float* noc, const float* depth, const uint8_t* mask, const int32_t* bbox_xy0,
                    const double* kinv, int kinv_per_object, int n_objects, int height, int width,
                    double* pose, double* ctx, int32_t* status, int32_t* n_valid,
                    void* workspace, size_t workspace_bytes, void* stream);

/* RANSAC fit with host-supplied sample indices.
 * Replaces, per object: the same front end, then estimateSimilarityTransform
 * (PoseEst/pose_utils.py:86-117): PassT/StopT heuristics (:91-96), getRANSACInliers (:63-83)
 * with hypothesis h fitted by estimateSimilarityUmeyama on the n_samp correspondences
 * sample_idx[b][h][:] (indices into the row-major list of valid pixels, replacing
 * np.random.randint at :73; values >= n_valid[b] are clamped), scored as evaluateModel does
 * (:5-14), first-minimum selection with early stop (:76-81), the ratio gate (:105-107) and the
 * refit on the winner's inliers (:109).
 * ref_compat bit 0 (POSEFIT_REF_COMPAT) set reproduces the reference exactly: hypotheses are scored with the
 * transposed rotation block the reference builds (:58) and point 0 is never counted as an inlier (:11);
 * clear: scores with s*R and counts every inlier.  Bit 1 (POSEFIT_SAMPLES_ARE_BITS): sample_idx holds uniform
 * 32-bit values instead of indices (device-side draws: no host round trip for the correspondence counts).
 * inlier_mask[b][H][W] (uint8) receives the winner's inlier set (all valid pixels when no
 * hypothesis was accepted); winner[b] (optional, may be NULL) the winning hypothesis or -1. */
int posefit_forward_ransac(const float* noc, const float* depth, const uint8_t* mask, const int32_t* bbox_xy0,
                           const double* kinv, int kinv_per_object, const int32_t* sample_idx,
                           int n_objects, int height, int width, int n_hyp, int n_samp,
                           double ratio_adapt, int ref_compat,
                           double* pose, double* ctx, int32_t* status, int32_t* n_valid,
                           uint8_t* inlier_mask, int32_t* winner,
                           void* workspace, size_t workspace_bytes, void* stream);

/* The same two fits with the extra outputs the autograd operator wants written by the kernels themselves (no
 * separate conversion / masking passes over the batch):  scale_f32 [B], rot_f32 [B][9], trans_f32 [B][3] -- the
 * pose rounded to float32 by the solve kernel -- and, for the plain fit, valid_mask [B][H][W] (uint8): 1 where the
 * pixel took part in the fit (mask != 0 and depth > 0, PoseEst/pose_estimation.py:23-25), i.e. what inlier_mask is for
 * the RANSAC entry.  Each of them may be NULL. */
int posefit_forward_ex(const float* noc, const float* depth, const uint8_t* mask, const int32_t* bbox_xy0,
                       const double* kinv, int kinv_per_object, int n_objects, int height, int width,
                       double* pose, double* ctx, int32_t* status, int32_t* n_valid,
                       float* scale_f32, float* rot_f32, float* trans_f32, uint8_t* valid_mask,
                       void* workspace, size_t workspace_bytes, void* stream);

int posefit_forward_ransac_ex(const float* noc, const float* depth, const uint8_t* mask, const int32_t* bbox_xy0,
                              const double* kinv, int kinv_per_object, const int32_t* sample_idx,
                              int n_objects, int height, int width, int n_hyp, int n_samp,
                              double ratio_adapt, int ref_compat,
                              double* pose, double* ctx, int32_t* status, int32_t* n_valid,
                              uint8_t* inlier_mask, int32_t* winner,
                              float* scale_f32, float* rot_f32, float* trans_f32,
                              void* workspace, size_t workspace_bytes, void* stream);

/* The plain fit and its gradient fed by the NOC HEAD OUTPUT instead of a materialised NOC crop.
 * Replaces, per object: the per-instance roi_align resize of the 3 x 28 x 28 head output to the instance's box
 * (Detection/tracker/postprocess.py:141-147, head: Detection/roi_heads/nocs_head.py:232-235) AND the fit it feeds
 * (posefit_forward): head[b][3][head_h][head_w] (float32) is sampled on the fly with torchvision's roi_align taps
 * (aligned, whole map, output size roi_hw[b] = (h_b, w_b), zero padding to height x width) -- bit-identical values to
 * posefit_resample_noc -- so no h x w x 3 patch is ever written or read.  depth / mask / bbox_xy0 / kinv and all
 * outputs as for posefit_forward_ex.  posefit_backward_head is the matching adjoint: gradient w.r.t. the head output
 * (grad_head [B][3][head_h][head_w], fully written) and optionally the depth crop. */
size_t posefit_head_workspace_bytes(int n_objects);
int posefit_forward_head(const float* head, const int32_t* roi_hw, const float* depth, const uint8_t* mask,
                         const int32_t* bbox_xy0, const double* kinv, int kinv_per_object,
                         int n_objects, int head_h, int head_w, int height, int width,
                         double* pose, double* ctx, int32_t* status, int32_t* n_valid,
                         float* scale_f32, float* rot_f32, float* trans_f32,
                         void* workspace, size_t workspace_bytes, void* stream);
int posefit_backward_head(const float* head, const int32_t* roi_hw, const float* depth, const uint8_t* mask,
                          const uint8_t* inlier_mask, const int32_t* bbox_xy0, const double* kinv, int kinv_per_object,
                          int n_objects, int head_h, int head_w, int height, int width,
                          const double* ctx, const int32_t* status,
                          const float* grad_scale, const float* grad_R, const float* grad_t,
                          float* grad_head, float* grad_depth,
                          void* workspace, size_t workspace_bytes, void* stream);

/* Points mode: the same two fits on explicit correspondences, for callers that hold point sets
 * rather than crops -- the argument form of estimateSimilarityUmeyama / estimateSimilarityTransform
 * themselves (PoseEst/pose_utils.py:16, :86; run_pose calls them on filtered clouds,
 * PoseEst/pose_estimation.py:363).  src[b][3][N] is the source cloud (NOC - 0.5), dst[b][3][N]
 * the target cloud, both float64 planar (rows 0..2 of the reference's homogeneous [4,N] arrays);
 * mask[b][N] (uint8) selects the points that exist (objects are padded to a common N).
 * sample_idx indexes the list of selected points.  Outputs as for the crop entries;
 * inlier_mask is [b][N].  pass_threshold / stop_threshold > 0 replace the data-derived PassT /
 * StopT (the explicit arguments of getRANSACInliers, pose_utils.py:63); pass <= 0 to derive them. */
int posefit_points_forward(const double* src, const double* dst, const uint8_t* mask, int n_objects, int n_points,
                           double* pose, double* ctx, int32_t* status, int32_t* n_valid,
                           void* workspace, size_t workspace_bytes, void* stream);

int posefit_points_forward_ransac(const double* src, const double* dst, const uint8_t* mask,
                                  const int32_t* sample_idx, int n_objects, int n_points, int n_hyp, int n_samp,
                                  double ratio_adapt, double pass_threshold, double stop_threshold,
                                  int ref_compat,
                                  double* pose, double* ctx, int32_t* status, int32_t* n_valid,
                                  uint8_t* inlier_mask, int32_t* winner,
                                  void* workspace, size_t workspace_bytes, void* stream);

/* Gradient of  sum_b ( grad_scale[b]*s + <grad_R[b], R> + <grad_t[b], t> )  with respect to
 * the NOC crop (and, when grad_depth != NULL, the depth crop).  No reference function exists:
 * the upstream code detaches before the fit (Detection/tracker/postprocess.py:151).
 * inlier_mask may be NULL (plain fit: every valid correspondence has weight 1), otherwise it is
 * the mask written by posefit_forward_ransac (selection is treated as a constant).
 * grad_scale [B], grad_R [B][9] (w.r.t. the TRUE rotation), grad_t [B][3] are float32 and may
 * each be NULL (= zeros).  grad_noc [B][3][H][W] and grad_depth [B][H][W] are fully written. */
int posefit_backward(const float* noc, const float* depth, const uint8_t* mask, const uint8_t* inlier_mask,
                     const int32_t* bbox_xy0, const double* kinv, int kinv_per_object,
                     int n_objects, int height, int width,
                     const double* ctx, const int32_t* status,
                     const float* grad_scale, const float* grad_R, const float* grad_t,
                     float* grad_noc, float* grad_depth,
                     void* workspace, size_t workspace_bytes, void* stream);

/* Bytes of device scratch posefit_backward needs (per-object adjoint coefficients, 16-byte aligned). */
size_t posefit_backward_workspace_bytes(int n_objects);

/* Stable row-major compaction of every crop: the correspondences exactly as the reference builds
 * them before the fit -- `backproject` (PoseEst/pose_estimation.py:16-43: points and the np.where
 * index arrays) and the NOC gather of run_pose (:323).  Frame-level use: one "crop" = the whole
 * frame, bbox (0,0).  dst[b][P][3] / src[b][P][3] (float64, interleaved like the reference's [N,3]
 * arrays; src and noc may be NULL), rows/cols[b][P] frame coordinates (int32), count[b].  Only the
 * first count[b] entries of each object are written. */
int posefit_compact(const float* noc, const float* depth, const uint8_t* mask, const int32_t* bbox_xy0,
                    const double* kinv, int kinv_per_object, int n_objects, int height, int width,
                    double* src, double* dst, int32_t* rows, int32_t* cols, int32_t* count, void* stream);

/* evaluateModel (PoseEst/pose_utils.py:5-14) for one explicit 4x4 row-major transform per object on
 * points-mode inputs.  stats[b][4] = { Residual, number of inliers, 1 if the first selected point is
 * an inlier (the reference's count_nonzero skips it, :11), number of selected points };
 * inlier_mask[b][N].  pass_threshold: one value (pass_per_object = 0) or one per object. */
int posefit_points_evaluate(const double* transform, const double* src, const double* dst, const uint8_t* mask,
                            const double* pass_threshold, int pass_per_object, int n_objects, int n_points,
                            double* stats, uint8_t* inlier_mask, void* stream);

/* out = A p + t on interleaved float64 [N][3] points, matrix = [A | t] row-major 3x4 (one, or one
 * per object): transform_pc (PoseEst/pose_estimation.py:45-57) and cam2world (:59-70). */
int posefit_transform_points(const double* matrix, int matrix_per_object, const double* points, double* out,
                             int n_objects, int n_points, void* stream);

/* Batched tail of run_pose (PoseEst/pose_estimation.py:367-412) plus the Euler conversion of its
 * callers (Detection/tracker/postprocess.py:158-160): for every object
 *   out[b][0..8]   global_rot  = (campose @ [s R | t])[:3,:3]   (scale embedded, :404-406)
 *   out[b][9..11]  global_trans
 *   out[b][12]     global_scale = s
 *   out[b][13..15] XYZ Euler angles of global_rot / column norms (get_scale + mathutils to_euler)
 *   out[b][16..39] world-space axis-aligned box of the object's depth points, 8 corners in the
 *                  reference's sort_bbox order (:72-93, :373-380); zeros when status != 0.
 * campose: n_campose row-major 4x4 camera-to-world matrices; cam_index[b] picks one per object
 * (NULL: n_campose == 1 -> shared, else one per object); campose == NULL keeps camera space
 * (run_pose_office, :501-512). */
int posefit_epilogue(const float* depth, const uint8_t* mask, const int32_t* bbox_xy0, const double* kinv,
                     int kinv_per_object, const double* pose, const int32_t* status, const double* campose,
                     int n_campose, const int32_t* cam_index, int n_objects, int height, int width, double* out,
                     void* stream);

/* GT-box pre-filter of run_pose: clean_depth (PoseEst/pose_estimation.py:107-134) applied as at
 * :293-299.  out_mask[b] = mask & depth>0 & (world-space depth point strictly inside the axis-aligned
 * extent of gt_box[b] (8x3 float64 corners)) when more than min_keep (reference: 20) points survive,
 * else the unclipped validity mask.  campose / cam_index as for posefit_epilogue (campose required).
 * kept[b] (optional) receives the number of surviving correspondences.  Feed out_mask to the fit
 * entries in place of mask. */
int posefit_clip_mask(const float* depth, const uint8_t* mask, const int32_t* bbox_xy0, const double* kinv,
                      int kinv_per_object, const double* campose, int n_campose, const int32_t* cam_index,
                      const double* gt_box, int min_keep, int n_objects, int height, int width,
                      uint8_t* out_mask, int32_t* kept, void* stream);

/* Statistical outlier removal as a mask filter: the Open3D remove_statistical_outlier(20, 2.0)
 * passes of run_pose (PoseEst/pose_estimation.py:311-318 on the depth cloud = source 0, :341-349 on
 * the NOC cloud = source 1; applied only when the cloud has more than min_points = 100 points).
 * UNPINNED: open3d==0.10.0.0 is not vendored; the semantics are restated from Open3D's
 * PointCloud::RemoveStatisticalOutliers (mean distance to the nb_neighbors nearest points, the query
 * itself included; threshold = mean + std_ratio * sample std).  nb_neighbors must be 20.
 * out_mask[b] = the surviving subset of (mask & depth > 0); chain two calls (depth, then NOC on the
 * first call's out_mask) to reproduce run_pose. */
size_t posefit_sor_workspace_bytes(int n_objects, int height, int width);
int posefit_sor_mask(const float* noc, const float* depth, const uint8_t* mask, const int32_t* bbox_xy0,
                     const double* kinv, int kinv_per_object, int source, int nb_neighbors, double std_ratio,
                     int min_points, int n_objects, int height, int width, uint8_t* out_mask,
                     void* workspace, size_t workspace_bytes, void* stream);

/* Batched front end of the per-instance loop (Detection/tracker/postprocess.py:131-152).
 *
 * posefit_resample_noc: the ROI-align resize of the NOC head output head[b][3][head_h][head_w]
 * (3 x 28 x 28, Detection/roi_heads/nocs_head.py:232-235) to each instance's box size
 * roi_hw[b] = (h_b, w_b) (postprocess.py:141-147: roi_align over the whole map, aligned=True,
 * sampling_ratio -1), written zero-padded into noc[b][3][H][W].  Same arithmetic as
 * torchvision.ops.roi_align, which detectron2.layers.roi_align wraps.
 * posefit_resample_noc_backward: its adjoint, grad_head[b][3][head_h][head_w] (fully written).
 * posefit_gather_crops: depth[b][H][W] / mask[b][H][W] = the bbox window of frame frame_of[b] of
 * depth_frames[F][FH][FW] and of the instance's full-frame mask mask_frames[b][FH][FW]
 * (PoseEst/pose_estimation.py:260-262, :290), zero outside the box; bbox_xyxy[b] = integer
 * (x0, y0, x1, y1), upper corner exclusive; also writes bbox_xy0[b] and roi_hw[b]. */
int posefit_resample_noc(const float* head, const int32_t* roi_hw, int n_objects, int head_h, int head_w,
                         int height, int width, float* noc, void* stream);
int posefit_resample_noc_backward(const float* grad_noc, const int32_t* roi_hw, int n_objects, int head_h,
                                  int head_w, int height, int width, float* grad_head, void* stream);
int posefit_gather_crops(const float* depth_frames, const uint8_t* mask_frames, const int32_t* frame_of,
                         const int32_t* bbox_xyxy, int n_objects, int frame_h, int frame_w, int height, int width,
                         float* depth, uint8_t* mask, int32_t* bbox_xy0, int32_t* roi_hw, void* stream);

/* Instance masks as they cross PCIe: one bit per pixel.  The masks the pose path receives are boolean
 * (Detection/tracker/postprocess.py:134-139; PoseEst/pose_estimation.py:23-25 only tests them for truth), so a host
 * that stages them may pack them -- numpy.packbits(mask, bitorder='little') over the flattened [B][H][W] array: pixel i
 * is bit (i & 7) of byte (i >> 3) -- and expand them on the device: mask[i] = 0 / 1 for i < n_pixels. */
int posefit_unpack_mask(const uint8_t* bits, long long n_pixels, uint8_t* mask, void* stream);

/* Tracker graph edges from pose tensors: GraphDataset.get_edge_data and
 * get_edge_data_office (Tracking/datasets/graph_dataset.py:30-199, :232-330), batched over
 * n_sequences sequences of n_frames frames each.  Nodes are the detections in (sequence, frame)
 * order; frame_start[s * n_frames + f] is the index of the first node of frame f of sequence s
 * (n_sequences * n_frames + 1 entries).  translations / rotations (XYZ Euler) [N][3] and
 * scales [N][scale_dim] are float64, the per-frame hdf5 records of
 * Detection/inference_detector.py:352-371 concatenated as Tracking/mpn_trainer.py:440-442 does.
 * node_id[n] = what train_utils.check_pair returned for the node (ground-truth object id, < 0 for
 * None); NULL keeps every pair (the _office variant).  Candidate pairs (n in frame t, m in frame
 * t+1 .. t+max_frame_dist, frame < min(max_seq_len, n_frames)) are emitted in the reference's loop
 * order:  edge_index[0][e], edge_index[1][e] (row stride max_edges, node indices local to the
 * sequence), edge_attr[e] = { t_m - t_n (3), euler_m - euler_n (3), log(scale_m / scale_n)
 * (scale_dim), frame distance } as float32 (float64 arithmetic, rounded once, :166-199),
 * targets[e] = (id_n == id_m), consecutive[e] = (frame == t + 1), edge_seq[e] = sequence.
 * totals[0] = number of (directed) edges, totals[1] = the reference's false_positives count
 * (:95-96, :133-136).  Edges beyond max_edges are dropped (totals still counts them).
 * targets / consecutive / edge_seq may be NULL.  The undirected duplication of :203-206 is a plain
 * concatenation and is left to the caller. */
size_t posefit_edge_workspace_bytes(int n_sequences, int n_frames, int n_nodes, int max_frame_dist);
int posefit_edge_features(const double* translations, const double* rotations, const double* scales, int scale_dim,
                          const int32_t* frame_start, const int32_t* node_id, int n_sequences, int n_frames,
                          int n_nodes, int max_frame_dist, int max_seq_len, long long max_edges,
                          long long* edge_index, float* edge_attr, float* targets, int8_t* consecutive,
                          int32_t* edge_seq, long long* totals, void* workspace, size_t workspace_bytes,
                          void* stream);

/* Number of kernels this library has launched in the calling process (for bench.py's
 * gpu_launches claim). */
unsigned long long posefit_launch_count(void);

/* The POSEFIT_* launch knobs are read from the environment once, when the library is
 * loaded.  Tests and tuning tools that change one afterwards call this to re-read them; the launch
 * path itself never touches the environment. */
void posefit_debug_reload_env(void);

#ifdef __cplusplus
}
#endif
#endif /* POSEFIT_H_ */
