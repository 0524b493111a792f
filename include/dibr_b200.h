/*
 * dibr_b200.h -- C ABI of libdibr_b200.so, the B200 (sm_100a) implementation of the DIB-R
 * differentiable rasterizer on Self6D++'s render-and-compare path.
 *
 * What each entry point replaces in the reference (paths under /root/reference):
 *
 *   dibr_setup_faces      lib/dr_utils/dib_renderer_x/rasterizer/rasterizer.py:36-70  (prepare_tfpoints:
 *                         x multiplier, per-face bbox, expanded bbox) at the linear_rasterizer seam.
 *   dibr_setup_meshes     renderer/vertex_shaders/perpsective.py:71-111 (view transform, 4x4 projection,
 *                         divide, per-face gather, face normal) + renderer/vcrender_batch.py:49-88
 *                         (per-sample Python loop, attribute gather + ones channel) + prepare_tfpoints,
 *                         for a whole ragged batch in one launch.
 *   dibr_forward          kaolin.graphics.dib_renderer.cuda.rasterizer.forward as called at
 *                         rasterizer.py:152-172 (kernels dr_cuda_forward_render_batch and
 *                         dr_cuda_forward_prob_batch of the un-vendored kaolin v0.1).
 *   dibr_backward_faces   kaolin...rasterizer.backward as called at rasterizer.py:249-269
 *                         (dr_cuda_backward_color_batch + dr_cuda_backward_prob_batch), returning
 *                         dldp2 + dldp2_prob and dldc exactly like rasterizer.py:278-291 -- but as a
 *                         deterministic per-face gather instead of fp32 global atomics.
 *   dibr_backward_meshes  the torch autograd tail the reference runs after the extension returns:
 *                         index_select/cat backward (scatter-add to vertices), division and matmul
 *                         backward down to cam_view_R / cam_view_pos (renderer/base.py:169-170).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless it says host; fp32 contiguous; the library never
 *     allocates, frees or retains caller memory (the reference's Python owns every buffer too,
 *     rasterizer.py:124-148,198-211) and never synchronises the device.
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*).
 *   - return 0 on success, non-zero on error; dibr_last_error() gives a thread-local message.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point returns an error.
 *   - one DibrPass describes one rasterisation: set-up, forward and backward must be called with the SAME sizes
 *     and scalar parameters (the set-up kernel already bins the faces into screen tiles with the expanded
 *     bboxes, so `expand`, `multiplier`, `height`, `width` are baked into the workspace).
 *   - limits: image sides <= 16384, batch <= 4096, num_attr <= DIBR_MAX_ATTR, knum <= 250; the workspace
 *     (dibr_workspace_bytes) grows with tiles x faces / 8 bytes for the tile bitmaps.
 *
 * Image/face layout
 *   images b = 0..batch-1; image b owns faces [face_offsets[b], face_offsets[b+1]) of one global
 *   face array (total_faces entries).  If face_offsets is NULL every image owns faces_per_image
 *   faces (the reference's dense b x f layout).  Face index buffers store the face number LOCAL to
 *   the image plus one (0 = no face), like the reference's imidx (rasterizer.py:124).
 */
#ifndef DIBR_B200_H
#define DIBR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DIBR_ABI_VERSION 3
#define DIBR_MAX_ATTR 12 /* interpolated channels per pixel (the reference uses 4: rgb/xyz/normal + ones) */
#define DIBR_MAX_OUTPUTS 6

typedef struct DibrPass {
    /* ---- problem size ------------------------------------------------------------------ */
    int32_t batch;           /* images */
    int32_t height, width;   /* pixels */
    int32_t num_attr;        /* D: channels per vertex in face_attr (3*D floats per face) */
    int32_t knum;            /* soft-silhouette cap K (reference default 30) */
    int32_t multiplier;      /* reference default 1000 (passed as a C int, rasterizer.py:170) */
    int32_t delta;           /* reference default 7000 ("sigmainv") */
    float expand;            /* reference default 0.02 */
    int32_t total_faces;     /* faces over all images.  With face_offsets this is the CAPACITY of the per-face arrays (and fixes the
                                workspace layout); the faces in use are face_offsets[batch] <= total_faces, read on the device */
    int32_t faces_per_image; /* used when face_offsets == NULL */
    const int32_t *face_offsets; /* [batch+1] or NULL */

    /* ---- operator-seam inputs (dibr_setup_faces) ------------------------------------------ */
    const float *points3d;   /* [total_faces, 9] view-space xyz of the 3 corners (only z is read) */
    const float *points2d;   /* [total_faces, 6] NDC xy of the 3 corners, NOT yet multiplied */
    const float *normalz;    /* [total_faces]    z of the face normal; < 0 = back face */

    /* ---- fused inputs (dibr_setup_meshes / dibr_backward_meshes) ---------------------------- */
    int32_t num_instances;          /* objects with a pose; the batch modes have one per image */
    const int32_t *inst_desc;       /* [num_instances, 12]: vert_base, num_verts, mesh_face_base, num_faces,
                                       out_face_base (ascending), cam_index, proj_index, attr_base,
                                       gvert_base (row of this instance in grad_verts / grad_vert_attr),
                                       image index, adj_base (row of the mesh's vertex 0 in vert_face_ptr), 1 reserved */
    const float *verts;             /* [sum verts, 3] object-space vertices (packed meshes) */
    const int32_t *mesh_faces;      /* [sum mesh faces, 3] vertex ids local to the mesh */
    const float *vert_attr;         /* [sum verts(attr space), vert_attr_dim] per-vertex attributes */
    int32_t vert_attr_dim;          /* channels read per vertex */
    int32_t attr_flags;             /* bit0: append a ones channel (vcrender_batch.py:87-88);
                                       bit1: append view depth -z_view (= z of R v + t,
                                       renderer_dibr.py:296-301) as one more channel;
                                       bit2: grad_face_attr is SCRATCH between dibr_backward_faces and dibr_backward_meshes
                                       (the caller never reads it; ignored when grad_vert_attr is set): only the depth
                                       channel's corner gradients are kept, as [total_faces, 3] at the start of the buffer */
    const float *cam_rot;           /* [ncam, 9]  cam_view_R = diag(1,-1,-1) R   (base.py:169) */
    const float *cam_pos;           /* [ncam, 3]  cam_view_pos = -(R^T t)        (base.py:170) */
    const float *cam_proj;          /* [nproj, 16] row-major 4x4 used as [p,1] @ proj (perspective.py:122-129) */

    /* ---- pose mode (optional): cameras given as object pose + intrinsics --------------------------
     * When pose_R is non-NULL the library derives cam_rot / cam_pos / cam_proj itself (into the workspace),
     * restating renderer/base.py:131-191 + utils/perspective.py:95-130 in one tiny kernel, and
     * dibr_backward_meshes chains the gradients down to R and t.  cam_index of instance i is i and its
     * proj_index column selects the row of pose_K. */
    const float *pose_R;            /* [num_instances, 9] rotation object -> camera (OpenCV convention) */
    const float *pose_t;            /* [num_instances, 3] */
    const float *pose_K;            /* [num_K, 9] intrinsics */
    int32_t num_K;
    float znear, zfar;              /* only shape the unused clip-z column (perspective.py:98-99) */
    float *grad_pose_R;             /* [num_instances, 9] */
    float *grad_pose_t;             /* [num_instances, 3] */

    /* ---- workspace written by set-up, read by forward/backward (sizes: dibr_workspace_bytes) - */
    void *workspace;
    size_t workspace_bytes;
    float *face_attr;        /* [total_faces, 3, D] per-face corner attributes.  Seam mode: the caller's
                                vertex_attr_bxfx3d (read-only).  Fused mode: written by set-up. */
    float *face_normal;      /* optional [total_faces, 3]: normalised face normal (vcrender_batch.py:79) */

    /* ---- forward outputs ----------------------------------------------------------------- */
    float *im;               /* [batch, H, W, D] interpolated attributes (0 where uncovered) */
    float *improb;           /* [batch, H, W]    soft silhouette probability */
    int32_t *imidx;          /* [batch, H, W]    >0: local face id + 1 of the covering face;
                                                 0: uncovered, every near face counted;
                                                 <0: uncovered, only faces with id+1 <= -value counted
                                                 (the K-th accepted face: first-K-in-index-order rule) */
    float *imcomp;           /* [batch, H, W]    prod_k (1 - p_k) over the accepted faces of an uncovered pixel
                                                 (= 1 - improb, kept separately at full relative precision;
                                                 saved for the backward, 0 where covered).  Written on the 16x16 tiles
                                                 some face's expanded bbox reaches -- the only pixels the backward reads;
                                                 NOT written elsewhere (where the empty product would be 1) */

    /* ---- backward inputs / outputs --------------------------------------------------------- */
    const float *grad_im;     /* [batch, H, W, D] or NULL */
    const float *grad_improb; /* [batch, H, W]    or NULL */
    float *grad_points2d;     /* [total_faces, 6]  dL/d points2d (un-multiplied NDC), colour + soft parts */
    float *grad_face_attr;    /* [total_faces, 3, D] */
    float *grad_verts;        /* optional [sum over instances of num_verts, 3] (fused), row gvert_base + v */
    float *grad_vert_attr;    /* optional [sum over instances of num_verts, vert_attr_dim] (fused) */
    float *grad_cam_rot;      /* [num_instances, 9] (fused): dL/d cam_view_R of the instance's camera */
    float *grad_cam_pos;      /* [num_instances, 3] (fused) */
    /* ---- optional split of the D channels over several output tensors --------------------- */
    int32_t num_outputs;                       /* 0: one tensor `im` / `grad_im` with D channels (operator seam).
                                                  n>0: channel groups out_channels[0..n) (sum = D) go to separate
                                                  tensors out[g] of shape [batch,H,W,out_channels[g]]; grad_out[g]
                                                  may be NULL (= no gradient for that group). */
    int32_t out_channels[DIBR_MAX_OUTPUTS];
    float *out[DIBR_MAX_OUTPUTS];
    const float *grad_out[DIBR_MAX_OUTPUTS];
    /* ---- optional batch-global minimum of one output group (renderer_dibr.py:284 `_ren_norms.min()`) ---- */
    int32_t min_output;                        /* index of the output group to minimise over, or -1 */
    uint32_t *out_min_ordered;                 /* [1] order-preserving uint32 encoding of the float minimum over
                                                  ALL pixels of that group (uncovered pixels count as 0); reset by
                                                  the set-up call, decoded by dibr_normal_map */
    const int32_t *vert_face_ptr;  /* CSR vertex -> incident (face,corner) list, per packed mesh vertex: [sum verts + 1] */
    const int32_t *vert_face_idx;  /* [3 * sum mesh faces] entries face*3+corner, ascending */
    int32_t num_cams;
    int32_t verts_stride;          /* floats per row of verts: 0 or 3 = packed [.,3]; 4 = rows padded to 16 B (one 128-bit gather per vertex) */
    int32_t vert_attr_stride;      /* floats per row of vert_attr: 0 = vert_attr_dim; a multiple of 4 (and a 16 B aligned base) turns the
                                      per-corner attribute gather into 128-bit loads */
    int32_t reserved0;
} DibrPass;

int dibr_abi_version(void);
/* sizeof(DibrPass) as compiled, so a binding can verify its struct mirror */
int dibr_sizeof_pass(void);
int dibr_sizeof_step(void);
const char *dibr_last_error(void);

/* number of CUDA devices visible to the library (0 when there is no driver/GPU) */
int dibr_device_count(void);

/* bytes of workspace needed for a pass of this size (host-side arithmetic only) */
int dibr_workspace_bytes(const DibrPass *pass, size_t *bytes);

int dibr_setup_faces(const DibrPass *pass, void *stream);
int dibr_setup_meshes(const DibrPass *pass, void *stream);
int dibr_forward(const DibrPass *pass, void *stream);
int dibr_backward_faces(const DibrPass *pass, void *stream);
int dibr_backward_meshes(const DibrPass *pass, void *stream);

/* normal-map post-processing of Renderer_dibr.render_batch(mode=["norm"]) (renderer_dibr.py:281-286):
 *   out = (n - min) / (||n - min||_2 + 1e-5) * mask      per pixel, n = interpolated vertex normal (3 channels),
 * min = the batch-global minimum accumulated by dibr_forward into min_ordered. */
int dibr_normal_map(const float *normals_nx3, const float *mask_n, const uint32_t *min_ordered, float *out_nx3,
                    long long num_pixels, void *stream);
/* The same map over a pass that dibr_forward has rasterised: only the tiles some face reaches are read and normalised (the
 * plan of the pass lists them), the others are zero-filled -- or left alone when out_nx3 == normals_nx3 (in place over the
 * forward's own zero fill).  normals / mask are output groups of the pass, the minimum is pass->out_min_ordered. */
int dibr_normal_map_pass(const DibrPass *pass, const float *normals_nx3, const float *mask_n, float *out_nx3, void *stream);

/* ---------------------------------------------------------------------------------------------------------------
 * One render-and-compare step of Self6D++'s compute_self_loss_pose (core/self6dpp/engine/self_engine_utils.py:426-447)
 * from HOST buffers, enqueued with a single call:
 *   1. one H2D copy of the packed per-step inputs (poses, intrinsics, instance table) from pinned host memory,
 *   2. student pass  : dibr_setup_meshes + dibr_forward (+ dibr_normal_map when student_normal_* are set),
 *   3. teacher pass  : the same for the pseudo-label pose (skipped when teacher.num_instances == 0),
 *   4. backward      : dibr_backward_faces + dibr_backward_meshes of the student pass with the caller's upstream
 *                      gradients (student.grad_out[] / grad_improb), skipped when run_backward == 0,
 *   5. one D2H copy of dL/dR, dL/dt (num_instances x 12 floats) into pinned host memory.
 * Both passes are ordinary pose-mode DibrPass structs whose DEVICE pointers for pose_R / pose_t / pose_K /
 * inst_desc / face_offsets point INTO `staging_device`, laid out exactly like `staging_host`.
 * Nothing is synchronised: the caller waits on the stream before reading host_grad_pose.
 * The two host transfers are a few KB each.  When staging_host / host_grad_pose are pinned AND mapped (any cudaHostAlloc'd
 * buffer under unified addressing) they cross the bus without a copy-engine node: a small kernel reads the staging block, and
 * the kernel that finalises the pose gradients stores them into host_grad_pose itself (e2e 0.296 -> 0.281 ms per step);
 * any other host pointer goes through cudaMemcpyAsync. */
typedef struct DibrStep {
    DibrPass student;
    DibrPass teacher;
    const void *staging_host;      /* pinned */
    void *staging_device;
    size_t staging_bytes;
    /* optional normal-map post-processing (renderer_dibr.py:281-286); inputs are output groups of the pass */
    const float *student_normal_in, *student_mask_in; float *student_normal_out;
    const float *teacher_normal_in, *teacher_mask_in; float *teacher_normal_out;
    int32_t run_backward;          /* bit 0 (dibr_render_step): run the backward after the forward.
                                      bit 1: the forward call also PREPARES the backward of the student pass (zeroes
                                      grad_points2d / the scratch grad_face_attr and builds the work lists, in the shadow of the
                                      teacher rasterisation); dibr_render_backward / the backward half of dibr_render_step with
                                      bit 1 set rely on that and must be the FIRST backward after such a forward -- clear the bit
                                      for a repeated backward over the same forward (it then prepares itself) */
    int32_t grad_pose_sum;         /* != 0: device_grad_pose has num_instances + 1 rows and the last one receives the column sums
                                      of the others (added in instance order by the backward itself): the 12-float vector a
                                      data-parallel step all-reduces over NCCL, ready without another launch */
    float *host_grad_pose;         /* pinned [num_instances, 12]: 9 of dL/dR then 3 of dL/dt per instance, or NULL */
    float *device_grad_pose;       /* [num_instances, 12] packed dL/dR | dL/dt on the device (always written when non-NULL);
                                      the D2H copy reads from it */
    void *overlap;                 /* handle from dibr_overlap_create (a side stream + two events owned by the CALLER, one per
                                      session / host thread), or NULL: both passes run one after the other in `stream` */
} DibrStep;

/* Side stream + fork/join events that let the teacher rasterisation run next to the student chain.  One handle per
 * session (or per host thread that steps on a device): the library keeps no stream or event of its own, so two
 * sessions stepping concurrently never share one.  Create and destroy with the session's device current. */
int dibr_overlap_create(void **handle);
int dibr_overlap_destroy(void *handle);

/* The step in the two halves a training loop needs (the upstream gradients of the backward depend on the forward's
 * images, self_engine_utils.py:541-558, 736-813):
 *   dibr_render_forward   steps 1-3 above (H2D, student and teacher rasterisation, normal maps),
 *   dibr_render_backward  steps 4-5 (backward of the student pass with student.grad_out[] / grad_improb, D2H of the
 *                         pose gradients) on the forward's saved buffers; may be called again with other gradients.
 * dibr_render_step = dibr_render_forward, then dibr_render_backward when run_backward != 0 (the student chain stays on
 * the side stream throughout).  On every path, error returns included, `stream` has been made to wait for whatever was
 * enqueued on the side stream. */
int dibr_render_forward(const DibrStep *step, void *stream);
int dibr_render_backward(const DibrStep *step, void *stream);
int dibr_render_step(const DibrStep *step, void *stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Chamfer nearest-neighbour distance -- replaces the reference's torch_nndistance extension
 * (core/csrc/torch_nndistance/src/nnd_cuda.cpp:86-89 nnd_forward_cuda / nnd_backward_cuda; CPU twin nnd_cpu.cpp).
 * Clouds are padded: sample b of cloud i occupies rows [b*stride_i, b*stride_i + count_i[b]) of xyz_i ([batch*stride_i, 3]);
 * count_i == NULL means every row is a point (the reference's dense [b, n, 3] layout).  dist = squared distance to the
 * nearest point of the other cloud, idx = its row within the sample (first minimum in ascending order).  Rows beyond
 * count are not written by the forward and get zero gradient.
 * With a workspace (dibr_nnd_workspace_bytes, 256-byte aligned, the SAME buffer for the forward and its backward) the
 * search runs on a uniform grid over the target cloud and the backward on an inverse index -- same answers bit for
 * bit (coordinates must be finite); with workspace == NULL both are the reference's exhaustive O(n m) loops. */
typedef struct DibrNnd {
    int32_t batch, stride1, stride2, reserved;
    const int32_t *count1, *count2;          /* device [batch] or NULL */
    const float *xyz1, *xyz2;
    float *dist1, *dist2;                    /* [batch*stride_i] */
    int32_t *idx1, *idx2;
    const float *graddist1, *graddist2;      /* backward inputs */
    float *gradxyz1, *gradxyz2;              /* backward outputs [batch*stride_i, 3] */
    void *workspace;                         /* optional scratch, see above */
    size_t workspace_bytes;
} DibrNnd;
int dibr_nnd_workspace_bytes(const DibrNnd *p, size_t *bytes);
int dibr_nnd_forward(const DibrNnd *p, void *stream);
int dibr_nnd_backward(const DibrNnd *p, void *stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Depth map -> compacted point cloud: backproject_th (lib/pysixd/misc.py:350-367) followed by the per-sample boolean
 * mask `pc[pc[:, :, 2] > 0]` of core/self6dpp/losses/depth_bp_chamfer_loss.py:27-36, for the whole batch, without a
 * host sync.  points rows [0, count[b]) hold (X, Y, Z) of the pixels with depth > 0 in row-major pixel order (the
 * order boolean indexing gives); slot maps a pixel to its row (-1: dropped); chunk_count is scratch of
 * batch * ceil(H*W / 1024) ints.  The backward turns d loss / d points into d loss / d depth. */
typedef struct DibrBackproject {
    int32_t batch, height, width, num_K;     /* num_K = 1 (shared intrinsics) or batch */
    const float *depth;                      /* [batch, H, W] */
    const float *K;                          /* [num_K, 3, 3] */
    float *points;                           /* [batch, H*W, 3] */
    int32_t *count;                          /* [batch] */
    int32_t *slot;                           /* [batch, H*W] */
    int32_t *chunk_count;                    /* scratch */
    const float *grad_points;                /* backward in:  [batch, H*W, 3] */
    float *grad_depth;                       /* backward out: [batch, H, W] */
} DibrBackproject;
int dibr_backproject_compact(const DibrBackproject *p, void *stream);
int dibr_backproject_compact_backward(const DibrBackproject *p, void *stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Re-weighted BCE on probabilities: weighted_ex_loss_probs (core/self6dpp/losses/mask_losses.py:63-108), the loss the
 * rendered soft mask feeds (self_engine_utils.py:541-545).  scratch: dibr_mask_loss_scratch_floats(n) floats followed by
 * one uint32 that must be ZERO before the first call (the kernel re-arms it).  out[0] = loss, out[1] = |target > 0|,
 * out[2] = |target == 0|; the backward reads out[1..2] and grad_out[0] (d L / d loss) and writes d L / d probs. */
typedef struct DibrMaskLoss {
    int64_t n;                               /* elements of probs / target / weight */
    const float *probs, *target, *weight;    /* weight may be NULL */
    float *scratch;                          /* see above */
    float *out;                              /* [3] */
    const float *grad_out;                   /* backward in:  [1] */
    float *grad_probs;                       /* backward out: [n] */
} DibrMaskLoss;
int dibr_mask_loss_scratch_floats(int64_t n);
int dibr_mask_loss_forward(const DibrMaskLoss *p, void *stream);
int dibr_mask_loss_backward(const DibrMaskLoss *p, void *stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Chamfer distances -> the depth loss of core/self6dpp/losses/depth_bp_chamfer_loss.py:38-62: per sample the mean of the
 * distances below `threshold` (threshold <= 0: all) in both directions, samples with an empty selection skipped (the
 * reference's NaN test), the sum divided by max(#valid samples, 1).  stats: [batch, 4] scratch that the backward reads;
 * ticket: one uint32, ZERO before the first call (re-armed by the kernel); out[0] = loss, out[1] = #valid samples. */
typedef struct DibrChamferReduce {
    int32_t batch, stride1, stride2;
    float threshold;
    const int32_t *count1, *count2;          /* device [batch] or NULL (all rows) */
    const float *dist1, *dist2;              /* [batch, stride_i] */
    float *stats;                            /* [batch, 4] */
    uint32_t *ticket;
    float *out;                              /* [2] */
    const float *grad_out;                   /* backward in:  [1] */
    float *grad_dist1, *grad_dist2;          /* backward out: [batch, stride_i] */
} DibrChamferReduce;
int dibr_chamfer_reduce_forward(const DibrChamferReduce *p, void *stream);
int dibr_chamfer_reduce_backward(const DibrChamferReduce *p, void *stream);

/* ---------------------------------------------------------------------------------------------------------------
 * L1 loss in normalised CIE-Lab between the real crop and the rendered crop: core/self6dpp/engine/self_engine_utils.py:
 * 745-773 over lib/torch_utils/color/lab.py:16-82 (rgb_to_lab + normalize_lab) and color/xyz.py:28-30:
 *     loss = sum |lab(gt) * m - lab(ren) * m| / max(1, sum m)      (a and b channels only when no_l, the LAB_NO_L switch)
 * gt / ren are planar [n_img, 3, H*W]; bgr != 0 when the planes are B,G,R (the reference's [:, [2,1,0]] flip is then done
 * in the load).  mask is [n_img, H*W] or NULL (ones).  scratch: dibr_lab_loss_scratch_floats(n_img * hw) floats followed
 * by one uint32 that must be ZERO before the first call (re-armed by the kernel).  out[0] = loss, out[1] = sum |diff|,
 * out[2] = max(1, sum m).  The backward reads out[2] and grad_out[0] and writes d L / d ren (same layout as ren); like the
 * reference's autograd it yields NaN at exactly-black rendered pixels (pow(0, 1/3) backward). */
typedef struct DibrLabLoss {
    int32_t n_img, hw, bgr, no_l;
    const float *gt, *ren, *mask;
    float *scratch;
    float *out;                              /* [3] */
    const float *grad_out;                   /* backward in:  [1] */
    float *grad_ren;                         /* backward out: [n_img, 3, hw] */
} DibrLabLoss;
int dibr_lab_loss_scratch_floats(int64_t pixels);
int dibr_lab_loss_forward(const DibrLabLoss *p, void *stream);
int dibr_lab_loss_backward(const DibrLabLoss *p, void *stream);

/* ---------------------------------------------------------------------------------------------------------------
 * MS-SSIM: core/self6dpp/losses/ssim.py:58-160 (ssim + ms_ssim with use_padding=False), the MS_SSIM module the
 * self-supervised loop builds at core/self6dpp/engine/self_engine.py:352 and applies to the real and the rendered crop at
 * self_engine_utils.py:777-785.  x, y: [n_img, channels, height, width]; window: the 11 taps of ssim.py:13-30
 * (create_window(11, sigma)); weights: `levels` level weights (ssim.py:222-229); out: [n_img] MS-SSIM values (the
 * reference's product, last level's term raised to levels-1, normalisation (v+1)/2 when `normalize`).  Every level
 * must keep at least 11x11 pixels.  The forward fills `workspace` (dibr_ms_ssim_workspace_bytes) with the pyramid and,
 * when want_grad, the per-level coefficient maps; the backward must be given the same workspace untouched and writes
 * d L / d y for grad_out = d L / d out ([n_img]).  x is data on this path (no gradient). */
typedef struct DibrMsSsim {
    int32_t n_img, channels, height, width, levels, normalize, want_grad;
    int32_t use_padding;                     /* ssim.py:42-55: zero-pad the 11-tap windows (maps of the input's size) instead of valid windows */
    float data_range;
    float window[11];
    float weights[8];
    const float *x, *y;
    void *workspace;
    size_t workspace_bytes;
    float *out;                              /* [n_img] */
    const float *grad_out;                   /* backward in:  [n_img] */
    float *grad_y;                           /* backward out: [n_img, channels, height, width] */
} DibrMsSsim;
int dibr_ms_ssim_workspace_bytes(const DibrMsSsim *p, size_t *bytes);
int dibr_ms_ssim_forward(const DibrMsSsim *p, void *stream);
int dibr_ms_ssim_backward(const DibrMsSsim *p, void *stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Soft dice loss on probabilities: soft_dice_loss (core/self6dpp/losses/mask_losses.py:444-463), the other loss the
 * rendered soft mask can feed (MASK_INIT_REN_LOSS_TYPE == "dice", self_engine_utils.py:546-549).  probs / labels:
 * [num, per]; reduction 0 = mean (1 - sum(score) / num), 1 = sum, 2 = none (out has num entries).  stats: [num, 3] scratch
 * the backward reads; ticket: one uint32, ZERO before the first call (re-armed by the kernel).  The backward writes
 * d L / d probs for grad_out = d L / d out ([1], or [num] for reduction 2); labels are data (no gradient). */
typedef struct DibrDiceLoss {
    int32_t num, reduction;
    int64_t per;
    float smooth, eps;
    const float *probs, *labels;
    float *stats;                            /* [num, 3] */
    uint32_t *ticket;
    float *out;
    const float *grad_out;                   /* backward in */
    float *grad_probs;                       /* backward out: [num, per] */
} DibrDiceLoss;
int dibr_dice_loss_forward(const DibrDiceLoss *p, void *stream);
int dibr_dice_loss_backward(const DibrDiceLoss *p, void *stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Normal-map loss: NORMLoss (core/self6dpp/losses/vf_norm_loss.py:56-103), which compares the network's normals with the
 * cropped teacher render (self_engine_utils.py:667-680):  a = mask * out, b = mask * gt;
 *     loss = [with_l1] mean |a - b|  +  [with_cs] sum mask * (1 - cos(a, b)) / #(mask != 0)
 * out_norm / gt_norm: planar [n_img, 3, hw]; mask: [n_img, hw].  scratch: dibr_norm_loss_scratch_floats(n_img * hw) floats
 * followed by one uint32 that must be ZERO before the first call (re-armed by the kernel).  out[0] = loss, out[1] =
 * #(mask != 0); the backward reads out[1] and grad_out[0] and writes d L / d out_norm.  gt_norm and mask are data. */
typedef struct DibrNormLoss {
    int32_t n_img, hw, with_l1, with_cs;
    const float *out_norm, *gt_norm, *mask;
    float *scratch;
    float *out;                              /* [2] */
    const float *grad_out;                   /* backward in:  [1] */
    float *grad_out_norm;                    /* backward out: [n_img, 3, hw] */
} DibrNormLoss;
int dibr_norm_loss_scratch_floats(int64_t pixels);
int dibr_norm_loss_forward(const DibrNormLoss *p, void *stream);
int dibr_norm_loss_backward(const DibrNormLoss *p, void *stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Crop & resize of rendered images: batch_crop_resize (core/utils/zoom_utils.py:80-95) = detectron2's
 * ROIAlign(output_size, spatial_scale, sampling_ratio, aligned), the op Self6D++ applies to the rendered colour image and
 * the teacher normals (self_engine_utils.py:528-533, 662-666, 690-692).  input: [num_images, channels, height, width]
 * addressed through ELEMENT strides (any dense layout, e.g. the renderer's channels-last images, is read in place);
 * rois: [num_rois, 5] = (image index, x1, y1, x2, y2); output: [num_rois, channels, pooled_h, pooled_w] contiguous.
 * sampling_ratio <= 0: ceil(roi size / output size) samples per bin and axis (the reference's setting).  The backward
 * writes EVERY element of grad_input (same strides as input; no memset needed) by a fixed-order gather: bit-reproducible,
 * unlike the atomicAdd scatter it replaces. */
typedef struct DibrRoiAlign {
    int32_t num_rois, num_images, channels, height, width, pooled_h, pooled_w, sampling_ratio, aligned, reserved0;
    float spatial_scale, reserved1;
    int64_t stride_n, stride_c, stride_h, stride_w;
    const float *input;
    const float *rois;
    float *output;
    const float *grad_output;                /* backward in */
    float *grad_input;                       /* backward out */
} DibrRoiAlign;
int dibr_roi_align_forward(const DibrRoiAlign *p, void *stream);
int dibr_roi_align_backward(const DibrRoiAlign *p, void *stream);

/* batch_crop_resize(interpolation="nearest") (core/utils/zoom_utils.py:91-92) = torchvision.ops.RoIPool(output_size,
 * spatial_scale): maximum over quantised bins.  Same addressing as DibrRoiAlign; argmax: [num_rois, channels, pooled_h,
 * pooled_w] int32 scratch the forward fills (position h * width + w of the maximum, -1 for an empty bin) and the backward
 * reads.  The backward writes EVERY element of grad_input by a fixed-order gather (torchvision scatters with atomicAdd). */
typedef struct DibrRoiPool {
    int32_t num_rois, num_images, channels, height, width, pooled_h, pooled_w, reserved0;
    float spatial_scale, reserved1;
    int64_t stride_n, stride_c, stride_h, stride_w;
    const float *input;
    const float *rois;
    float *output;
    int32_t *argmax;
    const float *grad_output;                /* backward in */
    float *grad_input;                       /* backward out */
} DibrRoiPool;
int dibr_roi_pool_forward(const DibrRoiPool *p, void *stream);
int dibr_roi_pool_backward(const DibrRoiPool *p, void *stream);

/* how many kernels the library has launched on this thread since the last reset (bench evidence) */
long long dibr_launch_count(int reset);
/* a caller that replays a captured CUDA graph of library calls adds the kernels of the replay here (the count above only sees
 * the launches of the capture) */
void dibr_launch_count_add(long long n);

#ifdef __cplusplus
}
#endif
#endif /* DIBR_B200_H */
