// Internal launch-parameter structs shared by the kernels and the C-ABI translation unit.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "dibr_common.cuh"

#define DIBR_MAX_ATTR_INTERNAL 12

namespace dibr {

// workspace carved out of DibrPass::workspace by dibr_abi.cu
struct Workspace {
    FaceRec* recs;      // [total_faces]
    int4* fvid;         // fused mode: [total_faces] rows of the three corners in vert_attr (attr_base + vertex id); the forward and the
                        // backward gather the corner attributes through it instead of a materialised [F, 3, D] array
    uint32_t* bins;     // per image, per 16x16 tile: bitmap over the image's faces (bit = face may reach the tile); layout in bin_words()
    size_t bins_bytes;
    size_t clear_bytes;  // from order_cnt to the end of face_flags: what the set-up call clears with one memset
    int* order_cnt;     // [4 * ORDER_BUCKETS] tiles per cost bucket (bucket = ceil(listed faces / 32), capped) + the plan summary
    int* tile_count;    // [batch * tiles] faces listed in the tile's bitmap (bumped by the binning; directly after order_cnt: same memset)
    int* img_done;      // [batch] faces of the image that are binned (the CTA that completes an image plans its tiles; same memset)
    int* order_seg;     // [ORDER_BUCKETS, batch * tiles] tile ids of each bucket: the forward kernel works heaviest bucket first
    int4* order_desc;   // same shape: {tile id, first face of the image, one past its last face, offset of the tile's bitmap in bins (words)}:
                        // everything a tile CTA needs to start reading its bitmap, in one load
    float* xs;          // [width]  pixel-centre x
    float* ys;          // [height] pixel-centre y
    float* pose_part;   // [num_instances * POSE_BLOCKS * 12] partial pose-gradient sums
    float* cam_rot;     // pose mode: [num_instances, 9]
    float* cam_pos;     // pose mode: [num_instances, 3]
    float* cam_proj;    // pose mode: [num_K, 16]
    int* list_counts;   // [2] entries in color_list / soft_list (+ padding), zeroed ahead of the forward, filled by the backward
    unsigned int* pose_done;    // [num_instances + 1] vertex blocks that have delivered their pose sums (zeroed by dibr_forward with the lists);
                                // last entry: instances that are final
    unsigned int* face_flags;   // [total_faces] zeroed by dibr_forward.  bit 0: the face won a pixel, bit 1: it entered a soft product (set by the forward); bits 2, 3: listed (set by the backward)
    unsigned char* open8;       // [batch, H, ceil(W/8)] bit x%8 of byte x/8: pixel (y, x) is uncovered.  Written by the forward for every
                                // touched tile (the only ones a face's expanded pixel range can reach), read by the backward's soft part
    unsigned char* closed8;     // same layout: the uncovered pixel was closed by its K-th face (imidx < 0): the only uncovered pixels
                                // where the backward has to look at imidx
    int* color_list;    // [total_faces] global face ids that won at least one pixel (compacted from face_flags by the backward, arbitrary order)
    int* soft_list;     // [total_faces] global face ids that entered at least one soft-silhouette product
    size_t bytes;
};

#ifndef DIBR_POSE_BLOCKS
#define DIBR_POSE_BLOCKS 12
#endif
constexpr int POSE_BLOCKS = DIBR_POSE_BLOCKS;   // vertex blocks per instance in the pose-gradient reduction (12 x 32 instances of 256 threads x 80 registers = one wave)

struct SetupParams {
    int batch, height, width, multiplier;
    int total_faces, faces_per_image;
    const int32_t* face_offsets;
    // seam mode
    const float* points3d;
    const float* points2d;
    const float* normalz;
    // fused mode
    int num_instances;
    const int32_t* inst_desc;
    const float* verts;
    const int32_t* mesh_faces;
    const float* vert_attr;
    int vert_attr_dim, attr_flags, num_attr;
    int verts_stride, vert_attr_stride;      // floats per row (3 / vert_attr_dim when packed)
    const float* cam_rot;
    const float* cam_pos;
    const float* cam_proj;
    float* face_attr;
    float* face_normal;
    // pose mode
    const float* pose_R;
    const float* pose_t;
    const float* pose_K;
    int num_K;
    float q, qn;             // -(f+n)/(f-n), -2fn/(f-n)
    float expand_mul;        // expand * multiplier: the bins cover the EXPANDED bboxes
    unsigned int* out_min;   // the pass's batch-global minimum (ordered-uint encoding), reset to 'nothing seen' by the set-up kernel; or null
    Workspace ws;
};

// fused mode: the per-corner attributes of a face are [row of vert_attr | 1 (flags & 1) | -view z (flags & 2)]
// (vcrender_batch.py:84-88, renderer_dibr.py:296-301), gathered through the face's three row ids
struct VertexAttr {
    const int4* fvid;          // [total_faces] rows of the three corners, or null: seam mode, attributes come from face_attr
    const float* table;
    int dim, stride, flags, vec;
};

struct FwdParams {
    int batch, height, width, num_attr, knum, multiplier, delta;
    float expand_mul;
    int faces_per_image;
    const int32_t* face_offsets;
    int total_faces;
    const FaceRec* recs;
    const uint32_t* bins;
    const int* order_cnt;
    const int* order_seg;
    const int4* order_desc;
    const float* xs;           // [width], [height] pixel-centre tables
    const float* ys;
    const float* face_attr;
    VertexAttr va;
    int n_out;                 // channel groups (>= 1)
    int out_ch[6];
    float* out[6];             // [batch,H,W,out_ch[g]]
    float* chan_out[DIBR_MAX_ATTR_INTERNAL];    // per attribute channel d: its output tensor, pre-offset by the channel's slot in the group
    int chan_stride[DIBR_MAX_ATTR_INTERNAL];    // floats per pixel of that tensor
    int chan_off32[DIBR_MAX_ATTR_INTERNAL];     // channel d of lane l sits at chan_off32[d] + l * chan_stride[d] in a warp's 32-pixel transpose buffer
    int chan_off[DIBR_MAX_ATTR_INTERNAL];       // channel d of pixel p sits at chan_off[d] + p * chan_stride[d] in the tile's shared-memory copy
    unsigned min_mask;         // channels of the output group whose batch-global minimum is accumulated
    int vec_out;               // every output group can be written with 16 B stores (aligned base, W * channels % 4 == 0)
    float* improb;
    float* imcomp;
    int32_t* imidx;
    int* list_counts;
    unsigned int* face_flags;
    int* color_list;
    int* soft_list;
    unsigned char* open8;
    unsigned char* closed8;
    int min_group;             // output group whose batch-global minimum is accumulated, or -1
    unsigned int* out_min;     // ordered-uint encoding
};

struct BwdParams {
    int batch, height, width, num_attr, knum, multiplier, delta;
    float expand_mul;
    int total_faces, faces_per_image;
    const int32_t* face_offsets;
    const FaceRec* recs;
    const float* xs;
    const float* ys;
    const float* face_attr;
    VertexAttr va;
    const float* improb;
    const float* imcomp;
    const int32_t* imidx;
    const unsigned char* open8;
    const unsigned char* closed8;
    int* list_counts;          // [0] colour list length, [1] soft list length
    const unsigned char* face_flags;
    int* color_list;
    int* soft_list;
    const float* chan_grad[DIBR_MAX_ATTR_INTERNAL];   // per channel d: upstream gradient base (pre-offset) or null
    int chan_stride[DIBR_MAX_ATTR_INTERNAL];          // floats per pixel of the tensor that holds channel d
    int any_grad_im;
    const float* grad_improb;
    float* grad_points2d;
    float* grad_face_attr;
    int attr_compact;          // attr_flags bit 2: grad_face_attr is scratch for dibr_backward_meshes alone -- only the depth channel's
                               // corner gradients are kept, as a [total_faces, 3] array at its start (nothing when there is no depth channel)
};

struct MeshBwdParams {
    int num_instances;
    const int32_t* inst_desc;
    const float* verts;
    int verts_stride;
    const float* cam_rot;
    const float* cam_pos;
    const float* cam_proj;
    int vert_attr_dim, attr_flags, num_attr;
    const float* grad_points2d;
    const float* grad_face_attr;
    const int32_t* vert_face_ptr;
    const int32_t* vert_face_idx;
    float* grad_verts;
    float* grad_vert_attr;
    float* grad_cam_rot;
    float* grad_cam_pos;
    float* pose_part;
    unsigned int* pose_done;
    // pose mode: chain to R, t
    const float* pose_R;
    const float* pose_t;
    float* grad_pose_R;
    float* grad_pose_t;
    float* grad_pose_packed;   // optional [num_instances, 12]: dL/dR then dL/dt, the layout dibr_render_step copies back to the host
    float* host_pose_packed;   // optional device view of the caller's pinned, mapped [num_instances, 12] buffer: the finalising blocks
                               // write the rows there as well (no device-to-host copy behind the kernel)
    int pose_sum;              // the packed buffer has one more row that receives the column sums
    int attr_compact;          // grad_face_attr holds the depth channel alone, [total_faces, 3] (see BwdParams)
};

// Tile bins.  Image b owns the global 32-face words [f_lo >> 5, (f_hi - 1) >> 5]; its bitmaps start at word
// tiles * ((f_lo >> 5) + b) (images never overlap: consecutive images share at most one boundary word, and the "+ b"
// pays for it), one run of nw words per tile.  Total: tiles * (ceil(total_faces / 32) + batch + 1) words.
constexpr int ORDER_BUCKETS = 32;
// plan summary, written behind the 32 bucket counters by the set-up CTA that plans the last image:
constexpr int PLAN_TOUCHED = 64;        // order_cnt[64]: tiles with a non-empty bitmap
constexpr int PLAN_WORK_CTAS = 65;      // order_cnt[65]: CTAs of the tile-CTA forward that have work (touched tiles + ceil(untouched / 8))
constexpr int PLAN_TICKET = 66;         // order_cnt[66]: images whose tiles are planned
constexpr int PLAN_START = 96;          // order_cnt[96 + l]: first heaviest-first position of bucket 31 - l
__host__ __device__ inline size_t bin_total_words(int width, int height, int batch, int total_faces) {
    const size_t tiles = (size_t)((width + TILE - 1) / TILE) * (size_t)((height + TILE - 1) / TILE);
    return tiles * ((size_t)((total_faces + 31) / 32) + (size_t)batch + 1);
}

constexpr int INST_STRIDE = 12;
// inst_desc columns
enum { I_VERT_BASE = 0, I_NUM_VERTS, I_MESH_FACE_BASE, I_NUM_FACES, I_OUT_FACE_BASE, I_CAM, I_PROJ, I_ATTR_BASE,
       I_GVERT_BASE, I_IMAGE, I_ADJ_BASE, I_RESERVED };

// chamfer nearest-neighbour op: cloud i of sample b = rows [b*stride_i, b*stride_i + count_i[b]) (count NULL: all rows)
struct NndParams {
    int batch, stride1, stride2;
    const int* count1;
    const int* count2;
    const float* xyz1;
    const float* xyz2;
    float* dist1;
    float* dist2;
    int* idx1;
    int* idx2;
    const float* graddist1;
    const float* graddist2;
    float* gradxyz1;
    float* gradxyz2;
};

// re-weighted BCE on probabilities (dibr_maskloss.cu)
struct MaskLossParams {
    long long n;
    const float* probs;
    const float* target;
    const float* weight;       // or null
    float* partial;            // [4 * CTAs] scratch
    unsigned int* ticket;      // [1], zero between calls
    float* out;                // [3]: loss, |pos|, |neg|
    const float* grad_out;     // backward: [1]
    float* grad_probs;         // backward: [n]
};
int mask_loss_partial_floats(long long n);
int launch_mask_loss_forward(const MaskLossParams& P, cudaStream_t stream);
int launch_mask_loss_backward(const MaskLossParams& P, cudaStream_t stream);
// soft dice loss on probabilities (dibr_maskloss.cu)
struct DiceLossParams {
    int num;                   // samples
    long long per;             // elements per sample
    int reduction;             // 0 mean, 1 sum, 2 none
    float smooth, eps;
    const float* probs;        // [num, per]
    const float* labels;       // [num, per]
    float* stats;              // [num, 3]: sum p l, sum p, sum l (kept for the backward)
    unsigned int* ticket;      // [1], zero between calls
    float* out;                // [1] (mean, sum) or [num] (none)
    const float* grad_out;     // backward: [1] or [num]
    float* grad_probs;         // backward: [num, per]
};
int launch_dice_loss_forward(const DiceLossParams& P, cudaStream_t stream);
int launch_dice_loss_backward(const DiceLossParams& P, cudaStream_t stream);
// L1 + cosine loss between predicted and rendered normals (dibr_maskloss.cu)
struct NormLossParams {
    int n_img;                 // images
    int hw;                    // pixels per plane
    int with_l1, with_cs;
    const float* out_norm;     // [n_img, 3, hw]
    const float* gt_norm;      // [n_img, 3, hw]
    const float* mask;         // [n_img, hw]
    float* partial;            // [3 * CTAs] scratch
    unsigned int* ticket;      // [1], zero between calls
    float* out;                // [2]: loss, #(mask != 0)
    const float* grad_out;     // backward: [1]
    float* grad_out_norm;      // backward: [n_img, 3, hw]
};
int norm_loss_partial_floats(long long pixels);
int launch_norm_loss_forward(const NormLossParams& P, cudaStream_t stream);
int launch_norm_loss_backward(const NormLossParams& P, cudaStream_t stream);
// ROIAlign crop & resize of rendered images (dibr_roialign.cu)
struct RoiAlignParams {
    const float* input;        // [num_images, channels, height, width] through element strides
    const float* rois;         // [num_rois, 5]: image index, x1, y1, x2, y2
    float* output;             // [num_rois, channels, pooled_h, pooled_w] contiguous
    const float* grad_output;  // backward in, same layout as output
    float* grad_input;         // backward out, same strides as input; every element is written
    int num_rois, num_images, channels, height, width, pooled_h, pooled_w, sampling_ratio, aligned;
    float spatial_scale;
    long long stride_n, stride_c, stride_h, stride_w;
};
int launch_roi_align_forward(const RoiAlignParams& P, cudaStream_t stream);
int launch_roi_align_backward(const RoiAlignParams& P, cudaStream_t stream);
// RoIPool (max over quantised bins), the 'nearest' mode of batch_crop_resize
struct RoiPoolParams {
    const float* input;        // [num_images, channels, height, width] through element strides
    const float* rois;         // [num_rois, 5]
    float* output;             // [num_rois, channels, pooled_h, pooled_w]
    int* argmax;               // same shape: h * width + w of the maximum, -1 for an empty bin
    const float* grad_output;
    float* grad_input;         // same strides as input; every element is written
    int num_rois, num_images, channels, height, width, pooled_h, pooled_w;
    float spatial_scale;
    long long stride_n, stride_c, stride_h, stride_w;
};
int launch_roi_pool_forward(const RoiPoolParams& P, cudaStream_t stream);
int launch_roi_pool_backward(const RoiPoolParams& P, cudaStream_t stream);
// L1 in normalised CIE-Lab between the real and the rendered crop (dibr_photometric.cu)
struct LabLossParams {
    int n_img;                 // images
    int hw;                    // pixels per plane
    int bgr;                   // planes are B,G,R (the reference flips with [:, [2,1,0]]) instead of R,G,B
    int no_l;                  // LAB_NO_L: only the a and b channels count
    const float* gt;           // [n_img, 3, hw]
    const float* ren;          // [n_img, 3, hw]
    const float* mask;         // [n_img, hw] or null (all ones)
    float* partial;            // [2 * CTAs] scratch
    unsigned int* ticket;      // [1], zero between calls
    float* out;                // [3]: loss, sum |diff|, max(1, sum mask)
    const float* grad_out;     // backward: [1]
    float* grad_ren;           // backward: [n_img, 3, hw]
};
int lab_loss_partial_floats(long long pixels);
int launch_lab_loss_forward(const LabLossParams& P, cudaStream_t stream);
int launch_lab_loss_backward(const LabLossParams& P, cudaStream_t stream);
// MS-SSIM (dibr_photometric.cu): one level of the pyramid
constexpr int SSIM_MAX_LEVELS = 8;
struct SsimLevelParams {
    int H, W;                  // this level's image size
    int channels;
    int use_ssim;              // last level: the maps are d ssim, otherwise d cs
    int pad;                   // 0: valid windows; 5: zero padding (ssim.py use_padding)
    float C1, C2;
    float win[11];
    const float* x;            // [planes, H, W]
    const float* y;
    float* maps;               // [planes, 3, H-10, W-10] or null (forward without gradient)
    float* partial;            // forward: [planes, tiles, 2]
    const float* scale;        // backward: [n_img] of this level
    const float* grad_out;     // backward: [n_img]
    const float* grad_coarse;  // backward: [planes, Hc, Wc] gradient of the next (pooled) level, or null
    int Hc, Wc;
    float* grad_y;             // backward: [planes, H, W]
};
struct SsimCombineParams {
    int n_img, channels, levels, normalize;
    float weights[SSIM_MAX_LEVELS];
    int tiles[SSIM_MAX_LEVELS];
    int map_pixels[SSIM_MAX_LEVELS];
    long long partial_off[SSIM_MAX_LEVELS];   // in floats from `partial`
    const float* partial;
    float* out;                // [n_img]
    float* scale;              // [levels, n_img] or null
};
int launch_ssim_pool(const float* x, const float* y, float* px, float* py, int planes, int H, int W, int Ho, int Wo, cudaStream_t stream);
int ssim_forward_tiles(int H, int W, int pad);
int launch_ssim_level_forward(const SsimLevelParams& P, int planes, cudaStream_t stream);
int launch_ssim_combine(const SsimCombineParams& P, cudaStream_t stream);
int launch_ssim_level_backward(const SsimLevelParams& P, int planes, cudaStream_t stream);
// chamfer distances -> depth loss (dibr_maskloss.cu)
struct ChamferReduceParams {
    int batch, stride1, stride2;
    float threshold;
    const int* count1;
    const int* count2;
    const float* dist1;
    const float* dist2;
    float* stats;              // [batch, 4]: sum1, n1, sum2, n2 of the selected distances
    unsigned int* ticket;      // [1], zero between calls
    float* out;                // [2]: loss, number of valid samples
    const float* grad_out;     // backward: [1]
    float* grad_dist1;
    float* grad_dist2;
};
int launch_chamfer_reduce_forward(const ChamferReduceParams& P, cudaStream_t stream);
int launch_chamfer_reduce_backward(const ChamferReduceParams& P, cudaStream_t stream);
// depth map -> compacted cloud (dibr_backproject.cu)
struct BackprojectParams {
    int batch, height, width, num_K;
    const float* depth;        // [batch, H, W]
    const float* K;            // [num_K, 9], num_K = 1 or batch
    float* points;             // [batch, H*W, 3] rows [0, count[b]) valid, row-major pixel order
    int* count;                // [batch]
    int* slot;                 // [batch, H*W] pixel -> row, -1 where depth <= 0
    int* chunk_count;          // [batch, ceil(H*W / 1024)] scratch
    const float* grad_points;  // backward
    float* grad_depth;
};
int launch_backproject(const BackprojectParams& P, cudaStream_t stream);
int launch_backproject_backward(const BackprojectParams& P, cudaStream_t stream);
int launch_nnd_forward(const NndParams& P, cudaStream_t stream);
size_t nnd_grid_workspace_bytes(int batch, int stride1, int stride2);
int launch_nnd_forward_grid(const NndParams& P, void* workspace, cudaStream_t stream);
int launch_nnd_backward_grid(const NndParams& P, void* workspace, cudaStream_t stream);
int launch_nnd_backward(const NndParams& P, cudaStream_t stream);
int launch_setup_faces(const SetupParams& P, cudaStream_t stream);
int launch_setup_meshes(const SetupParams& P, cudaStream_t stream);
int launch_forward(const FwdParams& P, cudaStream_t stream);
// parts: bit 0 = prepare (zero the gradients, build the work lists from the forward's flags), bit 1 = the face kernel
int launch_backward_faces(const BwdParams& P, cudaStream_t stream, int parts = 3);
int launch_backward_meshes(const MeshBwdParams& P, cudaStream_t stream);
int launch_normal_map_tiles(int batch, int height, int width, const int* order_cnt, const int* order_seg, const float* n, const float* mask,
                            const unsigned int* min_ordered, float* out, cudaStream_t stream);
int launch_normal_map(const float* n, const float* mask, const unsigned int* min_ordered, float* out, long long npix, cudaStream_t stream);

}  // namespace dibr
