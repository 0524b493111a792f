// extern "C" boundary of libdibr_b200.so (see include/dibr_b200.h): argument validation, workspace
// carving and launch of the kernels.  No torch types, no allocation, no device synchronisation.
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "../../include/dibr_b200.h"
#include "dibr_internal.h"

namespace {

thread_local char g_err[512] = "";
thread_local long long g_launches = 0;

int fail(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return 1;
}

int cuda_fail(const char* what, int code) {
    if (code == 0) return 0;
    return fail("%s: CUDA error %d (%s)", what, code, cudaGetErrorString((cudaError_t)code));
}

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

dibr::Workspace carve(const DibrPass* p, void* base) {
    dibr::Workspace w;
    size_t off = 0;
    char* b = (char*)base;
    auto take = [&](size_t bytes) { void* r = b ? (void*)(b + off) : nullptr; off += align_up(bytes, 256); return r; };
    w.recs = (dibr::FaceRec*)take(sizeof(dibr::FaceRec) * (size_t)p->total_faces);
    w.fvid = (int4*)take(sizeof(int4) * (size_t)p->total_faces);
    {
        const size_t ntiles = (size_t)p->batch * ((p->height + dibr::TILE - 1) / dibr::TILE) * ((p->width + dibr::TILE - 1) / dibr::TILE);
        w.order_seg = (int*)take(sizeof(int) * dibr::ORDER_BUCKETS * ntiles);
        w.order_desc = (int4*)take(sizeof(int4) * dibr::ORDER_BUCKETS * ntiles);
        // counters (+ the plan summary), per-tile face counters, per-image progress, then the bins: one memset clears all
        w.order_cnt = (int*)take(sizeof(int) * 4 * dibr::ORDER_BUCKETS);
        w.tile_count = (int*)take(sizeof(int) * ntiles);
        w.img_done = (int*)take(sizeof(int) * (size_t)p->batch);
    }
    w.bins_bytes = sizeof(uint32_t) * dibr::bin_total_words(p->width, p->height, p->batch, p->total_faces);
    w.bins = (uint32_t*)take(w.bins_bytes);
    const size_t ni = (size_t)(p->num_instances > 0 ? p->num_instances : 0);
    // the backward's list counters, ticket counters and face flags sit directly behind the bitmaps: the set-up call clears
    // plan + bitmaps + lists with ONE memset (dibr_forward clears the lists again when it is called on its own)
    w.list_counts = (int*)take(sizeof(int) * 64);
    w.pose_done = (unsigned int*)take(sizeof(unsigned int) * (ni + 1));
    w.face_flags = (unsigned int*)take(sizeof(unsigned int) * (size_t)p->total_faces);
    w.clear_bytes = w.order_cnt ? (size_t)((char*)w.face_flags - (char*)w.order_cnt) + sizeof(unsigned int) * (size_t)p->total_faces : 0;
    w.xs = (float*)take(sizeof(float) * (size_t)p->width);
    w.ys = (float*)take(sizeof(float) * (size_t)p->height);
    w.pose_part = (float*)take(sizeof(float) * 12 * dibr::POSE_BLOCKS * (size_t)(p->num_instances > 0 ? p->num_instances : 0));
    w.cam_rot = (float*)take(sizeof(float) * 9 * ni);
    w.cam_pos = (float*)take(sizeof(float) * 3 * ni);
    w.cam_proj = (float*)take(sizeof(float) * 16 * (size_t)(p->num_K > 0 ? p->num_K : 0));
    w.color_list = (int*)take(sizeof(int) * (size_t)p->total_faces);
    w.soft_list = (int*)take(sizeof(int) * (size_t)p->total_faces);
    w.open8 = (unsigned char*)take((size_t)p->batch * p->height * ((p->width + 7) / 8));
    w.closed8 = (unsigned char*)take((size_t)p->batch * p->height * ((p->width + 7) / 8));
    w.bytes = off;
    return w;
}

int check_common(const DibrPass* p, bool need_ws) {
    if (!p) return fail("null DibrPass");
    if (p->batch <= 0 || p->height <= 0 || p->width <= 0) return fail("bad image size b=%d h=%d w=%d", p->batch, p->height, p->width);
    if (p->height > dibr::MAX_IMAGE_SIDE || p->width > dibr::MAX_IMAGE_SIDE) return fail("image side above %d", dibr::MAX_IMAGE_SIDE);
    if (p->batch > 4096) return fail("batch above 4096");
    if (p->num_attr <= 0 || p->num_attr > DIBR_MAX_ATTR) return fail("num_attr=%d outside [1,%d]", p->num_attr, DIBR_MAX_ATTR);
    if (p->knum < 0 || p->knum > 250) return fail("knum=%d outside [0,250]", p->knum);
    if (p->multiplier <= 0 || p->delta < 0) return fail("bad multiplier/delta");
    if (p->total_faces < 0) return fail("total_faces < 0");
    if (!p->face_offsets && (long long)p->faces_per_image * p->batch != (long long)p->total_faces)
        return fail("faces_per_image*batch != total_faces and no face_offsets given");
    if ((long long)p->batch * p->height * p->width >= (1ll << 31)) return fail("image batch too large for 32-bit pixel ids");
    if (dibr::bin_total_words(p->width, p->height, p->batch, p->total_faces) >= (1ull << 31)) return fail("tiles x faces too large: the tile bitmaps exceed 2^31 words");
    if (need_ws) {
        if (!p->workspace) return fail("workspace is null");
        if (((uintptr_t)p->workspace & 255u) != 0) return fail("workspace must be 256-byte aligned");
        const dibr::Workspace w = carve(p, nullptr);
        if (p->workspace_bytes < w.bytes) return fail("workspace too small: %zu < %zu", p->workspace_bytes, w.bytes);
    }
    if (dibr_device_count() <= 0) return fail("no CUDA device: libdibr_b200 has no CPU fallback");
    return 0;
}

int check_outputs(const DibrPass* p) {
    if (p->num_outputs < 0 || p->num_outputs > DIBR_MAX_OUTPUTS) return fail("num_outputs=%d outside [0,%d]", p->num_outputs, DIBR_MAX_OUTPUTS);
    if (p->num_outputs == 0) return p->im ? 0 : fail("forward: im is null");
    int d = 0;
    for (int g = 0; g < p->num_outputs; g++) {
        if (p->out_channels[g] <= 0 || !p->out[g]) return fail("forward: output group %d is empty or null", g);
        d += p->out_channels[g];
    }
    return d == p->num_attr ? 0 : fail("out_channels sum %d != num_attr %d", d, p->num_attr);
}

// a pass set up by dibr_setup_meshes: the corner attributes are gathered from the vertex table, face_attr is not used
bool is_fused(const DibrPass* p) { return p->num_instances > 0 && p->inst_desc && p->verts && p->mesh_faces; }
// attr_flags bit 2 (fused passes): the caller never reads grad_face_attr -- it is scratch between dibr_backward_faces and
// dibr_backward_meshes, which needs the depth channel alone.  Ignored when the per-vertex attribute gradient is asked for.
bool attr_compact(const DibrPass* p) { return is_fused(p) && (p->attr_flags & 4) != 0 && !p->grad_vert_attr; }

void set_vertex_attr(dibr::VertexAttr& va, const DibrPass* p, const dibr::Workspace& w) {
    va.fvid = w.fvid;
    va.table = p->vert_attr;
    va.dim = p->vert_attr_dim;
    va.stride = p->vert_attr_stride > 0 ? p->vert_attr_stride : p->vert_attr_dim;
    va.flags = p->attr_flags;
    va.vec = va.dim > 0 && (va.stride & 3) == 0 && (((uintptr_t)va.table & 15u) == 0);
}

dibr::SetupParams setup_params(const DibrPass* p) {
    dibr::SetupParams s;
    memset(&s, 0, sizeof(s));
    s.batch = p->batch; s.height = p->height; s.width = p->width; s.multiplier = p->multiplier;
    s.total_faces = p->total_faces; s.faces_per_image = p->faces_per_image; s.face_offsets = p->face_offsets;
    s.points3d = p->points3d; s.points2d = p->points2d; s.normalz = p->normalz;
    s.num_instances = p->num_instances; s.inst_desc = p->inst_desc; s.verts = p->verts; s.mesh_faces = p->mesh_faces;
    s.vert_attr = p->vert_attr; s.vert_attr_dim = p->vert_attr_dim; s.attr_flags = p->attr_flags; s.num_attr = p->num_attr;
    s.verts_stride = p->verts_stride > 0 ? p->verts_stride : 3;
    s.vert_attr_stride = p->vert_attr_stride > 0 ? p->vert_attr_stride : p->vert_attr_dim;
    s.cam_rot = p->cam_rot; s.cam_pos = p->cam_pos; s.cam_proj = p->cam_proj;
    s.face_attr = p->face_attr; s.face_normal = p->face_normal;
    s.ws = carve(p, p->workspace);
    s.pose_R = p->pose_R; s.pose_t = p->pose_t; s.pose_K = p->pose_K; s.num_K = p->num_K;
    s.out_min = (p->min_output >= 0) ? p->out_min_ordered : nullptr;
    const double nc = p->znear, fc = p->zfar;
    s.expand_mul = (float)((double)p->expand * (double)p->multiplier);
    s.q = (float)(-(fc + nc) / (fc - nc));
    s.qn = (float)(-2.0 * (fc * nc) / (fc - nc));
    return s;
}

}  // namespace

extern "C" {

int dibr_abi_version(void) { return DIBR_ABI_VERSION; }

int dibr_sizeof_pass(void) { return (int)sizeof(DibrPass); }

int dibr_sizeof_step(void) { return (int)sizeof(DibrStep); }

const char* dibr_last_error(void) { return g_err; }

int dibr_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

static int dl_params(const DibrDiceLoss* p, dibr::DiceLossParams& q, bool backward) {
    if (!p) return fail("null DibrDiceLoss");
    if (p->num < 0 || p->per < 0) return fail("dice_loss: negative size");
    if (p->num > 65535) return fail("dice_loss: more than 65535 samples");
    if (p->reduction < 0 || p->reduction > 2) return fail("dice_loss: reduction must be 0 (mean), 1 (sum) or 2 (none)");
    if (!p->out || !p->stats) return fail("dice_loss: out / stats required");
    if (p->num > 0 && p->per > 0 && (!p->probs || !p->labels)) return fail("dice_loss: probs / labels required");
    if (!backward && !p->ticket) return fail("dice_loss: ticket required");
    if (backward && (!p->grad_out || (p->num > 0 && p->per > 0 && !p->grad_probs))) return fail("dice_loss backward: grad_out / grad_probs required");
    if (dibr_device_count() <= 0) return fail("no CUDA device: libdibr_b200 has no CPU fallback");
    q.num = p->num; q.per = p->per; q.reduction = p->reduction; q.smooth = p->smooth; q.eps = p->eps;
    q.probs = p->probs; q.labels = p->labels; q.stats = p->stats; q.ticket = p->ticket; q.out = p->out;
    q.grad_out = p->grad_out; q.grad_probs = p->grad_probs;
    return 0;
}
int dibr_dice_loss_forward(const DibrDiceLoss* p, void* stream) {
    dibr::DiceLossParams q;
    if (int e = dl_params(p, q, false)) return e;
    g_launches += 1;
    return cuda_fail("dibr_dice_loss_forward", dibr::launch_dice_loss_forward(q, (cudaStream_t)stream));
}
int dibr_dice_loss_backward(const DibrDiceLoss* p, void* stream) {
    dibr::DiceLossParams q;
    if (int e = dl_params(p, q, true)) return e;
    g_launches += 1;
    return cuda_fail("dibr_dice_loss_backward", dibr::launch_dice_loss_backward(q, (cudaStream_t)stream));
}

int dibr_norm_loss_scratch_floats(int64_t pixels) { return dibr::norm_loss_partial_floats(pixels) + 1; }

static int nl_params(const DibrNormLoss* p, dibr::NormLossParams& q, bool backward) {
    if (!p) return fail("null DibrNormLoss");
    if (p->n_img < 0 || p->hw < 0) return fail("norm_loss: negative size");
    if (!p->with_l1 && !p->with_cs) return fail("norm_loss: with_l1 or with_cs (vf_norm_loss.py:59)");
    if (!p->out) return fail("norm_loss: out required");
    const long long pixels = (long long)p->n_img * p->hw;
    if (pixels > 0 && (!p->out_norm || !p->gt_norm || !p->mask)) return fail("norm_loss: out_norm / gt_norm / mask required");
    if (!backward && !p->scratch) return fail("norm_loss: scratch required");
    if (backward && (!p->grad_out || (pixels > 0 && !p->grad_out_norm))) return fail("norm_loss backward: grad_out / grad_out_norm required");
    if (dibr_device_count() <= 0) return fail("no CUDA device: libdibr_b200 has no CPU fallback");
    q.n_img = p->n_img; q.hw = p->hw; q.with_l1 = p->with_l1; q.with_cs = p->with_cs;
    q.out_norm = p->out_norm; q.gt_norm = p->gt_norm; q.mask = p->mask;
    q.partial = p->scratch; q.ticket = p->scratch ? (unsigned int*)(p->scratch + dibr::norm_loss_partial_floats(pixels)) : nullptr;
    q.out = p->out; q.grad_out = p->grad_out; q.grad_out_norm = p->grad_out_norm;
    return 0;
}
int dibr_norm_loss_forward(const DibrNormLoss* p, void* stream) {
    dibr::NormLossParams q;
    if (int e = nl_params(p, q, false)) return e;
    g_launches += 1;
    return cuda_fail("dibr_norm_loss_forward", dibr::launch_norm_loss_forward(q, (cudaStream_t)stream));
}
int dibr_norm_loss_backward(const DibrNormLoss* p, void* stream) {
    dibr::NormLossParams q;
    if (int e = nl_params(p, q, true)) return e;
    g_launches += 1;
    return cuda_fail("dibr_norm_loss_backward", dibr::launch_norm_loss_backward(q, (cudaStream_t)stream));
}

static int ra_params(const DibrRoiAlign* p, dibr::RoiAlignParams& q, bool backward) {
    if (!p) return fail("null DibrRoiAlign");
    if (p->num_rois < 0 || p->num_images < 0 || p->channels < 0) return fail("roi_align: negative size");
    if (p->height <= 0 || p->width <= 0 || p->pooled_h <= 0 || p->pooled_w <= 0) return fail("roi_align: bad image / output size");
    if (p->num_rois > 0 && !p->rois) return fail("roi_align: rois required");
    if (!backward && p->num_rois > 0 && p->channels > 0 && (!p->input || !p->output)) return fail("roi_align: input / output required");
    if (backward && p->num_images > 0 && p->channels > 0 && (!p->grad_input || (p->num_rois > 0 && !p->grad_output)))
        return fail("roi_align backward: grad_output / grad_input required");
    if ((long long)p->num_images * ((p->width + 63) / 64) * ((p->height + 3) / 4) >= (1ll << 31)) return fail("roi_align: input too large");
    if (dibr_device_count() <= 0) return fail("no CUDA device: libdibr_b200 has no CPU fallback");
    q.input = p->input; q.rois = p->rois; q.output = p->output; q.grad_output = p->grad_output; q.grad_input = p->grad_input;
    q.num_rois = p->num_rois; q.num_images = p->num_images; q.channels = p->channels; q.height = p->height; q.width = p->width;
    q.pooled_h = p->pooled_h; q.pooled_w = p->pooled_w; q.sampling_ratio = p->sampling_ratio; q.aligned = p->aligned;
    q.spatial_scale = p->spatial_scale;
    q.stride_n = p->stride_n; q.stride_c = p->stride_c; q.stride_h = p->stride_h; q.stride_w = p->stride_w;
    return 0;
}
int dibr_roi_align_forward(const DibrRoiAlign* p, void* stream) {
    dibr::RoiAlignParams q;
    if (int e = ra_params(p, q, false)) return e;
    g_launches += 1;
    return cuda_fail("dibr_roi_align_forward", dibr::launch_roi_align_forward(q, (cudaStream_t)stream));
}
int dibr_roi_align_backward(const DibrRoiAlign* p, void* stream) {
    dibr::RoiAlignParams q;
    if (int e = ra_params(p, q, true)) return e;
    g_launches += 1;
    return cuda_fail("dibr_roi_align_backward", dibr::launch_roi_align_backward(q, (cudaStream_t)stream));
}

static int rp_params(const DibrRoiPool* p, dibr::RoiPoolParams& q, bool backward) {
    if (!p) return fail("null DibrRoiPool");
    if (p->num_rois < 0 || p->num_images < 0 || p->channels < 0) return fail("roi_pool: negative size");
    if (p->height <= 0 || p->width <= 0 || p->pooled_h <= 0 || p->pooled_w <= 0) return fail("roi_pool: bad image / output size");
    if ((long long)p->height * p->width >= (1ll << 31)) return fail("roi_pool: image too large for 32-bit positions");
    if (p->num_rois > 0 && (!p->rois || !p->argmax)) return fail("roi_pool: rois / argmax required");
    if (!backward && p->num_rois > 0 && p->channels > 0 && (!p->input || !p->output)) return fail("roi_pool: input / output required");
    if (backward && p->num_images > 0 && p->channels > 0 && (!p->grad_input || (p->num_rois > 0 && !p->grad_output)))
        return fail("roi_pool backward: grad_output / grad_input required");
    if (dibr_device_count() <= 0) return fail("no CUDA device: libdibr_b200 has no CPU fallback");
    q.input = p->input; q.rois = p->rois; q.output = p->output; q.argmax = p->argmax; q.grad_output = p->grad_output; q.grad_input = p->grad_input;
    q.num_rois = p->num_rois; q.num_images = p->num_images; q.channels = p->channels; q.height = p->height; q.width = p->width;
    q.pooled_h = p->pooled_h; q.pooled_w = p->pooled_w; q.spatial_scale = p->spatial_scale;
    q.stride_n = p->stride_n; q.stride_c = p->stride_c; q.stride_h = p->stride_h; q.stride_w = p->stride_w;
    return 0;
}
int dibr_roi_pool_forward(const DibrRoiPool* p, void* stream) {
    dibr::RoiPoolParams q;
    if (int e = rp_params(p, q, false)) return e;
    g_launches += 1;
    return cuda_fail("dibr_roi_pool_forward", dibr::launch_roi_pool_forward(q, (cudaStream_t)stream));
}
int dibr_roi_pool_backward(const DibrRoiPool* p, void* stream) {
    dibr::RoiPoolParams q;
    if (int e = rp_params(p, q, true)) return e;
    g_launches += 1;
    return cuda_fail("dibr_roi_pool_backward", dibr::launch_roi_pool_backward(q, (cudaStream_t)stream));
}

void dibr_launch_count_add(long long n) { g_launches += n; }

long long dibr_launch_count(int reset) {
    const long long v = g_launches;
    if (reset) g_launches = 0;
    return v;
}

int dibr_workspace_bytes(const DibrPass* p, size_t* bytes) {
    if (!p || !bytes) return fail("null argument");
    if (p->batch <= 0 || p->height <= 0 || p->width <= 0 || p->total_faces < 0) return fail("bad sizes");
    *bytes = carve(p, nullptr).bytes;
    return 0;
}

int dibr_setup_faces(const DibrPass* p, void* stream) {
    if (int e = check_common(p, true)) return e;
    if (p->total_faces > 0 && (!p->points3d || !p->points2d || !p->normalz)) return fail("setup_faces: points3d/points2d/normalz required");
    const dibr::SetupParams s = setup_params(p);
    g_launches += 1;
    return cuda_fail("dibr_setup_faces", dibr::launch_setup_faces(s, (cudaStream_t)stream));
}

int dibr_setup_meshes(const DibrPass* p, void* stream) {
    if (int e = check_common(p, true)) return e;
    if (p->num_instances <= 0 || !p->inst_desc || !p->verts || !p->mesh_faces) return fail("setup_meshes: instances, verts, mesh_faces required");
    if (p->pose_R) {
        if (!p->pose_t || !p->pose_K || p->num_K <= 0 || !(p->zfar > p->znear)) return fail("setup_meshes: pose mode needs pose_t, pose_K, num_K > 0 and zfar > znear");
    } else if (!p->cam_rot || !p->cam_pos || !p->cam_proj) return fail("setup_meshes: cameras required");
    const int d = p->vert_attr_dim + ((p->attr_flags & 1) ? 1 : 0) + ((p->attr_flags & 2) ? 1 : 0);
    if (d != p->num_attr) return fail("setup_meshes: vert_attr_dim + flags = %d but num_attr = %d", d, p->num_attr);
    if (p->vert_attr_dim > 0 && !p->vert_attr) return fail("setup_meshes: vert_attr is null");
    if (p->verts_stride != 0 && p->verts_stride < 3) return fail("setup_meshes: verts_stride %d < 3", p->verts_stride);
    if (p->vert_attr_stride != 0 && p->vert_attr_stride < p->vert_attr_dim) return fail("setup_meshes: vert_attr_stride %d < vert_attr_dim %d", p->vert_attr_stride, p->vert_attr_dim);
    const dibr::SetupParams s = setup_params(p);
    g_launches += 1;
    return cuda_fail("dibr_setup_meshes", dibr::launch_setup_meshes(s, (cudaStream_t)stream));
}

// lists_clean: the set-up call of this pass has just run on the same stream (it cleared the work lists and reset the
// minimum): the resets below are skipped (dibr_render_forward; two graph nodes less per pass)
static int forward_impl(const DibrPass* p, void* stream, bool lists_clean) {
    if (int e = check_common(p, true)) return e;
    const bool fused = is_fused(p);
    if ((p->total_faces > 0 && !fused && !p->face_attr) || !p->improb || !p->imidx || !p->imcomp) return fail("forward: face_attr/improb/imidx/imcomp required");
    if (fused && p->vert_attr_dim > 0 && !p->vert_attr) return fail("forward: vert_attr is null");
    if (int e = check_outputs(p)) return e;
    const dibr::Workspace w = carve(p, p->workspace);
    dibr::FwdParams f;
    memset(&f, 0, sizeof(f));
    f.batch = p->batch; f.height = p->height; f.width = p->width; f.num_attr = p->num_attr; f.knum = p->knum;
    f.multiplier = p->multiplier; f.delta = p->delta;
    f.expand_mul = (float)((double)p->expand * (double)p->multiplier);
    f.faces_per_image = p->faces_per_image; f.face_offsets = p->face_offsets;
    f.total_faces = p->total_faces;
    f.recs = w.recs; f.bins = w.bins; f.order_cnt = w.order_cnt; f.order_seg = w.order_seg; f.order_desc = w.order_desc; f.xs = w.xs; f.ys = w.ys; f.face_attr = p->face_attr;
    if (fused) set_vertex_attr(f.va, p, w);
    if (p->num_outputs == 0) { f.n_out = 1; f.out_ch[0] = p->num_attr; f.out[0] = p->im; }
    else { f.n_out = p->num_outputs; for (int g = 0; g < f.n_out; g++) { f.out_ch[g] = p->out_channels[g]; f.out[g] = p->out[g]; } }
    {
        int d = 0, plane = 0;
        for (int g = 0; g < f.n_out; g++) {
            for (int c = 0; c < f.out_ch[g] && d < DIBR_MAX_ATTR_INTERNAL; c++, d++) {
                f.chan_out[d] = f.out[g] + c; f.chan_stride[d] = f.out_ch[g]; f.chan_off[d] = plane + c;
                f.chan_off32[d] = plane / (dibr::TILE * dibr::TILE) * 32 + c;
            }
            plane += f.out_ch[g] * dibr::TILE * dibr::TILE;
        }
        f.vec_out = 1;
        for (int g = 0; g < f.n_out; g++)
            if (((uintptr_t)f.out[g] & 15u) != 0 || (((long long)p->width * f.out_ch[g]) & 3) != 0) f.vec_out = 0;
    }
    f.improb = p->improb; f.imcomp = p->imcomp; f.imidx = p->imidx;
    f.list_counts = w.list_counts; f.face_flags = w.face_flags; f.color_list = w.color_list; f.soft_list = w.soft_list; f.open8 = w.open8; f.closed8 = w.closed8;
    if (!lists_clean) {
        const size_t nbytes = (size_t)((char*)w.face_flags - (char*)w.list_counts) + sizeof(unsigned int) * (size_t)p->total_faces;
        cudaError_t e = cudaMemsetAsync(w.list_counts, 0, nbytes, (cudaStream_t)stream);
        if (e != cudaSuccess) return cuda_fail("dibr_forward (reset face lists)", (int)e);
    }
    f.min_group = -1; f.out_min = nullptr;
    if (p->min_output >= 0 && p->out_min_ordered) {
        if (p->min_output >= f.n_out) return fail("forward: min_output=%d but only %d output groups", p->min_output, f.n_out);
        f.min_group = p->min_output; f.out_min = p->out_min_ordered;
        for (int g = 0, d = 0; g < f.n_out; d += f.out_ch[g], g++)
            if (g == f.min_group) f.min_mask = ((1u << f.out_ch[g]) - 1u) << d;
        if (!lists_clean) {
            cudaError_t e = cudaMemsetAsync(p->out_min_ordered, 0xff, sizeof(uint32_t), (cudaStream_t)stream);
            if (e != cudaSuccess) return cuda_fail("dibr_forward (reset min)", (int)e);
        }
    }
    g_launches += 1;
    return cuda_fail("dibr_forward", dibr::launch_forward(f, (cudaStream_t)stream));
}

int dibr_forward(const DibrPass* p, void* stream) { return forward_impl(p, stream, false); }

// parts: bit 0 = the prepare kernel (zero dL/dpoints2d, work lists from the forward's flags), bit 1 = the face kernel.
// dibr_render_forward runs the first part behind the student rasterisation when DibrStep.run_backward has bit 1 set, the
// backward entry points then run the second part alone.
static int backward_faces_impl(const DibrPass* p, void* stream, int parts) {
    if (int e = check_common(p, true)) return e;
    if (p->total_faces == 0) return 0;                      // nothing to differentiate
    const bool fused = is_fused(p);
    if ((!fused && !p->face_attr) || !p->improb || !p->imidx || !p->imcomp) return fail("backward_faces: saved forward buffers required");
    if (fused && p->vert_attr_dim > 0 && !p->vert_attr) return fail("backward_faces: vert_attr is null");
    if (!p->grad_points2d || !p->grad_face_attr) return fail("backward_faces: grad outputs required");
    const dibr::Workspace w = carve(p, p->workspace);
    dibr::BwdParams b;
    memset(&b, 0, sizeof(b));
    b.batch = p->batch; b.height = p->height; b.width = p->width; b.num_attr = p->num_attr; b.knum = p->knum;
    b.multiplier = p->multiplier; b.delta = p->delta;
    b.expand_mul = (float)((double)p->expand * (double)p->multiplier);
    b.total_faces = p->total_faces; b.faces_per_image = p->faces_per_image; b.face_offsets = p->face_offsets;
    b.recs = w.recs; b.xs = w.xs; b.ys = w.ys; b.face_attr = p->face_attr;
    if (fused) set_vertex_attr(b.va, p, w);
    b.improb = p->improb; b.imcomp = p->imcomp; b.imidx = p->imidx;
    b.list_counts = w.list_counts; b.face_flags = (const unsigned char*)w.face_flags; b.color_list = w.color_list; b.soft_list = w.soft_list; b.open8 = w.open8; b.closed8 = w.closed8;
    if (p->num_outputs < 0 || p->num_outputs > DIBR_MAX_OUTPUTS) return fail("num_outputs=%d outside [0,%d]", p->num_outputs, DIBR_MAX_OUTPUTS);
    b.any_grad_im = 0;
    if (p->num_outputs == 0) {
        for (int d = 0; d < p->num_attr; d++) { b.chan_grad[d] = p->grad_im ? p->grad_im + d : nullptr; b.chan_stride[d] = p->num_attr; }
        b.any_grad_im = p->grad_im != nullptr;
    } else {
        int d = 0;
        for (int g = 0; g < p->num_outputs; g++) {
            for (int c = 0; c < p->out_channels[g]; c++, d++) {
                if (d >= p->num_attr) return fail("out_channels sum exceeds num_attr");
                b.chan_grad[d] = p->grad_out[g] ? p->grad_out[g] + c : nullptr;
                b.chan_stride[d] = p->out_channels[g];
            }
            if (p->grad_out[g]) b.any_grad_im = 1;
        }
        if (d != p->num_attr) return fail("out_channels sum %d != num_attr %d", d, p->num_attr);
    }
    b.grad_improb = p->grad_improb;
    b.grad_points2d = p->grad_points2d; b.grad_face_attr = p->grad_face_attr;
    b.attr_compact = attr_compact(p) ? 1 : 0;
    g_launches += ((parts & 1) ? 1 : 0) + ((parts & 2) ? 1 : 0);     // prepare_backward_kernel, backward_faces_kernel
    return cuda_fail("dibr_backward_faces", dibr::launch_backward_faces(b, (cudaStream_t)stream, parts));
}

int dibr_backward_faces(const DibrPass* p, void* stream) { return backward_faces_impl(p, stream, 3); }

static int backward_meshes_impl(const DibrPass* p, void* stream, float* packed, int pose_sum = 0, float* host_view = nullptr) {
    if (int e = check_common(p, true)) return e;
    if (p->num_instances <= 0 || !p->inst_desc || !p->verts) return fail("backward_meshes: instances and verts required");
    if (!p->pose_R && (!p->cam_rot || !p->cam_pos || !p->cam_proj)) return fail("backward_meshes: cameras required");
    if (p->pose_R && (!p->pose_t || !p->grad_pose_R || !p->grad_pose_t)) return fail("backward_meshes: pose mode needs pose_t, grad_pose_R, grad_pose_t");
    if (!p->grad_points2d || !p->grad_face_attr || !p->vert_face_ptr || !p->vert_face_idx) return fail("backward_meshes: face grads and vertex adjacency required");
    if (!p->pose_R && (!p->grad_cam_rot || !p->grad_cam_pos)) return fail("backward_meshes: grad_cam_rot/grad_cam_pos required");
    const dibr::Workspace w = carve(p, p->workspace);
    dibr::MeshBwdParams m;
    memset(&m, 0, sizeof(m));
    m.num_instances = p->num_instances; m.inst_desc = p->inst_desc; m.verts = p->verts;
    m.verts_stride = p->verts_stride > 0 ? p->verts_stride : 3;
    m.cam_rot = p->pose_R ? w.cam_rot : p->cam_rot; m.cam_pos = p->pose_R ? w.cam_pos : p->cam_pos;
    m.cam_proj = p->pose_R ? w.cam_proj : p->cam_proj;
    m.pose_R = p->pose_R; m.pose_t = p->pose_t; m.grad_pose_R = p->grad_pose_R; m.grad_pose_t = p->grad_pose_t;
    m.vert_attr_dim = p->vert_attr_dim; m.attr_flags = p->attr_flags; m.num_attr = p->num_attr;
    m.grad_points2d = p->grad_points2d; m.grad_face_attr = p->grad_face_attr; m.attr_compact = attr_compact(p) ? 1 : 0;
    m.vert_face_ptr = p->vert_face_ptr; m.vert_face_idx = p->vert_face_idx;
    m.grad_verts = p->grad_verts; m.grad_vert_attr = p->grad_vert_attr;
    m.grad_cam_rot = p->grad_cam_rot; m.grad_cam_pos = p->grad_cam_pos; m.pose_part = w.pose_part; m.pose_done = w.pose_done;
    m.grad_pose_packed = p->pose_R ? packed : nullptr;
    m.host_pose_packed = (p->pose_R && packed) ? host_view : nullptr;
    m.pose_sum = pose_sum;
    g_launches += 1;
    return cuda_fail("dibr_backward_meshes", dibr::launch_backward_meshes(m, (cudaStream_t)stream));
}

int dibr_backward_meshes(const DibrPass* p, void* stream) { return backward_meshes_impl(p, stream, nullptr); }

int dibr_normal_map(const float* normals, const float* mask, const uint32_t* min_ordered, float* out, long long npix, void* stream) {
    if (!normals || !mask || !min_ordered || !out || npix < 0) return fail("normal_map: null argument");
    if (dibr_device_count() <= 0) return fail("no CUDA device: libdibr_b200 has no CPU fallback");
    g_launches += 1;
    return cuda_fail("dibr_normal_map", dibr::launch_normal_map(normals, mask, min_ordered, out, npix, (cudaStream_t)stream));
}

int dibr_normal_map_pass(const DibrPass* p, const float* normals, const float* mask, float* out, void* stream) {
    if (int e = check_common(p, true)) return e;
    if (!normals || !mask || !out || !p->out_min_ordered) return fail("normal_map_pass: null argument (the pass needs out_min_ordered)");
    const dibr::Workspace w = carve(p, p->workspace);
    g_launches += 1;
    return cuda_fail("dibr_normal_map_pass", dibr::launch_normal_map_tiles(p->batch, p->height, p->width, w.order_cnt, w.order_seg, normals, mask,
                                                                           p->out_min_ordered, out, (cudaStream_t)stream));
}

static int nnd_params(const DibrNnd* p, dibr::NndParams& n, bool backward) {
    if (!p) return fail("null DibrNnd");
    if (p->batch < 0 || p->stride1 < 0 || p->stride2 < 0) return fail("nnd: negative sizes");
    if (!p->xyz1 || !p->xyz2 || !p->idx1 || !p->idx2) return fail("nnd: xyz / idx buffers required");
    if (!backward && (!p->dist1 || !p->dist2)) return fail("nnd_forward: dist buffers required");
    if (backward && (!p->graddist1 || !p->graddist2 || !p->gradxyz1 || !p->gradxyz2)) return fail("nnd_backward: gradient buffers required");
    if (dibr_device_count() <= 0) return fail("no CUDA device: libdibr_b200 has no CPU fallback");
    n.batch = p->batch; n.stride1 = p->stride1; n.stride2 = p->stride2; n.count1 = p->count1; n.count2 = p->count2;
    n.xyz1 = p->xyz1; n.xyz2 = p->xyz2; n.dist1 = p->dist1; n.dist2 = p->dist2; n.idx1 = p->idx1; n.idx2 = p->idx2;
    n.graddist1 = p->graddist1; n.graddist2 = p->graddist2; n.gradxyz1 = p->gradxyz1; n.gradxyz2 = p->gradxyz2;
    return 0;
}

int dibr_nnd_workspace_bytes(const DibrNnd* p, size_t* bytes) {
    if (!p || !bytes) return fail("null argument");
    if (p->batch < 0 || p->stride1 < 0 || p->stride2 < 0) return fail("nnd: negative sizes");
    *bytes = dibr::nnd_grid_workspace_bytes(p->batch, p->stride1, p->stride2);
    return 0;
}

static int nnd_check_ws(const DibrNnd* p) {
    if (((uintptr_t)p->workspace & 255u) != 0) return fail("nnd: workspace must be 256-byte aligned");
    const size_t need = dibr::nnd_grid_workspace_bytes(p->batch, p->stride1, p->stride2);
    if (p->workspace_bytes < need) return fail("nnd: workspace too small: %zu < %zu", p->workspace_bytes, need);
    return 0;
}

int dibr_nnd_forward(const DibrNnd* p, void* stream) {
    dibr::NndParams n;
    if (int e = nnd_params(p, n, false)) return e;
    if (p->workspace) {
        if (int e = nnd_check_ws(p)) return e;
        g_launches += 5;
        return cuda_fail("dibr_nnd_forward", dibr::launch_nnd_forward_grid(n, p->workspace, (cudaStream_t)stream));
    }
    g_launches += 2;
    return cuda_fail("dibr_nnd_forward", dibr::launch_nnd_forward(n, (cudaStream_t)stream));
}

int dibr_nnd_backward(const DibrNnd* p, void* stream) {
    dibr::NndParams n;
    if (int e = nnd_params(p, n, true)) return e;
    if (p->workspace) {
        if (int e = nnd_check_ws(p)) return e;
        g_launches += 5;
        return cuda_fail("dibr_nnd_backward", dibr::launch_nnd_backward_grid(n, p->workspace, (cudaStream_t)stream));
    }
    g_launches += 2;
    return cuda_fail("dibr_nnd_backward", dibr::launch_nnd_backward(n, (cudaStream_t)stream));
}

static int bp_params(const DibrBackproject* p, dibr::BackprojectParams& q, bool backward) {
    if (!p) return fail("null DibrBackproject");
    if (p->batch < 0 || p->height <= 0 || p->width <= 0) return fail("backproject: bad sizes");
    if (p->num_K != 1 && p->num_K != p->batch) return fail("backproject: num_K must be 1 or batch");
    if (!p->K || !p->slot) return fail("backproject: K / slot required");
    if (!backward && (!p->depth || !p->points || !p->count || !p->chunk_count)) return fail("backproject: depth, points, count, chunk_count required");
    if (backward && (!p->grad_points || !p->grad_depth)) return fail("backproject backward: grad_points / grad_depth required");
    if ((long long)p->batch * p->height * p->width >= (1ll << 31)) return fail("backproject: too many pixels");
    if (dibr_device_count() <= 0) return fail("no CUDA device: libdibr_b200 has no CPU fallback");
    q.batch = p->batch; q.height = p->height; q.width = p->width; q.num_K = p->num_K;
    q.depth = p->depth; q.K = p->K; q.points = p->points; q.count = p->count; q.slot = p->slot; q.chunk_count = p->chunk_count;
    q.grad_points = p->grad_points; q.grad_depth = p->grad_depth;
    return 0;
}

int dibr_backproject_compact(const DibrBackproject* p, void* stream) {
    dibr::BackprojectParams q;
    if (int e = bp_params(p, q, false)) return e;
    g_launches += 2;
    return cuda_fail("dibr_backproject_compact", dibr::launch_backproject(q, (cudaStream_t)stream));
}

int dibr_backproject_compact_backward(const DibrBackproject* p, void* stream) {
    dibr::BackprojectParams q;
    if (int e = bp_params(p, q, true)) return e;
    g_launches += 1;
    return cuda_fail("dibr_backproject_compact_backward", dibr::launch_backproject_backward(q, (cudaStream_t)stream));
}

int dibr_mask_loss_scratch_floats(int64_t n) { return dibr::mask_loss_partial_floats(n) + 1; }

static int ml_params(const DibrMaskLoss* p, dibr::MaskLossParams& q, bool backward) {
    if (!p) return fail("null DibrMaskLoss");
    if (p->n < 0) return fail("mask_loss: negative size");
    if (!p->probs || !p->target || !p->out) return fail("mask_loss: probs / target / out required");
    if (!backward && !p->scratch) return fail("mask_loss: scratch required");
    if (backward && (!p->grad_out || !p->grad_probs)) return fail("mask_loss backward: grad_out / grad_probs required");
    if (dibr_device_count() <= 0) return fail("no CUDA device: libdibr_b200 has no CPU fallback");
    q.n = p->n; q.probs = p->probs; q.target = p->target; q.weight = p->weight;
    q.partial = p->scratch; q.ticket = p->scratch ? (unsigned int*)(p->scratch + dibr::mask_loss_partial_floats(p->n)) : nullptr;
    q.out = p->out; q.grad_out = p->grad_out; q.grad_probs = p->grad_probs;
    return 0;
}
int dibr_mask_loss_forward(const DibrMaskLoss* p, void* stream) {
    dibr::MaskLossParams q;
    if (int e = ml_params(p, q, false)) return e;
    g_launches += 1;
    return cuda_fail("dibr_mask_loss_forward", dibr::launch_mask_loss_forward(q, (cudaStream_t)stream));
}
int dibr_mask_loss_backward(const DibrMaskLoss* p, void* stream) {
    dibr::MaskLossParams q;
    if (int e = ml_params(p, q, true)) return e;
    g_launches += 1;
    return cuda_fail("dibr_mask_loss_backward", dibr::launch_mask_loss_backward(q, (cudaStream_t)stream));
}

int dibr_lab_loss_scratch_floats(int64_t pixels) { return dibr::lab_loss_partial_floats(pixels) + 1; }

static int lab_params(const DibrLabLoss* p, dibr::LabLossParams& q, bool backward) {
    if (!p) return fail("null DibrLabLoss");
    if (p->n_img < 0 || p->hw < 0) return fail("lab_loss: negative size");
    if (!p->gt || !p->ren || !p->out) return fail("lab_loss: gt / ren / out required");
    if (!backward && !p->scratch) return fail("lab_loss: scratch required");
    if (backward && (!p->grad_out || !p->grad_ren)) return fail("lab_loss backward: grad_out / grad_ren required");
    if (dibr_device_count() <= 0) return fail("no CUDA device: libdibr_b200 has no CPU fallback");
    const long long pixels = (long long)p->n_img * p->hw;
    q.n_img = p->n_img; q.hw = p->hw; q.bgr = p->bgr; q.no_l = p->no_l;
    q.gt = p->gt; q.ren = p->ren; q.mask = p->mask;
    q.partial = p->scratch; q.ticket = p->scratch ? (unsigned int*)(p->scratch + dibr::lab_loss_partial_floats(pixels)) : nullptr;
    q.out = p->out; q.grad_out = p->grad_out; q.grad_ren = p->grad_ren;
    return 0;
}
int dibr_lab_loss_forward(const DibrLabLoss* p, void* stream) {
    dibr::LabLossParams q;
    if (int e = lab_params(p, q, false)) return e;
    g_launches += 1;
    return cuda_fail("dibr_lab_loss_forward", dibr::launch_lab_loss_forward(q, (cudaStream_t)stream));
}
int dibr_lab_loss_backward(const DibrLabLoss* p, void* stream) {
    dibr::LabLossParams q;
    if (int e = lab_params(p, q, true)) return e;
    g_launches += 1;
    return cuda_fail("dibr_lab_loss_backward", dibr::launch_lab_loss_backward(q, (cudaStream_t)stream));
}

namespace {
struct SsimPlan {
    int levels, planes, pad;
    int H[dibr::SSIM_MAX_LEVELS], W[dibr::SSIM_MAX_LEVELS], tiles[dibr::SSIM_MAX_LEVELS];
    size_t px[dibr::SSIM_MAX_LEVELS], py[dibr::SSIM_MAX_LEVELS], gp[dibr::SSIM_MAX_LEVELS];      // float offsets (level 0 unused)
    size_t maps[dibr::SSIM_MAX_LEVELS], partial[dibr::SSIM_MAX_LEVELS], scale;
    size_t total_floats;
};
}  // namespace

static int ssim_plan(const DibrMsSsim* p, SsimPlan& s) {
    if (!p) return fail("null DibrMsSsim");
    if (p->n_img < 0 || p->channels <= 0) return fail("ms_ssim: bad n_img / channels");
    if (p->levels < 2 || p->levels > dibr::SSIM_MAX_LEVELS) return fail("ms_ssim: levels must be in [2, 8]");
    if ((long long)p->n_img * p->channels > 65535) return fail("ms_ssim: n_img * channels must be <= 65535");
    s.levels = p->levels; s.planes = p->n_img * p->channels;
    int h = p->height, w = p->width;
    const int pad = p->use_padding ? 5 : 0;               // ssim.py:42-45: window_size // 2
    s.pad = pad;
    size_t off = 0;
    for (int l = 0; l < s.levels; l++) {
        if (!pad && (h < 11 || w < 11)) return fail("ms_ssim: every level must keep at least 11x11 pixels (ssim.py valid convolution)");
        if (h < 1 || w < 1) return fail("ms_ssim: a level has no pixels");
        s.H[l] = h; s.W[l] = w; s.tiles[l] = dibr::ssim_forward_tiles(h, w, pad);
        const size_t img = (size_t)s.planes * h * w;
        s.px[l] = off; if (l) off += img;
        s.py[l] = off; if (l) off += img;
        s.gp[l] = off; if (l && p->want_grad) off += img;
        s.maps[l] = off; if (p->want_grad) off += (size_t)s.planes * 3 * (h - 10 + 2 * pad) * (w - 10 + 2 * pad);
        off = (off + 3) & ~(size_t)3;                // the combine kernel reads the partials as float2
        s.partial[l] = off; off += (size_t)s.planes * s.tiles[l] * 2;
        h = (h + 2 * (h & 1) - 2) / 2 + 1;          // avg_pool2d(kernel 2, stride 2, padding h % 2)
        w = (w + 2 * (w & 1) - 2) / 2 + 1;
    }
    s.scale = off; off += (size_t)s.levels * (p->n_img > 0 ? p->n_img : 1);
    s.total_floats = off;
    return 0;
}
int dibr_ms_ssim_workspace_bytes(const DibrMsSsim* p, size_t* bytes) {
    SsimPlan s;
    if (int e = ssim_plan(p, s)) return e;
    if (!bytes) return fail("ms_ssim: null bytes");
    *bytes = s.total_floats * sizeof(float);
    return 0;
}
static void ssim_level(const DibrMsSsim* p, const SsimPlan& s, int l, dibr::SsimLevelParams& q) {
    float* ws = (float*)p->workspace;
    q = dibr::SsimLevelParams{};
    q.H = s.H[l]; q.W = s.W[l]; q.channels = p->channels; q.use_ssim = (l == s.levels - 1); q.pad = s.pad;
    const double k1 = 0.01 * (double)p->data_range, k2 = 0.03 * (double)p->data_range;      // ssim.py:72-73, python doubles
    q.C1 = (float)(k1 * k1);
    q.C2 = (float)(k2 * k2);
    for (int k = 0; k < 11; k++) q.win[k] = p->window[k];
    q.x = l ? ws + s.px[l] : p->x;
    q.y = l ? ws + s.py[l] : p->y;
    q.maps = p->want_grad ? ws + s.maps[l] : nullptr;
    q.partial = ws + s.partial[l];
}
int dibr_ms_ssim_forward(const DibrMsSsim* p, void* stream) {
    SsimPlan s;
    if (int e = ssim_plan(p, s)) return e;
    if (!p->x || !p->y || !p->out || !p->workspace) return fail("ms_ssim: x / y / out / workspace required");
    if (p->workspace_bytes < s.total_floats * sizeof(float)) return fail("ms_ssim: workspace too small");
    if (dibr_device_count() <= 0) return fail("no CUDA device: libdibr_b200 has no CPU fallback");
    if (p->n_img == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    float* ws = (float*)p->workspace;
    dibr::SsimCombineParams c{};
    c.n_img = p->n_img; c.channels = p->channels; c.levels = s.levels; c.normalize = p->normalize;
    for (int l = 0; l < s.levels; l++) {
        dibr::SsimLevelParams q;
        ssim_level(p, s, l, q);
        if (l) {
            dibr::SsimLevelParams prev;
            ssim_level(p, s, l - 1, prev);
            g_launches += 1;
            if (int e = cuda_fail("dibr_ms_ssim_forward(pool)", dibr::launch_ssim_pool(prev.x, prev.y, ws + s.px[l], ws + s.py[l], s.planes,
                                                                                     s.H[l - 1], s.W[l - 1], s.H[l], s.W[l], st))) return e;
        }
        g_launches += 1;
        if (int e = cuda_fail("dibr_ms_ssim_forward(level)", dibr::launch_ssim_level_forward(q, s.planes, st))) return e;
        c.weights[l] = p->weights[l]; c.tiles[l] = s.tiles[l]; c.map_pixels[l] = (s.H[l] - 10 + 2 * s.pad) * (s.W[l] - 10 + 2 * s.pad);
        c.partial_off[l] = (long long)s.partial[l];
    }
    c.partial = ws; c.out = p->out; c.scale = p->want_grad ? ws + s.scale : nullptr;
    g_launches += 1;
    return cuda_fail("dibr_ms_ssim_forward(combine)", dibr::launch_ssim_combine(c, st));
}
int dibr_ms_ssim_backward(const DibrMsSsim* p, void* stream) {
    SsimPlan s;
    if (int e = ssim_plan(p, s)) return e;
    if (!p->want_grad) return fail("ms_ssim backward: the forward must have run with want_grad");
    if (!p->x || !p->y || !p->workspace || !p->grad_out || !p->grad_y) return fail("ms_ssim backward: x / y / workspace / grad_out / grad_y required");
    if (p->workspace_bytes < s.total_floats * sizeof(float)) return fail("ms_ssim: workspace too small");
    if (dibr_device_count() <= 0) return fail("no CUDA device: libdibr_b200 has no CPU fallback");
    if (p->n_img == 0) return 0;
    float* ws = (float*)p->workspace;
    for (int l = s.levels - 1; l >= 0; l--) {
        dibr::SsimLevelParams q;
        ssim_level(p, s, l, q);
        q.scale = ws + s.scale + (size_t)l * p->n_img;
        q.grad_out = p->grad_out;
        if (l < s.levels - 1) { q.grad_coarse = ws + s.gp[l + 1]; q.Hc = s.H[l + 1]; q.Wc = s.W[l + 1]; }
        q.grad_y = l ? ws + s.gp[l] : p->grad_y;
        g_launches += 1;
        if (int e = cuda_fail("dibr_ms_ssim_backward(level)", dibr::launch_ssim_level_backward(q, s.planes, (cudaStream_t)stream))) return e;
    }
    return 0;
}

static int cr_params(const DibrChamferReduce* p, dibr::ChamferReduceParams& q, bool backward) {
    if (!p) return fail("null DibrChamferReduce");
    if (p->batch < 0 || p->stride1 < 0 || p->stride2 < 0) return fail("chamfer_reduce: negative sizes");
    if (!p->dist1 || !p->dist2 || !p->stats || !p->out) return fail("chamfer_reduce: dist / stats / out required");
    if (!backward && !p->ticket) return fail("chamfer_reduce: ticket required");
    if (backward && (!p->grad_out || !p->grad_dist1 || !p->grad_dist2)) return fail("chamfer_reduce backward: gradient buffers required");
    if (dibr_device_count() <= 0) return fail("no CUDA device: libdibr_b200 has no CPU fallback");
    q.batch = p->batch; q.stride1 = p->stride1; q.stride2 = p->stride2; q.threshold = p->threshold;
    q.count1 = p->count1; q.count2 = p->count2; q.dist1 = p->dist1; q.dist2 = p->dist2;
    q.stats = p->stats; q.ticket = p->ticket; q.out = p->out;
    q.grad_out = p->grad_out; q.grad_dist1 = p->grad_dist1; q.grad_dist2 = p->grad_dist2;
    return 0;
}
int dibr_chamfer_reduce_forward(const DibrChamferReduce* p, void* stream) {
    dibr::ChamferReduceParams q;
    if (int e = cr_params(p, q, false)) return e;
    g_launches += 1;
    return cuda_fail("dibr_chamfer_reduce_forward", dibr::launch_chamfer_reduce_forward(q, (cudaStream_t)stream));
}
int dibr_chamfer_reduce_backward(const DibrChamferReduce* p, void* stream) {
    dibr::ChamferReduceParams q;
    if (int e = cr_params(p, q, true)) return e;
    g_launches += 1;
    return cuda_fail("dibr_chamfer_reduce_backward", dibr::launch_chamfer_reduce_backward(q, (cudaStream_t)stream));
}

// The step's two host transfers are a few KB each (5.9 KB of poses / intrinsics / instance table in, 1.5 KB of pose gradients
// out).  When the host buffer is pinned and mapped (every cudaHostAlloc'd buffer under unified addressing, torch's pinned
// tensors included) a tiny kernel moves the words over the bus itself: the copy-engine hand-over of a cudaMemcpyAsync node
// costs more than the transfer.  Anything else (pageable memory, unregistered pointers) goes through cudaMemcpyAsync.
__global__ void copy_words_kernel(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[i];
}

static bool device_visible_host(const void* host, const void** dev_view) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, host) != cudaSuccess) { cudaGetLastError(); return false; }
    if (a.type != cudaMemoryTypeHost || !a.devicePointer) return false;
    *dev_view = a.devicePointer;
    return true;
}

// bytes: multiple of 4, both pointers 4-byte aligned (the staging block and the gradient rows are arrays of 32-bit words)
static int small_transfer(void* dst, const void* src, size_t bytes, cudaMemcpyKind kind, cudaStream_t s) {
    const void* view = nullptr;
    const bool words = (bytes & 3) == 0 && bytes <= (1u << 20) && (((uintptr_t)dst | (uintptr_t)src) & 3) == 0;
    if (words && device_visible_host(kind == cudaMemcpyHostToDevice ? src : dst, &view)) {
        const int n = (int)(bytes >> 2);
        const uint32_t* sp = (const uint32_t*)(kind == cudaMemcpyHostToDevice ? view : src);
        uint32_t* dp = (uint32_t*)(kind == cudaMemcpyHostToDevice ? dst : const_cast<void*>(view));
        copy_words_kernel<<<(n + 255) / 256, 256, 0, s>>>(sp, dp, n);
        g_launches += 1;
        return (int)cudaGetLastError();
    }
    return (int)cudaMemcpyAsync(dst, src, bytes, kind, s);
}

// side stream + fork/join events of one session (dibr_overlap_create): owned by the caller, nothing process-global
struct Overlap { cudaStream_t stream; cudaEvent_t fork, join; };

int dibr_overlap_create(void** handle) {
    if (!handle) return fail("overlap_create: null handle");
    if (dibr_device_count() <= 0) return fail("no CUDA device: libdibr_b200 has no CPU fallback");
    Overlap* o = new Overlap();
    int least = 0, greatest = 0;
    cudaDeviceGetStreamPriorityRange(&least, &greatest);
    // the side stream carries the LONG chain (student forward + backward) at high priority; the short teacher chain stays on
    // the caller's stream and fills in behind it
    cudaError_t e = cudaStreamCreateWithPriority(&o->stream, cudaStreamNonBlocking, greatest);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&o->fork, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&o->join, cudaEventDisableTiming);
    if (e != cudaSuccess) { delete o; return cuda_fail("overlap_create", (int)e); }
    *handle = o;
    return 0;
}

int dibr_overlap_destroy(void* handle) {
    if (!handle) return 0;
    Overlap* o = (Overlap*)handle;
    cudaEventDestroy(o->fork); cudaEventDestroy(o->join); cudaStreamDestroy(o->stream);
    delete o;
    return 0;
}

namespace {
// Fork the caller's stream into the session's side stream; the destructor joins again, so every return path of the entry
// points below (errors included) leaves `cs` waiting for whatever was enqueued on the side stream.
struct ForkJoin {
    Overlap* o;
    cudaStream_t cs;
    int err;
    ForkJoin(Overlap* o_, cudaStream_t cs_) : o(o_), cs(cs_), err(0) {
        if (!o) return;
        cudaError_t e = cudaEventRecord(o->fork, cs);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(o->stream, o->fork, 0);
        if (e != cudaSuccess) { err = (int)e; o = nullptr; }
    }
    ~ForkJoin() {
        if (!o) return;
        if (cudaEventRecord(o->join, o->stream) == cudaSuccess) cudaStreamWaitEvent(cs, o->join, 0);
    }
    void* side(void* fallback) const { return o ? (void*)o->stream : fallback; }
};

int render_forward_on(const DibrStep* st, void* stream, const ForkJoin& fj) {
    const DibrPass* passes[2] = {&st->student, &st->teacher};
    const float* nin[2] = {st->student_normal_in, st->teacher_normal_in};
    const float* nmask[2] = {st->student_mask_in, st->teacher_mask_in};
    float* nout[2] = {st->student_normal_out, st->teacher_normal_out};
    for (int k = 0; k < 2; k++) {
        const DibrPass* p = passes[k];
        if (p->num_instances <= 0) continue;
        void* ks = (k == 0) ? fj.side(stream) : stream;          // student chain on the side stream, teacher on the caller's
        if (int e = dibr_setup_meshes(p, ks)) return e;
        if (int e = forward_impl(p, ks, true)) return e;          // the set-up call above cleared the lists and the minimum
        // the backward's preparation (zeroed gradients, work lists from the flags the forward just set) depends on no upstream
        // gradient: here it runs in the shadow of the teacher rasterisation instead of at the head of the backward call
        if (k == 0 && (st->run_backward & 2))
            if (int e = backward_faces_impl(p, ks, 1)) return e;
        if (nin[k]) {
            if (!p->out_min_ordered || p->min_output < 0) return fail("render_step: normal map needs min_output/out_min_ordered");
            if (int e = dibr_normal_map_pass(p, nin[k], nmask[k], nout[k], ks)) return e;
        }
    }
    return 0;
}

int render_backward_on(const DibrStep* st, void* ls) {
    const DibrPass* p = &st->student;
    if (int e = backward_faces_impl(p, ls, (st->run_backward & 2) ? 2 : 3)) return e;
    if (st->device_grad_pose && (!p->grad_pose_R || !p->grad_pose_t)) return fail("render_step: pose-gradient buffers are null");
    // the kernel's finalising block writes the [n,12] layout itself: no packing launch -- and, when the host buffer is pinned
    // and mapped, straight into it: no copy behind the kernel either
    const void* hv = nullptr;
    const bool direct = st->device_grad_pose && st->host_grad_pose && p->pose_R && ((uintptr_t)st->host_grad_pose & 3) == 0 &&
                        device_visible_host(st->host_grad_pose, &hv);
    if (int e = backward_meshes_impl(p, ls, st->device_grad_pose, st->grad_pose_sum, direct ? (float*)const_cast<void*>(hv) : nullptr)) return e;
    if (st->device_grad_pose && st->host_grad_pose && !direct) {
        const int e = small_transfer(st->host_grad_pose, st->device_grad_pose, sizeof(float) * 12 * (size_t)p->num_instances,
                                     cudaMemcpyDeviceToHost, (cudaStream_t)ls);
        if (e != 0) return cuda_fail("render_step pose gradients", e);
    }
    return 0;
}

int render_upload(const DibrStep* st, cudaStream_t cs) {
    if (st->staging_bytes > 0) {
        if (!st->staging_host || !st->staging_device) return fail("render_step: staging buffers are null");
        const int e = small_transfer(st->staging_device, st->staging_host, st->staging_bytes, cudaMemcpyHostToDevice, cs);
        if (e != 0) return cuda_fail("render_step H2D", e);
    }
    return 0;
}
}  // namespace

// The teacher rasterisation depends on nothing the student pass or the backward produce (and vice versa).  The long
// chain (student set-up, forward, backward, pose-gradient read-back) goes to the session's high-priority side stream,
// forked after the staging copy; the teacher pass stays on the caller's stream and its CTAs fill the tails and stalls of
// the long chain.  The caller's stream waits for the side stream before the call returns, so for the caller everything
// is ordered in `stream` as before.
int dibr_render_forward(const DibrStep* st, void* stream) {
    if (!st) return fail("null DibrStep");
    if (int e = render_upload(st, (cudaStream_t)stream)) return e;
    const bool both = st->student.num_instances > 0 && st->teacher.num_instances > 0;
    ForkJoin fj(both ? (Overlap*)st->overlap : nullptr, (cudaStream_t)stream);
    if (fj.err) return cuda_fail("render_forward fork", fj.err);
    return render_forward_on(st, stream, fj);
}

int dibr_render_backward(const DibrStep* st, void* stream) {
    if (!st) return fail("null DibrStep");
    return render_backward_on(st, stream);
}

int dibr_render_step(const DibrStep* st, void* stream) {
    if (!st) return fail("null DibrStep");
    if (int e = render_upload(st, (cudaStream_t)stream)) return e;
    const bool both = st->student.num_instances > 0 && st->teacher.num_instances > 0;
    ForkJoin fj(both ? (Overlap*)st->overlap : nullptr, (cudaStream_t)stream);
    if (fj.err) return cuda_fail("render_step fork", fj.err);
    if (int e = render_forward_on(st, stream, fj)) return e;
    if (st->run_backward & 1) return render_backward_on(st, fj.side(stream));   // the student chain's stream
    return 0;
}

}  // extern "C"
