"""ctypes binding of libdibr_b200.so (C ABI declared in include/dibr_b200.h).

The library is built in-tree by ``self6dpp_b200/csrc/Makefile`` (``__graft_entry__.build()`` runs
it).  There is deliberately no fallback of any kind: if the shared object is missing, or the
process has no CUDA device, the compute entry points raise.
"""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DIBR_B200_LIB") or os.path.join(_HERE, "lib", "libdibr_b200.so")   # override: debug builds (tools/variants.py)
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "dibr_b200.h")

_c_f32p = ctypes.c_void_p
_c_i32p = ctypes.c_void_p


class DibrPass(ctypes.Structure):
    """Mirror of ``struct DibrPass`` -- field order and types must match include/dibr_b200.h."""

    _fields_ = [
        ("batch", ctypes.c_int32), ("height", ctypes.c_int32), ("width", ctypes.c_int32),
        ("num_attr", ctypes.c_int32), ("knum", ctypes.c_int32), ("multiplier", ctypes.c_int32),
        ("delta", ctypes.c_int32), ("expand", ctypes.c_float), ("total_faces", ctypes.c_int32),
        ("faces_per_image", ctypes.c_int32), ("face_offsets", _c_i32p),
        ("points3d", _c_f32p), ("points2d", _c_f32p), ("normalz", _c_f32p),
        ("num_instances", ctypes.c_int32), ("inst_desc", _c_i32p), ("verts", _c_f32p),
        ("mesh_faces", _c_i32p), ("vert_attr", _c_f32p), ("vert_attr_dim", ctypes.c_int32),
        ("attr_flags", ctypes.c_int32), ("cam_rot", _c_f32p), ("cam_pos", _c_f32p), ("cam_proj", _c_f32p),
        ("pose_R", _c_f32p), ("pose_t", _c_f32p), ("pose_K", _c_f32p), ("num_K", ctypes.c_int32),
        ("znear", ctypes.c_float), ("zfar", ctypes.c_float), ("grad_pose_R", _c_f32p), ("grad_pose_t", _c_f32p),
        ("workspace", ctypes.c_void_p), ("workspace_bytes", ctypes.c_size_t),
        ("face_attr", _c_f32p), ("face_normal", _c_f32p),
        ("im", _c_f32p), ("improb", _c_f32p), ("imidx", _c_i32p), ("imcomp", _c_f32p),
        ("grad_im", _c_f32p), ("grad_improb", _c_f32p), ("grad_points2d", _c_f32p),
        ("grad_face_attr", _c_f32p), ("grad_verts", _c_f32p), ("grad_vert_attr", _c_f32p),
        ("grad_cam_rot", _c_f32p), ("grad_cam_pos", _c_f32p),
        ("num_outputs", ctypes.c_int32), ("out_channels", ctypes.c_int32 * 6),
        ("out", ctypes.c_void_p * 6), ("grad_out", ctypes.c_void_p * 6),
        ("min_output", ctypes.c_int32), ("out_min_ordered", ctypes.c_void_p),
        ("vert_face_ptr", _c_i32p), ("vert_face_idx", _c_i32p),
        ("num_cams", ctypes.c_int32), ("verts_stride", ctypes.c_int32), ("vert_attr_stride", ctypes.c_int32),
        ("reserved0", ctypes.c_int32),
    ]


class DibrStep(ctypes.Structure):
    """Mirror of ``struct DibrStep`` (include/dibr_b200.h)."""

    _fields_ = [
        ("student", DibrPass), ("teacher", DibrPass),
        ("staging_host", ctypes.c_void_p), ("staging_device", ctypes.c_void_p), ("staging_bytes", ctypes.c_size_t),
        ("student_normal_in", _c_f32p), ("student_mask_in", _c_f32p), ("student_normal_out", _c_f32p),
        ("teacher_normal_in", _c_f32p), ("teacher_mask_in", _c_f32p), ("teacher_normal_out", _c_f32p),
        ("run_backward", ctypes.c_int32), ("grad_pose_sum", ctypes.c_int32),
        ("host_grad_pose", _c_f32p), ("device_grad_pose", _c_f32p),
        ("overlap", ctypes.c_void_p),
    ]


class DibrNnd(ctypes.Structure):
    """Mirror of ``struct DibrNnd`` (include/dibr_b200.h)."""

    _fields_ = [
        ("batch", ctypes.c_int32), ("stride1", ctypes.c_int32), ("stride2", ctypes.c_int32), ("reserved", ctypes.c_int32),
        ("count1", _c_i32p), ("count2", _c_i32p), ("xyz1", _c_f32p), ("xyz2", _c_f32p),
        ("dist1", _c_f32p), ("dist2", _c_f32p), ("idx1", _c_i32p), ("idx2", _c_i32p),
        ("graddist1", _c_f32p), ("graddist2", _c_f32p), ("gradxyz1", _c_f32p), ("gradxyz2", _c_f32p),
        ("workspace", ctypes.c_void_p), ("workspace_bytes", ctypes.c_size_t),
    ]


class DibrBackproject(ctypes.Structure):
    """Mirror of ``struct DibrBackproject`` (include/dibr_b200.h)."""

    _fields_ = [
        ("batch", ctypes.c_int32), ("height", ctypes.c_int32), ("width", ctypes.c_int32), ("num_K", ctypes.c_int32),
        ("depth", _c_f32p), ("K", _c_f32p), ("points", _c_f32p), ("count", _c_i32p), ("slot", _c_i32p),
        ("chunk_count", _c_i32p), ("grad_points", _c_f32p), ("grad_depth", _c_f32p),
    ]


class DibrMaskLoss(ctypes.Structure):
    """Mirror of ``struct DibrMaskLoss`` (include/dibr_b200.h)."""

    _fields_ = [
        ("n", ctypes.c_int64), ("probs", _c_f32p), ("target", _c_f32p), ("weight", _c_f32p),
        ("scratch", _c_f32p), ("out", _c_f32p), ("grad_out", _c_f32p), ("grad_probs", _c_f32p),
    ]


class DibrLabLoss(ctypes.Structure):
    """Mirror of ``struct DibrLabLoss`` (include/dibr_b200.h)."""

    _fields_ = [
        ("n_img", ctypes.c_int32), ("hw", ctypes.c_int32), ("bgr", ctypes.c_int32), ("no_l", ctypes.c_int32),
        ("gt", _c_f32p), ("ren", _c_f32p), ("mask", _c_f32p),
        ("scratch", _c_f32p), ("out", _c_f32p), ("grad_out", _c_f32p), ("grad_ren", _c_f32p),
    ]


class DibrMsSsim(ctypes.Structure):
    """Mirror of ``struct DibrMsSsim`` (include/dibr_b200.h)."""

    _fields_ = [
        ("n_img", ctypes.c_int32), ("channels", ctypes.c_int32), ("height", ctypes.c_int32), ("width", ctypes.c_int32),
        ("levels", ctypes.c_int32), ("normalize", ctypes.c_int32), ("want_grad", ctypes.c_int32), ("use_padding", ctypes.c_int32),
        ("data_range", ctypes.c_float), ("window", ctypes.c_float * 11), ("weights", ctypes.c_float * 8),
        ("x", _c_f32p), ("y", _c_f32p), ("workspace", ctypes.c_void_p), ("workspace_bytes", ctypes.c_size_t),
        ("out", _c_f32p), ("grad_out", _c_f32p), ("grad_y", _c_f32p),
    ]


class DibrDiceLoss(ctypes.Structure):
    """Mirror of ``struct DibrDiceLoss`` (include/dibr_b200.h)."""

    _fields_ = [
        ("num", ctypes.c_int32), ("reduction", ctypes.c_int32), ("per", ctypes.c_int64),
        ("smooth", ctypes.c_float), ("eps", ctypes.c_float),
        ("probs", _c_f32p), ("labels", _c_f32p), ("stats", _c_f32p), ("ticket", ctypes.c_void_p),
        ("out", _c_f32p), ("grad_out", _c_f32p), ("grad_probs", _c_f32p),
    ]


class DibrNormLoss(ctypes.Structure):
    """Mirror of ``struct DibrNormLoss`` (include/dibr_b200.h)."""

    _fields_ = [
        ("n_img", ctypes.c_int32), ("hw", ctypes.c_int32), ("with_l1", ctypes.c_int32), ("with_cs", ctypes.c_int32),
        ("out_norm", _c_f32p), ("gt_norm", _c_f32p), ("mask", _c_f32p),
        ("scratch", _c_f32p), ("out", _c_f32p), ("grad_out", _c_f32p), ("grad_out_norm", _c_f32p),
    ]


class DibrRoiPool(ctypes.Structure):
    """Mirror of ``struct DibrRoiPool`` (include/dibr_b200.h)."""

    _fields_ = [
        ("num_rois", ctypes.c_int32), ("num_images", ctypes.c_int32), ("channels", ctypes.c_int32), ("height", ctypes.c_int32),
        ("width", ctypes.c_int32), ("pooled_h", ctypes.c_int32), ("pooled_w", ctypes.c_int32), ("reserved0", ctypes.c_int32),
        ("spatial_scale", ctypes.c_float), ("reserved1", ctypes.c_float),
        ("stride_n", ctypes.c_int64), ("stride_c", ctypes.c_int64), ("stride_h", ctypes.c_int64), ("stride_w", ctypes.c_int64),
        ("input", _c_f32p), ("rois", _c_f32p), ("output", _c_f32p), ("argmax", _c_i32p), ("grad_output", _c_f32p), ("grad_input", _c_f32p),
    ]


class DibrRoiAlign(ctypes.Structure):
    """Mirror of ``struct DibrRoiAlign`` (include/dibr_b200.h)."""

    _fields_ = [
        ("num_rois", ctypes.c_int32), ("num_images", ctypes.c_int32), ("channels", ctypes.c_int32), ("height", ctypes.c_int32),
        ("width", ctypes.c_int32), ("pooled_h", ctypes.c_int32), ("pooled_w", ctypes.c_int32), ("sampling_ratio", ctypes.c_int32),
        ("aligned", ctypes.c_int32), ("reserved0", ctypes.c_int32), ("spatial_scale", ctypes.c_float), ("reserved1", ctypes.c_float),
        ("stride_n", ctypes.c_int64), ("stride_c", ctypes.c_int64), ("stride_h", ctypes.c_int64), ("stride_w", ctypes.c_int64),
        ("input", _c_f32p), ("rois", _c_f32p), ("output", _c_f32p), ("grad_output", _c_f32p), ("grad_input", _c_f32p),
    ]


class DibrChamferReduce(ctypes.Structure):
    """Mirror of ``struct DibrChamferReduce`` (include/dibr_b200.h)."""

    _fields_ = [
        ("batch", ctypes.c_int32), ("stride1", ctypes.c_int32), ("stride2", ctypes.c_int32), ("threshold", ctypes.c_float),
        ("count1", _c_i32p), ("count2", _c_i32p), ("dist1", _c_f32p), ("dist2", _c_f32p),
        ("stats", _c_f32p), ("ticket", ctypes.c_void_p), ("out", _c_f32p),
        ("grad_out", _c_f32p), ("grad_dist1", _c_f32p), ("grad_dist2", _c_f32p),
    ]


EXPORTS = ["dibr_abi_version", "dibr_sizeof_pass", "dibr_last_error", "dibr_device_count", "dibr_workspace_bytes",
           "dibr_setup_faces", "dibr_setup_meshes", "dibr_forward", "dibr_backward_faces",
           "dibr_backward_meshes", "dibr_normal_map", "dibr_normal_map_pass", "dibr_launch_count_add", "dibr_render_step", "dibr_render_forward", "dibr_render_backward", "dibr_overlap_create", "dibr_overlap_destroy", "dibr_sizeof_step", "dibr_nnd_forward", "dibr_nnd_backward", "dibr_nnd_workspace_bytes", "dibr_backproject_compact",
           "dibr_backproject_compact_backward", "dibr_mask_loss_scratch_floats", "dibr_mask_loss_forward",
           "dibr_mask_loss_backward", "dibr_chamfer_reduce_forward", "dibr_chamfer_reduce_backward",
           "dibr_lab_loss_scratch_floats", "dibr_lab_loss_forward", "dibr_lab_loss_backward",
           "dibr_ms_ssim_workspace_bytes", "dibr_ms_ssim_forward", "dibr_ms_ssim_backward",
           "dibr_roi_align_forward", "dibr_roi_align_backward", "dibr_roi_pool_forward", "dibr_roi_pool_backward", "dibr_dice_loss_forward", "dibr_dice_loss_backward",
           "dibr_norm_loss_scratch_floats", "dibr_norm_loss_forward", "dibr_norm_loss_backward",
           "dibr_launch_count"]

_lib = None


def build(verbose=False):
    """Compile libdibr_b200.so for sm_100a with nvcc (cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", os.path.join(_HERE, "csrc")], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout[-4000:], r.stderr[-4000:])
    if r.returncode != 0:
        raise RuntimeError("building libdibr_b200.so failed")
    return LIB_PATH


def load():
    """Load the shared library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(self6dpp_b200 has no CPU or PyTorch fallback)")
    lib = ctypes.CDLL(LIB_PATH)
    lib.dibr_abi_version.restype = ctypes.c_int
    lib.dibr_last_error.restype = ctypes.c_char_p
    lib.dibr_device_count.restype = ctypes.c_int
    lib.dibr_launch_count.restype = ctypes.c_longlong
    lib.dibr_launch_count.argtypes = [ctypes.c_int]
    lib.dibr_launch_count_add.restype = None
    lib.dibr_launch_count_add.argtypes = [ctypes.c_longlong]
    lib.dibr_workspace_bytes.argtypes = [ctypes.POINTER(DibrPass), ctypes.POINTER(ctypes.c_size_t)]
    for name in ("dibr_setup_faces", "dibr_setup_meshes", "dibr_forward", "dibr_backward_faces", "dibr_backward_meshes"):
        fn = getattr(lib, name)
        fn.argtypes = [ctypes.POINTER(DibrPass), ctypes.c_void_p]
        fn.restype = ctypes.c_int
    lib.dibr_normal_map.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                    ctypes.c_longlong, ctypes.c_void_p]
    lib.dibr_normal_map.restype = ctypes.c_int
    lib.dibr_normal_map_pass.argtypes = [ctypes.POINTER(DibrPass), ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    lib.dibr_normal_map_pass.restype = ctypes.c_int
    for fn in (lib.dibr_render_step, lib.dibr_render_forward, lib.dibr_render_backward):
        fn.argtypes = [ctypes.POINTER(DibrStep), ctypes.c_void_p]
        fn.restype = ctypes.c_int
    lib.dibr_overlap_create.argtypes = [ctypes.POINTER(ctypes.c_void_p)]
    lib.dibr_overlap_create.restype = ctypes.c_int
    lib.dibr_overlap_destroy.argtypes = [ctypes.c_void_p]
    lib.dibr_overlap_destroy.restype = ctypes.c_int
    if lib.dibr_sizeof_step() != ctypes.sizeof(DibrStep):
        raise RuntimeError("DibrStep mirror out of date")
    for name in ("dibr_backproject_compact", "dibr_backproject_compact_backward"):
        fn = getattr(lib, name)
        fn.argtypes = [ctypes.POINTER(DibrBackproject), ctypes.c_void_p]
        fn.restype = ctypes.c_int
    for name in ("dibr_chamfer_reduce_forward", "dibr_chamfer_reduce_backward"):
        fn = getattr(lib, name)
        fn.argtypes = [ctypes.POINTER(DibrChamferReduce), ctypes.c_void_p]
        fn.restype = ctypes.c_int
    lib.dibr_mask_loss_scratch_floats.argtypes = [ctypes.c_int64]
    lib.dibr_mask_loss_scratch_floats.restype = ctypes.c_int
    for name in ("dibr_mask_loss_forward", "dibr_mask_loss_backward"):
        fn = getattr(lib, name)
        fn.argtypes = [ctypes.POINTER(DibrMaskLoss), ctypes.c_void_p]
        fn.restype = ctypes.c_int
    lib.dibr_lab_loss_scratch_floats.argtypes = [ctypes.c_int64]
    lib.dibr_lab_loss_scratch_floats.restype = ctypes.c_int
    for name in ("dibr_lab_loss_forward", "dibr_lab_loss_backward"):
        fn = getattr(lib, name)
        fn.argtypes = [ctypes.POINTER(DibrLabLoss), ctypes.c_void_p]
        fn.restype = ctypes.c_int
    lib.dibr_ms_ssim_workspace_bytes.argtypes = [ctypes.POINTER(DibrMsSsim), ctypes.POINTER(ctypes.c_size_t)]
    lib.dibr_ms_ssim_workspace_bytes.restype = ctypes.c_int
    for name in ("dibr_ms_ssim_forward", "dibr_ms_ssim_backward"):
        fn = getattr(lib, name)
        fn.argtypes = [ctypes.POINTER(DibrMsSsim), ctypes.c_void_p]
        fn.restype = ctypes.c_int
    lib.dibr_norm_loss_scratch_floats.argtypes = [ctypes.c_int64]
    lib.dibr_norm_loss_scratch_floats.restype = ctypes.c_int
    for name in ("dibr_norm_loss_forward", "dibr_norm_loss_backward"):
        fn = getattr(lib, name)
        fn.argtypes = [ctypes.POINTER(DibrNormLoss), ctypes.c_void_p]
        fn.restype = ctypes.c_int
    for name in ("dibr_dice_loss_forward", "dibr_dice_loss_backward"):
        fn = getattr(lib, name)
        fn.argtypes = [ctypes.POINTER(DibrDiceLoss), ctypes.c_void_p]
        fn.restype = ctypes.c_int
    for name in ("dibr_roi_pool_forward", "dibr_roi_pool_backward"):
        fn = getattr(lib, name)
        fn.argtypes = [ctypes.POINTER(DibrRoiPool), ctypes.c_void_p]
        fn.restype = ctypes.c_int
    for name in ("dibr_roi_align_forward", "dibr_roi_align_backward"):
        fn = getattr(lib, name)
        fn.argtypes = [ctypes.POINTER(DibrRoiAlign), ctypes.c_void_p]
        fn.restype = ctypes.c_int
    lib.dibr_nnd_workspace_bytes.argtypes = [ctypes.POINTER(DibrNnd), ctypes.POINTER(ctypes.c_size_t)]
    lib.dibr_nnd_workspace_bytes.restype = ctypes.c_int
    for name in ("dibr_nnd_forward", "dibr_nnd_backward"):
        fn = getattr(lib, name)
        fn.argtypes = [ctypes.POINTER(DibrNnd), ctypes.c_void_p]
        fn.restype = ctypes.c_int
    if lib.dibr_abi_version() != 3:
        raise RuntimeError("libdibr_b200.so ABI version mismatch")
    if lib.dibr_sizeof_pass() != ctypes.sizeof(DibrPass):
        raise RuntimeError("DibrPass mirror out of date: C sizeof %d != ctypes %d"
                           % (lib.dibr_sizeof_pass(), ctypes.sizeof(DibrPass)))
    _lib = lib
    return lib


def check(code, what):
    if code != 0:
        raise RuntimeError(f"{what} failed: {load().dibr_last_error().decode()}")


def workspace_bytes(p):
    n = ctypes.c_size_t(0)
    check(load().dibr_workspace_bytes(ctypes.byref(p), ctypes.byref(n)), "dibr_workspace_bytes")
    return n.value


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    if t is None:
        return None
    return ctypes.c_void_p(t.data_ptr())
