"""ctypes binding of libspnerf_sm100a.so (the C ABI declared in include/spnerf_b200.h)."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SPNERF_LIB") or os.path.join(_HERE, "lib", "libspnerf_sm100a.so")   # SPNERF_LIB: experiment builds (tools/build_variant.sh)

SELFTEST_MAX_KSTEPS = 64


class SpnerfError(RuntimeError):
    pass


class UmmaSelftest(ctypes.Structure):
    _fields_ = [
        ("a_img", ctypes.c_void_p),
        ("b_img", ctypes.c_void_p),
        ("d_out", ctypes.c_void_p),
        ("a_bytes", ctypes.c_uint32),
        ("b_bytes", ctypes.c_uint32),
        ("n", ctypes.c_uint32),
        ("ksteps", ctypes.c_uint32),
        ("idesc", ctypes.c_uint32),
        ("_pad", ctypes.c_uint32),
        ("a_desc_template", ctypes.c_uint64),
        ("b_desc_template", ctypes.c_uint64),
        ("a_off", ctypes.c_uint32 * SELFTEST_MAX_KSTEPS),
        ("b_off", ctypes.c_uint32 * SELFTEST_MAX_KSTEPS),
    ]


_lib = None


def lib():
    """Load the shared library once.  Raises (never falls back) if it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SpnerfError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU or PyTorch fallback)")
        _lib = ctypes.CDLL(LIB_PATH)
        _declare(_lib)
    return _lib


def _declare(L):
    L.spnerf_abi_version.restype = ctypes.c_int
    L.spnerf_abi_version.argtypes = []
    L.spnerf_watchdog_code.restype = ctypes.c_uint
    L.spnerf_watchdog_code.argtypes = []
    L.spnerf_selftest_umma.restype = ctypes.c_int
    L.spnerf_selftest_umma.argtypes = [ctypes.POINTER(UmmaSelftest), ctypes.c_void_p]


def check(rc, what):
    if rc == 0:
        return
    if rc > 0:
        names = {1: "bad argument", 2: "unsupported configuration", 3: "workspace too small"}
        raise SpnerfError(f"{what}: {names.get(rc, rc)}")
    raise SpnerfError(f"{what}: CUDA error {-rc}")


# ------------------------------------------------------------------------------------------------
# point network
# ------------------------------------------------------------------------------------------------
NUM_PARAMS = 45

# reference state_dict key -> parameter slot (include/spnerf_b200.h SPNERF_P_*)
PARAM_SLOTS = {"semantic_embedding.weight": 0}
for _i in range(8):
    PARAM_SLOTS[f"fc_net.{2 * _i}.weight"] = 1 + 2 * _i
    PARAM_SLOTS[f"fc_net.{2 * _i}.bias"] = 2 + 2 * _i
PARAM_SLOTS.update({
    "sigma_from_xyz.0.weight": 17, "sigma_from_xyz.0.bias": 18,
    "feats_from_xyz.weight": 19, "feats_from_xyz.bias": 20,
    "logit_from_label.0.weight": 21, "logit_from_label.0.bias": 22,
    "logit_from_label.2.weight": 23, "logit_from_label.2.bias": 24,
    "rgb_from_xyzdir.0.weight": 25, "rgb_from_xyzdir.0.bias": 26,
    "rgb_from_xyzdir.2.weight": 27, "rgb_from_xyzdir.2.bias": 28,
    "sky_color.0.weight": 37, "sky_color.0.bias": 38, "sky_color.2.weight": 39, "sky_color.2.bias": 40,
    "beta_from_xyz.0.weight": 41, "beta_from_xyz.0.bias": 42,
    "beta_from_xyz.2.weight": 43, "beta_from_xyz.2.bias": 44,
})
for _j in range(4):
    PARAM_SLOTS[f"sun_v_net.{2 * _j}.weight"] = 29 + 2 * _j
    PARAM_SLOTS[f"sun_v_net.{2 * _j}.bias"] = 30 + 2 * _j


class NetConfig(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in
                ("feat", "layers", "skip_layer", "mapping", "sem", "num_sem_classes", "emb_dim", "beta", "t_dim", "relu")]


class NetSizes(ctypes.Structure):
    _fields_ = [("fwd_blob_bytes", ctypes.c_int64), ("bwd_blob_bytes", ctypes.c_int64),
                ("small_floats", ctypes.c_int64), ("steps_bytes", ctypes.c_int64),
                ("fwd_steps", ctypes.c_int32), ("bwd_steps", ctypes.c_int32),
                ("save_slabs_per_tile", ctypes.c_int32), ("grad_slabs_per_tile", ctypes.c_int32),
                ("n_out", ctypes.c_int32),
                ("in_dim", ctypes.c_int32), ("tile_points", ctypes.c_int32)]


class MlpFwd(ctypes.Structure):
    _fields_ = [("cfg", NetConfig),
                ("rays", ctypes.c_void_p), ("z", ctypes.c_void_p), ("xyz", ctypes.c_void_p),
                ("dir_override", ctypes.c_void_p), ("labels", ctypes.c_void_p), ("t_emb", ctypes.c_void_p),
                ("sky", ctypes.c_void_p),
                ("n_rays", ctypes.c_int64), ("n_samples", ctypes.c_int32), ("n_steps", ctypes.c_int32),
                ("blob", ctypes.c_void_p), ("steps", ctypes.c_void_p), ("small", ctypes.c_void_p),
                ("out", ctypes.c_void_p), ("saves", ctypes.c_void_p),
                ("debug_flags", ctypes.c_int32), ("_pad", ctypes.c_int32)]


def _declare_net(L):
    L.spnerf_net_sizes.restype = ctypes.c_int
    L.spnerf_net_sizes.argtypes = [ctypes.POINTER(NetConfig), ctypes.POINTER(NetSizes)]
    L.spnerf_net_pack_workspace_bytes.restype = ctypes.c_int64
    L.spnerf_net_pack_workspace_bytes.argtypes = [ctypes.POINTER(NetConfig)]
    L.spnerf_net_prepare.restype = ctypes.c_int
    L.spnerf_net_prepare.argtypes = [ctypes.POINTER(NetConfig), ctypes.POINTER(ctypes.c_void_p), ctypes.c_void_p,
                                     ctypes.c_int64] + [ctypes.c_void_p] * 6
    L.spnerf_net_pack.restype = ctypes.c_int
    L.spnerf_net_pack.argtypes = [ctypes.POINTER(NetConfig)] + [ctypes.c_void_p] * 5
    L.spnerf_sky_fwd.restype = ctypes.c_int
    L.spnerf_sky_fwd.argtypes = [ctypes.c_void_p, ctypes.POINTER(NetConfig), ctypes.c_void_p, ctypes.c_int64,
                                 ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    L.spnerf_mlp_fwd.restype = ctypes.c_int
    L.spnerf_mlp_fwd.argtypes = [ctypes.POINTER(MlpFwd), ctypes.c_void_p]


VP, I32, I64, F32 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_float


class CompositeFwd(ctypes.Structure):
    _fields_ = [("out", VP), ("z", VP), ("noise", VP), ("n_rays", I64), ("n_samples", I32), ("n_out", I32),
                ("col_sem", I32), ("n_sem", I32), ("noise_std", F32), ("_pad", I32),
                ("weights", VP), ("transparency", VP), ("rgb", VP), ("rgb_raw", VP), ("depth", VP),
                ("sem_logits", VP), ("ray_aux", VP), ("sem_argmax", VP), ("col_beta", I32), ("_pad2", I32)]


class CompositeBwd(ctypes.Structure):
    _fields_ = [("out", VP), ("z", VP), ("noise", VP), ("weights", VP), ("transparency", VP), ("rgb_raw", VP),
                ("g_rgb", VP), ("g_depth", VP), ("g_sem_logits", VP), ("g_weights", VP), ("g_transparency", VP),
                ("g_out_ext", VP), ("n_rays", I64), ("n_samples", I32), ("n_out", I32), ("col_sem", I32),
                ("n_sem", I32), ("noise_std", F32), ("_pad", I32), ("g_out", VP), ("g_sky_ray", VP),
                ("g_absmax", VP)]


class Losses(ctypes.Structure):
    _fields_ = [("n_rays", I64), ("n_samples", I32), ("n_sem", I32),
                ("rgb", VP), ("rgb_target", VP), ("g_rgb", VP),
                ("depth", VP), ("z", VP), ("weights", VP), ("target_depth", VP), ("target_weight", VP),
                ("target_std", VP), ("valid_depth", VP), ("lambda_ds", F32), ("use_all_depth", I32),
                ("g_depth", VP),
                ("sem_logits", VP), ("labels", VP), ("lambda_ss", F32), ("_pad", I32), ("g_sem_logits", VP),
                ("losses", VP), ("workspace", VP), ("gnll", I32), ("_pad2", I32), ("g_weights", VP),
                ("target_stride", I64)]


class LossSolar(ctypes.Structure):
    _fields_ = [("n_rays", I64), ("n_samples", I32), ("_pad", I32), ("transparency_sc", VP), ("weights_sc", VP),
                ("sun_sc", VP), ("sun_stride", I64), ("lambda_sc", F32), ("_pad2", I32), ("upstream", VP),
                ("g_sun", VP), ("losses", VP), ("workspace", VP)]


class LossUncertainty(ctypes.Structure):
    _fields_ = [("n_rays", I64), ("n_samples", I32), ("_pad", I32), ("rgb", VP), ("rgb_target", VP), ("weights", VP),
                ("beta", VP), ("beta_stride", I64), ("beta_min", F32), ("_pad2", I32), ("upstream", VP),
                ("beta_ray", VP), ("g_rgb", VP), ("g_weights", VP), ("g_beta", VP), ("losses", VP), ("workspace", VP)]


F64 = ctypes.c_double


class RaysFromGeodetic(ctypes.Structure):
    _fields_ = [("lon_near", VP), ("lat_near", VP), ("lon_far", VP), ("lat_far", VP), ("alt_near", F64), ("alt_far", F64),
                ("center", F32 * 3), ("range", F32), ("normalize", I32), ("has_sun", I32), ("sun_dir", F32 * 3),
                ("row_stride", I32), ("n_rays", I64), ("rays", VP)]


class PointsToGeodetic(ctypes.Structure):
    _fields_ = [("rays", VP), ("row_stride", I32), ("_pad", I32), ("depth", VP), ("center", F32 * 3), ("range", F32),
                ("n_rays", I64), ("lat", VP), ("lon", VP), ("alt", VP)]


class Guided(ctypes.Structure):
    _fields_ = [("rays", VP), ("z", VP), ("weights", VP), ("depth", VP), ("valid_depth", VP),
                ("target_depth", VP), ("target_depth_stride", I64), ("target_std", VP), ("u_pred", VP),
                ("u_gt", VP), ("t_table", VP), ("gauss_table", VP), ("n_rays", I64), ("n_samples", I32),
                ("_pad", I32), ("z_unsort", VP), ("z_sorted", VP), ("searchsorted_out", VP)]


class MlpBwd(ctypes.Structure):
    _fields_ = [("cfg", NetConfig), ("g_out", VP), ("out", VP), ("rays", VP), ("labels", VP), ("t_emb", VP),
                ("n_rays", I64), ("n_samples", I32), ("n_steps", I32), ("blob", VP), ("steps", VP), ("small", VP),
                ("saves", VP), ("grad_saves", VP), ("g_absmax", VP), ("scale_out", VP), ("g_emb", VP),
                ("g_small_bias", VP), ("g_t_emb", VP), ("debug_flags", I32), ("_pad", I32)]


class MlpWgrad(ctypes.Structure):
    _fields_ = [("cfg", NetConfig), ("n_points", I64), ("saves", VP), ("grad_saves", VP), ("scale", VP),
                ("grads_host", ctypes.POINTER(VP)), ("workspace", VP), ("workspace_bytes", I64),
                ("accum", VP), ("absmax_reset", VP)]


# offsets (floats) into the self-cleaning accumulator block (include/spnerf_b200.h SPNERF_ACC_*)
ACC_SMALL_BIAS, ACC_EMB, ACC_SKY_W0, ACC_SKY_B0, ACC_SKY_W2, ACC_SKY_B2, ACCUM_FLOATS = 0, 16, 96, 864, 1120, 1888, 1892


def _declare_rest(L):
    for name, st in (("spnerf_composite_fwd", CompositeFwd), ("spnerf_composite_bwd", CompositeBwd),
                     ("spnerf_losses", Losses), ("spnerf_sample_guided", Guided),
                     ("spnerf_mlp_bwd_data", MlpBwd), ("spnerf_mlp_wgrad_prepare", MlpWgrad),
                     ("spnerf_mlp_bwd_weights", MlpWgrad)):
        f = getattr(L, name)
        f.restype = ctypes.c_int
        f.argtypes = [ctypes.POINTER(st), VP]
    for name, st in (("spnerf_rays_from_geodetic", RaysFromGeodetic), ("spnerf_points_to_geodetic", PointsToGeodetic)):
        f = getattr(L, name)
        f.restype = ctypes.c_int
        f.argtypes = [ctypes.POINTER(st), VP]
    for name, st in (("spnerf_loss_solar", LossSolar), ("spnerf_loss_uncertainty", LossUncertainty)):
        f = getattr(L, name)
        f.restype = ctypes.c_int
        f.argtypes = [ctypes.POINTER(st), I32, VP]
    L.spnerf_losses_workspace_bytes.restype = I64
    L.spnerf_losses_workspace_bytes.argtypes = []
    L.spnerf_mlp_wgrad_workspace_bytes.restype = I64
    L.spnerf_mlp_wgrad_workspace_bytes.argtypes = [ctypes.POINTER(NetConfig)]
    L.spnerf_sample_coarse.restype = ctypes.c_int
    L.spnerf_sample_coarse.argtypes = [VP, VP, VP, I64, I32, VP, VP]
    L.spnerf_sample_coarse_rng.restype = ctypes.c_int
    L.spnerf_sample_coarse_rng.argtypes = [VP, VP, VP, I64, I32, VP, VP]
    L.spnerf_sky_bwd.restype = ctypes.c_int
    L.spnerf_sky_bwd.argtypes = [VP, ctypes.POINTER(NetConfig), VP, VP, VP, VP, I64, VP, VP, VP, VP, VP]
    L.spnerf_adam_step.restype = I32
    L.spnerf_adam_step.argtypes = [VP, VP, VP, VP, I64, I64] + [ctypes.c_double] * 4 + [VP]
    L.spnerf_struct_sizes.restype = None
    L.spnerf_struct_sizes.argtypes = [ctypes.POINTER(I32)]


STRUCTS = (UmmaSelftest, NetConfig, NetSizes, MlpFwd, CompositeFwd, CompositeBwd, Losses, Guided, MlpBwd, MlpWgrad,
           LossSolar, LossUncertainty, RaysFromGeodetic, PointsToGeodetic)

_declare_base = _declare


def _declare(L):  # noqa: F811
    _declare_base(L)
    _declare_net(L)
    _declare_rest(L)
    sizes = (I32 * len(STRUCTS))()
    L.spnerf_struct_sizes(sizes)
    for st, n in zip(STRUCTS, sizes):
        if ctypes.sizeof(st) != n:
            raise SpnerfError(f"ABI mismatch: {st.__name__} is {ctypes.sizeof(st)} bytes here, {n} in the library")
