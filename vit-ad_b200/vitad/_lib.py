"""ctypes binding of libvitad.so (the C ABI declared in include/vitad.h).

The library is built in-tree by ``__graft_entry__.build()`` / ``make -C vit-ad_b200``; there is no
fallback: if the shared object is missing, importing this module raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# VITAD_LIB selects another build of the same library (diagnostic variants such as lib/libvitad_tl.so)
LIB_PATH = os.environ.get("VITAD_LIB") or os.path.join(os.path.dirname(_HERE), "lib", "libvitad.so")


class VitadError(RuntimeError):
    """Raised when a C-ABI call returns a negative vitad_status."""


def _load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(the CUDA extension is mandatory; there is no CPU path)"
        )
    return C.CDLL(LIB_PATH)


lib = _load()
lib.vitad_last_error.restype = C.c_char_p
lib.vitad_abi_version.restype = C.c_int
lib.vitad_launch_count.restype = C.c_uint64
lib.vitad_set_cta_pair.argtypes = [C.c_int]
lib.vitad_set_cta_pair.restype = None
lib.vitad_set_pdl.argtypes = [C.c_int]
lib.vitad_set_pdl.restype = None
lib.vitad_set_epilogue_warps.argtypes = [C.c_int]
lib.vitad_set_epilogue_warps.restype = None
lib.vitad_set_gmm_cluster4.argtypes = [C.c_int]
lib.vitad_set_gmm_cluster4.restype = None
lib.vitad_set_gmm_split.argtypes = [C.c_int]
lib.vitad_set_gmm_split.restype = None
if os.environ.get("VITAD_GMM_SPLIT") is not None:  # diagnostics: features given to the side CTA-pair launch (0 = off)
    lib.vitad_set_gmm_split(int(os.environ["VITAD_GMM_SPLIT"]))
if os.environ.get("VITAD_GMM_CLUSTER4") == "0":  # diagnostics: fused GMM kernel on CTA pairs
    lib.vitad_set_gmm_cluster4(0)
if os.environ.get("VITAD_PDL") == "0":  # diagnostics: plain stream order between kernels
    lib.vitad_set_pdl(0)

EPI_BIAS_F16 = 0
EPI_BIAS_GELU_F16 = 1
EPI_RESIDUAL_F32 = 2
EPI_QKV = 3
EPI_PATCH_EMBED = 4
EPI_F32 = 5
EPI_BIAS_RELU_F16 = 6
EPI_CONVT_RELU_F16 = 7
EPI_RES16_RELU_F16 = 8
EPI_TANH_PIX4_F32 = 9


class LinearArgs(C.Structure):
    _fields_ = [
        ("a", C.c_void_p),
        ("w", C.c_void_p),
        ("bias", C.c_void_p),
        ("m", C.c_int),
        ("n", C.c_int),
        ("k", C.c_int),
        ("lda", C.c_int),
        ("ldw", C.c_int),
        ("epilogue", C.c_int),
        ("block_n", C.c_int),
        ("out", C.c_void_p),
        ("ldo", C.c_int),
        ("resid", C.c_void_p),
        ("q", C.c_void_p),
        ("kmat", C.c_void_p),
        ("vt", C.c_void_p),
        ("tokens", C.c_int),
        ("tokens_pad", C.c_int),
        ("heads", C.c_int),
        ("q_scale", C.c_float),
        ("pos", C.c_void_p),
        ("patches", C.c_int),
        ("prefix", C.c_int),
        ("head_dim", C.c_int),
        ("windows", C.c_int),
        ("win_tokens", C.c_int),
        ("tok2win", C.c_void_p),
        ("convt_w", C.c_int),
        ("resid16", C.c_void_p),
        ("ldr", C.c_int),
        ("res_grid", C.c_int),
        ("conv_grid", C.c_int),
        ("out_pad_grid", C.c_int),
        ("split_c", C.c_int),
        ("a_taps", C.c_int),
        ("split_out", C.c_int),
        ("v_natural", C.c_int),
    ]


lib.vitad_linear_f16.argtypes = [C.POINTER(LinearArgs), C.c_void_p]
lib.vitad_linear_f16.restype = C.c_int


class LinearLnArgs(C.Structure):
    _fields_ = [("a", C.c_void_p), ("w", C.c_void_p), ("bias", C.c_void_p), ("m", C.c_int), ("k", C.c_int),
                ("lda", C.c_int), ("ldw", C.c_int), ("x", C.c_void_p), ("gamma", C.c_void_p), ("beta", C.c_void_p),
                ("eps", C.c_float), ("h", C.c_void_p), ("ldh", C.c_int)]


lib.vitad_linear_resid_ln_f16.argtypes = [C.POINTER(LinearLnArgs), C.c_void_p]
lib.vitad_linear_resid_ln_f16.restype = C.c_int
lib.vitad_set_fused_ln.argtypes = [C.c_int]
lib.vitad_set_fused_ln.restype = None
lib.vitad_layernorm768_tree.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float,
                                        C.c_void_p]
lib.vitad_layernorm768_tree.restype = C.c_int
lib.vitad_set_v_natural.argtypes = [C.c_int]
lib.vitad_set_v_natural.restype = None
if os.environ.get("VITAD_VNAT") == "0":  # diagnostics: transposed, zero-padded V for the attention kernel
    lib.vitad_set_v_natural(0)
if os.environ.get("VITAD_FUSED_LN") == "0":  # diagnostics: separate residual GEMM + LayerNorm launches
    lib.vitad_set_fused_ln(0)


def check(rc: int) -> None:
    if rc != 0:
        raise VitadError(f"vitad status {rc}: {lib.vitad_last_error().decode()}")


def launch_count() -> int:
    return int(lib.vitad_launch_count())


# ---------------------------------------------------------------------------------- encoder kernels
_vp, _i, _f, _sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t
lib.vitad_layernorm.argtypes = [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _f, _i, _vp]
lib.vitad_layernorm.restype = _i
lib.vitad_patchify.argtypes = [_vp, _vp, _i, _i, _i, _i, _vp]
lib.vitad_patchify.restype = _i
lib.vitad_prefix_tokens.argtypes = [_vp, _vp, _vp, _i, _i, _i, _i, _vp]
lib.vitad_prefix_tokens.restype = _i


class AttentionArgs(C.Structure):
    _fields_ = [("q", _vp), ("k", _vp), ("vt", _vp), ("out", _vp), ("batch_windows", _i), ("heads", _i),
                ("tokens", _i), ("tokens_pad", _i), ("head_dim", _i), ("windows", _i), ("bias", _vp),
                ("region", _vp), ("win2tok", _vp), ("v", _vp)]


lib.vitad_attention_f16.argtypes = [C.POINTER(AttentionArgs), _vp]
lib.vitad_attention_f16.restype = _i

DEIT_MAX_DEPTH = 24


class DeitLayer(C.Structure):
    _fields_ = [(n, _vp) for n in ("ln1_w", "ln1_b", "qkv_w", "qkv_b", "proj_w", "proj_b", "ln2_w", "ln2_b", "fc1_w",
                                   "fc1_b", "fc2_w", "fc2_b")]


class DeitWeights(C.Structure):
    _fields_ = [(n, _i) for n in ("img", "patch", "dim", "heads", "hidden", "depth", "tokens", "prefix")] + [
        ("patch_w", _vp), ("patch_b", _vp), ("prefix_tokens", _vp), ("pos", _vp), ("norm_w", _vp), ("norm_b", _vp),
        ("layers", C.POINTER(DeitLayer)),
    ]


lib.vitad_deit_workspace_bytes.argtypes = [C.POINTER(DeitWeights), _i]
lib.vitad_deit_workspace_bytes.restype = _sz
lib.vitad_deit_forward.argtypes = [C.POINTER(DeitWeights), _vp, _i, _i, _vp, _sz, _vp, _vp, _vp, _i, _vp]
lib.vitad_deit_forward.restype = _i
lib.vitad_deit_forward_u8.argtypes = lib.vitad_deit_forward.argtypes
lib.vitad_deit_forward_u8.restype = _i
lib.vitad_patchify_u8.argtypes = [_vp, _vp, _i, _i, _i, _i, _vp]
lib.vitad_patchify_u8.restype = _i

# ------------------------------------------------------------------------------------ Swin / EsViT
class SwinBlock(C.Structure):
    _fields_ = [(n, _vp) for n in ("ln1_w", "ln1_b", "qkv_w", "qkv_b", "attn_bias", "proj_w", "proj_b", "ln2_w",
                                   "ln2_b", "fc1_w", "fc1_b", "fc2_w", "fc2_b")] + [("shift", _i)]


class SwinStage(C.Structure):
    _fields_ = [(n, _i) for n in ("dim", "heads", "res", "window", "depth")] + [
        ("blocks", C.POINTER(SwinBlock)), ("tok2win", _vp * 2), ("win2tok", _vp * 2), ("region", _vp),
        ("merge_ln_w", _vp), ("merge_ln_b", _vp), ("merge_w", _vp)]


class SwinWeights(C.Structure):
    _fields_ = [(n, _i) for n in ("img", "patch", "embed", "stages")] + [
        ("patch_w", _vp), ("patch_b", _vp), ("patch_ln_w", _vp), ("patch_ln_b", _vp), ("norm_w", _vp),
        ("norm_b", _vp), ("stage", C.POINTER(SwinStage))]


lib.vitad_swin_workspace_bytes.argtypes = [C.POINTER(SwinWeights), _i]
lib.vitad_swin_workspace_bytes.restype = _sz
lib.vitad_swin_forward.argtypes = [C.POINTER(SwinWeights), _vp, _i, _vp, _sz, _vp, _vp, _vp, _i, _vp]
lib.vitad_swin_forward.restype = _i

# ---------------------------------------------------------------------------------------- GMM head
lib.vitad_gmm_plan.argtypes = [_i, C.POINTER(_i), C.POINTER(_i), C.POINTER(_i)]
lib.vitad_gmm_plan.restype = _i
lib.vitad_gmm_packed_weight_bytes.argtypes = [_i, _i]
lib.vitad_gmm_packed_weight_bytes.restype = _sz
lib.vitad_gmm_pack_weights.argtypes = [_vp, _vp, _vp, _vp, _i, _i, _vp, _vp]
lib.vitad_gmm_pack_weights.restype = _i
lib.vitad_gmm_make_operand.argtypes = [_vp, _i, _vp, _i, _i, _vp]
lib.vitad_gmm_make_operand.restype = _i
lib.vitad_gmm_log_pi.argtypes = [_vp, _i, _vp, _vp, _vp, _vp, _i, _i, _i, _vp]
lib.vitad_gmm_log_pi.restype = _i
lib.vitad_gmm_log_pi_seeded.argtypes = [_vp, _i, _vp, _vp, C.c_uint64, C.c_uint32, _vp, _i, _i, _i, _vp]
lib.vitad_gmm_log_pi_seeded.restype = _i
lib.vitad_gumbel_noise.argtypes = [C.c_uint64, C.c_uint32, _vp, _i, _i, _vp]
lib.vitad_gumbel_noise.restype = _i
lib.vitad_gmm_pi_packed_bytes.argtypes = [_i, _i]
lib.vitad_gmm_pi_packed_bytes.restype = _sz
lib.vitad_gmm_pack_pi.argtypes = [_vp, _i, _i, _vp, _vp]
lib.vitad_gmm_pack_pi.restype = _i
lib.vitad_gmm_log_pi_workspace_bytes.argtypes = [_i, _i, _i]
lib.vitad_gmm_log_pi_workspace_bytes.restype = _sz
lib.vitad_gmm_log_pi_tc.argtypes = [_vp, _i, _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _sz, _vp]
lib.vitad_gmm_log_pi_tc.restype = _i
lib.vitad_gmm_patch_loglik.argtypes = [_vp, _vp, _vp, _vp, _i, _vp, _i, _vp, _i, _i, _i, _vp]
lib.vitad_gmm_patch_loglik.restype = _i
lib.vitad_gmm_finish.argtypes = [_vp, _vp, _vp, _i, _i, _vp]
lib.vitad_gmm_finish.restype = _i

# ------------------------------------------------------------------------------------- score maps
lib.vitad_bilinear_up.argtypes = [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]
lib.vitad_bilinear_up.restype = _i
lib.vitad_l2_map_score.argtypes = [_vp, _vp, _vp, _vp, _i, _i, _i, _vp]
lib.vitad_l2_map_score.restype = _i

# ---------------------------------------------------------------------------- normalizing flow
class NfStep(C.Structure):
    _fields_ = [(n, _vp) for n in ("w0p", "b0p", "w2p", "b2p", "scale", "offset", "inv_perm")] + [("ksize", _i)]


class NfWeights(C.Structure):
    _fields_ = [("channels", _i), ("grid", _i), ("hidden_pad", _i), ("steps", _i), ("clamp", _f),
                ("logdet_const", _f), ("step", C.POINTER(NfStep))]


lib.vitad_nf_workspace_bytes.argtypes = [C.POINTER(NfWeights), _i]
lib.vitad_nf_workspace_bytes.restype = _sz
lib.vitad_nf_forward.argtypes = [C.POINTER(NfWeights), _vp, _i, _vp, _sz, _vp, _vp, _vp]
lib.vitad_nf_forward.restype = _i

# ------------------------------------------------------------------------- small CNN decoder
class CnnDecoderWeights(C.Structure):
    _fields_ = [("latent", _i), ("hidden", _i), ("grid0", _i), ("last_cin", _i), ("chan", _i * 5),
                ("lin1_w", _vp), ("lin1_b", _vp), ("lin2_w", _vp), ("lin2_b", _vp), ("conv_w", _vp * 4),
                ("conv_b", _vp * 4), ("last_w", _vp), ("last_b", _vp)]


lib.vitad_cnn_decoder_workspace_bytes.argtypes = [C.POINTER(CnnDecoderWeights), _i]
lib.vitad_cnn_decoder_workspace_bytes.restype = _sz
lib.vitad_cnn_decoder_forward.argtypes = [C.POINTER(CnnDecoderWeights), _vp, _i, _vp, _sz, _vp, _vp]
lib.vitad_cnn_decoder_forward.restype = _i

# ------------------------------------------------------------------- reverse-ResNet decoder
RESNET_MAX_BLOCKS = 24


class ResnetBlock(C.Structure):
    _fields_ = [("cin", _i), ("width", _i), ("cout", _i), ("stride", _i), ("w3", _vp), ("b3", _vp), ("w2", _vp),
                ("b2", _vp), ("w1", _vp), ("b1", _vp), ("wup", _vp), ("bup", _vp)]


class ResnetDecoderWeights(C.Structure):
    _fields_ = [("latent", _i), ("hidden", _i), ("feat", _i), ("grid0", _i), ("n_blocks", _i), ("last_c", _i),
                ("fc1_w", _vp), ("fc1_b", _vp), ("fc2_w", _vp), ("fc2_b", _vp), ("blocks", ResnetBlock * RESNET_MAX_BLOCKS),
                ("last_w", _vp), ("last_b", _vp), ("split", _i)]


lib.vitad_resnet_decoder_workspace_bytes.argtypes = [C.POINTER(ResnetDecoderWeights), _i]
lib.vitad_resnet_decoder_workspace_bytes.restype = _sz
lib.vitad_resnet_decoder_forward.argtypes = [C.POINTER(ResnetDecoderWeights), _vp, _i, _vp, _sz, _vp, _vp]
lib.vitad_resnet_decoder_forward.restype = _i

# ------------------------------------------------------------------------------ input resize
lib.vitad_resize_ksize.argtypes = [_i, _i]
lib.vitad_resize_ksize.restype = _i
lib.vitad_resize_plan.argtypes = [_i, _i, _vp]
lib.vitad_resize_plan.restype = _i
lib.vitad_resize_bilinear_u8.argtypes = [_vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]
lib.vitad_resize_bilinear_u8.restype = _i

MDN_KA = 784  # K extent of the packed MDN operands (768 + 16)


def gmm_plan(num_gaussians: int):
    n_kc, kc, kcv = _i(), _i(), _i()
    check(lib.vitad_gmm_plan(num_gaussians, C.byref(n_kc), C.byref(kc), C.byref(kcv)))
    return n_kc.value, kc.value, kcv.value
