"""ctypes binding of libdmc_b200.so (C ABI declared in include/dmc.h).

There is NO fallback: if the shared library is missing, cannot be loaded, or no sm_100 device is usable,
every product entry point raises ``DmcError``.  (The oracle under ``oracle/`` is test infrastructure and
is never imported from here.)
"""

from __future__ import annotations

import ctypes as C
import os
import threading

PKG = os.path.dirname(os.path.abspath(__file__))
# DMC_LIB: an alternative build of the same ABI (A/B measurements of kernel variants only; symbols it lacks are skipped)
LIB_PATH = os.environ.get("DMC_LIB") or os.path.join(PKG, "libdmc_b200.so")


class DmcError(RuntimeError):
    pass


c_f32p = C.POINTER(C.c_float)
vp = C.c_void_p


class DdimCoef(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("sqrt_one_minus_a", "sqrt_a", "sqrt_a_next", "dir_coef", "sigma")]


class DdpmCoef(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("sqrt_recip_a", "sqrt_recipm1_a", "coef1", "coef2", "noise_scale")]


class Guidance(C.Structure):
    _fields_ = [("cfg_scale", C.c_float), ("clip_mode", C.c_int32), ("q_lo", C.c_int32), ("q_hi", C.c_int32),
                ("q_weight", C.c_float)]


class CondDesc(C.Structure):
    _fields_ = [("t", vp), ("y", vp), ("B", C.c_int32), ("uniform_t", C.c_int32), ("num_classes", C.c_int32),
                ("half", C.c_int32), ("temb", C.c_int32), ("ncols", C.c_int32), ("freqs", vp), ("w1", vp), ("b1", vp),
                ("w2", vp), ("b2", vp), ("wt_all", vp), ("bt_all", vp), ("ytab", vp), ("scratch", vp), ("cond", vp)]


class StemDesc(C.Structure):
    _fields_ = [("x", vp), ("x_batch", C.c_int32), ("B", C.c_int32), ("Cin", C.c_int32), ("H", C.c_int32),
                ("W", C.c_int32), ("Cout", C.c_int32), ("weight", vp), ("bias", vp), ("out", vp), ("out_lo", vp)]


class GnStatsDesc(C.Structure):
    _fields_ = [("src", vp), ("B", C.c_int32), ("HW", C.c_int32), ("C", C.c_int32), ("stats", vp), ("src_lo", vp)]


class GnApplyDesc(C.Structure):
    _fields_ = [("nsrc", C.c_int32), ("src", vp * 2), ("src_c", C.c_int32 * 2), ("stats", vp * 2), ("stats_slots", C.c_int32 * 2),
                ("B", C.c_int32), ("HW", C.c_int32), ("groups", C.c_int32), ("gamma", vp), ("beta", vp), ("eps", C.c_float),
                ("silu", C.c_int32), ("out", vp), ("src_lo", vp * 2), ("out_lo", vp), ("drop_p", C.c_float),
                ("seed", C.c_uint32), ("seed_dev", vp)]


class ConvDesc(C.Structure):
    _fields_ = [("nsrc", C.c_int32), ("src", vp * 3), ("src_c", C.c_int32 * 3), ("src_taps", C.c_int32 * 3),
                ("B", C.c_int32), ("Hin", C.c_int32), ("Win", C.c_int32), ("stride", C.c_int32),
                ("up_phase", C.c_int32), ("weight", vp), ("Cout", C.c_int32), ("Cout_pad", C.c_int32),
                ("Ktot", C.c_int32), ("bias", vp), ("cond", vp), ("cond_stride", C.c_int32), ("residual", vp),
                ("out_bf16", vp), ("out_f32_nchw", vp), ("stats", vp), ("stats_slots", C.c_int32), ("impl", C.c_int32),
                ("act", C.c_int32), ("gate", vp), ("gate_stride", C.c_int32), ("residual_f32", vp), ("out_f32_nhwc", vp),
                ("out_lo", vp), ("residual_lo", vp), ("unpatch_p", C.c_int32),
                ("gn_nver", C.c_int32), ("gn_out", vp * 2), ("gn_pitch", C.c_int32 * 2), ("gn_coff", C.c_int32 * 2),
                ("gn_gamma", vp * 2), ("gn_beta", vp * 2), ("gn_gsize", C.c_int32 * 2), ("gn_silu", C.c_int32 * 2),
                ("gn_eps", C.c_float), ("gn_counters", vp), ("a_affine", vp)]


class DitCondDesc(C.Structure):
    _fields_ = [("t", vp), ("y", vp), ("B", C.c_int32), ("uniform_t", C.c_int32), ("num_classes", C.c_int32),
                ("freq_dim", C.c_int32), ("hidden", C.c_int32), ("ncols", C.c_int32), ("freqs", vp), ("w1", vp), ("b1", vp),
                ("w2", vp), ("b2", vp), ("emb", vp), ("w_all", vp), ("b_all", vp), ("scratch", vp), ("mod", vp)]


class PatchEmbedDesc(C.Structure):
    _fields_ = [("x", vp), ("x_batch", C.c_int32), ("B", C.c_int32), ("Cin", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
                ("patch", C.c_int32), ("hidden", C.c_int32), ("weight", vp), ("bias", vp), ("pos", vp), ("out", vp)]


class GnBwdDesc(C.Structure):
    _fields_ = [("nsrc", C.c_int32), ("src", vp * 2), ("src_c", C.c_int32 * 2), ("stats", vp * 2),
                ("stats_slots", C.c_int32 * 2), ("dout", vp), ("dsrc", vp * 2), ("accumulate", C.c_int32 * 2),
                ("B", C.c_int32), ("HW", C.c_int32), ("groups", C.c_int32), ("gamma", vp), ("beta", vp), ("eps", C.c_float),
                ("silu", C.c_int32), ("drop_p", C.c_float), ("seed", C.c_uint32), ("seed_dev", vp), ("dgamma", vp),
                ("dbeta", vp), ("scratch", vp)]


class AttnBwdDesc(C.Structure):
    _fields_ = [("qkv", vp), ("out", vp), ("dout", vp), ("dqkv", vp), ("B", C.c_int32), ("L", C.c_int32),
                ("heads", C.c_int32), ("C", C.c_int32)]


class DitGlmDesc(C.Structure):
    _fields_ = [("x_in", vp), ("y", vp), ("gate", vp), ("x_out", vp), ("h", vp), ("shift", vp), ("scale", vp),
                ("mod_stride", C.c_int32), ("gate_stride", C.c_int32), ("B", C.c_int32), ("L", C.c_int32), ("C", C.c_int32),
                ("eps", C.c_float), ("drop_p", C.c_float), ("seed", C.c_uint32)]


class DitGlmBwdDesc(C.Structure):
    _fields_ = [("x", vp), ("dh", vp), ("dx_out", vp), ("y", vp), ("gate", vp), ("scale", vp), ("mod_stride", C.c_int32),
                ("gate_stride", C.c_int32), ("dx_in", vp), ("dy", vp), ("dgate", vp), ("dshift", vp), ("dscale", vp), ("B", C.c_int32), ("L", C.c_int32),
                ("C", C.c_int32), ("eps", C.c_float), ("drop_p", C.c_float), ("seed", C.c_uint32), ("scratch", vp)]


DIT_GLM_BWD_SLICES = 4  # DMC_DIT_GLM_BWD_SLICES of include/dmc.h


class WgradDesc(C.Structure):
    _fields_ = [("x", vp), ("dy", vp), ("B", C.c_int32), ("Hin", C.c_int32), ("Win", C.c_int32), ("Cin", C.c_int32),
                ("Cout", C.c_int32), ("stride", C.c_int32), ("taps", C.c_int32), ("splits", C.c_int32), ("partial", vp),
                ("dw", vp), ("accumulate", C.c_int32), ("dw_cout", C.c_int32), ("dw_cin", C.c_int32), ("dw_cin_total", C.c_int32),
                ("dw_ci0", C.c_int32)]


class PackItem(C.Structure):
    _fields_ = [("src", vp), ("dst", vp)] + [(n, C.c_int32) for n in ("cout", "cin_total", "ci0", "cin", "taps", "mode", "ld", "col0",
                                                                     "cpad", "pad_")]


def pack_table(items, device):
    """device copy of an array of PackItem (dmc_pack_weights reads its items from device memory)"""
    import torch

    arr = (PackItem * len(items))(*items)
    host = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
    return host.to(device), len(items)


class OptItem(C.Structure):
    _fields_ = [("p", vp), ("g", vp), ("m", vp), ("v", vp), ("ema", vp), ("n", C.c_int64)]


class OptChunk(C.Structure):
    _fields_ = [("item", C.c_int32), ("pad_", C.c_int32), ("start", C.c_int64)]


class AdamWDesc(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("lr", "beta1", "beta2", "eps", "weight_decay", "bias_correction1", "bias_correction2",
                                         "max_norm", "ema_decay")]


class HeadDesc(C.Structure):
    _fields_ = [("src", vp), ("stats", vp), ("stats_slots", C.c_int32), ("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
                ("C", C.c_int32), ("Cout", C.c_int32), ("groups", C.c_int32), ("gamma", vp), ("beta", vp), ("eps", C.c_float),
                ("weight", vp), ("bias", vp), ("out", vp), ("wfrag", vp)]


class LnModDesc(C.Structure):
    _fields_ = [("x", vp), ("out", vp), ("B", C.c_int32), ("L", C.c_int32), ("C", C.c_int32), ("shift", vp), ("scale", vp),
                ("mod_stride", C.c_int32), ("eps", C.c_float), ("out_lo", vp)]


class AttnDesc(C.Structure):
    _fields_ = [("qkv", vp), ("out", vp), ("B", C.c_int32), ("L", C.c_int32), ("heads", C.c_int32), ("C", C.c_int32),
                ("impl", C.c_int32), ("qkv_lo", vp), ("out_lo", vp)]


class UpsampleDesc(C.Structure):
    _fields_ = [("src", vp), ("out", vp), ("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("C", C.c_int32)]


class GnCoeffDesc(C.Structure):
    _fields_ = [("stats", vp), ("stats_slots", C.c_int32), ("B", C.c_int32), ("HW", C.c_int32), ("C", C.c_int32),
                ("groups", C.c_int32), ("gamma", vp), ("beta", vp), ("eps", C.c_float), ("out", vp)]


class StemColsDesc(C.Structure):
    _fields_ = [("x", vp), ("x_batch", C.c_int32), ("B", C.c_int32), ("Cin", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
                ("out", vp)]


class HeadTapsDesc(C.Structure):
    _fields_ = [("y", vp), ("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("Cout", C.c_int32), ("ypitch", C.c_int32),
                ("bias", vp), ("out", vp)]


class StepDesc(C.Structure):
    _fields_ = [("x", vp), ("eps_c", vp), ("eps_u", vp), ("noise", vp), ("x_out", vp), ("B", C.c_int32),
                ("n_per_sample", C.c_int32), ("coef_dev", vp), ("g", Guidance), ("step_index_dev", vp)]


# every symbol include/dmc.h declares: (restype, argtypes)
SYMBOLS = {
    "dmc_last_error": (C.c_char_p, []),
    "dmc_abi_version": (C.c_int, []),
    "dmc_init": (C.c_int, []),
    "dmc_ddim_step": (C.c_int, [vp, vp, vp, vp, vp, C.c_int32, C.c_int32, vp, C.POINTER(Guidance), vp]),
    "dmc_ddpm_step": (C.c_int, [vp, vp, vp, vp, vp, C.c_int32, C.c_int32, vp, C.POINTER(Guidance), vp]),
    "dmc_ddim_step_at": (C.c_int, [vp, vp, vp, vp, vp, C.c_int32, C.c_int32, vp, vp, C.POINTER(Guidance), vp]),
    "dmc_ddpm_step_at": (C.c_int, [vp, vp, vp, vp, vp, C.c_int32, C.c_int32, vp, vp, C.POINTER(Guidance), vp]),
    "dmc_advance": (C.c_int, [vp, vp, vp, C.c_int32, vp]),
    "dmc_q_sample": (C.c_int, [vp, vp, vp, vp, vp, vp, C.c_int32, C.c_int32, vp]),
    "dmc_conv_wgrad_splits": (C.c_int, [C.POINTER(WgradDesc)]),
    "dmc_conv_wgrad": (C.c_int, [C.POINTER(WgradDesc), vp]),
    "dmc_gn_backward": (C.c_int, [C.POINTER(GnBwdDesc), vp]),
    "dmc_gn_backward_scratch": (C.c_int64, [C.POINTER(GnBwdDesc)]),
    "dmc_attention_backward": (C.c_int, [C.POINTER(AttnBwdDesc), vp]),
    "dmc_channel_sum": (C.c_int, [vp, vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, vp, vp]),
    "dmc_dilate2x": (C.c_int, [vp, vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, vp]),
    "dmc_dit_gate_ln_mod": (C.c_int, [C.POINTER(DitGlmDesc), vp]),
    "dmc_dit_gate_ln_mod_backward": (C.c_int, [C.POINTER(DitGlmBwdDesc), vp]),
    "dmc_gelu_forward": (C.c_int, [vp, vp, C.c_int64, C.c_float, C.c_uint32, vp]),
    "dmc_gelu_backward": (C.c_int, [vp, vp, vp, C.c_int64, C.c_float, C.c_uint32, vp]),
    "dmc_pack_weights": (C.c_int, [vp, C.c_int32, vp]),
    "dmc_opt_grad_norm": (C.c_int, [vp, vp, C.c_int32, vp, vp, vp]),
    "dmc_opt_adamw_step": (C.c_int, [vp, vp, C.c_int32, C.POINTER(AdamWDesc), vp, vp]),
    "dmc_add_bf16": (C.c_int, [vp, vp, C.c_int64, C.c_int32, vp]),
    "dmc_block_sum2x2": (C.c_int, [vp, vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, vp]),
    "dmc_nchw_f32_to_nhwc_bf16": (C.c_int, [vp, vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, vp]),
    "dmc_conv_dgrad_strided": (C.c_int, [vp, vp, vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                         C.c_int32, vp]),
    "dmc_plan_create": (C.c_int, [C.POINTER(vp)]),
    "dmc_plan_destroy": (C.c_int, [vp]),
    "dmc_plan_run": (C.c_int, [vp, vp]),
    "dmc_plan_run_op": (C.c_int, [vp, C.c_int32, vp]),
    "dmc_plan_num_launches": (C.c_int, [vp]),
    "dmc_plan_gemm_flops": (C.c_double, [vp]),
    "dmc_plan_set_seed": (C.c_int, [vp, C.c_int32, C.c_uint32]),
    "dmc_plan_rebind": (C.c_int, [vp, C.c_int32, C.c_int32, vp]),
    "dmc_plan_time_ops": (C.c_int, [vp, vp, C.c_int32, c_f32p, C.c_int32]),
    "dmc_plan_num_ops": (C.c_int, [vp]),
    "dmc_plan_op_kind": (C.c_int, [vp, C.c_int32]),
    "dmc_plan_op_flops": (C.c_double, [vp, C.c_int32]),
    "dmc_plan_op_bytes": (C.c_double, [vp, C.c_int32]),
    "dmc_plan_add_memset": (C.c_int, [vp, vp, C.c_size_t]),
    "dmc_plan_add_cond": (C.c_int, [vp, C.POINTER(CondDesc)]),
    "dmc_plan_add_stem": (C.c_int, [vp, C.POINTER(StemDesc)]),
    "dmc_plan_add_gn_stats": (C.c_int, [vp, C.POINTER(GnStatsDesc)]),
    "dmc_plan_add_gn_apply": (C.c_int, [vp, C.POINTER(GnApplyDesc)]),
    "dmc_plan_add_conv": (C.c_int, [vp, C.POINTER(ConvDesc)]),
    "dmc_plan_add_attention": (C.c_int, [vp, C.POINTER(AttnDesc)]),
    "dmc_plan_add_dit_cond": (C.c_int, [vp, C.POINTER(DitCondDesc)]),
    "dmc_plan_add_patch_embed": (C.c_int, [vp, C.POINTER(PatchEmbedDesc)]),
    "dmc_plan_add_ln_modulate": (C.c_int, [vp, C.POINTER(LnModDesc)]),
    "dmc_plan_add_head": (C.c_int, [vp, C.POINTER(HeadDesc)]),
    "dmc_head_supported": (C.c_int, [C.POINTER(HeadDesc)]),
    "dmc_conv_gn_supported": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "dmc_plan_add_upsample": (C.c_int, [vp, C.POINTER(UpsampleDesc)]),
    "dmc_plan_add_head_taps": (C.c_int, [vp, C.POINTER(HeadTapsDesc)]),
    "dmc_plan_add_stem_cols": (C.c_int, [vp, C.POINTER(StemColsDesc)]),
    "dmc_plan_add_gn_coeff": (C.c_int, [vp, C.POINTER(GnCoeffDesc)]),
    "dmc_conv_affine_supported": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "dmc_plan_add_ddim_step": (C.c_int, [vp, C.POINTER(StepDesc)]),
    "dmc_plan_add_ddpm_step": (C.c_int, [vp, C.POINTER(StepDesc)]),
}

OP_KINDS = ["memset", "cond", "stem", "gn_stats", "gn_apply", "conv", "attention", "upsample", "ddim", "ddpm", "dit_cond",
            "patch_embed", "ln_modulate", "head", "head_taps", "stem_cols", "gn_coeff"]

_lock = threading.Lock()
_lib = None
_inited = False


def load(require_device: bool = True):
    """Returns the loaded library; raises DmcError when it (or, if require_device, an sm_100 GPU) is missing."""
    global _lib, _inited
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise DmcError(
                    f"{LIB_PATH} is missing: build it with `python -m diffusion_models_collection_b200.build` "
                    "(there is no CPU / PyTorch fallback for the denoising hot path)")
            try:
                lib = C.CDLL(LIB_PATH)
            except OSError as e:  # pragma: no cover
                raise DmcError(f"cannot load {LIB_PATH}: {e}") from e
            for name, (res, args) in SYMBOLS.items():
                fn = getattr(lib, name, None)
                if fn is None:
                    if os.environ.get("DMC_LIB"):
                        continue
                    raise DmcError(f"{LIB_PATH} does not export {name}")
                fn.restype, fn.argtypes = res, args
            if lib.dmc_abi_version() != 1:
                raise DmcError("libdmc_b200.so ABI version mismatch")
            _lib = lib
        if require_device and not _inited:
            import torch

            if not torch.cuda.is_available():
                raise DmcError("no CUDA device: the B200 hot path has no CPU fallback")
            r = _lib.dmc_init()
            if r < 0:
                raise DmcError("dmc_init failed: " + _lib.dmc_last_error().decode())
            _inited = True
        return _lib


def check(rc: int, what: str = "dmc call"):
    if rc < 0:
        raise DmcError(f"{what} failed: {_lib.dmc_last_error().decode()}")
    return rc


def stream_ptr():
    import torch

    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t, byte_offset: int = 0):
    """device pointer of a torch tensor (None -> NULL)"""
    if t is None:
        return None
    return t.data_ptr() + byte_offset
