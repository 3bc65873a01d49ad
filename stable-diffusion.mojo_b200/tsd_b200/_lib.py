"""ctypes binding of libtsd_b200.so (the C ABI declared in include/tsd_b200.h).

The library is the product: if it cannot be loaded, or no sm_100 device is present, every
entry point fails loudly - there is no CPU / PyTorch fallback anywhere in this package.
"""
from __future__ import annotations

import ctypes as C
import os
import re
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(os.path.dirname(_HERE), "csrc")
LIB_PATH = os.environ.get("TSD_LIB", os.path.join(CSRC, "libtsd_b200.so"))  # TSD_LIB: lab builds (e.g. the in-kernel trace variant)
HEADER = os.path.join(os.path.dirname(os.path.dirname(_HERE)), "include", "tsd_b200.h")

c_float_p = C.POINTER(C.c_float)
c_double_p = C.POINTER(C.c_double)
c_i32_p = C.POINTER(C.c_int32)
c_i64_p = C.POINTER(C.c_int64)


class TsdError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"tsd_b200 error {code}: {msg}")
        self.code = code
        self.msg = msg


class DiffusionConfig(C.Structure):
    _fields_ = [
        ("latent_h", C.c_int32),
        ("latent_w", C.c_int32),
        ("max_batch", C.c_int32),
        ("context_len", C.c_int32),
        ("context_dim", C.c_int32),
        ("mojo_alias_time", C.c_int32),
        ("norm_affine", C.c_int32),
    ]


class LoopParams(C.Structure):
    _fields_ = [
        ("steps", C.c_int32),
        ("cfg", C.c_int32),
        ("cfg_scale", C.c_float),
        ("timesteps", c_i32_p),
        ("time_emb", c_float_p),
        ("coef", c_float_p),
        ("noise", c_float_p),
    ]


def build(force: bool = False) -> str:
    """Compile the CUDA sources in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    if force:
        subprocess.run(["rm", "-rf", os.path.join(CSRC, "build"), LIB_PATH], check=True)
    subprocess.run(["bash", os.path.join(CSRC, "build.sh")], check=True)
    return LIB_PATH


def declared_symbols() -> list[str]:
    """Every function name include/tsd_b200.h declares."""
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tsd_[a-z0-9_]+)\s*\(", text)))


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise TsdError(-1, f"{LIB_PATH} is missing: run __graft_entry__.build() / csrc/build.sh "
                               "(no fallback path exists)")
        _lib = C.CDLL(LIB_PATH)
        _declare(_lib)
    return _lib


def _declare(L: C.CDLL) -> None:
    vp = C.c_void_p
    i32, i64, f32 = C.c_int32, C.c_int64, C.c_float
    fp = C.c_void_p  # float* passed as raw addresses (numpy .ctypes.data / torch .data_ptr())

    def sig(name, restype, *argtypes):
        if hasattr(L, name):
            fn = getattr(L, name)
            fn.restype = restype
            fn.argtypes = list(argtypes)

    sig("tsd_init", i32, i32, C.POINTER(vp))
    sig("tsd_shutdown", i32, vp)
    sig("tsd_last_error", C.c_char_p, vp)
    sig("tsd_synchronize", i32, vp)
    sig("tsd_set_option", i32, vp, C.c_char_p, i32)
    sig("tsd_get_option", i32, vp, C.c_char_p, c_i32_p)
    sig("tsd_launch_count", i64, vp)
    sig("tsd_timer_start", i32, vp)
    sig("tsd_timer_stop", i32, vp, c_double_p)
    sig("tsd_conv2d", i32, vp, fp, i32, i32, i32, i32, fp, fp, i32, i32, i32, i32, fp)
    sig("tsd_conv2d_pad", i32, vp, fp, i32, i32, i32, i32, fp, fp, i32, i32, i32, i32, i32, fp)
    sig("tsd_linear", i32, vp, fp, i32, i32, i32, fp, fp, i32, fp)
    sig("tsd_matmul", i32, vp, fp, fp, i32, i32, i32, i32, fp)
    sig("tsd_groupnorm", i32, vp, fp, i32, i32, i32, i32, i32, f32, fp, fp, fp)
    sig("tsd_layernorm", i32, vp, fp, i32, i32, fp)
    sig("tsd_silu", i32, vp, fp, i64, fp)
    sig("tsd_gelu", i32, vp, fp, i64, fp)
    sig("tsd_upsample2x", i32, vp, fp, i32, i32, i32, fp)
    sig("tsd_softmax", i32, vp, fp, i32, i32, i32, i32, fp)
    sig("tsd_self_attention", i32, vp, fp, i32, i32, i32, fp, fp, fp, fp, fp)
    sig("tsd_cross_attention", i32, vp, fp, i32, i32, fp, i32, i32, i32, fp, fp, fp, fp, fp, fp, fp,
        fp, fp)
    sig("tsd_attention_core", i32, vp, fp, fp, fp, i32, i32, i32, i32, fp)
    sig("tsd_attention_core_dev", i32, vp, fp, fp, fp, i32, i32, i32, i32, fp)
    sig("tsd_sampler_step", i32, vp, fp, fp, fp, f32, fp, f32, f32, f32, f32, f32, i64, fp)
    sig("tsd_sampler_add_noise", i32, vp, fp, fp, f32, f32, i64, fp)
    sig("tsd_sampler_step_dev", i32, vp, fp, fp, fp, f32, fp, f32, f32, f32, f32, f32, i64, fp)
    sig("tsd_diffusion_create", i32, vp, C.POINTER(DiffusionConfig), C.POINTER(vp))
    sig("tsd_diffusion_destroy", i32, vp)
    sig("tsd_diffusion_num_params", i64, vp)
    sig("tsd_diffusion_load_weights", i32, vp, fp, i64)
    sig("tsd_diffusion_init_random", i32, vp, C.c_uint64)
    sig("tsd_diffusion_param_count", i32, vp)
    sig("tsd_diffusion_param_name", C.c_char_p, vp, i32, c_i64_p, c_i64_p)
    sig("tsd_diffusion_get_param", i32, vp, i32, fp)
    sig("tsd_diffusion_forward", i32, vp, fp, fp, i32, fp, i32, i32, fp)
    sig("tsd_diffusion_forward_dev", i32, vp, fp, fp, i32, fp, i32, i32, fp)
    sig("tsd_diffusion_step", i32, vp, fp, fp, i32, fp, fp, i32, f32, f32, f32, f32, f32, f32, i32, fp)
    sig("tsd_diffusion_profile", i32, vp, fp, fp, i32, fp, i32, i32, fp, c_double_p, c_double_p,
        c_i64_p)
    i32p = C.POINTER(C.c_int32)
    sig("tsd_tokenizer_load", i32, C.c_char_p, i32, C.POINTER(vp))
    sig("tsd_tokenizer_from_memory", i32, vp, i64, i32, C.POINTER(vp))
    sig("tsd_tokenizer_destroy", i32, vp)
    sig("tsd_tokenizer_vocab_size", i32, vp)
    sig("tsd_tokenizer_max_token_length", i32, vp)
    sig("tsd_tokenizer_token", vp, vp, i32, i32p, C.POINTER(C.c_float))
    sig("tsd_tokenizer_find", i32, vp, C.c_char_p, i32)
    sig("tsd_tokenizer_encode", i32, vp, C.c_char_p, i32, i32, i32p, i32, i32p)
    sig("tsd_safetensors_open", i32, C.c_char_p, C.POINTER(vp))
    sig("tsd_safetensors_from_memory", i32, C.c_char_p, i64, C.POINTER(vp))
    sig("tsd_safetensors_close", i32, vp)
    sig("tsd_safetensors_count", i32, vp)
    sig("tsd_safetensors_name", C.c_char_p, vp, i32)
    sig("tsd_safetensors_find", i32, vp, C.c_char_p)
    sig("tsd_safetensors_info", i32, vp, i32, C.c_char_p, i32p, C.POINTER(i64), C.POINTER(i64))
    sig("tsd_safetensors_read_f32", i32, vp, i32, fp, i64)
    sig("tsd_png_encode", i32, fp, i32, i32, i32, vp, i64, C.POINTER(i64))
    sig("tsd_png_write", i32, C.c_char_p, fp, i32, i32, i32)
    sig("tsd_decoder_create", i32, vp, i32, i32, i32, C.POINTER(vp))
    sig("tsd_decoder_create_ex", i32, vp, i32, i32, i32, C.c_uint32, C.POINTER(vp))
    sig("tsd_decoder_destroy", i32, vp)
    sig("tsd_decoder_num_params", i64, vp)
    sig("tsd_decoder_load_weights", i32, vp, fp, i64)
    sig("tsd_decoder_init_random", i32, vp, C.c_uint64)
    sig("tsd_decoder_param_count", i32, vp)
    sig("tsd_decoder_param_name", C.c_char_p, vp, i32, c_i64_p, c_i64_p)
    sig("tsd_decoder_get_param", i32, vp, i32, fp)
    sig("tsd_decoder_forward", i32, vp, fp, i32, i32, fp)
    sig("tsd_decoder_forward_dev", i32, vp, fp, i32, i32, fp)
    sig("tsd_encoder_create", i32, vp, i32, i32, i32, C.POINTER(vp))
    sig("tsd_encoder_create_ex", i32, vp, i32, i32, i32, C.c_uint32, C.POINTER(vp))
    sig("tsd_encoder_destroy", i32, vp)
    sig("tsd_encoder_num_params", i64, vp)
    sig("tsd_encoder_load_weights", i32, vp, fp, i64)
    sig("tsd_encoder_init_random", i32, vp, C.c_uint64)
    sig("tsd_encoder_param_count", i32, vp)
    sig("tsd_encoder_param_name", C.c_char_p, vp, i32, c_i64_p, c_i64_p)
    sig("tsd_encoder_get_param", i32, vp, i32, fp)
    sig("tsd_encoder_forward", i32, vp, fp, fp, i32, i32, fp)
    sig("tsd_encoder_forward_dev", i32, vp, fp, fp, i32, i32, fp)
    sig("tsd_clip_create", i32, vp, i32, i32, C.POINTER(vp))
    sig("tsd_clip_create_ex", i32, vp, i32, i32, C.c_uint32, C.POINTER(vp))
    sig("tsd_clip_destroy", i32, vp)
    sig("tsd_clip_num_params", i64, vp)
    sig("tsd_clip_load_weights", i32, vp, fp, i64)
    sig("tsd_clip_init_random", i32, vp, C.c_uint64)
    sig("tsd_clip_param_count", i32, vp)
    sig("tsd_clip_param_name", C.c_char_p, vp, i32, c_i64_p, c_i64_p)
    sig("tsd_clip_get_param", i32, vp, i32, fp)
    sig("tsd_clip_forward", i32, vp, vp, i32, fp)
    sig("tsd_clip_forward_dev", i32, vp, vp, i32, fp)
    sig("tsd_generate_latents", i32, vp, C.POINTER(LoopParams), fp, fp, i32, i32, fp)
    sig("tsd_dist_unique_id", i32, C.c_char_p)
    sig("tsd_dist_init", i32, vp, i32, i32, C.c_char_p, C.c_char_p, C.POINTER(vp))
    sig("tsd_dist_shutdown", i32, vp)
    sig("tsd_dist_rank", i32, vp)
    sig("tsd_dist_size", i32, vp)
    sig("tsd_dist_broadcast_context", i32, vp, fp, i64, i32)
    sig("tsd_dist_gather", i32, vp, fp, i64, fp, i32)
    sig("tsd_dist_generate", i32, vp, vp, C.POINTER(LoopParams), fp, fp, i32, i32, i32, fp)
    sig("tsd_bench_gemm", i32, vp, i32, i32, i32, i32, i32, i32, i32, i32, c_double_p)
    sig("tsd_bench_attention", i32, vp, i32, i32, i32, i32, i32, c_double_p)
    sig("tsd_bench_norm", i32, vp, i32, i32, i32, i32, i32, i32, i32, c_double_p)
    sig("tsd_bench_conv", i32, vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, c_double_p)
