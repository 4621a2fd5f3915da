"""ctypes binding of lib/libfadernets_b200.so (the C ABI declared in include/fadernets_b200.h).

There is NO fallback: if the shared library is missing or a call fails, an exception is
raised.  torch is used only for device memory, streams and autograd bookkeeping.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "lib", "libfadernets_b200.so")

c_f32p = C.c_void_p
c_ll = C.c_longlong


class FnGruChain(C.Structure):
    """Mirror of `struct FnGruChain` (include/fadernets_b200.h)."""
    _fields_ = [
        ("w_hh", C.c_void_p), ("b_hh", C.c_void_p),
        ("emb", C.c_void_p), ("ids", C.c_void_p), ("proj", C.c_void_p), ("proj_ld", c_ll),
        ("dense", C.c_void_p), ("h0", C.c_void_p), ("reverse", C.c_int32), ("_pad0", C.c_int32),
        ("hs", C.c_void_p), ("gates", C.c_void_p), ("h_final", C.c_void_p), ("h_final_ld", c_ll),
        ("dhs", C.c_void_p), ("dh_final", C.c_void_p), ("dh_final_ld", c_ll),
        ("dgh", C.c_void_p), ("dgin", C.c_void_p), ("dh0", C.c_void_p), ("dh_carry", C.c_void_p),
    ]


class FnGruChainBf16(C.Structure):
    """Mirror of `struct FnGruChainBf16` (include/fadernets_b200.h)."""
    _fields_ = [
        ("w_hh", C.c_void_p), ("w_hh_t", C.c_void_p), ("b_hh", C.c_void_p),
        ("emb", C.c_void_p), ("ids", C.c_void_p), ("proj", C.c_void_p), ("proj_ld", c_ll),
        ("dense", C.c_void_p), ("reverse", C.c_int32), ("dhs_f32", C.c_int32),
        ("hsx", C.c_void_p), ("gates", C.c_void_p), ("h_final", C.c_void_p), ("h_final_ld", c_ll),
        ("dhs", C.c_void_p), ("dh_final", C.c_void_p), ("dh_final_ld", c_ll),
        ("dg", C.c_void_p), ("dh0", C.c_void_p),
    ]


V, I, LL, F, SZ = C.c_void_p, C.c_int, C.c_longlong, C.c_float, C.c_size_t

# name -> (restype, argtypes); int-returning functions are status codes and are checked.
_SIGNATURES = {
    "fn_last_error": (C.c_char_p, []),
    "fn_abi_version": (I, []),
    "fn_source_hash": (C.c_char_p, []),
    "fn_device_info": (I, [C.POINTER(I)] * 4),
    "fn_gemm_f32": (I, [V, LL, LL, V, LL, LL, V, LL, V, I, I, I, I, V]),
    "fn_gemm_f32_splitk_ws_bytes": (SZ, [I, I, I]),
    "fn_gemm_f32_splitk": (I, [V, LL, LL, V, LL, LL, V, LL, V, I, I, I, I, I, V, SZ, V]),
    "fn_tc_gemm_bf16": (I, [V, LL, I, V, LL, I, V, LL, I, V, I, I, I, I, V]),
    "fn_tc_gemm_splitk_ws_bytes": (SZ, [I, I, I]),
    "fn_tc_gemm_bf16_splitk": (I, [V, LL, I, V, LL, I, V, LL, I, V, I, I, I, I, I, V, SZ, V]),
    "fn_gru_seq_fwd_f32": (I, [C.POINTER(FnGruChain), I, I, I, I, V, SZ, V]),
    "fn_gru_seq_bwd_f32": (I, [C.POINTER(FnGruChain), I, I, I, I, V, SZ, V]),
    "fn_gru_seq_ctas_per_chain": (I, [I]),
    "fn_gru_seq_fwd_bf16": (I, [C.POINTER(FnGruChainBf16), I, I, I, I, V, SZ, V]),
    "fn_gru_seq_bwd_bf16": (I, [C.POINTER(FnGruChainBf16), I, I, I, I, V, SZ, V]),
    "fn_gru_seq_fwd_bf16x3": (I, [C.POINTER(FnGruChainBf16), I, I, I, I, V, SZ, V]),
    "fn_gru_seq_bwd_bf16x3": (I, [C.POINTER(FnGruChainBf16), I, I, I, I, V, SZ, V]),
    "fn_tc_gemm_bf16x3": (I, [V, LL, LL, I, V, LL, LL, I, V, LL, I, V, I, I, I, I, I, V, SZ, V]),
    "fn_split_bf16": (I, [V, LL, LL, V, LL, LL, LL, LL, LL, V]),
    "fn_time_sum_bf16x3": (I, [V, I, I, I, V, V, V]),
    "fn_gru_debug_timeline": (I, [V]),
    "fn_decode_greedy_ws_bytes": (SZ, [I, I, I, I]),
    "fn_decode_greedy_bf16": (I, [V, V, V, V, V, V, V, V, V, V, V, V, V, I, I, I, I, I, V, V, V, SZ, V]),
    "fn_cast_bf16": (I, [V, LL, LL, V, LL, LL, LL, V]),
    "fn_ids_to_onehot_bf16": (I, [V, LL, I, LL, V, V]),
    "fn_time_sum_bf16": (I, [V, I, I, I, V, V, V]),
    "fn_col_sum_bf16": (I, [V, LL, LL, I, V, I, V, SZ, V]),
    "fn_add_f32_to_bf16": (I, [V, V, LL, V]),
    "fn_onehot_to_ids": (I, [V, I, I, I, V, V]),
    "fn_ids_to_onehot": (I, [V, I, I, I, V, V]),
    "fn_ids_to_time_major": (I, [V, I, I, I, I, V, V]),
    "fn_transpose_f32": (I, [V, LL, V, LL, I, I, I, V]),
    "fn_relu_mask_fwd": (I, [V, V, V, LL, V]),
    "fn_relu_mask_bwd": (I, [V, V, V, V, LL, V]),
    "fn_scale_f32": (I, [V, V, F, LL, V]),
    "fn_mse_mean_fwd": (I, [V, V, LL, V, V]),
    "fn_mse_mean_bwd": (I, [V, V, LL, V, V, V]),
    "fn_check_index_i64": (I, [V, LL, LL, V, V]),
    "fn_clamp_index_i32": (I, [V, LL, I, V, V]),
    "fn_clean_tokens": (I, [V, I, I, V, V, V]),
    "fn_add_f32": (I, [V, V, LL, V]),
    "fn_emb_grad_scratch_bytes": (SZ, [I, I, I, I]),
    "fn_emb_grad_f32": (I, [V, V, V, I, I, I, I, V, V, SZ, V]),
    "fn_time_sum_f32": (I, [V, V, I, I, I, V, V, V]),
    "fn_col_sum_scratch_bytes": (SZ, [LL, I]),
    "fn_col_sum_f32": (I, [V, LL, LL, I, V, I, V, SZ, V]),
    "fn_vocab_logsoftmax_fwd": (I, [V, I, I, I, V, V]),
    "fn_vocab_logsoftmax_bwd": (I, [V, V, I, I, I, V, V]),
    "fn_vocab_nll_fwd": (I, [V, V, I, I, I, V, V, V, V]),
    "fn_vocab_nll_bwd": (I, [V, V, V, V, F, I, I, I, V, V]),
    "fn_time_logsoftmax_fwd": (I, [V, I, I, I, V, V]),
    "fn_time_logsoftmax_bwd": (I, [V, V, I, I, I, V, V]),
    "fn_nll_mean_fwd": (I, [V, V, LL, I, V, V, SZ, V]),
    "fn_nll_mean_bwd": (I, [V, LL, I, V, V, I, V]),
    "fn_reduce_scratch_bytes": (SZ, [LL]),
    "fn_sum_f32": (I, [V, LL, F, V, V, SZ, V]),
    "fn_reparam_fwd": (I, [V, V, V, LL, V, V, V]),
    "fn_reparam_bwd": (I, [V, V, V, V, V, LL, V, V, V]),
    "fn_qy_x_fwd": (I, [V, V, V, I, I, I, V, V, V, V]),
    "fn_qy_x_bwd": (I, [V, V, V, V, V, V, I, I, I, V, V, V, SZ, V]),
    "fn_gm_kl_fwd": (I, [V, V, V, V, V, V, V, I, I, I, I, V, V, SZ, V]),
    "fn_latent_scratch_bytes": (SZ, [I, I, I]),
    "fn_gm_kl_bwd": (I, [V, V, V, V, V, V, V, I, I, I, I, V, V, V, V, V, V, V, SZ, V]),
    "fn_std_kl_fwd": (I, [V, V, LL, V, V, SZ, V]),
    "fn_std_kl_bwd": (I, [V, V, LL, V, V, V, V]),
    "fn_latent_reg_fwd": (I, [V, LL, V, I, V, V, V, V]),
    "fn_latent_reg_bwd": (I, [V, V, I, I, V, V]),
    "fn_grad_norm": (I, [V, LL, V, V, SZ, V]),
    "fn_clip_adam": (I, [V, V, V, V, LL, V, F, F, F, F, F, I, V]),
}
_UNCHECKED = {"fn_last_error", "fn_abi_version", "fn_source_hash", "fn_latent_scratch_bytes", "fn_gru_seq_ctas_per_chain", "fn_emb_grad_scratch_bytes", "fn_tc_gemm_splitk_ws_bytes", "fn_gemm_f32_splitk_ws_bytes", "fn_decode_greedy_ws_bytes",
              "fn_col_sum_scratch_bytes", "fn_reduce_scratch_bytes"}


class FaderNetsError(RuntimeError):
    pass


class _Lib:
    def __init__(self):
        self._dll = None
        self.launches = 0          # number of C-ABI compute calls issued (bench's gpu_launches claim)

    def load(self):
        if self._dll is not None:
            return self
        if not os.path.exists(LIB_PATH):
            raise FaderNetsError(
                f"{LIB_PATH} is missing -- build it with `python __graft_entry__.py build` "
                "(nvcc, sm_100a).  There is no CPU / PyTorch fallback for this path.")
        dll = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(dll, name)          # AttributeError if the symbol is not exported
            fn.restype, fn.argtypes = res, args
        # the binary must have been built from the sources it ships with (no stale .so on the GPU box)
        built = dll.fn_source_hash().decode()
        try:
            import importlib.util
            root = os.path.dirname(os.path.dirname(_HERE))
            spec = importlib.util.spec_from_file_location("_fn_graft_entry", os.path.join(root, "__graft_entry__.py"))
            ge = importlib.util.module_from_spec(spec); spec.loader.exec_module(ge)
            want = ge.source_hash()
        except Exception:                       # package used outside the repo tree: nothing to compare with
            want = None
        if want is not None and built != want:
            raise FaderNetsError(f"{LIB_PATH} was built from other sources (library {built}, tree {want}): "
                                 "run `python __graft_entry__.py build`")
        self.source_hash = built
        self._dll = dll
        return self

    @property
    def dll(self):
        return self.load()._dll

    def call(self, name, *args):
        fn = getattr(self.dll, name)
        rc = fn(*args)
        if name not in _UNCHECKED:
            self.launches += 1
            if rc != 0:
                raise FaderNetsError(f"{name} failed ({rc}): {self.dll.fn_last_error().decode()}")
        return rc


LIB = _Lib()


def symbols():
    """All C-ABI symbols the Python side binds (== those declared in the header)."""
    return sorted(_SIGNATURES)


def ptr(t):
    """Raw device pointer of a tensor (None -> NULL)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr(device=None):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise FaderNetsError("fadernets_b200 runs on CUDA (B200) only: got a CPU tensor. "
                                 "There is no CPU fallback; move the model and inputs to cuda.")
