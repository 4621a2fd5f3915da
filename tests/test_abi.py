"""CPU-side checks of the boundary: the C-ABI library builds, loads and exports every symbol that
include/fadernets_b200.h declares; the Python mirror keeps the reference's surface; and the
product path refuses to run without CUDA instead of falling back."""
import os
import re
import subprocess

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "fadernets_b200.h")


def header_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fn_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(lib):
    import fadernets_b200 as fn
    out = subprocess.run(["nm", "-D", "--defined-only", fn.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r"\bT (fn_[a-z0-9_]+)", out))
    declared = header_symbols()
    assert declared, "no declarations parsed"
    missing = [s for s in declared if s not in exported]
    assert not missing, f"declared but not exported: {missing}"
    # and the Python binding covers the whole header (and nothing else)
    assert sorted(fn.symbols()) == declared


def test_abi_version_and_error_string(lib):
    assert lib.call("fn_abi_version") == 3
    assert isinstance(lib.dll.fn_last_error(), bytes)


def test_library_is_sm100a_only(lib):
    import fadernets_b200 as fn
    out = subprocess.run(["cuobjdump", "-lelf", fn.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_struct_layout_matches_header():
    import ctypes
    from fadernets_b200._lib import FnGruChain
    src = open(HEADER).read()
    body = src[src.index("typedef struct FnGruChain {"):src.index("} FnGruChain;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = re.findall(r"(\w+)\s*;", body)
    assert names == [f[0] for f in FnGruChain._fields_]
    assert ctypes.sizeof(FnGruChain) == 8 * len(names) - 4 * 0 - 0 or ctypes.sizeof(FnGruChain) % 8 == 0


def test_bf16_struct_layout_matches_header():
    import ctypes
    from fadernets_b200._lib import FnGruChainBf16
    src = open(HEADER).read()
    body = src[src.index("typedef struct FnGruChainBf16 {"):src.index("} FnGruChainBf16;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = re.findall(r"(\w+)\s*;", body)
    assert names == [f[0] for f in FnGruChainBf16._fields_]
    # two int32 share one 8-byte slot; everything else is pointer / long long
    assert ctypes.sizeof(FnGruChainBf16) == 8 * (len(names) - 1)


def test_model_surface_matches_reference_contract():
    import fadernets_b200 as fn
    from oracle import fader_oracle as fo
    torch.manual_seed(0)
    m = fn.MusicAttrRegGMVAE(342, 3, 16, 24, 16, 8, 32, n_component=3)
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == fo.param_shapes(16, 8, "gmvae", 3)
    assert list(m.state_dict()) == list(fo.param_shapes(16, 8, "gmvae", 3))
    assert not m.logvar_r_lookup.weight.requires_grad and m.mu_r_lookup.weight.requires_grad
    assert float(m.logvar_n_lookup.weight[0, 0]) == pytest.approx(-4.0)
    for attr in ("n_component", "latent_dim", "roll_dims", "eps", "sample", "forward", "encode", "sub_decoders",
                 "global_decoder", "approx_qy_x", "_sampling", "mu_r_lookup", "logvar_n_lookup"):
        assert hasattr(m, attr), attr
    v = fn.MusicAttrRegVAE(342, 3, 16, 24, 16, 8, 32)
    assert list(v.state_dict()) == list(fo.param_shapes(16, 8, "vae"))
    for attr in ("encoder", "sub_decoders", "global_decoder", "iteration", "k", "n_step", "z_dims", "hidden_dims"):
        assert hasattr(v, attr), attr
    live = sorted(n for n, _ in m.live_parameters())
    assert live == sorted(k for k in m.state_dict() if fo.is_live(k))


def test_trainer_mirror_signatures():
    import inspect
    from fadernets_b200 import trainer, trainer_gmm
    assert list(inspect.signature(trainer_gmm.train).parameters) == [
        "step", "d_oh", "r_oh", "n_oh", "d", "r", "n", "c", "r_density", "n_density", "is_supervised", "y_label"]
    assert list(inspect.signature(trainer_gmm.loss_function).parameters) == [
        "out", "d", "r_out", "r", "n_out", "n", "dis", "qy_x_out", "logLogit_out", "step", "beta", "is_supervised",
        "y_label"]
    assert list(inspect.signature(trainer.train).parameters) == [
        "step", "d_oh", "r_oh", "n_oh", "d", "r", "n", "c", "r_density", "n_density"]
    assert list(inspect.signature(trainer.evaluate).parameters) == [
        "d_oh", "r_oh", "n_oh", "d", "r", "n", "c", "r_density", "n_density"]
    assert list(inspect.signature(trainer.loss_function).parameters) == [
        "out", "d", "r_out", "r", "n_out", "n", "dis", "beta"]


def test_no_cpu_fallback():
    import fadernets_b200 as fn
    m = fn.MusicAttrRegVAE(342, 3, 16, 24, 16, 8, 32)
    with pytest.raises(fn.FaderNetsError):
        m(torch.zeros(2, 3, 342), torch.zeros(2, 3, 3), torch.zeros(2, 3, 16), torch.zeros(2, 24))
    with pytest.raises(fn.FaderNetsError):
        m.global_decoder(torch.zeros(2, 40), 3)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "music-fader-nets_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                assert "oracle" not in open(os.path.join(dp, f)).read(), f
