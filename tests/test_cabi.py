"""CPU: the C-ABI library builds, loads, and exports every symbol include/smbv_b200.h declares; the host binding
fails loudly (no fallback) without a GPU; the module keeps the reference checkpoint ABI and error behaviour."""
import ctypes
import os
import re

import pytest
import torch

import __graft_entry__ as ge
from oracle import videomae_oracle as vo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    ge.build()
    import smb_vision_b200

    return smb_vision_b200.load()


def header_symbols():
    src = open(os.path.join(ROOT, "include", "smbv_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(smbv_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(lib):
    syms = header_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/smbv_b200.h but not exported"


def test_binding_covers_header(lib):
    from smb_vision_b200 import _lib

    declared = set(header_symbols()) - {"smbv_last_error"}
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)


def test_version_and_arch(lib):
    assert lib.smbv_version() >= 100
    assert lib.smbv_sm_arch() == 100


def test_argument_errors_are_reported_before_launch(lib):
    # negative return + message, no CUDA call needed
    rc = lib.smbv_mask_index(None, 1, 8, None, None, None, None, None)
    assert rc < 0 and b"null" in lib.smbv_last_error()
    rc = lib.smbv_layernorm_fwd(ctypes.c_void_p(16), ctypes.c_void_p(16), ctypes.c_void_p(16), 1e-5, 4, 6, ctypes.c_void_p(16), None, None, None)
    assert rc < 0 and b"d % 4" in lib.smbv_last_error()
    from smb_vision_b200._lib import GemmArgs

    g = GemmArgs()
    g.A = g.W = g.out = 16
    g.M, g.N, g.K, g.lda, g.ldw, g.ldo = 128, 100, 64, 64, 64, 100
    assert lib.smbv_gemm_bf16(ctypes.byref(g), None) < 0 and b"multiple of 32" in lib.smbv_last_error()


def test_no_cpu_fallback():
    from smb_vision_b200 import SmbvError, ops

    with pytest.raises(SmbvError):
        ops.layernorm_fwd(torch.zeros(4, 8), torch.ones(8), torch.zeros(8), 1e-5)
    with pytest.raises(SmbvError):
        ops.gemm(torch.zeros(4, 8, dtype=torch.bfloat16), torch.zeros(32, 8, dtype=torch.bfloat16))


def test_missing_library_fails_loudly(monkeypatch):
    from smb_vision_b200 import _lib

    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libsmbv_b200.so")
    with pytest.raises(_lib.SmbvError, match="not built"):
        _lib.load()


def test_product_package_never_imports_oracle():
    pkg = os.path.join(ROOT, "smb-vision_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f


@pytest.mark.parametrize("cfgd", [vo.TINY, ge.SMALL64, {}])
def test_checkpoint_abi_matches_reference(cfgd):
    """state-dict keys/shapes == the reference's, so load_state_dict(strict=True) works both ways."""
    from smb_vision_b200.modeling import B200VideoMAEForPreTraining

    ocfg = vo.OracleConfig(**cfgd)
    full = {k: getattr(ocfg, k) for k in ocfg.__dataclass_fields__}
    with torch.device("meta"):
        m = B200VideoMAEForPreTraining(ge.hf_config(full))
    got = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    want = {k: tuple(v) for k, v in vo.param_shapes(ocfg).items()}
    assert got == want
    if not cfgd:
        assert sum(v.numel() for v in m.state_dict().values()) == 97_161_088
        assert sum(v.numel() for v in m.videomae.state_dict().values()) == 88_191_744


def test_reference_error_behaviour():
    from smb_vision_b200.modeling import B200VideoMAEForPreTraining

    m = B200VideoMAEForPreTraining(ge.hf_config(ge.SMALL64))
    x = torch.zeros(1, 96, 1, 96, 96)
    with pytest.raises(ValueError, match="channel dimension"):  # modeling_videomae.py:181-184
        m.videomae(x.repeat(1, 1, 3, 1, 1))
    with pytest.raises(ValueError, match="doesn't match model"):  # :185-188
        m.videomae(x[..., :80])
    with pytest.raises(ValueError, match="boolean mask"):  # :807-808
        m(x, None)
    with pytest.raises(ValueError):
        m(x, torch.zeros(1, 216, dtype=torch.bool), head_mask=torch.ones(2))
    ragged = torch.zeros(2, 216, dtype=torch.bool)
    ragged[0, :8] = True
    with pytest.raises(RuntimeError):  # reshape failure of :137
        m(x.repeat(2, 1, 1, 1, 1), ragged)
