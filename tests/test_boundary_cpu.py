"""CPU checks of the drop-in boundary: parameter names/shapes equal the reference's, ctor kwargs are swallowed,
the library exports every declared symbol, and the product refuses to run without CUDA (no silent fallback)."""
import ctypes
import json
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def _meta(name):
    with open(os.path.join(GOLDEN, name + ".json")) as f:
        return json.load(f)


@pytest.mark.parametrize("case", ["t2v_small", "t2v_config1", "i2v_small", "i2v_config1"])
def test_state_dict_matches_reference(case):
    from videomv_b200 import unet
    meta = _meta(case)
    cls = unet.UNetSD_T2VBase if meta["kind"] == "t2v" else unet.UNetSD_I2VGen
    with torch.device("meta"):
        model = cls(**meta["kwargs"])
    mine = {k: list(v.shape) for k, v in model.state_dict().items()}
    ref = meta["shapes"]
    assert set(mine) == set(ref), (sorted(set(ref) - set(mine))[:5], sorted(set(mine) - set(ref))[:5])
    bad = [k for k in ref if mine[k] != ref[k]]
    assert not bad, bad[:5]
    assert list(mine) == list(ref)      # same ordering too (checkpoint tools sometimes zip by position)


def test_vae_state_dict_matches_reference():
    """AutoencoderKL drop-in: same keys, shapes and order as tools/modules/autoencoder.py:32."""
    from videomv_b200 import vae
    meta = _meta("vae_small")
    with torch.device("meta"):
        model = vae.AutoencoderKL(**meta["kwargs"])
    mine = {k: list(v.shape) for k, v in model.state_dict().items()}
    assert mine == meta["shapes"] and list(mine) == list(meta["shapes"])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        vae.AutoencoderKL(**meta["kwargs"]).decode(torch.zeros(1, 4, 8, 8))


def test_ctor_swallows_unknown_kwargs_and_yaml_types():
    from videomv_b200 import unet
    kw = dict(_meta("t2v_small")["kwargs"], some_future_flag=3, use_lgm_refine=True)
    with torch.device("meta"):
        m = unet.UNetSD_T2VBase(**kw)
    assert m.use_lgm_refine is True


def test_library_exports_every_declared_symbol():
    from videomv_b200 import _lib
    _lib.build()
    header = open(os.path.join(ROOT, "include", "videomv_b200.h")).read()
    declared = set(re.findall(r"\b(vmv_[a-z0-9_]+)\s*\(", header))
    declared -= {"vmv_gemm_params", "vmv_attn_params"}
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    L = ctypes.CDLL(_lib.LIBPATH)
    for sym in declared:
        assert hasattr(L, sym), sym
    assert _lib.lib().vmv_abi_version() == 6
    assert ctypes.sizeof(_lib.GemmParams) == _lib.lib().vmv_sizeof_gemm_params()
    assert ctypes.sizeof(_lib.AttnParams) == _lib.lib().vmv_sizeof_attn_params()
    assert ctypes.sizeof(_lib.PeerExchangeParams) == _lib.lib().vmv_sizeof_peer_exchange_params()
    assert ctypes.sizeof(_lib.PeerAllreduceParams) == _lib.lib().vmv_sizeof_peer_allreduce_params()
    assert ctypes.sizeof(_lib.GnPeer) == _lib.lib().vmv_sizeof_gn_peer()
    assert ctypes.sizeof(_lib.GemmScatter) == _lib.lib().vmv_sizeof_gemm_scatter()
    assert ctypes.sizeof(_lib.PeerAllgatherParams) == _lib.lib().vmv_sizeof_peer_allgather_params()


def test_cpu_call_fails_loudly():
    from videomv_b200 import unet
    kw = _meta("t2v_small")["kwargs"]
    m = unet.UNetSD_T2VBase(**kw)
    x = torch.zeros(1, 4, 2, 8, 8)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(x, torch.tensor([1]), y=torch.zeros(1, 77, 1024))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "videomv_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in src.replace("no oracle", ""), fn


@pytest.mark.parametrize("ci,co", [(8, 6), (64, 16)])
def test_upsample_conv_phase_weights_reproduce_interpolate_then_conv(ci, co):
    """packing.pack_upconv3x3 (host logic of the UPCONV3X3 GEMM mode): four 2x2 phase convs with pre-summed taps on the
    original image == F.interpolate(scale 2, nearest) followed by Conv2d 3x3 pad 1 (util.py:604-606), here in fp64 on the CPU."""
    import torch.nn.functional as F
    from videomv_b200 import packing
    g = torch.Generator().manual_seed(0)
    w = torch.randn(co, ci, 3, 3, generator=g, dtype=torch.float64)
    x = torch.randn(2, ci, 5, 7, generator=g, dtype=torch.float64)
    ref = F.conv2d(F.interpolate(x, scale_factor=2.0, mode="nearest"), w, padding=1)
    # the packed matrix in fp64 (pack_upconv3x3 rounds to fp16: redo its summation without the final cast)
    wp = packing.pack_upconv3x3(w.float()).double().reshape(4, co, 2, 2, ci)
    exact = torch.empty_like(wp)
    groups = {0: ([0], [1, 2]), 1: ([0, 1], [2])}
    for py in (0, 1):
        for px in (0, 1):
            for ty in (0, 1):
                for tx in (0, 1):
                    exact[2 * py + px, :, ty, tx] = sum(w[:, :, ky, kx] for ky in groups[py][ty] for kx in groups[px][tx])
    assert torch.allclose(wp, exact, rtol=2e-3, atol=2e-3)          # fp16 rounding of the packed weights only
    out = torch.zeros_like(ref)
    xp = F.pad(x, (1, 1, 1, 1))
    for py in (0, 1):
        for px in (0, 1):
            acc = torch.zeros(2, co, 5, 7, dtype=torch.float64)
            for ty in (0, 1):
                for tx in (0, 1):
                    dy, dx = ty - 1 + py, tx - 1 + px                 # input offset of this tap
                    patch = xp[:, :, 1 + dy:1 + dy + 5, 1 + dx:1 + dx + 7]
                    acc += torch.einsum("bchw,oc->bohw", patch, exact[2 * py + px, :, ty, tx])
            out[:, :, py::2, px::2] = acc
    assert torch.allclose(out, ref, rtol=1e-12, atol=1e-12)


def test_staged_reference_files_are_verbatim_copies():
    """oracle/stage_ref.py (baseline arms on the GPU box): the staged files are byte-identical to the reference tree."""
    from oracle import stage_ref
    if not os.path.isfile(os.path.join(stage_ref.SRC, stage_ref.FILES[0])):
        pytest.skip("reference tree not present")
    assert stage_ref.stage()
    for rel in stage_ref.FILES:
        assert open(os.path.join(stage_ref.SRC, rel), "rb").read() == open(os.path.join(stage_ref.DST, rel), "rb").read(), rel
