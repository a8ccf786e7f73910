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
