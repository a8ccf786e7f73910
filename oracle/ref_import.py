"""TEST INFRASTRUCTURE ONLY -- import the *unmodified* reference UNet classes.

Works only where /root/reference exists (the build container; NOT the GPU box).
Recipe from SURVEY.md section 8(c): stub the missing third-party packages, then load
`tools/modules/unet/{util,unet_t2v,unet_i2vgen}.py` directly from the read-only tree
(nothing is copied; `sys.dont_write_bytecode` keeps the tree clean).

Stubs (none of them carries arithmetic that the restatement does not restate):
  xformers.ops.memory_efficient_attention(q,k,v) := softmax(q k^T / sqrt(d)) v
      (xformers==0.0.13 semantics; same maths as util.py:396-427)
  fairscale.nn.checkpoint.checkpoint_wrapper      := identity
  rotary_embedding_torch / open_clip / kiui / easydict ... := empty modules (unused here)
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

REF_ROOT = os.environ.get("VIDEOMV_REFERENCE", "/root/reference")
if not os.path.isfile(os.path.join(REF_ROOT, "tools/modules/unet/unet_t2v.py")):
    # the GPU box: the unmodified files staged by oracle/stage_ref.py (git-ignored, travels with the snapshot)
    REF_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "tools/modules/unet/unet_t2v.py"))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _install_stubs():
    import torch
    import torch.nn.functional as F

    def memory_efficient_attention(q, k, v, attn_bias=None, op=None, p=0.0, scale=None):
        assert attn_bias is None
        if q.dim() == 3:
            return F.scaled_dot_product_attention(q, k, v)
        # [B, M, H, K] layout (core/attention.py) -- not on this path
        o = F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2))
        return o.transpose(1, 2)

    if "xformers" not in sys.modules:
        ops = _stub("xformers.ops", memory_efficient_attention=memory_efficient_attention,
                    LowerTriangularMask=type("LowerTriangularMask", (), {}),
                    unbind=torch.unbind)
        xf = _stub("xformers", ops=ops)
        xf.__path__ = []
    if "fairscale" not in sys.modules:
        ck = _stub("fairscale.nn.checkpoint", checkpoint_wrapper=lambda m, *a, **k: m)
        nn_ = _stub("fairscale.nn", checkpoint=ck)
        nn_.__path__ = []
        fs = _stub("fairscale", nn=nn_)
        fs.__path__ = []
    if "rotary_embedding_torch" not in sys.modules:
        _stub("rotary_embedding_torch", RotaryEmbedding=type("RotaryEmbedding", (), {"__init__": lambda s, *a, **k: None}))
    for name in ("open_clip", "kiui"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                _stub(name)


def _load(modname, relpath):
    spec = importlib.util.spec_from_file_location(modname, os.path.join(REF_ROOT, relpath))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[modname] = mod
    spec.loader.exec_module(mod)
    return mod


_CACHE = {}


def load_reference():
    """Returns (UNetSD_T2VBase, UNetSD_I2VGen) classes of the reference."""
    if _CACHE:
        return _CACHE["t2v"], _CACHE["i2v"]
    assert available(), f"reference tree not found at {REF_ROOT}"
    sys.dont_write_bytecode = True
    _install_stubs()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)           # for `utils.registry_class`
    for pkg in ("tools", "tools.modules", "tools.modules.unet"):
        if pkg not in sys.modules:
            m = types.ModuleType(pkg)
            m.__path__ = [os.path.join(REF_ROOT, *pkg.split("."))]
            sys.modules[pkg] = m
    util = _load("tools.modules.unet.util", "tools/modules/unet/util.py")
    sys.modules["tools.modules.unet"].util = util
    t2v = _load("tools.modules.unet.unet_t2v", "tools/modules/unet/unet_t2v.py")
    i2v = _load("tools.modules.unet.unet_i2vgen", "tools/modules/unet/unet_i2vgen.py")
    _CACHE["t2v"], _CACHE["i2v"] = t2v.UNetSD_T2VBase, i2v.UNetSD_I2VGen
    return _CACHE["t2v"], _CACHE["i2v"]


def load_reference_vae():
    """The reference's AutoencoderKL class (tools/modules/autoencoder.py), unmodified."""
    if "vae" in _CACHE:
        return _CACHE["vae"]
    load_reference()                      # stubs + sys.path for utils.registry_class
    mod = _load("tools.modules.autoencoder", "tools/modules/autoencoder.py")
    _CACHE["vae"] = mod.AutoencoderKL
    return _CACHE["vae"]


# cfg.auto_encoder of tools/modules/config.py:110-127 (the SD-1.x VAE every shipped config uses)
VAE_KWARGS = dict(ddconfig=dict(double_z=True, z_channels=4, resolution=256, in_channels=3, out_ch=3, ch=128,
                                ch_mult=[1, 2, 4, 4], num_res_blocks=2, attn_resolutions=[], dropout=0.0,
                                video_kernel_size=[3, 1, 1]), embed_dim=4)


def load_reference_ddim():
    """The reference's DiffusionDDIM class (tools/modules/diffusions/diffusion_ddim.py), unmodified."""
    if "ddim" in _CACHE:
        return _CACHE["ddim"]
    load_reference()                      # stubs + sys.path for utils.registry_class
    pk = "tools.modules.diffusions"
    if pk not in sys.modules:
        m = types.ModuleType(pk)
        m.__path__ = [os.path.join(REF_ROOT, "tools/modules/diffusions")]
        sys.modules[pk] = m
    for name in ("schedules", "losses", "diffusion_ddim"):
        if f"{pk}.{name}" not in sys.modules:
            _load(f"{pk}.{name}", f"tools/modules/diffusions/{name}.py")
    _CACHE["ddim"] = sys.modules[f"{pk}.diffusion_ddim"].DiffusionDDIM
    return _CACHE["ddim"]


# Resolved ctor kwargs (tools/modules/config.py:88-106 overlaid by configs/t2v_infer.yaml:20-41),
# with use_lgm_refine=False (LGM needs kiui/diff_gaussian_rasterization; not on this path).
T2V_KWARGS = dict(in_dim=4, dim=320, y_dim=1024, context_dim=1024, out_dim=4, dim_mult=[1, 2, 4, 4],
                  num_heads=8, head_dim=64, num_res_blocks=2, attn_scales=[1.0, 0.5, 0.25], dropout=0.1,
                  misc_dropout=0.4, temporal_attention=True, temporal_attn_times=1, use_checkpoint=True,
                  use_fps_condition=False, use_camera_condition=True, use_lgm_refine=False,
                  use_sim_mask=False, upper_len=128, default_fps=8)
I2V_KWARGS = dict(T2V_KWARGS, concat_dim=4)

# Reduced-width config used by the fast CPU tests (same topology, 1/5 the channels).
SMALL_KWARGS = dict(T2V_KWARGS, dim=64)   # context_dim stays 1024: decoder STs hard-code it (unet_t2v.py:237)
SMALL_I2V_KWARGS = dict(SMALL_KWARGS, concat_dim=4)
