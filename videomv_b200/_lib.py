"""Build + load the C-ABI shared library (include/videomv_b200.h) with ctypes.

The library is built IN-TREE (videomv_b200/lib/libvideomv_b200.so) by plain nvcc for sm_100a only.
There is no fallback: if the library is missing or fails to load, every op raises.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIBDIR = os.path.join(_HERE, "lib")
# VMV_LIB: load another build of the same ABI instead (same-box A/B runs of two source revisions; tools/ab_build.sh)
LIBPATH = os.environ.get("VMV_LIB") or os.path.join(LIBDIR, "libvideomv_b200.so")
SOURCES = ["gemm_tc.cu", "norm.cu", "attention.cu", "attention_tc.cu", "misc.cu", "peer.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isfile(cand) or cand == "nvcc"):
            return cand
    return "nvcc"


def _stale() -> bool:
    if not os.path.isfile(LIBPATH):
        return True
    t = os.path.getmtime(LIBPATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(_HERE, "..", "include", "videomv_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps if os.path.isfile(d))


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a and link the shared library. Returns its path."""
    if not force and not _stale():
        return LIBPATH
    os.makedirs(LIBDIR, exist_ok=True)
    objdir = os.path.join(LIBDIR, "obj")
    os.makedirs(objdir, exist_ok=True)
    nvcc = _nvcc()

    def compile_one(src):
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(" ".join(cmd))
        return obj

    with ThreadPoolExecutor(max_workers=4) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [nvcc, "-shared", "-o", LIBPATH, *objs, "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIBPATH


c_i32, c_i64, c_f32, c_vp = ctypes.c_int32, ctypes.c_int64, ctypes.c_float, ctypes.c_void_p


PEER_MAX = 8


class GemmScatter(ctypes.Structure):
    """Mirror of `vmv_gemm_scatter`."""
    _fields_ = [("world", c_i32), ("rank", c_i32), ("direction", c_i32), ("nowait", c_i32),
                ("B", c_i32), ("Fl", c_i32), ("HWl", c_i32), ("pad_", c_i32),
                ("dst", c_vp * PEER_MAX), ("flags", c_vp * PEER_MAX), ("epoch", c_vp), ("done", c_vp)]


class GemmParams(ctypes.Structure):
    """Mirror of `vmv_gemm_params` (include/videomv_b200.h)."""
    _fields_ = [
        ("mode", c_i32), ("M", c_i32), ("N", c_i32), ("K1", c_i32), ("K2", c_i32),
        ("A1", c_vp), ("lda1", c_i64), ("A2", c_vp), ("lda2", c_i64),
        ("W", c_vp), ("ldw", c_i64), ("D", c_vp), ("ldd", c_i64),
        ("B", c_i32), ("F", c_i32), ("H", c_i32), ("Wd", c_i32),
        ("bias", c_vp), ("rowbias", c_vp), ("ld_rowbias", c_i64), ("rows_per_group", c_i32),
        ("residual", c_vp), ("ldr", c_i64), ("act", c_i32), ("ln_stats", c_vp), ("ln_colsum", c_vp),
        ("ln_stats_src_n", c_i32), ("ln_stats_src_bn", c_i32), ("ln_eps", c_f32), ("rowstats_out", c_vp),
        ("scatter", ctypes.POINTER(GemmScatter)),
        ("block_n", c_i32), ("stages", c_i32), ("split_k", c_i32), ("variant", c_i32), ("w_static", c_i32),
        ("workspace", c_vp), ("workspace_bytes", c_i64),
    ]


class AttnParams(ctypes.Structure):
    """Mirror of `vmv_attn_params`."""
    _fields_ = [
        ("q", c_vp), ("k", c_vp), ("v", c_vp), ("o", c_vp),
        ("outer", c_i32), ("inner", c_i32), ("heads", c_i32), ("nq", c_i32), ("nk", c_i32),
        ("q_bs_outer", c_i64), ("q_bs_inner", c_i64), ("q_rs", c_i64),
        ("k_bs_outer", c_i64), ("k_bs_inner", c_i64), ("k_rs", c_i64),
        ("v_bs_outer", c_i64), ("v_bs_inner", c_i64), ("v_rs", c_i64),
        ("o_bs_outer", c_i64), ("o_bs_inner", c_i64), ("o_rs", c_i64),
        ("kv_group", c_i32), ("scale", c_f32), ("impl", c_i32),
    ]


PEER_MAX = 8


class PeerExchangeParams(ctypes.Structure):
    """Mirror of `vmv_peer_exchange_params`."""
    _fields_ = [("src", c_vp), ("dst", c_vp * PEER_MAX), ("flags", c_vp * PEER_MAX), ("epoch", c_vp), ("done", c_vp),
                ("world", c_i32), ("rank", c_i32), ("direction", c_i32), ("B", c_i32), ("Fl", c_i32), ("HWl", c_i32),
                ("C", c_i32), ("nowait", c_i32)]


class GnPeer(ctypes.Structure):
    """Mirror of `vmv_gn_peer`."""
    _fields_ = [("world", c_i32), ("rank", c_i32), ("slots", c_vp * PEER_MAX), ("flags", c_vp * PEER_MAX), ("epoch", c_vp),
                ("stat_rows", c_i64)]


class PeerAllreduceParams(ctypes.Structure):
    """Mirror of `vmv_peer_allreduce_params`."""
    _fields_ = [("data", c_vp), ("slots", c_vp * PEER_MAX), ("flags", c_vp * PEER_MAX), ("epoch", c_vp),
                ("world", c_i32), ("rank", c_i32), ("n", c_i32), ("nowait", c_i32)]


class PeerAllgatherParams(ctypes.Structure):
    """Mirror of `vmv_peer_allgather_params`."""
    _fields_ = [("src", c_vp), ("dst", c_vp * PEER_MAX), ("flags", c_vp * PEER_MAX), ("epoch", c_vp), ("done", c_vp),
                ("world", c_i32), ("rank", c_i32), ("nowait", c_i32), ("pad_", c_i32),
                ("nouter", c_i64), ("inner_bytes", c_i64), ("dst_offset_bytes", c_i64), ("dst_outer_stride_bytes", c_i64)]


# every symbol include/videomv_b200.h declares: (restype, argtypes)
SYMBOLS = {
    "vmv_last_error": (ctypes.c_char_p, []),
    "vmv_abi_version": (ctypes.c_int, []),
    "vmv_launch_count": (ctypes.c_longlong, []),
    "vmv_sizeof_gemm_params": (ctypes.c_int, []),
    "vmv_sizeof_gemm_scatter": (ctypes.c_int, []),
    "vmv_sizeof_attn_params": (ctypes.c_int, []),
    "vmv_sizeof_peer_exchange_params": (ctypes.c_int, []),
    "vmv_sizeof_peer_allreduce_params": (ctypes.c_int, []),
    "vmv_sizeof_gn_peer": (ctypes.c_int, []),
    "vmv_sizeof_peer_allgather_params": (ctypes.c_int, []),
    "vmv_gemm": (ctypes.c_int, [ctypes.POINTER(GemmParams), c_vp]),
    "vmv_gemm_workspace_bytes": (c_i64, [ctypes.POINTER(GemmParams)]),
    "vmv_gemm_block_n": (ctypes.c_int, [ctypes.POINTER(GemmParams)]),
    "vmv_gemm_epilogue_split": (ctypes.c_int, []),
    "vmv_groupnorm_scratch_bytes": (c_i64, [c_i32, c_i64, c_i32]),
    "vmv_groupnorm_stats": (ctypes.c_int, [c_vp, c_i64, c_i32, c_vp, c_i64, c_i32, c_i64, c_i32, c_vp, c_vp, c_vp, c_vp]),
    "vmv_groupnorm_apply": (ctypes.c_int, [c_vp, c_i64, c_i32, c_vp, c_i64, c_i32, c_i64, c_i32, c_vp, c_i64, c_vp, c_vp,
                                           c_f32, c_i32, c_vp, c_i64, c_vp]),
    "vmv_groupnorm_fused_fits_smem": (ctypes.c_int, [c_i32, c_i64, c_i32]),
    "vmv_groupnorm_fused": (ctypes.c_int, [c_vp, c_i64, c_i32, c_vp, c_i64, c_i32, c_i64, c_i32, c_vp, c_vp, c_vp, c_vp,
                                           c_f32, c_i32, c_vp, c_i64, c_vp]),
    "vmv_groupnorm_fused_peer": (ctypes.c_int, [c_vp, c_i64, c_i32, c_vp, c_i64, c_i32, c_i64, c_i32, c_vp, c_vp, c_vp, c_vp,
                                                c_f32, c_i32, c_vp, c_i64, ctypes.POINTER(GnPeer), c_vp]),
    "vmv_layernorm_stats": (ctypes.c_int, [c_vp, c_i64, c_i64, c_i32, c_f32, c_vp, c_vp]),
    "vmv_layernorm": (ctypes.c_int, [c_vp, c_i64, c_i64, c_i32, c_vp, c_vp, c_f32, c_vp, c_i64, c_vp]),
    "vmv_attention": (ctypes.c_int, [ctypes.POINTER(AttnParams), c_vp]),
    "vmv_upsample_nearest2x": (ctypes.c_int, [c_vp, c_i32, c_i32, c_i32, c_i32, c_vp, c_vp]),
    "vmv_im2col_3x3_s2": (ctypes.c_int, [c_vp, c_i32, c_i32, c_i32, c_i32, c_vp, c_vp]),
    "vmv_conv3x3_in": (ctypes.c_int, [c_vp, c_i32, c_vp, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp, c_vp, c_i32, c_vp, c_vp]),
    "vmv_conv3x3_out": (ctypes.c_int, [c_vp, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp, c_vp, c_i32, c_vp, c_vp]),
    "vmv_softmax_rows": (ctypes.c_int, [c_vp, c_i64, c_i64, c_i32, c_f32, c_vp, c_i64, c_vp]),
    "vmv_rows_to_ncfhw": (ctypes.c_int, [c_vp, c_i64, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp, c_vp]),
    "vmv_sinusoidal_embedding": (ctypes.c_int, [c_vp, c_i32, c_i32, c_vp, c_vp]),
    "vmv_embed_combine_silu": (ctypes.c_int, [c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_vp, c_vp]),
    "vmv_cfg_ddim_step": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_i64, c_vp, c_vp]),
    "vmv_peer_exchange": (ctypes.c_int, [ctypes.POINTER(PeerExchangeParams), c_vp]),
    "vmv_peer_allreduce_f64": (ctypes.c_int, [ctypes.POINTER(PeerAllreduceParams), c_vp]),
    "vmv_peer_allgather": (ctypes.c_int, [ctypes.POINTER(PeerAllgatherParams), c_vp]),
    "vmv_ipc_export": (ctypes.c_int, [c_vp, c_vp, ctypes.POINTER(c_i64)]),
    "vmv_ipc_import": (ctypes.c_int, [c_vp, c_i64, ctypes.POINTER(c_vp)]),
}

_LIB = None


def lib() -> ctypes.CDLL:
    """Load (never build implicitly on a GPU box without nvcc) the shared library; raise loudly if absent."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.isfile(LIBPATH):
        raise RuntimeError(
            f"videomv_b200: native library {LIBPATH} is missing. Run `python -c 'import __graft_entry__ as g; g.build()'`"
            " (nvcc, sm_100a). There is no CPU or PyTorch fallback for the UNet hot path.")
    L = ctypes.CDLL(LIBPATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(L, name)           # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    if L.vmv_abi_version() != 6:
        raise RuntimeError("videomv_b200: ABI version mismatch between _lib.py and the built library")
    if (L.vmv_sizeof_gemm_params() != ctypes.sizeof(GemmParams) or
            L.vmv_sizeof_gemm_scatter() != ctypes.sizeof(GemmScatter) or
            L.vmv_sizeof_attn_params() != ctypes.sizeof(AttnParams) or
            L.vmv_sizeof_peer_exchange_params() != ctypes.sizeof(PeerExchangeParams) or
            L.vmv_sizeof_peer_allreduce_params() != ctypes.sizeof(PeerAllreduceParams) or
            L.vmv_sizeof_gn_peer() != ctypes.sizeof(GnPeer) or
            L.vmv_sizeof_peer_allgather_params() != ctypes.sizeof(PeerAllgatherParams)):
        raise RuntimeError("videomv_b200: ctypes struct mirrors do not match the compiled vmv_*_params layouts")
    _LIB = L
    return L


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().vmv_last_error().decode(errors="replace")
        raise RuntimeError(f"videomv_b200 {what} failed (code {rc}): {msg}")


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
