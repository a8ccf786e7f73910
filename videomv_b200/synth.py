"""Deterministic synthetic weights / inputs for benchmarks and parity tests.

No checkpoints are reachable offline, and a freshly initialised reference UNet is
input-independent (every block tail is zero-initialised: SURVEY.md section 4, trap 1).
This recipe fills *every* parameter -- named exactly like the reference state_dict --
with seeded values that keep activations O(1) through the ~60 residual blocks.

Each tensor is generated from its own generator seeded by crc32(name) ^ seed, so the
result does not depend on iteration order and is identical on every machine with the
same torch build.
"""
from __future__ import annotations

import math
import zlib
from typing import Dict, Iterable, Mapping, Tuple

import torch


def _gen(name: str, seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(name.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    return g


def synth_tensor(name: str, shape: Tuple[int, ...], seed: int = 0, gain: float = 0.7) -> torch.Tensor:
    g = _gen(name, seed)
    shape = tuple(int(s) for s in shape)
    if len(shape) == 1:
        r = torch.randn(shape, generator=g, dtype=torch.float32)
        if name.endswith(".weight"):          # GroupNorm / LayerNorm scale
            return 1.0 + 0.1 * r
        return 0.05 * r                        # biases
    fan_in = 1
    for s in shape[1:]:
        fan_in *= s
    return torch.randn(shape, generator=g, dtype=torch.float32) * (gain / math.sqrt(fan_in))


def synth_state_dict(shapes: Mapping[str, Iterable[int]], seed: int = 0) -> Dict[str, torch.Tensor]:
    """shapes: {reference parameter name: shape}.  Returns fp32 CPU tensors."""
    return {k: synth_tensor(k, tuple(v), seed) for k, v in shapes.items()}


def fill_module_fast(module: torch.nn.Module, seed: int = 0, gain: float = 0.7) -> None:
    """Same distribution as synth_state_dict but generated on the module's own device (seconds instead of a minute
    for 1.4 B parameters). Values differ from the CPU recipe; use it where only the distribution matters (bench)."""
    with torch.no_grad():
        for name, p in module.state_dict().items():
            g = torch.Generator(device=p.device)
            g.manual_seed((zlib.crc32(name.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
            r = torch.randn(p.shape, generator=g, device=p.device, dtype=torch.float32)
            if p.dim() == 1:
                p.copy_(1.0 + 0.1 * r if name.endswith(".weight") else 0.05 * r)
            else:
                p.copy_(r * (gain / math.sqrt(p[0].numel())))


def synth_inputs(batch: int, frames: int, h: int, w: int, ctx_len: int = 77, ctx_dim: int = 1024,
                 in_dim: int = 4, seed: int = 1, t_value: int = 500):
    """Random (x, t, y, camera_data) shaped like the sampler's call (diffusion_ddim.py:149-155)."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    x = torch.randn(batch, in_dim, frames, h, w, generator=g)
    y = torch.randn(batch, ctx_len, ctx_dim, generator=g)
    cam = torch.randn(batch, frames, 16, generator=g)
    t = torch.full((batch,), t_value, dtype=torch.long)
    return x, t, y, cam


def orbit_cameras(frames: int = 24, elevation: float = 15.0, camera_distance: float = 2.0) -> torch.Tensor:
    """Orbit camera matrices [1,F,16] in the layout the T2V engine feeds the UNet.

    Same construction as utils/camera_utils.py:4-62 (`get_camera`: camera-to-world from
    elevation/azimuth on a sphere, look-at origin, y-up, then OpenGL->Blender axis change)
    followed by the row flips of tools/inferences/inference_text2video_entrance.py:186-191
    (negate row 1, swap rows 0 and 1).  Deterministic; no weights.
    """
    flip_yz = torch.tensor([[1, 0, 0, 0], [0, 0, -1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=torch.float64)
    cams = []
    for i in range(frames):
        az = math.radians(360.0 * i / frames)
        el = math.radians(elevation)
        pos = torch.tensor([camera_distance * math.cos(el) * math.sin(az),
                            camera_distance * math.sin(el),
                            camera_distance * math.cos(el) * math.cos(az)], dtype=torch.float64)
        fwd = -pos / pos.norm()
        up = torch.tensor([0.0, 1.0, 0.0], dtype=torch.float64)
        right = torch.linalg.cross(fwd, up)
        right = right / right.norm()
        new_up = torch.linalg.cross(right, fwd)
        new_up = new_up / new_up.norm()
        m = torch.eye(4, dtype=torch.float64)
        m[:3, 0], m[:3, 1], m[:3, 2], m[:3, 3] = right, new_up, -fwd, pos
        m = flip_yz @ m
        m[1, :] *= -1
        m = m[[1, 0, 2, 3], :]
        cams.append(m.reshape(16))
    return torch.stack(cams).float().unsqueeze(0)
