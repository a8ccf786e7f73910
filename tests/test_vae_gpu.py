"""VAE decoder on the GPU: videomv_b200.vae.AutoencoderKL.decode through the C ABI vs the
reference's golden outputs (oracle/gen_golden.py `vae`: the UNMODIFIED tools/modules/autoencoder.py on the same weights)."""
import pytest
import torch

from tests.test_oracle_cpu import load_case

pytestmark = pytest.mark.gpu


def _build(meta):
    from videomv_b200 import synth, vae
    model = vae.AutoencoderKL(**meta["kwargs"])
    sd = synth.synth_state_dict(meta["shapes"], seed=meta["seed_w"])
    model.load_state_dict(sd, strict=True)
    return model.cuda().eval(), sd


@pytest.mark.parametrize("case", ["vae_small", "vae_256"])
def test_vae_decode_matches_reference_golden(case):
    meta, d, _ = load_case(case)
    model, _ = _build(meta)
    out = model.decode(d["z"].cuda())
    ref = d["ref"]
    assert out.shape == ref.shape and out.dtype == torch.float32 and out.is_cuda
    rel = ((out.cpu() - ref).norm() / ref.norm()).item()
    mx = (out.cpu() - ref).abs().max().item()
    print(f"[vae] {case}: rel_l2={rel:.3e} max_abs={mx:.3e} max|ref|={ref.abs().max().item():.3f}")
    # ~30 convolutions / GroupNorms with fp16 activation storage (same bound as the UNet end-to-end tests)
    assert rel < 4e-3 and mx < 3e-2
    assert torch.equal(model.decode(d["z"].cuda()), out), "decode must be bit-reproducible"


def test_vae_decode_chunk_of_frames_like_the_reference_engine():
    """inference_text2video_entrance.py:279-290 decodes the 24 frames in chunks of `decoder_bs`: a chunk must decode each
    frame exactly as if it were alone (per-image GroupNorm / attention)."""
    meta, d, _ = load_case("vae_small")
    model, _ = _build(meta)
    g = torch.Generator().manual_seed(5)
    z = torch.randn(4, 4, 8, 8, generator=g).cuda() * 5
    full = model.decode(z)
    for i in range(4):
        one = model.decode(z[i:i + 1])
        assert ((full[i:i + 1] - one).norm() / one.norm()).item() < 3e-3


def test_softmax_rows():
    from videomv_b200 import ops
    x = (torch.randn(300, 1024, device="cuda") * 8).half()
    out = ops.softmax_rows(x, 512 ** -0.5)
    ref = torch.softmax(x.float() * 512 ** -0.5, dim=1)
    assert torch.allclose(out.float(), ref, rtol=2e-3, atol=1e-6)
    assert torch.allclose(out.float().sum(1), torch.ones(300, device="cuda"), atol=2e-3)
