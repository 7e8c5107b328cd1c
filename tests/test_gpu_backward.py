"""GPU parity of the transform's backward (evfeat_backward.cu) against torch autograd
through the CPU oracle's transform, i.e. through the same torch.stft / mel-basis graph the reference trains through
(HiFiGAN: hfgl/model.py:581-590, 719-721 -- mel of the generated audio, log compression, L1 loss * 45)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

CONFIG = (22050, 1024, 1024, 256, 80, 0, 8000)
RTOL_GRAD = 2e-3   # of the largest |gradient| of the batch; the forward agrees to 1e-3 in the log domain


def _pair(spec_type, win=1024, hop=256):
    import everyvoice_b200 as ev
    from oracle import ev_oracle as O

    sr, n_fft, _, _, n_mels, f_min, f_max = CONFIG
    return (ev.get_spectral_transform(spec_type, n_fft, win, hop, sr, n_mels, f_min, f_max),
            O.get_spectral_transform(spec_type, n_fft, win, hop, sr, n_mels, f_min, f_max))


def _inputs(B, L, seed):
    from everyvoice_b200 import synth

    return np.stack([synth.speech_like(L, 22050, seed=seed + b) * np.float32(0.7) for b in range(B)])


@pytest.mark.parametrize("spec_type", ["mel", "mel-librosa", "linear"])
@pytest.mark.parametrize("B,L", [(3, 8192), (2, 5000 + 13), (1, 700)])
@pytest.mark.parametrize("win,hop", [(1024, 256), (800, 200)])
def test_linear_domain_backward_matches_autograd(cuda_device, spec_type, B, L, win, hop):
    """d/dx of sum(R * transform(x)) for a random R: exactly J^T R, no dependence on loss non-smoothness.  L = 8192
    is HiFiGAN's segment; 5013 is not a multiple of the hop; 700 has both reflect margins inside every frame."""
    tf, otf = _pair(spec_type, win, hop)
    x = _inputs(B, L, 700)
    xr = torch.tensor(x, requires_grad=True)
    y_ref = otf(xr)
    R = torch.from_numpy(np.random.default_rng(L).normal(size=tuple(y_ref.shape)).astype(np.float32))
    (y_ref * R).sum().backward()
    xg = torch.tensor(x, device=cuda_device, requires_grad=True)
    y = tf(xg)
    assert tuple(y.shape) == tuple(y_ref.shape) == (B, y_ref.shape[1], L // hop + 1)
    (y * R.to(cuda_device)).sum().backward()
    scale = float(xr.grad.abs().max())
    err = float((xg.grad.cpu() - xr.grad).abs().max())
    assert err <= RTOL_GRAD * scale, (err, scale)
    assert float((y.detach().cpu() - y_ref.detach()).abs().max()) <= 2e-3 * float(y_ref.detach().abs().max())


@pytest.mark.parametrize("spec_type", ["mel", "mel-librosa"])
def test_hifigan_mel_loss_backward(cuda_device, spec_type):
    """The reference's training use: log-mel of generated audio, L1 against a target mel, times 45."""
    import everyvoice_b200 as ev
    from oracle import ev_oracle as O

    tf, otf = _pair(spec_type)
    B, L = 4, 8192
    x = _inputs(B, L, 800)
    rng = np.random.default_rng(5)
    xr = torch.tensor(x, requires_grad=True)
    y_ref = O.dynamic_range_compression_torch(otf(xr))
    target = y_ref.detach() + torch.from_numpy(rng.normal(0, 1.0, size=tuple(y_ref.shape)).astype(np.float32))
    (torch.nn.functional.l1_loss(y_ref, target) * 45).backward()
    xg = torch.tensor(x, device=cuda_device, requires_grad=True)
    y = ev.dynamic_range_compression_torch(tf(xg))
    loss = torch.nn.functional.l1_loss(y, target.to(cuda_device)) * 45
    loss.backward()
    assert float((y.detach().cpu() - y_ref.detach()).abs().max()) <= 1e-3
    scale = float(xr.grad.abs().max())
    assert float((xg.grad.cpu() - xr.grad).abs().max()) <= 5e-3 * scale   # a few sign(y - target) may differ
    # the fused differentiable path (features(normalize=True)) is the same composition
    xg2 = torch.tensor(x, device=cuda_device, requires_grad=True)
    y2 = tf.features(xg2, normalize=True, keep_last=True)
    (torch.nn.functional.l1_loss(y2, target.to(cuda_device)) * 45).backward()
    assert torch.equal(y2, y) and torch.equal(xg2.grad, xg.grad)


def test_log_compression_backward_and_clamp_mask(cuda_device):
    import everyvoice_b200 as ev

    v = torch.tensor([2.0, 1e-5, 9.9e-6, 0.0, 3e-3, 1e-7], device=cuda_device, requires_grad=True)
    ev.dynamic_range_compression_torch(v).sum().backward()
    vr = v.detach().cpu().clone().requires_grad_(True)
    torch.log(torch.clamp(vr, min=1e-5) * 1).sum().backward()
    assert torch.allclose(v.grad.cpu(), vr.grad, rtol=1e-6, atol=0)
    assert v.grad[2] == 0 and v.grad[3] == 0 and v.grad[1] > 0


def test_backward_is_deterministic_and_rejects_unsupported(cuda_device):
    import everyvoice_b200 as ev

    tf, _ = _pair("mel")
    x = torch.tensor(_inputs(2, 4096, 900), device=cuda_device)
    grads = []
    for _ in range(2):
        xg = x.clone().requires_grad_(True)
        tf(xg).square().sum().backward()
        grads.append(xg.grad.clone())
    assert torch.equal(grads[0], grads[1])
    raw, _ = _pair("raw")
    with pytest.raises(NotImplementedError):
        raw(x.clone().requires_grad_(True))


@pytest.mark.parametrize("spec_type", ["mel", "mel-librosa", "linear"])
@pytest.mark.parametrize("win,hop,f_max", [(2048, 512, 8000), (2048, 512, 22050), (1200, 300, 8000)])
def test_n_fft_2048_backward_matches_autograd(cuda_device, spec_type, win, hop, f_max):
    """44.1 kHz vocoder configuration (BASELINE configs[2]): one frame per FFT job, half-size inverse; also a window
    shorter than n_fft with a hop that is not a power of two."""
    import everyvoice_b200 as ev
    from everyvoice_b200 import synth
    from oracle import ev_oracle as O

    args = (spec_type, 2048, win, hop, 44100, 128, 0, f_max)
    tf, otf = ev.get_spectral_transform(*args), O.get_spectral_transform(*args)
    B, L = 2, 16384 + 77
    x = np.stack([synth.speech_like(L, 44100, seed=950 + b) * np.float32(0.6) for b in range(B)])
    xr = torch.tensor(x, requires_grad=True)
    y_ref = otf(xr)
    R = torch.from_numpy(np.random.default_rng(hop).normal(size=tuple(y_ref.shape)).astype(np.float32))
    (y_ref * R).sum().backward()
    xg = torch.tensor(x, device=cuda_device, requires_grad=True)
    y = tf(xg)
    assert tuple(y.shape) == tuple(y_ref.shape)
    (y * R.to(cuda_device)).sum().backward()
    scale = float(xr.grad.abs().max())
    assert float((xg.grad.cpu() - xr.grad).abs().max()) <= RTOL_GRAD * scale
