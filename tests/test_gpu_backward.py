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
    # (the log sits in the kernel's epilogue -- one MUFU.LG2 -- and its derivative exp(-y) in the backward's gradient
    # load: the same numbers as the two-step composition up to float32 round-off)
    assert float((y2 - y).detach().abs().max()) <= 2e-6 * float(y.detach().abs().max())
    assert float((xg2.grad - xg.grad).abs().max()) <= 1e-5 * float(xg.grad.abs().max())


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


# ------------------------------------------------------------------------------------------------------------------
# every other transform size (any-size kernels), and the warp kernel cross-checked against them
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("config", ["S512", "W512", "S256", "W256", "R3", "N400", "O1001", "OddHop", "BigHop", "Gap"])
@pytest.mark.parametrize("spec_type", ["mel", "mel-librosa", "linear"])
def test_any_size_backward_matches_autograd(cuda_device, config, spec_type):
    """d/dx of sum(R * transform(x)) for n_fft 512 / 3072 / 400 / 1001 (odd: 7 * 11 * 13, float64 direct-DFT stages),
    an odd hop at 2048, hop = n_fft and hop > n_fft (samples between frames get no gradient)."""
    import everyvoice_b200 as ev
    from conftest import CONFIGS
    from everyvoice_b200 import synth
    from oracle import ev_oracle as O

    sr, n_fft, win, hop, n_mels, f_min, f_max = CONFIGS[config]
    tf = ev.get_spectral_transform(spec_type, n_fft, win, hop, sr, n_mels, f_min, f_max)
    otf = O.get_spectral_transform(spec_type, n_fft, win, hop, sr, n_mels, f_min, f_max)
    B, L = 2, max(4 * n_fft, 9 * hop) + 29
    x = np.stack([synth.speech_like(L, sr, seed=1100 + b) * np.float32(0.6) for b in range(B)])
    xr = torch.tensor(x, requires_grad=True)
    y_ref = otf(xr)
    R = torch.from_numpy(np.random.default_rng(n_fft + hop).normal(size=tuple(y_ref.shape)).astype(np.float32))
    (y_ref * R).sum().backward()
    xg = torch.tensor(x, device=cuda_device, requires_grad=True)
    y = tf(xg)
    assert tuple(y.shape) == tuple(y_ref.shape)
    (y * R.to(cuda_device)).sum().backward()
    scale = float(xr.grad.abs().max())
    assert float((xg.grad.cpu() - xr.grad).abs().max()) <= RTOL_GRAD * scale


@pytest.mark.parametrize("win,hop,n_mels", [(512, 128, 80), (240, 50, 128), (400, 160, 80), (512, 512, 40)])
@pytest.mark.parametrize("spec_type", ["mel", "linear"])
def test_packed_job_backward_agrees_with_the_any_size_backward(cuda_device, spec_type, win, hop, n_mels):
    """n_fft 512: the warp backward runs two packed jobs per warp (register re-layout between the recomputed forward
    and the inverse transform, csrc/evfeat_backward.cu) -- against the any-size kernels on batches of several tiles,
    with the tile-sum overlap-add (4 to 11 rounds) and without overlap (hop = n_fft), two-step and fused-log paths."""
    import everyvoice_b200 as ev

    sr = 24000
    B, L = 3, 512 * 41 + 77
    g = torch.Generator(device="cpu").manual_seed(win + hop)
    x = (torch.rand(B, L, generator=g) * 1.6 - 0.8).to(cuda_device)
    fast = ev.SpectralTransform(spec_type, 512, win, hop, sr, n_mels, 0, 8000)
    slow = ev.SpectralTransform(spec_type, 512, win, hop, sr, n_mels, 0, 8000, fft_path="generic")
    grads = []
    for tf in (fast, slow):
        for fused in (False, True):
            xg = x.clone().requires_grad_(True)
            y = tf.features(xg, normalize=True, keep_last=True) if fused else ev.dynamic_range_compression_torch(tf(xg))
            R = torch.sin(torch.arange(y.numel(), device=cuda_device, dtype=torch.float32) * 0.37).view_as(y)
            (y * R).sum().backward()
            grads.append(xg.grad)
    scale = float(grads[2].abs().max())
    diffs = [float((gq - grads[2]).abs().max()) / scale for gq in (grads[0], grads[1], grads[3])]
    # [warp two-step, warp fused, any-size fused] against any-size two-step: two float32 FFTs through 1 / x
    assert max(diffs) <= 2e-4 and diffs[2] <= 2e-5, diffs


@pytest.mark.parametrize("n_fft,win,hop,sr,n_mels", [(1024, 1024, 256, 22050, 80), (2048, 1200, 300, 16000, 80)])
def test_backward_of_both_fft_kernels_agree(cuda_device, n_fft, win, hop, sr, n_mels):
    import everyvoice_b200 as ev
    from everyvoice_b200 import synth

    x = torch.from_numpy(np.stack([synth.speech_like(3 * n_fft + 11, sr, seed=1200 + b) for b in range(3)])).to(cuda_device)
    grads = []
    for path in ("auto", "generic"):
        tf = ev.SpectralTransform("mel", n_fft, win, hop, sr, n_mels, 0, 8000, fft_path=path)
        xg = x.clone().requires_grad_(True)
        tf.features(xg, normalize=True, keep_last=True).square().sum().backward()
        grads.append(xg.grad)
    assert float((grads[0] - grads[1]).abs().max()) <= 1e-3 * float(grads[0].abs().max())


def test_styletts2_mel_parameterisation_forward_and_backward(cuda_device):
    """StyleTTS2's mels (styletts2/utils.py:12-21 ``make_mel_transform``: T.MelSpectrogram(n_mels=80, n_fft=2048,
    win_length=1200, hop_length=300) -- torchaudio defaults: 16 kHz, htk scale, NO norm, f_max = sr / 2 -- followed by
    ``(log(1e-5 + mel) - (-4)) / 4``, utils.py:37 / losses.py:61-62) and its multi-resolution STFT loss transforms
    (losses.py:42-48, 24 kHz, 128 mels: 1024/120/600, 2048/240/1200, 512/50/240).  The mel basis is checked against
    torchaudio's own; forward and backward against torch autograd through the same graph built from torch.stft."""
    import everyvoice_b200 as ev
    from everyvoice_b200 import filterbanks, synth
    from oracle import ev_oracle as O

    for sr, n_fft, win, hop, n_mels in ((16000, 2048, 1200, 300, 80), (24000, 1024, 600, 120, 128),
                                        (24000, 2048, 1200, 240, 128), (24000, 512, 240, 50, 128)):
        fb = filterbanks.melscale_fbanks(n_fft // 2 + 1, 0.0, float(sr // 2), n_mels, sr, None, "htk")
        ref_fb = O.torchaudio_melscale_fbanks(n_fft // 2 + 1, 0.0, float(sr // 2), n_mels, sr, None, "htk")
        assert torch.equal(fb, ref_fb)
        try:
            import torchaudio

            assert torch.equal(fb, torchaudio.functional.melscale_fbanks(n_fft // 2 + 1, 0.0, float(sr // 2), n_mels, sr))
        except ImportError:
            pass
        tf = ev.SpectralTransform("mel", n_fft, win, hop, sr, n_mels, 0, None, norm=None, mel_scale="htk")
        L = 6 * n_fft + 17
        x = np.stack([synth.speech_like(L, sr, seed=1300 + b) * np.float32(0.5) for b in range(2)])

        def graph(xt, dev):
            if dev == "cpu":
                spec = O._spectrogram(xt, n_fft, win, hop, 2.0)
                mel = torch.matmul(spec.transpose(-1, -2), ref_fb).transpose(-1, -2)
            else:
                mel = tf(xt)
            return (torch.log(1e-5 + mel) - (-4.0)) / 4.0

        xr = torch.tensor(x, requires_grad=True)
        y_ref = graph(xr, "cpu")
        R = torch.from_numpy(np.random.default_rng(hop).normal(size=tuple(y_ref.shape)).astype(np.float32))
        (y_ref * R).sum().backward()
        xg = torch.tensor(x, device=cuda_device, requires_grad=True)
        y = graph(xg, "cuda")
        (y * R.to(cuda_device)).sum().backward()
        assert tuple(y.shape) == tuple(y_ref.shape)
        assert float((y.detach().cpu() - y_ref.detach()).abs().max()) <= 1e-3, (n_fft, hop)
        scale = float(xr.grad.abs().max())
        assert float((xg.grad.cpu() - xr.grad).abs().max()) <= 5e-3 * scale, (n_fft, hop)


def test_hifigan_training_mel_drops_the_first_frame(cuda_device):
    """hfgl/model.py:719-721: dynamic_range_compression_torch(spectral_transform(generated_wav).squeeze(1)[:, :, 1:]) --
    the same expression on the B200 transform, and SpectralTransform.training_mel as its fused form; gradients against
    autograd through the oracle, including a sampling-rate change of 2 (n_fft 2048 / hop 512 at the output rate)."""
    import everyvoice_b200 as ev
    from oracle import ev_oracle as O

    for k in (1, 2):
        args = ("mel", 1024 * k, 1024 * k, 256 * k, 22050 * k, 80, 0, 8000)
        tf, otf = ev.get_spectral_transform(*args), O.get_spectral_transform(*args)
        x = _inputs(4, 8192 * k, 1400)[:, None, :]                               # generated_wav [B, 1, L]
        xr = torch.tensor(x, requires_grad=True)
        y_ref = O.dynamic_range_compression_torch(otf(xr).squeeze(1)[:, :, 1:])
        target = y_ref.detach() * 0.9 - 0.3
        ((y_ref - target) ** 2).mean().backward()                                 # a smooth loss: no sign flips
        xg = torch.tensor(x, device=cuda_device, requires_grad=True)
        y = ev.dynamic_range_compression_torch(tf(xg).squeeze(1)[:, :, 1:])
        assert tuple(y.shape) == tuple(y_ref.shape) == (4, 80, 32)
        ((y - target.to(cuda_device)) ** 2).mean().backward()
        assert float((y.detach().cpu() - y_ref.detach()).abs().max()) <= 1e-3
        scale = float(xr.grad.abs().max())
        assert float((xg.grad.cpu() - xr.grad).abs().max()) <= RTOL_GRAD * scale
        xg2 = torch.tensor(x, device=cuda_device, requires_grad=True)
        y2 = tf.training_mel(xg2)                                                 # log fused forward and backward
        assert float((y2 - y).detach().abs().max()) <= 2e-6 * float(y.detach().abs().max())
        ((y2 - target.to(cuda_device)) ** 2).mean().backward()
        assert float((xg2.grad - xg.grad).abs().max()) <= 1e-5 * float(xg.grad.abs().max())
        assert float((xg2.grad.cpu() - xr.grad).abs().max()) <= RTOL_GRAD * scale


def test_full_size_backward_with_a_smooth_loss(cuda_device):
    """1 000 x 5 s at 22.05 kHz (profiles/r01n_backward_bench.json's batch_1000x5s, where an L1 loss against a random
    target showed 1.26 % of the largest gradient -- sign(y - target) flips where the two forwards differ by 1e-5).
    With a SMOOTH loss (mean squared log-mel difference) the same batch agrees with torch autograd on the same GPU
    (torch.stft / matmul graph built from the package's own window and filterbank) to 2e-3 of the largest gradient,
    and the L1 discrepancy is reproduced and attributed: it vanishes on the elements whose sign agrees."""
    import everyvoice_b200 as ev

    sr, n_fft, hop, n_mels = 22050, 1024, 256, 80
    B, L = 1000, 5 * sr // hop * hop
    g = torch.Generator(device=cuda_device)
    g.manual_seed(77)
    x = (torch.rand((B, L), device=cuda_device, generator=g) * 1.9 - 0.95) * torch.rand((B, 1), device=cuda_device, generator=g)
    tf = ev.get_spectral_transform("mel", n_fft, n_fft, hop, sr, n_mels, 0, 8000).to(cuda_device)
    win, fb = tf.window.to(cuda_device), tf.mel_fb.to(cuda_device)

    def torch_logmel(xt):
        spec = torch.stft(xt, n_fft, hop, n_fft, win, center=True, pad_mode="reflect", return_complex=True).abs().pow(2)
        return torch.log(torch.clamp(torch.matmul(spec.transpose(-1, -2), fb).transpose(-1, -2), min=1e-5))

    xr = x.clone().requires_grad_(True)
    y_ref = torch_logmel(xr)
    target = (y_ref.detach() * 0.8 + 0.1 * torch.randn(y_ref.shape, device=cuda_device, generator=g))
    ((y_ref - target) ** 2).mean().backward()
    xg = x.clone().requires_grad_(True)
    y = tf.features(xg, normalize=True, keep_last=True)
    ((y - target) ** 2).mean().backward()
    assert float((y.detach() - y_ref.detach()).abs().max()) <= 1e-3
    scale = float(xr.grad.abs().max())
    assert float((xg.grad - xr.grad).abs().max()) <= RTOL_GRAD * scale
    # the L1 case: differences come from sign(y - target) where |y - target| is below the forward difference
    xr1, xg1 = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    y1r, y1 = torch_logmel(xr1), tf.features(xg1, normalize=True, keep_last=True)
    flips = (torch.sign(y1r.detach() - target) != torch.sign(y1.detach() - target))
    (torch.nn.functional.l1_loss(y1r, target) * 45).backward()
    (torch.nn.functional.l1_loss(y1, target) * 45).backward()
    rel_l1 = float((xg1.grad - xr1.grad).abs().max()) / float(xr1.grad.abs().max())
    # same upstream gradient for both (the torch graph's sign pattern): the discrepancy disappears
    up = (torch.sign(y1r.detach() - target) * (45.0 / y1r.numel()))
    xr2, xg2 = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    torch_logmel(xr2).backward(up)
    tf.features(xg2, normalize=True, keep_last=True).backward(up)
    rel_same_sign = float((xg2.grad - xr2.grad).abs().max()) / float(xr2.grad.abs().max())
    assert rel_same_sign <= RTOL_GRAD, rel_same_sign
    assert int(flips.sum()) > 0 or rel_l1 <= RTOL_GRAD   # any excess of rel_l1 over rel_same_sign is the flipped signs


def test_backward_ex_argument_checks_and_layouts(cuda_device):
    """evf_features_backward_ex through the C ABI: both gradient layouts give the same gradient, with and without the
    folded log; a plan with apply_log needs the forward's log output; unknown layouts and the plain entry point on a
    log plan are refused with a status, nothing is launched."""
    import ctypes as C

    import everyvoice_b200 as ev
    from everyvoice_b200 import _lib
    from everyvoice_b200.heavy import _ptr, _stream_ptr

    tf = ev.get_spectral_transform("mel", 1024, 1024, 256, 22050, 80, 0, 8000).to(cuda_device)
    B, L = 3, 6000
    x = (torch.rand(B, L, device=cuda_device) * 1.6 - 0.8).contiguous()
    lib = _lib.load()
    outs = {}
    for apply_log in (False, True):
        batch = tf.uniform_batch(B, L, cuda_device, apply_log=apply_log, keep_last=True)
        plan = batch.plan
        spec, _ = tf.run(batch, x.reshape(-1), want_energy=False)          # [B * T, 80] frame-major
        T = spec.shape[0] // B
        g_fm = torch.cos(torch.arange(spec.numel(), device=cuda_device, dtype=torch.float32) * 0.11).view_as(spec)
        g_bm = g_fm.view(B, T, 80).transpose(1, 2).contiguous()            # per utterance [F][T]
        n = int(lib.evf_features_backward_scratch_floats(plan.handle, batch.handle))
        scratch = torch.empty(n, dtype=torch.float32, device=cuda_device)
        res = []
        for g, layout in ((g_fm, _lib.GRAD_FRAME_MAJOR), (g_bm, _lib.GRAD_BIN_MAJOR)):
            gx = torch.empty_like(x)
            rc = lib.evf_features_backward_ex(plan.handle, batch.handle, _ptr(x), _ptr(g), layout,
                                              _ptr(spec if apply_log else None), _ptr(scratch), _ptr(gx),
                                              _stream_ptr(cuda_device))
            assert rc == 0, lib.evf_last_error()
            res.append(gx)
        assert torch.equal(res[0], res[1])                                  # the layout only changes the addressing
        outs[apply_log] = res[0]
        gx = torch.empty_like(x)
        if apply_log:
            assert lib.evf_features_backward_ex(plan.handle, batch.handle, _ptr(x), _ptr(g_fm), 0, C.c_void_p(0),
                                                _ptr(scratch), _ptr(gx), _stream_ptr(cuda_device)) == _lib.EVF_ERR_INVALID_ARGUMENT
            assert lib.evf_features_backward(plan.handle, batch.handle, _ptr(x), _ptr(g_fm), _ptr(scratch), _ptr(gx),
                                             _stream_ptr(cuda_device)) == _lib.EVF_ERR_UNSUPPORTED
        assert lib.evf_features_backward_ex(plan.handle, batch.handle, _ptr(x), _ptr(g_fm), 7, C.c_void_p(0),
                                            _ptr(scratch), _ptr(gx), _stream_ptr(cuda_device)) == _lib.EVF_ERR_INVALID_ARGUMENT
    # d loss / d log-mel folded in == d loss / d mel of the same upstream gradient divided by the mel where it passes
    torch.cuda.synchronize()
    assert torch.isfinite(outs[True]).all() and float(outs[True].abs().max()) > 0
