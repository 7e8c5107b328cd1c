"""CPU restatement of the pitch tracker behind ``Preprocessor.extract_pitch``
(everyvoice/preprocessor/preprocessor.py:244-285): ``pyworld.dio(x, sr, frame_period=hop / sr * 1000, speed=4)``
followed by ``pyworld.stonemask(x, f0, t, sr)``.

TEST INFRASTRUCTURE -- and **PARITY UNPINNED**: the arithmetic lives in a third-party dependency that is absent from
/root/reference AND from this image: pyworld-prebuilt 0.3.4.4 (pyproject.toml: ``pyworld-prebuilt``), a wrapper of
M. Morise's WORLD vocoder (C++, ``src/dio.cpp``, ``src/stonemask.cpp``, ``src/matlabfunctions.cpp``, ``src/common.cpp``).
Neither the wheel nor the sources are available offline, so this file restates WORLD's published algorithm
(Morise, Kawahara, Katayose: "Fast and reliable F0 estimation method based on the period extraction of vocal fold
vibration of singing voice and speech", AES 35th Int. Conf., 2009; WORLD 0.2.x sources as published) function by
function, in float64 like WORLD, with WORLD's own formulation (FFT-domain filtering) -- the CUDA kernels use the
equivalent time-domain form, so the two are independent derivations of the same numbers.  It could not be checked
against pyworld itself.  One cross-check is possible and is made in tests/test_pitch_oracle.py: WORLD's decimation
filter is a hard-coded table that equals MATLAB's / scipy's ``cheby1(3, 0.05, 0.8 / r)``.

Defaults of pyworld.dio: f0_floor 71, f0_ceil 800, channels_in_octave 2, allowed_range 0.1.
"""

from __future__ import annotations

import math

import numpy as np

K_CUTOFF = 50.0            # world::kCutOff (Hz): low-cut filter of DIO
K_MAXIMUM_VALUE = 100000.0  # world::kMaximumValue: score of a rejected candidate
K_SAFE_GUARD = 1e-12        # world::kMySafeGuardMinimum
K_FLOOR_F0_STONEMASK = 40.0

# FilterForDecimate (matlabfunctions.cpp): 3rd-order Chebyshev type I low-pass, 0.05 dB ripple, cut-off 0.8 / r,
# stored as a = {-a1, -a2, -a3}, b = {b0, b1} (b = b0 * [1, 3, 3, 1]).  WORLD hard-codes the table (from MATLAB's
# cheby1); the values here are scipy.signal.cheby1(3, 0.05, 0.8 / r), the same design (the one entry of WORLD's table
# recalled with confidence, r = 11: a = {2.450743295230728, -2.06794904601978, 0.59574774438332101}, b =
# {0.0026822508007163792, 0.0080467524021491377}, agrees to 1e-14).  The reference uses r = 4 (speed=4).
DECIMATE_COEFFS = {
    2: ((0.04115673456775716, -0.4259911245918959, 0.04103721547996115), (0.1679746468180222, 0.5039239404540666)),
    3: ((0.9503937898323742, -0.674291467415268, 0.15412211621346472), (0.07122194517117862, 0.21366583551353585)),
    4: ((1.4499664446880223, -0.9894349708095054, 0.245782523406902), (0.03671075033932264, 0.11013225101796792)),
    5: ((1.761093965428056, -1.255491484385977, 0.32371865077882145), (0.02133485852238745, 0.06400457556716235)),
    6: ((1.971535274951214, -1.4686795689225343, 0.38939084349657005), (0.013469181309343806, 0.04040754392803142)),
    7: ((2.12252390195347, -1.6395144861046296, 0.44469707800587344), (0.009036688268160781, 0.027110064804482345)),
    8: ((2.2357462340187593, -1.7780899984041356, 0.491525553659687), (0.006352276340711179, 0.01905682902213354)),
    9: ((2.323600349175958, -1.89215456174636, 0.5314892813372907), (0.004633116404138924, 0.013899349212416773)),
    10: ((2.3936475118069382, -1.9873904075111852, 0.5658879979027052), (0.0034818622251927374, 0.010445586675578211)),
    11: ((2.450743295230728, -2.0679490460197805, 0.5957477443833211), (0.002682250800716404, 0.008046752402149212)),
    12: ((2.4981398605924205, -2.1368928194784025, 0.6218751381622148), (0.002109727590470877, 0.006329182771412631)),
}


def matlab_round(x: float) -> int:
    return int(x + 0.5) if x > 0 else int(x - 0.5)


def suitable_fft_size(sample: int) -> int:
    return int(2.0 ** (int(math.log(sample) / math.log(2.0)) + 1.0))


def filter_for_decimate(x: np.ndarray, r: int) -> np.ndarray:
    a, b = DECIMATE_COEFFS[r]
    y = np.empty_like(x)
    w0 = w1 = w2 = 0.0
    for i in range(len(x)):
        wt = x[i] + a[0] * w0 + a[1] * w1 + a[2] * w2
        y[i] = b[0] * wt + b[1] * w0 + b[1] * w1 + b[0] * w2
        w2, w1, w0 = w1, w0, wt
    return y


def decimate(x: np.ndarray, r: int) -> np.ndarray:
    """matlabfunctions.cpp ``decimate``: reflected 9-sample margins, the IIR forwards and backwards, every r-th sample."""
    n_fact = 9
    n = len(x)
    tmp1 = np.empty(n + 2 * n_fact)
    for i in range(n_fact):
        tmp1[i] = 2 * x[0] - x[n_fact - i]
    tmp1[n_fact : n_fact + n] = x
    for i in range(n_fact + n, 2 * n_fact + n):
        tmp1[i] = 2 * x[n - 1] - x[n - 2 - (i - (n_fact + n))]
    tmp2 = filter_for_decimate(tmp1, r)
    tmp1 = tmp2[::-1].copy()
    tmp2 = filter_for_decimate(tmp1, r)
    tmp1 = tmp2[::-1].copy()
    nout = (n - 1) // r + 1
    nbeg = r - r * nout + n
    return tmp1[np.arange(nbeg, n + n_fact, r) + n_fact - 1]


def nuttall_window(n: int) -> np.ndarray:
    t = np.arange(n) / (n - 1.0)
    return 0.355768 - 0.487396 * np.cos(2 * np.pi * t) + 0.144232 * np.cos(4 * np.pi * t) - 0.012604 * np.cos(6 * np.pi * t)


def low_cut_filter(n: int) -> np.ndarray:
    """DesignLowCutFilter: minus a normalised Hann window of n taps, plus a unit impulse at its centre."""
    i = np.arange(1, n + 1)
    h = 0.5 - 0.5 * np.cos(i * 2.0 * np.pi / (n + 1))
    h = -h / h.sum()
    h[(n - 1) // 2] += 1.0
    return h


def spectrum_for_estimation(x, fs, decimation_ratio, fft_size, y_length):
    y = np.zeros(fft_size)
    if decimation_ratio != 1:
        d = decimate(x, decimation_ratio)
        y[: len(d)] = d
    else:
        y[: len(x)] = x
    mean_y = y[:y_length].sum() / y_length
    y[:y_length] -= mean_y
    y[y_length:] = 0.0
    spec = np.fft.rfft(y)
    cutoff_in_sample = matlab_round(fs / decimation_ratio / K_CUTOFF)
    n = cutoff_in_sample * 2 + 1
    lc = np.zeros(fft_size)
    h = low_cut_filter(n)
    half = (n - 1) // 2
    lc[: half + 1] = h[half:]          # zero-phase arrangement: centre tap at index 0 ...
    lc[fft_size - half :] = h[:half]   # ... the first half wrapped around
    return spec * np.fft.rfft(lc)


def zero_crossing_engine(sig: np.ndarray, fs: float):
    n = len(sig)
    neg_going = np.nonzero((sig[:-1] > 0.0) & (sig[1:] <= 0.0))[0] + 1
    if len(neg_going) < 2:
        return np.zeros(0), np.zeros(0)
    fine = neg_going - sig[neg_going - 1] / (sig[neg_going] - sig[neg_going - 1])
    intervals = fs / (fine[1:] - fine[:-1])
    locations = (fine[:-1] + fine[1:]) / 2.0 / fs
    return locations, intervals


def interp1(x, y, xi):
    """matlabfunctions.cpp ``interp1`` (linear, with linear EXTRApolation from the first / last segment; histc bins)."""
    k = np.searchsorted(x, xi, side="right")           # smallest index with x[k] > xi
    k = np.clip(k, 1, len(x) - 1)
    s = (xi - x[k - 1]) / (x[k] - x[k - 1])
    return y[k - 1] + s * (y[k] - y[k - 1])


def dio(x, fs, frame_period=5.0, speed=1, f0_floor=71.0, f0_ceil=800.0, channels_in_octave=2.0, allowed_range=0.1):
    """``pyworld.dio``: returns ``(f0[f0_length], temporal_positions[f0_length])`` float64."""
    x = np.asarray(x, dtype=np.float64)
    x_length = len(x)
    number_of_bands = 1 + int(math.log(f0_ceil / f0_floor) / math.log(2.0) * channels_in_octave)
    boundary_f0_list = [f0_floor * 2.0 ** ((i + 1) / channels_in_octave) for i in range(number_of_bands)]
    decimation_ratio = max(min(int(speed), 12), 1)
    y_length = 1 + x_length // decimation_ratio
    actual_fs = fs / decimation_ratio
    fft_size = suitable_fft_size(y_length + matlab_round(actual_fs / K_CUTOFF) * 2 + 1
                                 + 4 * int(1.0 + actual_fs / boundary_f0_list[0] / 2.0))
    y_spectrum = spectrum_for_estimation(x, fs, decimation_ratio, fft_size, y_length)
    f0_length = int(1000.0 * x_length / fs / frame_period) + 1
    temporal_positions = np.arange(f0_length) * frame_period / 1000.0
    cand = np.zeros((number_of_bands, f0_length))
    score = np.zeros((number_of_bands, f0_length))
    for b, boundary_f0 in enumerate(boundary_f0_list):
        half = matlab_round(actual_fs / boundary_f0 / 2.0)
        lp = np.zeros(fft_size)
        lp[: half * 4] = nuttall_window(half * 4)
        filtered = np.fft.irfft(y_spectrum * np.fft.rfft(lp), fft_size)[half * 2 : half * 2 + y_length]
        events = [zero_crossing_engine(filtered, actual_fs), zero_crossing_engine(-filtered, actual_fs)]
        d = filtered[:-1] - filtered[1:]
        events += [zero_crossing_engine(d, actual_fs), zero_crossing_engine(-d, actual_fs)]
        if any(len(loc) - 2 <= 0 for loc, _ in events):   # CheckEvent(n - 2) for all four kinds
            cand[b], score[b] = 0.0, K_MAXIMUM_VALUE
            continue
        f = np.stack([interp1(loc, itv, temporal_positions) for loc, itv in events])
        c = (f[0] + f[1] + f[2] + f[3]) / 4.0
        s = np.sqrt(((f - c) ** 2).sum(axis=0) / 3.0)
        bad = (c > boundary_f0) | (c < boundary_f0 / 2.0) | (c > f0_ceil) | (c < f0_floor)
        c[bad], s[bad] = 0.0, K_MAXIMUM_VALUE
        cand[b], score[b] = c, s
    best = cand[np.argmin(score, axis=0), np.arange(f0_length)]   # first minimum over the bands, like the strict '>'
    return fix_f0_contour(frame_period, cand, best, f0_floor, allowed_range), temporal_positions


def _select_best_f0(current_f0, past_f0, cand, target_index, allowed_range):
    reference_f0 = (current_f0 * 3.0 - past_f0) / 2.0
    err = np.abs(reference_f0 - cand[:, target_index])
    best_f0 = cand[int(np.argmin(err)), target_index]
    if abs(1.0 - best_f0 / reference_f0) > allowed_range:
        return 0.0
    return best_f0


def fix_f0_contour(frame_period, cand, best, f0_floor, allowed_range):
    f0_length = len(best)
    vrm = int(0.5 + 1000.0 / frame_period / f0_floor) * 2 + 1     # voice_range_minimum
    out = np.zeros(f0_length)
    if f0_length <= vrm:
        return out
    # step 1: remove jumps
    base = np.zeros(f0_length)
    base[vrm : f0_length - vrm] = best[vrm : f0_length - vrm]
    s1 = np.zeros(f0_length)
    for i in range(vrm, f0_length):
        s1[i] = base[i] if abs((base[i] - base[i - 1]) / (K_SAFE_GUARD + base[i])) < allowed_range else 0.0
    # step 2: remove voiced sections shorter than the minimum
    s2 = s1.copy()
    center = (vrm - 1) // 2
    for i in range(center, f0_length - center):
        if (s1[i - center : i + center + 1] == 0).any():
            s2[i] = 0.0
    positive = [i for i in range(1, f0_length) if s2[i - 1] == 0 and s2[i] != 0]
    negative = [i - 1 for i in range(1, f0_length) if s2[i] == 0 and s2[i - 1] != 0]
    # step 3: extend every section forwards
    s3 = s2.copy()
    for n, start in enumerate(negative):
        limit = f0_length - 1 if n == len(negative) - 1 else negative[n + 1]
        for j in range(start, limit):
            s3[j + 1] = _select_best_f0(s3[j], s3[j - 1], cand, j + 1, allowed_range)
            if s3[j + 1] == 0:
                break
    # step 4: ... and backwards
    s4 = s3.copy()
    for n in range(len(positive) - 1, -1, -1):
        limit = 1 if n == 0 else positive[n - 1]
        for j in range(positive[n], limit, -1):
            s4[j - 1] = _select_best_f0(s4[j], s4[j + 1], cand, j - 1, allowed_range)
            if s4[j - 1] == 0:
                break
    return s4


def _fix_f0(power, numerator_i, fft_size, fs, initial_f0, n_harmonics):
    num = den = 0.0
    for i in range(n_harmonics):
        index = matlab_round(initial_f0 * fft_size / fs * (i + 1))
        inst = 0.0 if power[index] == 0.0 else index * fs / fft_size + numerator_i[index] / power[index] * fs / 2.0 / math.pi
        amp = math.sqrt(power[index])
        num += amp * inst
        den += amp * (i + 1.0)
    return num / (den + K_SAFE_GUARD)


def _refined_f0(x, fs, current_position, initial_f0):
    if initial_f0 <= K_FLOOR_F0_STONEMASK or initial_f0 > fs / 12.0:
        return 0.0
    half = int(1.5 * fs / initial_f0 + 1.0)
    window_length_in_time = (2.0 * half + 1.0) / fs
    n = 2 * half + 1
    base_time = (np.arange(n) - half) / fs
    fft_size = int(2.0 ** (2.0 + int(math.log(half * 2.0 + 1.0) / math.log(2.0))))
    basic_index = matlab_round((current_position + base_time[0]) * fs + 0.001)
    index_raw = basic_index + np.arange(n)
    t = (index_raw - 1.0) / fs - current_position
    main_w = 0.42 + 0.5 * np.cos(2.0 * np.pi * t / window_length_in_time) + 0.08 * np.cos(4.0 * np.pi * t / window_length_in_time)
    diff_w = np.empty(n)
    diff_w[0] = -main_w[1] / 2.0
    diff_w[1:-1] = -(main_w[2:] - main_w[:-2]) / 2.0
    diff_w[-1] = main_w[-2] / 2.0
    base_index = np.clip(index_raw - 1, 0, len(x) - 1)
    seg = x[base_index]
    main = np.fft.rfft(seg * main_w, fft_size)
    diff = np.fft.rfft(seg * diff_w, fft_size)
    numerator_i = main.real * diff.imag - main.imag * diff.real
    power = main.real ** 2 + main.imag ** 2
    tentative = _fix_f0(power, numerator_i, fft_size, fs, initial_f0, 2)
    if tentative <= 0.0 or tentative > initial_f0 * 2:   # GetTentativeF0: an overlarge fix is rejected
        mean_f0 = 0.0
    else:
        mean_f0 = _fix_f0(power, numerator_i, fft_size, fs, tentative, min(int(fs / 2.0 / tentative), 6))
    if abs(mean_f0 - initial_f0) > initial_f0 * 0.2:   # correction of more than 20 %: keep the initial estimate
        mean_f0 = initial_f0
    return mean_f0


def stonemask(x, f0, temporal_positions, fs):
    """``pyworld.stonemask``: instantaneous-frequency refinement of every voiced frame."""
    x = np.asarray(x, dtype=np.float64)
    return np.array([_refined_f0(x, fs, float(t), float(f)) for f, t in zip(f0, temporal_positions)], dtype=np.float64)


def extract_pitch(audio: np.ndarray, sr: int, hop: int) -> np.ndarray:
    """``Preprocessor.extract_pitch`` (preprocessor.py:244-285) for a mono float waveform: dio (speed 4, frame period
    hop / sr * 1000 ms) -> stonemask -> unvoiced frames filled by interpolation (ev_oracle.postprocess_pitch)."""
    from .ev_oracle import postprocess_pitch

    x = np.asarray(audio, dtype=np.float64)
    f0, t = dio(x, sr, frame_period=hop / sr * 1000, speed=4)
    f0 = stonemask(x, f0, t, sr)
    return postprocess_pitch(f0)
