"""Constants of the MFCC front-end (utils/common.py:340-349 -> librosa.feature.mfcc defaults: sr 16 kHz, n_fft 2048,
hop 512, 128 slaney-normalised mel filters on the Slaney mel scale, DCT-II ortho).  Like net/utils/graph.py these are
parameter tables generated once on the host; the arithmetic on the audio runs in csrc/frontend.cu."""
import numpy as np

N_FFT, HOP, N_MELS, TOP_DB = 2048, 512, 128, 80.0

_F_SP = 200.0 / 3          # Slaney scale: linear below 1 kHz ...
_BREAK_HZ = 1000.0
_BREAK_MEL = _BREAK_HZ / _F_SP
_LOGSTEP = np.log(6.4) / 27.0   # ... logarithmic above


def _hz_to_mel(f):
    f = np.atleast_1d(np.asarray(f, dtype=np.float64))
    out = f / _F_SP
    hi = f >= _BREAK_HZ
    out[hi] = _BREAK_MEL + np.log(f[hi] / _BREAK_HZ) / _LOGSTEP
    return out


def _mel_to_hz(m):
    m = np.atleast_1d(np.asarray(m, dtype=np.float64))
    out = m * _F_SP
    hi = m >= _BREAK_MEL
    out[hi] = _BREAK_HZ * np.exp(_LOGSTEP * (m[hi] - _BREAK_MEL))
    return out


def mel_bank(sr=16000, n_fft=N_FFT, n_mels=N_MELS):
    """-> (bank float32 [n_mels, 1 + n_fft/2], span int32 [n_mels, 2]): triangular filters with unit-area (slaney)
    normalisation and the [lo, hi) range of bins where each filter is non-zero."""
    bins = np.arange(1 + n_fft // 2) * (sr / float(n_fft))
    edges = _mel_to_hz(np.linspace(_hz_to_mel(0.0)[0], _hz_to_mel(sr / 2.0)[0], n_mels + 2))
    left, centre, right = edges[:-2, None], edges[1:-1, None], edges[2:, None]
    rising = (bins[None, :] - left) / (centre - left)
    falling = (right - bins[None, :]) / (right - centre)
    bank = np.clip(np.minimum(rising, falling), 0.0, None) * (2.0 / (right - left))
    bank = bank.astype(np.float32)
    span = np.zeros((n_mels, 2), dtype=np.int32)
    for m in range(n_mels):
        nz = np.nonzero(bank[m])[0]
        if len(nz):
            span[m] = (nz[0], nz[-1] + 1)
    return bank, span


def dct_rows(n_out, n_in=N_MELS):
    """first n_out rows of the orthonormal DCT-II matrix of size n_in, float32"""
    k = np.arange(n_out, dtype=np.float64)[:, None]
    n = np.arange(n_in, dtype=np.float64)[None, :]
    m = np.cos(np.pi * k * (2.0 * n + 1.0) / (2.0 * n_in)) * np.sqrt(2.0 / n_in)
    m[0] /= np.sqrt(2.0)
    return m.astype(np.float32)
