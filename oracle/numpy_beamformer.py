"""The reference's numpy beamformer prototype, restated for Python 3 / numpy 2.

TEST / BASELINE INFRASTRUCTURE (the CPU path BASELINE.json names).  Follows
/root/reference/sandbox/Beamformer Theory.ipynb cells 2-3 and
sandbox/2D Beamformer.ipynb cells 5 and 8 (serial branch; `abs`, not the 1-D
notebook's `real`).  All float64, np.round (half-to-even), so
it agrees with the C++/GPU path only to ~1e-3 (README.md:211) -- it is the timed
CPU baseline and a sanity check, not the bit-exact oracle.
"""
import numpy as np

C = 299792458.0
TOT_CHANNELS = 2048
N_GPUS = 8
BW_PER_CHANNEL = (1.53 - 1.28) / TOT_CHANNELS


def freqs(gpu=0, n_freq=256, zero_pt=0):
    # Python-2 integer division in the notebook: gpu * tot_channels/(n_gpus-1)
    return np.array([1.53 - (zero_pt + gpu * TOT_CHANNELS // (N_GPUS - 1) + i) * BW_PER_CHANNEL for i in range(n_freq)])


def fourier_matrix(pos, theta, phi, gpu=0, n_freq=256):
    """A[k, beam, ant] = round(127 exp(-2j pi (x sin th + y sin ph)/lambda_k))/127"""
    pos = np.asarray(pos, dtype=np.float64)
    g = np.outer(np.sin(theta), pos[:, 0]) + np.outer(np.sin(phi), pos[:, 1])      # [beam, ant]
    lam = C / (freqs(gpu, n_freq) * 1e9)
    ph = -2.0 * np.pi * g[None, :, :] / lam[:, None, None]
    z = 127.0 * np.exp(1j * ph)
    return (np.round(z.real) + 1j * np.round(z.imag)) / 127.0


def signals(pos, theta_s, phi_s, gpu=0, n_freq=256):
    """S[k, ant, src] = round(7 exp(+2j pi (x sin th + y sin ph)/lambda_k))"""
    pos = np.asarray(pos, dtype=np.float64)
    g = np.outer(pos[:, 0], np.sin(theta_s)) + np.outer(pos[:, 1], np.sin(phi_s))  # [ant, src]
    lam = C / (freqs(gpu, n_freq) * 1e9)
    z = 7.0 * np.exp(2j * np.pi * g[None, :, :] / lam[:, None, None])
    return np.round(z.real) + 1j * np.round(z.imag)


def beamform_power(A, S, n_averaging=1):
    """out[beam, src] = sum_k 2*N_AVERAGING*|A[k] @ S[k]|^2   (vectorised per channel)"""
    out = np.zeros((A.shape[1], S.shape[2]))
    for k in range(A.shape[0]):
        v = A[k] @ S[k]
        out += 2 * n_averaging * (v.real ** 2 + v.imag ** 2)
    return out


def beamform_power_literal(A, pos, theta_s, phi_s, gpu=0, n_averaging=1):
    """The notebook's literal per-source Python loop (slow; small cases only)."""
    pos = np.asarray(pos, dtype=np.float64)
    n_freq = A.shape[0]
    f = freqs(gpu, n_freq)
    out = np.zeros((A.shape[1], len(theta_s)))
    for k in range(n_freq):
        wavelength = C / (f[k] * 1e9)
        for jj in range(len(theta_s)):
            signal = [np.round(7 * np.exp(2 * np.pi * 1j * (pos[i, 0] * np.sin(theta_s[jj]) + pos[i, 1] * np.sin(phi_s[jj])) / wavelength))
                      for i in range(pos.shape[0])]
            out[:, jj] += 2 * n_averaging * np.abs(np.dot(A[k, :, :], signal)) ** 2
    return out
