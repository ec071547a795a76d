"""Sky covariance models ``C_l(nu, nu')`` for the KL transform.

Mirrors the interface of ``drift.core.skymodel`` (reference
drift/core/skymodel.py:23-78): ``foreground_model`` and ``im21cm_model`` return
real arrays packed ``[pol, pol, l, freq, freq]``.  The reference evaluates the
angular power spectra with ``cora`` (cosmology + galaxy models), which is not
installable offline; this module evaluates the same *functional forms* the cora
models are built on with fixed, documented parameters and no cosmology:

* foregrounds: Santos, Cooray & Knox (2005) power laws
  ``A (l / 100)^-alpha (nu nu' / nu_0^2)^-beta exp(-ln^2(nu / nu') / (2 zeta^2))``
  for Galactic synchrotron (unpolarised and polarised) and point sources;
* 21 cm signal: a flat-sky power law in ``l`` with a Gaussian frequency
  correlation of width ``sigma_nu``.

They are SYNTHETIC stand-ins; users with ``cora`` can pass
their own arrays through ``KLTransform.signal`` / ``foreground`` overrides.
"""

import warnings

import numpy as np

_reionisation = False

# Written into the KL product files (attribute `skymodel`) so that products made with these
# stand-ins cannot be mistaken for ones made with cora's models.
MODEL_TAG = "driftscan_b200 synthetic power laws (Santos et al. 2005 forms; NOT cora)"
_warned = False


def _warn_once():
    global _warned
    if not _warned:
        _warned = True
        warnings.warn(
            "driftscan_b200.core.skymodel: cora is not available; the KL stage uses synthetic power-law "
            "C_l(nu, nu') stand-ins.  Eigenvalues, mode counts at the default thresholds and everything "
            "derived from them are NOT comparable with products of the reference (which uses cora's "
            "FullSkySynchrotron / PointSources / Corr21cm).  Pass your own covariances through "
            "KLTransform.signal / foreground overrides for physical results.", stacklevel=3)

# (A [K^2], alpha, beta, zeta), nu_0 = 130 MHz, l_0 = 100 (Santos et al. 2005 table 1, in K^2)
_SYNC = (7.00e-4, 2.80, 2.8, 4.0)
_PSRC = (5.70e-5, 1.10, 2.07, 1.0)
_POL = (1.55e-5, 2.80, 2.8, 0.64)  # polarised synchrotron: Faraday-decorrelated (small zeta)
_NU0 = 130.0
_L0 = 100.0


def _power_law(lmax, frequencies, A, alpha, beta, zeta):
    ell = np.maximum(np.arange(lmax + 1, dtype=np.float64), 1.0)
    f = np.asarray(frequencies, dtype=np.float64)
    f1, f2 = f[:, np.newaxis], f[np.newaxis, :]
    fpart = (f1 * f2 / _NU0**2) ** (-beta) * np.exp(-0.5 * (np.log(f1 / f2) / zeta) ** 2)
    lpart = A * (ell / _L0) ** (-alpha)
    return lpart[:, np.newaxis, np.newaxis] * fpart[np.newaxis, :, :]


def foreground_model(lmax, frequencies, npol, pol_frac=1.0, pol_length=None):
    """Foreground covariance ``[npol, npol, lmax+1, nfreq, nfreq]`` (skymodel.py:23-50)."""
    _warn_once()
    nfreq = np.asarray(frequencies).size
    cv_fg = np.zeros((npol, npol, lmax + 1, nfreq, nfreq))
    cv_fg[0, 0] = _power_law(lmax, frequencies, *_SYNC)
    if npol >= 3:
        A, alpha, beta, zeta = _POL
        if pol_length is not None:
            zeta = pol_length
        cv_fg[1, 1] = pol_frac * _power_law(lmax, frequencies, A, alpha, beta, zeta)
        cv_fg[2, 2] = pol_frac * _power_law(lmax, frequencies, A, alpha, beta, zeta)
    cv_fg[0, 0] += _power_law(lmax, frequencies, *_PSRC)
    return cv_fg


def im21cm_model(lmax, frequencies, npol, cr=None, temponly=False):
    """21 cm signal covariance (skymodel.py:53-78): TT only.  ``cr`` may be a callable
    ``cr(l, nu1, nu2)`` returning the angular power spectrum (as cora's
    ``angular_powerspectrum`` does)."""
    f = np.asarray(frequencies, dtype=np.float64)
    nfreq = f.size
    ell = np.arange(lmax + 1, dtype=np.float64)
    if cr is None:
        _warn_once()
    if cr is not None:
        cv_t = np.asarray(cr(ell[:, np.newaxis, np.newaxis], f[np.newaxis, :, np.newaxis], f[np.newaxis, np.newaxis, :]))
    else:
        amp = 1.0e-7 if not _reionisation else 1.0e-5  # K^2 (mK^2-level fluctuations)
        sigma_nu = 1.0  # MHz
        lpart = amp / (1.0 + ell / 200.0) ** 1.2
        fpart = np.exp(-0.5 * ((f[:, np.newaxis] - f[np.newaxis, :]) / sigma_nu) ** 2)
        cv_t = lpart[:, np.newaxis, np.newaxis] * fpart[np.newaxis, :, :]
    if temponly:
        return cv_t
    cv_sg = np.zeros((npol, npol, lmax + 1, nfreq, nfreq))
    cv_sg[0, 0] = cv_t
    return cv_sg
