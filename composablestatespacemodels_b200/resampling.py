"""Resampling schemes (reference: model/Resampling.scala:63-96) as `Resample[A]` values.

Each function has the reference's signature `(particles, weights) => particles`, draws its
uniforms the way the reference does (host RNG; one for systematic, n for stratified/multinomial)
and computes the ancestor indices on the GPU (`cssm_resample`).  Inside a filter the same kernels
run without leaving the device; the functions below are the stand-alone seam (S2 in SURVEY.md).
"""
import numpy as np

from . import _abi

_rng = np.random.default_rng()


def seed(s):
    """Seed the host generator that plays the role of scala.util.Random."""
    global _rng
    _rng = np.random.default_rng(s)


def ancestors(kind, weights, u, device=0):
    """Ancestor indices for given weights and uniforms (bit-exact against the oracle)."""
    w = np.ascontiguousarray(weights, dtype=np.float64)
    u = np.ascontiguousarray(np.atleast_1d(u), dtype=np.float64)
    if w.size == 0:
        return np.empty(0, dtype=np.int32)
    anc = np.empty(w.size, dtype=np.int32)
    _abi.check(_abi.lib().cssm_resample(kind, _abi.dptr(w), w.size, _abi.dptr(u), u.size,
                                        anc.ctypes.data_as(_abi.c_int32_p), device))
    return anc


def _take(particles, anc):
    if isinstance(particles, np.ndarray):
        return particles[anc]
    return [particles[i] for i in anc]


class Resampling:
    @staticmethod
    def systematicResampling(particles, weights):
        """model/Resampling.scala:63-72"""
        return _take(particles, ancestors(_abi.RESAMPLE_SYSTEMATIC, weights, _rng.random(1)))

    @staticmethod
    def stratifiedResampling(s, w):
        """model/Resampling.scala:78-86"""
        return _take(s, ancestors(_abi.RESAMPLE_STRATIFIED, w, _rng.random(len(w))))

    @staticmethod
    def multinomialResampling(particles, weights):
        """model/Resampling.scala:92-96"""
        return _take(particles, ancestors(_abi.RESAMPLE_MULTINOMIAL, weights, _rng.random(len(weights))))

    @staticmethod
    def residualResampling(particles, weights, return_ancestors=False):
        """model/Resampling.scala:130-146, CORRECTED.  `weights` are LOG-likelihoods (the reference normalises them with
        expNormalise).  Particle i is copied k_i = floor(n w_i) times; the remaining m = n - sum k_i places are drawn with
        probabilities proportional to the residuals n w_i - k_i.  The reference's last two lines do not run --
        `multinomialResampling(Vector.range(1, m), residualWeights)` draws n indices over the n residual weights and looks
        them up in a vector of m - 1 items, then indexes the particles with those items -- so there is no behaviour to be
        faithful to; this is the algorithm its doc comment (and Liu & Chen 1998) describes: m multinomial draws over the
        residual weights, as ancestors of the ORIGINAL particles.  The draws run on the device (cssm_resample, the same
        Multinomial.draw inverse-CDF walk as multinomialResampling); the deterministic copies are host arithmetic."""
        w = Resampling.expNormalise(weights)
        n = w.size
        ki = np.floor(w * n).astype(np.int64)
        anc = np.repeat(np.arange(n, dtype=np.int32), ki)
        m = n - anc.size
        if m > 0:
            residual = n * w - ki
            draws = ancestors(_abi.RESAMPLE_MULTINOMIAL, residual, _rng.random(n))[:m]  # draw j uses uniform j
            anc = np.concatenate([anc, draws])
        return anc if return_ancestors else _take(particles, anc)

    @staticmethod
    def kind_of(resample):
        """Map a Resample value (or a kind constant / name) to the device's resample kind."""
        table = {
            Resampling.systematicResampling: _abi.RESAMPLE_SYSTEMATIC,
            Resampling.stratifiedResampling: _abi.RESAMPLE_STRATIFIED,
            Resampling.multinomialResampling: _abi.RESAMPLE_MULTINOMIAL,
            "systematic": _abi.RESAMPLE_SYSTEMATIC, "stratified": _abi.RESAMPLE_STRATIFIED,
            "multinomial": _abi.RESAMPLE_MULTINOMIAL,
            _abi.RESAMPLE_SYSTEMATIC: _abi.RESAMPLE_SYSTEMATIC, _abi.RESAMPLE_STRATIFIED: _abi.RESAMPLE_STRATIFIED,
            _abi.RESAMPLE_MULTINOMIAL: _abi.RESAMPLE_MULTINOMIAL,
        }
        try:
            return table[resample]
        except (KeyError, TypeError):
            raise Exception("a filter on the GPU needs one of Resampling.systematicResampling, "
                            "stratifiedResampling or multinomialResampling") from None

    @staticmethod
    def normalise(prob):
        """model/Resampling.scala:21-24 (host helper)"""
        prob = np.asarray(prob, dtype=np.float64)
        return prob / prob.sum()

    # ---- the small host helpers of model/Resampling.scala (used around the filter, never on N-sized clouds) ----
    @staticmethod
    def indentity(samples, weights):
        """model/Resampling.scala:29 (sic)"""
        return samples

    @staticmethod
    def expNormalise(prob):
        """model/Resampling.scala:102-108"""
        prob = np.asarray(prob, dtype=np.float64)
        w1 = np.exp(prob - prob.max())
        return w1 / w1.sum()

    @staticmethod
    def cumSum(l):
        """model/Resampling.scala:113-115: scanLeft from zero, n + 1 values"""
        return np.concatenate([[0.0], np.cumsum(np.asarray(l, dtype=np.float64))])

    @staticmethod
    def empDist(w):
        """model/Resampling.scala:120-122"""
        return Resampling.cumSum(Resampling.normalise(w))

    @staticmethod
    def sampleOne(s):
        """model/Resampling.scala:151-154"""
        return s[int(_rng.integers(0, len(s)))]

    @staticmethod
    def sampleMany(n, s):
        """model/Resampling.scala:159-162: uniformly without replacement"""
        idx = _rng.permutation(len(s))[:n]
        return _take(s, idx)
