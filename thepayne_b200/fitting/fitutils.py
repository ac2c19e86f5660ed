"""Host helpers that sit next to the likelihood (mirror of Payne/fitting/fitutils.py:11-37)."""
import numpy as np
from numpy.polynomial.chebyshev import chebval


def polycalc(coef, inwave):
    """Chebyshev continuum on the [-1, 1]-normalised wavelength (fitutils.py:11-20).
    Host-side convenience; inside the batched path this runs in the fused CUDA tail."""
    x = inwave - inwave.min()
    x = 2.0 * (x / x.max()) - 1.0
    return chebval(x, coef)


def airtovacuum(inwave):
    """Ciddor (1996) air -> vacuum, as used by fitstar.py:96-98 (fitutils.py:22-37)."""
    w = inwave * 1e-4
    d = 0.0 + (5.792105e-2 / (238.0185 - (1.0 / w ** 2.0))) + (1.67917e-3 / (57.362 - (1.0 / w ** 2.0)))
    return (w * (d + 1)) * 1e4
