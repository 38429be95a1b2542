"""Lock indicators and C/N0 estimators of sydr/dsp/lockindicator.py (scalar, once per 1 ms epoch,
host side): same names, arguments and expression order as the reference."""
from __future__ import annotations

import numpy as np


def lowPassFilter(new: float, old: float, alpha: float):
    """sydr/dsp/lockindicator.py:103-122."""
    return (1 - alpha) * old + alpha * new


def FLL_Lock_Borre(iprompt, iprompt_prev, qprompt, qprompt_prev, fll_lock_prev, alpha=0.01):
    """sydr/dsp/lockindicator.py:6-24."""
    lock = iprompt * iprompt_prev - qprompt * qprompt_prev
    lock *= np.sign(iprompt * iprompt_prev + qprompt * qprompt_prev)
    lock /= (iprompt ** 2 + qprompt ** 2)
    lock = abs(lock)
    return (1 - alpha) * fll_lock_prev + alpha * lock


def PLL_Lock_Borre(iprompt, qprompt, pll_lock_prev, alpha=0.01):
    """sydr/dsp/lockindicator.py:28-44."""
    nbd = iprompt ** 2 - qprompt ** 2
    nbp = iprompt ** 2 + qprompt ** 2
    return (1 - alpha) * pll_lock_prev + alpha * (nbd / nbp)


def CN0_NWPR(iPromptSum: float, qPromptSum: float, iPromptSum2: float, qPromptSum2: float, nbAccum=20,
             integrationPeriod=1e-3):
    """Narrow-band / wide-band power ratio, sydr/dsp/lockindicator.py:48-72."""
    nbp = iPromptSum ** 2 + qPromptSum ** 2
    wbp = iPromptSum2 + qPromptSum2
    normalisedPower = nbp / wbp
    return 10 * np.log10(1 / integrationPeriod * (normalisedPower - 1) / (nbAccum - normalisedPower))


def CN0_Beaulieu(ratio: float, N: int, T: float, old: float):
    """Beaulieu's estimator [Falletti, 2011], sydr/dsp/lockindicator.py:76-99."""
    lambda_c = 1 / (ratio / N)
    cn0 = lambda_c * (1 / T)
    return lowPassFilter(cn0, old, alpha=0.1)
