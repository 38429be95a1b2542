"""The three navigation-bit helpers the Borre channel calls after the correlators
(sydr/dsp/decoding.py: Prompt2Bit L16-27, LNAV_CheckPreambule L215-243, LNAV_DecodeTOW L247-275).

Host-side bit logic at 50 bit/s (stays Python).  The parity
equations are those of IS-GPS-200 Table 20-XIV, written here as XOR masks over a 32-bit window
(D29*, D30*, d1..d24, D25..D30) instead of the reference's +-1 products.
"""
from __future__ import annotations

import numpy as np

from ..utils.constants import (LNAV_PREAMBULE_BITS, LNAV_PREAMBULE_BITS_INV, LNAV_PREAMBULE_SIZE, LNAV_WORD_SIZE)

# Window positions (0 = D29*, 1 = D30*, 2..25 = d1..d24) entering each parity bit D25..D30.
_PARITY_TAPS = (
    (0, 2, 3, 4, 6, 7, 11, 12, 13, 14, 15, 18, 19, 21, 24),
    (1, 3, 4, 5, 7, 8, 12, 13, 14, 15, 16, 19, 20, 22, 25),
    (0, 2, 4, 5, 6, 8, 9, 13, 14, 15, 16, 17, 20, 21, 23),
    (1, 3, 5, 6, 7, 9, 10, 14, 15, 16, 17, 18, 21, 22, 24),
    (1, 2, 4, 6, 7, 8, 10, 11, 15, 16, 17, 18, 19, 22, 23, 25),
    (0, 4, 6, 7, 9, 10, 11, 12, 14, 16, 20, 23, 24, 25),
)


def Prompt2Bit(prompt: float, bit0: int = 0):
    return 1 if prompt > 0 else bit0


def word_parity_ok(window32) -> bool:
    """`window32`: 32 bits 0/1 = D29*, D30*, then the 30 received bits of a word.  True when the
    six received parity bits match (the data bits arrive XORed with D30*)."""
    w = [int(b) for b in window32]
    d30s = w[1]
    src = w[:2] + [b ^ d30s for b in w[2:26]]          # undo the D30* inversion of d1..d24
    for k, taps in enumerate(_PARITY_TAPS):
        p = 0
        for t in taps:
            p ^= src[t]
        if p != w[26 + k]:
            return False
    return True


def LNAV_CheckPreambule(bits):
    """bits[i-2 : i+62]: two bits of the previous word, then the first two words of a candidate
    subframe.  True when the preamble (or its inverse) is there and both words pass parity."""
    bits = np.asarray(bits)
    head = list(bits[2:2 + LNAV_PREAMBULE_SIZE])
    if head != LNAV_PREAMBULE_BITS and head != LNAV_PREAMBULE_BITS_INV:
        return False
    return word_parity_ok(bits[:LNAV_WORD_SIZE + 2]) and word_parity_ok(bits[LNAV_WORD_SIZE:2 * LNAV_WORD_SIZE + 2])


def LNAV_DecodeTOW(subframeBits, d30star: int):
    """Polarity-correct the ten words of a subframe in place (data bits of a word are inverted
    when the last bit of the previous word is 1), then read the truncated TOW count (x6 s) and
    the subframe id from the hand-over word.  Returns (tow, subframeID, bits-as-string)."""
    for j in range(10):
        if d30star == 1:
            seg = subframeBits[30 * j:30 * j + 24]
            subframeBits[30 * j:30 * j + 24] = 1 - seg
        d30star = subframeBits[30 * (j + 1) - 1]
    text = ''.join(str(int(b)) for b in subframeBits)
    return int(text[30:47], 2) * 6, int(text[49:52], 2), text
