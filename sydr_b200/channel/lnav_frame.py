"""LNAV subframe synchronisation shared by the Borre and Kaplan channel classes.

Both reference channels run the same search once a navigation bit has been appended to their bit
buffer (sydr/channel/channel_l1ca_borre.py:493-573, sydr/channel/channel_l1ca_kaplan.py:760-823): look
for the preamble (either polarity, both words parity-clean) in the newest 62 bits; the second
preamble found exactly one subframe after the first declares subframe synchronisation; from then
on every full buffer (300 + 62 bits) must end with a preamble, else synchronisation is dropped;
a full, valid buffer yields TOW, subframe id and the subframe's bits.

`advance_frame(ch)` performs one such step on the channel's own members (`navBitsBuffer`,
`navBitsCounter`, `navBitBufferSize`, `preambuleFound`, `trackFlags`) and returns the decoded
`(tow, subframe_id, bits)` or None.  The buffer is re-seated (newest 62 bits at the front, counter
62) after a decode, and the returned TOW is the one *decoded*; callers add the age of those 62 bits.
"""
from __future__ import annotations

import numpy as np

from ..dsp.decoding import LNAV_CheckPreambule, LNAV_DecodeTOW
from ..utils.constants import LNAV_SUBFRAME_SIZE, LNAV_WORD_SIZE
from ..utils.enumerations import TrackingFlags

HEAD_BITS = 2 + 2 * LNAV_WORD_SIZE          # two bits of the previous word + TLM + HOW


def _reseat(ch, first: int, zero_rest: bool):
    """Newest HEAD_BITS bits (starting at `first`) to the front of a fresh buffer."""
    fresh = np.empty_like(ch.navBitsBuffer)
    if zero_rest:
        fresh[HEAD_BITS:] = 0
    fresh[:HEAD_BITS] = ch.navBitsBuffer[first:first + HEAD_BITS]
    ch.navBitsBuffer = fresh
    ch.navBitsCounter = HEAD_BITS


def _newest_head_is_preamble(ch) -> bool:
    first = ch.navBitsCounter - HEAD_BITS
    return LNAV_CheckPreambule(ch.navBitsBuffer[first:first + HEAD_BITS])


def advance_frame(ch):
    if ch.navBitsCounter < HEAD_BITS:
        return None
    if not (ch.trackFlags & TrackingFlags.SUBFRAME_SYNC):
        first = ch.navBitsCounter - HEAD_BITS
        if not _newest_head_is_preamble(ch):
            if ch.navBitsCounter == ch.navBitBufferSize:       # buffer full while searching: drop the oldest bit
                slid = np.empty_like(ch.navBitsBuffer)
                slid[:-1] = ch.navBitsBuffer[1:]
                ch.navBitsBuffer = slid
                ch.navBitsCounter -= 1
            return None
        if ch.preambuleFound and first == LNAV_SUBFRAME_SIZE:
            ch.trackFlags |= TrackingFlags.SUBFRAME_SYNC
        else:                                                  # first sighting: restart the buffer at this preamble
            _reseat(ch, first, zero_rest=True)
            ch.preambuleFound = True
    if ch.navBitsCounter < ch.navBitBufferSize:
        return None
    first = ch.navBitsCounter - HEAD_BITS
    if not _newest_head_is_preamble(ch):                       # the next subframe does not start where it should
        ch.navBitsCounter = 0
        ch.trackFlags ^= TrackingFlags.SUBFRAME_SYNC
        return None
    decoded = LNAV_DecodeTOW(ch.navBitsBuffer[2:2 + LNAV_SUBFRAME_SIZE], ch.navBitsBuffer[1])
    _reseat(ch, first, zero_rest=False)
    return decoded
