"""The constants of sydr/utils/constants.py that the hot path and its callers use."""
PI = 3.1415926535898            # GPS ICD value of pi (constants.py:4); PLL_costa divides by 2*PI
HALF_PI = PI / 2.0
TWO_PI = PI * 2.0

LNAV_PREAMBULE_BITS = [1, 0, 0, 0, 1, 0, 1, 1]
LNAV_PREAMBULE_BITS_INV = [0, 1, 1, 1, 0, 1, 0, 0]
LNAV_PREAMBULE_SIZE = 8
LNAV_MS_PER_BIT = 20
LNAV_SUBFRAME_SIZE = 300
LNAV_WORD_SIZE = 30

GPS_L1CA_CODE_SIZE_BITS = 1023
GPS_L1CA_CODE_FREQ = 1.023e6
GPS_L1CA_CODE_MS = 1

# Digital loop filter constants [Kaplan, 2006, p180] (constants.py:80-85)
W0_BANDWIDTH_1 = 0.25
W0_BANDWIDTH_2 = 0.53
W0_BANDWIDTH_3 = 0.7845
W0_SCALE_A2 = 1.414
W0_SCALE_A3 = 1.1
W0_SCALE_B3 = 2.4
