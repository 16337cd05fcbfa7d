"""epgpy_b200 -- B200-native (sm_100a) execution of the epgpy EPG operator-chain hot path.

Drop-in for the path `epg.simulate` over T / Phi / E / P / R / S / D / X / ADC (+ order-1
derivatives) of py-baudin/epgpy: same operator API on the host, one fused CUDA kernel per sequence on
the device, reached through the C ABI of include/epgx.h (libepgx.so).  No CPU fallback.
"""

from . import common, core, engine, exchange, functions, lowering, operators, rfpulse, sharding, statematrix, utils
from . import core as epg

__version__ = "0.1.0"
