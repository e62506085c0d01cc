"""Restatement of the random streams the reference's tests draw their data from, so that the
RNG-dependent goldens become executable without Julia.  TEST INFRASTRUCTURE ONLY.

* ``StableRNG(seed)``: StableRNGs.jl ``LehmerRNG`` -- 128-bit multiplicative congruential generator,
  state = (seed << 1) | 1, multiplier 0x45a31efc5a35d971261fd0407a968add, output = high 64 bits.
* ``randn``: Julia's ``Random.randn(::AbstractRNG, Float64)`` -- 256-layer Marsaglia-Tsang
  ziggurat on a 52-bit integer (stdlib Random/src/normal.jl); tables regenerated here from the
  ziggurat recursion (the stdlib hard-codes them; first entries checked in the tests:
  wi[1] = 1.7367254121602630e-15, wi[2] = 9.5586603514556339e-17, fi[2] = 9.7710170126767082e-01,
  ki[1] = 0x0007799ec012f7b2).  Table values may differ from the hard-coded literals in the last
  ulp, which moves a sample by one ulp -- far below the goldens' 0.01 tolerance.
* ``rand`` (Float64 in [0,1)): ``reinterpret(Float64, 0x3ff0... | (u & 2^52-1)) - 1``.

Used by tests/test_reference_rng_goldens.py to regenerate
test/models/statespace/ulgssm_tests.jl:27-32 and mlgssm_test.jl:70-97 data.
"""
from __future__ import annotations

import math
import struct

import numpy as np

_MASK128 = (1 << 128) - 1
_MULT = 0x45A31EFC5A35D971261FD0407A968ADD
ZIG_R = 3.6541528853610087963519472518


def _tables():
    r = ZIG_R
    f = lambda x: math.exp(-0.5 * x * x)
    v = r * f(r) + math.sqrt(math.pi / 2.0) * math.erfc(r / math.sqrt(2.0))
    x = [0.0] * 257
    x[255] = r
    x[256] = v / f(r)               # pseudo-edge of the base strip
    for i in range(255, 1, -1):
        x[i - 1] = math.sqrt(-2.0 * math.log(v / x[i] + f(x[i])))
    x[0] = 0.0
    two51 = float(1 << 51)
    wi = [0.0] * 256
    ki = [0] * 256
    fi = [0.0] * 256
    # index 0 (Julia 1): base strip; index i (Julia i+1): layer with right edge x_i
    wi[0] = x[256] / two51
    ki[0] = int(math.floor(two51 * r / x[256]))
    fi[0] = 1.0
    for i in range(1, 256):
        wi[i] = x[i] / two51
        ki[i] = int(math.floor(two51 * x[i - 1] / x[i]))
        fi[i] = f(x[i])
    return ki, wi, fi


KI, WI, FI = _tables()


class StableRNG:
    def __init__(self, seed: int):
        assert seed >= 0
        self.state = ((seed << 1) | 1) & _MASK128

    def u64(self) -> int:
        self.state = (self.state * _MULT) & _MASK128
        return self.state >> 64

    def rand(self) -> float:
        bits = 0x3FF0000000000000 | (self.u64() & 0x000FFFFFFFFFFFFF)
        return struct.unpack("<d", struct.pack("<Q", bits))[0] - 1.0

    def randn(self) -> float:
        while True:
            r = self.u64() & 0x000FFFFFFFFFFFFF
            rabs = r >> 1
            idx = rabs & 0xFF
            x = (-rabs if (r & 1) else rabs) * WI[idx]
            if rabs < KI[idx]:
                return x
            if idx == 0:
                while True:
                    xx = -(1.0 / ZIG_R) * math.log(self.rand())
                    yy = -math.log(self.rand())
                    if yy + yy > xx * xx:
                        return (-ZIG_R - xx) if ((rabs >> 8) & 1) else (ZIG_R + xx)
            elif (FI[idx - 1] - FI[idx]) * self.rand() + FI[idx] < math.exp(-0.5 * x * x):
                return x
            # else: draw again (tail-recursive randn(rng) in the stdlib)

    def randn_vec(self, n: int) -> np.ndarray:
        return np.array([self.randn() for _ in range(n)])
