"""CPU check of the table-based exp used for the RBF factors (csrc/common.cuh::exp_nonpos): the 64-entry 2^(j/64)
table in the CUDA source is correctly rounded, and the reduction + degree-5 polynomial it is combined with (restated
here with the constants parsed from the source) stays within 2 ulp of exp on the argument range of the Gaussian mode
(interpolatory use: Modes/kernels.py:251-257, exp(-(x-y)^2 / (2 l^2)))."""
import decimal
import math
import os
import re

import numpy as np

SRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                   "computationalhypergraphdiscovery_b200", "csrc", "common.cuh")


def _parse():
    src = open(SRC).read()
    body = src[src.index("g_exp2_tab[64] = {"):]
    body = body[:body.index("};")]
    tab = [float.fromhex(h) for h in re.findall(r"0x[0-9a-f.]+p[+-]?\d+", body)]
    fn = src[src.index("double exp_nonpos("):]
    fn = fn[:fn.index("\n}\n")]
    consts = [float.fromhex(h) for h in re.findall(r"-?0x[0-9a-f.]+p[+-]?\d+", fn)]
    return tab, consts, fn


def test_table_is_correctly_rounded():
    tab, _, _ = _parse()
    assert len(tab) == 64
    decimal.getcontext().prec = 60
    ln2 = decimal.Decimal(2).ln()
    for j, v in enumerate(tab):
        exact = (ln2 * decimal.Decimal(j) / 64).exp()
        assert abs(decimal.Decimal(v) - exact) <= decimal.Decimal(math.ulp(v)) / 2, j


def test_scheme_is_within_two_ulp():
    tab, consts, fn = _parse()
    # constants in order of appearance: 64/ln2, 1.5*2^52 (x2), ln2/64 hi, ln2/64 lo, 1/120, 1/24, 1/6
    l2e64, magic = consts[0], consts[1]
    hi, lo = abs(consts[3]), abs(consts[4])
    c5, c4, c3 = consts[5], consts[6], consts[7]
    assert magic == 1.5 * 2.0 ** 52 and abs(l2e64 - 64 / math.log(2)) < 1e-13
    assert abs(hi + lo - math.log(2) / 64) < 1e-18
    assert int(hi.hex().split("p")[0].split(".")[1], 16) % (1 << 20) == 0      # k * hi is exact for |k| < 2^20
    assert abs(c5 - 1 / 120) < 1e-17 and abs(c4 - 1 / 24) < 1e-17 and abs(c3 - 1 / 6) < 1e-16

    def expn(x):
        x = max(x, -700.0)
        k = int(np.rint(x * l2e64))
        kd = float(k)
        r = (x - kd * hi) - kd * lo
        p = c5 * r + c4
        p = p * r + c3
        p = p * r + 0.5
        p = p * r + 1.0
        p = p * r + 1.0
        q, j = k >> 6, k & 63
        return tab[j] * p * 2.0 ** q

    rng = np.random.default_rng(0)
    xs = -np.abs(rng.standard_normal(100000)) * rng.choice([1e-3, 0.1, 1.0, 10.0, 100.0], 100000)
    worst = 0.0
    for x in xs:
        ref = math.exp(x)
        if ref > 1e-300:
            worst = max(worst, abs(expn(x) - ref) / math.ulp(ref))
    assert worst <= 2.0, worst
    assert expn(0.0) == 1.0 and 0.0 < expn(-1e9) < 1e-300
