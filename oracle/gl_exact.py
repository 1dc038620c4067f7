"""TEST INFRASTRUCTURE -- exact oracle for the Gruenwald-Letnikov EXTENSION rule (not part of the reference, so
there is no reference to pin it on: SURVEY.md section 8(f)(3) asks for validation against mpmath).

    w_0 = 1, w_x = w_{x-1} * (1 - ap1 / x),  ap1 = fl(alpha + 1)  (the rounded sum is part of the definition)
    stencil weight at offset -x: fl( fl((lf/p) ** (-alpha)) * fl(w_x) )

Products are evaluated with mpmath at 200 bits and rounded once."""
import mpmath as mp
import numpy as np


def weights(alpha, count):
    mp.mp.prec = 200
    ap1 = mp.mpf(float(alpha) + 1.0)
    w = mp.mpf(1)
    out = [1.0]
    for x in range(1, count):
        w = w * (1 - ap1 / x)
        out.append(float(w))
    return np.array(out)


def multiplier(alpha, lf, p):
    mp.mp.prec = 200
    step = float(lf) / float(p)
    return float(mp.power(mp.mpf(step), mp.mpf(-float(alpha))))


def left_arrays(alpha, lf, p):
    w = weights(alpha, p + 1)
    m = multiplier(alpha, lf, p)
    return -p, (m * w)[::-1].copy()        # offsets -p..0
