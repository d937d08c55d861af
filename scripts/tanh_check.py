"""
Accuracy of the device tanh (csrc/internal.h `tanh_accurate`: Eigen's degree-13/6 rational, evaluated in float32) against
float64 over 2.4e6 points, next to round 1's 1 - 2/(1 + 2^(2x log2 e)) form.  Pure numpy; run anywhere.

    python scripts/tanh_check.py
"""
import numpy as np

f = np.float32
A = [4.89352455891786e-03, 6.37261928875436e-04, 1.48572235717979e-05, 5.12229709037114e-08, -8.60467152213735e-11,
     2.00018790482477e-13, -2.76076847742355e-16]
B = [4.89352518554385e-03, 2.26843463243900e-03, 1.18534705686654e-04, 1.19825839466702e-06]


def tanh_rational(x):
    x = np.clip(x.astype(f), f(-9), f(9))
    x2 = x * x
    p = f(A[6])
    for c in A[5::-1]:
        p = p * x2 + f(c)
    q = f(B[3])
    for c in B[2::-1]:
        q = q * x2 + f(c)
    return (x * p / q).astype(f)


def tanh_exp_form(x):
    x = x.astype(f)
    e = np.exp2((x * f(2.8853900817779268)).astype(f)).astype(f)
    return (f(-2) * (f(1) / (e + f(1))).astype(f) + f(1)).astype(f)


if __name__ == '__main__':
    xs = np.concatenate([np.linspace(-10, 10, 2000001), np.logspace(-12, 1, 200000), -np.logspace(-12, 1, 200000)]).astype(f)
    ref = np.tanh(xs.astype(np.float64))
    for name, fn in (('rational 13/6 (current)', tanh_rational), ('1 - 2/(1+e^2x) (round 1)', tanh_exp_form)):
        got = fn(xs).astype(np.float64)
        rel = np.abs(got - ref) / np.maximum(np.abs(ref), 1e-300)
        small = np.abs(xs) < 1e-2
        print('%-28s max abs err %.3e   max rel err %.3e   max rel err for |x| < 1e-2: %.3e'
              % (name, np.abs(got - ref).max(), rel.max(), rel[small].max()))
