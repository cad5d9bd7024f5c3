"""CPU: unit checks of the SPIR-V interpreter's arithmetic helpers (oracle/spirv_vm). The end-to-end evidence that the
interpreter executes the shipped shaders correctly is that six independently written programs (reference SPIR-V in the
VM, C++ oracle, CUDA kernels) agree on every fixture; these tests cover the pieces that could agree by accident."""
import ctypes
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle", "spirv_vm"))


def test_fma32_is_correctly_rounded():
    from spirv_vm import fma32
    libm = ctypes.CDLL("libm.so.6")
    libm.fmaf.restype = ctypes.c_float
    libm.fmaf.argtypes = [ctypes.c_float] * 3
    rng = np.random.default_rng(3)
    cases = []
    for _ in range(4000):
        a, b = np.float32(rng.normal()), np.float32(rng.normal() * 10.0 ** rng.integers(-6, 6))
        c = np.float32(-a * b * (1.0 + rng.normal() * 2.0 ** -rng.integers(10, 30)))       # heavy cancellation
        cases.append((a, b, c))
    # exact ties in binary64: a*b = 1 + 2^-24 (halfway between two binary32 values), c breaks the tie by a hair
    for k in range(25, 60):
        cases.append((np.float32(1 + 2.0 ** -12), np.float32(1 + 2.0 ** -12), np.float32(2.0 ** -k)))
        cases.append((np.float32(1 + 2.0 ** -12), np.float32(1 + 2.0 ** -12), np.float32(-2.0 ** -k)))
        cases.append((np.float32(3.0), np.float32(1 + 2.0 ** -23), np.float32(2.0 ** -k)))
    for a, b, c in cases:
        got = fma32(a, b, c)
        want = np.float32(libm.fmaf(float(a), float(b), float(c)))
        assert struct.pack("<f", got) == struct.pack("<f", want), (a, b, c, got, want)


def test_sampler_footprint_and_level_selection():
    from spirv_vm import VM, Image, Sampler
    vm = VM.__new__(VM)                     # only the sampler helper is exercised
    lv0 = np.arange(32, dtype=np.float32).reshape(4, 8)
    lv1 = np.full((2, 4), -1.0, np.float32)
    img, smp = Image([lv0, lv1]), Sampler(True)
    f = np.float32
    # centre of texel (2,1): footprint is texels (1..2, 0..1) -> min = texel (1,0) = 1
    assert vm._sample(img, smp, [f(2.0 / 8), f(1.0 / 4)], f(0.0))[0] == 1.0
    # clamp to edge
    assert vm._sample(img, smp, [f(0.0), f(0.0)], f(0.0))[0] == 0.0
    assert vm._sample(img, smp, [f(1.0), f(1.0)], f(0.0))[0] == 31.0
    # nearest mip: lod 0.5 stays on level 0 (ceil(lod + 0.5) - 1), lod 0.51 selects level 1
    assert vm._sample(img, smp, [f(0.5), f(0.5)], f(0.5))[0] >= 0.0
    assert vm._sample(img, smp, [f(0.5), f(0.5)], f(0.51))[0] == -1.0
