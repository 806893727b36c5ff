"""CPU: the 4-operation exact divide of csrc/common.cuh (two-term reciprocal, one Markstein
correction) equals the IEEE divide -- the same operation sequence with the host's FMA on the
adversarial generator of the GPU test (tests/test_div_gpu.py runs the device code on 2^31 pairs)."""
import ctypes as C
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _lib():
    src = os.path.join(HERE, 'emul', 'div_check.c')
    so = os.path.join(HERE, 'emul', 'libdiv_check.so')
    if not os.path.exists(so) or os.path.getmtime(src) > os.path.getmtime(so):
        # -ffp-contract=off: only the explicit fma() calls fuse, as in the -fmad=false device build
        subprocess.check_call(['gcc', '-O2', '-mfma', '-ffp-contract=off', '-fPIC', '-shared',
                               '-o', so, src, '-lm'])
    L = C.CDLL(so)
    L.div_check.restype = C.c_long
    L.div_check.argtypes = [C.c_uint64, C.c_long, C.POINTER(C.c_long), C.POINTER(C.c_double)]
    return L


def test_four_operation_divide_is_correctly_rounded():
    try:
        L = _lib()
    except (OSError, subprocess.CalledProcessError) as e:      # no FMA on this host
        pytest.skip('cannot build the host check: %s' % e)
    for seed in (1, 2, 3):
        q0_bad = C.c_long(0)
        ex = (C.c_double * 4)()
        bad = L.div_check(seed, 4000000, C.byref(q0_bad), ex)
        assert bad == 0, (bad, list(ex))
