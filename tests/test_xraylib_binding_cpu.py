"""xmb_xrl_from_library: the provider that forwards to a libxrl loaded at run time (host_xraylib.cpp).  xraylib is not
in the image, so the binding is exercised against a stand-in library with xraylib 4's signatures whose return values
encode their arguments (tests/fixtures/fake_xrl.c, compiled here with gcc)."""
import ctypes as C
import os
import subprocess

import pytest

from xmimsim_b200 import abi

HERE = os.path.dirname(os.path.abspath(__file__))


def test_missing_library_fails_loudly():
    L = abi.lib()
    p = L.xmb_xrl_from_library(b"/nonexistent/libxrl.so")
    assert not p
    assert b"cannot load" in abi.last_error().encode()


def test_thunks_forward_their_arguments(tmp_path):
    so = str(tmp_path / "libfakexrl.so")
    subprocess.run(["gcc", "-shared", "-fPIC", "-O1", "-o", so, os.path.join(HERE, "fixtures", "fake_xrl.c")], check=True)
    L = abi.lib()
    pp = L.xmb_xrl_from_library(so.encode())
    assert pp, abi.last_error()
    p = pp.contents
    Z, E = 26, 7.5
    assert p.AtomicWeight(Z) == 52.0
    assert p.EdgeEnergy(Z, 3) == pytest.approx(26.03)
    assert p.LineEnergy(Z, -3) == pytest.approx(26.003)            # lines keep xraylib's negative macros
    assert p.CosKronTransProb(Z, 0) == pytest.approx(0.01 + 26e-5)  # XMB_FL12 = 0 -> FL12_TRANS = 1
    assert p.CosKronTransProb(Z, 12) == pytest.approx(0.13 + 26e-5)
    assert p.CS_Total_Kissel(Z, E) == 26007.5 and p.CS_Photo_Partial(Z, 2, E) == 2627.5
    assert p.FF_Rayl(Z, 1.5) == 24.5 and p.ComptonProfile_Partial(Z, 4, 0.5) == 30.5
    assert not p.AugerRate
    P = (C.c_double * 9)(*[10.0 ** -k for k in range(1, 10)])       # PK .. PM5 handed over
    w = [2, 3, 4, 5, 6, 7, 8, 9]
    def expect(tag, args):
        return tag + Z + E + sum(wi * a for wi, a in zip(w, args))
    # K: the partial photo-ionisation cross section
    assert p.VacancyCS(Z, 0, E, 4, P) == 2607.5
    # no cascade: only the shells of the same principal quantum number
    assert p.VacancyCS(Z, 1, E, 1, P) == pytest.approx(expect(1e-3, []))
    assert p.VacancyCS(Z, 3, E, 1, P) == pytest.approx(expect(3e-3, [P[1], P[2]]))
    assert p.VacancyCS(Z, 4, E, 1, P) == pytest.approx(expect(4e-3, []))
    assert p.VacancyCS(Z, 8, E, 1, P) == pytest.approx(expect(8e-3, [P[4], P[5], P[6], P[7]]))
    # cascades: every deeper shell from K on
    for mode, t in ((2, 0.1), (3, 0.2), (4, 0.3)):
        for shell in range(1, 9):
            assert p.VacancyCS(Z, shell, E, mode, P) == pytest.approx(expect(t + shell * 1e-3, list(P[:shell])), rel=1e-14)
