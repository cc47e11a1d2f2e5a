"""First-order fundamental-parameter intensity of a fluorescence line by direct quadrature -- an anchor for the
forced-detection estimator that shares NO code with the oracle's or the engine's history loop (the role the "XAS tool"
value plays in the reference's only physics known-answer test, tests/test-xmimsim-main-CaSO4.c:19-20, set-up
tests/libxmimsim-test.c:182-285).

For a pencil beam of I0 photons along z that enters a stack of parallel layers, the expected first-order intensity of
line `line` of element Z excited in layer c is

    I = I0 * w_Z * sigma_shell(E0) * omega_shell * rate_line * rho_c
        * Integral_0^S ds  exp(-tau_in(s)) * Omega(p(s)) / (4 pi) * < exp(-tau_out(p(s), q)) >_q

with p(s) the point at path length s inside layer c, tau_in the optical depth of the beam up to p (at E0), Omega the
solid angle of the detector disc seen from p, and <.>_q the mean over points q of the detector window of the attenuation
(at the line energy) along the straight path p -> q.  The reference's estimator (src/xmi_variance_reduction.F90:29-726)
draws q uniformly in AREA and multiplies by Omega / 4 pi (`estimator=True`); the physical intensity weights q by solid
angle (`estimator=False`).  Both are computed by Gauss-Legendre / polar quadrature here; cross sections come straight
from the provider's functions.  Only the geometry of tests/inputs.py::caso4-like inputs is handled: no collimator, the
detector window inside the first (air) layer or in front of the stack, the beam on the z axis."""
import math

import numpy as np

K_SHELL = 0
KL3_LINE = 3          # |KL3_LINE| of xraylib's line macros (include/xmb_lines.h)


def _mu_layer(prov, layer, E):
    w = np.asarray(layer.weight, float)
    w = w / w.sum()
    return float(sum(wi * prov.CS_Total_Kissel(int(z), float(E)) for z, wi in zip(layer.Z, w)))


def first_order_line_intensity(inp, prov, Z, line=KL3_LINE, shell=K_SHELL, layer_index=None, estimator=True,
                               n_s=96, n_rad=24, n_phi=48):
    """Expected var_red_history[Z-1][line-1][0] of xmi_main_msim (x live_time), see the module docstring."""
    n = np.asarray(inp.n_sample_orientation, float)
    n = n / np.linalg.norm(n)
    nd = np.asarray(inp.n_detector_orientation, float)
    nd = nd / np.linalg.norm(nd)
    pw = np.asarray(inp.p_detector_window, float)
    R = math.sqrt(inp.area_detector / math.pi)
    layers = inp.layers
    ref = inp.reference_layer - 1
    # layer boundaries along the beam (z axis), as xmi_init_input derives them (src/xmi_main.F90:1741-1918)
    tz = [abs(l.thickness / n[2]) for l in layers]
    zb = [0.0] * len(layers)
    ze = [0.0] * len(layers)
    zb[ref] = inp.d_sample_source
    ze[ref] = zb[ref] + tz[ref]
    for j in range(ref + 1, len(layers)):
        zb[j] = ze[j - 1]; ze[j] = zb[j] + tz[j]
    for j in range(ref - 1, -1, -1):
        ze[j] = zb[j + 1]; zb[j] = ze[j] - tz[j]
    c = layer_index if layer_index is not None else max(i for i, l in enumerate(layers) if Z in l.Z)
    lay = layers[c]
    w = np.asarray(lay.weight, float)
    w_Z = float(w[list(lay.Z).index(Z)] / w.sum())
    assert len(inp.discrete) == 1 and not inp.continuous
    src = inp.discrete[0]
    E0 = src.energy
    I0 = (src.horizontal_intensity + src.vertical_intensity) * inp.live_time
    E_line = prov.LineEnergy(Z, -line)
    mu0 = [_mu_layer(prov, l, E0) for l in layers]
    mu1 = [_mu_layer(prov, l, E_line) for l in layers]
    const = w_Z * prov.CS_Photo_Partial(Z, shell, float(E0)) * prov.FluorYield(Z, shell) * prov.RadRate(Z, -line) * lay.density
    # optical depth of the beam in front of layer c
    tau_front = sum(mu0[j] * layers[j].density * tz[j] for j in range(c))
    # detector disc: orthonormal basis (u, v) of the window plane, polar Gauss-Legendre x uniform-phi rule
    a = np.array([1.0, 0.0, 0.0]) if abs(nd[0]) < 0.9 else np.array([0.0, 1.0, 0.0])
    u = np.cross(nd, a); u /= np.linalg.norm(u)
    v = np.cross(nd, u)
    xr, wr = np.polynomial.legendre.leggauss(n_rad)
    rad = 0.5 * R * (xr + 1.0); w_rad = 0.5 * R * wr * rad            # dA = r dr dphi
    phi = (np.arange(n_phi) + 0.5) * 2.0 * math.pi / n_phi
    q = pw[None, None, :] + rad[:, None, None] * (np.cos(phi)[None, :, None] * u + np.sin(phi)[None, :, None] * v)
    dA = np.broadcast_to(w_rad[:, None] * (2.0 * math.pi / n_phi), (n_rad, n_phi))
    # plane offsets of the layer boundaries along the sample normal (planes pass through (0, 0, z_boundary))
    hb = [zb_j * n[2] for zb_j in zb]
    he = [ze_j * n[2] for ze_j in ze]
    xs, ws = np.polynomial.legendre.leggauss(n_s)
    S = tz[c]
    total = 0.0
    for x_, w_ in zip(xs, ws):
        s = 0.5 * S * (x_ + 1.0)
        p = np.array([0.0, 0.0, zb[c] + s])
        d = q - p
        dist = np.linalg.norm(d, axis=2)
        dirv = d / dist[:, :, None]
        cosn = dirv @ n                                             # along the sample normal
        hp = p @ n
        hq = q @ n
        # path length inside every layer between the heights hp and hq (straight line, |dh| = |cosn| * length)
        tau = np.zeros_like(dist)
        lo = np.minimum(hp, hq); hi = np.maximum(hp, hq)
        for j in range(len(layers)):
            seg = np.clip(np.minimum(hi, he[j]) - np.maximum(lo, hb[j]), 0.0, None)
            tau += mu1[j] * layers[j].density * seg / np.abs(cosn)
        att = np.exp(-tau)
        dOmega = np.abs(dirv @ nd) / dist ** 2 * dA                  # solid-angle element of each window patch
        Omega = dOmega.sum()
        if estimator:
            mean_att = (att * dA).sum() / dA.sum()                  # q uniform in area, times Omega / 4 pi
            f = Omega / (4.0 * math.pi) * mean_att
        else:
            f = (att * dOmega).sum() / (4.0 * math.pi)
        total += 0.5 * S * w_ * math.exp(-tau_front - mu0[c] * lay.density * s) * f
    return I0 * const * total
