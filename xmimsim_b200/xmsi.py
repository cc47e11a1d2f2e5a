"""XMSI input reader / XMSO reader+writer (stand-alone, xml.etree) and the ctypes input tree.

Follows the reference's reader conventions (src/xmi_xml.c): elements with |w| < 1e-20 dropped,
elements sorted by ascending Z inside a layer and weights normalised to sum 1 (:1209-1263), discrete
lines and continuous points sorted by energy (:966,969), detector_type strings SiLi / Ge / Si_SDD
(:1073-1082), optional scale_parameter with distribution_type (:777-796), nchannels default 2048
(:1068).
"""
import ctypes as C
import xml.etree.ElementTree as ET
from dataclasses import dataclass, field
from typing import List

from . import abi


@dataclass
class LayerD:
    Z: List[int]
    weight: List[float]
    density: float
    thickness: float


@dataclass
class DiscreteD:
    energy: float
    horizontal_intensity: float
    vertical_intensity: float
    sigma_x: float = 0.0
    sigma_xp: float = 0.0
    sigma_y: float = 0.0
    sigma_yp: float = 0.0
    distribution_type: int = 0
    scale_parameter: float = 0.0


@dataclass
class ContinuousD:
    energy: float
    horizontal_intensity: float
    vertical_intensity: float
    sigma_x: float = 0.0
    sigma_xp: float = 0.0
    sigma_y: float = 0.0
    sigma_yp: float = 0.0


@dataclass
class InputD:
    outputfile: str = "out.xmso"
    n_photons_interval: int = 10000
    n_photons_line: int = 100000
    n_interactions_trajectory: int = 4
    comments: str = ""
    layers: List[LayerD] = field(default_factory=list)
    reference_layer: int = 1
    d_sample_source: float = 100.0
    n_sample_orientation: List[float] = field(default_factory=lambda: [0.0, 0.707107, 0.707107])
    p_detector_window: List[float] = field(default_factory=lambda: [0.0, -1.0, 100.0])
    n_detector_orientation: List[float] = field(default_factory=lambda: [0.0, 1.0, 0.0])
    area_detector: float = 0.3
    collimator_height: float = 0.0
    collimator_diameter: float = 0.0
    d_source_slit: float = 100.0
    slit_size_x: float = 0.001
    slit_size_y: float = 0.001
    discrete: List[DiscreteD] = field(default_factory=list)
    continuous: List[ContinuousD] = field(default_factory=list)
    exc_layers: List[LayerD] = field(default_factory=list)
    det_layers: List[LayerD] = field(default_factory=list)
    detector_type: int = 2
    live_time: float = 1.0
    pulse_width: float = 1e-5
    gain: float = 0.02
    zero: float = 0.0
    fano: float = 0.12
    noise: float = 0.1
    nchannels: int = 2048
    crystal_layers: List[LayerD] = field(default_factory=list)


_DETECTOR_TYPES = {"SiLi": 0, "Ge": 1, "Si_SDD": 2}
_DISTRIBUTIONS = {"monochromatic": 0, "gaussian": 1, "lorentzian": 2}


def _f(node, tag, default=None):
    t = node.find(tag)
    if t is None or t.text is None:
        if default is None:
            raise ValueError("missing <%s>" % tag)
        return default
    return float(t.text)


def _layer(node):
    pairs = []
    for el in node.findall("element"):
        z = int(el.find("atomic_number").text)
        w = float(el.find("weight_fraction").text)
        if abs(w) < 1e-20:
            continue
        pairs.append((z, w))
    pairs.sort(key=lambda p: p[0])
    tot = sum(w for _, w in pairs)
    return LayerD([z for z, _ in pairs], [w / tot for _, w in pairs], _f(node, "density"), _f(node, "thickness"))


def _vec(node, tag):
    v = node.find(tag)
    return [float(v.find(a).text) for a in "xyz"]


def read_xmsi_root(root) -> InputD:
    d = InputD()
    g = root.find("general")
    d.outputfile = (g.find("outputfile").text or "").strip()
    d.n_photons_interval = int(g.find("n_photons_interval").text)
    d.n_photons_line = int(g.find("n_photons_line").text)
    d.n_interactions_trajectory = int(g.find("n_interactions_trajectory").text)
    d.comments = (g.find("comments").text or "") if g.find("comments") is not None else ""
    comp = root.find("composition")
    d.layers = [_layer(l) for l in comp.findall("layer")]
    d.reference_layer = int(comp.find("reference_layer").text)
    geo = root.find("geometry")
    d.d_sample_source = _f(geo, "d_sample_source")
    d.n_sample_orientation = _vec(geo, "n_sample_orientation")
    d.p_detector_window = _vec(geo, "p_detector_window")
    d.n_detector_orientation = _vec(geo, "n_detector_orientation")
    d.area_detector = _f(geo, "area_detector")
    d.collimator_height = _f(geo, "collimator_height")
    d.collimator_diameter = _f(geo, "collimator_diameter")
    d.d_source_slit = _f(geo, "d_source_slit")
    ss = geo.find("slit_size")
    d.slit_size_x = _f(ss, "slit_size_x")
    d.slit_size_y = _f(ss, "slit_size_y")
    exc = root.find("excitation")
    for n in exc.findall("discrete"):
        dd = DiscreteD(_f(n, "energy"), _f(n, "horizontal_intensity"), _f(n, "vertical_intensity"),
                       _f(n, "sigma_x", 0.0), _f(n, "sigma_xp", 0.0), _f(n, "sigma_y", 0.0), _f(n, "sigma_yp", 0.0))
        sp = n.find("scale_parameter")
        if sp is not None:
            dd.distribution_type = _DISTRIBUTIONS[sp.get("distribution_type", "monochromatic")]
            dd.scale_parameter = float(sp.text)
        d.discrete.append(dd)
    for n in exc.findall("continuous"):
        d.continuous.append(ContinuousD(_f(n, "energy"), _f(n, "horizontal_intensity"), _f(n, "vertical_intensity"),
                                        _f(n, "sigma_x", 0.0), _f(n, "sigma_xp", 0.0), _f(n, "sigma_y", 0.0),
                                        _f(n, "sigma_yp", 0.0)))
    d.discrete.sort(key=lambda e: e.energy)
    d.continuous.sort(key=lambda e: e.energy)
    ab = root.find("absorbers")
    if ab is not None:
        ep, dp = ab.find("excitation_path"), ab.find("detector_path")
        d.exc_layers = [_layer(l) for l in ep.findall("layer")] if ep is not None else []
        d.det_layers = [_layer(l) for l in dp.findall("layer")] if dp is not None else []
    det = root.find("detector")
    d.detector_type = _DETECTOR_TYPES[det.find("detector_type").text.strip()]
    d.live_time = _f(det, "live_time")
    d.pulse_width = _f(det, "pulse_width")
    d.nchannels = int(_f(det, "nchannels", 2048))
    d.gain = _f(det, "gain")
    d.zero = _f(det, "zero")
    d.fano = _f(det, "fano")
    d.noise = _f(det, "noise")
    d.crystal_layers = [_layer(l) for l in det.find("crystal").findall("layer")]
    return d


def read_xmsi(path) -> InputD:
    return read_xmsi_root(ET.parse(path).getroot())


def read_xmso(path):
    """Returns dict(input=InputD, conv=[n_int][nch], unconv=[n_int][nch], history={(Z,line): [counts per order]})."""
    import numpy as np
    root = ET.parse(path).getroot()

    def spectrum(tag):
        rows = []
        for ch in root.find(tag).findall("channel"):
            rows.append([float(c.text) for c in ch.findall("counts")])
        return np.array(rows).T.copy()
    hist = {}
    vr = root.find("variance_reduction_history")
    if vr is not None:
        for fl in vr.findall("fluorescence_line_counts"):
            z = int(fl.get("atomic_number"))
            for ln in fl.findall("fluorescence_line"):
                per = {int(c.get("interaction_number")): float(c.text) for c in ln.findall("counts")}
                hist[(z, ln.get("type"))] = dict(energy=float(ln.get("energy")), total=float(ln.get("total_counts")), counts=per)
    inp = read_xmsi_root(root.find("xmimsim")) if root.find("xmimsim") is not None else None
    return dict(input=inp, conv=spectrum("spectrum_conv"), unconv=spectrum("spectrum_unconv"), history=hist)


class CInput:
    """ctypes xmb_input tree built from an InputD; keeps every buffer alive."""

    def __init__(self, d: InputD):
        self._keep = []
        k = self._keep
        self.general = abi.General(1.0, d.outputfile.encode(), d.n_photons_interval, d.n_photons_line,
                                   d.n_interactions_trajectory, d.comments.encode())
        self.composition = abi.Composition(len(d.layers), self._layers(d.layers), d.reference_layer)
        self.geometry = abi.Geometry(d.d_sample_source, (C.c_double * 3)(*d.n_sample_orientation),
                                     (C.c_double * 3)(*d.p_detector_window), (C.c_double * 3)(*d.n_detector_orientation),
                                     d.area_detector, d.collimator_height, d.collimator_diameter, d.d_source_slit,
                                     d.slit_size_x, d.slit_size_y)
        disc = (abi.EnergyDiscrete * max(1, len(d.discrete)))()
        for i, e in enumerate(d.discrete):
            disc[i] = abi.EnergyDiscrete(e.energy, e.horizontal_intensity, e.vertical_intensity, e.sigma_x, e.sigma_xp,
                                         e.sigma_y, e.sigma_yp, e.distribution_type, e.scale_parameter)
        cont = (abi.EnergyContinuous * max(1, len(d.continuous)))()
        for i, e in enumerate(d.continuous):
            cont[i] = abi.EnergyContinuous(e.energy, e.horizontal_intensity, e.vertical_intensity, e.sigma_x, e.sigma_xp,
                                           e.sigma_y, e.sigma_yp)
        k += [disc, cont]
        self.excitation = abi.Excitation(len(d.discrete), C.cast(disc, C.POINTER(abi.EnergyDiscrete)),
                                         len(d.continuous), C.cast(cont, C.POINTER(abi.EnergyContinuous)))
        self.absorbers = abi.Absorbers(len(d.exc_layers), self._layers(d.exc_layers), len(d.det_layers),
                                       self._layers(d.det_layers))
        self.detector = abi.Detector(d.detector_type, d.live_time, d.pulse_width, d.gain, d.zero, d.fano, d.noise,
                                     d.nchannels, len(d.crystal_layers), self._layers(d.crystal_layers))
        self.input = abi.Input(C.pointer(self.general), C.pointer(self.composition), C.pointer(self.geometry),
                               C.pointer(self.excitation), C.pointer(self.absorbers), C.pointer(self.detector))

    def _layers(self, layers):
        arr = (abi.Layer * max(1, len(layers)))()
        for i, l in enumerate(layers):
            z = (C.c_int * len(l.Z))(*l.Z)
            w = (C.c_double * len(l.weight))(*l.weight)
            self._keep += [z, w]
            arr[i] = abi.Layer(len(l.Z), C.cast(z, abi.c_int_p), C.cast(w, abi.c_double_p), l.density, l.thickness)
        self._keep.append(arr)
        return C.cast(arr, C.POINTER(abi.Layer))
