#!/usr/bin/env python3
"""Extract compact golden vectors from the reference's shipped examples (run once, in the build container).

examples/{srm1155,srm1412,srm1132,In}.xmso hold, per interaction order, the unconvoluted and convoluted
2048-channel spectra and the per-line variance-reduction history (printed %g, 6 significant digits;
src/xmi_xml.c:1480-1567).  They are the only numeric outputs of the hot path the reference repository pins
(SURVEY.md 8c).  Written to tests/golden/<name>_xmso.npz; nothing reads /root/reference at test time."""
import os
import sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import xmimsim_b200 as x   # noqa: E402
src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/examples"
for name in ("srm1155", "srm1412", "srm1132", "In"):
    o = x.read_xmso(os.path.join(src, name + ".xmso"))
    keys = sorted(o["history"].keys())
    hz = np.array([k[0] for k in keys], np.int32)
    hl = np.array([k[1] for k in keys])
    he = np.array([o["history"][k]["energy"] for k in keys])
    n_int = o["conv"].shape[0]
    hc = np.zeros((len(keys), n_int))
    for i, k in enumerate(keys):
        for order, v in o["history"][k]["counts"].items():
            hc[i, order - 1] = v
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + "_xmso.npz"), conv=o["conv"], unconv=o["unconv"],
                        hist_Z=hz, hist_line=hl, hist_energy=he, hist_counts=hc)
    print(name, o["conv"].shape, len(keys), "lines; sum unconv", o["unconv"].sum(axis=1))


def svg_fixture(name="srm1155"):
    """The <svg_graphs> block of a shipped .xmso as arrays (tests/golden/<name>_svg.npz): per graphic its (kind, interaction), the
    box's min / max energy, the tick positions of both axes and the points -- what tests/test_io_cpu.py asks of the XMSO writer."""
    import xml.etree.ElementTree as ET
    root = ET.parse(os.path.join(src, name + ".xmso")).getroot()
    out = {}
    for gi, g in enumerate(root.find("svg_graphs").findall("graphic")):
        kind = g.find("id/name").text
        order = int(g.find("id/interaction").text)
        size = g.find("rect/size")
        out["g%d_id" % gi] = np.array([0 if kind == "convoluted" else 1, order])
        out["g%d_box" % gi] = np.array([float(size.find(k).text) for k in ("width", "height", "min_energy", "max_energy")])
        out["g%d_xt" % gi] = np.array([[float(i.find("value").text), float(i.find("name").text)] for i in g.find("rect/x-axis").findall("index")])
        out["g%d_yt" % gi] = np.array([[float(i.find("value").text), float(i.find("name").text)] for i in g.find("rect/y-axis").findall("index")])
        out["g%d_pts" % gi] = np.array([[float(p.find("x").text), float(p.find("y").text)] for p in g.find("points").findall("point")], np.float32)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + "_svg.npz"), **out)
    print(name, "svg graphics:", len(out) // 5)


svg_fixture()
