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
