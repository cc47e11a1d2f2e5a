"""xmimsim_b200 -- B200-native XMI-MSIM photon-transport engine (host-side mirror of the C ABI).

Python here is plumbing: it parses XMSI, fills the reference-layout C structs and calls
libxmimsim_b200.so through ctypes.  All compute is in the CUDA library; it fails loudly when the
library or a GPU is missing.
"""
from . import abi            # noqa: F401
from .xmsi import InputD, LayerD, DiscreteD, ContinuousD, read_xmsi, read_xmso, CInput   # noqa: F401
from .engine import Simulation, main_options, tube_ebel   # noqa: F401
