"""lbm-b200: B200-native (sm_100a) D3Q19 BGK Lattice-Boltzmann solver, drop-in for the case
interface of Xinhuan-Imperial/Lattice-Boltzmann-Method-GPU.  All compute lives in the in-tree CUDA
library (csrc/ -> liblbm_b200.so); this package is the ctypes host mirror of the reference's
geo_pre / index_transform / read_vel / initialize / update / outputSave sequence."""
from .api import *  # noqa: F401,F403
from .api import Case, Group, CaseDesc, LbmError, load_library, make_case, case_defaults, p2p_open, voxelize, ABI_SYMBOLS, LIB_PATH  # noqa: F401
