"""ctypes front-end of the CPU oracle (oracle/lbm_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package never
imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / "liblbm_oracle.so"

CASE_LDC, CASE_POS, CASE_BIF, CASE_COR = 0, 1, 2, 3
Q = 19
CX = np.array([0, 1, -1, 0, 0, 0, 0, 1, 1, -1, -1, 1, 1, -1, -1, 0, 0, 0, 0])
CY = np.array([0, 0, 0, 1, -1, 0, 0, 1, -1, 1, -1, 0, 0, 0, 0, 1, -1, 1, -1])
CZ = np.array([0, 0, 0, 0, 0, 1, -1, 0, 0, 0, 0, 1, -1, 1, -1, 1, 1, -1, -1])
OPP = np.array([0, 2, 1, 4, 3, 6, 5, 10, 9, 8, 7, 14, 13, 12, 11, 18, 17, 16, 15])

# the reference's own opening-plane list for the coronary case (cor:77-141)
def cor_reference_rules(nx: int, ny: int, nz: int) -> np.ndarray:
    return np.array(
        [
            [0, 3, 1, ny - 2, 1, nz - 2, 1],
            [0, 272, 1, ny - 2, 1, nz - 2, 2],
            [2, 185, 217, 236, 113, 137, 4],
            [2, 191, 160, 205, 159, 199, 5],
            [2, 204, 1, nx - 2, 1, ny - 2, 6],
        ],
        dtype=np.int32,
    )


def build(force: bool = False) -> Path:
    """Compile the oracle (and oracle/_ref when /root/reference exists)."""
    newest = max((HERE / f).stat().st_mtime for f in ("lbm_oracle.c", "voxel_oracle.c"))
    if force or not LIB_PATH.exists() or LIB_PATH.stat().st_mtime < newest:
        subprocess.run(["make", "-C", str(HERE), "liblbm_oracle.so"], check=True, capture_output=True)
    return LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(str(LIB_PATH))
        _declare(_lib)
    return _lib


_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")


def _declare(L: C.CDLL) -> None:
    L.orc_geo_pre_ldc.argtypes = [C.c_int, C.c_int, C.c_int, _i32p]
    L.orc_geo_pre_pos.argtypes = [C.c_int, C.c_int, C.c_int, _i32p]
    L.orc_geo_pre_bif.argtypes = [C.c_int, C.c_int, C.c_int, _i32p, _i32p]
    L.orc_geo_pre_cor.argtypes = [C.c_int, C.c_int, C.c_int, _i32p, C.c_int, _i32p, _i32p]
    L.orc_index_transform.argtypes = [C.c_int, C.c_int, C.c_int, _i32p, _i32p]
    L.orc_index_transform.restype = C.c_int
    L.orc_index_dense.argtypes = [C.c_int, C.c_int, C.c_int, _i32p]
    L.orc_index_dense.restype = C.c_int
    L.orc_read_geo_file.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, _i32p]
    L.orc_read_geo_file.restype = C.c_long
    L.orc_read_vel_file.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, _i32p, _f32p, _f32p]
    L.orc_read_vel_file.restype = C.c_long
    for suf, dt in (("_f32", np.float32), ("_f64", np.float64)):
        rp = np.ctypeslib.ndpointer(dtype=dt, flags="C_CONTIGUOUS")
        g = lambda n: getattr(L, n + suf)
        g("orc_create").argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, _i32p, _i32p, C.c_int, C.c_double, C.c_double]
        g("orc_create").restype = C.c_void_p
        g("orc_destroy").argtypes = [C.c_void_p]
        g("orc_set_bc_planes").argtypes = [C.c_void_p, _f32p, _f32p]
        g("orc_set_cor_speeds").argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double]
        g("orc_set_pulse").argtypes = [C.c_void_p, C.c_double, C.c_double]
        g("orc_set_u_bc").argtypes = [C.c_void_p, C.c_double]
        g("orc_set_ldc_order").argtypes = [C.c_void_p, C.c_int]
        g("orc_initialize").argtypes = [C.c_void_p]
        g("orc_step").argtypes = [C.c_void_p, C.c_int]
        g("orc_get_fields").argtypes = [C.c_void_p, rp, rp, rp, rp]
        g("orc_get_populations").argtypes = [C.c_void_p, rp]
        g("orc_velsum").argtypes = [C.c_void_p]
        g("orc_velsum").restype = C.c_double
        g("orc_calc_res").argtypes = [C.c_void_p]
        g("orc_calc_res").restype = C.c_double
        g("orc_time_steps").argtypes = [C.c_void_p, C.c_int]
        g("orc_time_steps").restype = C.c_double
        g("orc_num_fluid").argtypes = [C.c_void_p]
        g("orc_num_fluid").restype = C.c_int


# ---------------------------------------------------------------- geometry
def geo_pre_ldc(nx, ny, nz):
    geo = np.zeros((nz, ny, nx), dtype=np.int32)
    lib().orc_geo_pre_ldc(nx, ny, nz, geo)
    return geo


def geo_pre_pos(nx, ny, nz):
    geo = np.zeros((nz, ny, nx), dtype=np.int32)
    lib().orc_geo_pre_pos(nx, ny, nz, geo)
    return geo


def geo_pre_bif(flag: np.ndarray):
    flag = np.ascontiguousarray(flag, dtype=np.int32)
    nz, ny, nx = flag.shape
    geo = np.zeros_like(flag)
    lib().orc_geo_pre_bif(nx, ny, nz, flag, geo)
    return geo


def geo_pre_cor(flag: np.ndarray, rules: np.ndarray):
    flag = np.ascontiguousarray(flag, dtype=np.int32)
    rules = np.ascontiguousarray(rules, dtype=np.int32).reshape(-1, 7)
    nz, ny, nx = flag.shape
    geo = np.zeros_like(flag)
    lib().orc_geo_pre_cor(nx, ny, nz, flag, len(rules), rules, geo)
    return geo


def index_transform(geo: np.ndarray):
    nz, ny, nx = geo.shape
    index = np.zeros_like(geo)
    nlat = lib().orc_index_transform(nx, ny, nz, np.ascontiguousarray(geo), index)
    return index, nlat


def index_dense(shape):
    nz, ny, nx = shape
    index = np.zeros(shape, dtype=np.int32)
    nlat = lib().orc_index_dense(nx, ny, nz, index)
    return index, nlat


def read_geo_file(path, nx, ny, nz, yfast=False):
    flag = np.zeros((nz, ny, nx), dtype=np.int32)
    n = lib().orc_read_geo_file(os.fsencode(str(path)), nx, ny, nz, int(yfast), flag)
    if n < 0:
        raise FileNotFoundError(path)
    return flag, n


def read_vel_file(path, geo: np.ndarray):
    nz, ny, nx = geo.shape
    inl = np.zeros((nz, nx), dtype=np.float32)
    out = np.zeros((nz, nx), dtype=np.float32)
    n = lib().orc_read_vel_file(os.fsencode(str(path)), nx, ny, nz, np.ascontiguousarray(geo), inl, out)
    if n < 0:
        raise FileNotFoundError(path)
    return inl, out, n


# ---------------------------------------------------------------- solver state
class Oracle:
    """One reference-style simulation (stored-node formulation) on the CPU."""

    def __init__(self, case_id, geo, index, nlat, tau, u_max=0.0, dtype=np.float32):
        self.dtype = np.dtype(dtype)
        self.suf = "_f32" if self.dtype == np.float32 else "_f64"
        self.geo = np.ascontiguousarray(geo, dtype=np.int32)
        self.index = np.ascontiguousarray(index, dtype=np.int32)
        self.nlat = int(nlat)
        self.case_id = case_id
        nz, ny, nx = self.geo.shape
        self.shape = (nz, ny, nx)
        self._h = self._fn("orc_create")(case_id, nx, ny, nz, self.geo, self.index, self.nlat, float(tau), float(u_max))

    def _fn(self, name):
        return getattr(lib(), name + self.suf)

    def set_bc_planes(self, inlety, outlety):
        self._fn("orc_set_bc_planes")(self._h, np.ascontiguousarray(inlety, dtype=np.float32),
                                      np.ascontiguousarray(outlety, dtype=np.float32))

    def set_cor_speeds(self, uin, uout, usub):
        self._fn("orc_set_cor_speeds")(self._h, uin, uout, usub)

    def set_u_bc(self, u_bc):
        self._fn("orc_set_u_bc")(self._h, float(u_bc))

    def set_ldc_order(self, mode):
        """0: walls bounce before fluid pulls (defined semantics); 1 / 2: the order ldc.cu's launch executes in,
        with the truly racy links (same koff iteration, different warp) taken fresh / stale"""
        self._fn("orc_set_ldc_order")(self._h, int(mode))

    def set_pulse(self, amp, period):
        self._fn("orc_set_pulse")(self._h, amp, period)

    def initialize(self):
        self._fn("orc_initialize")(self._h)

    def step(self, n=1):
        self._fn("orc_step")(self._h, int(n))

    def fields(self):
        out = [np.zeros(self.nlat, dtype=self.dtype) for _ in range(4)]
        self._fn("orc_get_fields")(self._h, *out)
        return out  # rho, ux, uy, uz (compact order)

    def populations(self):
        f = np.zeros((Q, self.nlat), dtype=self.dtype)
        self._fn("orc_get_populations")(self._h, f)
        return f

    def velsum(self):
        return self._fn("orc_velsum")(self._h)

    def calc_res(self):
        return self._fn("orc_calc_res")(self._h)

    def time_steps(self, n):
        return self._fn("orc_time_steps")(self._h, int(n))

    def num_fluid(self):
        return self._fn("orc_num_fluid")(self._h)

    def close(self):
        if self._h:
            self._fn("orc_destroy")(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---------------------------------------------------------------- STL voxeliser (voxel_oracle.c)
class VoxGrid(C.Structure):
    _fields_ = [("origin", C.c_double * 3), ("spacing", C.c_double), ("nx", C.c_int32), ("ny", C.c_int32),
                ("nz", C.c_int32), ("reserved", C.c_int32)]


def read_stl(path) -> np.ndarray:
    """binary STL -> float32 [ntri][3][3]"""
    raw = Path(path).read_bytes()
    n = int(np.frombuffer(raw, dtype="<u4", count=1, offset=80)[0])
    if 84 + 50 * n != len(raw):
        raise ValueError(f"{path}: not a binary STL")
    rec = np.frombuffer(raw, dtype=np.dtype([("n", "<f4", 3), ("v", "<f4", (3, 3)), ("a", "<u2")]), count=n, offset=84)
    return np.ascontiguousarray(rec["v"], dtype=np.float32)


def voxelize(tri: np.ndarray, origin, spacing: float, dims, z_range=None) -> np.ndarray:
    """uint8 [z1-z0][ny][nx], 1 inside; dims = (nx, ny, nz)"""
    L = lib()
    tri = np.ascontiguousarray(tri, dtype=np.float32).reshape(-1, 9)
    g = VoxGrid()
    g.origin[:] = [float(v) for v in origin]
    g.spacing = float(spacing)
    g.nx, g.ny, g.nz = (int(v) for v in dims)
    z0, z1 = (0, g.nz) if z_range is None else z_range
    out = np.zeros((z1 - z0, g.ny, g.nx), dtype=np.uint8)
    L.vox_oracle.restype = C.c_int
    L.vox_oracle.argtypes = [C.c_void_p, C.c_int64, C.POINTER(VoxGrid), C.c_int32, C.c_int32, C.c_void_p]
    rc = L.vox_oracle(tri.ctypes.data, tri.shape[0], C.byref(g), z0, z1, out.ctypes.data)
    if rc:
        raise MemoryError("vox_oracle")
    return out
