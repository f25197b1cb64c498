"""oracle/refnb.py -- TEST INFRASTRUCTURE.  ctypes access to the compiled, unmodified reference
(oracle/_ref/libref_nbabfs*.so, built by oracle/Makefile from /root/reference) through the flat driver
oracle/ref_driver.c.  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this.
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}

ENERGY_LABELS = ("MM/MM Elect.", "MM/MM LJ", "MM/MM 1-4 Elect.", "MM/MM 1-4 LJ", "MM/MM Image Elect.", "MM/MM Image LJ")


def _path(omp=False):
    # omp == "shim": the same driver and reference containers with the reference's NBModelABFS.c / NBModelABFSState.c replaced by
    # pdynamo-mirror_b200/csrc/compat_shim.c on top of libnbabfs_b200.so (oracle/Makefile: libshim_nbabfs.so) -- the boundary under test
    name = "libshim_nbabfs.so" if omp == "shim" else ("libref_nbabfs_omp.so" if omp else "libref_nbabfs.so")
    return os.path.join(_HERE, "_ref", name)


def available(omp=False):
    return os.path.exists(_path(omp))


def _lib(omp=False):
    key = omp if omp == "shim" else bool(omp)
    if key in _LIBS:
        return _LIBS[key]
    path = _path(omp)
    lib = C.CDLL(path, mode=C.RTLD_LOCAL)
    dp, ip, vp = C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_void_p
    lib.refnb_create.restype = vp
    lib.refnb_create.argtypes = [C.c_int, dp, ip, C.c_int, ip, dp, dp, C.c_int, ip, dp, dp,
                                 C.c_int, ip, C.c_int, ip, C.c_int, dp, dp]
    lib.refnb_destroy.argtypes = [vp]
    lib.refnb_set_centering.restype = C.c_int
    lib.refnb_set_centering.argtypes = [vp, C.c_int]
    lib.refnb_set_fixed.restype = C.c_int
    lib.refnb_set_fixed.argtypes = [vp, C.c_int, C.POINTER(C.c_int)]
    lib.refnb_set_options.argtypes = [vp] + [C.c_double] * 6 + [C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int]
    lib.refnb_energy.restype = C.c_int
    lib.refnb_energy.argtypes = [vp, dp, dp, C.c_int, dp, dp, dp, dp]
    lib.refnb_num_primary_pairs.restype = C.c_long
    lib.refnb_num_primary_pairs.argtypes = [vp]
    lib.refnb_num_image_pairs.restype = C.c_long
    lib.refnb_num_image_pairs.argtypes = [vp]
    lib.refnb_num_14_pairs.restype = C.c_long
    lib.refnb_num_14_pairs.argtypes = [vp]
    lib.refnb_num_images.restype = C.c_int
    lib.refnb_num_images.argtypes = [vp]
    lib.refnb_uses_grid.restype = C.c_int
    lib.refnb_uses_grid.argtypes = [vp]
    lib.refnb_num_threads.restype = C.c_int
    lib.refnb_get_primary_pairs.argtypes = [vp, ip]
    lib.refnb_get_image_info.argtypes = [vp, C.c_int, ip, dp]
    lib.refnb_get_image_pairs.argtypes = [vp, C.c_int, ip]
    lib.refnb_make_factors.argtypes = [C.c_double] * 3 + [dp]
    lib.refnb_lj_table.argtypes = [C.c_int, dp, dp, C.c_int, ip, dp, dp]
    lib.refmm_energy.argtypes = [C.c_int, dp, C.c_int, ip, dp, dp, C.c_int, ip, dp, dp, C.c_int, ip, dp, dp, C.c_int, ip, dp, ip, dp, C.c_int, ip, dp, dp, dp, dp]
    lib.refnb_make_M.argtypes = [dp, dp, dp]
    lib.refnb_set_interaction_form.argtypes = [vp, C.c_int, C.c_int]
    lib.refnb_make_spline.restype = C.c_int
    lib.refnb_make_spline.argtypes = [C.c_int] + [C.c_double] * 3 + [C.c_int, dp, dp, dp]
    lib.refnb_spline_evaluate.argtypes = [C.c_int, dp, dp, C.c_double, dp, dp]
    if hasattr(lib, "refqc_create"):
        lib.refqc_create.restype = vp
        lib.refqc_create.argtypes = lib.refnb_create.argtypes + [C.c_int, ip, ip, C.c_int]
        lib.refqc_destroy.argtypes = [vp]
        lib.refqc_energy.restype = C.c_int
        lib.refqc_energy.argtypes = [vp, dp, dp, dp, dp, dp, dp, dp, dp, dp]
        lib.refqc_counts.argtypes = [vp, C.POINTER(C.c_long)]
        lib.refqc_get_pairs.argtypes = [vp, C.c_int, ip]
    _LIBS[key] = lib
    return lib


def _d(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _i(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_int))


class RefNB:
    """The reference NBModelABFS + state for one system dict (see pdynamo-mirror_b200/workloads.py)."""

    def __init__(self, system, omp=False, **options):
        self.lib = _lib(omp)
        s = self.sys = system
        self.n = s["n"]
        q = np.ascontiguousarray(s["charges"], np.float64)
        lt = np.ascontiguousarray(s["ljtypes"], np.int32)
        ex = np.ascontiguousarray(s["exclusions"], np.int32).reshape(-1)
        p14 = np.ascontiguousarray(s["pairs14"], np.int32).reshape(-1)
        rot = np.ascontiguousarray(s["rot"], np.float64).reshape(-1)
        trn = np.ascontiguousarray(s["trans"], np.float64).reshape(-1)
        ntrans = 0 if s["box"] is None else len(s["trans"])
        self.h = self.lib.refnb_create(self.n, _d(q), _i(lt), s["ntypes"], _i(s["tableindex"]), _d(s["tableA"]), _d(s["tableB"]),
                                       s["ntypes"], _i(s["tableindex14"]), _d(s["tableA14"]), _d(s["tableB14"]),
                                       len(ex) // 2, _i(ex) if len(ex) else None, len(p14) // 2, _i(p14) if len(p14) else None,
                                       ntrans, _d(rot) if ntrans else None, _d(trn) if ntrans else None)
        if not self.h:
            raise RuntimeError("refnb_create failed")
        self.opts = dict(dampingCutoff=0.5, innerCutoff=8.0, outerCutoff=12.0, listCutoff=13.5, dielectric=1.0,
                         electrostaticScale14=s.get("electrostaticScale14", 1.0), checkForInverses=True,
                         imageExpandFactor=0, cutoffCellSizeFactor=0.5, method=0, useGridByCell=True, sortIndices=False,
                         useAnalyticForm=True, splinePointDensity=50)
        centering = bool(options.pop("useCentering", False))
        self.set_options(**options)
        if s.get("fixed") is not None and len(s["fixed"]) > 0:
            self.set_fixed(s["fixed"])
        if centering and not self.lib.refnb_set_centering(self.h, 1):
            raise RuntimeError("refnb_set_centering failed")

    def set_fixed(self, indices):
        """fixedAtoms of NBModelABFSState_SetUp (the state is created anew, as for a new configuration)."""
        idx = np.ascontiguousarray(indices, np.int32).reshape(-1)
        if not self.lib.refnb_set_fixed(self.h, len(idx), _i(idx) if len(idx) else None):
            raise RuntimeError("refnb_set_fixed failed")

    def set_options(self, **kw):
        for k in kw:
            if k not in self.opts:
                raise ValueError("unknown option " + k)
        self.opts.update(kw)
        o = self.opts
        self.lib.refnb_set_options(self.h, o["dampingCutoff"], o["innerCutoff"], o["outerCutoff"], o["listCutoff"],
                                   o["dielectric"], o["electrostaticScale14"], int(o["checkForInverses"]),
                                   int(o["imageExpandFactor"]), o["cutoffCellSizeFactor"], int(o["method"]),
                                   int(o["useGridByCell"]), int(o["sortIndices"]))
        self.lib.refnb_set_interaction_form(self.h, int(bool(o["useAnalyticForm"])), int(o["splinePointDensity"]))

    def energy(self, xyz=None, box=None, force_new=False, gradients=True):
        """Returns dict(energies[6], grad[n,3], dEdM[3,3], updated, t_update, t_energy)."""
        xyz = np.ascontiguousarray(self.sys["xyz"] if xyz is None else xyz, np.float64)
        box = self.sys["box"] if box is None else box
        boxa = None if box is None else np.ascontiguousarray(box, np.float64)
        e = np.zeros(6)
        g = np.zeros((self.n, 3)) if gradients else None
        dm = np.zeros((3, 3)) if gradients else None
        tm = np.zeros(2)
        upd = self.lib.refnb_energy(self.h, _d(xyz), _d(boxa), int(force_new), _d(e), _d(g), _d(dm), _d(tm))
        if upd < 0:
            raise RuntimeError("reference NBModelABFS_Update failed")
        return dict(energies=e, grad=g, dEdM=dm, updated=bool(upd), t_update=tm[0], t_energy=tm[1])

    def counts(self):
        return dict(primary=self.lib.refnb_num_primary_pairs(self.h), images=self.lib.refnb_num_images(self.h),
                    image_pairs=self.lib.refnb_num_image_pairs(self.h), pairs14=self.lib.refnb_num_14_pairs(self.h),
                    grid=bool(self.lib.refnb_uses_grid(self.h)))

    def primary_pairs(self):
        n = self.lib.refnb_num_primary_pairs(self.h)
        p = np.zeros((n, 2), np.int32)
        if n:
            self.lib.refnb_get_primary_pairs(self.h, _i(p))
        return p

    def images(self):
        """list of dict(t,a,b,c,scale,pairs[np,2])"""
        out = []
        for k in range(self.lib.refnb_num_images(self.h)):
            info = np.zeros(6, np.int32)
            sc = np.zeros(1)
            self.lib.refnb_get_image_info(self.h, k, _i(info), _d(sc))
            p = np.zeros((info[4], 2), np.int32)
            if info[4]:
                self.lib.refnb_get_image_pairs(self.h, k, _i(p))
            out.append(dict(t=int(info[0]), a=int(info[1]), b=int(info[2]), c=int(info[3]), scale=float(sc[0]), pairs=p))
        return out

    def num_threads(self):
        return self.lib.refnb_num_threads()

    def close(self):
        if self.h:
            self.lib.refnb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def make_factors(damp, inner, outer):
    out = np.zeros(21)
    _lib().refnb_make_factors(damp, inner, outer, _d(out))
    return out


def lj_table(eps, sigma, style):
    eps = np.ascontiguousarray(eps, np.float64)
    sigma = np.ascontiguousarray(sigma, np.float64)
    nt = len(eps)
    ti = np.zeros(nt * nt, np.int32)
    ta = np.zeros(nt * (nt + 1) // 2)
    tb = np.zeros(nt * (nt + 1) // 2)
    _lib().refnb_lj_table(nt, _d(eps), _d(sigma), 1 if style == "amber" else 0, _i(ti), _d(ta), _d(tb))
    return ti, ta, tb


def make_spline(which, damp=0.5, inner=8.0, outer=12.0, density=50):
    """(x, y, h) of the reference's PairwiseInteractionABFS_Make*Spline: which = 0 elect. (kJ/mol), 1 LJ-A, 2 LJ-B, 3 elect. (a.u.)"""
    n = _lib().refnb_make_spline(which, damp, inner, outer, density, None, None, None)
    x, y, h = np.zeros(n), np.zeros(n), np.zeros(n)
    _lib().refnb_make_spline(which, damp, inner, outer, density, _d(x), _d(y), _d(h))
    return x, y, h


def spline_evaluate(x, y, x0):
    """CubicSpline_Evaluate of the spline the reference builds from (x, y) with zero end slopes: (f, df/dx)"""
    f, g = np.zeros(1), np.zeros(1)
    x = np.ascontiguousarray(x, np.float64); y = np.ascontiguousarray(y, np.float64)
    _lib().refnb_spline_evaluate(len(x), _d(x), _d(y), float(x0), _d(f), _d(g))
    return f[0], g[0]


def make_M(box6):
    b = np.ascontiguousarray(box6, np.float64)
    m = np.zeros((3, 3))
    im = np.zeros((3, 3))
    _lib().refnb_make_M(_d(b), _d(m), _d(im))
    return m, im


def _mm_call(fn, b, xyz, gradients):
    """b: bonded-term dict (tests/golden/dhfr_bonded.npz layout, per-term parameters); returns (energies5, grad or None)"""
    xyz = np.ascontiguousarray(xyz, np.float64)
    n = len(xyz)
    ia = lambda k, w: np.ascontiguousarray(b.get(k, np.zeros((0, w), np.int32)), np.int32).reshape(-1, w)
    da = lambda k: np.ascontiguousarray(b.get(k, np.zeros(0)), np.float64)
    bonds, angles, ubs, dih, imp = ia("bonds", 2), ia("angles", 3), ia("ureybradleys", 2), ia("dihedrals", 4), ia("impropers", 4)
    per = np.ascontiguousarray(b.get("dihedral_period", np.zeros(0, np.int32)), np.int32)
    keep = [da(k) for k in ("bond_eq", "bond_fc", "angle_eq", "angle_fc", "ub_eq", "ub_fc", "dihedral_fc", "dihedral_phase", "improper_eq", "improper_fc")]
    e = np.zeros(5)
    g = np.zeros((n, 3)) if gradients else None
    fn(n, _d(xyz), len(bonds), _i(bonds), _d(keep[0]), _d(keep[1]), len(angles), _i(angles), _d(keep[2]), _d(keep[3]),
       len(ubs), _i(ubs), _d(keep[4]), _d(keep[5]), len(dih), _i(dih), _d(keep[6]), _i(per), _d(keep[7]),
       len(imp), _i(imp), _d(keep[8]), _d(keep[9]), _d(e), _d(g))
    return e, g


def mm_energy(bonded, xyz, gradients=True):
    """bonded MM terms {bond, angle, Urey-Bradley, dihedral, improper} and their gradient"""
    return _mm_call(_lib().refmm_energy, bonded, xyz, gradients)


class RefQC:
    """The reference NBModelABFS with a QC region (no boundary atoms, MM link-atom coupling): MM/MM energy, QC/MM LJ energy, QC/MM
    potentials and electrostatic gradients through the compiled reference (oracle/ref_driver.c, refqc_*).  Default options only
    (ABFS 0.5 / 8 / 12 / 13.5 A, dielectric 1, electrostaticScale14 1)."""

    COUNT_LABELS = ("nbmmmm", "nbqcmmlj", "nbqcmmel", "nbmmmm14", "nbqcmmlj14", "nbqcmmel14", "inbmmmm_images", "inbmmmm_pairs", "inbqcmmlj_images",
                    "inbqcmmlj_pairs", "inbqcmmel_images", "inbqcmmel_pairs", "inbqcqclj_images", "inbqcqclj_pairs", "inbqcqcel_images", "inbqcqcel_pairs")

    def __init__(self, system, qc_index, qc_atomic_numbers, spline_point_density=50, omp=False):
        self.lib = _lib(omp)
        s = self.sys = system
        self.n = s["n"]
        q = np.ascontiguousarray(s["charges"], np.float64)
        lt = np.ascontiguousarray(s["ljtypes"], np.int32)
        ex = np.ascontiguousarray(s["exclusions"], np.int32).reshape(-1)
        p14 = np.ascontiguousarray(s["pairs14"], np.int32).reshape(-1)
        rot = np.ascontiguousarray(s["rot"], np.float64).reshape(-1)
        trn = np.ascontiguousarray(s["trans"], np.float64).reshape(-1)
        self.ntrans = 0 if s["box"] is None else len(s["trans"])
        self.qc_index = np.ascontiguousarray(qc_index, np.int32)
        zs = np.ascontiguousarray(qc_atomic_numbers, np.int32)
        self.nqc = len(self.qc_index)
        self.h = self.lib.refqc_create(self.n, _d(q), _i(lt), s["ntypes"], _i(s["tableindex"]), _d(s["tableA"]), _d(s["tableB"]),
                                       s["ntypes"], _i(s["tableindex14"]), _d(s["tableA14"]), _d(s["tableB14"]),
                                       len(ex) // 2, _i(ex) if len(ex) else None, len(p14) // 2, _i(p14) if len(p14) else None,
                                       self.ntrans, _d(rot) if self.ntrans else None, _d(trn) if self.ntrans else None,
                                       self.nqc, _i(self.qc_index), _i(zs), int(spline_point_density))
        if not self.h:
            raise RuntimeError("refqc_create failed")

    def energy(self, qc_charges, xyz=None, box=None):
        xyz = np.ascontiguousarray(self.sys["xyz"] if xyz is None else xyz, np.float64)
        box = self.sys["box"] if box is None else box
        boxa = None if box is None else np.ascontiguousarray(box, np.float64)
        qcq = np.ascontiguousarray(qc_charges, np.float64)
        e, pot, qcqc = np.zeros(10), np.zeros(self.nqc), np.zeros(self.nqc * (self.nqc + 1) // 2)
        g_lj, g_el, dm = np.zeros((self.n, 3)), np.zeros((self.n, 3)), np.zeros((3, 3))
        rc = self.lib.refqc_energy(self.h, _d(xyz), _d(boxa), _d(qcq), _d(e), _d(pot), _d(qcqc), _d(g_lj), _d(g_el), _d(dm))
        if rc < 0:
            raise RuntimeError("refqc_energy failed")
        counts = (C.c_long * 16)()
        self.lib.refqc_counts(self.h, counts)
        return dict(energies=e, potentials=pot, qcqc_potentials=qcqc, grad_lj=g_lj, grad_el=g_el, dEdM=dm,
                    counts=dict(zip(self.COUNT_LABELS, [int(v) for v in counts])))

    def pairs(self, which):
        """which: 0 nbmmmm, 1 nbqcmmlj, 2 nbqcmmel (first column of a QC/MM pair = position in the QC container)."""
        counts = (C.c_long * 16)()
        self.lib.refqc_counts(self.h, counts)
        p = np.zeros((int(counts[which]), 2), np.int32)
        if len(p):
            self.lib.refqc_get_pairs(self.h, which, _i(p))
        return p

    def close(self):
        if self.h:
            self.lib.refqc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
