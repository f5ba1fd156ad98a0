"""ctypes access to the checkers under oracle/ (tests only).

    RefGPU  : oracle/_ref/libphoxref_{debugtag,production}.so - the REFERENCE's device headers compiled
              unmodified behind a brute-force traversal harness (oracle/ref_gpu_harness.cu)
"""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE = os.path.join(ROOT, "oracle")


class RefConfig(C.Structure):
    _fields_ = [("max_bounce", C.c_int), ("max_record", C.c_int), ("event_index", C.c_int), ("pad", C.c_int),
                ("tmin", C.c_float), ("tmin0", C.c_float), ("tmax", C.c_float), ("max_time", C.c_float),
                ("eps0mask", C.c_uint), ("pad1", C.c_uint),
                ("seed", C.c_uint64), ("offset", C.c_uint64), ("skipahead", C.c_uint64), ("photon_offset", C.c_uint64),
                ("refine", C.c_uint), ("refine_distance", C.c_float)]


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class RefGPU:
    def __init__(self, variant="debugtag"):
        path = os.path.join(ORACLE, "_ref", "libphoxref_%s.so" % variant)
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = C.CDLL(path)
        self.lib.phoxref_last_error.restype = C.c_char_p
        self.lib.phoxref_simulate.restype = C.c_int
        self.lib.phoxref_intersect.restype = C.c_int
        self.production = bool(self.lib.phoxref_is_production())

    @staticmethod
    def _geo(fd):
        a = {k: np.ascontiguousarray(fd[k], dtype=(np.int32 if k == "solid" else np.float32)) for k in ("solid", "prim", "node", "plan", "itra", "inst")}
        args = []
        for k in ("solid", "prim", "node", "plan", "itra", "inst"):
            args += [_p(a[k]) if len(a[k]) else None, C.c_int(len(a[k]))]
        return a, args

    def simulate(self, geom, gensteps, input_photons=None, event_id=0, photon_offset=0, max_bounce=31, max_record=32,
                 tmin=0.05, tmin0=0.05, tmax=1e6, max_time=1e27, eps0mask=0x37, seed=0, offset=0, skipahead=100000, hd_factor=20, tags=False,
                 refine=0, refine_distance=5000.0):
        fd = geom["foundry"]
        keep, gargs = self._geo(fd)
        bnd = np.ascontiguousarray(geom["bnd"], dtype=np.float32)
        optical = np.ascontiguousarray(geom["optical"], dtype=np.int32)
        icdf = geom.get("icdf")
        icdf = None if icdf is None else np.ascontiguousarray(icdf, dtype=np.float32)
        gs = np.ascontiguousarray(gensteps, dtype=np.float32).reshape(-1, 6, 4)
        n = int(gs.view(np.uint32)[:, 0, 3].sum())
        ip = None if input_photons is None else np.ascontiguousarray(input_photons, dtype=np.float32)
        cfg = RefConfig(max_bounce, max_record, event_id, 0, tmin, tmin0, tmax, max_time, eps0mask, 0, seed, offset, skipahead, photon_offset,
                        refine, refine_distance)
        photon = np.zeros((n, 4, 4), dtype=np.float32)
        dbg = not self.production
        record = np.zeros((n, max_record, 4, 4), dtype=np.float32) if dbg else None
        seq = np.zeros((n, 2, 2), dtype=np.uint64) if dbg else None
        prd = np.zeros((n, max_record, 2, 4), dtype=np.float32) if dbg else None
        nray = C.c_uint64(0)
        tag = np.zeros((n, 4), dtype=np.uint64) if (tags and dbg) else None
        flat = np.zeros((n, 64), dtype=np.float32) if (tags and dbg) else None
        self.lib.phoxref_set_tag_out(_p(tag), _p(flat))
        rc = self.lib.phoxref_simulate(*gargs, _p(bnd), C.c_int(bnd.shape[0]), C.c_int(bnd.shape[3]), C.c_float(60.0), C.c_float(1.0), _p(optical),
                                       _p(icdf), C.c_int(0 if icdf is None else icdf.shape[1]), C.c_int(hd_factor),
                                       _p(gs), C.c_int(len(gs)), _p(ip), C.c_int(0 if ip is None else len(ip)), C.byref(cfg),
                                       _p(photon), _p(record), _p(seq), _p(prd), C.byref(nray))
        self.lib.phoxref_set_tag_out(None, None)
        if rc != 0:
            raise RuntimeError("phoxref_simulate: " + self.lib.phoxref_last_error().decode())
        return dict(photon=photon, record=record, seq=seq, prd=prd, nray=nray.value, tag=tag, flat=flat)

    def simtrace(self, geom, gensteps, tmin=0.05, tmax=1e6, seed=0, offset=0):
        """the reference's generate_photon_simtrace_frame + add_simtrace around the brute-force trace (FRAME gensteps)"""
        keep, gargs = self._geo(geom["foundry"])
        gs = np.ascontiguousarray(gensteps, dtype=np.float32).reshape(-1, 6, 4)
        n = int(gs.view(np.uint32)[:, 0, 3].sum())
        out = np.zeros((n, 4, 4), dtype=np.float32)
        self.lib.phoxref_simtrace.restype = C.c_int
        rc = self.lib.phoxref_simtrace(*gargs, _p(gs), C.c_int(len(gs)), C.c_float(tmin), C.c_float(tmax), C.c_uint64(seed), C.c_uint64(offset), _p(out))
        if rc != n:
            raise RuntimeError("phoxref_simtrace: " + self.lib.phoxref_last_error().decode())
        return out

    def merge(self, photons, time_window, select_mask=0):
        """sphoton::select_pred / key_functor / reduce_op of the reference, driven on the host (no GPU involved)"""
        ph = np.ascontiguousarray(photons, dtype=np.float32).reshape(-1, 4, 4)
        out = np.zeros_like(ph)
        self.lib.phoxref_merge.restype = C.c_int
        m = self.lib.phoxref_merge(_p(ph), C.c_int(len(ph)), C.c_uint(select_mask), C.c_float(time_window), _p(out))
        return out[:m].copy()

    def merge_lite(self, lite, time_window, select_mask=0):
        a = np.ascontiguousarray(lite, dtype=np.uint32).reshape(-1, 4)
        out = np.zeros_like(a)
        self.lib.phoxref_merge_lite.restype = C.c_int
        m = self.lib.phoxref_merge_lite(_p(a), C.c_int(len(a)), C.c_uint(select_mask), C.c_float(time_window), _p(out))
        return out[:m].copy()

    def intersect(self, geom, origin, direction, tmin=0.0, tmax=1e6):
        keep, gargs = self._geo(geom["foundry"])
        o = np.zeros((len(origin), 4), dtype=np.float32); o[:, :3] = origin; o[:, 3] = tmin
        d = np.zeros((len(origin), 4), dtype=np.float32); d[:, :3] = direction
        out = np.zeros((len(o), 2, 4), dtype=np.float32)
        rc = self.lib.phoxref_intersect(*gargs, _p(o), _p(d), C.c_int(len(o)), C.c_float(tmax), _p(out))
        if rc != 0:
            raise RuntimeError("phoxref_intersect: " + self.lib.phoxref_last_error().decode())
        return out


class OracleConfig(C.Structure):
    _fields_ = [("max_bounce", C.c_int), ("max_record", C.c_int), ("event_index", C.c_int), ("debug_tag", C.c_int),
                ("tmin", C.c_float), ("tmin0", C.c_float), ("tmax", C.c_float), ("max_time", C.c_float),
                ("eps0mask", C.c_uint), ("hit_mask", C.c_uint),
                ("seed", C.c_uint64), ("offset", C.c_uint64), ("skipahead", C.c_uint64), ("photon_offset", C.c_uint64),
                ("use_boxes", C.c_int), ("nthreads", C.c_int), ("refine", C.c_uint), ("refine_distance", C.c_float)]


class Oracle:
    """oracle/liboracle.so : the CPU restatement (oracle/phox_oracle.cpp)"""

    def __init__(self):
        path = os.path.join(ORACLE, "liboracle.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = C.CDLL(path)
        self.lib.oracle_simulate.restype = C.c_int
        self.lib.oracle_num_threads.restype = C.c_int

    def num_threads(self):
        return self.lib.oracle_num_threads()

    def simulate(self, geom, gensteps, input_photons=None, event_id=0, photon_offset=0, max_bounce=31, max_record=32,
                 tmin=0.05, tmin0=0.05, tmax=1e6, max_time=1e27, eps0mask=0x37, hit_mask=0x40, seed=0, offset=0, skipahead=100000,
                 hd_factor=20, debug_tag=True, use_boxes=False, nthreads=0, arrays=True, lite=False, tags=False, refine=0, refine_distance=5000.0):
        fd = geom["foundry"]
        a = {k: np.ascontiguousarray(fd[k], dtype=(np.int32 if k == "solid" else np.float32)) for k in ("solid", "prim", "node", "plan", "itra", "inst")}
        bnd = np.ascontiguousarray(geom["bnd"], dtype=np.float32)
        optical = np.ascontiguousarray(geom["optical"], dtype=np.int32)
        icdf = geom.get("icdf")
        icdf = None if icdf is None else np.ascontiguousarray(icdf, dtype=np.float32)
        gs = np.ascontiguousarray(gensteps, dtype=np.float32).reshape(-1, 6, 4)
        n = int(gs.view(np.uint32)[:, 0, 3].sum())
        ip = None if input_photons is None else np.ascontiguousarray(input_photons, dtype=np.float32)
        cfg = OracleConfig(max_bounce, max_record, event_id, 1 if debug_tag else 0, tmin, tmin0, tmax, max_time, eps0mask, hit_mask,
                           seed, offset, skipahead, photon_offset, int(use_boxes), nthreads, refine, refine_distance)
        photon = np.zeros((n, 4, 4), dtype=np.float32)
        record = np.zeros((n, max_record, 4, 4), dtype=np.float32) if arrays else None
        seq = np.zeros((n, 2, 2), dtype=np.uint64) if arrays else None
        prd = np.zeros((n, max_record, 2, 4), dtype=np.float32) if arrays else None
        nray, nhit = C.c_uint64(0), C.c_uint64(0)
        lite_arr = np.zeros((n, 4), dtype=np.uint32) if lite else None
        self.lib.oracle_set_lite_out(_p(lite_arr))
        tag = np.zeros((n, 4), dtype=np.uint64) if tags else None
        flat = np.zeros((n, 64), dtype=np.float32) if tags else None
        self.lib.oracle_set_tag_out(_p(tag), _p(flat))
        rc = self.lib.oracle_simulate(_p(a["solid"]), C.c_int(len(a["solid"])), _p(a["prim"]), C.c_int(len(a["prim"])), _p(a["node"]), C.c_int(len(a["node"])),
                                      _p(a["plan"]) if len(a["plan"]) else None, C.c_int(len(a["plan"])), _p(a["itra"]), C.c_int(len(a["itra"])),
                                      _p(a["inst"]), C.c_int(len(a["inst"])),
                                      _p(bnd), C.c_int(bnd.shape[0]), C.c_int(bnd.shape[3]), C.c_float(60.0), C.c_float(1.0), _p(optical),
                                      _p(icdf), C.c_int(0 if icdf is None else icdf.shape[1]), C.c_int(hd_factor),
                                      _p(gs), C.c_int(len(gs)), _p(ip), C.c_int(0 if ip is None else len(ip)), C.byref(cfg),
                                      _p(photon), _p(record), _p(seq), _p(prd), C.byref(nray), C.byref(nhit))
        self.lib.oracle_set_lite_out(None)
        self.lib.oracle_set_tag_out(None, None)
        if rc != 0:
            raise RuntimeError("oracle_simulate failed")
        return dict(photon=photon, record=record, seq=seq, prd=prd, nray=nray.value, nhit=nhit.value, lite=lite_arr, tag=tag, flat=flat)

    def simtrace(self, geom, gensteps, input_simtrace=None, tmin=0.05, tmax=1e6, seed=0, offset=0, use_boxes=True):
        fd = geom["foundry"]
        a = {k: np.ascontiguousarray(fd[k], dtype=(np.int32 if k == "solid" else np.float32)) for k in ("solid", "prim", "node", "plan", "itra", "inst")}
        gs = np.ascontiguousarray(gensteps, dtype=np.float32).reshape(-1, 6, 4)
        n = int(gs.view(np.uint32)[:, 0, 3].sum())
        ip = None if input_simtrace is None else np.ascontiguousarray(input_simtrace, dtype=np.float32)
        out = np.zeros((n, 4, 4), dtype=np.float32)
        self.lib.oracle_simtrace.restype = C.c_int
        rc = self.lib.oracle_simtrace(_p(a["solid"]), C.c_int(len(a["solid"])), _p(a["prim"]), _p(a["node"]), _p(a["plan"]) if len(a["plan"]) else None,
                                      _p(a["itra"]), C.c_int(len(a["itra"])), _p(a["inst"]), C.c_int(len(a["inst"])), _p(gs), C.c_int(len(gs)),
                                      _p(ip), C.c_float(tmin), C.c_float(tmax), C.c_uint64(seed), C.c_uint64(offset), _p(out), C.c_int(int(use_boxes)))
        if rc != n:
            raise RuntimeError("oracle_simtrace failed")
        return out

    def intersect(self, geom, origin, direction, tmin=0.0, tmax=1e6, use_boxes=False):
        """closest hit of each ray over all prims (quad2 per ray), CPU"""
        fd = geom["foundry"]
        a = {k: np.ascontiguousarray(fd[k], dtype=(np.int32 if k == "solid" else np.float32)) for k in ("solid", "prim", "node", "plan", "itra", "inst")}
        o = np.zeros((len(origin), 4), dtype=np.float32); o[:, :3] = origin; o[:, 3] = tmin
        d = np.zeros((len(origin), 4), dtype=np.float32); d[:, :3] = direction
        out = np.zeros((len(o), 2, 4), dtype=np.float32)
        self.lib.oracle_intersect.restype = C.c_int
        rc = self.lib.oracle_intersect(_p(a["solid"]), C.c_int(len(a["solid"])), _p(a["prim"]), _p(a["node"]), _p(a["plan"]) if len(a["plan"]) else None,
                                       _p(a["itra"]), C.c_int(len(a["itra"])), _p(a["inst"]), C.c_int(len(a["inst"])), _p(o), _p(d), C.c_int(len(o)),
                                       C.c_float(tmax), _p(out), C.c_int(int(use_boxes)))
        if rc != 0:
            raise RuntimeError("oracle_intersect failed")
        return out

    def merge_lite(self, lite, time_window, select_mask=0):
        a = np.ascontiguousarray(lite, dtype=np.uint32).reshape(-1, 4)
        out = np.zeros_like(a)
        self.lib.oracle_merge_lite.restype = C.c_int
        m = self.lib.oracle_merge_lite(_p(a), C.c_int(len(a)), C.c_uint(select_mask), C.c_float(time_window), _p(out))
        return out[:m].copy()

    def merge(self, photons, time_window, select_mask=0):
        ph = np.ascontiguousarray(photons, dtype=np.float32).reshape(-1, 4, 4)
        out = np.zeros_like(ph)
        self.lib.oracle_merge.restype = C.c_int
        m = self.lib.oracle_merge(_p(ph), C.c_int(len(ph)), C.c_uint(select_mask), C.c_float(time_window), _p(out))
        return out[:m].copy()

    def intersect_prim_batch(self, fd, prim_idx, o, d, tmin):
        node = np.ascontiguousarray(fd["node"], dtype=np.float32)
        plan = np.ascontiguousarray(fd["plan"], dtype=np.float32)
        itra = np.ascontiguousarray(fd["itra"], dtype=np.float32).copy()
        itra[:, :3, 3] = 0.0; itra[:, 3, 3] = 1.0
        node_offset = int(np.asarray(fd["prim"]).view(np.int32)[prim_idx, 0, 1])
        o = np.ascontiguousarray(o, dtype=np.float32); d = np.ascontiguousarray(d, dtype=np.float32)
        tm = np.ascontiguousarray(np.broadcast_to(np.float32(tmin), (len(o),)), dtype=np.float32)
        out = np.zeros((len(o), 4), dtype=np.float32); valid = np.zeros(len(o), dtype=np.int32)
        self.lib.oracle_intersect_prim_batch(_p(node), C.c_int(node_offset), _p(plan) if len(plan) else None, _p(itra), C.c_int(len(itra)),
                                             _p(o), _p(d), _p(tm), C.c_int(len(o)), _p(out), _p(valid))
        return out, valid.astype(bool)

    def rng_sequence(self, ni, nv, id0=0, seed=0, offset=0):
        out = np.zeros((ni, nv), dtype=np.float32)
        self.lib.oracle_rng_sequence(_p(out), C.c_int(ni), C.c_int(nv), C.c_uint64(id0), C.c_uint64(seed), C.c_uint64(offset))
        return out

    def tex2d4(self, data, xy):
        data = np.ascontiguousarray(data, dtype=np.float32); xy = np.ascontiguousarray(xy, dtype=np.float32)
        out = np.zeros((len(xy), 4), dtype=np.float32)
        self.lib.oracle_tex2d4(_p(data), C.c_int(data.shape[1]), C.c_int(data.shape[0]), _p(xy), C.c_int(len(xy)), _p(out))
        return out


class RefCSGHost:
    """oracle/_ref/libcsgref.so : the reference's CSG headers compiled for the host"""

    def __init__(self):
        path = os.path.join(ORACLE, "_ref", "libcsgref.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = C.CDLL(path)

    def intersect_prim_batch(self, fd, prim_idx, o, d, tmin):
        node = np.ascontiguousarray(fd["node"], dtype=np.float32)
        plan = np.ascontiguousarray(fd["plan"], dtype=np.float32)
        itra = np.ascontiguousarray(fd["itra"], dtype=np.float32).copy()
        itra[:, :3, 3] = 0.0; itra[:, 3, 3] = 1.0
        node_offset = int(np.asarray(fd["prim"]).view(np.int32)[prim_idx, 0, 1])
        o = np.ascontiguousarray(o, dtype=np.float32); d = np.ascontiguousarray(d, dtype=np.float32)
        tm = np.ascontiguousarray(np.broadcast_to(np.float32(tmin), (len(o),)), dtype=np.float32)
        out = np.zeros((len(o), 4), dtype=np.float32); valid = np.zeros(len(o), dtype=np.int32)
        self.lib.csgref_intersect_prim_batch(_p(node), C.c_int(node_offset), _p(plan) if len(plan) else None, _p(itra), _p(o), _p(d), _p(tm),
                                             C.c_int(len(o)), _p(out), _p(valid))
        return out, valid.astype(bool)
