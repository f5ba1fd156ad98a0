"""Host-side mirror of the reference's simulate interface over the C ABI.

    Simulator            <-> CSGOptiX : SSimulator   (sysrap/SSimulator.h:16-35, CSGOptiX/CSGOptiX.h:59)
        Create(foundry)       CSGOptiX::Create(CSGFoundry*)                  CSGOptiX.cc:367
        simulate(eventID)     double simulate(int eventID, bool reset)       CSGOptiX.cc:798-803
        simulate_np(gs, id)   NP* simulate(const NP* gs, int eventID)        CSGOptiX.cc:823-826
        reset(eventID)        void reset(int eventID)
        desc()                const char* desc()
    Event                <-> the slice of SEvt the simulate path uses       (sysrap/SEvt.cc)
        add_genstep           SEvt::AddGenstep                               :2440-2548
        set_input_photon      SEvt::SetInputPhoton                           :2059
        get_num_hit / get_hit SEvt::GetNumHit / getHit                       :4924-4925, 4991

Every call goes through libphox.so; nothing here computes physics.
"""
import ctypes as C

import numpy as np

from . import lib as L
from .gensteps import input_photon_genstep


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class Event:
    """genstep / input-photon collection for one event (the EGPU SEvt role)"""

    def __init__(self):
        self.gensteps = []
        self.input_photon = None
        self.hits = None
        self.index = 0

    def add_genstep(self, gs):
        gs = np.ascontiguousarray(gs, dtype=np.float32).reshape(-1, 6, 4)
        self.gensteps.append(gs)

    def set_input_photon(self, photons):
        self.input_photon = np.ascontiguousarray(photons, dtype=np.float32).reshape(-1, 4, 4)

    def genstep_array(self):
        if self.input_photon is not None:
            return input_photon_genstep(len(self.input_photon))
        if not self.gensteps:
            return None
        return np.ascontiguousarray(np.concatenate(self.gensteps, axis=0))

    def get_num_hit(self):
        return 0 if self.hits is None else len(self.hits)

    def get_hit(self, idx):
        return self.hits[idx]

    def clear(self):
        self.gensteps, self.input_photon, self.hits = [], None, None


class Simulator:
    def __init__(self, device=0):
        self.lib = L.load()
        self.ctx = self.lib.phox_create(device)
        if not self.ctx:
            raise L.PhoxError(-5, self.lib.phox_last_error(None).decode())
        self.cfg = L.Config()
        self.lib.phox_default_config(C.byref(self.cfg))
        self.event = Event()
        self._keep = []

    # ---- lifecycle -------------------------------------------------------------------------
    @classmethod
    def Create(cls, foundry, bnd, optical, icdf=None, hd_factor=20, device=0, domain=(60.0, 1.0), **config):
        sim = cls(device)
        sim.set_geometry(foundry)
        sim.set_tables(bnd, optical, icdf, hd_factor, domain)
        if config:
            sim.set_config(**config)
        return sim

    def close(self):
        if self.ctx:
            self.lib.phox_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc < 0:
            raise L.PhoxError(rc, self.lib.phox_last_error(self.ctx).decode())
        return rc

    def desc(self):
        return self.lib.phox_desc(self.ctx).decode()

    # ---- setup -----------------------------------------------------------------------------
    def set_geometry(self, fd):
        arr = {}
        for k, dt in (("solid", np.int32), ("prim", np.float32), ("node", np.float32), ("plan", np.float32), ("itra", np.float32),
                      ("inst", np.float32)):
            arr[k] = np.ascontiguousarray(fd[k], dtype=dt)
        n = {k: len(arr[k]) for k in arr}
        self._check(self.lib.phox_set_geometry(self.ctx, _ptr(arr["solid"]), n["solid"], _ptr(arr["prim"]), n["prim"], _ptr(arr["node"]),
                                               n["node"], _ptr(arr["plan"]) if n["plan"] else None, n["plan"],
                                               _ptr(arr["itra"]) if n["itra"] else None, n["itra"], _ptr(arr["inst"]), n["inst"]))

    def set_tables(self, bnd, optical, icdf=None, hd_factor=20, domain=(60.0, 1.0)):
        bnd = np.ascontiguousarray(bnd, dtype=np.float32)
        assert bnd.ndim == 5 and bnd.shape[1:3] == (4, 2) and bnd.shape[4] == 4, bnd.shape
        optical = np.ascontiguousarray(optical, dtype=np.int32).reshape(-1, 4)
        assert len(optical) == 4 * bnd.shape[0]
        if icdf is not None:
            icdf = np.ascontiguousarray(icdf, dtype=np.float32).reshape(3, -1)
        self._check(self.lib.phox_set_tables(self.ctx, _ptr(bnd), bnd.shape[0], bnd.shape[3], domain[0], domain[1], _ptr(optical),
                                             _ptr(icdf) if icdf is not None else None, 3 if icdf is not None else 0,
                                             icdf.shape[1] if icdf is not None else 0, hd_factor))

    def set_stream(self, cuda_stream_ptr):
        """run on the caller's CUDA stream (e.g. torch.cuda.current_stream().cuda_stream); 0 restores the own stream"""
        self._check(self.lib.phox_set_stream(self.ctx, C.c_void_p(cuda_stream_ptr) if cuda_stream_ptr else None))

    def set_config(self, **kw):
        for k, v in kw.items():
            if not hasattr(self.cfg, k):
                raise AttributeError("phox_config has no field %r" % k)
            setattr(self.cfg, k, v)
        self._check(self.lib.phox_set_config(self.ctx, C.byref(self.cfg)))

    # ---- SSimulator high-level API ---------------------------------------------------------------
    def simulate(self, event_id=0, reset=False, photon_offset=0):
        """double simulate(int eventID, bool reset): runs the gensteps / input photons collected in
        self.event, returns launch seconds (-1. when there is nothing to simulate, QSim.cc:446)."""
        gs = self.event.genstep_array()
        if gs is None:
            return -1.0
        hits = self.simulate_np(gs, event_id, self.event.input_photon, photon_offset)
        self.event.hits = hits
        dt = self.last_launch_seconds
        if reset:
            self.reset(event_id)
        return dt

    def simulate_np(self, gensteps, event_id=0, input_photons=None, photon_offset=0):
        """NP* simulate(const NP* gs, int eventID): gensteps in, copy of the hit array out"""
        gs = np.ascontiguousarray(gensteps, dtype=np.float32).reshape(-1, 6, 4)
        ip = None
        if input_photons is not None:
            ip = np.ascontiguousarray(input_photons, dtype=np.float32).reshape(-1, 4, 4)
        dt = C.c_double(0.0)
        self._check(self.lib.phox_simulate(self.ctx, _ptr(gs), len(gs), _ptr(ip) if ip is not None else None,
                                           len(ip) if ip is not None else 0, event_id, photon_offset, C.byref(dt)))
        self.last_launch_seconds = dt.value
        return self.get_hits()

    def simulate_np_into(self, gensteps, event_id, input_photons, photon_offset, out):
        """simulate_np writing the hits into a caller buffer (e.g. pinned memory); returns the hit count"""
        dt = C.c_double(0.0)
        ip = input_photons
        self._check(self.lib.phox_simulate(self.ctx, _ptr(gensteps), len(gensteps), _ptr(ip) if ip is not None else None,
                                           len(ip) if ip is not None else 0, event_id, photon_offset, C.byref(dt)))
        self.last_launch_seconds = dt.value
        n = self.num_hit()
        assert n <= len(out)
        if n:
            self._check(self.lib.phox_get_hits(self.ctx, _ptr(out)))
        return n

    def get_hits_device(self, d_dst_ptr):
        """copy the hit records into device memory owned by the caller (async on the context's stream)"""
        self._check(self.lib.phox_get_hits_device(self.ctx, C.c_void_p(d_dst_ptr)))

    def simulate_device(self, d_genstep_ptr, ngs, d_input_ptr=0, ninput=0, event_id=0, photon_offset=0):
        """device-resident variant: pointers are raw CUDA addresses (e.g. torch.Tensor.data_ptr())"""
        dt = C.c_double(0.0)
        self._check(self.lib.phox_simulate_device(self.ctx, C.c_void_p(d_genstep_ptr), ngs, C.c_void_p(d_input_ptr) if d_input_ptr else None,
                                                  ninput, event_id, photon_offset, C.byref(dt)))
        self.last_launch_seconds = dt.value
        return dt.value

    def reset(self, event_id=0):
        self.lib.phox_reset(self.ctx)
        self.event.clear()

    # ---- results -------------------------------------------------------------------------------
    def num_hit(self):
        return self.lib.phox_num_hit(self.ctx)

    def num_photon(self):
        return self.lib.phox_num_photon(self.ctx)

    def hits_device_ptr(self):
        return self.lib.phox_hits_device(self.ctx) or 0

    def get_hits(self, out=None):
        n = self.num_hit()
        if out is None:
            out = np.empty((n, 4, 4), dtype=np.float32)
        if n:
            self._check(self.lib.phox_get_hits(self.ctx, _ptr(out)))
        return out[:n]

    def get_array(self, name):
        nbytes = self._check(self.lib.phox_get_array(self.ctx, name.encode(), None, 0))
        n = self.num_photon()
        if name in ("photon",):
            out = np.empty((nbytes // 64, 4, 4), dtype=np.float32)
        elif name == "record":
            out = np.empty((n, nbytes // 64 // max(n, 1), 4, 4), dtype=np.float32)
        elif name == "seq":
            out = np.empty((nbytes // 32, 2, 2), dtype=np.uint64)
        elif name == "prd":
            out = np.empty((n, nbytes // 32 // max(n, 1), 2, 4), dtype=np.float32)
        elif name == "hit":
            out = np.empty((nbytes // 64, 4, 4), dtype=np.float32)
        elif name == "tag":                                     # stag: 4 x u64 per photon, 16 4-bit tags each
            out = np.empty((nbytes // 32, 4), dtype=np.uint64)
        elif name == "flat":                                    # sflat: the first 64 tagged uniforms per photon
            out = np.empty((nbytes // 256, 64), dtype=np.float32)
        else:
            raise KeyError(name)
        if nbytes:
            self._check(self.lib.phox_get_array(self.ctx, name.encode(), _ptr(out), out.nbytes))
        return out

    def set_profiling(self, on=True):
        """per-kernel CUDA-event timing of the bounce loop (stats trace_/propagate_kernel_seconds)"""
        self._check(self.lib.phox_set_profiling(self.ctx, 1 if on else 0))

    def stats(self):
        st = L.Stats()
        self._check(self.lib.phox_get_stats(self.ctx, C.byref(st)))
        return {f: getattr(st, f) for f, _ in st._fields_}

    # ---- geometry queries / rng ---------------------------------------------------------------
    def intersect(self, origin, direction, tmin=0.0, accel=L.ACCEL_BVH):
        o = np.zeros((len(origin), 4), dtype=np.float32)
        o[:, :3] = origin
        o[:, 3] = tmin
        d = np.zeros((len(direction), 4), dtype=np.float32)
        d[:, :3] = direction
        out = np.empty((len(o), 2, 4), dtype=np.float32)
        self._check(self.lib.phox_intersect(self.ctx, _ptr(o), _ptr(d), len(o), _ptr(out), accel))
        return out

    def boundary_lookup(self, nm, line, k):
        """hardware-texture readback of the boundary table: (n,) wavelengths, lines, payload groups -> (n,4)"""
        nm = np.ascontiguousarray(nm, dtype=np.float32)
        line = np.ascontiguousarray(line, dtype=np.uint32)
        k = np.ascontiguousarray(k, dtype=np.uint32)
        out = np.empty((len(nm), 4), dtype=np.float32)
        self._check(self.lib.phox_boundary_lookup(self.ctx, _ptr(nm), _ptr(line), _ptr(k), len(nm), _ptr(out)))
        return out

    def simtrace(self, gensteps, input_simtrace=None):
        """SSimulator::simtrace: FRAME / INPUT_PHOTON_SIMTRACE gensteps -> (n,4,4) simtrace records
        (sevent::add_simtrace layout, sysrap/sevent.h:670-697)."""
        gs = np.ascontiguousarray(gensteps, dtype=np.float32).reshape(-1, 6, 4)
        ip = None if input_simtrace is None else np.ascontiguousarray(input_simtrace, dtype=np.float32).reshape(-1, 4, 4)
        n = self.lib.phox_simtrace(self.ctx, _ptr(gs), len(gs), _ptr(ip) if ip is not None else None, 0 if ip is None else len(ip), None, 0)
        if n < 0:
            self._check(int(n))
        out = np.empty((n, 4, 4), dtype=np.float32)
        m = self.lib.phox_simtrace(self.ctx, _ptr(gs), len(gs), _ptr(ip) if ip is not None else None, 0 if ip is None else len(ip), _ptr(out), n)
        if m < 0:
            self._check(int(m))
        return out

    def merge_hits(self, time_window):
        """hits of the current event merged per (identity, time bucket) on the device (QEvt::PerLaunchMerge role)"""
        m = self.lib.phox_merge_hits(self.ctx, time_window, None, 0)
        if m < 0:
            self._check(int(m))
        out = np.empty((m, 4, 4), dtype=np.float32)
        if m:
            r = self.lib.phox_merge_hits(self.ctx, time_window, _ptr(out), m)
            if r < 0:
                self._check(int(r))
        return out

    def get_hits_lite(self):
        """(num_hit, 4) uint32 view of the sphotonlite hits (mode_lite = 1): [hitcount<<16|identity, time bits,
        lposcost<<16|lposfphi, flagmask]"""
        n = self.num_hit()
        out = np.empty((n, 4), dtype=np.uint32)
        self._check(self.lib.phox_get_hits_lite(self.ctx, _ptr(out) if n else None))
        return out

    def merge_hits_lite(self, time_window):
        m = self.lib.phox_merge_hits_lite(self.ctx, time_window, None, 0)
        if m < 0:
            self._check(int(m))
        out = np.empty((m, 4), dtype=np.uint32)
        if m:
            r = self.lib.phox_merge_hits_lite(self.ctx, time_window, _ptr(out), m)
            if r < 0:
                self._check(int(r))
        return out

    def merge(self, photons, time_window, select_mask=0):
        """merge any (n,4,4) sphoton array (QEvt::FinalMerge role: concatenated per-launch / per-rank results)"""
        ph = np.ascontiguousarray(photons, dtype=np.float32).reshape(-1, 4, 4)
        out = np.empty_like(ph)
        m = self.lib.phox_merge(self.ctx, _ptr(ph), len(ph), select_mask, time_window, _ptr(out), len(out))
        if m < 0:
            self._check(int(m))
        return out[:m].copy()

    def rng_sequence(self, ni, nv, id0=0, event_id=0):
        out = np.empty((ni, nv), dtype=np.float32)
        self._check(self.lib.phox_rng_sequence(self.ctx, _ptr(out), ni, nv, id0, event_id))
        return out
