"""Genstep and input-photon producers (host side).

    quad6 gensteps (96 B) for torch / Cerenkov / scintillation / input-photon events, field for
    field the reference structs: storch (sysrap/storch.h:43-70), scerenkov
    (sysrap/scerenkov.h:29-58), sscint (sysrap/sscint.h:34-62); gencodes sysrap/OpticksGenstep.h.

    torch_config / torch_photons restate the gphox host path: JSON torch config
    (src/config.cpp:106-146) -> host-generated photons (src/torch.cpp:8-30: ONE Philox stream
    seed,0,0 consumed sequentially, storch::generate per photon) that ride into the simulation as
    input photons (src/GPUPhotonSourceMinimal.h:57-88) on one INPUT_PHOTON genstep
    (sysrap/SEvt.cc:1057-1064).

    photons_from_text / write_hits_text are the GPUPhotonFileSource formats
    (src/GPUPhotonFileSource.h:51-87, src/GPUPhotonSourceMinimal.h:124-144).
"""
import json

import numpy as np

GS_TORCH, GS_CARRIER, GS_CERENKOV, GS_SCINTILLATION, GS_INPUT_PHOTON = 6, 14, 15, 16, 19
TORCH_TYPES = {"undef": 0, "disc": 1, "line": 2, "point": 3, "circle": 4, "rectangle": 5, "sphere_marsaglia": 6, "sphere": 7}
F_TORCH = 1 << 2


# ---- Philox4x32-10, curand conventions (host restatement used for torch photons and RNG tests) ----
def philox_block(ctr, key):
    """ctr (n,4) uint32, key (2,) -> (n,4) uint32 outputs of one Philox4x32-10 block"""
    M0, M1, W0, W1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), 0x9E3779B9, 0xBB67AE85
    c = np.array(ctr, dtype=np.uint64).reshape(-1, 4)
    k0, k1 = int(key[0]), int(key[1])
    for _ in range(10):
        p0 = M0 * c[:, 0]
        p1 = M1 * c[:, 2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & np.uint64(0xffffffff)
        hi1, lo1 = p1 >> np.uint64(32), p1 & np.uint64(0xffffffff)
        n0 = hi1 ^ c[:, 1] ^ np.uint64(k0)
        n2 = hi0 ^ c[:, 3] ^ np.uint64(k1)
        c = np.stack([n0, lo1, n2, lo0], axis=1)
        k0 = (k0 + W0) & 0xffffffff
        k1 = (k1 + W1) & 0xffffffff
    return c.astype(np.uint32)


def curand_uniform_stream(seed, subsequence, offset, n):
    """first n curand_uniform floats of the Philox stream curand_init(seed, subsequence, offset)"""
    nblk = (offset % 4 + n + 3) // 4 + 1
    b0 = offset // 4
    blk = b0 + np.arange(nblk, dtype=np.uint64)
    ctr = np.zeros((nblk, 4), dtype=np.uint64)
    ctr[:, 0] = blk & np.uint64(0xffffffff)
    ctr[:, 1] = blk >> np.uint64(32)
    ctr[:, 2] = subsequence & 0xffffffff
    ctr[:, 3] = (subsequence >> 32) & 0xffffffff
    out = philox_block(ctr, (seed & 0xffffffff, (seed >> 32) & 0xffffffff)).reshape(-1)
    u = out[offset % 4: offset % 4 + n]
    return u.astype(np.float32) * np.float32(2.3283064365386963e-10) + np.float32(2.3283064365386963e-10 / 2.0)


def curand_uniform_matrix(seed, sub0, nsub, offset, nv):
    """(nsub, nv) uniforms: row i = stream of subsequence sub0+i (the precooked-sequence layout)"""
    nblk = (offset % 4 + nv + 3) // 4 + 1
    b0 = offset // 4
    blk = (b0 + np.arange(nblk, dtype=np.uint64))
    sub = (sub0 + np.arange(nsub, dtype=np.uint64))
    ctr = np.zeros((nsub, nblk, 4), dtype=np.uint64)
    ctr[:, :, 0] = (blk & np.uint64(0xffffffff))[None, :]
    ctr[:, :, 1] = (blk >> np.uint64(32))[None, :]
    ctr[:, :, 2] = (sub & np.uint64(0xffffffff))[:, None]
    ctr[:, :, 3] = (sub >> np.uint64(32))[:, None]
    out = philox_block(ctr.reshape(-1, 4), (seed & 0xffffffff, (seed >> 32) & 0xffffffff)).reshape(nsub, nblk * 4)
    u = out[:, offset % 4: offset % 4 + nv]
    return u.astype(np.float32) * np.float32(2.3283064365386963e-10) + np.float32(2.3283064365386963e-10 / 2.0)


# ---- gensteps -------------------------------------------------------------------------------------
def empty_gensteps(n):
    return np.zeros((n, 6, 4), dtype=np.float32)


def _normalize_f32(v):
    v = np.asarray(v, dtype=np.float32)
    inv = np.float32(1.0) / np.sqrt(np.float32(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]), dtype=np.float32)
    return (v * inv).astype(np.float32)


def torch_config(path_or_dict):
    """parse a gphox JSON config -> (torch dict with normalised mom, event dict)"""
    cfg = path_or_dict
    if isinstance(cfg, str):
        with open(cfg) as f:
            cfg = json.load(f)
    t = dict(cfg["torch"])
    t["mom"] = _normalize_f32(t["mom"])
    return t, dict(cfg.get("event", {}))


def torch_genstep(t, numphoton=None):
    """one storch genstep (1,6,4)"""
    gs = empty_gensteps(1)
    u = gs.view(np.uint32)
    u[0, 0] = (GS_TORCH, t.get("trackid", 0), t.get("matline", 0), numphoton if numphoton is not None else t["numphoton"])
    gs[0, 1, :3] = t["pos"]; gs[0, 1, 3] = t.get("time", 0.0)
    gs[0, 2, :3] = t["mom"]; gs[0, 2, 3] = t.get("weight", 0.0)
    gs[0, 3, :3] = t["pol"]; gs[0, 3, 3] = t["wavelength"]
    gs[0, 4, :2] = t.get("zenith", (0.0, 1.0)); gs[0, 4, 2:] = t.get("azimuth", (0.0, 1.0))
    gs[0, 5, 0] = t["radius"]; gs[0, 5, 1] = t.get("distance", 0.0)
    u[0, 5, 2] = t.get("mode", 255)
    ty = t.get("type", "disc")
    u[0, 5, 3] = TORCH_TYPES[ty] if isinstance(ty, str) else ty
    return gs


def _rotate_uz(d, u):
    """smath::rotateUz on (n,3) float32 arrays d with one unit vector u"""
    u = np.asarray(u, dtype=np.float32)
    up = np.float32(u[0] * u[0] + u[1] * u[1])
    d = d.astype(np.float32)
    if up > 0:
        up = np.sqrt(up, dtype=np.float32)
        px, py, pz = d[:, 0].copy(), d[:, 1].copy(), d[:, 2].copy()
        d[:, 0] = (u[0] * u[2] * px - u[1] * py) / up + u[0] * pz
        d[:, 1] = (u[1] * u[2] * px + u[0] * py) / up + u[1] * pz
        d[:, 2] = -up * px + u[2] * pz
    elif u[2] < 0:
        d[:, 0] = -d[:, 0]
        d[:, 2] = -d[:, 2]
    return d


def torch_photons(t, num_photons=0, seed=0, _offset=0):
    """generate_photons (src/torch.cpp:8-30) for the disc type the shipped configs use:
    one Philox stream (seed, subsequence 0, offset 0), two uniforms per photon."""
    n = num_photons or t["numphoton"]
    ty = t.get("type", "disc")
    ty = TORCH_TYPES[ty] if isinstance(ty, str) else ty
    if ty != TORCH_TYPES["disc"]:
        raise NotImplementedError("host torch generation covers the disc type of the shipped configs")
    f = np.float32
    chunk = 8_000_000
    if n > chunk:           # big arrays in pieces of the same stream (offset = 2 uniforms per photon): same numbers, bounded scratch
        out = np.empty((n, 4, 4), dtype=f)
        for s0 in range(0, n, chunk):
            out[s0:s0 + chunk] = torch_photons(t, min(chunk, n - s0), seed, _offset=2 * s0)
        return out
    uu = curand_uniform_stream(seed, 0, _offset, 2 * n).reshape(n, 2)
    zen, azi = t.get("zenith", (0.0, 1.0)), t.get("azimuth", (0.0, 1.0))
    u_zenith = f(zen[0]) + uu[:, 0] * f(zen[1] - zen[0])
    u_azimuth = f(azi[0]) + uu[:, 1] * f(azi[1] - azi[0])
    r = f(t["radius"]) * u_zenith
    phi = f(2.0) * f(np.pi) * u_azimuth
    sinPhi, cosPhi = np.sin(phi, dtype=f), np.cos(phi, dtype=f)
    mom = np.asarray(t["mom"], dtype=f)
    pos = np.stack([r * cosPhi, r * sinPhi, np.zeros(n, dtype=f)], axis=1)
    pos = _rotate_uz(pos, mom) + np.asarray(t["pos"], dtype=f)
    pol = np.stack([sinPhi, -cosPhi, np.zeros(n, dtype=f)], axis=1)
    pol = _rotate_uz(pol, mom)
    ph = np.zeros((n, 4, 4), dtype=f)
    ph[:, 0, :3] = pos; ph[:, 0, 3] = t.get("time", 0.0)
    ph[:, 1, :3] = mom
    ph[:, 2, :3] = pol; ph[:, 2, 3] = t["wavelength"]
    pu = ph.view(np.uint32)
    pu[:, 3, 0] = F_TORCH      # zero_flags(); set_flag(TORCH)
    pu[:, 3, 3] = F_TORCH
    return ph


def input_photon_genstep(n):
    """the single genstep input photons ride on (SEvt::addInputGenstep, sysrap/SEvt.cc:1057-1064)"""
    gs = empty_gensteps(1)
    u = gs.view(np.uint32)
    u[0, 0, 0] = GS_INPUT_PHOTON
    u[0, 0, 3] = n
    return gs


def cerenkov_gensteps(pos, direction, step_length, numphoton, matline, beta_inverse, wmin, wmax, n_max, time0=0.0,
                      pre_velocity=299.792458, post_velocity=299.792458, mean_photons=(2.0, 2.0), charge=-1.0):
    """scerenkov gensteps with the field semantics of U4::CollectGenstep_G4Cerenkov_modified
    (u4/U4.cc:196-252).  pos (n,3) start points, direction unit vector(s), numphoton (n,)."""
    pos = np.atleast_2d(np.asarray(pos, dtype=np.float32))
    n = len(pos)
    gs = empty_gensteps(n)
    u = gs.view(np.uint32)
    i = gs.view(np.int32)
    u[:, 0, 0] = GS_CERENKOV
    u[:, 0, 2] = matline
    u[:, 0, 3] = numphoton
    gs[:, 1, :3] = pos
    gs[:, 1, 3] = time0
    gs[:, 2, :3] = np.asarray(direction, dtype=np.float32) * np.float32(step_length)      # DeltaPosition, not normalised
    gs[:, 2, 3] = step_length
    i[:, 3, 0] = 11                                  # pdg code of the parent (e-)
    gs[:, 3, 1] = charge
    gs[:, 3, 2] = 1.0                                # weight
    gs[:, 3, 3] = pre_velocity
    max_cos = beta_inverse / n_max
    gs[:, 4, 0] = beta_inverse
    gs[:, 4, 1] = wmin
    gs[:, 4, 2] = wmax
    gs[:, 4, 3] = max_cos
    gs[:, 5, 0] = (1.0 - max_cos) * (1.0 + max_cos)  # maxSin2
    gs[:, 5, 1] = mean_photons[0]
    gs[:, 5, 2] = mean_photons[1]
    gs[:, 5, 3] = post_velocity
    return gs


def scint_gensteps(pos, direction, step_length, numphoton, matline, scintillation_time, time0=0.0, mean_velocity=299.792458,
                   charge=-1.0):
    """sscint gensteps with the field semantics of U4::CollectGenstep_DsG4Scintillation_r4695
    (u4/U4.cc:82-143)"""
    pos = np.atleast_2d(np.asarray(pos, dtype=np.float32))
    n = len(pos)
    gs = empty_gensteps(n)
    u = gs.view(np.uint32)
    i = gs.view(np.int32)
    u[:, 0, 0] = GS_SCINTILLATION
    u[:, 0, 2] = matline
    u[:, 0, 3] = numphoton
    gs[:, 1, :3] = pos
    gs[:, 1, 3] = time0
    gs[:, 2, :3] = np.asarray(direction, dtype=np.float32) * np.float32(step_length)
    gs[:, 2, 3] = step_length
    i[:, 3, 0] = 11
    gs[:, 3, 1] = charge
    gs[:, 3, 2] = 1.0
    gs[:, 3, 3] = mean_velocity
    i[:, 4, 0] = 1                                   # scnt
    gs[:, 5, 0] = scintillation_time
    return gs


# ---- simtrace gensteps (sysrap/SFrameGenstep.cc, SGenstep.h) -----------------------------------------------------------
GS_FRAME, GS_INPUT_PHOTON_SIMTRACE = 17, 20        # OpticksGenstep.h:38,41
AX_XYZ, AX_YZ, AX_XZ, AX_XY = 0, 1, 2, 3           # sxyz.h:3


def grid_axes(nx, ny, nz):
    """SGenstep::GridAxes (SGenstep.h:137-153): which plane a planar grid lies in, XYZ for anything else"""
    if nx == 0 and ny > 0 and nz > 0:
        return AX_YZ
    if nx > 0 and ny == 0 and nz > 0:
        return AX_XZ
    if nx > 0 and ny > 0 and nz == 0:
        return AX_XY
    return AX_XYZ


def standardize_cegs(cegs):
    """SFrameGenstep::StandardizeCEGS: nx:ny:nz:n or nx:ny:nz:dx:dy:dz:n -> ix0:ix1:iy0:iy1:iz0:iz1:n:high"""
    c = [int(v) for v in cegs]
    if len(c) == 4:
        nx, ny, nz, n = c
        return [-nx, nx, -ny, ny, -nz, nz, n, 1]
    if len(c) == 7:
        nx, ny, nz, dx, dy, dz, n = c
        return [-nx + dx, nx + dx, -ny + dy, ny + dy, -nz + dz, nz + dz, n, 1]
    if len(c) == 8:
        return c
    raise ValueError("cegs must have 4, 7 or 8 integers")


def frame_gensteps(ce, cegs, gridscale=1.0, geotran=None, ce_offset=((0.0, 0.0, 0.0),), ce_scale=True, radial_range=None):
    """SFrameGenstep::MakeCenterExtentGenstep (SFrameGenstep.cc:604-735): a grid of FRAME gensteps around a
    center-extent `ce` = (cx, cy, cz, extent).  Grid point (ix,iy,iz) sits at local offset i*gridscale*extent; each
    genstep carries the transform  translate(offset) . geotran  (row-vector convention, so the small local shift is
    applied first) in q2..q5, gridaxes in q0.y, the packed signed-char id (ix,iy,iz,plane) in q0.z and
    photons_per_genstep in q0.w.  geotran None = translation to ce.xyz (the usual frame of a target volume)."""
    c = standardize_cegs(cegs)
    high = c[7]
    assert 1 <= high <= 8
    scale = float(gridscale) / float(high)
    ix0, ix1, iy0, iy1, iz0, iz1 = [v * high for v in c[:6]]
    per = c[6]
    axes = grid_axes((ix1 - ix0) // 2, (iy1 - iy0) // 2, (iz1 - iz0) // 2)
    local_scale = scale * float(ce[3]) if ce_scale else scale
    if geotran is None:
        geotran = np.eye(4, dtype=np.float64)
        geotran[3, :3] = ce[:3]
    geotran = np.asarray(geotran, dtype=np.float64).reshape(4, 4)
    rmin, rmax = (radial_range if radial_range is not None else (0.0, np.inf))
    out = []
    for ip, off in enumerate(ce_offset):
        for ix in range(ix0, ix1 + 1):
            for iy in range(iy0, iy1 + 1):
                for iz in range(iz0, iz1 + 1):
                    t = np.array([ix, iy, iz], dtype=np.float64) * local_scale
                    if radial_range is not None and not (rmin <= np.sqrt((t * t).sum()) <= rmax):
                        continue
                    shift = np.eye(4, dtype=np.float64)
                    shift[3, :3] = t
                    m = (shift @ geotran).astype(np.float32)
                    g = np.zeros((6, 4), dtype=np.float32)
                    gi = g.view(np.int32)
                    gi[0, 0] = GS_FRAME
                    gi[0, 1] = axes
                    g.view(np.uint32)[0, 2] = np.array([ix, iy, iz, ip], dtype=np.int8).view(np.uint32)[0]      # SGenstep::GenstepID
                    gi[0, 3] = per
                    g[1] = (off[0], off[1], off[2], 1.0)
                    g[2:6] = m
                    out.append(g)
    return np.stack(out) if out else np.zeros((0, 6, 4), dtype=np.float32)


def input_simtrace_genstep(n):
    """one INPUT_PHOTON_SIMTRACE genstep carrying n caller-supplied rays (qsim.h:2455)"""
    g = np.zeros((1, 6, 4), dtype=np.float32)
    g.view(np.int32)[0, 0, 0] = GS_INPUT_PHOTON_SIMTRACE
    g.view(np.uint32)[0, 0, 3] = n
    return g


def partition_gensteps(gs, nrank):
    """Contiguous genstep ranges balanced by photon count, one per rank, with the absolute photon
    offset of each range - the concurrent form of SGenstep::GetGenstepSlices
    (sysrap/SGenstep.h:249-323).  Returns [(gs_start, gs_stop, photon_offset, photon_count)]."""
    num = gs.view(np.uint32)[:, 0, 3].astype(np.int64)
    total = int(num.sum())
    csum = np.concatenate([[0], np.cumsum(num)])
    out = []
    start = 0
    for r in range(nrank):
        target = total * (r + 1) // nrank
        stop = int(np.searchsorted(csum, target, side="left")) if r < nrank - 1 else len(num)
        stop = max(stop, start)
        stop = min(stop, len(num))
        out.append((start, stop, int(csum[start]), int(csum[stop] - csum[start])))
        start = stop
    return out


# ---- text formats -----------------------------------------------------------------------------------
def photons_from_text(path):
    """GPUPhotonFileSource input: 11 floats per line, '#' comments and blank lines skipped,
    malformed lines skipped with a warning (src/GPUPhotonFileSource.h:51-87). Flags stay zero."""
    rows = []
    with open(path) as f:
        for lineno, line in enumerate(f, 1):
            line = line.rstrip("\n")
            if not line or line[0] == "#":
                continue
            tok = line.split()
            try:
                vals = [float(x) for x in tok[:11]]
                if len(vals) < 11:
                    raise ValueError
            except ValueError:
                print("WARNING: skipping malformed line %d: %s" % (lineno, line))
                continue
            rows.append(vals)
    ph = np.zeros((len(rows), 4, 4), dtype=np.float32)
    for k, v in enumerate(rows):
        ph[k, 0] = v[0:4]
        ph[k, 1, :3] = v[4:7]
        ph[k, 2, :3] = v[7:10]
        ph[k, 2, 3] = v[10]
    return ph


def write_hits_text(hits, path, with_process=False):
    """opticks_hits_output.txt: `time wavelength  (x, y, z)  (mx, my, mz)  (px, py, pz)`
    (src/GPUPhotonSourceMinimal.h:124-144; CreationProcessID from flagmask, src/GPUCerenkov.h:388-391)"""
    fm = hits.view(np.uint32)[:, 3, 3]
    with open(path, "w") as f:
        for k, h in enumerate(hits):
            line = "%g %g  (%g, %g, %g)  (%g, %g, %g)  (%g, %g, %g)" % (h[0, 3], h[2, 3], h[0, 0], h[0, 1], h[0, 2], h[1, 0], h[1, 1],
                                                                     h[1, 2], h[2, 0], h[2, 1], h[2, 2])
            if with_process:
                pid = 0 if fm[k] & 1 else (1 if fm[k] & 2 else -1)
                line += "  CreationProcessID=%d" % pid
            f.write(line + "\n")
