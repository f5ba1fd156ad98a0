"""Boundary-property tables: bnd, optical and the scintillation ICDF.

Layouts and defaults follow the reference's standardised arrays:

    bnd     float32 (nbnd, 4, 2, 761, 4)   sysrap/sstandard.h:447-530
            species omat,osur,isur,imat ; material payload group 0 = (RINDEX, ABSLENGTH, RAYLEIGH,
            REEMISSIONPROB), group 1 = (GROUPVEL, 0, 0, 0) (sysrap/sproplist.h:32-53);
            surface payload group 0 = (detect, absorb, reflect_specular, reflect_diffuse)
            (u4/U4SurfaceArray.h:159-240); slots with no surface stay at -1
    optical int32 (nbnd*4, 4)              sysrap/sstandard.h:311-441
            .x 1-based material/surface index (0 = none), .y "ems" (sysrap/smatsur.h:8-16),
            .z finish, .w value percent
    domain  60..820 nm in 1 nm steps, energies via hc = 1239.84198433200208455673 eV nm
            (sysrap/sdomain.h:23-75); properties given against energy are sampled with
            G4PhysicsVector::Value semantics = linear interpolation, clamped at the ends
    icdf    float32 (3, 4096)              u4/U4Scint.h:406-470, qudarap/QScint.cc:84-120
            row 0 full range, row 1 the lowest 1/20, row 2 the highest 1/20 of the CDF (hd_factor 20)
"""
import numpy as np

HC_EVNM = 1239.84198433200208455673
DOMAIN_LOW, DOMAIN_HIGH, DOMAIN_STEP, DOMAIN_LENGTH = 60.0, 820.0, 1.0, 761
WAVELENGTH_NM = DOMAIN_LOW + DOMAIN_STEP * np.arange(DOMAIN_LENGTH, dtype=np.float64)
ENERGY_EV = HC_EVNM / WAVELENGTH_NM

MATERIAL_DEFAULTS = dict(RINDEX=1.0, ABSLENGTH=1e12, RAYLEIGH=1e12, REEMISSIONPROB=0.0, GROUPVEL=299.792458)
C_LIGHT = 299.792458

EMS_MATERIAL, EMS_NOSURFACE, EMS_SURFACE, EMS_SENSOR_A, EMS_CUSTOM_ART, EMS_ZMINUS = range(6)


def ems_from_name(optical_surface_name):
    """smatsur::TypeFromChar (sysrap/smatsur.h:43-56)"""
    c = optical_surface_name[:1]
    return {"": EMS_MATERIAL, "-": EMS_NOSURFACE, "@": EMS_CUSTOM_ART, "#": EMS_SENSOR_A, "!": EMS_ZMINUS}.get(c, EMS_SURFACE)


def sample(prop, default):
    """property -> values on the 761-sample wavelength domain.
    prop: None | scalar | (energy_eV ascending, values) pair"""
    if prop is None:
        return np.full(DOMAIN_LENGTH, default, dtype=np.float64)
    if np.isscalar(prop):
        return np.full(DOMAIN_LENGTH, float(prop), dtype=np.float64)
    e, v = np.asarray(prop[0], dtype=np.float64), np.asarray(prop[1], dtype=np.float64)
    return np.interp(ENERGY_EV, e, v)


def groupvel_from_rindex(energy_ev, rindex):
    """G4MaterialPropertiesTable::CalculateGROUPVEL (Geant4 11, the dependency the reference gets GROUPVEL from
    when a material gives only RINDEX): vg = c / (n + dn/dlogE); first and last points at the end energies,
    intermediate ones at bin mid-points; anything but 'normal dispersion' clamps to c/n."""
    import math
    c = C_LIGHT
    E, n = list(np.asarray(energy_ev, dtype=np.float64)), list(np.asarray(rindex, dtype=np.float64))
    if len(E) < 2:
        return np.array(E), np.array([c / n[0]])
    oe, ov = [], []
    E0, n0, E1, n1 = E[0], n[0], E[1], n[1]
    vg = c / (n0 + (n1 - n0) / math.log(E1 / E0))
    if vg < 0 or vg > c / n0:
        vg = c / n0
    oe.append(E0); ov.append(vg)
    for i in range(2, len(E)):
        vg = c / (0.5 * (n0 + n1) + (n1 - n0) / math.log(E1 / E0))
        if vg < 0 or vg > c / (0.5 * (n0 + n1)):
            vg = c / (0.5 * (n0 + n1))
        oe.append(0.5 * (E0 + E1)); ov.append(vg)
        E0, n0, E1, n1 = E1, n1, E[i], n[i]
    vg = c / (n1 + (n1 - n0) / math.log(E1 / E0))
    if vg < 0 or vg > c / n1:
        vg = c / n1
    oe.append(E1); ov.append(vg)
    return np.array(oe), np.array(ov)


class Material:
    def __init__(self, name, RINDEX=None, ABSLENGTH=None, RAYLEIGH=None, REEMISSIONPROB=None, GROUPVEL=None):
        self.name = name
        self.props = dict(RINDEX=RINDEX, ABSLENGTH=ABSLENGTH, RAYLEIGH=RAYLEIGH, REEMISSIONPROB=REEMISSIONPROB, GROUPVEL=GROUPVEL)

    @property
    def has_rindex(self):
        return self.props["RINDEX"] is not None

    def payload(self):
        out = np.zeros((2, DOMAIN_LENGTH, 4), dtype=np.float64)
        for l, key in enumerate(("RINDEX", "ABSLENGTH", "RAYLEIGH", "REEMISSIONPROB")):
            out[0, :, l] = sample(self.props[key], MATERIAL_DEFAULTS[key])
        gv = self.props["GROUPVEL"]
        if gv is None and self.props["RINDEX"] is not None and not np.isscalar(self.props["RINDEX"]):
            gv = groupvel_from_rindex(*self.props["RINDEX"])
        elif gv is None and self.props["RINDEX"] is not None:
            gv = C_LIGHT / float(self.props["RINDEX"])
        out[1, :, 0] = sample(gv, MATERIAL_DEFAULTS["GROUPVEL"])
        return out


class Surface:
    """Optical surface reduced to the four probabilities the GPU model uses.

    Either give detect/absorb/specular/diffuse directly, or REFLECTIVITY/EFFICIENCY (+ polished)
    and the rule of U4SurfaceArray::addSurface is applied: sensor (max EFFICIENCY > 0) ->
    (eff, 1-eff, 0, 0); else polished -> (0, 1-R, R, 0); else (0, 1-R, 0, R)."""

    def __init__(self, name, REFLECTIVITY=None, EFFICIENCY=None, polished=True, payload=None, optical_surface_name=None,
                 finish=0, value=1.0):
        self.name = name
        self.REFLECTIVITY, self.EFFICIENCY, self.polished, self._payload = REFLECTIVITY, EFFICIENCY, polished, payload
        self.optical_surface_name = optical_surface_name if optical_surface_name is not None else name
        self.finish, self.value = finish, value

    def payload(self):
        out = np.full((2, DOMAIN_LENGTH, 4), -1.0, dtype=np.float64)
        if self._payload is not None:
            out[0, :, :] = np.asarray(self._payload, dtype=np.float64)
            return out
        effi = sample(self.EFFICIENCY, 0.0)
        refl = sample(self.REFLECTIVITY, 0.0)
        if effi.max() > 0.0:
            out[0, :, 0], out[0, :, 1], out[0, :, 2], out[0, :, 3] = effi, 1.0 - effi, 0.0, 0.0
        elif self.polished:
            out[0, :, 0], out[0, :, 1], out[0, :, 2], out[0, :, 3] = 0.0, 1.0 - refl, refl, 0.0
        else:
            out[0, :, 0], out[0, :, 1], out[0, :, 2], out[0, :, 3] = 0.0, 1.0 - refl, 0.0, refl
        return out


def implicit_surface(name):
    """perfect absorber standing in for Geant4's RINDEX -> no-RINDEX fStopAndKill
    (u4/U4TreeBorder.h:198-250, u4/U4SurfaceArray.h addImplicit)"""
    return Surface(name, payload=(0.0, 1.0, 0.0, 0.0), optical_surface_name="X", finish=1, value=1.0)


class BoundaryTable:
    """Collects materials, surfaces and (omat,osur,isur,imat) boundaries; emits bnd + optical."""

    def __init__(self):
        self.materials, self.surfaces, self.boundaries = [], [], []

    def add_material(self, m):
        self.materials.append(m)
        return m

    def add_surface(self, s):
        self.surfaces.append(s)
        return s

    def _mat(self, name):
        return [m.name for m in self.materials].index(name)

    def _sur(self, name):
        return -1 if not name else [s.name for s in self.surfaces].index(name)

    def boundary(self, omat, osur, isur, imat):
        """index of the boundary, adding it when new (names; '' = no surface)"""
        key = (omat, osur or "", isur or "", imat)
        if key not in self.boundaries:
            self.boundaries.append(key)
        return self.boundaries.index(key)

    def material_line(self, name):
        """texture line (4*boundary + species) of a material, for genstep.matline: first boundary
        that has it as imat, else as omat (SEvt::addGenstep maps material index -> line,
        sysrap/SEvt.cc:2473-2498)"""
        for i, b in enumerate(self.boundaries):
            if b[3] == name:
                return 4 * i + 3
        for i, b in enumerate(self.boundaries):
            if b[0] == name:
                return 4 * i + 0
        raise KeyError(name)

    def names(self):
        return ["%s/%s/%s/%s" % b for b in self.boundaries]

    def arrays(self):
        nb = len(self.boundaries)
        bnd = np.full((nb, 4, 2, DOMAIN_LENGTH, 4), -1.0, dtype=np.float64)
        optical = np.zeros((nb * 4, 4), dtype=np.int32)
        mat_payload = [m.payload() for m in self.materials]
        sur_payload = [s.payload() for s in self.surfaces]
        for i, (omat, osur, isur, imat) in enumerate(self.boundaries):
            for j, name in enumerate((omat, osur, isur, imat)):
                row = optical[4 * i + j]
                if j in (0, 3):
                    k = self._mat(name)
                    bnd[i, j] = mat_payload[k]
                    row[:] = (k + 1, 0, 0, 0)
                else:
                    k = self._sur(name)
                    if k < 0:
                        row[:] = (0, EMS_NOSURFACE, 0, 0)
                    else:
                        s = self.surfaces[k]
                        bnd[i, j] = sur_payload[k]
                        row[:] = (k + 1, ems_from_name(s.optical_surface_name), s.finish, int(100.0 * s.value))
        return bnd.astype(np.float32), optical


def make_icdf(energy_ev, spectrum, nx=4096, hd_factor=20):
    """Scintillation inverse-CDF texture (3, nx), U4Scint::CreateGeant4InterpolatedInverseCDF
    (u4/U4Scint.h:406-470): the emission spectrum (vs ascending energy) is integrated with the trapezoid
    rule (G4Scintillation::BuildThePhysicsTable), and energy = integral.GetEnergy(u * max) is sampled at
    u = j/nx for the full range (row 0), j/(hd*nx) (row 1) and 1 - 1/hd + j/(hd*nx) (row 2); stored as
    wavelength = hc/energy, so wavelength DEscends with u."""
    e = np.asarray(energy_ev, dtype=np.float64)
    s = np.asarray(spectrum, dtype=np.float64)
    cdf = np.concatenate([[0.0], np.cumsum(0.5 * (s[1:] + s[:-1]) * np.diff(e))])
    mx = cdf[-1]
    j = np.arange(nx, dtype=np.float64)
    edge = 1.0 / hd_factor if hd_factor else 0.0
    out = np.zeros((3, nx), dtype=np.float64)
    us = (j / nx, j / (hd_factor * nx) if hd_factor else j / nx, 1.0 - edge + (j / (hd_factor * nx) if hd_factor else 0.0))
    for row, u in enumerate(us):
        out[row] = HC_EVNM / np.interp(u * mx, cdf, e)
    return out.astype(np.float32)
