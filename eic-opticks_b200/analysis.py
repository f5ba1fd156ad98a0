"""Validation outputs of SURVEY 8f rank 3: SEvt-style event directories, seqhis history tables and the
chi2 the reference's python analysis uses to call two simulations "consistent".

    save_event / load_event   NPFold-like directory  <base>/A000/{photon,record,seq,prd,hit,genstep,domain}.npy
                              (sysrap/SEvt.cc:4571 SEvt::save; component list of the event modes
                              sysrap/SEventConfig.cc:1528-1642) so an external Geant4+U4 install can be run
                              on the B side and compared.
    seqhis_table              unique seqhis values with counts, labelled "TO BT SD ..." (sysrap/sseq.h:54-184,
                              flag abbreviations sysrap/OpticksPhoton.hh)
    chi2_histories            chi2 = sum (a-b)^2/(a+b) over histories with a+b > cut (default 30), with the
                              number of degrees of freedom, as in ana/nbase.py:339-379 and ana/qcf.py:59-175
    save_rng_sequence         precooked random streams in the reference's file naming
                              rng_sequence_f_ni<NI>_nj<NJ>_nk<NK>_tranche<NI>/rng_sequence_f_ni..._ioffset<OFF>.npy
                              (sysrap/s_seq.h:30-147, qudarap/QSim.cu:43-68)
"""
import os

import numpy as np

FLAG_ABBREV = ["??", "CK", "SI", "TO", "AB", "RE", "SC", "SD", "SA", "DR", "SR", "BR", "BT", "NA", "EC", "EX", "MI"]   # index = FFS(flag)


def seqhis_label(seqhis):
    """(2,) uint64 or python int pair -> 'TO BT SD' ; 16 nibbles per 64-bit word, slot 0 in the low nibble"""
    words = [int(seqhis[0]), int(seqhis[1])] if hasattr(seqhis, "__len__") else [int(seqhis), 0]
    out = []
    for w in words:
        for k in range(16):
            nib = (w >> (4 * k)) & 0xF
            if nib == 0:
                return " ".join(out)
            out.append(FLAG_ABBREV[nib])
    return " ".join(out)


def seqhis_table(seq):
    """seq (N,2,2) uint64 -> list of (label, count, (hi, lo)) sorted by descending count"""
    seq = np.asarray(seq, dtype=np.uint64)
    his = np.ascontiguousarray(seq[:, 0, :])                      # seqhis[0], seqhis[1]
    key = his.view([("a", np.uint64), ("b", np.uint64)]).reshape(-1)
    u, c = np.unique(key, return_counts=True)
    order = np.argsort(-c, kind="stable")
    return [(seqhis_label((u[i]["a"], u[i]["b"])), int(c[i]), (int(u[i]["a"]), int(u[i]["b"]))) for i in order]


def chi2_histories(seq_a, seq_b, cut=30):
    """history-table chi2 between two simulations: bins = distinct seqhis, only bins with a+b > cut count.
    Returns (chi2, ndf, rows) with rows = [(label, a, b, contribution)]"""
    ta = {k: c for _, c, k in seqhis_table(seq_a)}
    tb = {k: c for _, c, k in seqhis_table(seq_b)}
    rows, chi2, ndf = [], 0.0, 0
    for k in sorted(set(ta) | set(tb), key=lambda kk: -(ta.get(kk, 0) + tb.get(kk, 0))):
        a, b = ta.get(k, 0), tb.get(k, 0)
        c2 = 0.0
        if a + b > cut:
            c2 = (a - b) ** 2 / float(a + b)
            chi2 += c2
            ndf += 1
        rows.append((seqhis_label(k), a, b, c2))
    return chi2, max(ndf - 1, 1), rows


TAG_NAMES = ["_", "to_sci", "to_bnd", "to_sca", "to_abs", "at_burn_sf_sd", "at_ref", "sf_burn", "sc", "to_ree", "re_wl", "re_mom_ph",
             "re_mom_ct", "re_pol_ph", "re_pol_ct", "hp_ph"]                       # sysrap/stag.h:17-35


def tag_slots(tag):
    """(n,4) uint64 stag array -> (n,64) int array of the 4-bit consumption tags, slot order (stag::get, sysrap/stag.h:344-350)"""
    tag = np.asarray(tag, dtype=np.uint64).reshape(-1, 4)
    return np.stack([(tag[:, k // 16] >> np.uint64(4 * (k % 16))) & np.uint64(0xf) for k in range(64)], axis=1).astype(np.int64)


def tag_desc(tag_row, flat_row=None):
    """one photon's consumption record as text, e.g. 'to_sci:0.1234 to_bnd:0.5678 ...' (the stagr::desc role)"""
    slots = tag_slots(np.asarray(tag_row).reshape(1, 4))[0]
    out = []
    for k, t in enumerate(slots):
        if t == 0:
            break
        out.append(TAG_NAMES[t] if flat_row is None else "%s:%.4f" % (TAG_NAMES[t], flat_row[k]))
    return " ".join(out)


def save_event(folder, index, arrays, meta=None):
    """arrays: dict name -> ndarray (photon, record, seq, prd, tag, flat, hit, genstep, inphoton ...).  Written as
    <folder>/A%03d/<name>.npy plus NPFold_index.txt, like an SEvt save directory of the EGPU ("A") event."""
    d = os.path.join(folder, "A%03d" % index)
    os.makedirs(d, exist_ok=True)
    names = []
    for name, a in arrays.items():
        if a is None:
            continue
        np.save(os.path.join(d, name + ".npy"), np.ascontiguousarray(a))
        names.append(name + ".npy")
    dom = np.zeros((2, 4, 4), dtype=np.float32)
    dom[0, 0] = (0.0, 0.0, 0.0, 1000.0)       # center_extent (sevent::init_domain defaults, sysrap/SEventConfig.cc:72-73)
    dom[0, 1, :2] = (0.0, 10.0)               # time domain
    dom[0, 2, :2] = (60.0, 820.0)             # wavelength domain (sysrap/sevent.h:111-114)
    np.save(os.path.join(d, "domain.npy"), dom)
    names.append("domain.npy")
    with open(os.path.join(d, "NPFold_index.txt"), "w") as f:
        f.write("\n".join(names) + "\n")
    if meta:
        with open(os.path.join(d, "NPFold_meta.txt"), "w") as f:
            for k, v in meta.items():
                f.write("%s:%s\n" % (k, v))
    return d


def load_event(folder, index):
    d = os.path.join(folder, "A%03d" % index)
    out = {}
    for line in open(os.path.join(d, "NPFold_index.txt")):
        n = line.strip()
        if n:
            out[n[:-4]] = np.load(os.path.join(d, n))
    return out


def save_rng_sequence(folder, uniforms, ioffset=0, nj=16, nk=16):
    """(ni, nj*nk) precooked uniforms -> the reference's rng_sequence file naming, shape (ni, nj, nk)"""
    u = np.asarray(uniforms, dtype=np.float32)
    ni = u.shape[0]
    assert u.shape[1] == nj * nk
    sub = os.path.join(folder, "rng_sequence_f_ni%d_nj%d_nk%d_tranche%d" % (ni, nj, nk, ni))
    os.makedirs(sub, exist_ok=True)
    path = os.path.join(sub, "rng_sequence_f_ni%d_nj%d_nk%d_ioffset%06d.npy" % (ni, nj, nk, ioffset))
    np.save(path, u.reshape(ni, nj, nk))
    return path


def compare_ab(record_a, record_b, atol=1e-5, shifted=True):
    """tests/compare_ab.py of the reference: photon indices whose step records differ between the GPU event (A) and the Geant4 event (B).
    With `shifted` the A records are compared from step 1 on against the B records up to the last but one (the U4Recorder B side
    has no separate generation point: compare_ab.py:11), otherwise step for step."""
    a, b = np.asarray(record_a), np.asarray(record_b)
    assert a.shape == b.shape, (a.shape, b.shape)
    if shifted:
        a, b = a[:, 1:], b[:, :-1]
    close = np.isclose(a, b, rtol=0.0, atol=atol).reshape(len(a), -1).all(axis=1)
    return [int(i) for i in np.nonzero(~close)[0]]
