"""history tables / chi2 / event directories on CPU-oracle events."""
import numpy as np

from eic_opticks_b200 import analysis as A, workloads
from _ref import Oracle


def test_seqhis_labels_and_table():
    seq = np.zeros((5, 2, 2), dtype=np.uint64)
    seq[:3, 0, 0] = 0x7c3          # TO BT SD
    seq[3:, 0, 0] = 0x83           # TO SA
    t = A.seqhis_table(seq)
    assert t[0][:2] == ("TO BT SD", 3) and t[1][:2] == ("TO SA", 2)
    s17 = (0x3 | sum(0xA << (4 * k) for k in range(1, 16)), 0xA)    # 17 steps spill into the second word
    assert A.seqhis_label(s17).split() == ["TO"] + ["SR"] * 16


def test_chi2_consistent_between_rng_modes_and_inconsistent_when_physics_changes(tmp_path):
    orc = Oracle()
    w = workloads.scintillator_tank(num_photon=40000, photons_per_genstep=200)
    a = orc.simulate(w["geom"], w["gensteps"], debug_tag=True)
    b = orc.simulate(w["geom"], w["gensteps"], debug_tag=False)          # different random streams, same physics
    chi2, ndf, rows = A.chi2_histories(a["seq"], b["seq"])
    assert ndf > 10 and chi2 / ndf < 2.0, (chi2, ndf)
    # change the physics (truncate histories at 2 bounces): the table must disagree
    c = orc.simulate(w["geom"], w["gensteps"], debug_tag=True, max_bounce=2)
    chi2c, ndfc, _ = A.chi2_histories(a["seq"], c["seq"])
    assert chi2c / ndfc > 10.0
    # the same event compared with itself
    assert A.chi2_histories(a["seq"], a["seq"])[0] == 0.0
    d = A.save_event(str(tmp_path), 0, dict(photon=a["photon"], seq=a["seq"], record=a["record"], genstep=w["gensteps"]), meta=dict(rng="DEBUG_TAG"))
    back = A.load_event(str(tmp_path), 0)
    assert (back["seq"] == a["seq"]).all() and back["photon"].shape == (40000, 4, 4) and back["domain"].shape == (2, 4, 4)
    assert d.endswith("A000")


def test_rng_sequence_file_naming(tmp_path):
    u = Oracle().rng_sequence(100, 256, 0)
    p = A.save_rng_sequence(str(tmp_path), u)
    assert p.endswith("rng_sequence_f_ni100_nj16_nk16_tranche100/rng_sequence_f_ni100_nj16_nk16_ioffset000000.npy")
    assert np.load(p).shape == (100, 16, 16)


def test_tag_decoding():
    from eic_opticks_b200 import analysis as A
    tag = np.zeros((1, 4), dtype=np.uint64)
    seq = [1, 2, 3, 4, 5, 7] + [1, 2, 3, 4, 5, 6] * 2 + [8] * 5           # 23 slots: spills into the second u64
    for k, t in enumerate(seq):
        tag[0, k // 16] |= np.uint64(t) << np.uint64(4 * (k % 16))
    s = A.tag_slots(tag)
    assert s.shape == (1, 64) and s[0, :len(seq)].tolist() == seq and (s[0, len(seq):] == 0).all()
    txt = A.tag_desc(tag[0])
    assert txt.startswith("to_sci to_bnd to_sca to_abs at_burn_sf_sd sf_burn") and txt.endswith("sc sc sc sc sc")
    assert A.tag_desc(tag[0], np.linspace(0, 1, 64)).split()[1] == "to_bnd:%.4f" % (1 / 63)


def test_compare_ab_lists_the_photons_whose_records_differ(tmp_path):
    """tests/compare_ab.py of the reference, as analysis.compare_ab and scripts/compare_ab.py"""
    import os, subprocess, sys
    orc = Oracle()
    w = workloads.scintillator_tank(num_photon=2000, photons_per_genstep=100)
    a = orc.simulate(w["geom"], w["gensteps"])
    rec = a["record"]
    b = np.zeros_like(rec)
    b[:, :-1] = rec[:, 1:]                          # a B side without the generation point, like the U4Recorder arrays
    b[[14, 22, 81], 0, 0, 0] += 1e-3                # three photons moved by a micron at their first step
    assert A.compare_ab(rec, b) == [14, 22, 81]
    assert A.compare_ab(rec, rec, shifted=False) == []
    da = A.save_event(str(tmp_path / "a"), 0, dict(record=rec, seq=a["seq"]))
    db = A.save_event(str(tmp_path / "b"), 0, dict(record=b, seq=a["seq"]))
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "scripts", "compare_ab.py"), da, db], capture_output=True, text=True)
    assert r.returncode == 0 and "[14, 22, 81]" in r.stdout and "chi2/ndf 0.00" in r.stdout, r.stdout + r.stderr
