"""CPU tests of the host side: C-ABI exports, array builders, genstep producers, sharding logic."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import eic_opticks_b200 as ph
from eic_opticks_b200 import foundry as F, gensteps as G, tables as T, workloads

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "phox.h")).read()
    declared = set(re.findall(r"\b(phox_[a-z_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    lib = C.CDLL(ph.lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), "libphox.so does not export %s" % name
    assert declared == set(ph.lib.SYMBOLS), (declared ^ set(ph.lib.SYMBOLS))


def test_config_struct_matches_header_defaults():
    cfg = ph.lib.default_config()
    assert C.sizeof(ph.lib.Config) == 96
    assert (cfg.max_bounce, cfg.event_mode, cfg.rng_mode, cfg.accel) == (31, 0, 1, 0)
    assert cfg.hit_mask == 0x40 and cfg.epsilon0_mask == 0x37            # SD ; TO|CK|SI|SC|RE
    assert abs(cfg.propagate_epsilon - 0.05) < 1e-9 and cfg.tmax == 1e6 and cfg.skipahead_event_offset == 100000


def test_no_gpu_means_loud_failure_not_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(ph.PhoxError) as e:
        ph.Simulator()
    assert "no CUDA device" in str(e.value) or "CUDA" in str(e.value)


def test_philox_host_matches_curand_golden():
    for line in open(os.path.join(GOLD, "curand_philox.txt")):
        f = line.split()
        seed, sub, off, skip, n = (int(x) for x in f[:5])
        want = np.array([int(x) for x in f[5:5 + n]], dtype=np.uint32)
        got = G.curand_uniform_stream(seed, sub, off + skip, n).view(np.uint32)
        assert (got == want).all(), (seed, sub, off, skip)


def test_torch_config_and_host_photons():
    t, ev = G.torch_config(os.path.join(GOLD, "config_dev.json"))
    assert ev["mode"] == "DebugLite" and t["numphoton"] == 100
    assert abs(np.linalg.norm(t["mom"]) - 1) < 1e-6
    p = G.torch_photons(t, 0, seed=0)
    assert p.shape == (100, 4, 4)
    # photons start on a disc of radius 15 about pos, perpendicular to mom, pol perpendicular to mom
    d = p[:, 0, :3] - np.array(t["pos"], dtype=np.float32)
    assert np.abs(d @ t["mom"]).max() < 1e-4 and np.linalg.norm(d, axis=1).max() <= 15.0 + 1e-4
    assert np.abs((p[:, 2, :3] * p[:, 1, :3]).sum(1)).max() < 1e-5
    assert (p.view(np.uint32)[:, 3, 3] == 4).all() and (p[:, 2, 3] == 420.0).all()
    gs = G.torch_genstep(t)
    assert gs.view(np.uint32)[0, 0, 0] == 6 and gs.view(np.uint32)[0, 0, 3] == 100 and gs.view(np.uint32)[0, 5, 3] == 1


def test_photons_from_text_skips_comments_blank_and_malformed(tmp_path):
    ph_ = G.photons_from_text(os.path.join(GOLD, "photons_file_source.txt"))
    assert ph_.shape == (10, 4, 4)
    assert sorted(ph_[:, 2, 3].tolist()) == [420.0] * 3 + [450.0] * 2 + [500.0] * 5
    assert (ph_.view(np.uint32)[:, 3, :] == 0).all()
    out = tmp_path / "hits.txt"
    G.write_hits_text(ph_, str(out))
    assert len(open(out).read().strip().split("\n")) == 10


def test_foundry_layouts_and_tree_conventions():
    g = ph.geometries.boolean_zoo()
    fd = g["foundry"]
    assert fd["prim"].dtype == np.float32 and fd["prim"].shape[1:] == (4, 4) and fd["solid"].shape[1:] == (3, 4)
    node_u = fd["node"].view(np.uint32).reshape(-1, 16)
    prim_i = fd["prim"].view(np.int32).reshape(-1, 16)
    for p in range(len(prim_i)):
        nn, no = prim_i[p, 0], prim_i[p, 1]
        root_tc = node_u[no, 14]
        if root_tc < 11:                                   # boolean tree: complete binary, subNum on the root
            assert node_u[no, 0] == ((1 << int(np.log2(node_u[no, 0] + 1))) - 1)
            assert node_u[no, 0] <= nn
            tcs = node_u[no:no + node_u[no, 0], 14]
            assert not (tcs == F.CSG_DIFFERENCE).any(), "trees must be positivised"
        for k in range(nn):                                # every leaf carries a valid 1-based transform
            tc = node_u[no + k, 14]
            if tc >= 101:
                ti = node_u[no + k, 15] & 0x7fffffff
                assert 1 <= ti <= len(fd["itra"])
    # tran * itra = identity
    prod = np.einsum("nij,njk->nik", fd["tran"].astype(np.float64), fd["itra"].astype(np.float64))
    assert np.abs(prod - np.eye(4)).max() < 1e-4          # float32 storage, translations ~ 1e3


def test_foundry_save_load_roundtrip(tmp_path):
    fd = ph.geometries.pmt_wall(3, 3)["foundry"]
    F.save_foundry(fd, str(tmp_path / "CSGFoundry"))
    back = F.load_foundry(str(tmp_path / "CSGFoundry"))
    for k in ("solid", "prim", "node", "tran", "itra", "inst"):
        assert (back[k].view(np.uint32) == fd[k].view(np.uint32)).all()
    inst_i = fd["inst"].view(np.int32)
    assert (inst_i[1:, 1, 3] == 1).all() and inst_i[0, 1, 3] == 0       # gas_idx
    assert (inst_i[1:, 2, 3] == np.arange(1, 10)).all()                # sensor_identifier + 1


def test_boundary_table_conventions():
    g = ph.geometries.sipm8x8()
    bnd, opt, names = g["bnd"], g["optical"], g["bnd_names"]
    assert bnd.shape == (len(names), 4, 2, 761, 4) and opt.shape == (4 * len(names), 4)
    i = names.index("Air/CrystalSkin/CrystalSkin/Crystal")
    assert np.allclose(bnd[i, 3, 0, :, 0], 1.82) and np.allclose(bnd[i, 3, 0, :, 1], 400.0)         # imat RINDEX, ABSLENGTH
    assert np.allclose(bnd[i, 0, 0, :, 0], 1.0) and np.allclose(bnd[i, 0, 1, :, 0], 299.792458)     # omat Air, GROUPVEL c/n
    assert np.allclose(bnd[i, 1, 0, 0], [0.0, 0.02, 0.98, 0.0], atol=1e-6)                         # polished non-sensor
    assert tuple(opt[4 * i + 1][:2]) == (opt[4 * i + 1][0], T.EMS_SURFACE) and opt[4 * i + 1][0] > 0
    j = names.index("Air///OpticalGrease")
    assert (bnd[j, 1] == -1).all() and tuple(opt[4 * j + 1]) == (0, T.EMS_NOSURFACE, 0, 0)
    k = names.index("Air/SiPMActiveSkin/SiPMActiveSkin/EntranceWindow")
    assert np.allclose(bnd[k, 2, 0, 0], [1.0, 0.0, 0.0, 0.0])                                      # sensor: detect = efficiency
    assert g["crystal_line"] == 4 * i + 3
    assert g["icdf"].shape == (3, 4096) and (np.diff(g["icdf"][0]) <= 1e-3).all()      # energy ascends with u, wavelength descends
    assert abs(g["icdf"][0, 0] - 1239.84198 / 1.5) < 0.01 and g["icdf"][2, -1] < 311.0


def test_partition_gensteps_and_shard_event():
    from eic_opticks_b200 import parallel
    w = workloads.sipm8x8_scint(num_photon=100000, photons_per_genstep=100)
    gs = w["gensteps"]
    num = gs.view(np.uint32)[:, 0, 3].astype(np.int64)
    assert num.sum() == 100000 == w["num_photon"]
    for world in (1, 2, 3, 8):
        parts = G.partition_gensteps(gs, world)
        assert parts[0][0] == 0 and parts[-1][1] == len(gs)
        off = 0
        for r, (s0, s1, o, c) in enumerate(parts):
            assert o == off and c == num[s0:s1].sum()
            if r:
                assert s0 == parts[r - 1][1]
            off += c
        assert off == 100000
        assert max(p[3] for p in parts) - min(p[3] for p in parts) <= 2 * num.max()
    ip = np.zeros((10, 4, 4), dtype=np.float32)
    got = [parallel.shard_event(G.input_photon_genstep(10), r, 3, ip) for r in range(3)]
    assert [g_[2] for g_ in got] == [0, 3, 6] and sum(g_[3] for g_ in got) == 10
    assert all(g_[0].view(np.uint32)[0, 0, 3] == g_[3] for g_ in got)


def test_cxx_adaptor_compiles_standalone_and_geometry_dir_roundtrip(tmp_path):
    import subprocess
    src = tmp_path / "t.cpp"
    src.write_text('#include "PhoxSimulator.h"\n#include "phox_npy.h"\nint main(){ SSimulator* s = nullptr; (void)s; return 0; }\n')
    r = subprocess.run(["/usr/bin/g++", "-std=c++17", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), str(src)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    g = ph.geometries.sipm8x8()
    F.save_geometry(g, str(tmp_path / "geom"))
    back = F.load_geometry(str(tmp_path / "geom"))
    assert (back["bnd"] == g["bnd"]).all() and (back["optical"] == g["optical"]).all() and (back["icdf"] == g["icdf"]).all()
    assert back["bnd_names"] == g["bnd_names"]
    assert os.path.exists(os.path.join(ROOT, "eic-opticks_b200", "apps", "PhoxPhotonFileSource"))


def test_cxx_torch_driver_generates_the_reference_photons(tmp_path):
    """apps/PhoxPhotonSourceMinimal.cpp: JSON config parsing (src/config.cpp:107-145) and host torch generation
    (src/torch.cpp:8-30) against the python restatement, without a GPU (--dump-photons)"""
    import subprocess
    from eic_opticks_b200 import gensteps as G
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "eic-opticks_b200", "apps", "PhoxPhotonSourceMinimal")
    assert os.path.exists(exe), "run __graft_entry__.build()"
    cfg = os.path.join(root, "tests", "golden", "config_dev.json")
    out = tmp_path / "photons.txt"
    r = subprocess.run([exe, "-c", cfg, "--dump-photons", str(out)], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stderr
    assert "Dumped 100 photons, event mode DebugLite maxslot 1000000" in r.stdout
    a = np.loadtxt(out)
    t, _ = G.torch_config(cfg)
    ref = G.torch_photons(t, seed=0)
    assert a.shape == (100, 16)
    assert np.abs(a[:, :12] - ref.reshape(100, 16)[:, :12]).max() < 2e-5          # sinf/cosf of glibc vs numpy differ by ulps of a 15 mm radius
    assert (a[:, 12] == 4).all() and (a[:, 15] == 4).all()
    bad = tmp_path / "bad.json"; bad.write_text('{"torch": {"gentype": "TORCH"}}')
    r2 = subprocess.run([exe, "-c", str(bad), "--dump-photons", str(out)], capture_output=True, text=True, timeout=60)
    assert r2.returncode != 0 and "missing key" in r2.stderr


def test_cxx_genstep_slices_equal_the_python_partition(tmp_path):
    """include/PhoxMultiGPU.h PhoxMultiGPU::slices (the C++ multi-GPU host's cut of an event into per-device genstep ranges, the
    concurrent form of SGenstep::GetGenstepSlices, sysrap/SGenstep.h:249-323) against gensteps.partition_gensteps: same ranges,
    absolute photon offsets and counts, for ragged gensteps, more ranks than gensteps, empty gensteps and one huge genstep."""
    import subprocess
    src = tmp_path / "s.cpp"
    src.write_text(r'''
#include "PhoxMultiGPU.h"
#include "phox_npy.h"
#include <cstdio>
int main(int argc, char** argv) {
    phoxnpy::Array gs = phoxnpy::load(argv[1]);
    const int64_t n = gs.count() / 24;
    for (int nr = 1; nr <= 9; nr++) {
        auto sl = PhoxMultiGPU::slices(gs.data.data(), n, nr);
        for (auto& s : sl) std::printf("%d %lld %lld %llu %lld\n", nr, (long long)s.gs_start, (long long)s.gs_stop, (unsigned long long)s.ph_offset, (long long)s.ph_count);
    }
    return 0;
}
''')
    exe = tmp_path / "s"
    csrc = os.path.join(ROOT, "eic-opticks_b200", "csrc")
    r = subprocess.run(["/usr/bin/g++", "-std=c++17", "-O1", "-pthread", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src), "-L", csrc, "-lphox",
                        "-Wl,-rpath," + csrc], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    rng = np.random.default_rng(5)
    cases = {"ragged": rng.integers(0, 3000, size=57), "few": np.array([10, 0, 7]), "one_huge": np.array([5, 1000000, 3]), "empty_event": np.array([0, 0])}
    for name, counts in cases.items():
        gs = G.empty_gensteps(len(counts))
        gs.view(np.uint32)[:, 0, 3] = counts
        f = tmp_path / (name + ".npy")
        np.save(f, gs)
        out = subprocess.run([str(exe), str(f)], capture_output=True, text=True)
        assert out.returncode == 0, out.stderr
        got = [tuple(int(x) for x in line.split()) for line in out.stdout.strip().splitlines()]
        want = [(nr,) + tuple(p) for nr in range(1, 10) for p in G.partition_gensteps(gs, nr)]
        assert got == want, name


def test_bvh_depth_guard_function(tmp_path):
    """phox_bvh.cuh bvh_tree_depth (host): what phox_set_geometry uses to refuse trees deeper than the traversal stack"""
    import subprocess
    src = tmp_path / "d.cpp"
    src.write_text(r'''
#include "phox_bvh.cuh"
#include <cstdio>
using namespace phox;
static BvhNode mk(int c0, int c1) { BvhNode n{}; n.d.x = c0; n.d.y = c1; return n; }
int main() {
    // one item
    BvhNode one[1] = {mk(~5, kBvhNoChild)};
    // chain: node i -> (leaf, node i+1), 6 internal nodes ; leaves carry the flag bits the engine sets (still negative)
    BvhNode chain[6];
    for (int i = 0; i < 6; i++) chain[i] = mk(~(i | 0x40000000), i < 5 ? i + 1 : ~(99 | 0x20000000));
    // balanced over 4 items: root(1,2), 1(l,l), 2(l,l)
    BvhNode bal[3] = {mk(1, 2), mk(~0, ~1), mk(~2, ~3)};
    // malformed: cycle, and child out of range
    BvhNode cyc[2] = {mk(1, ~0), mk(0, ~1)};
    BvhNode oor[1] = {mk(7, ~0)};
    std::printf("%d %d %d %d %d %d\n", bvh_tree_depth(one, 1), bvh_tree_depth(chain, 6), bvh_tree_depth(bal, 3), bvh_tree_depth(cyc, 2),
                bvh_tree_depth(oor, 1), bvh_tree_depth(nullptr, 0));
    return 0;
}
''')
    exe = tmp_path / "d"
    r = subprocess.run(["/usr/bin/g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "eic-opticks_b200", "csrc"), "-I", "/usr/local/cuda/include",
                        "-o", str(exe), str(src)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = subprocess.run([str(exe)], capture_output=True, text=True).stdout.split()
    assert out == ["1", "6", "2", "-1", "-1", "-1"], out
