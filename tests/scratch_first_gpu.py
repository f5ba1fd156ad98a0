"""first GPU bring-up: product vs reference-header harness on every hand-built geometry"""
import sys, time
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import eic_opticks_b200 as ph
from eic_opticks_b200 import gensteps as G
from _ref import RefGPU

def seqhis_str(s):
    names = ["??","CK","SI","TO","AB","RE","SC","SD","SA","DR","SR","BR","BT","NA","EC","EX","MI"]
    out=[]
    for w in range(2):
        v=int(s[w,0])
        for k in range(16):
            nib=(v>>(4*k))&0xf
            if nib==0: return " ".join(out)
            out.append(names[nib])
    return " ".join(out)

def compare(tag, geom, gs, ip=None, n_show=5, **kw):
    ref = RefGPU("debugtag")
    t0=time.time(); r = ref.simulate(geom, gs, ip, **kw); t1=time.time()
    sim = ph.Simulator.Create(geom["foundry"], geom["bnd"], geom["optical"], geom.get("icdf"), event_mode=ph.MODE_DEBUGHEAVY, max_record=32,
                              max_bounce=kw.get("max_bounce",31))
    for accel in (ph.ACCEL_BRUTE, ph.ACCEL_BVH):
        sim.set_config(accel=accel)
        t2=time.time(); hits = sim.simulate_np(gs, kw.get("event_id",0), ip, kw.get("photon_offset",0)); t3=time.time()
        p = sim.get_array("photon"); s = sim.get_array("seq"); rec = sim.get_array("record")
        same_seq = (s == r["seq"]).all(axis=(1,2))
        pu, ru = p.view(np.uint32), r["photon"].view(np.uint32)
        same_flags = (pu[:,3,:] == ru[:,3,:]).all(axis=1)
        good = same_seq
        dpos = np.abs(p[good,0,:] - r["photon"][good,0,:]).max() if good.any() else -1
        bit = (pu == ru).all(axis=(1,2))
        st = sim.stats()
        print(f"{tag:28s} accel={accel} N={len(p)} hits={len(hits)} seq_same={same_seq.mean():.6f} flags_same={same_flags.mean():.6f} bit_identical={bit.mean():.6f} max|dpos,t|={dpos:.3g} rays={st['num_ray']} ref_rays={r['nray']} t_ref={t1-t0:.2f}s t_phox={t3-t2:.3f}s", flush=True)
        bad = np.where(~same_seq)[0][:n_show]
        for b in bad:
            print("   idx", b, "phox:", seqhis_str(s[b]), "| ref:", seqhis_str(r["seq"][b]))
    u,c = np.unique(r["seq"][:,0,0], return_counts=True)
    order = np.argsort(-c)[:6]
    for o in order: print("      ", c[o], seqhis_str(np.array([[u[o],0],[0,0]],dtype=np.uint64)))
    sim.close()

N = int(sys.argv[1]) if len(sys.argv)>1 else 20000
g = ph.geometries.raindrop()
t,_ = G.torch_config("tests/golden/config_dev.json") if False else (dict(pos=[-10,-30,-90],time=0,mom=G._normalize_f32([0,0.3,1.0]),pol=[1,0,0],wavelength=420.0,radius=15.0,numphoton=N,type="disc"),{})
compare("raindrop torch genstep", g, G.torch_genstep(t, N))
ipn = G.torch_photons(t, N, seed=0)
compare("raindrop input photons", g, G.input_photon_genstep(N), ipn)
# cerenkov in water
rng = np.random.default_rng(1)
ngs=200
pos = np.stack([np.zeros(ngs), 0.2*np.linspace(-40,40,ngs), 0.8*np.linspace(-40,40,ngs)],axis=1)
d = np.array([0,0.2,0.8]); d/=np.linalg.norm(d)
gs = G.cerenkov_gensteps(pos, d, 1.0, rng.poisson(N/ngs, ngs), g["water_line"], 1.0, 80.0, 800.0, 1.333, pre_velocity=299.0, post_velocity=298.0, mean_photons=(2.0,1.5))
compare("raindrop cerenkov", g, gs)
g = ph.geometries.sphere_leak()
t2 = dict(pos=[0,0,0],time=0,mom=G._normalize_f32([0,0.3,1.0]),pol=[1,0,0],wavelength=420.0,radius=0.1,numphoton=N,type="disc")
compare("sphere_leak torch", g, G.torch_genstep(t2, N))
t3 = dict(t2); t3["pos"]=[0,0,20.0]; t3["radius"]=3.0
compare("sphere_leak torch outside glass", g, G.torch_genstep(t3, N))
g = ph.geometries.sipm8x8()
t4 = dict(pos=[-8.7,-8.7,4.0],time=0,mom=[0,0,1.0],pol=[1,0,0],wavelength=420.0,radius=0.5,numphoton=N,type="disc")
compare("sipm8x8 torch", g, G.torch_genstep(t4, N), max_bounce=32)
cc = g["crystal_centers"]; ngs=256
pos = cc[rng.integers(0,64,ngs)] + rng.uniform(-0.9,0.9,(ngs,3))*np.array([1,1,3.9])
dirs = rng.normal(size=(ngs,3)); dirs/=np.linalg.norm(dirs,axis=1)[:,None]
gs = G.scint_gensteps(pos, dirs, 0.05, rng.poisson(N/ngs, ngs), g["crystal_line"], g["scintillation_time"])
compare("sipm8x8 scint", g, gs, max_bounce=32)
g = ph.geometries.pmt_wall(10,10)
t5 = dict(pos=[0,0,800.0],time=0,mom=[0,0,-1.0],pol=[1,0,0],wavelength=420.0,radius=1200.0,numphoton=N,type="disc")
compare("pmt_wall torch", g, G.torch_genstep(t5, N))
g = ph.geometries.boolean_zoo()
t6 = dict(pos=[0,0,0.0],time=0,mom=[0,0,1.0],pol=[1,0,0],wavelength=420.0,radius=100.0,numphoton=N,type="sphere", zenith=[0,1],azimuth=[0,1])
t6["radius"]=950.0*-1  # inward from a sphere
compare("boolean_zoo inward sphere", g, G.torch_genstep(t6, N))
t7 = dict(t6); t7["radius"]=30.0; t7["pos"]=[0,0,0]
compare("boolean_zoo outward sphere", g, G.torch_genstep(t7, N))
