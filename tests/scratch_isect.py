import sys, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import eic_opticks_b200 as ph
from eic_opticks_b200 import foundry as F, tables as T
from _ref import RefGPU

def single(shape, frame=None):
    bt = T.BoundaryTable(); bt.add_material(T.Material("A")); b = bt.boundary("A","","","A")
    fd = F.Foundry(); fd.begin_solid("r0"); fd.add_prim(shape, b, frame); fd.end_solid()
    bnd, opt = bt.arrays()
    return dict(foundry=fd.arrays(), bnd=bnd, optical=opt, icdf=None)

shapes = {
 "sphere": F.sphere(100), "sphere_tr": F.sphere(100).placed(F.translate(10,20,30)),
 "zsphere": F.zsphere(100,-50,70), "box3": F.box3(100,150,200), "box3_rot": F.box3(100,150,200).placed(F.rotate_z(30)@F.translate(5,6,7)),
 "cylinder": F.cylinder(80,-100,100), "cone": F.cone(100,-80,30,80), "hyperboloid": F.hyperboloid(50,80,-100,100),
 "diff": F.difference(F.box3(200,200,200), F.sphere(120)), "inter": F.intersection(F.sphere(130), F.box3(200,200,200)),
 "union": F.union(F.cylinder(60,-100,0), F.cone(60,0,10,120)),
}
ref = RefGPU("debugtag")
rng = np.random.default_rng(7)
n = 200000
for name, shp in shapes.items():
    g = single(shp)
    o = rng.uniform(-250,250,(n,3)).astype(np.float32)
    tgt = rng.uniform(-90,90,(n,3)).astype(np.float32)
    d = tgt - o; d /= np.linalg.norm(d,axis=1)[:,None]; d = d.astype(np.float32)
    # half of the rays start inside-ish
    o[::2] = rng.uniform(-60,60,(n//2+n%2,3)).astype(np.float32)[:len(o[::2])]
    r = ref.intersect(g, o, d, tmin=0.05)
    sim = ph.Simulator.Create(g["foundry"], g["bnd"], g["optical"])
    out = {}
    for accel in (1,0):
        p = sim.intersect(o, d, tmin=0.05, accel=accel)
        pu, ru = p.view(np.uint32), r.view(np.uint32)
        hit_r = ru[:,1,3] != 0xffffffff; hit_p = pu[:,1,3] != 0xffffffff
        same_hit = (hit_r == hit_p).mean()
        both = hit_r & hit_p
        t_same = (pu[both,0,3] == ru[both,0,3]).mean()
        n_same = (pu[both,0,:3] == ru[both,0,:3]).all(axis=1).mean()
        lp_same = (pu[both,1,:2] == ru[both,1,:2]).all(axis=1).mean()
        dt = np.abs(p[both,0,3]-r[both,0,3]).max()
        out[accel]=(same_hit,t_same,n_same,lp_same,dt)
        print(f"{name:12s} accel={accel} hitfrac={hit_r.mean():.3f} same_hit={same_hit:.6f} t_bits={t_same:.6f} n_bits={n_same:.6f} lpos_bits={lp_same:.6f} max|dt|={dt:.3g}")
        if accel==1 and t_same<1:
            bad = np.where(both)[0][pu[both,0,3] != ru[both,0,3]][:3]
            for b in bad: print("    o",o[b],"d",d[b],"t",p[b,0,3],r[b,0,3], "n", p[b,0,:3], r[b,0,:3])
    sim.close()
