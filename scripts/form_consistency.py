"""Do the kernels agree with each other?  Photon arrays of the same event from the persistent and the wavefront form (production
kernels: physics compiled in place) and from the debug kernels (physics out of line): integer data and float bits.  nvcc fuses a*b+c
per compilation context, so identity between kernels is a property of the build that has to be looked at, not assumed."""
import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np
import eic_opticks_b200 as ph
from eic_opticks_b200 import workloads
from _parity import ulp_distance
for name,kw in (("box_maze_photons",dict(num_photon=400000)),("sipm8x8_scint",dict(num_photon=1000000, photons_per_genstep=100)),("boolean_zoo_torch",dict(num_photon=1000000)),("scintillator_tank",dict(num_photon=1000000,photons_per_genstep=100)),
                ("pmt_wall_torch",dict(num_photon=1000000,nx=20,ny=20)),("sphere_leak_torch",dict(num_photon=200000)),("raindrop_cerenkov",dict(num_photon=1000000)),("halfspace_zoo_torch",dict(num_photon=300000)),("pmt_wall_sensor_a",dict(num_photon=300000))):
    w=workloads.WORKLOADS[name](**kw); g=w["geom"]
    out={}
    for mode in (ph.KERNEL_PERSISTENT, ph.KERNEL_WAVEFRONT):
        for em in (ph.MODE_HITPHOTON, ph.MODE_DEBUGLITE):
            kwc=dict(w["config"]); 
            sim=ph.Simulator.Create(g["foundry"], g["bnd"], g["optical"], g["icdf"], event_mode=em, kernel_mode=mode, **kwc)
            sim.simulate_np(w["gensteps"], 2, w["input_photons"])
            out[(mode,em)]=sim.get_array("photon").copy(); sim.close()
    a=out[(ph.KERNEL_PERSISTENT,ph.MODE_HITPHOTON)]; b=out[(ph.KERNEL_WAVEFRONT,ph.MODE_HITPHOTON)]; c=out[(ph.KERNEL_WAVEFRONT,ph.MODE_DEBUGLITE)]
    for tag,x,y in (("persistent vs wavefront (production kernels)",a,b),("production vs debug kernels (wavefront)",b,c)):
        xi,yi=x.view(np.uint32),y.view(np.uint32)
        same_int=(xi[:,3,:]==yi[:,3,:]).all(axis=1)
        u=ulp_distance(x[:,:3,:],y[:,:3,:]).reshape(len(x),-1).max(axis=1)
        print(name,tag,"photons",len(x),"int-differ",int((~same_int).sum()),"float-differ",int((u>0).sum()),"max ulp among same-int",int(u[same_int].max()) if same_int.any() else -1)
