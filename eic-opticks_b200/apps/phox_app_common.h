// phox_app_common.h : what the standalone drivers share - load a persisted CSGFoundry directory
// (CSG/CSGFoundry.cc:2768-2802 layout, written by eic_opticks_b200.foundry.save_geometry or a reference install) into a
// PhoxSimulator, and write hits in the text format of the reference apps (src/GPUPhotonSourceMinimal.h:124-144).
#pragma once
#include <fstream>
#include <iostream>
#include <stdexcept>
#include <string>

#include "../../include/PhoxSimulator.h"
#include "../../include/phox_npy.h"

namespace phoxapp {

inline PhoxSimulator* create_from_geometry_dir(const std::string& geom, int device) {
    using phoxnpy::load;
    std::string fd = geom + "/CSGFoundry/", ss = fd + "SSim/stree/standard/";
    auto solid = load(fd + "solid.npy"), prim = load(fd + "prim.npy"), node = load(fd + "node.npy"), itra = load(fd + "itra.npy");
    auto inst = load(fd + "inst.npy"), plan = load(fd + "plan.npy", false);
    auto bnd = load(ss + "bnd.npy"), optical = load(ss + "optical.npy"), icdf = load(ss + "icdf.npy", false);
    if (bnd.dtype != "<f4" || bnd.shape.size() != 5) throw std::runtime_error("bnd.npy must be float32 (nbnd,4,2,nwl,4)");
    return PhoxSimulator::Create(solid.data.data(), solid.shape[0], prim.data.data(), prim.shape[0], node.data.data(), node.shape[0],
                                 plan.empty() ? nullptr : plan.data.data(), plan.empty() ? 0 : plan.shape[0], itra.data.data(), itra.shape[0],
                                 inst.data.data(), inst.shape[0], bnd.as<float>(), bnd.shape[0], bnd.shape[3], 60.f, 1.f, optical.as<int32_t>(),
                                 icdf.empty() ? nullptr : icdf.as<float>(), icdf.empty() ? 0 : 3, icdf.empty() ? 0 : icdf.count() / 3, 20, device);
}

// "time wavelength  (x, y, z)  (mx, my, mz)  (px, py, pz)" per hit
inline void write_hits_text(PhoxSimulator* cx, const std::string& path) {
    std::ofstream of(path);
    if (!of.is_open()) { std::cerr << "Error opening output file!" << std::endl; return; }
    unsigned nhit = cx->getNumHit();
    for (unsigned i = 0; i < nhit; i++) {
        PhoxPhoton h;
        cx->getHit(h, i);
        of << h.q[3] << " " << h.q[11] << "  (" << h.q[0] << ", " << h.q[1] << ", " << h.q[2] << ")  (" << h.q[4] << ", " << h.q[5] << ", " << h.q[6]
           << ")  (" << h.q[8] << ", " << h.q[9] << ", " << h.q[10] << ")" << std::endl;
    }
}

}  // namespace phoxapp
