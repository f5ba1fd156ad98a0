// PhoxPhotonFileSource : standalone C++ driver with the command-line contract of the reference's
// GPUPhotonFileSource (src/GPUPhotonFileSource.cpp, .h:51-87, 124-180): photons from a text file in,
// "Opticks: NumHits:  N" on stdout and opticks_hits_output.txt out.  Geant4 is not involved: the
// geometry comes from a persisted CSGFoundry directory (CSG/CSGFoundry.cc:2768-2802, written here by
// eic_opticks_b200.foundry.save_geometry or by a reference install) instead of a GDML file.
//
//   PhoxPhotonFileSource -g <geometry dir> -p <photons.txt> [-o opticks_hits_output.txt] [-d device]
//
// Host code is C++ on the C ABI (include/phox.h) through the SSimulator adaptor (include/PhoxSimulator.h).
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/PhoxSimulator.h"
#include "../../include/phox_npy.h"

static std::vector<PhoxPhoton> load_photons_txt(const std::string& path) {
    std::vector<PhoxPhoton> out;
    std::ifstream ifs(path);
    if (!ifs.is_open()) { std::cerr << "ERROR: cannot open photon file: " << path << std::endl; return out; }
    std::string line;
    int lineno = 0;
    while (std::getline(ifs, line)) {
        lineno++;
        if (line.empty() || line[0] == '#') continue;
        std::istringstream ss(line);
        float v[11];
        bool ok = true;
        for (int k = 0; k < 11; k++) if (!(ss >> v[k])) { ok = false; break; }
        if (!ok) { std::cerr << "WARNING: skipping malformed line " << lineno << ": " << line << std::endl; continue; }
        PhoxPhoton p = {};
        p.q[0] = v[0]; p.q[1] = v[1]; p.q[2] = v[2]; p.q[3] = v[3];
        p.q[4] = v[4]; p.q[5] = v[5]; p.q[6] = v[6];
        p.q[8] = v[7]; p.q[9] = v[8]; p.q[10] = v[9]; p.q[11] = v[10];
        out.push_back(p);
    }
    return out;
}

int main(int argc, char** argv) {
    std::string geom, photons, out = "opticks_hits_output.txt";
    int device = 0;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        if ((a == "-g" || a == "--geometry") && i + 1 < argc) geom = argv[++i];
        else if ((a == "-p" || a == "--photons") && i + 1 < argc) photons = argv[++i];
        else if ((a == "-o" || a == "--output") && i + 1 < argc) out = argv[++i];
        else if ((a == "-d" || a == "--device") && i + 1 < argc) device = std::atoi(argv[++i]);
    }
    if (geom.empty() || photons.empty()) {
        std::cerr << "usage: PhoxPhotonFileSource -g <geometry dir> -p <photons.txt> [-o hits.txt] [-d device]" << std::endl;
        return 1;                                  // the reference fails without -p (tests/test_GPUPhotonFileSource.sh:108-118)
    }
    try {
        using phoxnpy::load;
        std::string fd = geom + "/CSGFoundry/", ss = fd + "SSim/stree/standard/";
        auto solid = load(fd + "solid.npy"), prim = load(fd + "prim.npy"), node = load(fd + "node.npy"), itra = load(fd + "itra.npy");
        auto inst = load(fd + "inst.npy"), plan = load(fd + "plan.npy", false);
        auto bnd = load(ss + "bnd.npy"), optical = load(ss + "optical.npy"), icdf = load(ss + "icdf.npy", false);
        if (bnd.dtype != "<f4" || bnd.shape.size() != 5) throw std::runtime_error("bnd.npy must be float32 (nbnd,4,2,nwl,4)");
        std::vector<PhoxPhoton> ph = load_photons_txt(photons);
        if (ph.empty()) { std::cerr << "ERROR: no photons loaded from " << photons << std::endl; return 1; }
        std::cout << "Loaded " << ph.size() << " photons from " << photons << std::endl;

        PhoxSimulator* cx = PhoxSimulator::Create(solid.data.data(), solid.shape[0], prim.data.data(), prim.shape[0], node.data.data(), node.shape[0],
                                                  plan.empty() ? nullptr : plan.data.data(), plan.empty() ? 0 : plan.shape[0], itra.data.data(),
                                                  itra.shape[0], inst.data.data(), inst.shape[0], bnd.as<float>(), bnd.shape[0], bnd.shape[3], 60.f, 1.f,
                                                  optical.as<int32_t>(), icdf.empty() ? nullptr : icdf.as<float>(), icdf.empty() ? 0 : 3,
                                                  icdf.empty() ? 0 : icdf.count() / 3, 20, device);
        std::cout << cx->desc() << std::endl;
        cx->setInputPhoton(ph.data(), (int64_t)ph.size());
        double dt = cx->simulate(0, false);
        unsigned nhit = cx->getNumHit();
        std::cout << "Simulation time: " << dt << " seconds" << std::endl;
        std::cout << "Opticks: NumHits:  " << nhit << std::endl;
        std::ofstream of(out);
        for (unsigned i = 0; i < nhit; i++) {
            PhoxPhoton h;
            cx->getHit(h, i);
            of << h.q[3] << " " << h.q[11] << "  (" << h.q[0] << ", " << h.q[1] << ", " << h.q[2] << ")  (" << h.q[4] << ", " << h.q[5] << ", " << h.q[6]
               << ")  (" << h.q[8] << ", " << h.q[9] << ", " << h.q[10] << ")" << std::endl;
        }
        cx->reset(0);
        delete cx;
    } catch (const std::exception& e) {
        std::cerr << "ERROR: " << e.what() << std::endl;
        return 2;
    }
    return 0;
}
