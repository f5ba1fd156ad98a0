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

#include "phox_app_common.h"

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
        std::vector<PhoxPhoton> ph = load_photons_txt(photons);
        if (ph.empty()) { std::cerr << "ERROR: no photons loaded from " << photons << std::endl; return 1; }
        std::cout << "Loaded " << ph.size() << " photons from " << photons << std::endl;

        PhoxSimulator* cx = phoxapp::create_from_geometry_dir(geom, device);
        std::cout << cx->desc() << std::endl;
        cx->setInputPhoton(ph.data(), (int64_t)ph.size());
        double dt = cx->simulate(0, false);
        unsigned nhit = cx->getNumHit();
        std::cout << "Simulation time: " << dt << " seconds" << std::endl;
        std::cout << "Opticks: NumHits:  " << nhit << std::endl;
        phoxapp::write_hits_text(cx, out);
        cx->reset(0);
        delete cx;
    } catch (const std::exception& e) {
        std::cerr << "ERROR: " << e.what() << std::endl;
        return 2;
    }
    return 0;
}
