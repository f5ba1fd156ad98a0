// PhoxMultiGPU : C++ host for one box of GPUs - geometry directory + genstep file in, hits out, photons/s reported.
//
//   PhoxMultiGPU -g <geometry dir> -G <gensteps.npy> [-I <input_photons.npy>] [--gpus N | --devices 0,0,1] [--events E]
//                [-o hits.npy] [--max-bounce B] [--seed S]
//
// Every event is cut into one contiguous genstep range per GPU (include/PhoxMultiGPU.h, the concurrent form of
// SGenstep::GetGenstepSlices, sysrap/SGenstep.h:249-323), simulated by one thread + phox_context per device, and its
// hits gathered into one page-locked buffer in ascending photon index while the next event runs.  With --events E the
// same gensteps are run E times with event ids 0..E-1 (the reference's event loop, G4CXOpticks::simulate per event) and
// the LAST event's hits are written.  A JSON line with the timing goes to stdout.
#include <chrono>
#include <cstdlib>
#include <iostream>
#include <sstream>

#include "../../include/PhoxMultiGPU.h"
#include "../../include/phox_npy.h"

struct Geometry {
    phoxnpy::Array solid, prim, node, itra, inst, plan, bnd, optical, icdf;
    explicit Geometry(const std::string& geom) {
        using phoxnpy::load;
        const std::string fd = geom + "/CSGFoundry/", ss = fd + "SSim/stree/standard/";
        solid = load(fd + "solid.npy"); prim = load(fd + "prim.npy"); node = load(fd + "node.npy"); itra = load(fd + "itra.npy");
        inst = load(fd + "inst.npy"); plan = load(fd + "plan.npy", false);
        bnd = load(ss + "bnd.npy"); optical = load(ss + "optical.npy"); icdf = load(ss + "icdf.npy", false);
        if (bnd.dtype != "<f4" || bnd.shape.size() != 5) throw std::runtime_error("bnd.npy must be float32 (nbnd,4,2,nwl,4)");
    }
    int upload(phox_context* ctx) const {
        int rc = phox_set_geometry(ctx, solid.data.data(), solid.shape[0], prim.data.data(), prim.shape[0], node.data.data(), node.shape[0],
                                   plan.empty() ? nullptr : plan.data.data(), plan.empty() ? 0 : plan.shape[0], itra.data.data(), itra.shape[0],
                                   inst.data.data(), inst.shape[0]);
        if (rc) return rc;
        return phox_set_tables(ctx, bnd.as<float>(), bnd.shape[0], bnd.shape[3], 60.f, 1.f, optical.as<int32_t>(), icdf.empty() ? nullptr : icdf.as<float>(),
                               icdf.empty() ? 0 : 3, icdf.empty() ? 0 : icdf.count() / 3, 20);
    }
};

int main(int argc, char** argv) {
    std::string geom, gsfile, ipfile, out;
    std::vector<int> devices;
    int ngpu = 0, events = 1, max_bounce = -1;
    unsigned long long seed = 0;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto next = [&]() -> std::string { if (i + 1 >= argc) { std::cerr << "missing value for " << a << std::endl; std::exit(2); } return argv[++i]; };
        if (a == "-g") geom = next();
        else if (a == "-G") gsfile = next();
        else if (a == "-I") ipfile = next();
        else if (a == "-o") out = next();
        else if (a == "--gpus") ngpu = std::atoi(next().c_str());
        else if (a == "--devices") { std::stringstream ss(next()); std::string t; while (std::getline(ss, t, ',')) devices.push_back(std::atoi(t.c_str())); }
        else if (a == "--events") events = std::atoi(next().c_str());
        else if (a == "--max-bounce") max_bounce = std::atoi(next().c_str());
        else if (a == "--seed") seed = std::strtoull(next().c_str(), nullptr, 10);
        else { std::cerr << "unknown option " << a << std::endl; return 2; }
    }
    if (geom.empty() || gsfile.empty()) {
        std::cerr << "usage: PhoxMultiGPU -g <geometry dir> -G <gensteps.npy> [-I input_photons.npy] [--gpus N | --devices a,b,..] [--events E] [-o hits.npy]" << std::endl;
        return 2;
    }
    try {
        const int visible = phox_device_count();
        if (visible == 0) throw std::runtime_error("no CUDA device: this engine has no CPU path");
        if (devices.empty()) {
            if (ngpu <= 0) ngpu = visible;
            for (int d = 0; d < ngpu; d++) devices.push_back(d % visible);
        }
        Geometry G(geom);
        phoxnpy::Array gs = phoxnpy::load(gsfile), ip;
        if (gs.dtype != "<f4" || gs.count() % 24 != 0) throw std::runtime_error("gensteps must be float32 (n,6,4)");
        const int64_t ngs = gs.count() / 24;
        if (!ipfile.empty()) {
            ip = phoxnpy::load(ipfile);
            if (ip.dtype != "<f4" || ip.count() % 16 != 0) throw std::runtime_error("input photons must be float32 (n,4,4)");
        }
        const int64_t nip = ip.empty() ? 0 : ip.count() / 16;
        PhoxMultiGPU mg(devices, [&](phox_context* ctx) {
            int rc = G.upload(ctx);
            if (rc) return rc;
            phox_config c;
            phox_get_config(ctx, &c);
            if (max_bounce >= 0) c.max_bounce = max_bounce;
            c.rng_seed = seed;
            return phox_set_config(ctx, &c);
        });
        int64_t photons = 0;
        for (int64_t i = 0; i < ngs; i++) { uint32_t n; std::memcpy(&n, gs.data.data() + i * 96 + 12, 4); photons += n; }
        for (int k = 0; k < 2; k++) mg.submit(gs.data.data(), ngs, nip ? ip.data.data() : nullptr, nip, k);      // warm-up: both hit buffers and both staging buffers reach their size
        mg.wait();
        const auto t0 = std::chrono::steady_clock::now();
        int64_t nhit = 0;
        uint64_t rays = 0;
        for (int e = 0; e < events; e++) {
            nhit = mg.submit(gs.data.data(), ngs, nip ? ip.data.data() : nullptr, nip, e);
            rays += mg.stats().num_ray;
        }
        mg.wait();
        const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (!out.empty()) phoxnpy::save(out, "<f4", {nhit, 4, 4}, mg.hits(), (size_t)nhit * 64);
        std::cout << "{\"app\": \"PhoxMultiGPU\", \"gpus\": " << devices.size() << ", \"events\": " << events << ", \"photons_per_event\": " << photons
                  << ", \"hits_last_event\": " << nhit << ", \"seconds\": " << dt << ", \"photons_per_s\": " << (double)photons * events / dt
                  << ", \"rays_per_s\": " << (double)rays / dt << ", \"timing\": \"host wall clock around E events, hit copies of event k overlapped with event k+1, last copy waited for\"}"
                  << std::endl;
        std::cout << "Opticks: NumHits:  " << nhit << std::endl;
    } catch (const std::exception& e) {
        std::cerr << "PhoxMultiGPU: " << e.what() << std::endl;
        return 1;
    }
    return 0;
}
