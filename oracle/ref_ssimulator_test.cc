// oracle/ref_ssimulator_test.cc : TEST INFRASTRUCTURE - never linked into or called by the product.
//
// Drives include/PhoxSimulator.h the way the reference's QSim::simulate drives its back end, using the REFERENCE's own
// headers where they lie under /root/reference/sysrap (nothing copied): SSimulator.h (the launcher protocol), SComp.h
// (SCompProvider + component enum), NP.hh (arrays), sslice.h (launch slices).  The object is only touched through
// `SSimulator*` and `const SCompProvider*` after its creation, plus the one call that stands in for
// QEvt::setGenstepUpload_NP(igs, &sl) (qudarap/QSim.cc:479-486).
//
//   1. whole event in one launch      : cx->simulate_launch(); hit = provider->gatherComponent(SCOMP_HIT)
//   2. the same event sliced by max_slot with the rule of SGenstep::GetGenstepSlices (sysrap/SGenstep.h:249-323; that header
//      needs glm, which is absent here, so the ~20 lines are restated below and checked with sslice::TotalPhoton), one
//      simulate_launch + gather per slice, concatenated like NPFold::concat of QSim::simulate (QSim.cc:525-535)
//   3. bytes of 2 == bytes of 1, hit indices ascending; a second event with the same hit count returns ITS hits
//      (stale getHit cache); reset() empties the provider.
// Prints "PASS ..." and exits 0, or the first failed check and exits 1.  Built by oracle/Makefile into oracle/_ref/.
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "NP.hh"
#include "SComp.h"
#include "SSimulator.h"
#include "sslice.h"
#include "PhoxSimulator.h"

static unsigned numphoton(const NP* gs, size_t i) { return ((const unsigned*)gs->bytes())[24 * i + 3]; }

// SGenstep::GetGenstepSlices : greedy, whole gensteps, in order
static void genstep_slices(std::vector<sslice>& slice, const NP* gs, size_t max_slot) {
    sslice sl = {};
    size_t n = gs->shape[0];
    for (size_t i = 0; i < n; i++) {
        size_t num = numphoton(gs, i);
        if (sl.ph_count + num <= max_slot) { sl.gs_stop = i + 1; sl.ph_count += num; }
        else { sl.gs_stop = i; slice.push_back(sl); sl.ph_count = num; sl.gs_start = i; sl.gs_stop = i + 1; }
        if (i == n - 1) slice.push_back(sl);
    }
    sslice::SetOffset(slice);
}

#define CHECK(cond, msg) do { if (!(cond)) { std::printf("FAIL %s (%s:%d)\n", msg, __FILE__, __LINE__); return 1; } } while (0)

int main(int argc, char** argv) {
    if (argc < 4) { std::printf("usage: %s <geometry dir> <genstep.npy> <max_slot> [out_hit.npy]\n", argv[0]); return 2; }
    std::string geom = argv[1], fd = geom + "/CSGFoundry/", ss = fd + "SSim/stree/standard/";
    NP* solid = NP::Load((fd + "solid.npy").c_str()); NP* prim = NP::Load((fd + "prim.npy").c_str()); NP* node = NP::Load((fd + "node.npy").c_str());
    NP* itra = NP::Load((fd + "itra.npy").c_str()); NP* inst = NP::Load((fd + "inst.npy").c_str());
    NP* plan = NP::Exists((fd + "plan.npy").c_str()) ? NP::Load((fd + "plan.npy").c_str()) : nullptr;
    NP* bnd = NP::Load((ss + "bnd.npy").c_str()); NP* optical = NP::Load((ss + "optical.npy").c_str());
    NP* icdf = NP::Exists((ss + "icdf.npy").c_str()) ? NP::Load((ss + "icdf.npy").c_str()) : nullptr;
    NP* igs = NP::Load(argv[2]);
    CHECK(solid && prim && node && itra && inst && bnd && optical && igs, "loading arrays");
    size_t max_slot = std::stoull(argv[3]);

    SSimulator* cx = PhoxSimulator::Create(solid->bytes(), solid->shape[0], prim->bytes(), prim->shape[0], node->bytes(), node->shape[0],
                                           plan ? plan->bytes() : nullptr, plan ? plan->shape[0] : 0, itra->bytes(), itra->shape[0],
                                           inst->bytes(), inst->shape[0], bnd->cvalues<float>(), bnd->shape[0], bnd->shape[3], 60.f, 1.f,
                                           optical->cvalues<int>(), icdf ? icdf->cvalues<float>() : nullptr, icdf ? 3 : 0,
                                           icdf ? icdf->shape[1] : 0, 20);
    const SCompProvider* provider = dynamic_cast<const SCompProvider*>(cx);          // what SEvt::setCompProvider would be given
    PhoxSimulator* input = dynamic_cast<PhoxSimulator*>(cx);                         // only for the genstep hand-over
    CHECK(provider && input, "PhoxSimulator implements SSimulator and SCompProvider");
    CHECK(std::strcmp(provider->getTypeName(), "PhoxSimulator") == 0 && std::strlen(cx->desc()) > 0, "names");
    const int eventID = 3;

    // 1. one launch
    input->setGenstep(igs->bytes(), igs->shape[0]);
    double dt = cx->simulate(eventID, false);
    CHECK(dt >= 0., "simulate returns launch seconds");
    NP* hit_whole = provider->gatherComponent(SCOMP_HIT);
    CHECK(hit_whole && hit_whole->shape.size() == 3 && hit_whole->shape[1] == 4 && hit_whole->shape[2] == 4, "hit array shape (n,4,4)");
    NP* gs_back = provider->gatherComponent(SCOMP_GENSTEP);
    CHECK(gs_back && gs_back->shape[0] == igs->shape[0] && std::memcmp(gs_back->bytes(), igs->bytes(), igs->arr_bytes()) == 0, "genstep component");
    CHECK(provider->gatherComponent(SCOMP_PHOTON) == nullptr, "Minimal event mode keeps no photon array");
    cx->reset(eventID);
    CHECK(provider->gatherComponent(SCOMP_HIT) == nullptr, "reset empties the provider");
    CHECK(cx->simulate(eventID, false) == -1., "simulate without gensteps returns -1. like QSim::simulate");

    // 2. sliced like QSim::simulate
    std::vector<sslice> slices;
    genstep_slices(slices, igs, max_slot);
    size_t tot = 0;
    for (size_t i = 0; i < (size_t)igs->shape[0]; i++) tot += numphoton(igs, i);
    CHECK(sslice::TotalPhoton(slices) == tot && slices.size() > 1, "slices cover the event");
    std::vector<char> cat;
    size_t nhit = 0;
    for (const sslice& sl : slices) {
        CHECK(sl.ph_count <= max_slot, "slice within max_slot");
        input->setGenstepSlice(igs->bytes(), (int64_t)sl.gs_start, (int64_t)sl.gs_stop, sl.ph_offset);
        double d = cx->simulate(eventID, false);
        CHECK(d >= 0., "slice launch");
        NP* h = provider->gatherComponent(SCOMP_HIT);
        if (h) { cat.insert(cat.end(), (const char*)h->bytes(), (const char*)h->bytes() + h->arr_bytes()); nhit += h->shape[0]; delete h; }
    }
    CHECK(nhit == (size_t)hit_whole->shape[0], "sliced event has the same number of hits");
    CHECK(std::memcmp(cat.data(), hit_whole->bytes(), cat.size()) == 0, "sliced event == single launch, byte for byte");
    const unsigned* hu = (const unsigned*)hit_whole->bytes();
    for (size_t i = 1; i < nhit; i++) CHECK(hu[16 * i + 14] > hu[16 * (i - 1) + 14], "hit photon indices ascend");

    // 3. two events with the same number of hits: getHit must not serve the first event's copy
    input->setGenstep(igs->bytes(), igs->shape[0]);
    cx->simulate(eventID, false);
    PhoxPhoton a, b;
    input->getHit(a, 0);
    input->setGenstepSlice(igs->bytes(), 0, igs->shape[0], 1000000ull);              // same photons, other absolute indices -> other streams
    cx->simulate(eventID, false);
    input->getHit(b, 0);
    unsigned ia, ib; std::memcpy(&ia, &a.q[14], 4); std::memcpy(&ib, &b.q[14], 4);
    CHECK(ib >= 1000000u && ia < 1000000u, "getHit after a second simulate returns the second event's hits");

    if (argc > 4) hit_whole->save(argv[4]);
    std::printf("PASS hits %zu slices %zu photons %zu launch %.4f s : %s\n", nhit, slices.size(), tot, dt, cx->desc());
    delete cx;
    return 0;
}
