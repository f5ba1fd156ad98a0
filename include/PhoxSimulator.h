// PhoxSimulator.h : C++ host adaptor that puts libphox.so behind the reference's SSimulator protocol.
//
// The reference drives its GPU back end through the pure-virtual SSimulator interface
// (sysrap/SSimulator.h:16-35): QSim holds an `SSimulator* cx` and calls cx->simulate_launch() after it
// has uploaded the gensteps (qudarap/QSim.cc:360, 508); G4CXOpticks calls cx->simulate(eventID, reset)
// and cx->reset(eventID) (g4cx/G4CXOpticks.cc:480, 513); CSGOptiX is the one implementation
// (CSGOptiX/CSGOptiX.h:59, factory CSGOptiX::Create(CSGFoundry*) CSGOptiX.cc:367).
//
// PhoxSimulator is the replacement implementation.  Inside the reference tree it derives from the
// reference's own SSimulator (found via __has_include); standalone it derives from an identical local
// declaration so the header compiles anywhere.  It is header-only and speaks to the engine purely
// through the C ABI of phox.h, so the reference needs no CUDA code of ours at compile time.
//
//   PhoxSimulator* cx = PhoxSimulator::Create(solid, nsolid, prim, nprim, node, nnode, plan, nplan, itra, nitra, inst, ninst,
//                                             bnd, nbnd, nwl, 60.f, 1.f, optical, icdf, 3, 4096, 20);
//   cx->setGenstep(gs, ngs);  /* or setInputPhoton */   double dt = cx->simulate(eventID, false);
//   unsigned nhit = cx->getNumHit();  cx->getHit(hit, i);  cx->reset(eventID);
//
// Hits reach the apps through SEvt (SEvt::GetNumHit / getHit read the "hit" array of the event's NPFold, sysrap/SEvt.cc:4924-4991),
// and SEvt fills that fold by asking its SCompProvider for each component (SEvt::gather_components, SEvt.cc:4041-4078; the provider
// is QEvt on the GPU path, set in QEvt::init with SEvt::setCompProvider, SEvt.cc:1294).  Inside the reference tree PhoxSimulator
// therefore ALSO implements SCompProvider (sysrap/SComp.h:50-55): gatherComponent(SCOMP_HIT | PHOTON | RECORD | SEQ | PRD | TAG | FLAT |
// GENSTEP | HITLITE | HITMERGED | HITLITEMERGED | SIMTRACE) hands SEvt freshly allocated NP arrays copied from the engine, so
// SEvt::gather, SEvt::getHit, SEvt::save and the apps on top of them work unchanged.  docs/reference_side.patch shows the hunks
// in QSim.cc / G4CXOpticks.cc; oracle/ref_ssimulator_test.cc drives this class through SSimulator* and SCompProvider* only.
#pragma once
#include <cstdint>
#include <chrono>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "phox.h"

#if defined(__has_include)
#if __has_include("SSimulator.h")
#include "SSimulator.h"
#define PHOX_HAVE_REFERENCE_SSIMULATOR 1
#endif
#endif

#if defined(__has_include)
#if __has_include("SComp.h") && __has_include("NP.hh")
#include "NP.hh"
#include "SComp.h"
#define PHOX_HAVE_REFERENCE_SCOMP 1
#endif
#endif

#ifndef PHOX_HAVE_REFERENCE_SSIMULATOR
struct SSimulator {                       // sysrap/SSimulator.h:16-35, same virtuals in the same order
    virtual ~SSimulator() = default;
    virtual double render_launch() = 0;
    virtual double simtrace_launch() = 0;
    virtual double simulate_launch() = 0;
    virtual double launch() = 0;
    virtual const char* desc() const = 0;
    virtual double simulate(int eventID, bool reset = false) = 0;
    virtual double simtrace(int eventID) = 0;
    virtual double render(const char* stem = nullptr) = 0;
    virtual void reset(int eventID) = 0;
};
#endif

struct PhoxPhoton { float q[16]; };       // sphoton, 64 bytes

class PhoxSimulator : public SSimulator
#ifdef PHOX_HAVE_REFERENCE_SCOMP
    , public SCompProvider
#endif
{
public:
    static PhoxSimulator* Create(const void* solid, int64_t nsolid, const void* prim, int64_t nprim, const void* node, int64_t nnode,
                                 const void* plan, int64_t nplan, const void* itra, int64_t nitra, const void* inst, int64_t ninst,
                                 const float* bnd, int64_t nbnd, int64_t nwl, float domain_low, float domain_step, const int32_t* optical,
                                 const float* icdf, int64_t icdf_ny, int64_t icdf_nx, int32_t hd_factor, int device = 0) {
        phox_context* ctx = phox_create(device);
        if (!ctx) throw std::runtime_error(std::string("PhoxSimulator::Create: ") + phox_last_error(nullptr));
        PhoxSimulator* cx = new PhoxSimulator(ctx);                   // owns ctx from here on
        try {
            cx->check(phox_set_geometry(ctx, solid, nsolid, prim, nprim, node, nnode, plan, nplan, itra, nitra, inst, ninst));
            cx->check(phox_set_tables(ctx, bnd, nbnd, nwl, domain_low, domain_step, optical, icdf, icdf_ny, icdf_nx, hd_factor));
        } catch (...) {
            delete cx;                                                // destroys the context too
            throw;
        }
        return cx;
    }
    ~PhoxSimulator() override { phox_destroy(ctx_); }

    // event input: what SEvt::AddGenstep / SEvt::SetInputPhoton collect (sysrap/SEvt.cc:2059, 2440-2548)
    void setGenstep(const void* quad6, int64_t n) { gs_.assign((const char*)quad6, (const char*)quad6 + n * 96); ngs_ = n; ip_.clear(); nip_ = 0; ph_offset_ = 0; }
    // one launch slice of the event's gensteps, what QSim::simulate hands to QEvt::setGenstepUpload_NP(igs, &sl) (qudarap/QSim.cc:479-486):
    // gensteps [gs_start, gs_stop) of the whole array and the photons before them (sslice::ph_offset -> absolute photon indices)
    void setGenstepSlice(const void* quad6_all, int64_t gs_start, int64_t gs_stop, uint64_t ph_offset) {
        setGenstep((const char*)quad6_all + gs_start * 96, gs_stop - gs_start);
        ph_offset_ = ph_offset;
    }
    void setInputPhoton(const void* sphoton, int64_t n) {
        ip_.assign((const char*)sphoton, (const char*)sphoton + n * 64); nip_ = n;
        gs_.assign(96, 0); ngs_ = 1;                                  // one OpticksGenstep_INPUT_PHOTON genstep (SEvt.cc:1057-1064)
        int32_t code = 19; uint32_t num = (uint32_t)n;
        std::memcpy(gs_.data(), &code, 4); std::memcpy(gs_.data() + 12, &num, 4);
    }
    void setEventID(int eventID) { event_id_ = eventID; }             // QSim::simulate calls simulate_launch(), which has no eventID argument
    phox_config& config() { return cfg_; }
    void applyConfig() { check(phox_set_config(ctx_, &cfg_)); }

    // ---- SSimulator ------------------------------------------------------------------------------
    double simulate_launch() override {                               // low level: one event from the collected input
        if (ngs_ == 0) return -1.;                                    // QSim::simulate returns -1. without gensteps (QSim.cc:446)
        double dt = 0.;
        hits_valid_ = false;                                          // the cached copy belongs to the previous event
        check(phox_simulate(ctx_, gs_.data(), ngs_, nip_ ? ip_.data() : nullptr, nip_, event_id_, ph_offset_, &dt));
        return dt;
    }
    double launch() override { return simulate_launch(); }
    double simulate(int eventID, bool reset_ = false) override {
        event_id_ = eventID;
        double dt = simulate_launch();
        if (reset_) reset(eventID);
        return dt;
    }
    void reset(int /*eventID*/) override { phox_reset(ctx_); gs_.clear(); ip_.clear(); ngs_ = nip_ = 0; ph_offset_ = 0; hits_.clear(); hits_valid_ = false; }
    const char* desc() const override { return phox_desc(ctx_); }
    // simtrace (CSGOptiX7.cu:536-577): the collected gensteps must be FRAME / INPUT_PHOTON_SIMTRACE ones; the records
    // (sevent::add_simtrace layout) are kept for getSimtrace().  Returns wall seconds, -1. without gensteps.
    double simtrace_launch() override {
        if (ngs_ == 0) return -1.;
        auto t0 = std::chrono::steady_clock::now();
        int64_t n = phox_simtrace(ctx_, gs_.data(), ngs_, nip_ ? ip_.data() : nullptr, nip_, nullptr, 0);
        if (n < 0) check((int)n);
        simtrace_.resize((size_t)n * 16);
        n = phox_simtrace(ctx_, gs_.data(), ngs_, nip_ ? ip_.data() : nullptr, nip_, simtrace_.data(), n);
        if (n < 0) check((int)n);
        return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    }
    double simtrace(int eventID) override { event_id_ = eventID; return simtrace_launch(); }
    void setInputSimtrace(const void* quad4, int64_t n) {                // rays for an INPUT_PHOTON_SIMTRACE genstep (qsim.h:2455)
        ip_.assign((const char*)quad4, (const char*)quad4 + n * 64); nip_ = n;
        gs_.assign(96, 0); ngs_ = 1;
        int32_t code = 20; uint32_t num = (uint32_t)n;
        std::memcpy(gs_.data(), &code, 4); std::memcpy(gs_.data() + 12, &num, 4);
    }
    const std::vector<float>& getSimtrace() const { return simtrace_; }    // (n,4,4) float32
    // render is outside the path this library replaces (SURVEY 2.4 "OUT")
    double render_launch() override { return -1.; }
    double render(const char* = nullptr) override { return -1.; }

    // hits, as SEvt::GetNumHit / SEvt::getHit hand them to the apps (sysrap/SEvt.cc:4924-4925, 4991)
    unsigned getNumHit() const { return (unsigned)phox_num_hit(ctx_); }
    void gatherHits() { hits_.resize((size_t)phox_num_hit(ctx_)); if (!hits_.empty()) check(phox_get_hits(ctx_, hits_.data())); hits_valid_ = true; }
    void getHit(PhoxPhoton& p, unsigned idx) { if (!hits_valid_) gatherHits(); p = hits_.at(idx); }   // first call after a simulate re-gathers
    phox_context* context() { return ctx_; }

#ifdef PHOX_HAVE_REFERENCE_SCOMP
    // ---- SCompProvider (sysrap/SComp.h:50-55): what SEvt::gather_components calls for every component of its gather mask ----
    const char* getTypeName() const override { return "PhoxSimulator"; }
    std::string getMeta() const override { return std::string("provider:PhoxSimulator\n") + phox_desc(ctx_) + "\n"; }
    NP* gatherComponent(unsigned comp) const override {
        phox_context* c = ctx_;
        auto named = [&](const char* name, int itemsize_bytes, auto make) -> NP* {
            int64_t bytes = phox_get_array(c, name, nullptr, 0);
            if (bytes <= 0) return nullptr;                           // not kept in this event mode: SEvt skips null components
            NP* a = make(bytes / itemsize_bytes);
            if (phox_get_array(c, name, a->bytes(), bytes) != bytes) { delete a; return nullptr; }
            return a;
        };
        switch (comp) {
            case SCOMP_HIT:     return named("hit", 64, [](int64_t n) { return NP::Make<float>(n, 4, 4); });          // QEvt::gatherHit
            case SCOMP_PHOTON:  return named("photon", 64, [](int64_t n) { return NP::Make<float>(n, 4, 4); });
            case SCOMP_RECORD: {
                phox_config cf; phox_get_config(c, &cf);
                const int mr = cf.max_record > 0 ? cf.max_record : 1;
                return named("record", 64 * mr, [mr](int64_t n) { return NP::Make<float>(n, mr, 4, 4); });
            }
            case SCOMP_SEQ:     return named("seq", 32, [](int64_t n) { return NP::Make<unsigned long long>(n, 2, 2); });
            case SCOMP_PRD: {
                phox_config cf; phox_get_config(c, &cf);
                const int mr = cf.max_record > 0 ? cf.max_record : 1;
                return named("prd", 32 * mr, [mr](int64_t n) { return NP::Make<float>(n, mr, 2, 4); });
            }
            case SCOMP_TAG:     return named("tag", 32, [](int64_t n) { return NP::Make<unsigned long long>(n, 4); });
            case SCOMP_FLAT:    return named("flat", 256, [](int64_t n) { return NP::Make<float>(n, 64); });
            case SCOMP_GENSTEP: {
                if (ngs_ == 0) return nullptr;
                NP* a = NP::Make<float>(ngs_, 6, 4);
                std::memcpy(a->bytes(), gs_.data(), (size_t)ngs_ * 96);
                return a;
            }
            case SCOMP_INPHOTON: {
                if (nip_ == 0) return nullptr;
                NP* a = NP::Make<float>(nip_, 4, 4);
                std::memcpy(a->bytes(), ip_.data(), (size_t)nip_ * 64);
                return a;
            }
            case SCOMP_SIMTRACE: {
                if (simtrace_.empty()) return nullptr;
                NP* a = NP::Make<float>((int64_t)simtrace_.size() / 16, 4, 4);
                std::memcpy(a->bytes(), simtrace_.data(), simtrace_.size() * 4);
                return a;
            }
            case SCOMP_HITLITE: {                                       // sphotonlite (n,4) u32, QEvt::gatherHitLite_
                int64_t n = phox_num_hit(c);
                phox_config cf; phox_get_config(c, &cf);
                if (n <= 0 || !cf.mode_lite) return nullptr;
                NP* a = NP::Make<unsigned>(n, 4);
                if (phox_get_hits_lite(c, a->bytes()) < 0) { delete a; return nullptr; }
                return a;
            }
            case SCOMP_HITMERGED: {                                     // QEvt::PerLaunchMerge with the configured window
                int64_t m = phox_merge_hits(c, merge_window_, nullptr, 0);
                if (m <= 0) return nullptr;
                NP* a = NP::Make<float>(m, 4, 4);
                if (phox_merge_hits(c, merge_window_, a->bytes(), m) != m) { delete a; return nullptr; }
                return a;
            }
            case SCOMP_HITLITEMERGED: {
                phox_config cf; phox_get_config(c, &cf);
                if (!cf.mode_lite) return nullptr;
                int64_t m = phox_merge_hits_lite(c, merge_window_, nullptr, 0);
                if (m <= 0) return nullptr;
                NP* a = NP::Make<unsigned>(m, 4);
                if (phox_merge_hits_lite(c, merge_window_, a->bytes(), m) != m) { delete a; return nullptr; }
                return a;
            }
            default: return nullptr;                                    // seed, domain, aux, sup ...: not produced by this back end
        }
    }
    void setMergeWindow(float ns) { merge_window_ = ns; }               // SEventConfig::MergeWindow
#endif

private:
    explicit PhoxSimulator(phox_context* ctx) : ctx_(ctx) { phox_default_config(&cfg_); }
    void check(int rc) { if (rc < 0) throw std::runtime_error(std::string("phox: ") + phox_last_error(ctx_)); }
    phox_context* ctx_;
    phox_config cfg_;
    std::vector<char> gs_, ip_;
    int64_t ngs_ = 0, nip_ = 0;
    int event_id_ = 0;
    uint64_t ph_offset_ = 0;
    float merge_window_ = 1.f;
    bool hits_valid_ = false;
    std::vector<PhoxPhoton> hits_;
    std::vector<float> simtrace_;
};
