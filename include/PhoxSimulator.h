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

class PhoxSimulator : public SSimulator {
public:
    static PhoxSimulator* Create(const void* solid, int64_t nsolid, const void* prim, int64_t nprim, const void* node, int64_t nnode,
                                 const void* plan, int64_t nplan, const void* itra, int64_t nitra, const void* inst, int64_t ninst,
                                 const float* bnd, int64_t nbnd, int64_t nwl, float domain_low, float domain_step, const int32_t* optical,
                                 const float* icdf, int64_t icdf_ny, int64_t icdf_nx, int32_t hd_factor, int device = 0) {
        phox_context* ctx = phox_create(device);
        if (!ctx) throw std::runtime_error(std::string("PhoxSimulator::Create: ") + phox_last_error(nullptr));
        PhoxSimulator* cx = new PhoxSimulator(ctx);
        cx->check(phox_set_geometry(ctx, solid, nsolid, prim, nprim, node, nnode, plan, nplan, itra, nitra, inst, ninst));
        cx->check(phox_set_tables(ctx, bnd, nbnd, nwl, domain_low, domain_step, optical, icdf, icdf_ny, icdf_nx, hd_factor));
        return cx;
    }
    ~PhoxSimulator() override { phox_destroy(ctx_); }

    // event input: what SEvt::AddGenstep / SEvt::SetInputPhoton collect (sysrap/SEvt.cc:2059, 2440-2548)
    void setGenstep(const void* quad6, int64_t n) { gs_.assign((const char*)quad6, (const char*)quad6 + n * 96); ngs_ = n; ip_.clear(); nip_ = 0; }
    void setInputPhoton(const void* sphoton, int64_t n) {
        ip_.assign((const char*)sphoton, (const char*)sphoton + n * 64); nip_ = n;
        gs_.assign(96, 0); ngs_ = 1;                                  // one OpticksGenstep_INPUT_PHOTON genstep (SEvt.cc:1057-1064)
        int32_t code = 19; uint32_t num = (uint32_t)n;
        std::memcpy(gs_.data(), &code, 4); std::memcpy(gs_.data() + 12, &num, 4);
    }
    phox_config& config() { return cfg_; }
    void applyConfig() { check(phox_set_config(ctx_, &cfg_)); }

    // ---- SSimulator ------------------------------------------------------------------------------
    double simulate_launch() override {                               // low level: one event from the collected input
        if (ngs_ == 0) return -1.;                                    // QSim::simulate returns -1. without gensteps (QSim.cc:446)
        double dt = 0.;
        check(phox_simulate(ctx_, gs_.data(), ngs_, nip_ ? ip_.data() : nullptr, nip_, event_id_, 0, &dt));
        return dt;
    }
    double launch() override { return simulate_launch(); }
    double simulate(int eventID, bool reset_ = false) override {
        event_id_ = eventID;
        double dt = simulate_launch();
        if (reset_) reset(eventID);
        return dt;
    }
    void reset(int /*eventID*/) override { phox_reset(ctx_); gs_.clear(); ip_.clear(); ngs_ = nip_ = 0; }
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
    void gatherHits() { hits_.resize((size_t)phox_num_hit(ctx_)); if (!hits_.empty()) check(phox_get_hits(ctx_, hits_.data())); }
    void getHit(PhoxPhoton& p, unsigned idx) { if (hits_.size() != (size_t)phox_num_hit(ctx_)) gatherHits(); p = hits_.at(idx); }
    phox_context* context() { return ctx_; }

private:
    explicit PhoxSimulator(phox_context* ctx) : ctx_(ctx) { phox_default_config(&cfg_); }
    void check(int rc) { if (rc < 0) throw std::runtime_error(std::string("phox: ") + phox_last_error(ctx_)); }
    phox_context* ctx_;
    phox_config cfg_;
    std::vector<char> gs_, ip_;
    int64_t ngs_ = 0, nip_ = 0;
    int event_id_ = 0;
    std::vector<PhoxPhoton> hits_;
    std::vector<float> simtrace_;
};
