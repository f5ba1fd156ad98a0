// PhoxMultiGPU.h : one event over the GPUs of a box, in C++ on the C ABI (include/phox.h).
//
// The reference has no multi-GPU path; what it has is QSim::simulate slicing an event into sequential launches at
// genstep granularity, each launch told the absolute index of its first photon (sysrap/SGenstep.h:249-323
// SGenstep::GetGenstepSlices, qudarap/QSim.cc:479-528, CSGOptiX/CSGOptiX7.cu:415-419), so that the concatenated slices
// equal one launch.  This class runs the same slices CONCURRENTLY, one per GPU:
//
//   * one worker thread and one phox_context per device; geometry and tables are replicated (set through `setup`);
//   * gensteps are cut into contiguous ranges balanced by photon count (slices(), the concurrent form of
//     GetGenstepSlices: whole gensteps, in order), every range with its absolute photon offset; an input-photon event
//     (one INPUT_PHOTON genstep) is cut by photon range instead;
//   * hits land in ONE page-locked host buffer at the prefix offsets of the per-device hit counts, i.e. in ascending
//     photon index: byte for byte the array a single GPU returns (tests/test_parity_gpu.py drives the app built on this);
//   * the device-to-host copies of event k run on each context's second stream (phox_get_hits_async) while event k + 1
//     is being simulated: submit() returns as soon as the copies are posted, hits() of an event is valid after wait().
//
// No collective is involved: the only exchange is the hit counts between host threads.
#pragma once
#include <algorithm>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <functional>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "phox.h"

class PhoxMultiGPU {
public:
    struct Slice { int64_t gs_start, gs_stop; uint64_t ph_offset; int64_t ph_count; };

    // Contiguous genstep ranges balanced by photon count, one per rank, whole gensteps, in order.
    static std::vector<Slice> slices(const void* quad6, int64_t ngs, int nrank) {
        std::vector<int64_t> csum((size_t)ngs + 1, 0);
        for (int64_t i = 0; i < ngs; i++) {
            uint32_t n;
            std::memcpy(&n, (const char*)quad6 + i * 96 + 12, 4);            // q0.w = numphoton
            csum[i + 1] = csum[i] + n;
        }
        const int64_t total = csum[ngs];
        std::vector<Slice> out;
        int64_t start = 0;
        for (int r = 0; r < nrank; r++) {
            int64_t stop = ngs;
            if (r < nrank - 1) {
                const int64_t target = total * (r + 1) / nrank;
                stop = std::lower_bound(csum.begin(), csum.end(), target) - csum.begin();
                if (stop < start) stop = start;
                if (stop > ngs) stop = ngs;
            }
            out.push_back({start, stop, (uint64_t)csum[start], csum[stop] - csum[start]});
            start = stop;
        }
        return out;
    }

    // devices: CUDA device index of every rank (a device may appear twice: two contexts share it - what the single-GPU test does).
    // setup(ctx) uploads geometry + tables + config into one context and returns a PHOX_* code.
    PhoxMultiGPU(const std::vector<int>& devices, const std::function<int(phox_context*)>& setup) {
        if (devices.empty()) throw std::runtime_error("PhoxMultiGPU: no devices");
        ranks_.resize(devices.size());
        for (size_t r = 0; r < devices.size(); r++) {
            Rank& k = ranks_[r];
            k.ctx = phox_create(devices[r]);
            if (!k.ctx) { std::string m = phox_last_error(nullptr); close(); throw std::runtime_error("PhoxMultiGPU: " + m); }
            const int rc = setup(k.ctx);
            if (rc != PHOX_OK) { std::string m = phox_last_error(k.ctx); close(); throw std::runtime_error("PhoxMultiGPU: setup failed: " + m); }
        }
        for (size_t r = 0; r < ranks_.size(); r++) ranks_[r].th = std::thread([this, r] { worker((int)r); });
    }
    ~PhoxMultiGPU() { close(); }
    PhoxMultiGPU(const PhoxMultiGPU&) = delete;
    PhoxMultiGPU& operator=(const PhoxMultiGPU&) = delete;

    int num_rank() const { return (int)ranks_.size(); }
    phox_context* context(int r) { return ranks_[r].ctx; }

    // One event.  Blocks until every device has simulated its share and POSTED the copy of its hits into the buffer of
    // this event (two buffers alternate); returns the event's hit count.  The arrays must stay valid during the call.
    int64_t submit(const void* quad6, int64_t ngs, const void* input_photon, int64_t ninput, int event_id) {
        Job j;
        j.gs = quad6; j.ngs = ngs; j.ip = input_photon; j.nip = ninput; j.event_id = event_id;
        const int nr = num_rank();
        j.parts.resize(nr);
        if (input_photon && ninput > 0) {                     // one INPUT_PHOTON genstep: cut by photon range
            for (int r = 0; r < nr; r++) {
                const int64_t lo = ninput * r / nr, hi = ninput * (r + 1) / nr;
                j.parts[r] = {0, 1, (uint64_t)lo, hi - lo};
            }
        } else j.parts = slices(quad6, ngs, nr);
        buf_ ^= 1;
        {
            std::unique_lock<std::mutex> lk(m_);
            job_ = j; counts_.assign(nr, -1); arrived_ = 0; error_.clear();
            phase_++;                                         // workers: simulate
            cv_.notify_all();
            cv_.wait(lk, [&] { return arrived_ == nr; });
            if (!error_.empty()) throw std::runtime_error("PhoxMultiGPU: " + error_);
            // prefix offsets of the hit counts; grow the event's pinned buffer if needed (waits for its previous copies)
            int64_t total = 0;
            offsets_.assign(nr, 0);
            for (int r = 0; r < nr; r++) { offsets_[r] = total; total += counts_[r]; }
            num_hit_[buf_] = total;
            if (total > cap_[buf_]) {
                for (auto& k : ranks_) phox_hits_wait(k.ctx);
                if (hits_[buf_]) phox_host_free(hits_[buf_]);
                cap_[buf_] = total + total / 4 + 1024;
                hits_[buf_] = phox_host_alloc(cap_[buf_] * 64);
                if (!hits_[buf_]) { cap_[buf_] = 0; throw std::runtime_error("PhoxMultiGPU: cannot allocate the pinned hit buffer"); }
            }
            arrived_ = 0;
            phase_++;                                         // workers: post the copies
            cv_.notify_all();
            cv_.wait(lk, [&] { return arrived_ == nr; });
            if (!error_.empty()) throw std::runtime_error("PhoxMultiGPU: " + error_);
        }
        return num_hit_[buf_];
    }
    // every copy posted so far has landed
    void wait() { for (auto& k : ranks_) if (phox_hits_wait(k.ctx) != PHOX_OK) throw std::runtime_error(std::string("PhoxMultiGPU: ") + phox_last_error(k.ctx)); }
    // hits of the last submitted event (sphoton[num_hit()], ascending photon index); valid after wait() until the second-next submit()
    const void* hits() const { return hits_[buf_]; }
    int64_t num_hit() const { return num_hit_[buf_]; }
    const std::vector<int64_t>& counts() const { return counts_; }
    // sum over the devices of the last event's counters
    phox_stats stats() {
        phox_stats t;
        std::memset(&t, 0, sizeof(t));
        for (auto& k : ranks_) {
            phox_stats s;
            phox_get_stats(k.ctx, &s);
            t.num_photon += s.num_photon; t.num_hit += s.num_hit; t.num_ray += s.num_ray; t.num_launch += s.num_launch; t.num_kernel += s.num_kernel;
            t.num_home_ray += s.num_home_ray;
            if (s.simulate_kernel_seconds > t.simulate_kernel_seconds) t.simulate_kernel_seconds = s.simulate_kernel_seconds;
            if (s.launch_seconds > t.launch_seconds) t.launch_seconds = s.launch_seconds;
        }
        return t;
    }

private:
    struct Rank { phox_context* ctx = nullptr; std::thread th; };
    struct Job { const void* gs = nullptr; int64_t ngs = 0; const void* ip = nullptr; int64_t nip = 0; int event_id = 0; std::vector<Slice> parts; };

    void worker(int r) {
        Rank& k = ranks_[r];
        uint64_t seen = 0;
        std::vector<char> one_gs(96);
        while (true) {
            Job j;
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [&] { return phase_ != seen || quit_; });
                if (quit_) return;
                seen = phase_;
                j = job_;
            }
            // phase 1: simulate this rank's share
            const Slice& s = j.parts[r];
            int rc;
            if (j.ip && j.nip > 0) {
                std::memcpy(one_gs.data(), j.gs, 96);
                const uint32_t n = (uint32_t)s.ph_count;
                std::memcpy(one_gs.data() + 12, &n, 4);
                rc = phox_simulate(k.ctx, one_gs.data(), 1, (const char*)j.ip + s.ph_offset * 64, s.ph_count, j.event_id, s.ph_offset, nullptr);
            } else {
                rc = phox_simulate(k.ctx, (const char*)j.gs + s.gs_start * 96, s.gs_stop - s.gs_start, nullptr, 0, j.event_id, s.ph_offset, nullptr);
            }
            {
                std::unique_lock<std::mutex> lk(m_);
                if (rc != PHOX_OK && error_.empty()) error_ = phox_last_error(k.ctx);
                counts_[r] = rc == PHOX_OK ? phox_num_hit(k.ctx) : 0;
                arrived_++;
                cv_.notify_all();
                cv_.wait(lk, [&] { return phase_ != seen || quit_; });
                if (quit_) return;
                seen = phase_;
            }
            // phase 2: post the copy of this rank's hits to its place in the event's buffer
            rc = PHOX_OK;
            if (counts_[r] > 0 && error_.empty()) rc = phox_get_hits_async(k.ctx, (char*)hits_[buf_] + offsets_[r] * 64);
            {
                std::unique_lock<std::mutex> lk(m_);
                if (rc != PHOX_OK && error_.empty()) error_ = phox_last_error(k.ctx);
                arrived_++;
                cv_.notify_all();
            }
        }
    }

    void close() {
        {
            std::unique_lock<std::mutex> lk(m_);
            quit_ = true;
            cv_.notify_all();
        }
        for (auto& k : ranks_) if (k.th.joinable()) k.th.join();
        for (auto& k : ranks_) if (k.ctx) { phox_hits_wait(k.ctx); phox_destroy(k.ctx); k.ctx = nullptr; }
        for (int b = 0; b < 2; b++) if (hits_[b]) { phox_host_free(hits_[b]); hits_[b] = nullptr; }
    }

    std::vector<Rank> ranks_;
    std::mutex m_;
    std::condition_variable cv_;
    uint64_t phase_ = 0;
    int arrived_ = 0;
    bool quit_ = false;
    Job job_;
    std::vector<int64_t> counts_, offsets_;
    std::string error_;
    void* hits_[2] = {nullptr, nullptr};
    int64_t cap_[2] = {0, 0}, num_hit_[2] = {0, 0};
    int buf_ = 0;
};
