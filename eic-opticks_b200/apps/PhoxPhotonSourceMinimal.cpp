// PhoxPhotonSourceMinimal : standalone C++ driver with the contract of the reference's GPUPhotonSourceMinimal
// (src/GPUPhotonSourceMinimal.cpp/.h): a gphox JSON config (config/*.json: "torch" + "event" objects, parsed like
// src/config.cpp:107-145) -> torch photons generated on the host exactly like generate_photons (src/torch.cpp:8-30: one
// curand Philox stream, seed 0, subsequence 0, storch::generate per photon) -> input photons -> simulate ->
// "Opticks: NumHits:  N" and opticks_hits_output.txt.  With --genstep the torch genstep itself is handed to the GPU
// (storch::generate runs per photon on the device, every torch type) instead of host-made input photons.
//
//   PhoxPhotonSourceMinimal -g <geometry dir> -c <config.json> [-o opticks_hits_output.txt] [-s seed] [-d device] [--genstep]
//
// Geant4 is not involved: geometry comes from a persisted CSGFoundry directory.  Host code is C++ on the C ABI.
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <map>
#include <sstream>

#include "phox_app_common.h"

// ---- minimal JSON (objects, arrays of numbers, numbers, strings): all the gphox configs use -------------------------------
struct JValue {
    enum Kind { NUM, STR, ARR, OBJ } kind = NUM;
    double num = 0.;
    std::string str;
    std::vector<double> arr;
    std::map<std::string, JValue> obj;
    const JValue& at(const std::string& k) const {
        auto it = obj.find(k);
        if (it == obj.end()) throw std::runtime_error("config: missing key \"" + k + "\"");
        return it->second;
    }
    bool has(const std::string& k) const { return obj.count(k) != 0; }
};

struct JParser {
    const std::string& s;
    size_t i = 0;
    explicit JParser(const std::string& text) : s(text) {}
    void ws() { while (i < s.size() && (s[i] == ' ' || s[i] == '\n' || s[i] == '\t' || s[i] == '\r')) i++; }
    [[noreturn]] void fail(const char* what) { throw std::runtime_error(std::string("config: JSON parse error: ") + what + " at offset " + std::to_string(i)); }
    std::string string_() {
        if (s[i] != '"') fail("expected string");
        size_t j = s.find('"', i + 1);
        if (j == std::string::npos) fail("unterminated string");
        std::string r = s.substr(i + 1, j - i - 1);
        i = j + 1;
        return r;
    }
    double number_() {
        char* end = nullptr;
        double v = std::strtod(s.c_str() + i, &end);
        if (end == s.c_str() + i) fail("expected number");
        i = (size_t)(end - s.c_str());
        return v;
    }
    JValue value() {
        ws();
        if (i >= s.size()) fail("unexpected end");
        JValue v;
        if (s[i] == '{') {
            v.kind = JValue::OBJ; i++; ws();
            if (s[i] == '}') { i++; return v; }
            while (true) {
                ws(); std::string k = string_(); ws();
                if (s[i] != ':') fail("expected ':'");
                i++;
                v.obj[k] = value(); ws();
                if (s[i] == ',') { i++; continue; }
                if (s[i] == '}') { i++; break; }
                fail("expected ',' or '}'");
            }
        } else if (s[i] == '[') {
            v.kind = JValue::ARR; i++; ws();
            if (s[i] == ']') { i++; return v; }
            while (true) {
                ws(); v.arr.push_back(number_()); ws();
                if (s[i] == ',') { i++; continue; }
                if (s[i] == ']') { i++; break; }
                fail("expected ',' or ']'");
            }
        } else if (s[i] == '"') { v.kind = JValue::STR; v.str = string_(); }
        else { v.kind = JValue::NUM; v.num = number_(); }
        return v;
    }
};

// ---- curand Philox4_32_10, host side (published algorithm, curand conventions; cf. csrc/phox_philox.cuh) -------------------
struct HostPhilox {
    uint32_t ctr[4] = {0, 0, 0, 0}, key[2] = {0, 0}, out[4];
    int pos = 0;
    void block() {
        uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]}, k[2] = {key[0], key[1]};
        for (int r = 0; r < 10; r++) {
            uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
            uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k[0], n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k[1], n3 = (uint32_t)p0;
            c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
            k[0] += 0x9E3779B9u; k[1] += 0xBB67AE85u;
        }
        std::memcpy(out, c, 16);
    }
    void init(uint64_t seed) { key[0] = (uint32_t)seed; key[1] = (uint32_t)(seed >> 32); pos = 0; block(); }    // curand_init(seed, 0, 0)
    float uniform() {                                                                                            // curand_uniform
        uint32_t r = out[pos++];
        if (pos == 4) { if (++ctr[0] == 0 && ++ctr[1] == 0 && ++ctr[2] == 0) ++ctr[3]; block(); pos = 0; }
        return r * 2.3283064365386963e-10f + (2.3283064365386963e-10f / 2.0f);
    }
};

struct Torch {
    int gentype = 6, trackid = 0, matline = 0, numphoton = 0;
    float pos[3], time = 0.f, mom[3], weight = 0.f, pol[3], wavelength = 0.f, zenith[2] = {0.f, 1.f}, azimuth[2] = {0.f, 1.f}, radius = 0.f, distance = 0.f;
    unsigned mode = 255, type = 1;
};

static unsigned torch_type(const std::string& n) {      // storchtype::Type (sysrap/storchtype.h)
    static const char* names[] = {"undef", "disc", "line", "point", "circle", "rectangle", "sphere_marsaglia", "sphere"};
    for (unsigned k = 0; k < 8; k++) if (n == names[k]) return k;
    throw std::runtime_error("config: unknown torch type \"" + n + "\"");
}

static Torch read_config(const std::string& path, std::string* event_mode, long long* maxslot) {
    std::ifstream ifs(path);
    if (!ifs.is_open()) throw std::runtime_error("Could not find config file \"" + path + "\"");
    std::stringstream ss; ss << ifs.rdbuf();
    std::string text = ss.str();
    JParser jp(text);
    JValue root = jp.value();
    const JValue& t = root.at("torch");
    Torch q;
    if (t.at("gentype").str != "TORCH") throw std::runtime_error("config: torch.gentype must be TORCH");
    q.trackid = (int)t.at("trackid").num; q.matline = (int)t.at("matline").num; q.numphoton = (int)t.at("numphoton").num;
    for (int k = 0; k < 3; k++) { q.pos[k] = (float)t.at("pos").arr.at(k); q.mom[k] = (float)t.at("mom").arr.at(k); q.pol[k] = (float)t.at("pol").arr.at(k); }
    float inv = 1.f / sqrtf(q.mom[0] * q.mom[0] + q.mom[1] * q.mom[1] + q.mom[2] * q.mom[2]);       // normalize (scuda.h: v * (1/sqrtf(dot)))
    for (int k = 0; k < 3; k++) q.mom[k] *= inv;
    q.time = (float)t.at("time").num; q.weight = (float)t.at("weight").num; q.wavelength = (float)t.at("wavelength").num;
    for (int k = 0; k < 2; k++) { q.zenith[k] = (float)t.at("zenith").arr.at(k); q.azimuth[k] = (float)t.at("azimuth").arr.at(k); }
    q.radius = (float)t.at("radius").num; q.distance = (float)t.at("distance").num; q.mode = (unsigned)t.at("mode").num;
    q.type = torch_type(t.at("type").str);
    if (root.has("event")) {
        const JValue& e = root.at("event");
        if (e.has("mode")) *event_mode = e.at("mode").str;
        if (e.has("maxslot")) *maxslot = (long long)e.at("maxslot").num;
    }
    return q;
}

static void rotate_uz(float* d, const float* u) {       // smath::rotateUz (sysrap/smath.h:77-95)
    float up = u[0] * u[0] + u[1] * u[1];
    if (up > 0.f) {
        up = sqrtf(up);
        float px = d[0], py = d[1], pz = d[2];
        d[0] = (u[0] * u[2] * px - u[1] * py) / up + u[0] * pz;
        d[1] = (u[1] * u[2] * px + u[0] * py) / up + u[1] * pz;
        d[2] = -up * px + u[2] * pz;
    } else if (u[2] < 0.f) { d[0] = -d[0]; d[2] = -d[2]; }
}

// generate_photons (src/torch.cpp:8-30) for the disc type (storch::generate T_DISC, sysrap/storch.h:200-240)
static std::vector<PhoxPhoton> generate_photons(const Torch& t, unsigned seed) {
    if (t.type != 1) throw std::runtime_error("host torch generation covers the disc type; use --genstep for the others");
    HostPhilox rng; rng.init(seed);
    std::vector<PhoxPhoton> out((size_t)t.numphoton);
    for (auto& p : out) {
        std::memset(&p, 0, sizeof(p));
        float u_zenith = t.zenith[0] + rng.uniform() * (t.zenith[1] - t.zenith[0]);
        float u_azimuth = t.azimuth[0] + rng.uniform() * (t.azimuth[1] - t.azimuth[0]);
        float r = t.radius * u_zenith, phi = 2.f * (float)M_PI * u_azimuth;
        float sinPhi = sinf(phi), cosPhi = cosf(phi);
        float pos[3] = {r * cosPhi, r * sinPhi, 0.f}, pol[3] = {sinPhi, -cosPhi, 0.f};
        rotate_uz(pos, t.mom); rotate_uz(pol, t.mom);
        for (int k = 0; k < 3; k++) { p.q[k] = pos[k] + t.pos[k]; p.q[4 + k] = t.mom[k]; p.q[8 + k] = pol[k]; }
        p.q[3] = t.time; p.q[11] = t.wavelength;
        uint32_t flag = 4u;                                  // zero_flags(); set_flag(TORCH)
        std::memcpy(&p.q[12], &flag, 4); std::memcpy(&p.q[15], &flag, 4);
    }
    return out;
}

static void torch_genstep(const Torch& t, float* gs) {   // storch as quad6 (sysrap/storch.h:43-70)
    std::memset(gs, 0, 96);
    uint32_t u[4] = {(uint32_t)t.gentype, (uint32_t)t.trackid, (uint32_t)t.matline, (uint32_t)t.numphoton};
    std::memcpy(gs, u, 16);
    for (int k = 0; k < 3; k++) { gs[4 + k] = t.pos[k]; gs[8 + k] = t.mom[k]; gs[12 + k] = t.pol[k]; }
    gs[7] = t.time; gs[11] = t.weight; gs[15] = t.wavelength;
    gs[16] = t.zenith[0]; gs[17] = t.zenith[1]; gs[18] = t.azimuth[0]; gs[19] = t.azimuth[1];
    gs[20] = t.radius; gs[21] = t.distance;
    std::memcpy(&gs[22], &t.mode, 4); std::memcpy(&gs[23], &t.type, 4);
}

int main(int argc, char** argv) {
    std::string geom, config, out = "opticks_hits_output.txt";
    int device = 0;
    unsigned seed = 0;
    bool genstep = false;
    std::string dump;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        if ((a == "-g" || a == "--geometry") && i + 1 < argc) geom = argv[++i];
        else if ((a == "-c" || a == "--config") && i + 1 < argc) config = argv[++i];
        else if ((a == "-o" || a == "--output") && i + 1 < argc) out = argv[++i];
        else if ((a == "-s" || a == "--seed") && i + 1 < argc) seed = (unsigned)std::strtoul(argv[++i], nullptr, 10);
        else if ((a == "-d" || a == "--device") && i + 1 < argc) device = std::atoi(argv[++i]);
        else if (a == "--genstep") genstep = true;
        else if (a == "--dump-photons" && i + 1 < argc) dump = argv[++i];     // write the host-generated photons as 16 floats per line and stop (no GPU needed)
    }
    if (!dump.empty() && !config.empty()) {
        try {
            std::string em; long long ms = 0;
            Torch t = read_config(config, &em, &ms);
            std::vector<PhoxPhoton> ph = generate_photons(t, seed);
            std::ofstream of(dump);
            of.precision(9);
            for (const auto& p : ph) { for (int k = 0; k < 12; k++) of << p.q[k] << " "; uint32_t u[4]; std::memcpy(u, &p.q[12], 16); of << u[0] << " " << u[1] << " " << u[2] << " " << u[3] << "\n"; }
            std::cout << "Dumped " << ph.size() << " photons, event mode " << em << " maxslot " << ms << std::endl;
            return 0;
        } catch (const std::exception& e) { std::cerr << "ERROR: " << e.what() << std::endl; return 2; }
    }
    if (geom.empty() || config.empty()) {
        std::cerr << "usage: PhoxPhotonSourceMinimal -g <geometry dir> -c <config.json> [-o hits.txt] [-s seed] [-d device] [--genstep]" << std::endl;
        return 1;
    }
    try {
        std::string event_mode = "Minimal";
        long long maxslot = 0;
        Torch t = read_config(config, &event_mode, &maxslot);
        PhoxSimulator* cx = phoxapp::create_from_geometry_dir(geom, device);
        std::cout << cx->desc() << std::endl;
        cx->config().max_slot = maxslot;
        cx->config().event_mode = event_mode == "DebugLite" ? PHOX_MODE_DEBUGLITE : event_mode == "DebugHeavy" ? PHOX_MODE_DEBUGHEAVY
                                  : event_mode == "HitPhoton" ? PHOX_MODE_HITPHOTON : event_mode == "HitPhotonSeq" ? PHOX_MODE_HITPHOTONSEQ : PHOX_MODE_MINIMAL;
        cx->applyConfig();
        if (genstep) {
            float gs[24];
            torch_genstep(t, gs);
            cx->setGenstep(gs, 1);
            std::cout << "Torch genstep: " << t.numphoton << " photons generated on the device" << std::endl;
        } else {
            std::vector<PhoxPhoton> ph = generate_photons(t, seed);
            cx->setInputPhoton(ph.data(), (int64_t)ph.size());
            std::cout << "Generated " << ph.size() << " torch photons on the host (seed " << seed << ")" << std::endl;
        }
        double dt = cx->simulate(0, false);
        std::cout << "Simulation time: " << dt << " seconds" << std::endl;
        std::cout << "Opticks: NumHits:  " << cx->getNumHit() << std::endl;
        phoxapp::write_hits_text(cx, out);
        cx->reset(0);
        delete cx;
    } catch (const std::exception& e) {
        std::cerr << "ERROR: " << e.what() << std::endl;
        return 2;
    }
    return 0;
}
