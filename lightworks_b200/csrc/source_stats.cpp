// source_stats.cpp -- host-only enumeration of the imperfect-source input statistics.
//
// Replaces Source._build_statistics (lightworks/emulator/components/source.py:134-161): _build_statistics_basic
// (:163-224), _build_statistics_full / _full_distribution / _single_mode_distribution / _single_photon_distribution /
// _remap_distribution (:226-391) and apply_threshold (:394-406).  The reference multiplies out 6 labelled options
// per photon as Python dicts of AnnotatedState objects (6^n entries before merging: 131 s for 8 photons) and only
// then merges the ones that are equal up to a relabelling.  After that relabelling an annotated input is fully
// described by two occupation vectors: A_j = photons of mode j carrying the shared label 0 (mutually
// indistinguishable) and B_j = photons of mode j carrying labels of their own; a lone label-0 photon is the same
// state as a distinguishable one (_remap_distribution numbers labels by first appearance).  This file enumerates
// those classes directly: per photon the options (a, b) in the reference's order
//     [] (0,0) c0 | [0] (1,0) c1 | [d] (0,1) c1d | [m] (0,1) c1dp | [0,m] (1,1) c12d | [d,m] (0,2) c1d2d,
// summed per mode, multiplied out across modes (mode 0 outermost, as the reference nests its loops) and merged in
// first-occurrence order, which reproduces the reference's dict order; then p >= threshold and renormalisation.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/lwb200.h"

namespace lwb {
int set_error(int code, const char* msg);  // lwb_api.cu: message for lwb_last_error()
extern int g_source_string_keys;           // test hook: take the string-key path even when the packed form fits
}

namespace {

struct Option {
    int a, b;
    double p;
};
struct ModeClass {
    int a, b;
    double p;
};

// source.py:437-448
bool purity_to_prob(double purity, double* out) {
    if (purity <= 0.5) return false;
    if (purity < 1) {
        const double g2 = 1 - purity;
        const double b = 2 * (1 - (1 / g2));
        *out = 1 - (-b - std::pow(b * b - 4, 0.5)) / 2;
    } else {
        *out = 1;
    }
    return true;
}

// all sums of c per-photon options, first-occurrence order of the photon-major enumeration (source.py:318-351)
std::vector<ModeClass> mode_classes(const std::vector<Option>& opts, int c) {
    std::vector<ModeClass> cur{{0, 0, 1.0}};
    for (int ph = 0; ph < c; ++ph) {
        std::vector<ModeClass> next;
        for (const ModeClass& m : cur)
            for (const Option& o : opts) {
                const int a = m.a + o.a, b = m.b + o.b;
                const double p = m.p * o.p;
                bool found = false;
                for (ModeClass& n : next)
                    if (n.a == a && n.b == b) { n.p += p; found = true; break; }
                if (!found) next.push_back({a, b, p});
            }
        cur.swap(next);
    }
    return cur;
}

}  // namespace

extern "C" int lwb_source_statistics(int n_modes, const uint8_t* state, double purity, double brightness,
                                     double indistinguishability, double threshold, uint64_t cap, uint8_t* group_occ,
                                     uint8_t* single_occ, double* probs, uint64_t* count, int* annotated) {
    if (!state || !count || !annotated || n_modes < 1 || n_modes > 128) return lwb::set_error(LWB_E_ARG, "bad argument");
    auto bad = [](double v) { return !(v >= 0.0 && v <= 1.0); };
    if (bad(purity) || bad(brightness) || bad(indistinguishability) || bad(threshold))
        return lwb::set_error(LWB_E_ARG, "source quantities must lie in [0, 1]");
    std::vector<Option> opts;
    const bool basic = purity == 1 && indistinguishability == 1;  // source.py:155-157
    if (basic) {
        opts.push_back({1, 0, brightness});                             // :178-183
        if (brightness < 1) opts.push_back({0, 0, 1 - brightness});     // :185-186
    } else {
        double p1;
        if (!purity_to_prob(purity, &p1)) return lwb::set_error(LWB_E_ARG, "Purity values below 0.5 not supported.");
        const double nu = brightness;                                   // :365-377
        const double p_i = std::pow(indistinguishability, 0.5), p_d = 1 - p_i, p2 = 1 - p1;
        const double c0 = 1 - nu * (p1 + p2 * nu + 2 * (1 - nu) * p2);
        const double c1 = p_i * nu * (p1 + (1 - nu) * p2);
        const double c1d = p_d * nu * (p1 + (1 - nu) * p2);
        const double c1dp = nu * (1 - nu) * p2;
        const double c12d = nu * nu * p_i * p2;
        const double c1d2d = nu * nu * p_d * p2;
        const Option all[6] = {{0, 0, c0}, {1, 0, c1}, {0, 1, c1d}, {0, 1, c1dp}, {1, 1, c12d}, {0, 2, c1d2d}};
        for (const Option& o : all)
            if (o.p > 0) opts.push_back(o);                             // :391
    }
    *annotated = basic ? 0 : 1;

    std::vector<std::vector<ModeClass>> per_mode(n_modes);
    double combos = 1.0;
    for (int j = 0; j < n_modes; ++j) {
        per_mode[j] = mode_classes(opts, state[j]);
        combos *= (double)per_mode[j].size();
    }
    if (combos > 67108864.0) return lwb::set_error(LWB_E_LIMIT, "more than 2^26 input classes");

    // product over modes (mode 0 outermost, as the reference nests its loops), merged in first-occurrence order.
    // Only occupied modes take part (an empty mode has the single class (0, 0, 1)); a class tuple is packed into one
    // integer, (a_j + (c_j + 1) * b_j) in mixed radix, when that fits 62 bits (strings otherwise), and the running
    // product / key / label-0 count are kept per odometer position so a step recomputes only the changed suffix.
    std::vector<int> active;
    for (int j = 0; j < n_modes; ++j)
        if (state[j] > 0) active.push_back(j);
    const int K = (int)active.size();
    std::vector<uint64_t> stride(K + 1, 1), radix_a(K, 1);
    bool packed = !lwb::g_source_string_keys;
    for (int k = 0; packed && k < K; ++k) {
        const uint64_t c = state[active[k]];
        radix_a[k] = c + 1;
        const uint64_t r = (c + 1) * (2 * c + 2);
        if (stride[k] > (1ull << 62) / r) { packed = false; break; }
        stride[k + 1] = stride[k] * r;
    }
    std::vector<std::string> keys;  // string form of every kept key (decoded from the integer form at the end)
    std::vector<uint64_t> ikeys;
    std::vector<double> vals;
    std::unordered_map<std::string, size_t> sindex;
    std::unordered_map<uint64_t, size_t> iindex;
    if (!(K == 0 && basic)) {
        std::vector<size_t> it(K, 0);
        std::vector<double> pp(K + 1, 1.0);
        std::vector<uint64_t> pk(K + 1, 0);
        std::vector<int> pa(K + 1, 0), plone(K + 1, -1);
        std::string skey(2 * (size_t)n_modes, '\0');
        int from = 0;
        for (;;) {
            for (int k = from; k < K; ++k) {
                const ModeClass& m = per_mode[active[k]][it[k]];
                pp[k + 1] = pp[k] * m.p;
                pa[k + 1] = pa[k] + m.a;
                plone[k + 1] = m.a ? k : plone[k];
                if (packed) pk[k + 1] = pk[k] + ((uint64_t)m.a + radix_a[k] * (uint64_t)m.b) * stride[k];
            }
            const double p = pp[K];
            // a lone label-0 photon is a distinguishable photon after relabelling (:262-283)
            const bool lone = !basic && pa[K] == 1;
            if (packed) {
                uint64_t key = pk[K];
                if (lone) key += (radix_a[plone[K]] - 1) * stride[plone[K]];  // a: 1 -> 0, b: +1
                auto f = iindex.find(key);
                if (f == iindex.end()) {
                    iindex.emplace(key, vals.size());
                    ikeys.push_back(key);
                    vals.push_back(p);
                } else {
                    vals[f->second] += p;
                }
            } else {
                std::fill(skey.begin(), skey.end(), '\0');
                for (int k = 0; k < K; ++k) {
                    const ModeClass& m = per_mode[active[k]][it[k]];
                    skey[active[k]] = (char)m.a;
                    skey[n_modes + active[k]] = (char)m.b;
                }
                if (lone) {
                    const int j = active[plone[K]];
                    skey[j] = 0;
                    skey[n_modes + j] = (char)(skey[n_modes + j] + 1);
                }
                auto f = sindex.find(skey);
                if (f == sindex.end()) {
                    sindex.emplace(skey, vals.size());
                    keys.push_back(skey);
                    vals.push_back(p);
                } else {
                    vals[f->second] += p;
                }
            }
            int k = K - 1;
            while (k >= 0 && ++it[k] == per_mode[active[k]].size()) { it[k] = 0; --k; }
            if (k < 0) break;
            from = k;
        }
        if (packed) {  // decode the integer keys
            keys.resize(ikeys.size());
            for (size_t i = 0; i < ikeys.size(); ++i) {
                std::string sk(2 * (size_t)n_modes, '\0');
                uint64_t key = ikeys[i];
                for (int k = 0; k < K; ++k) {
                    const uint64_t r = radix_a[k] * (2 * (radix_a[k] - 1) + 2);
                    const uint64_t d = key % r;
                    key /= r;
                    sk[active[k]] = (char)(d % radix_a[k]);
                    sk[n_modes + active[k]] = (char)(d / radix_a[k]);
                }
                keys[i] = sk;
            }
        }
    }
    if (keys.empty()) {  // source.py:221-223: an empty distribution returns the input itself
        std::string k(2 * (size_t)n_modes, '\0');
        for (int j = 0; j < n_modes; ++j) k[j] = (char)state[j];
        keys.push_back(k);
        vals.push_back(1.0);
    }
    // apply_threshold, source.py:394-406
    double total = 0.0;
    uint64_t kept = 0;
    for (size_t i = 0; i < keys.size(); ++i)
        if (vals[i] >= threshold) { total += vals[i]; ++kept; }
    *count = kept;
    if (kept > cap) return lwb::set_error(LWB_E_LIMIT, "capacity too small for the source statistics");
    if (kept && (!group_occ || !single_occ || !probs)) return lwb::set_error(LWB_E_ARG, "null output buffer");
    uint64_t w = 0;
    for (size_t i = 0; i < keys.size(); ++i) {
        if (!(vals[i] >= threshold)) continue;
        memcpy(group_occ + w * (uint64_t)n_modes, keys[i].data(), (size_t)n_modes);
        memcpy(single_occ + w * (uint64_t)n_modes, keys[i].data() + n_modes, (size_t)n_modes);
        probs[w] = vals[i] / total;
        ++w;
    }
    return 0;
}
