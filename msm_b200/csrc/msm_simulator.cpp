// msm_simulator.cpp -- native host over the C ABI: the loop of `msm-simulator` (simulator/src/main.rs:21-89) in C++.
//
//   msm-simulator-b200 --params run.params [--out sim-data] [--verbose] [--test] [--max-updates N]
//
// The reference parses a TOML with serde (`msm_common::read_toml`); that crate stays untouched, so this program takes
// the RESOLVED scalars of `SimulationParameters` (simulation_object.rs:67-140) as `key = value` lines -- what
// `SimulationIter::next` (utils/io.rs:164-245) hands to `SimulationObject::new_from_params`.  `python -m msm_b200
// --toml X --export-params run.params` writes such a file from a reference TOML.  Keys:
//   dims size axis_length final_sim_time cfl num_data_dumps total_mass particle_mass hbar_ k2_cutoff alias_threshold
//   sim_name [expanding omega_matter_now omega_radiation_now h z0 max_dloga]
//   ics = ColdGauss m0 m1 m2 s0 s1 s2 | SphericalTophat radius delta slope | File <raw f64 interleaved, linear layout>
//   seeds = 1,2,3 (sampled streams "<sim>-stream%05d", then the un-sampled run "<sim>")   scheme = Wigner | Husimi | Poisson
// Everything that touches the grid goes through include/msm_b200.h; there is no other dependency.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/msm_b200.h"

static std::map<std::string, std::string> read_params(const char* path) {
    std::map<std::string, std::string> kv;
    FILE* f = fopen(path, "r");
    if (!f) {
        fprintf(stderr, "cannot open %s\n", path);
        exit(2);
    }
    char line[4096];
    while (fgets(line, sizeof line, f)) {
        std::string s(line);
        const size_t hash = s.find('#');
        if (hash != std::string::npos) s.erase(hash);
        const size_t eq = s.find('=');
        if (eq == std::string::npos) continue;
        auto trim = [](std::string t) {
            const char* ws = " \t\r\n\"";
            const size_t a = t.find_first_not_of(ws), b = t.find_last_not_of(ws);
            return a == std::string::npos ? std::string() : t.substr(a, b - a + 1);
        };
        kv[trim(s.substr(0, eq))] = trim(s.substr(eq + 1));
    }
    fclose(f);
    return kv;
}

static double num(const std::map<std::string, std::string>& kv, const char* k, double dflt, bool required = false) {
    auto it = kv.find(k);
    if (it == kv.end()) {
        if (required) {
            fprintf(stderr, "missing key %s\n", k);
            exit(2);
        }
        return dflt;
    }
    return atof(it->second.c_str());
}

#define CHECK_SIM(call)                                                                          \
    do {                                                                                         \
        int rc_ = (call);                                                                        \
        if (rc_ != MSM_OK) {                                                                     \
            fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, msm_sim_last_error(sim));        \
            return 1;                                                                            \
        }                                                                                        \
    } while (0)
#define CHECK_CTX(call)                                                                          \
    do {                                                                                         \
        int rc_ = (call);                                                                        \
        if (rc_ != MSM_OK) {                                                                     \
            fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, msm_last_error(ctx));            \
            return 1;                                                                            \
        }                                                                                        \
    } while (0)

int main(int argc, char** argv) {
    const char* params_path = nullptr;
    std::string out = "sim-data";
    bool verbose = false, test_only = false;
    long max_updates = -1;
    for (int i = 1; i < argc; ++i) {
        if (!strcmp(argv[i], "--params") && i + 1 < argc) params_path = argv[++i];
        else if (!strcmp(argv[i], "--out") && i + 1 < argc) out = argv[++i];
        else if (!strcmp(argv[i], "--verbose") || !strcmp(argv[i], "-v")) verbose = true;
        else if (!strcmp(argv[i], "--test")) test_only = true;                       // main.rs:16,59
        else if (!strcmp(argv[i], "--max-updates") && i + 1 < argc) max_updates = atol(argv[++i]);
    }
    if (!params_path) {
        fprintf(stderr, "usage: %s --params run.params [--out dir] [--verbose] [--test]\n", argv[0]);
        return 2;
    }
    const auto kv = read_params(params_path);

    // stream list of SimulationIter (utils/io.rs:164-245): seeds ascending, then the un-sampled mean-field run
    std::vector<long long> seeds;
    if (kv.count("seeds")) {
        std::stringstream ss(kv.at("seeds"));
        std::string tok;
        while (std::getline(ss, tok, ',')) if (!tok.empty()) seeds.push_back(atoll(tok.c_str()));
    }
    const std::string sim_name = kv.count("sim_name") ? kv.at("sim_name") : "sim";
    std::vector<std::string> names;
    for (long long s : seeds) {
        char buf[512];
        snprintf(buf, sizeof buf, "%s-stream%05lld", sim_name.c_str(), s);           // io.rs:199
        names.push_back(buf);
    }
    names.push_back(sim_name);                                                       // io.rs:214-240
    const int S = (int)names.size();

    msm_sim_params p;
    memset(&p, 0, sizeof p);
    p.struct_size = sizeof p;
    p.dims = (int)num(kv, "dims", 3, true);
    p.size = (int)num(kv, "size", 0, true);
    p.n_streams = S;
    p.expanding = (int)num(kv, "expanding", 0);
    p.coupling = MSM_COUPLING_INDEPENDENT;
    p.device = (int)num(kv, "device", 0);
    p.num_data_dumps = (uint32_t)num(kv, "num_data_dumps", 0, true);
    p.nranks = 1;
    p.axis_length = num(kv, "axis_length", 0, true);
    p.time = num(kv, "time", 0.0);
    p.final_sim_time = num(kv, "final_sim_time", 0, true);
    p.cfl = num(kv, "cfl", 0, true);
    p.total_mass = num(kv, "total_mass", 0, true);
    p.particle_mass = num(kv, "particle_mass", 0, true);
    p.hbar_ = num(kv, "hbar_", 0, true);
    p.k2_cutoff = num(kv, "k2_cutoff", 0.95);
    p.alias_threshold = num(kv, "alias_threshold", 0.02);
    if (p.expanding) {
        p.omega_matter_now = num(kv, "omega_matter_now", 0, true);
        p.omega_radiation_now = num(kv, "omega_radiation_now", 0.0);
        p.h = num(kv, "h", 0, true);
        p.z0 = num(kv, "z0", 0, true);
        p.has_max_dloga = kv.count("max_dloga") ? 1 : 0;
        p.max_dloga = num(kv, "max_dloga", 0.0);
    }

    msm_sim* sim = nullptr;
    int rc = msm_sim_create(&p, &sim);
    if (rc != MSM_OK) {
        fprintf(stderr, "msm_sim_create failed (%d): %s\n", rc, msm_sim_last_error(nullptr));
        return 1;
    }
    msm_ctx* ctx = msm_sim_ctx(sim);

    // new_from_params (simulation_object.rs:404-435): IC of the first stream, copies, then the sampler per seed
    std::stringstream ics(kv.count("ics") ? kv.at("ics") : "");
    std::string kind;
    ics >> kind;
    if (kind == "ColdGauss") {
        double mean[3] = {0, 0, 0}, sd[3] = {1, 1, 1};
        for (int a = 0; a < p.dims; ++a) ics >> mean[a];
        for (int a = 0; a < p.dims; ++a) ics >> sd[a];
        CHECK_CTX(msm_ic_cold_gauss(ctx, 0, mean, sd));
    } else if (kind == "SphericalTophat") {
        double radius, delta, slope;
        ics >> radius >> delta >> slope;
        CHECK_CTX(msm_ic_spherical_tophat(ctx, 0, p.axis_length, radius, delta, slope));
    } else if (kind == "File") {
        std::string path;
        ics >> path;
        size_t cells = 1;
        for (int d = 0; d < p.dims; ++d) cells *= (size_t)p.size;
        std::vector<double> buf(2 * cells);
        FILE* f = fopen(path.c_str(), "rb");
        if (!f || fread(buf.data(), sizeof(double), 2 * cells, f) != 2 * cells) {
            fprintf(stderr, "cannot read %zu doubles from %s\n", 2 * cells, path.c_str());
            return 1;
        }
        fclose(f);
        CHECK_SIM(msm_sim_set_psi(sim, 0, buf.data()));
    } else {
        fprintf(stderr, "unknown ics kind '%s'\n", kind.c_str());
        return 2;
    }
    for (int s = 1; s < S; ++s) CHECK_CTX(msm_ic_copy(ctx, s, 0));
    const std::string scheme = kv.count("scheme") ? kv.at("scheme") : "";
    const int scheme_id = scheme == "Wigner" ? MSM_SCHEME_WIGNER : scheme == "Husimi" ? MSM_SCHEME_HUSIMI
                          : scheme == "Poisson" ? MSM_SCHEME_POISSON : MSM_SCHEME_NONE;
    const double n_tot = p.total_mass / p.particle_mass;
    for (size_t i = 0; i < seeds.size(); ++i)
        if (scheme_id != MSM_SCHEME_NONE) CHECK_CTX(msm_sample_perturbation(ctx, (int)i, scheme_id, (uint64_t)seeds[i], n_tot));

    if (verbose) printf("%d streams, %d^%d grid, %s box\n", S, p.size, p.dims, p.expanding ? "expanding" : "static");
    const bool output_potential = num(kv, "output_potential", 0.0) != 0.0;                           // :1167-1180
    auto dump = [&](int s, uint32_t index) -> int {
        int rc_ = msm_sim_dump(sim, s, out.c_str(), names[s].c_str(), index);
        if (rc_ == MSM_OK && output_potential) rc_ = msm_sim_dump_potential(sim, s, out.c_str(), names[s].c_str(), index);
        return rc_;
    };
    if (!test_only) {
        CHECK_SIM(msm_sim_reserve_dump_buffers(sim, 2));                                                 // pinned staging, up front
        for (int s = 0; s < S; ++s) CHECK_SIM(dump(s, 0));                                               // main.rs:61
        long updates = 0;
        while (msm_sim_not_finished(sim) && (max_updates < 0 || updates < max_updates)) {               // main.rs:65
            rc = msm_sim_update(sim);
            if (rc == MSM_E_ALIASING) {                                                                  // :607-617
                fprintf(stderr, "simulation aliased: %s\n", msm_sim_last_error(sim));
                return 3;
            }
            if (rc != MSM_OK) {
                fprintf(stderr, "msm_sim_update failed (%d): %s\n", rc, msm_sim_last_error(sim));
                return 1;
            }
            ++updates;
            for (int s = 0; s < S; ++s) {
                msm_stream_state st;
                msm_sim_state(sim, s, &st);
                if (st.dumped) CHECK_SIM(dump(s, st.current_dumps));                                     // :620-631
            }
        }
        CHECK_SIM(msm_sim_wait_io(sim));                                                                 // RuntimeError::IOError
        if (verbose) {
            unsigned long long steps = 0;
            for (int s = 0; s < S; ++s) {
                msm_stream_state st;
                msm_sim_state(sim, s, &st);
                steps += st.n_steps;
            }
            printf("Finished all streams: %ld updates, %llu stream-steps\n", updates, steps);
        }
    }
    msm_sim_destroy(sim);
    return 0;
}
