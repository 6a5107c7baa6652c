// sim.cpp -- host-logic level of the C ABI (msm_sim_*): a C++ mirror of the reference's `SimulationObject`
// (simulator/src/simulation_object.rs:145-184) batched over streams.  Everything here is host scalars; the grid
// work is delegated to the grid-level calls msm_potential_max / msm_step (core.cu).
//
// Reference rows: a2 SimulationParameters::new (:223-315), a7 get_timestep static
// (:878-934), a8 get_timestep expanding (:939-990), a13 update bookkeeping (:590,:620-635 / :757-759,:828-844),
// a14 dump (utils/io.rs:34-88), a15 not_finished (:1226-1228), a16 ScaleFactorSolver (expanding.rs:12-118),
// rk4 (utils/mod.rs:14-43), calculate_dt_from_dtau (:1344-1388), get_tau (:1408-1453),
// get_supercomoving_boxsize (common/src/parameters.rs:205-220).
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>

#include <errno.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "../../include/msm_b200.h"

namespace {

// common/src/constants.rs:2-9
const double POIS_CONST = 4.0 * M_PI * 4.49e-12;
const double LITTLE_H_TO_BIG_H = 1.022e-4;
const double DEFAULT_MAX_DLOGA = 1e-3;   // expanding.rs:27

struct Cosmo {
    double om = 0, orad = 0, h = 0, z0 = 0, max_dloga = 0;
    bool has_max_dloga = false;
};

// utils/mod.rs:14-43
double rk4(const std::function<double(double, double)>& f, double tn, double yn, double h) {
    const double k1 = f(tn, yn);
    const double k2 = f(tn + h / 2.0, yn + h * k1 / 2.0);
    const double k3 = f(tn + h / 2.0, yn + h * k2 / 2.0);
    const double k4 = f(tn + h, yn + h * k3);
    return yn + h * (k1 + 2.0 * k2 + 2.0 * k3 + k4) / 6.0;
}

// expanding.rs:12-118.  The reference wraps crate `cosmology` 0.2.0 (not vendored).  Restated as flat FLRW,
// omega_de0 = 1 - om - or (expanding.rs:29-38), H0 = h * 1.022e-4 / Myr, a(0) = 1/(1+z0),
// da/dt = a H0 sqrt(om a^-3 + or a^-4 + ode), RK4 sub-steps bounded by max_dloga * a / (da/dt); negative dt
// steps backwards with the same rule.
struct ScaleFactorSolver {
    double om = 0, orad = 0, ode = 0, h0 = 0, max_dloga = DEFAULT_MAX_DLOGA, a = 1, t = 0;
    ScaleFactorSolver() {}
    explicit ScaleFactorSolver(const Cosmo& c) {
        om = c.om;
        orad = c.orad;
        ode = 1.0 - om - orad;
        h0 = c.h * LITTLE_H_TO_BIG_H;
        max_dloga = c.has_max_dloga ? c.max_dloga : DEFAULT_MAX_DLOGA;
        a = 1.0 / (1.0 + c.z0);
        t = 0.0;
    }
    double dadt(double x) const { return x * h0 * sqrt(om / (x * x * x) + orad / (x * x * x * x) + ode); }
    double step(double dt) {
        double remaining = dt;
        while (remaining != 0.0) {
            const double lim = max_dloga * a / dadt(a);
            const double h = fabs(remaining) <= lim ? remaining : copysign(lim, remaining);
            const double k1 = dadt(a);
            const double k2 = dadt(a + 0.5 * h * k1);
            const double k3 = dadt(a + 0.5 * h * k2);
            const double k4 = dadt(a + h * k3);
            a = a + h * (k1 + 2.0 * k2 + 2.0 * k3 + k4) / 6.0;
            t += h;
            remaining = (h == remaining) ? 0.0 : remaining - h;
        }
        return a;
    }
    double get_a() const { return a; }
    double get_dadt() const { return dadt(a); }
    double get_time() const { return t; }
};

// simulation_object.rs:1408-1453
double get_tau(double target_time, const Cosmo& c) {
    ScaleFactorSolver solver(c);
    const double pref = sqrt(1.5 * c.om * pow(LITTLE_H_TO_BIG_H * c.h, 2));
    auto dtau_dt = [&](double t, double) {
        const double a_at_t = solver.step(t - solver.get_time());
        return pref / (a_at_t * a_at_t);
    };
    double tau = 0.0, time = 0.0;
    while (time < target_time) {
        double dt = target_time / 1000.0;
        if (c.has_max_dloga) dt = fmin(target_time / 1000.0, solver.get_a() / solver.get_dadt() * c.max_dloga);
        dt = fmin(dt, target_time - time);
        tau = rk4(dtau_dt, time, tau, dt);
        time += dt;
    }
    return tau;
}

// common/src/parameters.rs:205-220
double supercomoving_boxsize(double hbar_, const Cosmo& c, double axis_length) {
    const double a0 = 1.0 / (1.0 + c.z0);
    const double comoving = axis_length / a0;
    return sqrt(sqrt(1.5 * c.om * pow(LITTLE_H_TO_BIG_H * c.h, 2)) / hbar_) * comoving;
}

struct Stream {
    double time = 0, tau = 0, dt = 0, potential_max = 0, alias_mass = 0;
    uint64_t n_steps = 0;
    uint32_t current_dumps = 0;
    int dumped = 0, aliased = 0;
    ScaleFactorSolver solver;
};

}  // namespace

// dump staging (row f-2): a small pool of pinned host buffers, 2 * cells doubles each (re plane | im plane).  The D2H
// copy of a dump runs on the library's copy stream while the step kernels continue; writer threads wait for their
// copy (msm_download_wait), write the NPY files and hand the buffer back.
struct DumpPool {
    std::mutex mu;
    std::condition_variable cv;
    std::vector<double*> all, free_;
    bool failed = false;          // a writer failed: reported by the next msm_sim_dump* / msm_sim_wait_io as MSM_E_IO
    std::string error;
};

struct msm_sim {
    msm_sim_params p{};
    msm_derived d{};
    Cosmo cosmo;
    msm_ctx* ctx = nullptr;
    std::vector<Stream> st;
    std::vector<std::thread> io;
    DumpPool pool;
    std::string err;
};

namespace {
std::string g_sim_error;
int sfail(msm_sim* s, int code, const std::string& m) {
    if (s) s->err = m; else g_sim_error = m;
    return code;
}

bool not_finished(const msm_sim* s, const Stream& st) { return st.time < s->p.final_sim_time; }   // :1226-1228

// simulation_object.rs:878-934 / :939-990
void get_timestep(const msm_sim* sim, const Stream& st, double potential_max, bool* dump, double* dt_out) {
    const msm_sim_params& p = sim->p;
    const double time_to_next_dump =
        ((double)(st.current_dumps + 1) * p.final_sim_time / (double)p.num_data_dumps) - st.time;   // :916-919
    if (!p.expanding) {
        const double kinetic_dt = p.cfl * 2.0 * p.axis_length / sqrt(sim->d.k2_max) / p.hbar_;        // :881-884
        const double potential_dt = p.cfl * (2.0 * M_PI) * p.hbar_ / (2.0 * potential_max);           // :906-909
        const double dt = fmin(fmin(kinetic_dt, potential_dt), time_to_next_dump);                    // :922
        *dump = (dt == time_to_next_dump);                                                            // :927
        *dt_out = dt;
        return;
    }
    const double kinetic_dtau = p.cfl * 2.0 * sim->d.comoving_boxsize / sqrt(sim->d.k2_max);          // :942-944
    const double potential_dtau = p.cfl * (2.0 * M_PI) / ((2.0 * st.solver.get_a()) * potential_max); // :957-959
    const double tau_to_next_dump = get_tau(st.time + time_to_next_dump, sim->cosmo) - st.tau;        // :970-975
    const double dtau = fmin(fmin(kinetic_dtau, potential_dtau), tau_to_next_dump);                   // :978
    *dump = (dtau == tau_to_next_dump);                                                               // :983
    *dt_out = dtau;
}

// simulation_object.rs:1344-1388
double calculate_dt_from_dtau(const msm_sim* sim, const Stream& st, double dtau) {
    ScaleFactorSolver solver = st.solver;   // clone
    const double pref = sqrt(1.5 * sim->cosmo.om * pow(LITTLE_H_TO_BIG_H * sim->cosmo.h, 2));
    auto dt_dtau = [&](double, double t) {
        const double a_at_t = solver.step(t - solver.get_time());
        return 1.0 / (pref / (a_at_t * a_at_t));
    };
    return rk4(dt_dtau, st.tau, st.time, dtau) - st.time;
}

// NPY v1.0 writer for one f64 plane, shape (n, n|1, n|1, 1)   (utils/io.rs:63-66,90-97)
bool write_npy(const std::string& path, const double* data, int dims, int n) {
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) return false;
    char dict[160];
    const int s1 = dims >= 2 ? n : 1, s2 = dims >= 3 ? n : 1;
    int len = snprintf(dict, sizeof dict, "{'descr': '<f8', 'fortran_order': False, 'shape': (%d, %d, %d, 1), }", n, s1, s2);
    const int unpadded = 10 + len + 1;
    const int pad = (64 - unpadded % 64) % 64;
    const uint16_t hlen = (uint16_t)(len + pad + 1);
    const unsigned char magic[8] = {0x93, 'N', 'U', 'M', 'P', 'Y', 1, 0};
    bool ok = fwrite(magic, 1, 8, f) == 8;
    ok = ok && fwrite(&hlen, 2, 1, f) == 1;
    ok = ok && fwrite(dict, 1, len, f) == (size_t)len;
    for (int i = 0; i < pad && ok; ++i) ok = fputc(' ', f) != EOF;
    ok = ok && fputc('\n', f) != EOF;
    size_t count = 1;
    for (int d = 0; d < dims; ++d) count *= (size_t)n;
    if (data) {
        ok = ok && fwrite(data, sizeof(double), count, f) == count;
    } else {   // an all-zero plane (the imaginary file of a potential dump, simulation_object.rs:1167-1180)
        std::vector<double> z(std::min<size_t>(count, 1 << 16), 0.0);
        for (size_t done = 0; done < count && ok; done += z.size()) {
            const size_t m = std::min(z.size(), count - done);
            ok = fwrite(z.data(), sizeof(double), m, f) == m;
        }
    }
    ok = (fclose(f) == 0) && ok;
    return ok;
}

// Stream groups of msm_sim_run_streams: the launch chunk when there are enough streams to pipeline, else smaller
// (even) groups.  The first and the last group are short (2 streams) so that the un-overlapped upload of the first
// group and download of the last one are short: [2, chunk-2, chunk, ..., chunk, chunk-2, 2].
std::vector<int> run_groups(int n, int chunk) {
    std::vector<int> bounds{0};
    if (n <= 0) return bounds;
    chunk = std::max(1, chunk);
    if (n >= 4 * chunk && chunk >= 4) {
        const int lead[2] = {2, chunk - 2};
        for (int k = 0; k < 2; ++k) bounds.push_back(bounds.back() + lead[k]);
        const int tail_begin = n - chunk;
        while (bounds.back() + chunk <= tail_begin) bounds.push_back(bounds.back() + chunk);
        if (bounds.back() < tail_begin) bounds.push_back(tail_begin);
        bounds.push_back(n - 2);
        bounds.push_back(n);
    } else {
        int g = chunk;
        if (n < 4 * g) g = std::max(2, (n / 4) & ~1);
        g = std::max(1, std::min(g, chunk));
        while (bounds.back() < n) bounds.push_back(std::min(n, bounds.back() + g));
    }
    return bounds;
}

bool mkdirs(const std::string& path) {
    std::string cur;
    for (size_t i = 0; i < path.size(); ++i) {
        cur.push_back(path[i]);
        if (path[i] == '/' || i + 1 == path.size())
            if (mkdir(cur.c_str(), 0777) != 0 && errno != EEXIST && !(cur == "/")) return false;
    }
    struct stat sb;
    return stat(path.c_str(), &sb) == 0 && S_ISDIR(sb.st_mode);
}

size_t sim_cells(const msm_sim* sim) {
    size_t c = 1;
    for (int d = 0; d < sim->p.dims; ++d) c *= (size_t)sim->p.size;
    return c;
}

int pool_error(msm_sim* sim) {
    std::lock_guard<std::mutex> lock(sim->pool.mu);
    if (!sim->pool.failed) return MSM_OK;
    sim->err = sim->pool.error;
    return MSM_E_IO;
}

size_t pool_cap() {
    size_t cap = 4;
    if (const char* e = getenv("MSM_B200_DUMP_BUFFERS")) cap = (size_t)std::max(1, atoi(e));
    return cap;
}

// A free staging buffer; grows the pool up to MSM_B200_DUMP_BUFFERS (default 4), then waits for a writer to finish.
// Pinning host memory is expensive (cudaMallocHost: 0.1-0.2 s per 256 MiB measured), so once the pool holds two buffers
// a momentary shortage first waits 20 ms for a writer before it pays for another buffer; msm_sim_reserve_dump_buffers
// moves the whole cost in front of the step loop.
int pool_acquire(msm_sim* sim, double** out) {
    DumpPool& P = sim->pool;
    const size_t cap = pool_cap();
    std::unique_lock<std::mutex> lock(P.mu);
    bool waited = false;
    for (;;) {
        if (!P.free_.empty()) {
            *out = P.free_.back();
            P.free_.pop_back();
            return MSM_OK;
        }
        if (P.all.size() >= 2 && P.all.size() < cap && !waited) {
            waited = true;
            P.cv.wait_for(lock, std::chrono::milliseconds(20));
            continue;
        }
        if (P.all.size() < cap) {
            void* p = nullptr;
            lock.unlock();
            const int rc = msm_host_alloc(sim->ctx, 2 * sizeof(double) * sim_cells(sim), &p);
            lock.lock();
            if (rc == MSM_OK) {
                P.all.push_back((double*)p);
                *out = (double*)p;
                return MSM_OK;
            }
            if (P.all.empty()) {
                sim->err = msm_last_error(sim->ctx);
                return rc;
            }
            // no more pinned memory: live with the buffers we have
        }
        P.cv.wait(lock);
    }
}

void pool_release(msm_sim* sim, double* buf) {
    {
        std::lock_guard<std::mutex> lock(sim->pool.mu);
        sim->pool.free_.push_back(buf);
    }
    sim->pool.cv.notify_all();
}

void pool_fail(msm_sim* sim, const std::string& what) {
    std::lock_guard<std::mutex> lock(sim->pool.mu);
    if (!sim->pool.failed) sim->pool.error = what;
    sim->pool.failed = true;
}

// two writer threads per dump, one per plane (utils/io.rs:58-60,72-74); the last one returns the staging buffer
void spawn_writers(msm_sim* sim, double* buf, bool has_ticket, uint64_t ticket, const std::string& path_re,
                   const std::string& path_im, bool imag_zero) {
    const int dims = sim->p.dims, n = sim->p.size;
    const size_t cells = sim_cells(sim);
    auto left = std::make_shared<std::atomic<int>>(2);
    auto job = [=](const std::string& path, const double* data) {
        bool ok = !has_ticket || msm_download_wait(sim->ctx, ticket) == MSM_OK;
        if (!ok) pool_fail(sim, "dump: the device-to-host copy failed for " + path);
        else if (!write_npy(path, data, dims, n)) pool_fail(sim, "dump: cannot write " + path + ": " + strerror(errno));
        if (left->fetch_sub(1) == 1) pool_release(sim, buf);
    };
    const double* im_plane = imag_zero ? (const double*)nullptr : (const double*)(buf + cells);
    // a thread that cannot be started (std::system_error) must not unwind through the C ABI: its plane is written here
    try {
        sim->io.emplace_back(job, path_re, (const double*)buf);
    } catch (const std::exception&) {
        job(path_re, (const double*)buf);
    }
    try {
        sim->io.emplace_back(job, path_im, im_plane);
    } catch (const std::exception&) {
        job(path_im, im_plane);
    }
}

}  // namespace

extern "C" {

const char* msm_sim_last_error(const msm_sim* sim) { return sim ? sim->err.c_str() : g_sim_error.c_str(); }

double msm_get_tau(double target_time, double om, double orad, double h, double z0, double max_dloga, int32_t has) {
    Cosmo c;
    c.om = om; c.orad = orad; c.h = h; c.z0 = z0; c.max_dloga = max_dloga; c.has_max_dloga = has != 0;
    return get_tau(target_time, c);
}

double msm_supercomoving_boxsize(double hbar_, double om, double h, double z0, double axis_length) {
    Cosmo c;
    c.om = om; c.h = h; c.z0 = z0;
    return supercomoving_boxsize(hbar_, c, axis_length);
}

double msm_scale_factor_after(double t, double om, double orad, double h, double z0, double max_dloga) {
    Cosmo c;
    c.om = om; c.orad = orad; c.h = h; c.z0 = z0; c.max_dloga = max_dloga; c.has_max_dloga = true;
    ScaleFactorSolver s(c);
    return s.step(t);
}

int msm_sim_create(const msm_sim_params* p, msm_sim** out) {
    if (!p || !out) return sfail(nullptr, MSM_E_ARG, "null argument");
    *out = nullptr;
    if (p->struct_size != (int32_t)sizeof(msm_sim_params)) return sfail(nullptr, MSM_E_ARG, "msm_sim_params.struct_size mismatch");
    if (p->num_data_dumps == 0) return sfail(nullptr, MSM_E_ARG, "num_data_dumps must be > 0");
    if (p->size < 2 || (p->size % 2)) return sfail(nullptr, MSM_E_ARG, "size must be even (utils/fft.rs:105)");
    msm_sim* s = new msm_sim();
    s->p = *p;
    s->cosmo.om = p->omega_matter_now;
    s->cosmo.orad = p->omega_radiation_now;
    s->cosmo.h = p->h;
    s->cosmo.z0 = p->z0;
    s->cosmo.max_dloga = p->max_dloga;
    s->cosmo.has_max_dloga = p->has_max_dloga != 0;

    // SimulationParameters::new (simulation_object.rs:243-274)
    msm_derived& d = s->d;
    d.tau0 = 0.0;
    d.final_sim_tau = 0.0;
    d.comoving_boxsize = 0.0;
    if (p->expanding) {
        if (p->omega_matter_now + p->omega_radiation_now > 1.0 || p->z0 < 0.0 || p->omega_matter_now < 0.0 ||
            p->omega_radiation_now < 0.0) {   // expanding.rs:62-80
            delete s;
            return sfail(nullptr, MSM_E_ARG, "only flat cosmologies with z0 >= 0 are supported");
        }
        d.tau0 = get_tau(p->time, s->cosmo);                                     // :246
        d.final_sim_tau = get_tau(p->final_sim_time, s->cosmo);                  // :248-249
        d.comoving_boxsize = supercomoving_boxsize(p->hbar_, s->cosmo, p->axis_length);   // :251-256
        d.dx = d.comoving_boxsize / (double)p->size;                             // :262
    } else {
        d.dx = p->axis_length / (double)p->size;                                 // :260
    }
    d.dk = d.dx;                                                                 // :263
    d.n_tot = p->total_mass / p->particle_mass;                                  // :264
    {   // k2_max = max(spec_grid): every axis at its Nyquist index, summed in spec_grid's order (fft.rs:141-160)
        const double kn = (double)(-(p->size / 2)) / ((double)p->size * d.dx);
        const double m = kn * kn;
        double sum = m;
        if (p->dims >= 2) sum = sum + m;
        if (p->dims >= 3) sum = sum + m;
        d.k2_max = sum * ((2.0 * M_PI) * (2.0 * M_PI));
    }
    if (p->expanding)   // calculate_density :1033-1048
        d.density_prefactor = p->total_mass * POIS_CONST *
                              pow(2.0 / (3.0 * pow(p->h * LITTLE_H_TO_BIG_H, 2) * p->omega_matter_now), 1.0 / 4.0) /
                              pow(p->hbar_, (double)p->dims / 2.0);
    else
        d.density_prefactor = p->total_mass;                                     // :1056
    d.poisson_coeff = p->expanding ? -1.0 : -POIS_CONST;                         // :1079-1086

    msm_config cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.struct_size = sizeof(msm_config);
    cfg.dims = p->dims;
    cfg.size = p->size;
    cfg.n_streams = p->n_streams;
    cfg.coupling = p->coupling;
    cfg.device = p->device;
    cfg.chunk_streams = p->chunk_streams;
    cfg.rank = p->rank;
    cfg.nranks = p->nranks > 0 ? p->nranks : 1;
    cfg.n_streams_global = p->n_streams_global;
    cfg.dx = d.dx;
    cfg.density_prefactor = d.density_prefactor;
    cfg.poisson_coeff = d.poisson_coeff;
    cfg.k2_cutoff = p->k2_cutoff;
    cfg.nccl_unique_id = p->nccl_unique_id;
    int rc = msm_create(&cfg, &s->ctx);
    if (rc) {
        g_sim_error = msm_last_error(nullptr);
        delete s;
        return rc;
    }
    s->st.resize(p->n_streams);
    for (auto& st : s->st) {
        st.time = p->time;
        st.tau = d.tau0;
        if (p->expanding) st.solver = ScaleFactorSolver(s->cosmo);               // new_from_params :437-438
    }
    *out = s;
    return MSM_OK;
}

int msm_sim_wait_io(msm_sim* sim) {
    if (!sim) return MSM_E_ARG;
    for (auto& t : sim->io)
        if (t.joinable()) t.join();
    sim->io.clear();
    const int rc = pool_error(sim);   // RuntimeError::IOError (utils/error.rs:5-27); the reference panics in the writer
    if (rc) {
        std::lock_guard<std::mutex> lock(sim->pool.mu);
        sim->pool.failed = false;     // reported once
    }
    return rc;
}

void msm_sim_destroy(msm_sim* sim) {
    if (!sim) return;
    msm_sim_wait_io(sim);
    for (double* b : sim->pool.all) msm_host_free(sim->ctx, b);
    msm_destroy(sim->ctx);
    delete sim;
}

msm_ctx* msm_sim_ctx(msm_sim* sim) { return sim ? sim->ctx : nullptr; }

int msm_sim_derived(const msm_sim* sim, msm_derived* out) {
    if (!sim || !out) return MSM_E_ARG;
    *out = sim->d;
    return MSM_OK;
}

int msm_sim_set_psi(msm_sim* sim, int32_t stream, const double* psi) {
    if (!sim) return MSM_E_ARG;
    int rc = msm_set_psi(sim->ctx, stream, psi);
    if (rc) sim->err = msm_last_error(sim->ctx);
    else sim->st[stream].aliased = 0;   // a new wavefunction
    return rc;
}

int msm_sim_not_finished(const msm_sim* sim) {
    if (!sim) return 0;
    // a stream that crossed alias_threshold has stopped for good (the reference panics there, :607-617)
    for (const auto& st : sim->st)
        if (not_finished(sim, st) && !st.aliased) return 1;
    return 0;
}

// one `update()` (:475 / :669) for every unfinished stream of `subset` (NULL = all streams)
static int update_streams(msm_sim* sim, const int32_t* subset) {
    const msm_sim_params& p = sim->p;
    const int S = p.n_streams;
    const bool summed = p.coupling == MSM_COUPLING_SUMMED;
    std::vector<int32_t> active(S, 0);
    int nact = 0, head = -1;
    for (int s = 0; s < S; ++s) {
        if (subset && !subset[s]) continue;
        sim->st[s].dumped = 0;
        // a stream that crossed alias_threshold stops there: the reference panics (:607-617)
        if (not_finished(sim, sim->st[s]) && !sim->st[s].aliased) {
            active[s] = 1;
            if (head < 0) head = s;
            ++nact;
        }
    }
    if (nact == 0) return MSM_OK;

    // calculate_potential at t + max_all(abs(phi))  (:497 -> :905)
    std::vector<double> pmax(S, 0.0), drift(S, 0.0), kick(S, 0.0), alias(S, 0.0), dts(S, 0.0);
    std::vector<char> dump(S, 0);
    int rc = msm_potential_max(sim->ctx, active.data(), pmax.data());
    if (rc) {
        sim->err = msm_last_error(sim->ctx);
        return rc;
    }
    // The host scalars of the step are staged in copies and committed only after msm_step succeeded, so that a failed
    // step leaves time / tau / a(t) consistent with psi on the device.
    std::vector<Stream> next(sim->st);
    for (int s = 0; s < S; ++s) {
        if (!active[s]) continue;
        Stream& st = next[s];
        bool dmp = false;
        double dt = 0.0;
        if (summed && s != head) {   // shared potential => identical scalars; keep the streams in lock step
            dmp = dump[head];
            dt = dts[head];
        } else {
            get_timestep(sim, st, pmax[s], &dmp, &dt);                           // :500 / :695
        }
        dump[s] = dmp;
        dts[s] = dt;
        st.dt = dt;
        st.potential_max = pmax[s];
        if (!p.expanding) {
            drift[s] = dt / 4.0 * p.hbar_;                                       // :508
            kick[s] = dt / p.hbar_;                                              // :537
        } else {
            drift[s] = dt / 4.0;                                                 // :701
            double ksum = 0.0;
            for (int j = 0; j < 2; ++j) {                                        // :726-760
                const double a = st.solver.get_a();                              // :728
                ksum += dt / 2.0 * a;                                            // :733 (both half kicks use the same phi)
                const double dt_half = calculate_dt_from_dtau(sim, st, dt / 2.0);   // :751-752
                st.solver.step(dt_half);                                         // :755-756
                st.time = st.time + dt_half;                                     // :757
                st.tau = st.tau + dt / 2.0;                                      // :759
            }
            kick[s] = ksum;
        }
    }
    rc = msm_step(sim->ctx, active.data(), drift.data(), kick.data(), alias.data());
    if (rc) {
        sim->err = msm_last_error(sim->ctx);
        return rc;
    }
    // Summed density: the streams of ALL ranks share one potential, so an alias event anywhere stops every stream at this
    // step (the reference would have panicked, :607-617); the ranks agree on the largest alias mass first.
    double alias_worst = 0.0;
    if (summed) {
        for (int s = 0; s < S; ++s)
            if (active[s]) alias_worst = fmax(alias_worst, alias[s]);
        rc = msm_allreduce_max(sim->ctx, &alias_worst);
        if (rc) {
            sim->err = msm_last_error(sim->ctx);
            return rc;
        }
    }
    int result = MSM_OK;
    for (int s = 0; s < S; ++s) {
        if (!active[s]) continue;
        Stream& st = sim->st[s];
        st = next[s];
        if (!p.expanding) st.time = st.time + dts[s];                            // :590
        st.alias_mass = alias[s];
        st.aliased = (alias[s] > p.alias_threshold || (summed && alias_worst > p.alias_threshold)) ? 1 : 0;   // :1288
        if (st.aliased) {
            result = MSM_E_ALIASING;                                             // the reference panics here (:607-617)
            char buf[200];
            snprintf(buf, sizeof buf, "simulation aliased: stream %d threshold %g k2_cutoff %g p_mass %g%s", s,
                     p.alias_threshold, p.k2_cutoff, alias[s],
                     alias[s] > p.alias_threshold ? "" : " (another stream of the summed ensemble crossed the threshold)");
            sim->err = buf;
        }
        if (dump[s]) {                                                           // :620-631 / :828-844
            st.current_dumps += 1;
            st.dumped = 1;
            st.time = (double)st.current_dumps * p.final_sim_time / (double)p.num_data_dumps;
            if (p.expanding) st.tau = get_tau(st.time, sim->cosmo);
        }
        st.n_steps += 1;                                                         // :635 / :797
    }
    return result;
}

int msm_sim_update(msm_sim* sim) {
    if (!sim) return MSM_E_ARG;
    return update_streams(sim, nullptr);
}

int msm_run_groups(int32_t n, int32_t chunk, int32_t* bounds, int32_t cap) {
    const std::vector<int> b = run_groups(n, chunk);
    if (bounds)
        for (size_t i = 0; i < b.size() && (int32_t)i < cap; ++i) bounds[i] = b[i];
    return (int)b.size() - 1;
}

int msm_sim_update_streams(msm_sim* sim, const int32_t* subset) {
    if (!sim) return MSM_E_ARG;
    if (subset && sim->p.coupling == MSM_COUPLING_SUMMED)
        return sfail(sim, MSM_E_ARG, "msm_sim_update_streams: summed coupling advances all streams together");
    return update_streams(sim, subset);
}

// The reference's outer loop (simulator/src/main.rs:43-85): for every stream of the TOML, build the IC, construct the
// SimulationObject (new_from_params, simulation_object.rs:404-449), `while not_finished() { update() }` (main.rs:65-69)
// and write the final state.  The reference runs the streams strictly one after another; here they advance in groups
// of up to chunk_streams, and while group c computes, the ICs of group c+1 cross PCIe on the upload stream and the final
// wavefunctions of group c-1 leave on the copy stream (one download enqueued after every update, so that the two
// staging buffers never stall the step kernels).
// `prepare(i)` puts the initial wavefunction of streams[i] in place (returns 0 / an error code; < 0 on error, 1 = the
// stream keeps its current state): an asynchronous upload from host memory, or a device-side build from a seed.
static int run_streams_impl(msm_sim* sim, int32_t n, const int32_t* streams, const std::function<int(int)>& prepare,
                            double* const* re_out, double* const* im_out, uint64_t max_updates);

int msm_sim_run_streams(msm_sim* sim, int32_t n, const int32_t* streams, const double* const* psi_in, double* const* re_out,
                        double* const* im_out, uint64_t max_updates) {
    if (!sim || n < 0 || (n > 0 && !streams)) return sfail(sim, MSM_E_ARG, "msm_sim_run_streams: bad argument");
    auto prepare = [&](int i) -> int {
        if (!psi_in || !psi_in[i]) return 1;
        return msm_upload_begin(sim->ctx, streams[i], psi_in[i]);
    };
    return run_streams_impl(sim, n, streams, prepare, re_out, im_out, max_updates);
}

int msm_sim_run_streams_seeded(msm_sim* sim, int32_t n, const int32_t* streams, int32_t scheme, const uint64_t* seeds,
                               double* const* re_out, double* const* im_out, uint64_t max_updates) {
    if (!sim || n < 0 || (n > 0 && (!streams || !seeds))) return sfail(sim, MSM_E_ARG, "msm_sim_run_streams_seeded: bad argument");
    // new_from_params (simulation_object.rs:404-435): the un-sampled initial condition (saved once by msm_ic_store), then
    // sample_quantum_perturbation with the stream's seed (ics.rs:434-648) -- both on the device, enqueued behind the
    // kernels of the previous group; only scalars cross PCIe on the way in.
    auto prepare = [&](int i) -> int {
        int rc = msm_ic_load(sim->ctx, streams[i]);
        if (rc == MSM_OK && seeds[i] != MSM_SEED_NONE)
            rc = msm_sample_perturbation(sim->ctx, streams[i], scheme, seeds[i], sim->d.n_tot);
        return rc;
    };
    return run_streams_impl(sim, n, streams, prepare, re_out, im_out, max_updates);
}

static int run_streams_impl(msm_sim* sim, int32_t n, const int32_t* streams, const std::function<int(int)>& prepare,
                            double* const* re_out, double* const* im_out, uint64_t max_updates) {
    const msm_sim_params& p = sim->p;
    const int S = p.n_streams;
    if (p.coupling == MSM_COUPLING_SUMMED)
        return sfail(sim, MSM_E_ARG, "msm_sim_run_streams: streams are coupled (summed density); use msm_sim_update");
    std::vector<char> seen(S, 0);
    for (int i = 0; i < n; ++i) {
        if (streams[i] < 0 || streams[i] >= S || seen[streams[i]])
            return sfail(sim, MSM_E_ARG, "msm_sim_run_streams: stream out of range or listed twice");
        seen[streams[i]] = 1;
    }
    if (n == 0) return MSM_OK;
    int32_t chunk = 1;
    msm_chunk_streams(sim->ctx, &chunk);
    const std::vector<int> bounds = run_groups(n, chunk);
    const int ngroups = (int)bounds.size() - 1;
    auto has_out = [&](int i) { return (re_out && re_out[i]) || (im_out && im_out[i]); };
    auto gfail = [&](int rc) {
        sim->err = msm_last_error(sim->ctx);
        msm_transfers_wait(sim->ctx);
        return rc;
    };
    auto upload_group = [&](int c) -> int {
        for (int i = bounds[c]; i < bounds[c + 1]; ++i) {
            const int rc = prepare(i);
            if (rc < 0) return rc;
            if (rc == 1) continue;                                               // keeps its current state
            Stream fresh;                                                        // a new SimulationObject (:404-449)
            fresh.time = p.time;
            fresh.tau = sim->d.tau0;
            if (p.expanding) fresh.solver = ScaleFactorSolver(sim->cosmo);
            sim->st[streams[i]] = fresh;
        }
        return MSM_OK;
    };
    std::vector<int> pending;   // indices whose final psi still has to leave the device
    size_t pending_pos = 0;
    auto download_some = [&](size_t count) -> int {
        for (; count > 0 && pending_pos < pending.size(); --count, ++pending_pos) {
            const int i = pending[pending_pos];
            if (int rc = msm_download_begin(sim->ctx, streams[i], re_out ? re_out[i] : nullptr, im_out ? im_out[i] : nullptr))
                return rc;
        }
        return MSM_OK;
    };
    // MSM_B200_TRACE=1: wall-clock milestones of the pipeline on stderr (diagnostics only)
    const bool trace = getenv("MSM_B200_TRACE") != nullptr;
    const auto t_begin = std::chrono::steady_clock::now();
    auto stamp = [&](const char* what, int c) {
        if (!trace) return;
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
        fprintf(stderr, "[msm_sim_run_streams] %9.2f ms  %s %d\n", ms, what, c);
    };
    int result = MSM_OK;
    std::string alias_text;
    int rc = upload_group(0);
    stamp("uploads enqueued for group", 0);
    if (rc) return gfail(rc);
    for (int c = 0; c < ngroups; ++c) {
        if (c + 1 < ngroups && (rc = upload_group(c + 1))) return gfail(rc);
        stamp("start of group", c);
        std::vector<int32_t> mask(S, 0);
        const int lo = bounds[c], hi = bounds[c + 1];
        // downloads of the previous group per update of this one (all of them when this group has nothing to do)
        const size_t left = pending.size() - pending_pos;
        const size_t per_update = max_updates ? (left + max_updates - 1) / max_updates : 1;
        for (uint64_t u = 0; max_updates == 0 || u < max_updates; ++u) {
            int nact = 0;
            for (int i = lo; i < hi; ++i) {
                const Stream& st = sim->st[streams[i]];
                mask[streams[i]] = (not_finished(sim, st) && !st.aliased) ? 1 : 0;   // the reference aborts on aliasing (:607-617)
                nact += mask[streams[i]];
            }
            if (!nact) break;
            rc = update_streams(sim, mask.data());
            if (rc == MSM_E_ALIASING) {
                result = rc;
                alias_text = sim->err;
            } else if (rc) {
                msm_transfers_wait(sim->ctx);
                return rc;
            }
            if (u == 0) stamp("first update done of group", c);
            if ((rc = download_some(per_update))) return gfail(rc);
        }
        stamp("updates done of group", c);
        if ((rc = download_some(pending.size()))) return gfail(rc);
        pending.clear();
        pending_pos = 0;
        for (int i = lo; i < hi; ++i)
            if (has_out(i)) pending.push_back(i);
    }
    if ((rc = download_some(pending.size()))) return gfail(rc);
    stamp("last downloads enqueued", ngroups);
    if ((rc = msm_transfers_wait(sim->ctx))) return gfail(rc);
    stamp("transfers done", ngroups);
    if (result == MSM_E_ALIASING) sim->err = alias_text;
    return result;
}

int msm_sim_state(const msm_sim* sim, int32_t s, msm_stream_state* out) {
    if (!sim || !out || s < 0 || s >= sim->p.n_streams) return MSM_E_ARG;
    const Stream& st = sim->st[s];
    out->time = st.time;
    out->tau = st.tau;
    out->dt = st.dt;
    out->potential_max = st.potential_max;
    out->alias_mass = st.alias_mass;
    out->scale_factor = sim->p.expanding ? st.solver.get_a() : 1.0;
    out->n_steps = st.n_steps;
    out->current_dumps = st.current_dumps;
    out->dumped = st.dumped;
    out->finished = not_finished(sim, st) ? 0 : 1;
    out->aliased = st.aliased;
    return MSM_OK;
}

int msm_sim_get_psi(msm_sim* sim, int32_t stream, double* re, double* im) {
    if (!sim) return MSM_E_ARG;
    int rc = msm_get_psi(sim->ctx, stream, re, im);
    if (rc) sim->err = msm_last_error(sim->ctx);
    return rc;
}

// `dump()` (simulation_object.rs:1113-1223) without stalling the step loop: the call only ENQUEUES the inverse transform
// + plane split on the compute stream and the D2H copy on the copy stream (msm_download_begin) into a pinned staging
// buffer; two writer threads wait for that copy and write the NPY files.  The reference's `array.host()` blocks
// (utils/io.rs:46-47).  A failed writer surfaces as MSM_E_IO from the next msm_sim_dump* or msm_sim_wait_io.
static int dump_field(msm_sim* sim, int32_t stream, const char* root_dir, const char* sim_name, uint32_t dump_index,
                      bool potential) {
    if (!sim || !root_dir || !sim_name || stream < 0 || stream >= sim->p.n_streams)
        return sfail(sim, MSM_E_ARG, "msm_sim_dump: bad argument");
    if (int rc = pool_error(sim)) return rc;
    const bool trace = getenv("MSM_B200_TRACE") != nullptr;
    const auto t_begin = std::chrono::steady_clock::now();
    auto stamp = [&](const char* what) {
        if (!trace) return;
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
        fprintf(stderr, "[msm_sim_dump] %8.2f ms  %s\n", ms, what);
    };
    // at most 2 * MAX_CONCURRENT_GRID_WRITES live writers (simulation_object.rs:39,:1123)
    if (sim->io.size() >= 32)
        if (int rc = msm_sim_wait_io(sim)) return rc;
    const std::string dir = std::string(root_dir) + "/" + sim_name;
    if (!mkdirs(dir)) return sfail(sim, MSM_E_IO, "dump: cannot create directory " + dir + ": " + strerror(errno));   // :1119
    stamp("directory ready");
    double* buf = nullptr;
    if (int rc = pool_acquire(sim, &buf)) return rc;
    stamp("staging buffer acquired");
    const size_t cells = sim_cells(sim);
    uint64_t ticket = 0;
    int rc;
    if (potential) {   // calculate_potential + dump of phi (:1167-1180); rare, computed synchronously
        rc = msm_get_potential(sim->ctx, stream, buf);
    } else {
        rc = msm_download_begin(sim->ctx, stream, buf, buf + cells);
        if (!rc) rc = msm_download_ticket(sim->ctx, &ticket);
    }
    if (rc) {
        sim->err = msm_last_error(sim->ctx);
        pool_release(sim, buf);
        return rc;
    }
    char base[64];
    snprintf(base, sizeof base, "/%s_%05u", potential ? "potential" : "psi", dump_index);   // :1155-1158, :1171-1174
    stamp("transfer enqueued");
    spawn_writers(sim, buf, !potential, ticket, dir + base + "_real", dir + base + "_imag", potential);   // io.rs:54-55
    stamp("writers started");
    return MSM_OK;
}

int msm_sim_reserve_dump_buffers(msm_sim* sim, int32_t n) {
    if (!sim || n < 0) return sfail(sim, MSM_E_ARG, "msm_sim_reserve_dump_buffers: bad argument");
    DumpPool& P = sim->pool;
    const size_t want = std::min<size_t>((size_t)n, pool_cap());
    for (;;) {
        {
            std::lock_guard<std::mutex> lock(P.mu);
            if (P.all.size() >= want) return MSM_OK;
        }
        void* p = nullptr;
        const int rc = msm_host_alloc(sim->ctx, 2 * sizeof(double) * sim_cells(sim), &p);
        if (rc) {
            sim->err = msm_last_error(sim->ctx);
            return rc;
        }
        {
            std::lock_guard<std::mutex> lock(P.mu);
            P.all.push_back((double*)p);
            P.free_.push_back((double*)p);
        }
        P.cv.notify_all();
    }
}

int msm_sim_dump(msm_sim* sim, int32_t stream, const char* root_dir, const char* sim_name, uint32_t dump_index) {
    return dump_field(sim, stream, root_dir, sim_name, dump_index, false);
}

int msm_sim_dump_potential(msm_sim* sim, int32_t stream, const char* root_dir, const char* sim_name, uint32_t dump_index) {
    return dump_field(sim, stream, root_dir, sim_name, dump_index, true);
}

}  // extern "C"
