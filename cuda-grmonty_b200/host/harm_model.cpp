/*
 * harm_model.cpp -- host side of the B200 path (see harm_model.hpp).
 *
 * Reference behaviour restated (cuda_grmonty/): units harm_model.cpp:64-79,139-141; dump format :81-232;
 * init_geometry :242-266; init_weight_table :268-306; init_nint_table :308-338; hot cross-section table
 * hotcross.cpp:60-79,108-179; emissivity tables jnu_mixed.cpp:57-73,127-148; report_spectrum
 * harm_model.cpp:416-471.  The transport itself is NOT here: run_simulation() calls the CUDA library through
 * include/grmonty_b200.h.
 */
#include "harm_model.hpp"

#include <dlfcn.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <barrier>
#include <charconv>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <filesystem>
#include <format>
#include <functional>
#include <fstream>
#include <numbers>
#include <stdexcept>
#include <thread>

#include "../../include/grmonty_b200.h"
#include "../csrc/gm_params.h" /* constants only (same values as the reference's consts.hpp) */

namespace harm {

using std::numbers::pi;

/* ---- logging (spdlog is not part of this build; same information, plain stderr) ------------------------- */
static int g_verbosity = 2;
void set_verbosity(int level) { g_verbosity = level; }
void log_info(const char *fmt, ...) {
    if (g_verbosity > 2)
        return;
    va_list ap;
    va_start(ap, fmt);
    std::fputs("[info] ", stderr);
    std::vfprintf(stderr, fmt, ap);
    std::fputc('\n', stderr);
    va_end(ap);
}

static unsigned hw_threads(int requested) {
    if (requested > 0)
        return (unsigned)requested;
    const unsigned n = std::thread::hardware_concurrency();
    return n ? n : 1;
}

template <typename F> static void parallel_for(int n, unsigned nthreads, F &&body) {
    nthreads = std::max(1u, std::min<unsigned>(nthreads, (unsigned)n));
    if (nthreads == 1) {
        for (int i = 0; i < n; ++i)
            body(i);
        return;
    }
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < nthreads; ++t)
        pool.emplace_back([&, t] {
            for (int i = (int)t; i < n; i += (int)nthreads)
                body(i);
        });
    for (auto &th : pool)
        th.join();
}

/* ---- construction / units (reference harm_model.cpp:64-79) --------------------------------------------- */
HARMModel::HARMModel(int photon_n, double mass_unit) : photon_n_(photon_n) {
    units_.mass_unit = mass_unit;
    units_.l_unit = gm::kGNewt * gm::kMBH / (gm::kCL * gm::kCL);
    units_.t_unit = units_.l_unit / gm::kCL;
    units_.rho_unit = units_.mass_unit / std::pow(units_.l_unit, 3);
    units_.u_unit = units_.rho_unit * gm::kCL * gm::kCL;
    units_.b_unit = gm::kCL * std::sqrt(4.0 * pi * units_.rho_unit);
    units_.n_e_unit = units_.rho_unit / (gm::kMP + gm::kME);
    max_tau_scatt_ = 6.0 * units_.l_unit * units_.rho_unit * 0.4;
    d_tau_k_ = 2.0 * pi * units_.l_unit / (gm::kME * gm::kCL * gm::kCL / gm::kHBAR);
    spectrum_.assign((size_t)kNThBins * kNEBins * kSpecFields, 0.0);
}

/* ---- dump loader ------------------------------------------------------------------------------------------ */
namespace {
struct Tokenizer {
    const char *p, *end;
    bool next(double &v) {
        while (p < end && (*p == ' ' || *p == '\t' || *p == '\r'))
            ++p;
        if (p >= end || *p == '\n')
            return false;
        if (*p == '+')
            ++p;
        auto [q, ec] = std::from_chars(p, end, v);
        if (ec != std::errc()) {
            /* tolerate tokens from_chars rejects (e.g. "nan"): skip to the next blank */
            v = std::nan("");
            while (p < end && *p != ' ' && *p != '\t' && *p != '\n')
                ++p;
            return true;
        }
        p = q;
        return true;
    }
    void end_line() {
        while (p < end && *p != '\n')
            ++p;
        if (p < end)
            ++p;
    }
};
} /* namespace */

/* ---- binary dump cache ------------------------------------------------------------------------
 * The 34-column text dump is 29 MB at 192^2 and ~0.9 GB at 1024^2, and the run needs 8 of the 34 columns plus one
 * scalar.  With `dump_cache` on, read_file() leaves `<dump>.b200cache` (or the same name under `dump_cache_dir`)
 * behind: the 26 header fields, bias_norm and the 8 primitive grids as raw doubles, stamped with the size, the
 * modification time and a hash of the first 64 KB of the text file it was made from.  The next read_file() of the same
 * dump maps straight into the grids; any mismatch (other file, edited file, truncated or foreign cache) falls back to
 * the text parser and rewrites the cache.  Little-endian IEEE doubles, as everything else this library reads. */
namespace {
constexpr char kDumpCacheMagic[8] = {'G', 'M', 'B', '2', 'D', 'U', 'M', 'P'};
constexpr uint32_t kDumpCacheVersion = 2;
struct DumpCacheHeader {
    char magic[8];
    uint32_t version, header_fields;
    uint64_t src_size;
    int64_t src_mtime_ns;
    uint64_t src_head_hash, n_zones;
    uint64_t payload_hash; /* of the 8 grids, as written: a torn or bit-rotten cache is not loaded */
    double bias_norm;
    double header[26];
};
/* 64-bit words, multiply-xor in four independent lanes (the multiply latency would otherwise bound it to ~4 GB/s):
 * an integrity check that runs near memory speed, not a cryptographic hash */
uint64_t hash_words(uint64_t h, const double *p, size_t n) {
    uint64_t a[4] = {h, h ^ 0x9E3779B97F4A7C15ull, h ^ 0xC2B2AE3D27D4EB4Full, h ^ 0x165667B19E3779F9ull};
    size_t i = 0;
    for (; i + 4 <= n; i += 4) {
        uint64_t w[4];
        std::memcpy(w, p + i, sizeof(w));
        for (int l = 0; l < 4; ++l) {
            a[l] = (a[l] ^ w[l]) * 0x9E3779B97F4A7C15ull;
            a[l] ^= a[l] >> 29;
        }
    }
    for (; i < n; ++i) {
        uint64_t w;
        std::memcpy(&w, p + i, sizeof(w));
        a[0] = (a[0] ^ w) * 0x9E3779B97F4A7C15ull;
        a[0] ^= a[0] >> 29;
    }
    uint64_t r = a[0];
    for (int l = 1; l < 4; ++l) {
        r = (r ^ a[l]) * 0x9E3779B97F4A7C15ull;
        r ^= r >> 32;
    }
    return r;
}
uint64_t fnv1a(const char *p, size_t n) {
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < n; ++i)
        h = (h ^ (unsigned char)p[i]) * 1099511628211ull;
    return h;
}
struct SourceStamp {
    uint64_t size = 0;
    int64_t mtime_ns = 0;
    uint64_t head_hash = 0;
};
SourceStamp stamp_of(const std::string &path) {
    SourceStamp st;
    st.size = (uint64_t)std::filesystem::file_size(path);
    st.mtime_ns = (int64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(
                      std::filesystem::last_write_time(path).time_since_epoch())
                      .count();
    std::ifstream in(path, std::ios::binary);
    std::string head((size_t)std::min<uint64_t>(st.size, 65536), '\0');
    in.read(head.data(), (std::streamsize)head.size());
    st.head_hash = fnv1a(head.data(), head.size());
    return st;
}
} /* namespace */

std::string HARMModel::dump_cache_path(const std::string &filepath) const {
    namespace fs = std::filesystem;
    if (dump_cache_dir.empty())
        return filepath + ".b200cache";
    return (fs::path(dump_cache_dir) / (fs::path(filepath).filename().string() + ".b200cache")).string();
}

void HARMModel::apply_header(const double h[26], const std::string &filepath) {
    Header &H = header_;
    H.t = h[0];
    H.n[0] = (int)h[1];
    H.n[1] = (int)h[2];
    H.x_start[0] = 0.0;
    H.x_start[1] = h[3];
    H.x_start[2] = h[4];
    H.x_start[3] = 0.0;
    H.dx[0] = 1.0;
    H.dx[1] = h[5];
    H.dx[2] = h[6];
    H.dx[3] = 2.0 * pi;
    H.x_stop[0] = 1.0;
    H.x_stop[1] = H.x_start[1] + H.n[0] * H.dx[1];
    H.x_stop[2] = H.x_start[2] + H.n[1] * H.dx[2];
    H.x_stop[3] = 2.0 * pi;
    H.t_final = h[7];
    H.n_step = (int)h[8];
    H.a = h[9];
    H.gamma = h[10];
    H.courant = h[11];
    H.dt_dump = h[12];
    H.dt_log = h[13];
    H.dt_img = h[14];
    H.dt_rdump = (int)h[15];
    H.cnt_dump = (int)h[16];
    H.cnt_img = (int)h[17];
    H.cnt_rdump = (int)h[18];
    H.dt = h[19];
    H.lim = (int)h[20];
    H.failed = (int)h[21];
    H.r_in = h[22];
    H.r_out = h[23];
    H.h_slope = h[24];
    H.r_0 = h[25];
    if (H.n[0] < 1 || H.n[1] < 1)
        throw std::runtime_error("Bad HARM dump header in " + filepath);
    /* reference :139-141 */
    const double two_temp_gamma = 0.5 * ((1. + 2. / 3. * (gm::kTpOverTe + 1.) / (gm::kTpOverTe + 2.)) + H.gamma);
    units_.theta_e_unit = (two_temp_gamma - 1.) * (gm::kMP / gm::kME) / (1. + gm::kTpOverTe);
    rh_ = 1.0 + std::sqrt(1.0 - H.a * H.a);
    x1_min_ = std::log(rh_);
}

bool HARMModel::load_dump_cache(const std::string &filepath) {
    const std::string cpath = dump_cache_path(filepath);
    std::error_code ec;
    if (!std::filesystem::exists(cpath, ec))
        return false;
    std::ifstream in(cpath, std::ios::binary | std::ios::ate);
    if (!in.is_open())
        return false;
    const uint64_t csize = (uint64_t)in.tellg();
    in.seekg(0);
    DumpCacheHeader ch;
    if (csize < sizeof(ch) || !in.read(reinterpret_cast<char *>(&ch), sizeof(ch)))
        return false;
    const SourceStamp st = stamp_of(filepath);
    if (std::memcmp(ch.magic, kDumpCacheMagic, 8) != 0 || ch.version != kDumpCacheVersion || ch.header_fields != 26 ||
        ch.src_size != st.size || ch.src_mtime_ns != st.mtime_ns || ch.src_head_hash != st.head_hash)
        return false;
    const int64_t n0 = (int64_t)ch.header[1], n1 = (int64_t)ch.header[2];
    if (n0 < 1 || n1 < 1 || ch.n_zones != (uint64_t)(n0 * n1) ||
        csize != sizeof(ch) + 8ull * ch.n_zones * sizeof(double))
        return false;
    std::vector<double> *grids[8] = {&data_.k_rho, &data_.u,   &data_.u_1, &data_.u_2,
                                     &data_.u_3,   &data_.b_1, &data_.b_2, &data_.b_3};
    /* read into scratch first: a cache that fails its check must leave the model untouched */
    std::vector<double> scratch[8];
    uint64_t h = 0x243F6A8885A308D3ull;
    for (auto &g : scratch) {
        g.resize(ch.n_zones);
        if (!in.read(reinterpret_cast<char *>(g.data()), (std::streamsize)(ch.n_zones * sizeof(double))))
            return false;
        h = hash_words(h, g.data(), g.size());
    }
    if (h != ch.payload_hash)
        return false;
    for (int k = 0; k < 8; ++k)
        grids[k]->swap(scratch[k]);
    apply_header(ch.header, filepath);
    bias_norm_ = ch.bias_norm;
    return true;
}

void HARMModel::store_dump_cache(const std::string &filepath, const double h[26]) const {
    /* best effort: a read-only dump directory must not fail the run */
    try {
        /* the temporary name is unique per process and call: the ranks of a multi-GPU job all read the same dump */
        static std::atomic<unsigned> serial{0};
        const std::string cpath = dump_cache_path(filepath),
                          tmp = cpath + ".tmp." + std::to_string((long long)getpid()) + "." + std::to_string(serial++);
        DumpCacheHeader ch{};
        std::memcpy(ch.magic, kDumpCacheMagic, 8);
        ch.version = kDumpCacheVersion;
        ch.header_fields = 26;
        const SourceStamp st = stamp_of(filepath);
        ch.src_size = st.size;
        ch.src_mtime_ns = st.mtime_ns;
        ch.src_head_hash = st.head_hash;
        ch.n_zones = data_.k_rho.size();
        ch.bias_norm = bias_norm_;
        std::memcpy(ch.header, h, sizeof(ch.header));
        const std::vector<double> *grids[8] = {&data_.k_rho, &data_.u,   &data_.u_1, &data_.u_2,
                                               &data_.u_3,   &data_.b_1, &data_.b_2, &data_.b_3};
        ch.payload_hash = 0x243F6A8885A308D3ull;
        for (auto *g : grids)
            ch.payload_hash = hash_words(ch.payload_hash, g->data(), g->size());
        {
            std::ofstream out(tmp, std::ios::binary | std::ios::trunc);
            if (!out.is_open())
                return;
            out.write(reinterpret_cast<const char *>(&ch), sizeof(ch));
            for (auto *g : grids)
                out.write(reinterpret_cast<const char *>(g->data()), (std::streamsize)(g->size() * sizeof(double)));
            if (!out.good()) {
                out.close();
                std::filesystem::remove(tmp);
                return;
            }
        }
        std::filesystem::rename(tmp, cpath); /* atomic: a concurrent reader sees the old or the new file */
    } catch (const std::exception &e) {
        log_info("Dump cache not written: %s", e.what());
    }
}

void HARMModel::read_file(std::string filepath) {
    log_info("Reading file %s", filepath.c_str());
    if (!std::filesystem::exists(filepath))
        throw std::runtime_error("File does not exist " + filepath);
    read_from_cache_ = false;
    if (dump_cache && load_dump_cache(filepath)) {
        read_from_cache_ = true;
        log_info("Reading file done (binary cache %s)", dump_cache_path(filepath).c_str());
        return;
    }
    std::ifstream in(filepath, std::ios::binary | std::ios::ate);
    if (!in.is_open())
        throw std::runtime_error("Cannot open file " + filepath);
    const std::streamsize size = in.tellg();
    in.seekg(0);
    std::string buf((size_t)size, '\0');
    in.read(buf.data(), size);
    Tokenizer tk{buf.data(), buf.data() + buf.size()};

    /* header: 26 fields (reference :99-137) */
    double h[26] = {0};
    for (int i = 0; i < 26; ++i)
        if (!tk.next(h[i]))
            break;
    tk.end_line();
    apply_header(h, filepath);
    const Header &H = header_;

    const size_t nz = (size_t)H.n[0] * H.n[1];
    std::vector<double> *grids[8] = {&data_.k_rho, &data_.u,   &data_.u_1, &data_.u_2,
                                     &data_.u_3,   &data_.b_1, &data_.b_2, &data_.b_3};
    for (auto *g : grids)
        g->assign(nz, 0.0);
    const double d_v = H.dx[1] * H.dx[2] * H.dx[3];
    double v = 0.0;
    bias_norm_ = 0.0;
    /* data: 34 columns per zone, i outer / j inner (reference :175-215); used: 4..11 primitives, 33 gdet */
    for (size_t z = 0; z < nz; ++z) {
        double c[34] = {0};
        for (int k = 0; k < 34; ++k)
            if (!tk.next(c[k]))
                break;
        tk.end_line();
        for (int k = 0; k < 8; ++k)
            (*grids[k])[z] = c[4 + k];
        const double g_det = c[33];
        bias_norm_ += d_v * g_det * std::pow(data_.u[z] / data_.k_rho[z] * units_.theta_e_unit, 2.);
        v += d_v * g_det;
    }
    bias_norm_ /= v;
    if (dump_cache)
        store_dump_cache(filepath, h);
    log_info("Reading file done");
}

/* ---- geometry (reference gcov_func :499-530, gcon_func :473-497, get_bl_coord :1632-1637) --------------- */
void HARMModel::gcov(const double x[4], double g[4][4]) const {
    std::memset(g, 0, sizeof(double) * 16);
    const double a = header_.a, hs = header_.h_slope;
    const double r = std::exp(x[1]) + header_.r_0;
    const double th = pi * x[2] + ((1.0 - hs) / 2.0) * std::sin(2.0 * pi * x[2]);
    const double st = std::fabs(std::sin(th)) + gm::kEps, ct = std::cos(th);
    const double s2 = st * st, rho2 = r * r + a * a * ct * ct;
    const double rfac = r - header_.r_0;
    const double hfac = pi + (1.0 - hs) * pi * std::cos(2.0 * pi * x[2]);
    g[0][0] = (-1.0 + 2.0 * r / rho2);
    g[0][1] = (2.0 * r / rho2) * rfac;
    g[0][3] = (-2.0 * a * r * s2 / rho2);
    g[1][0] = g[0][1];
    g[1][1] = (1.0 + 2.0 * r / rho2) * rfac * rfac;
    g[1][3] = (-a * s2 * (1.0 + 2.0 * r / rho2)) * rfac;
    g[2][2] = rho2 * hfac * hfac;
    g[3][0] = g[0][3];
    g[3][1] = g[1][3];
    g[3][3] = s2 * (rho2 + a * a * s2 * (1.0 + 2.0 * r / rho2));
}

void HARMModel::gcon(const double x[4], double g[4][4]) const {
    std::memset(g, 0, sizeof(double) * 16);
    const double a = header_.a, hs = header_.h_slope;
    const double r = std::exp(x[1]) + header_.r_0;
    const double th = pi * x[2] + ((1.0 - hs) / 2.0) * std::sin(2.0 * pi * x[2]);
    const double st = std::fabs(std::sin(th)) + gm::kEps, ct = std::cos(th);
    const double irho2 = 1.0 / (r * r + a * a * ct * ct);
    const double hfac = pi + (1.0 - hs) * pi * std::cos(2.0 * pi * x[2]);
    g[0][0] = -1.0 - 2.0 * r * irho2;
    g[0][1] = 2.0 * irho2;
    g[1][0] = g[0][1];
    g[1][1] = irho2 * (r * (r - 2.0) + a * a) / (r * r);
    g[1][3] = a * irho2 / r;
    g[2][2] = irho2 / (hfac * hfac);
    g[3][1] = g[1][3];
    g[3][3] = irho2 / (st * st);
}

/* Laplace expansion along the first row, 3x3 minors by cofactors */
static double det3(const double r0[3], const double r1[3], const double r2[3]) {
    return r0[0] * (r1[1] * r2[2] - r1[2] * r2[1]) - r0[1] * (r1[0] * r2[2] - r1[2] * r2[0]) +
           r0[2] * (r1[0] * r2[1] - r1[1] * r2[0]);
}
static double det4(const double m[4][4]) {
    double acc = 0.0, sign = 1.0;
    for (int c = 0; c < 4; ++c) {
        double minor[3][3];
        for (int r = 1; r < 4; ++r) {
            int cc = 0;
            for (int k = 0; k < 4; ++k)
                if (k != c)
                    minor[r - 1][cc++] = m[r][k];
        }
        const double term = m[0][c] * det3(minor[0], minor[1], minor[2]);
        acc = (c == 0) ? term : acc + sign * term;
        sign = -sign;
    }
    return acc;
}

void HARMModel::init_geometry() {
    log_info("Initializing HARM model geometry");
    const int n0 = header_.n[0], n1 = header_.n[1];
    det_.assign((size_t)n0 * n1, 0.0);
    parallel_for(n0, hw_threads(init_threads), [&](int i) {
        for (int j = 0; j < n1; ++j) {
            const double x[4] = {header_.x_start[0], header_.x_start[1] + (i + 0.5) * header_.dx[1],
                                 header_.x_start[2] + (j + 0.5) * header_.dx[2], header_.x_start[3]};
            double g[4][4];
            gcov(x, g);
            det_[(size_t)i * n1 + j] = std::sqrt(std::abs(det4(g)));
        }
    });
    log_info("Initializing HARM model geometry done");
}

/* ---- hot cross-section table (reference hotcross.cpp:60-79, :108-179) ------------------------------------- */
static double hc_klein_nishina(double w) {
    if (w < 1.0e-3)
        return (1.0 - 2.0 * w);
    return (3.0 / 4.0) * (2.0 / (w * w) + (1.0 / (2.0 * w) - (1.0 + w) / (w * w * w)) * std::log(1.0 + 2.0 * w) +
                          (1.0 + w) / ((1.0 + 2.0 * w) * (1.0 + 2.0 * w)));
}

/* total_compton_cross_num with the (gamma-independent) Bessel factor evaluated once per cell instead of once
 * per quadrature node -- the reference spends its 33 s start-up in std::cyl_bessel_k */
static double total_compton_cross_num(double w, double theta_e) {
    if (std::isnan(w))
        return 0.0;
    if (theta_e < gm::kHcMinT && w < gm::kHcMinW)
        return gm::kSigmaThomson;
    if (theta_e < gm::kHcMinT)
        return hc_klein_nishina(w) * gm::kSigmaThomson;
    const double k2f = (theta_e > 1.0e-2) ? std::cyl_bessel_k(2, 1.0 / theta_e) * std::exp(1.0 / theta_e)
                                          : std::sqrt(pi * theta_e / 2.0);
    double cross = 0.0;
    for (double mu_e = -1.0 + 0.5 * gm::kHcDMuE; mu_e < 1.0; mu_e += gm::kHcDMuE) {
        for (double gamma_e = 1.0 + 0.5 * theta_e * gm::kHcDGammaE; gamma_e < 1.0 + gm::kHcMaxGamma * theta_e;
             gamma_e += theta_e * gm::kHcDGammaE) {
            const double dnd = ((gamma_e * std::sqrt(gamma_e * gamma_e - 1.) / (theta_e * k2f)) *
                                std::exp(-(gamma_e - 1.) / theta_e));
            const double f = 0.5 * dnd;
            const double v = std::sqrt(gamma_e * gamma_e - 1.0) / gamma_e;
            const double we = w * gamma_e * (1.0 - mu_e * v);
            const double boostcross = hc_klein_nishina(we) * (1.0 - mu_e * v);
            cross += theta_e * gm::kHcDMuE * gm::kHcDGammaE * boostcross * f;
        }
    }
    return cross * gm::kSigmaThomson;
}

/* On-disk copy of the table: it depends on nothing but the grid constants of consts.hpp:97-112, which
 * the header repeats so that a file written for another grid is not used. */
namespace {
struct HotcrossCacheHeader {
    char magic[8]; /* "GMB2HOTX" */
    uint32_t version, n_w, n_t, reserved;
    double min_w, max_w, min_t, max_t;
    uint64_t payload_hash; /* of the table, filled in by store / checked by load */
};
HotcrossCacheHeader hotcross_cache_header() {
    HotcrossCacheHeader h{};
    std::memcpy(h.magic, "GMB2HOTX", 8);
    h.version = 2;
    h.n_w = kHcNW;
    h.n_t = kHcNT;
    h.min_w = gm::kHcMinW;
    h.max_w = gm::kHcMaxW;
    h.min_t = gm::kHcMinT;
    h.max_t = gm::kHcMaxT;
    return h;
}
} /* namespace */

bool HARMModel::load_hotcross_cache() {
    if (hotcross_cache.empty())
        return false;
    std::ifstream in(hotcross_cache, std::ios::binary | std::ios::ate);
    if (!in.is_open())
        return false;
    const size_t n = (size_t)(kHcNW + 1) * (kHcNT + 1);
    const HotcrossCacheHeader want = hotcross_cache_header();
    HotcrossCacheHeader got;
    if ((size_t)in.tellg() != sizeof(got) + n * sizeof(double))
        return false;
    in.seekg(0);
    if (!in.read(reinterpret_cast<char *>(&got), sizeof(got)))
        return false;
    const uint64_t stored_hash = got.payload_hash;
    got.payload_hash = 0;
    if (std::memcmp(&got, &want, sizeof(got)) != 0)
        return false;
    std::vector<double> table(n);
    if (!in.read(reinterpret_cast<char *>(table.data()), (std::streamsize)(n * sizeof(double))) ||
        hash_words(0x13198A2E03707344ull, table.data(), n) != stored_hash)
        return false;
    hotcross_.swap(table);
    return true;
}

void HARMModel::store_hotcross_cache() const {
    if (hotcross_cache.empty())
        return;
    try {
        static std::atomic<unsigned> serial{0};
        const std::string tmp =
            hotcross_cache + ".tmp." + std::to_string((long long)getpid()) + "." + std::to_string(serial++);
        HotcrossCacheHeader h = hotcross_cache_header();
        h.payload_hash = hash_words(0x13198A2E03707344ull, hotcross_.data(), hotcross_.size());
        {
            std::ofstream out(tmp, std::ios::binary | std::ios::trunc);
            if (!out.is_open())
                return;
            out.write(reinterpret_cast<const char *>(&h), sizeof(h));
            out.write(reinterpret_cast<const char *>(hotcross_.data()),
                      (std::streamsize)(hotcross_.size() * sizeof(double)));
            if (!out.good())
                return;
        }
        std::filesystem::rename(tmp, hotcross_cache);
    } catch (const std::exception &e) {
        log_info("Hotcross cache not written: %s", e.what());
    }
}

void HARMModel::init_hotcross_table() {
    log_info("Initializing HARM model hotcross");
    hotcross_from_cache_ = load_hotcross_cache();
    if (hotcross_from_cache_) {
        log_info("Initializing HARM model hotcross done (cache %s)", hotcross_cache.c_str());
        return;
    }
    hotcross_.assign((size_t)(kHcNW + 1) * (kHcNT + 1), 0.0);
    const double l_min_w = std::log10(gm::kHcMinW), l_min_t = std::log10(gm::kHcMinT);
    const double d_l_w = std::log10(gm::kHcMaxW / gm::kHcMinW) / kHcNW;
    const double d_l_t = std::log10(gm::kHcMaxT / gm::kHcMinT) / kHcNT;
    parallel_for(kHcNW + 1, hw_threads(init_threads), [&](int i) {
        for (int j = 0; j <= kHcNT; ++j) {
            const double l_w = l_min_w + i * d_l_w, l_t = l_min_t + j * d_l_t;
            hotcross_[(size_t)i * (kHcNT + 1) + j] =
                std::log10(total_compton_cross_num(std::pow(10.0, l_w), std::pow(10.0, l_t)));
        }
    });
    store_hotcross_cache();
    log_info("Initializing HARM model hotcross done");
}

/* ---- emissivity tables (reference jnu_mixed.cpp:57-73, :127-148) ---------------------------------------- */
/* Gauss-Legendre nodes/weights on [-1,1] by Newton iteration on P_n */
static void gauss_legendre(int n, std::vector<double> &x, std::vector<double> &w) {
    x.resize(n);
    w.resize(n);
    for (int i = 0; i < (n + 1) / 2; ++i) {
        double z = std::cos(pi * (i + 0.75) / (n + 0.5)), pp = 0.0;
        for (int it = 0; it < 100; ++it) {
            double p1 = 1.0, p2 = 0.0;
            for (int j = 1; j <= n; ++j) {
                const double p3 = p2;
                p2 = p1;
                p1 = ((2.0 * j - 1.0) * z * p2 - (j - 1.0) * p3) / j;
            }
            pp = n * (z * p1 - p2) / (z * z - 1.0);
            const double dz = p1 / pp;
            z -= dz;
            if (std::fabs(dz) < 1e-16)
                break;
        }
        x[i] = -z;
        x[n - 1 - i] = z;
        w[i] = w[n - 1 - i] = 2.0 / ((1.0 - z * z) * pp * pp);
    }
}

static double jnu_integrand(double th, double k) {
    const double sin_th = std::sin(th);
    const double x = k / sin_th;
    if (sin_th < 1.0e-150 || x > 2.0e8)
        return 0.0;
    return sin_th * sin_th * std::pow(std::sqrt(x) + gm::kJnuCst * std::pow(x, 1.0 / 6.0), 2.0) *
           std::exp(-std::pow(x, 1.0 / 3.0));
}

void HARMModel::init_emiss_tables() {
    log_info("Initializing HARM model emission tables");
    /* F(K) = 4 pi int_0^{pi/2} jnu_integrand d(theta).  The reference integrates adaptively (Gauss-Kronrod 61,
     * eps_rel = 1e-6); here: composite 32-point Gauss-Legendre on 64 panels, converged far below 1e-6. */
    std::vector<double> gx, gw;
    gauss_legendre(32, gx, gw);
    const int panels = 64;
    const double l_min_k = std::log(gm::kJnuMinK), d_l_k = std::log(gm::kJnuMaxK / gm::kJnuMinK) / kNESamp;
    parallel_for(kNESamp + 1, hw_threads(init_threads), [&](int i) {
        const double k = std::exp(i * d_l_k + l_min_k);
        double sum = 0.0;
        const double hp = (pi / 2.0) / panels;
        for (int p = 0; p < panels; ++p) {
            const double a = p * hp, mid = a + 0.5 * hp;
            double s = 0.0;
            for (size_t q = 0; q < gx.size(); ++q)
                s += gw[q] * jnu_integrand(mid + 0.5 * hp * gx[q], k);
            sum += 0.5 * hp * s;
        }
        f_[i] = std::log(4 * pi * sum);
    });
    const double l_min_t = std::log(gm::kThetaEMin), d_l_t = std::log(gm::kJnuMaxT / gm::kThetaEMin) / kNESamp;
    for (int i = 0; i <= kNESamp; ++i) {
        const double t = std::exp(i * d_l_t + l_min_t);
        k2_[i] = std::log(std::cyl_bessel_k(2, 1.0 / t));
    }
    log_info("Initializing HARM model emission tables done");
}

/* reference k2_eval jnu_mixed.cpp:102-111, f_eval :113-125 */
static double interp_exp(const double *tab, double lx, double l_min, double d_l) {
    double d_i = (lx - l_min) / d_l;
    int i = (int)d_i;
    i = std::min(i, kNESamp - 1);
    d_i -= i;
    return std::exp((1.0 - d_i) * tab[i] + d_i * tab[i + 1]);
}
double HARMModel::k2_eval(double theta_e) const {
    if (theta_e < gm::kThetaEMin)
        return 0.0;
    if (theta_e > gm::kJnuMaxT)
        return 2.0 * theta_e * theta_e;
    return interp_exp(k2_.data(), std::log(theta_e), std::log(gm::kThetaEMin),
                      std::log(gm::kJnuMaxT / gm::kThetaEMin) / kNESamp);
}
double HARMModel::f_eval(double theta_e, double b_mag, double nu) const {
    const double k = gm::kJnuKFac * nu / (b_mag * theta_e * theta_e);
    if (k > gm::kJnuMaxK)
        return 0.0;
    if (k < gm::kJnuMinK) {
        const double x = std::pow(k, 1.0 / 3.0);
        return x * (37.67503800178 + 2.240274341836 * x);
    }
    return interp_exp(f_.data(), std::log(k), std::log(gm::kJnuMinK), std::log(gm::kJnuMaxK / gm::kJnuMinK) / kNESamp);
}

/* zone-centre n_e, theta_e, |B| (reference get_fluid_zone :538-593; only what the weight table needs) */
HARMModel::ZoneFluid HARMModel::fluid_zone(int i, int j) const {
    const int n1 = header_.n[1];
    const size_t z = (size_t)i * n1 + j;
    const double x[4] = {header_.x_start[0], header_.x_start[1] + (i + 0.5) * header_.dx[1],
                         header_.x_start[2] + (j + 0.5) * header_.dx[2], header_.x_start[3]};
    double gc[4][4], gn[4][4];
    gcov(x, gc);
    gcon(x, gn);
    const double v_con[4] = {0.0, data_.u_1[z], data_.u_2[z], data_.u_3[z]};
    const double bp[4] = {0.0, data_.b_1[z], data_.b_2[z], data_.b_3[z]};
    ZoneFluid f;
    f.n_e = data_.k_rho[z] * units_.n_e_unit;
    f.theta_e = (data_.u[z] / f.n_e) * units_.n_e_unit * units_.theta_e_unit;
    double v_dot_v = 0.0;
    for (int a = 1; a < 4; ++a)
        for (int b = 1; b < 4; ++b)
            v_dot_v += gc[a][b] * v_con[a] * v_con[b];
    const double v_fac = std::sqrt(-1.0 / gn[0][0] * (1.0 + std::abs(v_dot_v)));
    double u_con[4], u_cov[4], b_con[4], b_cov[4];
    u_con[0] = -v_fac * gn[0][0];
    for (int a = 1; a < 4; ++a)
        u_con[a] = v_con[a] - v_fac * gn[0][a];
    for (int a = 0; a < 4; ++a)
        u_cov[a] = gc[a][0] * u_con[0] + gc[a][1] * u_con[1] + gc[a][2] * u_con[2] + gc[a][3] * u_con[3];
    double u_dot_b = 0.0;
    for (int a = 1; a < 4; ++a)
        u_dot_b += u_cov[a] * bp[a];
    b_con[0] = u_dot_b;
    for (int a = 1; a < 4; ++a)
        b_con[a] = (bp[a] + u_con[a] * u_dot_b) / u_con[0];
    for (int a = 0; a < 4; ++a)
        b_cov[a] = gc[a][0] * b_con[0] + gc[a][1] * b_con[1] + gc[a][2] * b_con[2] + gc[a][3] * b_con[3];
    f.b = std::sqrt(b_con[0] * b_cov[0] + b_con[1] * b_cov[1] + b_con[2] * b_cov[2] + b_con[3] * b_cov[3]) *
          units_.b_unit;
    return f;
}

/* reference init_weight_table :268-306 (zone sums kept in zone order so that the table does not depend on the
 * number of threads: each thread owns whole frequency columns) */
void HARMModel::init_weight_table() {
    log_info("Initializing super photon weight table");
    const int n0 = header_.n[0], n1 = header_.n[1];
    const double l_nu_min = std::log(gm::kNuMin);
    const double d_l_nu = (std::log(gm::kNuMax) - std::log(gm::kNuMin)) / kNESamp;
    const double s_fac = header_.dx[1] * header_.dx[2] * header_.dx[3] * units_.l_unit * units_.l_unit * units_.l_unit;
    /* per-zone prefactor and state, computed once */
    std::vector<double> fac((size_t)n0 * n1, 0.0), te((size_t)n0 * n1, 0.0), bb((size_t)n0 * n1, 0.0);
    parallel_for(n0, hw_threads(init_threads), [&](int i) {
        for (int j = 0; j < n1; ++j) {
            const size_t z = (size_t)i * n1 + j;
            const ZoneFluid f = fluid_zone(i, j);
            if (f.n_e == 0.0 || f.theta_e < gm::kThetaEMin)
                continue;
            const double k2 = k2_eval(f.theta_e);
            fac[z] = (gm::kJcst * f.n_e * f.b * f.theta_e * f.theta_e / k2) * s_fac * det_[z];
            te[z] = f.theta_e;
            bb[z] = f.b;
        }
    });
    parallel_for(kNESamp + 1, hw_threads(init_threads), [&](int k) {
        const double nu = std::exp(k * d_l_nu + l_nu_min);
        double sum = 0.0;
        for (size_t z = 0; z < fac.size(); ++z)
            if (te[z] != 0.0)
                sum += fac[z] * f_eval(te[z], bb[z], nu);
        weight_[k] = std::log(sum / (gm::kHPL * photon_n_));
    });
    log_info("Initializing super photon weight table done");
}

/* reference init_nint_table :308-338 */
void HARMModel::init_nint_table() {
    log_info("Initializing nint table");
    nint_.assign(kNint + 1, 0.0);
    dndlnu_max_.assign(kNint + 1, 0.0);
    const double l_nu_min = std::log(gm::kNuMin);
    const double d_l_nu = (std::log(gm::kNuMax) - std::log(gm::kNuMin)) / kNESamp;
    const double l_b_min = std::log(gm::kBthsqMin), d_l_b = std::log(gm::kBthsqMax / gm::kBthsqMin) / kNint;
    std::array<double, kNESamp> nu_j, ew_j;
    for (int j = 0; j < kNESamp; ++j) {
        nu_j[j] = std::exp(j * d_l_nu + l_nu_min);
        ew_j[j] = std::exp(weight_[j]) + 1.0e-100;
    }
    parallel_for(kNint + 1, hw_threads(init_threads), [&](int i) {
        double nint = 0.0, dndlnu_max = 0.0;
        const double b_mag = std::exp(i * d_l_b + l_b_min);
        for (int j = 0; j < kNESamp; ++j) {
            const double dn = f_eval(1.0, b_mag, nu_j[j]) / ew_j[j];
            if (dn > dndlnu_max)
                dndlnu_max = dn;
            nint += d_l_nu * dn;
        }
        nint *= header_.dx[1] * header_.dx[2] * header_.dx[3] * units_.l_unit * units_.l_unit * units_.l_unit *
                std::numbers::sqrt2 * gm::kEE * gm::kEE * gm::kEE / (27.0 * gm::kME * gm::kCL * gm::kCL) *
                (1.0 / gm::kHPL);
        nint_[i] = std::log(nint);
        dndlnu_max_[i] = std::log(dndlnu_max);
    });
    log_info("Initializing nint table done");
}

/* ---- the CUDA library, loaded on demand (run_simulation and the device table builders) ------------------- */
namespace {
struct CudaLib {
    void *h = nullptr;
    decltype(&grmonty_b200_create) create = nullptr;
    decltype(&grmonty_b200_run) run = nullptr;
    decltype(&grmonty_b200_set_progress) set_progress = nullptr;
    decltype(&grmonty_b200_nccl_comm_init_all) nccl_comm_init_all = nullptr;
    decltype(&grmonty_b200_nccl_comm_destroy) nccl_comm_destroy = nullptr;
    decltype(&grmonty_b200_allreduce) allreduce = nullptr;
    decltype(&grmonty_b200_result) result = nullptr;
    decltype(&grmonty_b200_destroy) destroy = nullptr;
    decltype(&grmonty_b200_last_error) last_error = nullptr;
    decltype(&grmonty_b200_hotcross_table) hotcross_table = nullptr;
    decltype(&grmonty_b200_init_tables) init_tables = nullptr;
};

std::string self_dir() {
    Dl_info info;
    if (dladdr((void *)&self_dir, &info) && info.dli_fname)
        return std::filesystem::path(info.dli_fname).parent_path().string();
    return ".";
}

CudaLib load_cuda_lib(const std::string &hint) {
    std::vector<std::string> candidates;
    if (!hint.empty())
        candidates.push_back(hint);
    if (const char *e = std::getenv("GRMONTY_B200_LIB"))
        candidates.push_back(e);
    candidates.push_back(self_dir() + "/libgrmonty_b200.so");
    candidates.push_back("libgrmonty_b200.so");
    CudaLib L;
    std::string tried;
    for (const auto &c : candidates) {
        L.h = dlopen(c.c_str(), RTLD_NOW | RTLD_LOCAL);
        if (L.h)
            break;
        tried += "\n  " + c + ": " + dlerror();
    }
    if (!L.h)
        throw std::runtime_error("cannot load the CUDA library libgrmonty_b200.so (there is no CPU fallback):" + tried);
#define SYM(name)                                                                 \
    L.name = reinterpret_cast<decltype(L.name)>(dlsym(L.h, "grmonty_b200_" #name)); \
    if (!L.name)                                                                  \
        throw std::runtime_error("libgrmonty_b200.so lacks grmonty_b200_" #name);
    SYM(create) SYM(run) SYM(set_progress) SYM(allreduce) SYM(nccl_comm_init_all) SYM(nccl_comm_destroy) SYM(result) SYM(destroy) SYM(last_error) SYM(hotcross_table) SYM(init_tables)
#undef SYM
    return L;
}
} /* namespace */

void HARMModel::fill_config(grmonty_b200_config &cfg) const {
    std::memset(&cfg, 0, sizeof(cfg));
    cfg.abi_version = GRMONTY_B200_ABI_VERSION;
    cfg.struct_size = sizeof(cfg);
    cfg.n0 = header_.n[0];
    cfg.n1 = header_.n[1];
    cfg.x_start1 = header_.x_start[1];
    cfg.x_start2 = header_.x_start[2];
    cfg.dx1 = header_.dx[1];
    cfg.dx2 = header_.dx[2];
    cfg.dx3 = header_.dx[3];
    cfg.x_stop1 = header_.x_stop[1];
    cfg.x_stop2 = header_.x_stop[2];
    cfg.a = header_.a;
    cfg.h_slope = header_.h_slope;
    cfg.r_0 = header_.r_0;
    cfg.mass_unit = units_.mass_unit;
    cfg.l_unit = units_.l_unit;
    cfg.t_unit = units_.t_unit;
    cfg.rho_unit = units_.rho_unit;
    cfg.u_unit = units_.u_unit;
    cfg.b_unit = units_.b_unit;
    cfg.theta_e_unit = units_.theta_e_unit;
    cfg.n_e_unit = units_.n_e_unit;
    cfg.k_rho = data_.k_rho.data();
    cfg.u = data_.u.data();
    cfg.u_1 = data_.u_1.data();
    cfg.u_2 = data_.u_2.data();
    cfg.u_3 = data_.u_3.data();
    cfg.b_1 = data_.b_1.data();
    cfg.b_2 = data_.b_2.data();
    cfg.b_3 = data_.b_3.data();
    cfg.geom_det = det_.data();
    cfg.hotcross = hotcross_.data();
    cfg.f = f_.data();
    cfg.k2 = k2_.data();
    cfg.weight = weight_.data();
    cfg.nint = nint_.data();
    cfg.dndlnu_max = dndlnu_max_.data();
    cfg.photon_n = photon_n_;
    cfg.bias_norm = bias_norm_;
    cfg.max_tau_scatt0 = max_tau_scatt_;
    cfg.seed = options.seed;
    cfg.rank = options.rank;
    cfg.world = options.world;
    cfg.device = options.device;
    cfg.threads_per_block = options.threads_per_block;
    cfg.kernel = options.kernel;
    cfg.gen_overlap = options.gen_overlap;
    cfg.blocks_per_sm = options.blocks_per_sm;
    cfg.queue_capacity = options.queue_capacity;
    cfg.gen0 = options.gen0;
    cfg.gen_cap = options.gen_cap;
    cfg.gen_budget = options.gen_budget;
    cfg.gen_fine_from = options.gen_fine_from;
    cfg.gen_fine_div = options.gen_fine_div;
}

/* The grid-dependent tables (geometry, weight, nint) and the hot cross-section table on options.device, through
 * grmonty_b200_init_tables / grmonty_b200_hotcross_table.  The F(K) and K2 tables (201 entries each)
 * stay on the host and must exist. */
void HARMModel::init_tables_on_device(bool with_hotcross) {
    log_info("Initializing geometry, weight and nint tables on device %d", options.device);
    CudaLib L = load_cuda_lib(options.cuda_library);
    const size_t nz = (size_t)header_.n[0] * header_.n[1];
    det_.assign(nz, 0.0);
    nint_.assign(kNint + 1, 0.0);
    dndlnu_max_.assign(kNint + 1, 0.0);
    if (with_hotcross) {
        hotcross_from_cache_ = load_hotcross_cache();
        if (!hotcross_from_cache_) {
            hotcross_.assign((size_t)(kHcNW + 1) * (kHcNT + 1), 0.0);
            if (L.hotcross_table(options.device, hotcross_.data()) != GRMONTY_B200_OK)
                throw std::runtime_error(std::string("grmonty_b200_hotcross_table: ") + L.last_error(nullptr));
            store_hotcross_cache();
        }
    }
    grmonty_b200_config cfg;
    fill_config(cfg);
    double ms = 0.0;
    if (L.init_tables(&cfg, det_.data(), weight_.data(), nint_.data(), dndlnu_max_.data(), &ms) != GRMONTY_B200_OK)
        throw std::runtime_error(std::string("grmonty_b200_init_tables: ") + L.last_error(nullptr));
    log_info("Initializing tables on device done (%.3f ms of kernels)", ms);
}

void HARMModel::init() {
    if (options.device_tables) {
        init_emiss_tables();
        init_tables_on_device(true);
        return;
    }
    init_geometry();
    init_hotcross_table();
    init_emiss_tables();
    init_weight_table();
    init_nint_table();
}

/* One GPU's share of the run through the C ABI: create -> run -> (all-reduce) -> result -> destroy.  `collect`: this
 * rank's results (after the all-reduce: the whole job's) go into spectrum_ / stats_.  `sync` (multi-GPU in one
 * process): every rank calls it once between the run and the collective and learns whether all ranks are still
 * alive -- a rank that failed must not leave the others blocked inside NCCL. */
void HARMModel::run_share(int rank, int world, int device, void *nccl_comm, bool collect,
                          const std::function<bool(bool)> &sync) {
    const auto start = std::chrono::steady_clock::now();
    CudaLib L = load_cuda_lib(options.cuda_library);
    grmonty_b200_config cfg;
    fill_config(cfg);
    cfg.rank = rank;
    cfg.world = world;
    cfg.device = device;

    grmonty_b200_ctx *ctx = nullptr;
    std::string error;
    if (L.create(&ctx, &cfg) != 0)
        error = std::string("grmonty_b200_create: ") + L.last_error(nullptr);
    auto fail = [&](const char *what) { error = std::string(what) + ": " + L.last_error(ctx); };
    /* progress like the reference's once-a-second "Rate" line (harm_model.cpp:397-403): the library calls back after
     * every generation; runs shorter than a second never print */
    struct Progress {
        std::chrono::steady_clock::time_point t_last;
        int64_t done_last = 0;
    } prog{start, 0};
    if (error.empty() && options.progress && collect)
        L.set_progress(ctx, [](void *user, int64_t done, int64_t total) {
            auto *p = static_cast<Progress *>(user);
            const auto now = std::chrono::steady_clock::now();
            const double dt = std::chrono::duration<double>(now - p->t_last).count();
            if (dt > 1.0) {
                log_info("Rate %.2f ph/s, position %lld of %lld", (double)(done - p->done_last) / dt, (long long)done,
                         (long long)total);
                p->t_last = now;
                p->done_last = done;
            }
        }, &prog);
    if (error.empty() && L.run(ctx) != 0)
        fail("grmonty_b200_run");
    const bool all_ok = sync ? sync(error.empty()) : error.empty();
    if (all_ok && world > 1 && nccl_comm && L.allreduce(ctx, nccl_comm, nullptr) != 0)
        fail("grmonty_b200_allreduce");
    uint64_t counts[3] = {0, 0, 0};
    grmonty_b200_stats st{};
    double max_tau = 0.0;
    if (all_ok && error.empty() && collect &&
        L.result(ctx, spectrum_.data(), counts, &max_tau, &st) != 0)
        fail("grmonty_b200_result");
    if (ctx)
        L.destroy(ctx);
    if (!error.empty())
        throw std::runtime_error(error);
    if (!all_ok)
        throw std::runtime_error("another rank of the job failed");
    if (!collect)
        return;
    stats_.max_tau_scatt = max_tau;
    stats_.created = counts[0];
    stats_.scattered = counts[1];
    stats_.recorded = counts[2];
    stats_.kernel_ms = st.kernel_ms;
    stats_.transport_ms = st.transport_ms;
    stats_.n_tracked = st.n_tracked;
    stats_.n_steps = st.n_steps;
    stats_.n_push_attempts = st.n_push_attempts;
    stats_.n_interactions = st.n_interactions;
    stats_.n_scatter_events = st.n_scatter_events;
    stats_.n_generations = st.n_generations;
    stats_.n_kernel_launches = st.n_kernel_launches;
}

void HARMModel::run_simulation() {
    const auto start = std::chrono::steady_clock::now();
    log_info("Starting main loop");
    if (options.gpus > 1) {
        /* one host thread per GPU of this box, each with its own context (rank g of options.gpus on device g) and a
         * communicator from ncclCommInitAll; rank 0 collects the all-reduced result */
        if (options.world != 1)
            throw std::runtime_error("run_simulation: options.gpus > 1 and options.world > 1 are mutually exclusive");
        CudaLib L = load_cuda_lib(options.cuda_library);
        const int n = options.gpus;
        std::vector<void *> comms((size_t)n, nullptr);
        if (L.nccl_comm_init_all(comms.data(), n, nullptr) != 0)
            throw std::runtime_error(std::string("grmonty_b200_nccl_comm_init_all: ") + L.last_error(nullptr));
        std::atomic<int> n_failed{0};
        std::barrier gate(n);
        std::vector<std::string> errors((size_t)n);
        std::vector<std::thread> pool;
        for (int g = 0; g < n; ++g)
            pool.emplace_back([&, g] {
                try {
                    run_share(g, n, g, comms[(size_t)g], g == 0, [&](bool ok) {
                        if (!ok)
                            n_failed.fetch_add(1);
                        gate.arrive_and_wait();
                        return n_failed.load() == 0;
                    });
                } catch (const std::exception &e) {
                    errors[(size_t)g] = e.what();
                }
            });
        for (auto &t : pool)
            t.join();
        for (void *c : comms)
            L.nccl_comm_destroy(c);
        for (int g = 0; g < n; ++g)
            if (!errors[(size_t)g].empty() && errors[(size_t)g] != "another rank of the job failed")
                throw std::runtime_error("GPU " + std::to_string(g) + ": " + errors[(size_t)g]);
    } else {
        if (options.world > 1 && !options.nccl_comm && !options.external_reduce)
            throw std::runtime_error(
                "run_simulation: world > 1 needs options.nccl_comm (or options.external_reduce if the caller sums the "
                "per-rank spectra itself): refusing to report a rank-partial spectrum");
        run_share(options.rank, options.world, options.device, options.nccl_comm, true, nullptr);
    }
    stats_.seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - start).count();
    /* same final summary as the reference (harm_model.cpp:409-413) */
    log_info("Final rate %.2f ph/s", stats_.created / stats_.seconds);
    log_info("Super photons:");
    log_info("\tcreated: %llu", (unsigned long long)stats_.created);
    log_info("\tscattered: %llu", (unsigned long long)stats_.scattered);
    log_info("\trecorded: %llu", (unsigned long long)stats_.recorded);
}

/* ---- spectrum file (reference report_spectrum :416-471, d_omega_func :532-536) ---------------------------- */
double HARMModel::d_omega(double x2i, double x2f) const {
    const double hs = header_.h_slope;
    return 2.0 * pi *
           (-std::cos(pi * x2f + 0.5 * (1.0 - hs) * std::sin(2 * pi * x2f)) +
            std::cos(pi * x2i + 0.5 * (1.0 - hs) * std::sin(2 * pi * x2i)));
}

void HARMModel::report_spectrum(std::string filepath) {
    const double dx2 = (header_.x_stop[2] - header_.x_start[2]) / (2 * kNThBins);
    log_info("Writing spectrum to file %s", filepath.c_str());
    std::ofstream out(filepath);
    if (!out.is_open()) {
        std::fprintf(stderr, "[error] Cannot open file %s\n", filepath.c_str());
        return;
    }
    enum { DN_DLE = 0, DE_DLE = 1, X1I_AV = 4, X2I_SQ = 5, X3F_SQ = 6, TAU_ABS = 7, TAU_SCATT = 8 };
    auto S = [&](int j, int i, int f) { return spectrum_[((size_t)j * kNEBins + i) * kSpecFields + f]; };
    const double l_e_0 = std::log(1.0e-12);
    double max_tau_scatt = 0.0, l = 0.0;
    std::string line;
    for (int i = 0; i < kNEBins; ++i) {
        line.clear();
        line += std::format("{:10.5g} ", (i * gm::kSpecDLE + l_e_0) / std::numbers::ln10);
        for (int j = 0; j < kNThBins; ++j) {
            const double d_om = 2.0 * d_omega(j * dx2, (j + 1) * dx2);
            double nu_lnu = (gm::kME * gm::kCL * gm::kCL) * (4.0 * pi / d_om) * (1.0 / gm::kSpecDLE);
            nu_lnu *= S(j, i, DE_DLE);
            nu_lnu /= gm::kLSun;
            const double dn = S(j, i, DN_DLE) + gm::kEps;
            const double tau_scatt = S(j, i, TAU_SCATT) / dn;
            line += std::format("{:10.5g} ", nu_lnu);
            line += std::format("{:10.5g} ", S(j, i, TAU_ABS) / dn);
            line += std::format("{:10.5g} ", tau_scatt);
            line += std::format("{:10.5g} ", S(j, i, X1I_AV) / dn);
            line += std::format("{:10.5g} ", std::sqrt(std::abs(S(j, i, X2I_SQ) / dn)));
            line += std::format("{:10.5g} ", std::sqrt(std::abs(S(j, i, X3F_SQ) / dn)));
            if (tau_scatt > max_tau_scatt)
                max_tau_scatt = tau_scatt;
            l += nu_lnu * d_om * gm::kSpecDLE;
        }
        out << line << '\n';
    }
    out.close();
    luminosity_ = l;
    max_tau_reported_ = max_tau_scatt;
    log_info("Writing spectrum done");
    log_info("\tlumosity: %g", l);
    log_info("\tmax_tau_scatt: %g", max_tau_scatt);
}

/* ---- full-field binary spectrum --------------------------------------------------------------
 * The text file keeps 6 derived columns per angle bin; 7 of the 13 accumulated fields (dn_dle, nph, nscatt, ne_0,
 * theta_e_0, b_0, e_0) never reach it (reference harm_model.cpp:443-455 vs :1323-1334).  This side-file is the whole
 * [6][200][13] accumulator in harm::Spectrum field order plus the run counters, for parity checks and for merging
 * runs (spectra are additive).  Layout: SpectrumBinHeader, then 6*200*13 little-endian doubles. */
namespace {
struct SpectrumBinHeader {
    char magic[8]; /* "GMB2SPEC" */
    uint32_t version, n_th, n_e, n_fields;
    uint64_t created, scattered, recorded;
    double max_tau_scatt, mass_unit, photon_n, a, h_slope, x_start2, x_stop2;
};
} /* namespace */

void HARMModel::report_spectrum_binary(std::string filepath) {
    log_info("Writing binary spectrum to file %s", filepath.c_str());
    SpectrumBinHeader h{};
    std::memcpy(h.magic, "GMB2SPEC", 8);
    h.version = 1;
    h.n_th = kNThBins;
    h.n_e = kNEBins;
    h.n_fields = kSpecFields;
    h.created = stats_.created;
    h.scattered = stats_.scattered;
    h.recorded = stats_.recorded;
    h.max_tau_scatt = stats_.max_tau_scatt;
    h.mass_unit = units_.mass_unit;
    h.photon_n = photon_n_;
    h.a = header_.a;
    h.h_slope = header_.h_slope;
    h.x_start2 = header_.x_start[2];
    h.x_stop2 = header_.x_stop[2];
    std::ofstream out(filepath, std::ios::binary | std::ios::trunc);
    if (!out.is_open())
        throw std::runtime_error("Cannot open file " + filepath);
    out.write(reinterpret_cast<const char *>(&h), sizeof(h));
    out.write(reinterpret_cast<const char *>(spectrum_.data()), (std::streamsize)(spectrum_.size() * sizeof(double)));
    if (!out.good())
        throw std::runtime_error("Short write to " + filepath);
}

} /* namespace harm */
