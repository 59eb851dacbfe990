/*
 * main.cpp -- command-line driver with the reference's flags (cuda_grmonty/main.cpp:20-56):
 *     --photon_n N   --mass_unit M   --harm_dump_path FILE   --spectrum_path FILE   --verbosity LEVEL
 * Both `--flag value`, `--flag=value` and the single-dash spellings abseil accepts (`-photon_n 5000000`,
 * reference README.md:30) work.  Extra flags of the B200 path: --seed, --device, --gpus N (shard the run over N GPUs
 * of this box, one host thread each, NCCL all-reduce of the spectrum at the end), --init_threads, --dump_cache 0|1
 * --device_tables 0|1 (build the init tables on the GPU), --hotcross_cache FILE (on-disk hot cross-section table),
 * (binary cache next to the dump, or under --dump_cache_dir), --spectrum_bin_path FILE (all 13 accumulated fields).
 * Call order is the reference's: HARMModel(photon_n, mass_unit) -> read_file -> init -> run_simulation ->
 * report_spectrum.  (The reference seeds its global mt19937 with 123 at this point; here the seed is the
 * Philox key and is part of the run options.)
 */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "harm_model.hpp"

static bool flag_value(int argc, char **argv, const char *name, std::string &out) {
    const std::string a1 = std::string("--") + name, a2 = std::string("-") + name;
    for (int i = 1; i < argc; ++i) {
        const std::string s = argv[i];
        if ((s == a1 || s == a2) && i + 1 < argc) {
            out = argv[i + 1];
            return true;
        }
        for (const std::string &a : {a1, a2})
            if (s.rfind(a + "=", 0) == 0) {
                out = s.substr(a.size() + 1);
                return true;
            }
    }
    return false;
}

static int parse_verbosity(const std::string &v) {
    const char *names[] = {"trace", "debug", "info", "warning", "error", "critical", "off"};
    for (int i = 0; i < 7; ++i)
        if (v == names[i] || (i == 3 && v == "warn"))
            return i;
    return 2;
}

int main(int argc, char **argv) {
    std::string v;
    int photon_n = 5000000; /* reference default, main.cpp:20 */
    double mass_unit = 4e19;
    std::string harm_dump_path, spectrum_path;
    if (flag_value(argc, argv, "photon_n", v))
        photon_n = (int)std::strtod(v.c_str(), nullptr);
    if (flag_value(argc, argv, "mass_unit", v))
        mass_unit = std::strtod(v.c_str(), nullptr);
    flag_value(argc, argv, "harm_dump_path", harm_dump_path);
    flag_value(argc, argv, "spectrum_path", spectrum_path);
    if (flag_value(argc, argv, "verbosity", v))
        harm::set_verbosity(parse_verbosity(v));

    harm::log_info("Parameters:");
    harm::log_info("\tphoton_n: %d", photon_n);
    harm::log_info("\tmass_unit: %g", mass_unit);
    harm::log_info("\tharm_dump_path: %s", harm_dump_path.c_str());
    harm::log_info("\tspectrum_path: %s", spectrum_path.c_str());

    try {
        harm::HARMModel model(photon_n, mass_unit);
        if (flag_value(argc, argv, "seed", v))
            model.options.seed = std::strtoull(v.c_str(), nullptr, 10);
        if (flag_value(argc, argv, "device", v))
            model.options.device = std::atoi(v.c_str());
        if (flag_value(argc, argv, "gpus", v))
            model.options.gpus = std::atoi(v.c_str());
        if (flag_value(argc, argv, "gen_overlap", v)) /* 1: overlapping generations (default), 2: one launch each */
            model.options.gen_overlap = std::atoi(v.c_str());
        if (flag_value(argc, argv, "queue_capacity", v))
            model.options.queue_capacity = std::atoll(v.c_str());
        if (flag_value(argc, argv, "init_threads", v))
            model.init_threads = std::atoi(v.c_str());
        if (flag_value(argc, argv, "device_tables", v))
            model.options.device_tables = std::atoi(v.c_str()) != 0;
        flag_value(argc, argv, "hotcross_cache", model.hotcross_cache);
        if (flag_value(argc, argv, "dump_cache", v))
            model.dump_cache = std::atoi(v.c_str()) != 0;
        flag_value(argc, argv, "dump_cache_dir", model.dump_cache_dir);
        std::string spectrum_bin_path;
        flag_value(argc, argv, "spectrum_bin_path", spectrum_bin_path);
        model.read_file(harm_dump_path);
        model.init();
        model.run_simulation();
        model.report_spectrum(spectrum_path);
        if (!spectrum_bin_path.empty())
            model.report_spectrum_binary(spectrum_bin_path);
    } catch (const std::exception &e) {
        std::fprintf(stderr, "[error] %s\n", e.what());
        return 1;
    }
    return 0;
}
