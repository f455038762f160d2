// main.cpp -- `raytracing_hw5 <scene.txt> <out.ppm>`: the reference's command line
// (src/main.cpp:6-19, run.sh) on top of the C-ABI.  Extra, optional environment knobs:
//   RTC_DEVICE=<n>   CUDA device (default 0)      RTC_SEED=<n>   RNG seed (default 0)
//   RTC_SAMPLES / RTC_WIDTH / RTC_HEIGHT / RTC_RAY_DEPTH   override the scene file
//   RTC_DEVICES=<list>   render on several devices of the machine, e.g. "0,1,2,3" or "0-7" (samples split over them,
//                        summed over NVLink peer access on the first): Scene::Render uses the whole machine too
// The same program serves the four earlier homework snapshots (hwN/run.sh -> build/raytracing_hwN): the dialect
// is the digit in the name it is called by (raytracing_hw1 .. raytracing_hw4 are links to this binary), or
// RTC_DIALECT=<1..5>.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include <string>
#include <vector>

#include "rtc_b200.h"

// "0,2,3" / "0-7" / "0-3,6" -> device indices
static std::vector<int> parse_devices(const char* text) {
    std::vector<int> out;
    const std::string t(text ? text : "");
    size_t i = 0;
    while (i < t.size()) {
        size_t j = t.find(',', i);
        if (j == std::string::npos) j = t.size();
        const std::string item = t.substr(i, j - i);
        const size_t dash = item.find('-');
        if (!item.empty()) {
            if (dash == std::string::npos) out.push_back(std::atoi(item.c_str()));
            else {
                const int a = std::atoi(item.substr(0, dash).c_str()), b = std::atoi(item.substr(dash + 1).c_str());
                for (int d = a; d <= b && (int)out.size() < 64; ++d) out.push_back(d);
            }
        }
        i = j + 1;
    }
    return out;
}

static int env_int(const char* name, int dflt) {
    const char* v = std::getenv(name);
    return v && *v ? std::atoi(v) : dflt;
}

int main(int argc, const char* argv[]) {
    if (argc < 3) {
        std::fprintf(stderr, "usage: %s <scene.txt> <out.ppm>\n", argv[0]);
        return 2;
    }
    int dialect = RTC_DIALECT_HW5;
    const char* base = std::strrchr(argv[0], '/');
    base = base ? base + 1 : argv[0];
    if (std::strncmp(base, "raytracing_hw", 13) == 0 && base[13] >= '1' && base[13] <= '5' && base[14] == 0) dialect = base[13] - '0';
    dialect = env_int("RTC_DIALECT", dialect);
    const bool timing = std::getenv("RTC_TIMING") != nullptr;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
        return std::chrono::duration<double, std::milli>(b - a).count();
    };
    const auto t_start = now();
    const std::vector<int> devices = parse_devices(std::getenv("RTC_DEVICES"));
    const int device0 = devices.empty() ? env_int("RTC_DEVICE", 0) : devices[0];
    rtc_scene* scene = rtc_scene_load_dialect(argv[1], device0, dialect);
    if (!scene) {
        std::fprintf(stderr, "raytracing_hw5: %s\n", rtc_last_error());
        return 1;
    }
    const auto t_loaded = now();
    int rc = rtc_scene_override(scene, env_int("RTC_WIDTH", -1), env_int("RTC_HEIGHT", -1), env_int("RTC_SAMPLES", -1),
                                env_int("RTC_RAY_DEPTH", -1));
    if (rc != RTC_OK) {
        std::fprintf(stderr, "raytracing_hw5: %s\n", rtc_last_error());
        rtc_scene_free(scene);
        return 1;
    }
    if (devices.size() > 1) rc = rtc_render_ppm_multi(scene, devices.data(), (int)devices.size(), (uint32_t)env_int("RTC_SEED", 0), argv[2]);
    else rc = rtc_render_ppm(scene, (uint32_t)env_int("RTC_SEED", 0), argv[2]);
    if (rc != RTC_OK) std::fprintf(stderr, "raytracing_hw5: %s\n", rtc_last_error());
    const auto t_rendered = now();
    if (timing)
        std::fprintf(stderr, "[rtc timing] %-28s %8.1f ms\n[rtc timing] %-28s %8.1f ms\n",
                     "load (file -> HBM, CUDA init)", ms(t_start, t_loaded), "render + PPM", ms(t_loaded, t_rendered));
    // The image is on disk: leave without unmapping 2 GB of device memory buffer by buffer (0.15 - 1.1 s measured on the
    // B200 box); the driver reclaims everything when the process ends, as it does for the reference's own heap.
    std::fflush(nullptr);
    std::_Exit(rc == RTC_OK ? 0 : 1);
    return rc == RTC_OK ? 0 : 1;
}
