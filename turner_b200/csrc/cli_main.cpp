// `pathtracer` / `raycaster` / `raytracer` executables: the reference's command line, stderr report and P3 output
// (main.cpp:88-244, pathtracer.h / raycaster.h USAGE, config.h, lib/output.h:101-113,
// lib/progress_bar.h) in front of the CUDA render loop in libturner_b200.so.
// One source, three binaries: -DTRN_CLI_PATHTRACER, -DTRN_CLI_RAYCASTER or -DTRN_CLI_RAYTRACER (the reference links
// main.cpp against one integrator TU the same way, CMakeLists.txt:43-67).
//
// Deliberately not replicated: ./kdtree.cache (main.cpp:142-167 loads whatever tree is lying in the CWD,
// for any scene); -t only sizes the host kd-tree build.
#include "../../include/turner_b200.h"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

namespace {

#if defined(TRN_CLI_RAYCASTER)
constexpr bool kPathtracer = false, kRaytracer = false;
const char* kProgram = "raycaster";
#elif defined(TRN_CLI_RAYTRACER)
constexpr bool kPathtracer = false, kRaytracer = true;
const char* kProgram = "raytracer";
#else
constexpr bool kPathtracer = true, kRaytracer = false;
const char* kProgram = "pathtracer";
#endif

std::string usage_text() {
    std::ostringstream u;
    u << "Usage: " << kProgram << " <filename> [options]\n\n"
      << "Options:\n"
         "  -w --width=<px>                   Image width in pixels [default: 640].\n"
         "  -a --aspect=<num>                 Aspect ratio; used when the scene camera does not\n"
         "                                    define one [default: 1].\n"
         "  --background=<3x float>           Background colour, one or three floats [default: 0 0 0].\n"
         "  -t --threads=<int>                Host threads (kd-tree build) [default: 1].\n"
         "  --inverse-gamma=<float>           1/gamma of the gamma correction [default: 0.454545].\n"
         "  --no-gamma-correction             Switch gamma correction off.\n"
         "  --exposure=<float>                Exposure [default: 1].\n"
         "  -v --verbose                      Print the configuration.\n"
         "  --gpus=<int>                      GPUs to split the pixel samples over [default: 1].\n"
         "  --seed=<int>                      Run seed of the hemisphere streams [default: 1].\n"
         "  --dump-linear=<file>              Also write the linear RGBA sums (raw float32).\n"
         "  --dump-hits=<file>                Also write primary-hit triangle ids (raw uint32).\n"
         "  --kdtree-cache=<file>             Load the triangles + kd-tree from this kdtree.cache if it exists (refused\n"
         "                                    when stale), else build them and write it [default: none].\n"
         "  --kd-builder=<host|gpu>           Build the kd-tree on the host (the reference's tree, node for node) or on the\n"
         "                                    GPU (binned SAH, ~20x sooner at 1 M triangles) [default: host].\n"
         "  -h --help                         Show this text.\n\n";
    if (kPathtracer) {
        u << "Pathtracer options:\n"
             "  -d --max-depth=<int>              Maximum recursion depth [default: 3].\n"
             "  -p --pixel-samples=<int>          Samples per pixel [default: 1].\n"
             "  -m --monte-carlo-samples=<int>    Monte Carlo samples per hit [default: 8].\n";
    } else if (kRaytracer) {
        u << "Raytracer options:\n"
             "  -d --max-depth=<int>              Maximum recursion depth [default: 3].\n"
             "  --shadow=<float>                  Intensity of shadow [default: 0.5].\n";
    } else {
        u << "Raycaster options:\n"
             "  --max-visibility=<float>          Anything farther away is dark [default: 2.0].\n";
    }
    return u.str();
}

struct Options {
    std::string filename;
    long width = 640;
    float aspect = 1;
    std::string background = "0 0 0";
    long threads = 1;
    float inverse_gamma = 0.454545f;
    bool gamma = true;
    float exposure = 1;
    bool verbose = false;
    long gpus = 1;
    unsigned long long seed = 1;
    std::string dump_linear, dump_hits, kdtree_cache, kd_builder = "host";
    // TracerConfig defaults (config.h:106-117); the pathtracer USAGE overrides -m to 8 (pathtracer.h:24)
    long max_depth = 3;
    float max_visibility = 2;
    float shadow_intensity = 0.5f;
    long pixel_samples = 1;
    long mc_samples = kPathtracer ? 8 : 1;
};

[[noreturn]] void usage_error(const std::string& msg) {
    // docopt.cpp prints the message and the usage text and exits with -1
    std::cerr << msg << std::endl << usage_text();
    std::exit(-1);
}

struct OptSpec {
    const char* shortname; // "-w" or nullptr
    const char* longname;  // "--width"
    bool takes_value;
    int programs; // bit 0 pathtracer, bit 1 raycaster, bit 2 raytracer
};
constexpr int kThisProgram = kPathtracer ? 1 : (kRaytracer ? 4 : 2);

const OptSpec kSpecs[] = {
    {"-w", "--width", true, 7},          {"-a", "--aspect", true, 7},
    {nullptr, "--background", true, 7},  {"-t", "--threads", true, 7},
    {nullptr, "--inverse-gamma", true, 7}, {nullptr, "--no-gamma-correction", false, 7},
    {nullptr, "--exposure", true, 7},    {"-v", "--verbose", false, 7},
    {nullptr, "--gpus", true, 7},        {nullptr, "--seed", true, 7},
    {nullptr, "--dump-linear", true, 7}, {nullptr, "--dump-hits", true, 7},
    {nullptr, "--kdtree-cache", true, 7}, {nullptr, "--kd-builder", true, 7},
    {"-h", "--help", false, 7},          {"-d", "--max-depth", true, 1 | 4},
    {"-p", "--pixel-samples", true, 1},  {"-m", "--monte-carlo-samples", true, 1},
    {nullptr, "--max-visibility", true, 2}, {nullptr, "--shadow", true, 4},
};

const OptSpec* find_spec(const std::string& name) {
    for (const auto& s : kSpecs) {
        if (!(s.programs & kThisProgram)) continue;
        if (name == s.longname || (s.shortname && name == s.shortname)) return &s;
    }
    return nullptr;
}

void apply(Options& o, const OptSpec& s, const std::string& v) {
    const std::string n = s.longname;
    try {
        if (n == "--width") o.width = std::stol(v);
        else if (n == "--aspect") o.aspect = std::stof(v);
        else if (n == "--background") o.background = v;
        else if (n == "--threads") o.threads = std::stol(v);
        else if (n == "--inverse-gamma") o.inverse_gamma = std::stof(v);
        else if (n == "--no-gamma-correction") o.gamma = false;
        else if (n == "--exposure") o.exposure = std::stof(v);
        else if (n == "--verbose") o.verbose = true;
        else if (n == "--gpus") o.gpus = std::stol(v);
        else if (n == "--seed") o.seed = std::stoull(v);
        else if (n == "--dump-linear") o.dump_linear = v;
        else if (n == "--dump-hits") o.dump_hits = v;
        else if (n == "--kdtree-cache") o.kdtree_cache = v;
        else if (n == "--kd-builder") o.kd_builder = v;
        else if (n == "--max-depth") o.max_depth = std::stol(v);
        else if (n == "--pixel-samples") o.pixel_samples = std::stol(v);
        else if (n == "--monte-carlo-samples") o.mc_samples = std::stol(v);
        else if (n == "--max-visibility") o.max_visibility = std::stof(v);
        else if (n == "--shadow") o.shadow_intensity = std::stof(v);
        else if (n == "--help") {
            std::cout << usage_text();
            std::exit(0);
        }
    } catch (const std::exception&) {
        usage_error(std::string("bad value for ") + n + ": " + v);
    }
}

// docopt's option grammar: --name=value, --name value, -n value, -nvalue; options anywhere around <filename>
Options parse_args(int argc, const char* const* argv) {
    Options o;
    bool have_file = false;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        if (a.size() > 2 && a[0] == '-' && a[1] == '-') {
            std::string name = a, value;
            bool has_eq = false;
            size_t eq = a.find('=');
            if (eq != std::string::npos) {
                name = a.substr(0, eq);
                value = a.substr(eq + 1);
                has_eq = true;
            }
            const OptSpec* s = find_spec(name);
            if (!s) usage_error("Unexpected argument: " + a);
            if (s->takes_value) {
                if (!has_eq) {
                    if (i + 1 >= argc) usage_error(name + " requires an argument");
                    value = argv[++i];
                }
            } else if (has_eq) {
                usage_error(name + " must not have an argument");
            }
            apply(o, *s, value);
        } else if (a.size() >= 2 && a[0] == '-' && a != "--") {
            // possibly stacked short options: -vw100
            size_t k = 1;
            while (k < a.size()) {
                std::string name = std::string("-") + a[k];
                const OptSpec* s = find_spec(name);
                if (!s) usage_error("Unexpected argument: " + a);
                ++k;
                if (s->takes_value) {
                    std::string value;
                    if (k < a.size()) value = a.substr(k);
                    else if (i + 1 < argc) value = argv[++i];
                    else usage_error(name + " requires an argument");
                    apply(o, *s, value);
                    break;
                }
                apply(o, *s, "");
            }
        } else {
            if (have_file) usage_error("Unexpected argument: " + a);
            o.filename = a;
            have_file = true;
        }
    }
    if (!have_file) usage_error("Arguments did not match expected patterns");
    return o;
}

// Config::parse_color, config.h:45-60: one or three space-separated floats used as they are, alpha 1
bool parse_color(const std::string& s, float out[4]) {
    std::vector<float> vals;
    std::stringstream ss(s);
    std::string item;
    try {
        while (std::getline(ss, item, ' ')) vals.push_back(std::stof(item));
    } catch (const std::exception&) {
        return false;
    }
    if (vals.size() == 1) {
        out[0] = out[1] = out[2] = vals[0];
    } else if (vals.size() == 3) {
        out[0] = vals[0];
        out[1] = vals[1];
        out[2] = vals[2];
    } else {
        return false;
    }
    out[3] = 1;
    return true;
}

// the reference validates with assert(); here a failed check is a message + exit code 2, never an abort
void require(bool ok, const char* what) {
    if (!ok) {
        std::cerr << kProgram << ": invalid configuration: " << what << std::endl;
        std::exit(2);
    }
}

void print_config(const Options& o) { // operator<<(Config) + operator<<(TracerConfig), config.h:85-97,155-165
    std::cerr << "Filename: " << o.filename << std::endl;
    std::cerr << "Aspect ratio: " << o.aspect << std::endl;
    std::cerr << "Image width: " << o.width << std::endl;
    std::cerr << std::endl;
    std::cerr << "Common parameters:" << std::endl;
    std::cerr << "  Number of threads: " << o.threads << std::endl;
    std::cerr << "  Inverse gamma: " << o.inverse_gamma << std::endl;
    std::cerr << "  Exposure: " << o.exposure << std::endl;
    std::cerr << "  Background color: " << o.exposure << std::endl; // sic, config.h:95 prints the exposure here
    std::cerr << "  Gamma correction enabled: " << o.gamma << std::endl;
    std::cerr << std::endl;
    std::cerr << "Tracer parameters (not all applicable):" << std::endl;
    std::cerr << "  Max recursion depth: " << o.max_depth << std::endl;
    std::cerr << "  Max visibility: " << o.max_visibility << std::endl;
    std::cerr << "  Shadow intensity: " << o.shadow_intensity << std::endl;
    std::cerr << "  Number of pixel samples: " << o.pixel_samples << std::endl;
    std::cerr << "  Number of Monte-Carlo samples: " << o.mc_samples << std::endl;
}

void progress(const char* label, double fraction) { // lib/progress_bar.h:14-28
    const float p = static_cast<float>(fraction);
    int bar = static_cast<int>(p * 20);
    if (bar > 20) bar = 20;
    std::cerr << "\r" << std::setw(20) << std::setfill(' ') << std::left << label;
    for (int s = 0; s < bar; ++s) std::cerr << "\xE2\x96\xA0";
    for (int s = 0; s < 20 - bar; ++s) std::cerr << "\xE2\x96\xA1";
    std::cerr << std::setw(7) << std::setfill(' ') << std::right << std::fixed << std::setprecision(2) << (p * 100.0) << '%';
    std::cerr.flush();
}

size_t ms_since(std::chrono::steady_clock::time_point t0) {
    return static_cast<size_t>(std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::steady_clock::now() - t0).count());
}

} // namespace

int main(int argc, char const* argv[]) {
    Options o = parse_args(argc, argv);
    float bg[4];
    require(parse_color(o.background, bg), "--background takes one or three floats");
    require(0 < o.aspect, "0 < aspect");                 // config.h:30
    require(1 <= o.threads, "1 <= threads");             // config.h:31
    require(0 <= o.exposure, "0 <= exposure");           // config.h:32
    require(0 < o.max_depth, "0 < max-depth");           // config.h:121
    require(0 <= o.max_visibility, "0 <= max-visibility"); // config.h:122
    require(1 <= o.pixel_samples, "1 <= pixel-samples"); // config.h:124
    require(1 <= o.mc_samples, "1 <= monte-carlo-samples (0 divides by zero in the reference, pathtracer.cpp:88)");
    require(0 <= o.shadow_intensity && o.shadow_intensity <= 1, "0 <= shadow <= 1"); // config.h:123
    require(1 <= o.width, "1 <= width");
    require(1 <= o.gpus, "1 <= gpus");
    require(o.kd_builder == "host" || o.kd_builder == "gpu", "--kd-builder is host or gpu");
    if (o.verbose) print_config(o);

    std::cerr << "Loading scene..." << std::endl; // main.cpp:97
    trn_loaded_scene ls;
    const bool is_blend = o.filename.size() >= 6 && o.filename.compare(o.filename.size() - 6, 6, ".blend") == 0;
    if ((is_blend ? trn_load_blend(o.filename.c_str(), &ls) : trn_load_soup(o.filename.c_str(), &ls)) != TRN_OK) {
        std::cout << trn_last_error() << std::endl; // main.cpp:104-107: import errors go to stdout, exit code 1
        return 1;
    }
    const float aspect = ls.cam_aspect > 0 ? ls.cam_aspect : o.aspect; // main.cpp:112-116
    trn_camera cam;
    int32_t height = 0;
    if (trn_camera_setup(ls.cam_trafo4x4, ls.cam_hfov, aspect, static_cast<int32_t>(o.width), &cam, &height) != TRN_OK) {
        std::cerr << trn_last_error() << std::endl;
        return 2;
    }

    std::cerr << "Loading triangles and building kd-tree..." << std::endl; // main.cpp:139
    const auto t_load = std::chrono::steady_clock::now();
    setenv("TRN_BUILD_THREADS", std::to_string(o.threads).c_str(), 0);
    trn_scene* scene = nullptr;
    const auto t_kd = std::chrono::steady_clock::now();
    // main.cpp:142-167 reads ./kdtree.cache whenever it exists, whatever scene was asked for; here the file is named
    // explicitly, and a cache whose tree does not belong to its triangles is refused (trn_scene_load_cache)
    bool from_cache = false;
    if (!o.kdtree_cache.empty()) {
        if (FILE* probe = std::fopen(o.kdtree_cache.c_str(), "rb")) {
            std::fclose(probe);
            if (trn_scene_load_cache(o.kdtree_cache.c_str(), &scene) != TRN_OK) {
                std::cerr << trn_last_error() << std::endl;
                return 2;
            }
            from_cache = true;
        }
    }
    if (!from_cache) {
        const int32_t crc = o.kd_builder == "gpu"
                                ? trn_scene_create_gpu(ls.verts, ls.normals, ls.diffuse, ls.reflective, ls.reflectivity, ls.num_triangles, 0, &scene)
                                : trn_scene_create_ex(ls.verts, ls.normals, ls.diffuse, ls.reflective, ls.reflectivity, ls.num_triangles, &scene);
        if (crc != TRN_OK) {
            std::cerr << trn_last_error() << std::endl;
            return 2;
        }
        if (!o.kdtree_cache.empty() && trn_scene_save_cache(scene, o.kdtree_cache.c_str()) != TRN_OK) {
            std::cerr << trn_last_error() << std::endl;
            return 2;
        }
    }
    std::cerr << "KDTree runtime: " << ms_since(t_kd) << std::endl; // main.cpp:168
    trn_scene_info info;
    trn_scene_get_info(scene, &info);
    const size_t loading_ms = ms_since(t_load);

    trn_render_config cfg;
    std::memset(&cfg, 0, sizeof cfg);
    cfg.width = static_cast<int32_t>(o.width);
    cfg.height = height;
    cfg.max_depth = static_cast<int32_t>(o.max_depth);
    cfg.mc_samples = static_cast<int32_t>(o.mc_samples);
    cfg.pixel_samples = static_cast<int32_t>(o.pixel_samples);
    cfg.integrator = kPathtracer ? TRN_PATHTRACER : (kRaytracer ? TRN_RAYTRACER : TRN_RAYCASTER);
    cfg.shadow_intensity = o.shadow_intensity;
    std::memcpy(cfg.bg_rgba, bg, sizeof bg);
    cfg.max_visibility = o.max_visibility;
    cfg.num_lights = ls.num_lights;
    cfg.light = ls.light;
    cfg.seed = o.seed;
    cfg.sample_begin = 0;
    cfg.sample_stride = 1;

    const size_t npix = static_cast<size_t>(cfg.width) * cfg.height;
    std::vector<float> sum(npix * 4), image(npix * 4);
    trn_stats st;
    std::memset(&st, 0, sizeof st);
    std::cerr << "Rendering ";
    progress("Rendering", 0.0);
    const auto t_render = std::chrono::steady_clock::now();
    std::vector<int32_t> devices;
    for (long g = 0; g < o.gpus; ++g) devices.push_back(static_cast<int32_t>(g));
    const int rc = trn_render_multi(scene, devices.data(), static_cast<int32_t>(devices.size()), &cam, &cfg, sum.data(), &st);
    if (rc != TRN_OK) {
        std::cerr << std::endl << kProgram << ": " << trn_last_error() << std::endl;
        return 3;
    }
    // image(x,y) /= pps; exposure; gamma (main.cpp:216-223) stay on the host
    trn_tonemap(sum.data(), npix, cfg.pixel_samples, o.exposure, o.gamma ? 1 : 0, o.inverse_gamma, image.data());
    const size_t runtime_ms = ms_since(t_render);
    progress("Rendering", 1.0);
    std::cerr << std::endl;

    if (!o.dump_linear.empty()) {
        std::ofstream f(o.dump_linear, std::ios::binary);
        f.write(reinterpret_cast<const char*>(sum.data()), static_cast<std::streamsize>(sum.size() * sizeof(float)));
    }
    if (!o.dump_hits.empty()) {
        std::vector<uint32_t> ids(npix * cfg.pixel_samples);
        std::vector<float> rst(ids.size() * 3);
        if (trn_primary_hits(scene, devices[0], &cam, &cfg, ids.data(), rst.data()) == TRN_OK) {
            std::ofstream f(o.dump_hits, std::ios::binary);
            f.write(reinterpret_cast<const char*>(ids.data()), static_cast<std::streamsize>(ids.size() * sizeof(uint32_t)));
        }
    }

    // Stats block, lib/output.h:101-113
    std::cerr.unsetf(std::ios::floatfield);
    std::cerr << std::setprecision(6);
    std::cerr << "Triangles      : " << info.num_triangles << std::endl
              << "Kd-Tree Height : " << info.kdtree_height << std::endl
              << "Rays           : " << st.rays << std::endl
              << "Rays (primary) : " << st.prim_rays << std::endl
              << "Rays/sec       : " << (runtime_ms ? 1000 * st.rays / runtime_ms : 0) << std::endl
              << "Loading time   : " << 1.0 * loading_ms / 1000 << " sec" << std::endl
              << "Rendering time : " << 1.0 * runtime_ms / 1000 << " sec" << std::endl;

    // the image, lib/raster.h:79-100 + std::endl (main.cpp:242)
    const uint64_t need = trn_write_p3(image.data(), cfg.width, cfg.height, nullptr, 0);
    std::string text(need, '\0');
    trn_write_p3(image.data(), cfg.width, cfg.height, &text[0], need);
    std::cout.write(text.data(), static_cast<std::streamsize>(text.size()));
    std::cout.flush();

    trn_scene_destroy(scene);
    trn_loaded_scene_free(&ls);
    return 0;
}
