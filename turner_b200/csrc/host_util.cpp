// Host-side pieces of the reference's main() that stay on the CPU: camera set-up,
// tone mapping and the P3 writer. libm's expf/powf are used on purpose so that, given the
// same linear buffer, the 8-bit output equals the reference's (lib/effects.h:15-48).
#include "../../include/turner_b200.h"
#include "host_util.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <string>
#include <thread>
#include <vector>

extern "C" {

// Camera(trafo, aiCamera), lib/types.h:92-105: position = trafo * (0,0,0), rotation = upper 3x3,
// delta_x = tan(hfov) evaluated in double (the unqualified tan() there binds to ::tan(double)),
// delta_y = delta_x / aspect; image height = int(width / aspect), main.cpp:178-179.
int32_t trn_camera_setup(const float* trafo4x4, float hfov, float aspect, int32_t width, trn_camera* cam, int32_t* height) {
    if (!trafo4x4 || !cam) return trn::fail(TRN_ERR_INVALID, "null argument");
    if (!(aspect > 0)) return trn::fail(TRN_ERR_INVALID, "aspect must be > 0 (config.h:30)");
    const float* m = trafo4x4;
    for (int r = 0; r < 3; ++r) {
        cam->pos[r] = m[4 * r] * 0.f + m[4 * r + 1] * 0.f + m[4 * r + 2] * 0.f + m[4 * r + 3];
        for (int c = 0; c < 3; ++c) cam->rot[3 * r + c] = m[4 * r + c];
    }
    cam->delta_x = static_cast<float>(std::tan(static_cast<double>(hfov)));
    cam->delta_y = cam->delta_x / aspect;
    if (height) *height = static_cast<int32_t>(width / aspect);
    return TRN_OK;
}

int32_t trn_tonemap(const float* rgba_sum, uint64_t npix, int32_t pixel_samples, float exposure, int32_t gamma_enabled,
                    float inverse_gamma, float* rgba_out) {
    if (!rgba_sum || !rgba_out || pixel_samples < 1) return trn::fail(TRN_ERR_INVALID, "bad argument");
    const float n = static_cast<float>(pixel_samples);
    // per pixel independent (the reference does it inside each row task, main.cpp:216-223): split over the host threads;
    // same libm calls per pixel whatever the split, so the output does not depend on the thread count
    auto span = [=](uint64_t lo, uint64_t hi) {
        for (uint64_t i = lo; i < hi; ++i) {
            float c[4];
            for (int k = 0; k < 4; ++k) c[k] = rgba_sum[4 * i + k] / n;           // main.cpp:216
            // a channel that is exactly +0 stays +0 through both steps (1 - expf(-0 * e) = 0, powf(0, g) = 0 for g > 0):
            // same bits without the two libm calls -- most of a frame whose background is black
            const bool zero_ok = exposure >= 0.f && std::isfinite(exposure) && (!gamma_enabled || inverse_gamma > 0.f);
            for (int k = 0; k < 3; ++k) {
                if (zero_ok && c[k] == 0.f && !std::signbit(c[k])) continue;
                c[k] = 1 - expf(-c[k] * exposure);                                // effects.h:15-17
                if (gamma_enabled) c[k] = powf(c[k], inverse_gamma);              // effects.h:36-38
            }
            std::memcpy(rgba_out + 4 * i, c, sizeof c);
        }
    };
    unsigned nt = std::thread::hardware_concurrency();
    if (const char* e = std::getenv("TRN_HOST_THREADS")) nt = static_cast<unsigned>(std::atoi(e));
    nt = static_cast<unsigned>(std::min<uint64_t>(std::max(1u, std::min(nt, 64u)), std::max<uint64_t>(1, npix / 65536)));
    if (nt <= 1) {
        span(0, npix);
        return TRN_OK;
    }
    // chunks of 8 Ki pixels handed out by an atomic counter: the work per pixel is very uneven (a black background pixel
    // costs nothing), contiguous equal shares would leave most threads idle
    constexpr uint64_t kChunk = 8192;
    std::atomic<uint64_t> next{0};
    auto worker = [&]() {
        for (;;) {
            const uint64_t lo = next.fetch_add(kChunk);
            if (lo >= npix) break;
            span(lo, std::min<uint64_t>(npix, lo + kChunk));
        }
    };
    std::vector<std::thread> pool;
    for (unsigned t = 1; t < nt; ++t) pool.emplace_back(worker);
    worker();
    for (auto& th : pool) th.join();
    return TRN_OK;
}

uint64_t trn_write_p3(const float* rgba, int32_t width, int32_t height, char* buf, uint64_t cap) {
    // lib/raster.h:79-100: header, then per pixel "\n" at row starts else " ", three setw(3) ints of
    // int(clamp(255*c*a, 0, 255)); main.cpp:242 appends std::endl.
    std::string head = "P3\n" + std::to_string(width) + " " + std::to_string(height) + "\n255";
    const uint64_t npix = static_cast<uint64_t>(width) * height;
    const uint64_t need = head.size() + npix * 12 + 1;
    if (!buf || cap == 0) return need;
    uint64_t pos = 0;
    auto put = [&](const char* s, size_t n) {
        for (size_t i = 0; i < n && pos < cap; ++i) buf[pos++] = s[i];
    };
    put(head.data(), head.size());
    // every field is one of 256 right-aligned 3-character strings: table lookup instead of a printf per pixel
    // (the image of BASELINE config 4 is 436 MB of text)
    char lut[256][3];
    for (int v = 0; v < 256; ++v) {
        lut[v][0] = v >= 100 ? static_cast<char>('0' + v / 100) : ' ';
        lut[v][1] = v >= 10 ? static_cast<char>('0' + (v / 10) % 10) : ' ';
        lut[v][2] = static_cast<char>('0' + v % 10);
    }
    // clamp(255*c*a, 0, 255) -> int (raster.h:91-96). NaN (powf of a negative colour, a degenerate vertex normal) compares
    // false both ways: the reference then prints whatever int(NaN) is; here it becomes 0 -- the table index stays in range.
    auto q = [](float v) {
        const float c = (v > 0.f) ? (v < 255.f ? v : 255.f) : 0.f;
        return static_cast<int>(c);
    };
    char tmp[12];
    for (uint64_t i = 0; i < npix; ++i) {
        const float* p = rgba + 4 * i;
        tmp[0] = (i % static_cast<uint64_t>(width) == 0) ? '\n' : ' ';
        const int c0 = q(255 * p[0] * p[3]), c1 = q(255 * p[1] * p[3]), c2 = q(255 * p[2] * p[3]);
        std::memcpy(tmp + 1, lut[c0], 3);
        tmp[4] = ' ';
        std::memcpy(tmp + 5, lut[c1], 3);
        tmp[8] = ' ';
        std::memcpy(tmp + 9, lut[c2], 3);
        put(tmp, 12);
    }
    put("\n", 1);
    return need;
}

} // extern "C"
