// sbx_gradients.h -- the 19 gradients of default_color_schemes() (src/colorscheme.rs:129-149) as host functions
// t in [0, 1] -> (r, g, b), for sbx_builtin_lut / sbx_builtin_gradient.  Host only; the kernels read tables.
//
// [unvendored] colorous 1.0.12 -- the crate the reference evaluates gradients with (Gradient::eval_continuous,
// src/colorscheme.rs:43,49,51,66,69,82,85) -- is not under /root/reference and no registry is reachable from this
// image.  colorous is a port of d3-scale-chromatic, so the definitions below restate d3-scale-chromatic's published
// ones: 256-entry ramps (magma, viridis, inferno, plasma: luts.inc), closed-form polynomials (turbo, cividis),
// cubehelix interpolation (cubehelix default, cool) and the uniform B-spline ("basis") through the largest ColorBrewer
// class of each scheme (9 colours for the sequential, 11 for the diverging ones).  Channel values are rounded to the
// nearest integer and clamped to [0, 255] as d3 formats them.  A Rust caller fills the tables from colorous itself
// (INTEGRATION.md), which removes the assumption; these built-ins serve C / Python callers, the tests and the bench.
// Second source available here: OpenCV's COLORMAP_TURBO / COLORMAP_CIVIDIS tables (tests/test_oracle_kat.py compares).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

namespace sbx {
namespace gradients {

enum Kind { RAMP, TURBO, CIVIDIS, CUBEHELIX_LONG, BASIS };

struct Def {
    const char *name;            // sbx_builtin_lut name
    const char *reference_name;  // colorous constant, src/colorscheme.rs:130-148
    bool stereo;                 // ColorScheme::new_stereo in the reference (a diverging gradient indexed by pan)
    Kind kind;
    const unsigned char *ramp;   // RAMP: 256 x 3
    double ch[6];                // CUBEHELIX_LONG: (h, s, l) of both ends
    int n_colors;                // BASIS
    uint32_t colors[11];         // BASIS: 0xRRGGBB
};

inline uint8_t to_u8(double v)
{
    if (!(v > 0.0)) return 0;
    if (v >= 255.0) return 255;
    return (uint8_t)std::floor(v + 0.5);
}

// d3-interpolate basis(): uniform cubic B-spline through values[], end points clamped by reflection
inline double basis(const double *v, int n_values, double t)
{
    const int n = n_values - 1;
    int i;
    if (t <= 0.0) { t = 0.0; i = 0; }
    else if (t >= 1.0) { t = 1.0; i = n - 1; }
    else i = (int)std::floor(t * n);
    const double v1 = v[i], v2 = v[i + 1];
    const double v0 = i > 0 ? v[i - 1] : 2.0 * v1 - v2;
    const double v3 = i < n - 1 ? v[i + 2] : 2.0 * v2 - v1;
    const double t1 = (t - (double)i / n) * n, t2 = t1 * t1, t3 = t2 * t1;
    return ((1.0 - 3.0 * t1 + 3.0 * t2 - t3) * v0 + (4.0 - 6.0 * t2 + 3.0 * t3) * v1 + (1.0 + 3.0 * t1 + 3.0 * t2 - 3.0 * t3) * v2 + t3 * v3) / 6.0;
}

inline void eval(const Def &d, double t, uint8_t rgb[3])
{
    if (!(t >= 0.0)) t = 0.0;                                 // NaN and negatives
    if (t > 1.0) t = 1.0;
    switch (d.kind) {
        case RAMP: {                                          // d3 ramp(): range[max(0, min(n - 1, floor(t n)))]
            int i = (int)std::floor(t * 256.0);
            i = i > 255 ? 255 : i;
            std::memcpy(rgb, d.ramp + 3 * i, 3);
            break;
        }
        case TURBO:                                           // d3 interpolateTurbo
            rgb[0] = to_u8(34.61 + t * (1172.33 - t * (10793.56 - t * (33300.12 - t * (38394.49 - t * 14825.05)))));
            rgb[1] = to_u8(23.31 + t * (557.33 + t * (1225.33 - t * (3574.96 - t * (1073.77 + t * 707.56)))));
            rgb[2] = to_u8(27.2 + t * (3211.1 - t * (15327.97 - t * (27814.0 - t * (22569.18 - t * 6838.66)))));
            break;
        case CIVIDIS:                                         // d3 interpolateCividis
            rgb[0] = to_u8(-4.54 - t * (35.34 - t * (2381.73 - t * (6402.7 - t * (7024.72 - t * 2710.57)))));
            rgb[1] = to_u8(32.49 + t * (170.73 + t * (52.82 - t * (131.46 - t * (176.58 - t * 67.37)))));
            rgb[2] = to_u8(81.24 + t * (442.36 - t * (2482.43 - t * (6167.24 - t * (6614.94 - t * 2475.67)))));
            break;
        case CUBEHELIX_LONG: {                                // d3 interpolateCubehelixLong + cubehelix -> rgb
            const double h = (d.ch[0] + t * (d.ch[3] - d.ch[0]) + 120.0) * (3.14159265358979323846 / 180.0);
            const double s = d.ch[1] + t * (d.ch[4] - d.ch[1]);
            const double l = d.ch[2] + t * (d.ch[5] - d.ch[2]);
            const double a = s * l * (1.0 - l), c = std::cos(h), sn = std::sin(h);
            rgb[0] = to_u8(255.0 * (l + a * (-0.14861 * c + 1.78277 * sn)));
            rgb[1] = to_u8(255.0 * (l + a * (-0.29227 * c - 0.90649 * sn)));
            rgb[2] = to_u8(255.0 * (l + a * (1.97294 * c)));
            break;
        }
        case BASIS: {                                         // d3 interpolateRgbBasis(scheme[largest class])
            double ch[3][11];
            for (int i = 0; i < d.n_colors; ++i) {
                ch[0][i] = (double)((d.colors[i] >> 16) & 255u);
                ch[1][i] = (double)((d.colors[i] >> 8) & 255u);
                ch[2][i] = (double)(d.colors[i] & 255u);
            }
            for (int c = 0; c < 3; ++c) rgb[c] = to_u8(basis(ch[c], d.n_colors, t));
            break;
        }
    }
}

#define SBX_BASIS9(...) BASIS, nullptr, {0, 0, 0, 0, 0, 0}, 9, {__VA_ARGS__}
#define SBX_BASIS11(...) BASIS, nullptr, {0, 0, 0, 0, 0, 0}, 11, {__VA_ARGS__}

// in the order of default_color_schemes(), src/colorscheme.rs:129-149
static const Def kDefs[19] = {
    {"red_yellow_blue", "RED_YELLOW_BLUE", true, SBX_BASIS11(0xa50026, 0xd73027, 0xf46d43, 0xfdae61, 0xfee090, 0xffffbf, 0xe0f3f8, 0xabd9e9, 0x74add1, 0x4575b4, 0x313695)},
    {"magma", "MAGMA", false, RAMP, SBX_LUT_MAGMA, {0, 0, 0, 0, 0, 0}, 0, {0}},
    {"viridis", "VIRIDIS", false, RAMP, SBX_LUT_VIRIDIS, {0, 0, 0, 0, 0, 0}, 0, {0}},
    {"red_blue", "RED_BLUE", true, SBX_BASIS11(0x67001f, 0xb2182b, 0xd6604d, 0xf4a582, 0xfddbc7, 0xf7f7f7, 0xd1e5f0, 0x92c5de, 0x4393c3, 0x2166ac, 0x053061)},
    {"spectral", "SPECTRAL", true, SBX_BASIS11(0x9e0142, 0xd53e4f, 0xf46d43, 0xfdae61, 0xfee08b, 0xffffbf, 0xe6f598, 0xabdda4, 0x66c2a5, 0x3288bd, 0x5e4fa2)},
    {"red_yellow_green", "RED_YELLOW_GREEN", true, SBX_BASIS11(0xa50026, 0xd73027, 0xf46d43, 0xfdae61, 0xfee08b, 0xffffbf, 0xd9ef8b, 0xa6d96a, 0x66bd63, 0x1a9850, 0x006837)},
    {"pink_green", "PINK_GREEN", true, SBX_BASIS11(0x8e0152, 0xc51b7d, 0xde77ae, 0xf1b6da, 0xfde0ef, 0xf7f7f7, 0xe6f5d0, 0xb8e186, 0x7fbc41, 0x4d9221, 0x276419)},
    {"purple_orange", "PURPLE_ORANGE", true, SBX_BASIS11(0x2d004b, 0x542788, 0x8073ac, 0xb2abd2, 0xd8daeb, 0xf7f7f7, 0xfee0b6, 0xfdb863, 0xe08214, 0xb35806, 0x7f3b08)},
    {"inferno", "INFERNO", false, RAMP, SBX_LUT_INFERNO, {0, 0, 0, 0, 0, 0}, 0, {0}},
    {"plasma", "PLASMA", false, RAMP, SBX_LUT_PLASMA, {0, 0, 0, 0, 0, 0}, 0, {0}},
    {"cividis", "CIVIDIS", false, CIVIDIS, nullptr, {0, 0, 0, 0, 0, 0}, 0, {0}},
    {"cubehelix", "CUBEHELIX", false, CUBEHELIX_LONG, nullptr, {300.0, 0.5, 0.0, -240.0, 0.5, 1.0}, 0, {0}},
    {"turbo", "TURBO", false, TURBO, nullptr, {0, 0, 0, 0, 0, 0}, 0, {0}},
    {"cool", "COOL", false, CUBEHELIX_LONG, nullptr, {260.0, 0.75, 0.35, 80.0, 1.50, 0.8}, 0, {0}},
    {"reds", "REDS", false, SBX_BASIS9(0xfff5f0, 0xfee0d2, 0xfcbba1, 0xfc9272, 0xfb6a4a, 0xef3b2c, 0xcb181d, 0xa50f15, 0x67000d)},
    {"blues", "BLUES", false, SBX_BASIS9(0xf7fbff, 0xdeebf7, 0xc6dbef, 0x9ecae1, 0x6baed6, 0x4292c6, 0x2171b5, 0x08519c, 0x08306b)},
    {"greens", "GREENS", false, SBX_BASIS9(0xf7fcf5, 0xe5f5e0, 0xc7e9c0, 0xa1d99b, 0x74c476, 0x41ab5d, 0x238b45, 0x006d2c, 0x00441b)},
    {"greys", "GREYS", false, SBX_BASIS9(0xffffff, 0xf0f0f0, 0xd9d9d9, 0xbdbdbd, 0x969696, 0x737373, 0x525252, 0x252525, 0x000000)},
    {"oranges", "ORANGES", false, SBX_BASIS9(0xfff5eb, 0xfee6ce, 0xfdd0a2, 0xfdae6b, 0xfd8d3c, 0xf16913, 0xd94801, 0xa63603, 0x7f2704)},
};
#undef SBX_BASIS9
#undef SBX_BASIS11

inline const Def *find(const char *name)
{
    for (const Def &d : kDefs)
        if (!std::strcmp(name, d.name)) return &d;
    return nullptr;
}

// table[i] = gradient((i + 0.5) / entries): the value at the middle of the step a table of `entries` steps
// replaces the continuous gradient with (for the 256-entry ramps and entries = 256 exactly the ramp)
inline void table(const Def &d, uint32_t entries, uint8_t *rgb)
{
    for (uint32_t i = 0; i < entries; ++i) eval(d, ((double)i + 0.5) / (double)entries, rgb + 3 * (size_t)i);
}

}  // namespace gradients
}  // namespace sbx
