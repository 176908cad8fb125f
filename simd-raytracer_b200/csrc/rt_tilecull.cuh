// csrc/rt_tilecull.cuh - can any camera ray of a raster rectangle reach an axis-aligned box?  (k_tile_cull, rt_stream.cuh)
//
// The reference builds a camera ray from the raster position (x, y) as normalized(transpose(M) * (sx, sy, -1)) with
// sx = (2 x / W - 1) * aspect * tan(fov / 2), sy = (1 - 2 y / H) * tan(fov / 2)  (render/render.hpp:47-60): before the
// normalisation the direction is AFFINE in (x, y).  Every ray through the rectangle [x0, x1] x [y0, y1] - any sample position
// inside its pixels - therefore lies in the pyramid spanned by the directions through the rectangle's four corners, apex at
// the camera.  If the box lies outside one of the pyramid's four side planes, no ray of the rectangle enters it.
//
// The test is conservative: a tile is reported as missing only if the box is outside a side plane by more than MARGIN times
// the camera-to-box reach (L1 norm) - 1e-3, about half a pixel at 1080p, three orders of magnitude above the rounding of this
// arithmetic and of the reference's own slab test (core/math/aabb3.hpp:74-90), which is the first thing the reference does
// with such a ray.  Tiles that graze the box are traced as before.  A NaN anywhere keeps the tile.
//
// Compiles as CUDA device code and as plain C++ (tests/helpers/kd8_host.cpp checks it on the CPU against exact ray-box tests).
#pragma once

#include <cmath>

#if defined(__CUDACC__)
#define RT_TC_HD __host__ __device__ __forceinline__
#else
#define RT_TC_HD inline
#endif

namespace rtb {

constexpr float TILE_CULL_MARGIN = 1e-3f;

struct TileCamera {
    float m[9];                 // camera matrix, row-major as the scene file has it (the ray uses its transpose)
    float pos[3];
    float width, height;        // image size in pixels
    float tan_half_fov;
};

// un-normalised direction of the camera ray through raster position (x, y)
RT_TC_HD void tile_corner_dir(const TileCamera& c, float x, float y, float d[3]) {
    const float aspect = c.width / c.height;
    const float sx = ((2.0f * (x / c.width)) - 1.0f) * aspect * c.tan_half_fov;
    const float sy = (1.0f - (2.0f * (y / c.height))) * c.tan_half_fov;
    d[0] = c.m[0] * sx + c.m[3] * sy - c.m[6];
    d[1] = c.m[1] * sx + c.m[4] * sy - c.m[7];
    d[2] = c.m[2] * sx + c.m[5] * sy - c.m[8];
}

// true: no ray through the raster rectangle [x0, x1] x [y0, y1] can reach the box [lo, hi]
RT_TC_HD bool tile_misses_box(const TileCamera& c, const float box_lo[3], const float box_hi[3], float x0, float y0, float x1, float y1) {
    float d[4][3];
    tile_corner_dir(c, x0, y0, d[0]); tile_corner_dir(c, x1, y0, d[1]); tile_corner_dir(c, x1, y1, d[2]); tile_corner_dir(c, x0, y1, d[3]);
    const float centre[3] = {(d[0][0] + d[1][0]) + (d[2][0] + d[3][0]), (d[0][1] + d[1][1]) + (d[2][1] + d[3][1]),
                             (d[0][2] + d[1][2]) + (d[2][2] + d[3][2])};
    const float lo[3] = {box_lo[0] - c.pos[0], box_lo[1] - c.pos[1], box_lo[2] - c.pos[2]};
    const float hi[3] = {box_hi[0] - c.pos[0], box_hi[1] - c.pos[1], box_hi[2] - c.pos[2]};
    const float reach = fmaxf(fabsf(lo[0]), fabsf(hi[0])) + fmaxf(fabsf(lo[1]), fabsf(hi[1])) + fmaxf(fabsf(lo[2]), fabsf(hi[2]));
    bool outside = false;
    for (int e = 0; e < 4; ++e) {
        const float* a = d[e];
        const float* b = d[(e + 1) & 3];
        float n[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
        if (n[0] * centre[0] + n[1] * centre[1] + n[2] * centre[2] > 0.0f) { n[0] = -n[0]; n[1] = -n[1]; n[2] = -n[2]; }   // pyramid: n . p <= 0
        // the box corner that is deepest on the pyramid's side of the plane
        const float deepest = n[0] * (n[0] > 0.0f ? lo[0] : hi[0]) + n[1] * (n[1] > 0.0f ? lo[1] : hi[1]) + n[2] * (n[2] > 0.0f ? lo[2] : hi[2]);
        outside = outside || (deepest > TILE_CULL_MARGIN * reach * sqrtf(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]));
    }
    return outside;
}

}  // namespace rtb
