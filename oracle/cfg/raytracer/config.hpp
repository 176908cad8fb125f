#pragma once
// Build-time config shim for oracle/_ref (test infrastructure, not product code).
//
// The reference takes all render knobs from <raytracer/config.hpp> as constexpr values
// (/root/reference/include/raytracer/config.hpp:6-17; render.hpp includes it with <> at :5).  Putting this
// directory BEFORE the reference include dir on the command line lets one unmodified reference tree be compiled
// once per configuration: -DRT_CFG_SPP=.. -DRT_CFG_MAX_DEPTH=.. -DRT_CFG_GI_RAYS=..   Defaults = the reference's.

#include <cstddef>
#include <optional>

#ifndef RT_CFG_SPP
#define RT_CFG_SPP 1
#endif
#ifndef RT_CFG_MAX_DEPTH
#define RT_CFG_MAX_DEPTH 5
#endif
#ifndef RT_CFG_GI_RAYS
#define RT_CFG_GI_RAYS 0
#endif

constexpr double fov_degrees = 90.;

constexpr double epsilon = 1e-6;
constexpr double shadow_bias = 1e-4;
constexpr double reflection_bias = 1e-4;
constexpr double refraction_bias = 1e-4;

constexpr std::size_t samples_per_pixel = RT_CFG_SPP;
constexpr std::size_t max_ray_depth = RT_CFG_MAX_DEPTH;
constexpr std::size_t diffuse_reflection_ray_count = RT_CFG_GI_RAYS;

constexpr std::optional fixed_rng_seed = std::make_optional(42);
