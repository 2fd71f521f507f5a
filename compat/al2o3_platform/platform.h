// Compatibility shim for DeanoC/al2o3 `al2o3_platform/platform.h` (not vendored by the reference;
// its CMake fetches al2o3@master, which is unreachable offline). Only the macros the BCn encode
// path touches are provided. This file is OUR definition of those semantics.
#pragma once
#include <stdint.h>
#include <stdbool.h>
#include <stddef.h>
#include <string.h>
#include <stdlib.h>
#include <math.h>

#ifdef __cplusplus
#define AL2O3_EXTERN_C extern "C"
#else
#define AL2O3_EXTERN_C
#endif

#ifndef ASSERT
#define ASSERT(x) ((void)0)
#endif

#define AL2O3_DEFINE_ALIGNED(def, a) __attribute__((aligned(a))) def
#define AL2O3_FORCE_INLINE inline __attribute__((always_inline))
