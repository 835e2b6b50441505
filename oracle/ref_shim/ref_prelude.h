/* ref_prelude.h -- force-included before the reference's translation units (TEST INFRASTRUCTURE ONLY) */
#pragma once
#include <climits>
#include <cstdlib>
#include <cstring>
#include <cmath>
#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif
