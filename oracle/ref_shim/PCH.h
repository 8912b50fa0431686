// Shim standing in for terrainwatersim/source/PCH.h when compiling the reference's
// NoiseGenerator.cpp / Random.cpp unmodified (oracle/Makefile, target _ref).  Only the
// ezEngine math headers those two files use are pulled in.  TEST INFRASTRUCTURE.
#pragma once
#include <algorithm>
#include <cmath>
#include <Foundation/Basics.h>
#include <Foundation/Math/Vec2.h>
#include <Foundation/Math/Vec3.h>
