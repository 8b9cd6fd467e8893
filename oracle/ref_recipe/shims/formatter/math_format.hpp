// Stand-in for engine/modules/Formatter/include/formatter/math_format.hpp: Render/src/render.cpp includes it
// but formats nothing on the interpolateFrame path. TEST INFRASTRUCTURE ONLY.
#pragma once
#include <algorithm>
#include <chrono>
