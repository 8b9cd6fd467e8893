// Part of the GLM-API stand-in; see refglm.hpp. TEST INFRASTRUCTURE ONLY.
#pragma once
#include <glm/refglm.hpp>
