// The reference's umbrella header ecs/ecs.hpp also pulls in ScriptSystem (Lua/sol2) and InputSystem (Win32 input),
// which are outside the path.  This includes the reference's own headers for the four systems on it.
// TEST INFRASTRUCTURE ONLY.
#pragma once
#include "ecs/registry.hpp"
#include "ecs/system/transform_system.hpp"
#include "ecs/system/camera_system.hpp"
#include "ecs/system/visual_system.hpp"
#include "ecs/system/light_system.hpp"
