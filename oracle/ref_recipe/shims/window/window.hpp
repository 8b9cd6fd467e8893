// Stand-in for engine/modules/Window/include/window/window.hpp (Win32-only). render.hpp only names
// window::IWindow in the declaration of createRenderer. TEST INFRASTRUCTURE ONLY.
#pragma once
namespace astre::window { class IWindow; }
