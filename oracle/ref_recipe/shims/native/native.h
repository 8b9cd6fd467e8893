// Stand-in for engine/modules/Native/include/native/native.h (platform typedefs; nothing on this path uses them).
#pragma once
