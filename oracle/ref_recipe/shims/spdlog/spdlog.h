// Stand-in for spdlog 1.17 (fetched by the reference's build, absent here): logging calls become no-ops.
// TEST INFRASTRUCTURE ONLY.
#pragma once
#include <cassert>
namespace spdlog
{
    template<class... A> inline void trace(A && ...) {}
    template<class... A> inline void debug(A && ...) {}
    template<class... A> inline void info(A && ...) {}
    template<class... A> inline void warn(A && ...) {}
    template<class... A> inline void error(A && ...) {}
    template<class... A> inline void critical(A && ...) {}
}
