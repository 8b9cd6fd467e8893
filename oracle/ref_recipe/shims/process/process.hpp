// Stand-in for engine/modules/Process/include/process/process.hpp (Win32-only module). The ECS Registry only
// needs IProcess::execution_context_type and getExecutionContext() (ECS/src/registry.cpp:7-11).
// TEST INFRASTRUCTURE ONLY.
#pragma once
#include <asio.hpp>
#include "async/async.hpp"
#include "type/type.hpp"

namespace astre::process
{
    class IProcess
    {
    public:
        using execution_context_type = asio::thread_pool;
        execution_context_type & getExecutionContext() { return _pool; }
    private:
        execution_context_type _pool;
    };
}
