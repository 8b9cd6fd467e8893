// Stand-in for the reference's Async module header (engine/modules/Async/include/async/async.hpp:46-72),
// which is built on asio strands. Only AsyncContext is needed by the ECS sources; on this single-threaded
// path "being on the strand" is always true, so ensureOnStrand() is an already-finished awaitable
// (the reference returns noop() in that case, async.hpp:66-70). TEST INFRASTRUCTURE ONLY.
#pragma once
#include <chrono>
#include <stdexcept>
#include <utility>
#include <variant>
#include <asio.hpp>

namespace astre::async
{
    inline asio::awaitable<void> noop() { co_return; }

    template<class AsioContext>
    class AsyncContext
    {
    public:
        explicit AsyncContext(AsioContext &) {}
        AsyncContext(const AsyncContext &) = delete;
        AsyncContext & operator=(const AsyncContext &) = delete;
        AsyncContext(AsyncContext &&) noexcept = default;
        AsyncContext & operator=(AsyncContext &&) noexcept = default;

        asio::awaitable<void> ensureOnStrand() const { return noop(); }
    };
}
