// Stand-in for <asio.hpp> (asio 1.34.2 is fetched at configure time by the reference's build --
// cmake/dependencies.cmake -- and is not in this image). TEST INFRASTRUCTURE ONLY: it exists so that the
// reference's own ECS / Render sources compile unmodified into oracle/_ref (see ../README.md).
//
// Only what those sources name is provided: asio::awaitable<T> (a coroutine return type) and the two
// execution-context type names. Nothing on this path ever suspends -- every awaited operation in the sources
// we compile is an already-finished awaitable -- so the coroutine starts eagerly, runs to its co_return on
// the caller's stack, and `co_await` on it just hands over the stored result.
#pragma once
#include <coroutine>
#include <exception>
#include <optional>
#include <utility>

namespace asio
{
    class io_context {};
    class thread_pool {};

    template<class T> class awaitable;

    namespace detail
    {
        template<class T>
        struct promise_result
        {
            std::optional<T> value;
            template<class U> void return_value(U && v) { value.emplace(std::forward<U>(v)); }
            T take() { return std::move(*value); }
        };
        template<>
        struct promise_result<void>
        {
            void return_void() noexcept {}
            void take() noexcept {}
        };
    }

    template<class T = void>
    class awaitable
    {
    public:
        struct promise_type : detail::promise_result<T>
        {
            std::exception_ptr error;
            awaitable get_return_object() { return awaitable(std::coroutine_handle<promise_type>::from_promise(*this)); }
            std::suspend_never initial_suspend() noexcept { return {}; }
            std::suspend_always final_suspend() noexcept { return {}; }   // frame lives until the awaitable dies
            void unhandled_exception() { error = std::current_exception(); }
        };

        awaitable(awaitable && o) noexcept : _h(std::exchange(o._h, {})) {}
        awaitable & operator=(awaitable && o) noexcept
        {
            if (this != &o) { reset(); _h = std::exchange(o._h, {}); }
            return *this;
        }
        awaitable(const awaitable &) = delete;
        awaitable & operator=(const awaitable &) = delete;
        ~awaitable() { reset(); }

        // awaiter interface: the coroutine has already run to completion
        bool await_ready() const noexcept { return true; }
        void await_suspend(std::coroutine_handle<>) const noexcept {}
        T await_resume() { return get(); }

        // used by the driver in place of co_spawn + a blocking wait
        T get()
        {
            if (_h.promise().error) std::rethrow_exception(_h.promise().error);
            return _h.promise().take();
        }

    private:
        explicit awaitable(std::coroutine_handle<promise_type> h) : _h(h) {}
        void reset() { if (_h) { _h.destroy(); _h = {}; } }
        std::coroutine_handle<promise_type> _h;
    };
}
