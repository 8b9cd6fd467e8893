// Stand-in for abseil's flat_hash_map (abseil 20250127.1 is fetched by the reference's build, absent here).
// Iteration order differs from abseil's (which is salted per process anyway -- SURVEY.md section 0), so the
// driver never depends on it: results are read back per entity id. TEST INFRASTRUCTURE ONLY.
#pragma once
#include <cassert>
#include <unordered_map>
#include <typeindex>
#include <string>
namespace absl
{
    template<class K, class V> using flat_hash_map = std::unordered_map<K, V>;
}
