// Stand-in for abseil's flat_hash_set; see flat_hash_map.h. TEST INFRASTRUCTURE ONLY.
#pragma once
#include <unordered_set>
namespace absl
{
    template<class K> using flat_hash_set = std::unordered_set<K>;
}
