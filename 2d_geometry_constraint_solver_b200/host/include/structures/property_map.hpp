// MathUtils::PropertyMap: id -> value, iterated in ascending id order (reference:
// src/structures/include/structures/property_map.hpp:33-180, a std::flat_map there; the sorted
// vector of simple_graph.hpp here - g++ 13 has no <flat_map>).
#pragma once

#include <cstddef>
#include <expected>
#include <functional>
#include <utility>

#include <structures/simple_graph.hpp>

namespace MathUtils {

enum class PropertyMapError { KeyNotFound };

template <typename KeyId, typename Value>
class PropertyMap {
public:
    using KeyType = KeyId;
    using ValueType = Value;

    Value& operator[](const KeyId& id)
    {
        auto it = m_data.find(id);
        return it != m_data.end() ? it->second : m_data.put(id, Value {}).second;
    }
    std::expected<std::reference_wrapper<const Value>, PropertyMapError> get(const KeyId& id) const
    {
        const auto it = m_data.find(id);
        if (it == m_data.end()) return std::unexpected(PropertyMapError::KeyNotFound);
        return std::cref(it->second);
    }
    std::expected<std::reference_wrapper<Value>, PropertyMapError> get(const KeyId& id)
    {
        auto it = m_data.find(id);
        if (it == m_data.end()) return std::unexpected(PropertyMapError::KeyNotFound);
        return std::ref(it->second);
    }
    void set(const KeyId& id, Value value) { m_data.put(id, std::move(value)); }
    std::expected<void, PropertyMapError> erase(const KeyId& id)
    {
        if (!m_data.erase(id)) return std::unexpected(PropertyMapError::KeyNotFound);
        return {};
    }
    void clear() { m_data.clear(); }
    void reserve(std::size_t n) { m_data.reserve(n); }
    bool contains(const KeyId& id) const { return m_data.contains(id); }
    std::size_t size() const { return m_data.size(); }
    bool empty() const { return m_data.empty(); }
    auto begin() { return m_data.begin(); }
    auto end() { return m_data.end(); }
    auto begin() const { return m_data.begin(); }
    auto end() const { return m_data.end(); }

private:
    detail::FlatTable<KeyId, Value> m_data;  // sorted by id, like the reference's std::flat_map
};

template <typename Value>
using NodePropertyMap = PropertyMap<NodeId, Value>;
template <typename Value>
using EdgePropertyMap = PropertyMap<EdgeId, Value>;

}  // namespace MathUtils
