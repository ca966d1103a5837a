// shim for MSVC PPL <concurrent_unordered_map.h>: std::map guarded by a mutex.
// std::map iterators stay valid across inserts, which is all the reference relies on.
#pragma once
#include <map>
#include <mutex>
namespace concurrency
{
template <typename K, typename V>
class concurrent_unordered_map
{
	typedef std::map<K, V> Map;
public:
	typedef typename Map::value_type value_type;
	typedef typename Map::iterator iterator;
	typedef typename Map::const_iterator const_iterator;

	iterator find(const K& k) { std::lock_guard<std::mutex> g(m_Lock); return m_Map.find(k); }
	iterator end() { std::lock_guard<std::mutex> g(m_Lock); return m_Map.end(); }
	const_iterator cbegin() const { return m_Map.cbegin(); }
	const_iterator cend() const { return m_Map.cend(); }
	std::pair<iterator, bool> insert(const value_type& v) { std::lock_guard<std::mutex> g(m_Lock); return m_Map.insert(v); }
	void clear() { std::lock_guard<std::mutex> g(m_Lock); m_Map.clear(); }
private:
	Map m_Map;
	std::mutex m_Lock;
};
}
