// Test-infrastructure shim: lets the UNMODIFIED reference sources (MSVC/Win32 dialect)
// compile with g++ on Linux.  Nothing here is reference code.
#pragma once
#include <cstdlib>
#include <cstring>
#include <cstdio>
#include <cfloat>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <limits>
#include <atomic>
#include <memory>
#include <vector>
#include <algorithm>
#include <functional>
#include <utility>

#ifndef _countof
#define _countof(a) (sizeof(a) / sizeof((a)[0]))
#endif

// __declspec(x) is mapped to __VX_DS_##x on the command line.
#define __VX_DS_thread thread_local
#define __VX_DS_dllexport
#define __VX_DS_dllimport

inline void* _aligned_malloc(size_t size, size_t alignment)
{
	void* p = nullptr;
	if (alignment < sizeof(void*)) alignment = sizeof(void*);
	if (posix_memalign(&p, alignment, size ? size : 1) != 0) return nullptr;
	return p;
}
inline void _aligned_free(void* p) { free(p); }

inline int GetCurrentThreadId()
{
	static std::atomic<int> next(1);
	static thread_local int mine = 0;
	if (!mine) mine = next.fetch_add(1);
	return mine;
}
