// oracle/ref_shims/core/utils/memory_utils.h — TEST INFRASTRUCTURE ONLY (see core/xmipp_error.h in this directory).
// Stand-in for xmippCore's memoryUtils::page_aligned_alloc as used by reconstruct_fourier_buffer_data.h:55-69:
// page-aligned host allocation released with free(), optionally zeroed.
#pragma once
#include <cstdlib>
#include <cstring>
#include <unistd.h>
namespace memoryUtils {
template <typename T>
T* page_aligned_alloc(size_t elems, bool initToZero) {
    const size_t page = (size_t)sysconf(_SC_PAGESIZE);
    size_t bytes = ((elems * sizeof(T) + page - 1) / page) * page;
    if (bytes == 0) bytes = page;
    void* p = aligned_alloc(page, bytes);
    if (p && initToZero) memset(p, 0, bytes);
    return (T*)p;
}
}  // namespace memoryUtils
