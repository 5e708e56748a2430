/* Shim <volk/volk_alloc.hh>: volk::vector<T> = std::vector with 64-byte aligned storage.
 * TEST INFRASTRUCTURE ONLY. */
#ifndef ORACLE_SHIM_VOLK_ALLOC_HH
#define ORACLE_SHIM_VOLK_ALLOC_HH
#include <cstdlib>
#include <new>
#include <vector>
namespace volk {
template <class T>
struct alloc {
    typedef T value_type;
    alloc() = default;
    template <class U>
    constexpr alloc(alloc<U> const&) noexcept {}
    T* allocate(std::size_t n)
    {
        void* p = nullptr;
        const std::size_t bytes = n * sizeof(T);
        if (posix_memalign(&p, 64, bytes > 0 ? bytes : 64)) throw std::bad_alloc();
        return static_cast<T*>(p);
    }
    void deallocate(T* p, std::size_t) noexcept { free(p); }
};
template <class T, class U>
bool operator==(alloc<T> const&, alloc<U> const&) { return true; }
template <class T, class U>
bool operator!=(alloc<T> const&, alloc<U> const&) { return false; }
template <class T>
using vector = std::vector<T, alloc<T>>;
} // namespace volk
#endif
