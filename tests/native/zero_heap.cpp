// TEST INFRASTRUCTURE: every C++ allocation zero-filled.  Linked only into the reference-side build of tests/native/
// client_presets.cpp so that the reference's reads of members it never initialises (Neuron::scheduledFireTime, NeuCor.h:241;
// see oracle/ref_harness.cpp for the details) do not depend on what the allocator hands back.
#include <cstdlib>
#include <new>
void* operator new(std::size_t n) {
    void* p = calloc(1, n ? n : 1);
    if (!p) throw std::bad_alloc();
    return p;
}
void* operator new[](std::size_t n) { return operator new(n); }
void operator delete(void* p) noexcept { free(p); }
void operator delete[](void* p) noexcept { free(p); }
void operator delete(void* p, std::size_t) noexcept { free(p); }
void operator delete[](void* p, std::size_t) noexcept { free(p); }
