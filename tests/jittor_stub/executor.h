// Stub of Jittor's <executor.h> for tests/test_jittor_adapter.py: just enough of the JIT glue that a `jt.code`
// CUDA body sees (exe.allocator, LOGf) to type-check rs_detection_b200/jittor_adapter.py's sources with nvcc.
// TEST INFRASTRUCTURE ONLY -- the real header comes with Jittor.
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <sstream>
namespace jittor {
struct Allocator {
    void* alloc(size_t size, size_t& allocation) { void* p = nullptr; allocation = 0; cudaMalloc(&p, size); return p; }
    void free(void* p, size_t, const size_t&) { cudaFree(p); }
};
struct Executor { Allocator* allocator; };
extern Executor exe;
struct LogFatal {   // `LOGf << a << b;` throws in Jittor; here it aborts when the temporary dies
    std::ostringstream s;
    template <class T> LogFatal& operator<<(const T& v) { s << v << ' '; return *this; }
    ~LogFatal() { fprintf(stderr, "%s\n", s.str().c_str()); abort(); }
};
struct Var { size_t size; };
}
using namespace jittor;
#define LOGf jittor::LogFatal()
