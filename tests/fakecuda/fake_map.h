// what tests/fakecuda/fake_cuda.cpp stores inside the opaque 128-byte CUtensorMap: enough geometry for range checks and for the CPU
// model of the conv kernel (fake_kernels.cpp) to address the tensor like the TMA unit would
#pragma once
#include <stdint.h>
struct FakeMap {
    uint64_t magic;       // 0x54454e534d4150
    const void* base;
    uint64_t extent;      // bytes from base to the end of the tensor
    uint32_t elem_bytes, rank;
    uint64_t dims[5];     // dims[0] innermost
    uint64_t strides[4];  // bytes, for dims[1..]
};
static_assert(sizeof(FakeMap) <= 128, "must fit a CUtensorMap");
constexpr uint64_t FAKE_MAP_MAGIC = 0x54454e534d4150ull;
