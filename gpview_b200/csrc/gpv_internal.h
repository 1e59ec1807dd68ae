// gpview_b200/csrc/gpv_internal.h -- shared between the host (.cpp) and device (.cu) halves of libgpview_b200.so
#pragma once
#include <string>

namespace gpv {
// records the message for gpv_last_error() (thread-local) and returns a non-zero error code
int fail(const std::string& msg);
}

#include <cstddef>
#include <cstdint>
namespace gpv {
// gpv_expand.cpp: 2-bit packed Level-2 words (uint2 per 32 sub-voxels: inside mask, boundary mask) -> file bytes 0 / 127 / 254
void expand_packed(const void* packed, uint8_t* out, size_t nWords);
struct ExpandPool;                                   // host threads that expand chunks while the next chunk is on the bus
ExpandPool* expand_pool_get();                       // process-wide, created on first use
int expand_pool_threads(ExpandPool* p);
void expand_pool_begin(ExpandPool* p);               // one call at a time owns the pool (blocks until the previous owner is done)
void expand_pool_submit(ExpandPool* p, const void* packed, uint8_t* out, size_t nWords);
void expand_pool_end(ExpandPool* p);                 // returns when everything submitted has been expanded
}
