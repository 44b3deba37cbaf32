// tests/simt_emu/cuda_pipeline.h -- see cuda_runtime.h in this directory.  cp.async modelled as an immediate copy.
#pragma once
#include <cstring>
#include <cstddef>
static inline void __pipeline_memcpy_async(void* dst, const void* src, size_t n) { std::memcpy(dst, src, n); }
static inline void __pipeline_commit() {}
static inline void __pipeline_wait_prior(int) {}
