// Shared plumbing for the smgx library: error type, CUDA checks, small device-buffer helper.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <string>

#include "../../include/smgx.h"

namespace smgx {

struct Error : std::runtime_error {
    smgx_status code;
    Error(smgx_status c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define SMGX_CUDA(expr)                                                                                        \
    do {                                                                                                       \
        cudaError_t _e = (expr);                                                                               \
        if (_e != cudaSuccess)                                                                                 \
            throw ::smgx::Error(SMGX_DEVICE_ERROR, std::string(#expr) + ": " + cudaGetErrorString(_e) +         \
                                                       " (" __FILE__ ":" + std::to_string(__LINE__) + ")");    \
    } while (0)

#define SMGX_REQUIRE(cond, msg)                                              \
    do {                                                                     \
        if (!(cond)) throw ::smgx::Error(SMGX_INVALID_ARGUMENT, (msg));      \
    } while (0)

// Growable device allocation (never shrinks); contents are NOT preserved on growth.
struct DevBuf {
    void* ptr = nullptr;
    size_t cap = 0;
    void reserve(size_t bytes) {
        if (bytes <= cap) return;
        if (ptr) cudaFree(ptr);
        ptr = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        SMGX_CUDA(cudaMalloc(&ptr, want));
        cap = want;
    }
    void release() {
        if (ptr) cudaFree(ptr);
        ptr = nullptr;
        cap = 0;
    }
    template <class T> T* as() const { return reinterpret_cast<T*>(ptr); }
};

// Growable pinned host allocation.
struct PinBuf {
    void* ptr = nullptr;
    size_t cap = 0;
    void reserve(size_t bytes) {
        if (bytes <= cap) return;
        if (ptr) cudaFreeHost(ptr);
        ptr = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        SMGX_CUDA(cudaMallocHost(&ptr, want));
        cap = want;
    }
    void release() {
        if (ptr) cudaFreeHost(ptr);
        ptr = nullptr;
        cap = 0;
    }
    template <class T> T* as() const { return reinterpret_cast<T*>(ptr); }
};

}  // namespace smgx
