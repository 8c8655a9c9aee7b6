// devbuf.hpp -- tiny helpers for device memory and error checking (host side of the CUDA backend).
#pragma once
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>

#define MPB_CUDA_OK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
	fprintf(stderr, "[miniprot_b200] CUDA error %s at %s:%d: %s\n", cudaGetErrorName(e_), __FILE__, __LINE__, cudaGetErrorString(e_)); abort(); } } while (0)

namespace mpb {
namespace cuda {

// Grow-only device buffer: the arenas of a context live as long as the context, so steady-state batches
// allocate nothing (cudaMalloc is a device-wide synchronisation point).
struct DevBuf {
	void *p = 0;
	size_t cap = 0;
	void reserve(size_t bytes)
	{
		if (bytes <= cap) return;
		if (p) MPB_CUDA_OK(cudaFree(p));
		cap = bytes + bytes / 4 + 4096;
		MPB_CUDA_OK(cudaMalloc(&p, cap));
	}
	void release() { if (p) cudaFree(p); p = 0, cap = 0; }
	template <class T> T *as() const { return (T*)p; }
};

// Pinned host staging buffer (grow-only).
struct PinBuf {
	void *p = 0;
	size_t cap = 0;
	void reserve(size_t bytes)
	{
		if (bytes <= cap) return;
		if (p) MPB_CUDA_OK(cudaFreeHost(p));
		cap = bytes + bytes / 4 + 4096;
		MPB_CUDA_OK(cudaMallocHost(&p, cap));
	}
	void release() { if (p) cudaFreeHost(p); p = 0, cap = 0; }
	template <class T> T *as() const { return (T*)p; }
};

} // namespace cuda
} // namespace mpb
