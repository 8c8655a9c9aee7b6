// seg_sort.cu -- segmented ascending sort of 64-bit keys (anchors per protein / per window, sketch.c:95 and
// map.c:156,177 radix_sort_mp64: full-key sorts, so any correct sort gives the reference's result).
// This round it is served by CUB's DeviceSegmentedSort, a library primitive compiled for sm_100a; it is NOT
// one of the hand-written hot kernels and is listed as such in DESIGN.md.
#include <cub/device/device_segmented_sort.cuh>
#include "devbuf.hpp"
#include "seed_dev.hpp"

namespace mpb {
namespace cuda {

void seg_sort_u64(cudaStream_t st, uint64_t *keys, uint64_t *tmp, int64_t n_items, int n_seg, const int64_t *seg_begin, const int64_t *seg_end,
                  void **scratch, size_t *scratch_cap)
{
	if (n_items <= 0 || n_seg <= 0) return;
	cub::DoubleBuffer<uint64_t> db(keys, tmp);
	size_t need = 0;
	MPB_CUDA_OK(cub::DeviceSegmentedSort::SortKeys(0, need, db, n_items, n_seg, seg_begin, seg_end, st));
	if (need > *scratch_cap) {
		if (*scratch) MPB_CUDA_OK(cudaFree(*scratch));
		*scratch_cap = need + need / 4 + 4096;
		MPB_CUDA_OK(cudaMalloc(scratch, *scratch_cap));
	}
	MPB_CUDA_OK(cub::DeviceSegmentedSort::SortKeys(*scratch, need, db, n_items, n_seg, seg_begin, seg_end, st));
	if (db.Current() != keys) MPB_CUDA_OK(cudaMemcpyAsync(keys, db.Current(), sizeof(uint64_t) * (size_t)n_items, cudaMemcpyDeviceToDevice, st));
}

} // namespace cuda
} // namespace mpb
