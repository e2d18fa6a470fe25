// TMA (cp.async.bulk.tensor) support: tile-mode tensor maps over the chunk-planar f16 scene tensors
//   [outer][8 chunks][rows][cols][8 halves]
// encoded on the host through the driver entry point (no libcuda link), and the 4-D load that drops a
// (rows x 32 columns x 8 chunks) tile into shared memory as [chunk][row][col][16 B] -- exactly the UMMA
// no-swizzle K-major operand layout the scene kernels issue their tcgen05.mma on.  Out-of-range rows /
// columns (map borders, negative halo coordinates) are zero-filled by the hardware.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace cmlpl {

// tensor [outer][8][rows][cols][8 halves] f16 at `base`; box = (box_cols x 8 halves, box_rows, 8 chunks, 1)
int make_scene_tmap(CUtensorMap* out, const void* base, int outer, int rows, int cols, int box_rows, int box_cols);

// tile at (column x, row y) of chunk-plane group `outer`: coordinates (x*8 halves, y, chunk 0, outer)
__device__ __forceinline__ void tma_load_tile(uint32_t dst_smem, const CUtensorMap* tm, int x, int y, int outer, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
      ::"r"(dst_smem), "l"(tm), "r"(x * 8), "r"(y), "r"(0), "r"(outer), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tm) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tm) : "memory");
}

}  // namespace cmlpl
