// Fused tile kernels: placeholder plan until the tiled path is built.
#pragma once
#include <cstddef>
namespace lfm {
struct TilePlan {
	bool ready = false;
	int n_tiles = 0, tile_cells = 0;
	size_t smem_bytes = 0;
	double halo_face_ratio = 0.0;
};
}  // namespace lfm
