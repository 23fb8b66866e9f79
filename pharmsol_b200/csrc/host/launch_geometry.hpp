// launch_geometry.hpp — grid / block shape of a psi launch for the one-thread-per-pair kernels, shared by the launcher
// (runtime.cu) and the host-compiled test double of the kernels (tests/hostsim), so both walk the same index space.
#pragma once
#include <algorithm>
#include <cstdint>

namespace pharmsol {

// Particle workspace of one SDE CTA, in doubles (psi_sde.cuh): two state buffers [nstate][np] that swap roles at every
// resampling — the weights and their running sum live in the first np slots of the idle one — and np 32-bit ancestors.
inline int64_t sde_workspace_doubles(int nstate, int np) {
    return (2LL * nstate * np + (np + 1) / 2 + 31) / 32 * 32;
}

struct LaunchGeometry {
    unsigned grid_x = 1, grid_y = 1, block = 128;
    int warp_tasks = 0;
};

// ncols support points x nsub subjects.  Normal case: blockIdx.x * block + threadIdx.x runs over columns, blockIdx.y
// strides over subjects.  Few support points (< 128): the warps of a 1-D grid take (subject, 32-column chunk) tasks.
inline LaunchGeometry psi_launch_geometry(int64_t nsub, int64_t ncols, bool diagonal, int sm_count, int block_threads) {
    LaunchGeometry g;
    g.block = (unsigned)block_threads;
    g.grid_x = (unsigned)((ncols + block_threads - 1) / block_threads);
    g.grid_y = diagonal ? 1u : (unsigned)std::min<int64_t>(nsub, 65535);
    if (!diagonal && ncols < 128) {
        g.warp_tasks = 1;
        const int64_t ntask = ((ncols + 31) / 32) * nsub, wpb = block_threads / 32;
        g.grid_x = (unsigned)std::max<int64_t>(1, std::min<int64_t>((ntask + wpb - 1) / wpb, (int64_t)sm_count * 64));
        g.grid_y = 1u;
    }
    return g;
}

}  // namespace pharmsol
