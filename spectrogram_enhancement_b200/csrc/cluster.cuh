// Thread-block cluster helpers (barrier, rank, distributed shared memory) with their stand-ins in the emulation build.
#pragma once
#include "common.cuh"

#ifndef SPECGPU_EMULATE
#include <cooperative_groups.h>
#define SPECGPU_CLUSTER_SYNC() cooperative_groups::this_cluster().sync()
#define SPECGPU_CLUSTER_RANK() ((int)cooperative_groups::this_cluster().block_rank())
#define SPECGPU_MAP_SHARED(p, r) cooperative_groups::this_cluster().map_shared_rank((p), (r))
#else
#define SPECGPU_CLUSTER_SYNC() emu::cluster_sync()
#define SPECGPU_CLUSTER_RANK() ((int)emu::cluster_ctarank())
#define SPECGPU_MAP_SHARED(p, r) emu::map_shared_rank((p), (r))
#endif
