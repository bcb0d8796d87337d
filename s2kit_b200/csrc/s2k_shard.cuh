// s2k_shard.cuh -- state of one rank's share of a single field split over several GPUs (shard.cu, multi.cu).
#pragma once
#include "s2k_internal.cuh"

struct ShardState {
    int nr = 0;           // rings per rank = rows per rank
    int nrows_real = 0;   // rows of this rank that exist (one less on the rank that owns order 0)
    long block = 0;       // doubles per (src, dst) block
    long* d_rowbase = nullptr;
    int* d_rowlist = nullptr;
    int* d_orders = nullptr;
    int norders = 0;
    s2k::PlaneView ring_view, order_view;
};


inline ShardState* shard_of(const s2kit_cuda_plan* p) { return reinterpret_cast<ShardState*>(p->shard); }
