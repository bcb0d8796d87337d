// shard.cu -- single-field transform split over the GPUs of one box.  See below.
#include "s2k_internal.cuh"

static int not_yet(const char* what) {
    (void)what;
    return 3;
}

extern "C" int s2kit_cuda_plan_create_sharded(s2kit_cuda_plan** out, int, int, int, int, int) {
    if (out) *out = nullptr;
    return not_yet("plan_create_sharded");
}
extern "C" int s2kit_cuda_fst_rings(s2kit_cuda_plan*, const double*, const double*, double*) { return not_yet("fst_rings"); }
extern "C" int s2kit_cuda_fst_orders(s2kit_cuda_plan*, const double*, double*, double*) { return not_yet("fst_orders"); }
extern "C" int s2kit_cuda_inv_fst_orders(s2kit_cuda_plan*, const double*, const double*, double*) { return not_yet("inv_fst_orders"); }
extern "C" int s2kit_cuda_inv_fst_rings(s2kit_cuda_plan*, const double*, double*, double*) { return not_yet("inv_fst_rings"); }
extern "C" int s2kit_cuda_shard_info(const s2kit_cuda_plan*, long*, int*, int*) { return not_yet("shard_info"); }
