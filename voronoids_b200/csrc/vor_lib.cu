// vor_lib.cu -- the single translation unit of libvoronoids_b200.so (sm_100a).
// nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -lineinfo ... (see voronoids_b200/build.py)
#include "backend_cuda.cuh"
#include "engine.cuh"

namespace vor { namespace be { std::atomic<unsigned long long> g_launches{0}; Pool g_pool; HostPool g_hostpool; Staging g_staging_dev[MAX_DEVICES]; } }

#include "capi.inl"
