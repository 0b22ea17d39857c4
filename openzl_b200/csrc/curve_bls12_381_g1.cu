#define OZL_F ozl::Fp<ozl_params::Bls12381Fq>
#define OZL_C ozl_params::Bls12381G1
#define OZL_BASE ozl_params::Bls12381Fq
#define OZL_OPS ozl_ops_bls12_381_g1
#define OZL_FP64_BENCH 1
#include "curve_inst.cuh"
