#define OZL_F ozl::Fp<ozl_params::Bn254Fq>
#define OZL_C ozl_params::Bn254G1
#define OZL_BASE ozl_params::Bn254Fq
#define OZL_OPS ozl_ops_bn254_g1
#include "curve_inst.cuh"
