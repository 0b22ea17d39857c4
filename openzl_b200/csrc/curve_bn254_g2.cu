#define OZL_F ozl::Fp2<ozl_params::Bn254Fq>
#define OZL_C ozl_params::Bn254G2
#define OZL_BASE ozl_params::Bn254Fq
#define OZL_OPS ozl_ops_bn254_g2
#include "curve_inst.cuh"
