#define OZL_F ozl::Fp2<ozl_params::Bls12381Fq>
#define OZL_C ozl_params::Bls12381G2
#define OZL_BASE ozl_params::Bls12381Fq
#define OZL_OPS ozl_ops_bls12_381_g2
#include "curve_inst.cuh"
