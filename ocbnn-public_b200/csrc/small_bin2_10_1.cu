// Instantiation of the register-resident kernels for the [2,10,1] rbf net (repro toy shape).
#include "small_kernels.cuh"
OCBNN_DEFINE_SMALL_NET(net_bin2_10_1, 2, 10, 1, OCBNN_ACT_RBF)
