// Instantiation of the register-resident kernels for the [2,10,3] rbf net (repro toy shape).
#include "small_kernels.cuh"
OCBNN_DEFINE_SMALL_NET(net_cls2_10_3, 2, 10, 3, OCBNN_ACT_RBF)
