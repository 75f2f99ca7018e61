"""Gaussian AOCP learning loop (bnn/base.py:164-188).  The objective + gradient kernel (K6) is not built yet."""


def learn_gaussian_aocp(model, verbose=True):
    raise NotImplementedError("learn_gaussian_aocp: the AOCP objective kernel (K6) is scheduled after the sampler path; "
                              "use load_gaussian_aocp_parameters() with learnt parameters")
