"""magprop_b200 -- B200-native likelihood hot path of sgibson91/magprop.

    from magprop_b200 import magnetar            # drop-in for the reference's `magnetar` package
    from magprop_b200.synthetic import funcs, mcmc_eqns   # drop-in for code/synthetic_datasets/*
    from magprop_b200.engine import Likelihood   # batched handle over the C ABI

All arithmetic runs in magprop_b200/libmagprop_b200.so (CUDA, sm_100a); there
is no CPU fallback.
"""
__version__ = "0.1.0"
