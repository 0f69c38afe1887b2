"""ssmtoybox_b200 -- B200-native (sm_100a) Monte-Carlo sigma-point / Bayesian-quadrature Kalman
filtering behind the SSMToybox Python API.  CUDA-only: importing the compute modules requires the
built C-ABI library (python -m ssmtoybox_b200.build)."""
__version__ = '0.1.0'
