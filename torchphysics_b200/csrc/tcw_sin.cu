// The wide tensor-core kernels for tp.models.Sinus (reference models/activation_fn.py:76-85): tcw_kernels.cu compiled
// with the sin activation (namespace tcws, entry points tcws_forward / tcws_backward).
#define TCW_ACT_SIN 1
#include "tcw_kernels.cu"
