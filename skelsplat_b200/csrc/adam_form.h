// The second-moment update of torch.optim.Adam's foreach path, exp_avg_sq.mul_(beta2).addcmul_(g, g, value = 1 - beta2), as ONE
// expression whose rounding matches torch's CUDA kernel bit for bit (tests/test_gpu_losses.py::test_adam_kernel_is_bit_exact...
// compares 125 steps against torch.optim.Adam on the GPU).  _foreach_addcmul_ evaluates  self + value * (t1 * t2)  and nvcc
// contracts the outer multiply-add:  fma(value, g * g, self).   SSB_ADAM_SQ_FIRST=0 selects the other association,
// fma(value * g, g, self), which differs in the last bit for ~1/3 of the inputs (kept for the A/B test only).
#pragma once
#ifndef SSB_ADAM_SQ_FIRST
#define SSB_ADAM_SQ_FIRST 1
#endif
#if SSB_ADAM_SQ_FIRST
#define SSB_ADAM_SECOND_MOMENT(omb2, g, decayed) fmaf((omb2), __fmul_rn((g), (g)), (decayed))
#else
#define SSB_ADAM_SECOND_MOMENT(omb2, g, decayed) fmaf(__fmul_rn((omb2), (g)), (g), (decayed))
#endif
