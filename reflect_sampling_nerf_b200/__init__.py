"""reflect_sampling_nerf_b200 -- B200-native (sm_100a) implementation of the per-ray rendering hot path of
the nerfstudio method `reflect-sampling-nerf`.  Host code is Python/PyTorch (device memory, streams,
torch.distributed); every stage of the path is a hand-written CUDA kernel behind the C-ABI declared in
include/rsn_b200.h.  See DESIGN.md."""
__version__ = "0.1.0"
