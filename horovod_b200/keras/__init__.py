"""horovod_b200.keras — not available.

The reference ships a keras binding (horovod/keras); this build targets PyTorch on B200 only and keras is not installed in
the build image, so there is nothing to bind against. The native runtime is framework-neutral (csrc/common/engine.h takes
raw device pointers + CUDA events): a keras adapter would mirror csrc/torch/binding.cc."""
raise ImportError('horovod_b200.keras is not built: only the PyTorch binding (horovod_b200.torch) exists in this build')
