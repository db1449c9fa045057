"""horovod_b200.tensorflow — not available.

The reference ships a tensorflow binding (horovod/tensorflow); this build targets PyTorch on B200 only and tensorflow is not installed in
the build image, so there is nothing to bind against. The native runtime is framework-neutral (csrc/common/engine.h takes
raw device pointers + CUDA events): a tensorflow adapter would mirror csrc/torch/binding.cc."""
raise ImportError('horovod_b200.tensorflow is not built: only the PyTorch binding (horovod_b200.torch) exists in this build')
