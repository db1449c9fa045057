"""horovod_b200.mxnet — not available.

The reference ships a mxnet binding (horovod/mxnet); this build targets PyTorch on B200 only and mxnet is not installed in
the build image, so there is nothing to bind against. The native runtime is framework-neutral (csrc/common/engine.h takes
raw device pointers + CUDA events): a mxnet adapter would mirror csrc/torch/binding.cc."""
raise ImportError('horovod_b200.mxnet is not built: only the PyTorch binding (horovod_b200.torch) exists in this build')
