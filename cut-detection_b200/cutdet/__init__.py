"""cutdet: the engine under the frameID mirror -- ctypes binding of libcutdet_b200.so, the frame
pipeline, multi-GPU sharding and the synthetic workload generator.  No CPU fallback lives here: every
numeric entry point raises if the CUDA library cannot be loaded."""
