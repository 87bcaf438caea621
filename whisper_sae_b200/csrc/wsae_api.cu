// wsae_api.cu — ABI identification for libwsae_sm100.so (see include/wsae.h).
extern "C" int wsae_abi_version(void) { return 111; }

#include <cuda_runtime.h>
// Launch an instantiated CUDA graph (cudaGraphExec_t) on `stream`: the trainer's captured step.  PyTorch's
// CUDAGraph.replay() spends ~25 us of host time per call around this one driver call.
extern "C" int wsae_graph_launch(void* graph_exec, void* stream) {
  if (!graph_exec) return -1;
  return static_cast<int>(cudaGraphLaunch(static_cast<cudaGraphExec_t>(graph_exec), static_cast<cudaStream_t>(stream)));
}

// cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, stream): the 64-byte control block of the captured step
// (pinned host -> device) without the ~6 us of tensor-copy dispatch in front of it.
extern "C" int wsae_memcpy_async(void* dst, const void* src, size_t bytes, void* stream) {
  if (!dst || !src) return -1;
  return static_cast<int>(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, static_cast<cudaStream_t>(stream)));
}
