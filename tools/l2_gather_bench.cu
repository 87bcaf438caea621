// l2_gather_bench.cu - measured ceiling of the access pattern K23 (wsae_decode_backward.cu) is bound by:
// warps gathering whole rows of an L2-resident bf16 table at random row indices (k = 32 rows per
// activation row, d * 2 bytes each), every byte consumed exactly once.  No arithmetic besides an
// XOR that keeps the loads alive.  Sweeps bytes per lane per load (8 = K23's current form, 16) and
// resident warps per SM, prints one JSON object; `ceiling_gbs` is the best configuration.
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/l2_gather_bench tools/l2_gather_bench.cu
//   ./tools/l2_gather_bench [F=3072] [d=384] [rows=75776] > profiles/r2_l2_gather.json
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

template <int VEC>   // 8 or 16 bytes per lane per load
__global__ void gather_kernel(const uint8_t* __restrict__ table, const int* __restrict__ idx, int rows,
                              int k, int row_bytes, unsigned* __restrict__ sink) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  unsigned acc = 0;
  for (int r = warp; r < rows; r += nwarps) {
    const int my = idx[static_cast<size_t>(r) * k + (lane % k)];
    for (int c0 = 0; c0 < row_bytes; c0 += 32 * VEC) {
      const int off = c0 + lane * VEC;
      if (VEC == 8) {
        uint2 v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int f = __shfl_sync(0xffffffffu, my, j);
          v[j] = make_uint2(0u, 0u);
          if (off < row_bytes) v[j] = __ldg(reinterpret_cast<const uint2*>(table + static_cast<size_t>(f) * row_bytes + off));
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) acc ^= v[j].x ^ v[j].y;
      } else {
        uint4 v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int f = __shfl_sync(0xffffffffu, my, j);
          v[j] = make_uint4(0u, 0u, 0u, 0u);
          if (off < row_bytes) v[j] = __ldg(reinterpret_cast<const uint4*>(table + static_cast<size_t>(f) * row_bytes + off));
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) acc ^= v[j].x ^ v[j].y ^ v[j].z ^ v[j].w;
      }
    }
  }
  if (acc == 0x12345678u) sink[0] = acc;   // never true in practice; keeps the loads observable
}

int main(int argc, char** argv) {
  const int F = argc > 1 ? atoi(argv[1]) : 3072;
  const int d = argc > 2 ? atoi(argv[2]) : 384;
  const int rows = argc > 3 ? atoi(argv[3]) : 75776;
  const int k = 32, row_bytes = d * 2;
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  std::vector<int> h_idx(static_cast<size_t>(rows) * k);
  uint64_t s = 88172645463325252ull;
  for (auto& v : h_idx) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; v = static_cast<int>(s % F); }
  uint8_t* table; int* idx; unsigned* sink;
  cudaMalloc(&table, static_cast<size_t>(F) * row_bytes);
  cudaMemset(table, 1, static_cast<size_t>(F) * row_bytes);
  cudaMalloc(&idx, h_idx.size() * sizeof(int));
  cudaMemcpy(idx, h_idx.data(), h_idx.size() * sizeof(int), cudaMemcpyHostToDevice);
  cudaMalloc(&sink, 4);
  const double gathered = static_cast<double>(rows) * k * row_bytes;
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  double best = 0; int best_vec = 0, best_wps = 0;
  printf("{\"table\": \"bf16 [%d, %d] = %.2f MB (L2 resident)\", \"rows\": %d, \"k\": %d, \"sms\": %d, \"sm_clock_khz\": %d,\n \"gathered_bytes\": %.0f, \"sweep\": [", F, d, F * row_bytes / 1e6, rows, k, sms, clk_khz, gathered);
  bool first = true;
  for (int vec : {8, 16}) {
    for (int wps : {8, 16, 24, 32, 48, 64}) {          // resident warps per SM (blocks of 128 threads)
      const int blocks = sms * wps / 4;
      auto launch = [&]() {
        if (vec == 8) gather_kernel<8><<<blocks, 128>>>(table, idx, rows, k, row_bytes, sink);
        else gather_kernel<16><<<blocks, 128>>>(table, idx, rows, k, row_bytes, sink);
      };
      for (int i = 0; i < 3; ++i) launch();
      cudaDeviceSynchronize();
      float best_ms = 1e9f;
      for (int rep = 0; rep < 10; ++rep) {
        cudaEventRecord(a);
        launch();
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (ms < best_ms) best_ms = ms;
      }
      if (cudaGetLastError() != cudaSuccess) { printf("]}\n"); return 1; }
      const double gbs = gathered / (best_ms * 1e-3) / 1e9;
      printf("%s\n  {\"bytes_per_lane\": %d, \"warps_per_sm\": %d, \"ms\": %.4f, \"gbs\": %.1f}", first ? "" : ",", vec, wps, best_ms, gbs);
      first = false;
      if (gbs > best) { best = gbs; best_vec = vec; best_wps = wps; }
    }
  }
  printf("],\n \"ceiling_gbs\": %.1f, \"ceiling_config\": {\"bytes_per_lane\": %d, \"warps_per_sm\": %d},\n"
         " \"note\": \"best of 10 launches per cell, CUDA events; registers limit the resident warps (ptxas), so high warps_per_sm cells may run in waves\"}\n",
         best, best_vec, best_wps);
  return 0;
}
