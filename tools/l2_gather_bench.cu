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


// ---- cp.async (LDGSTS) staging variants: the same gather, L2 -> shared memory ring -> registers ----
// MODE 0: address of every copy computed from a shuffled offset right before it (few address
//         registers, rewritten for every copy);
// MODE 1: 8 row pointers per lane computed once per activation row, every copy addresses
//         [pointer + immediate] (slices unrolled), so no address register is rewritten while copies
//         that read it are in flight.
__device__ __forceinline__ void cp16(unsigned dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
template <int MODE, int NSL, int STAGES>
__global__ void __launch_bounds__(128, 2)
gather_cpasync_kernel(const uint8_t* __restrict__ table, const int* __restrict__ idx, int rows, int k,
                      unsigned* __restrict__ sink) {
  extern __shared__ __align__(16) uint8_t ring_raw[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int row_bytes = NSL * 256;
  const unsigned ring = static_cast<unsigned>(__cvta_generic_to_shared(ring_raw)) + wib * (STAGES * 8192);
  unsigned acc = 0;
  int stage = 0;
  // flattened (row, slice) sequence; prologue fills STAGES - 1 slices of the first row
  int r = warp;
  unsigned cur = r < rows ? static_cast<unsigned>(idx[static_cast<size_t>(r) * k + lane]) * row_bytes : 0u;
  unsigned nxt = r + nwarps < rows ? static_cast<unsigned>(idx[static_cast<size_t>(r + nwarps) * k + lane]) * row_bytes : 0u;
  auto issue0 = [&](unsigned offs, int sl, int st) {
    const unsigned dst0 = ring + st * 8192 + (lane >> 4) * 256 + (lane & 15) * 16;
    const uint8_t* src0 = table + sl * 256 + (lane & 15) * 16;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const unsigned off = __shfl_sync(0xffffffffu, offs, 2 * i + (lane >> 4));
      cp16(dst0 + i * 512, src0 + off);
    }
  };
  const uint8_t* pc[8];
  const uint8_t* pn[8];
  auto make_ptrs = [&](unsigned offs, const uint8_t* (&p)[8]) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      p[i] = table + __shfl_sync(0xffffffffu, offs, 4 * i + (lane >> 3)) + (lane & 7) * 16;
  };
  auto issue1 = [&](const uint8_t* (&p)[8], int sl_imm, int st) {
    const unsigned dst0 = ring + st * 8192 + (lane >> 3) * 256 + (lane & 7) * 16;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      cp16(dst0 + i * 1024, p[i] + sl_imm * 256);
      cp16(dst0 + i * 1024 + 128, p[i] + sl_imm * 256 + 128);
    }
  };
  if (MODE == 1) { make_ptrs(cur, pc); make_ptrs(nxt, pn); }
#pragma unroll
  for (int q = 0; q < STAGES - 1; ++q) {
    if (MODE == 0) issue0(cur, q, q); else issue1(pc, q, q);
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  for (; r < rows; r += nwarps) {
    unsigned nn = r + 2 * nwarps < rows ? static_cast<unsigned>(idx[static_cast<size_t>(r + 2 * nwarps) * k + lane]) * row_bytes : 0u;
#pragma unroll
    for (int sl = 0; sl < NSL; ++sl) {
      asm volatile("cp.async.wait_group %0;" ::"n"(STAGES - 2) : "memory");
      __syncwarp();
      int pst = stage + STAGES - 1; if (pst >= STAGES) pst -= STAGES;
      const int ps = sl + STAGES - 1;
      if (ps < NSL) { if (MODE == 0) issue0(cur, ps, pst); else issue1(pc, ps, pst); }
      else if (r + nwarps < rows) { if (MODE == 0) issue0(nxt, ps - NSL, pst); else issue1(pn, ps - NSL, pst); }
      asm volatile("cp.async.commit_group;" ::: "memory");
      const unsigned sb = ring + stage * 8192 + lane * 8;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        unsigned a, b;
        asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(a), "=r"(b) : "r"(sb + j * 256));
        acc ^= a ^ b;
      }
      if (++stage == STAGES) stage = 0;
    }
    cur = nxt; nxt = nn;
    if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < 8; ++i) pc[i] = pn[i];
      make_ptrs(nxt, pn);
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  if (acc == 0x12345678u) sink[0] = acc;
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
  printf("],\n \"cp_async\": [");
  if (d == 384) {
    bool f2 = true;
    auto time_it = [&](auto kern, int smem, const char* name) {
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      const int blocks = sms * 2;
      for (int i = 0; i < 3; ++i) kern<<<blocks, 128, smem>>>(table, idx, rows, k, sink);
      cudaDeviceSynchronize();
      float best_ms = 1e9f;
      for (int rep = 0; rep < 10; ++rep) {
        cudaEventRecord(a);
        kern<<<blocks, 128, smem>>>(table, idx, rows, k, sink);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (ms < best_ms) best_ms = ms;
      }
      printf("%s\n  {\"variant\": \"%s\", \"ms\": %.4f, \"gbs\": %.1f, \"err\": \"%s\"}", f2 ? "" : ",", name, best_ms,
             gathered / (best_ms * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
      f2 = false;
    };
    time_it(gather_cpasync_kernel<0, 3, 3>, 4 * 3 * 8192, "shuffled address per copy, 3 stages, 8 warps/SM");
    time_it(gather_cpasync_kernel<1, 3, 3>, 4 * 3 * 8192, "row pointers + immediates, 3 stages, 8 warps/SM");
    time_it(gather_cpasync_kernel<1, 3, 2>, 4 * 2 * 8192, "row pointers + immediates, 2 stages, 8 warps/SM");
  }
  printf("],\n \"ceiling_gbs\": %.1f, \"ceiling_config\": {\"bytes_per_lane\": %d, \"warps_per_sm\": %d},\n"
         " \"note\": \"best of 10 launches per cell, CUDA events; registers limit the resident warps (ptxas), so high warps_per_sm cells may run in waves\"}\n",
         best, best_vec, best_wps);
  return 0;
}
