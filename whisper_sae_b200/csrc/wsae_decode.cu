// wsae_decode.cu — K2: k-sparse decode with fused bias adds, residual, MSE numerator, L0 count and
// dead-feature "fired" stamps.
//
// Replaces (reference, /root/reference/src/whisper_sae/sae/model.py):
//   :115-116  hidden = zeros; hidden.scatter_(relu(topk_values))   (never densified here)
//   :129      recon = decoder(hidden) + b_pre                      (gather of k decoder rows)
//   :145      mse_loss(recon, x)                                   (sum of squares -> stats[0])
//   :148      l0 = (hidden > 0).sum(-1).mean()                     (integer count -> stats[1])
//   :174-181  step_count += 1; last_activated[fired] = step_count  (stamp written per fired idx;
//                                                                   the +1 itself is in wsae_elementwise.cu)
//
// Layout: the decoder is held feature-major, W_decT[F, d] (one feature's decoder vector is one
// contiguous row), either fp32 (master weights) or bf16 (shadow).  One warp owns one activation
// row at a time; lane l owns elements {128*c + 4*l .. +3} so every gather is a coalesced
// 16-byte (fp32) or 8-byte (bf16) load per lane.
#include "wsae_common.cuh"

namespace wsae {

struct DecodeStats {
  double sse;                     // sum over rows/cols of (recon - target)^2
  unsigned long long l0_count;    // number of selected values > 0
};

template <typename WT>
__device__ __forceinline__ float4 load_w4(const WT* p);
template <>
__device__ __forceinline__ float4 load_w4<float>(const float* p) {
  return __ldg(reinterpret_cast<const float4*>(p));
}
template <>
__device__ __forceinline__ float4 load_w4<__nv_bfloat16>(const __nv_bfloat16* p) {
  return bf16x4_to_float4(__ldg(reinterpret_cast<const uint2*>(p)));
}

// NV = number of float4 column groups per lane (d <= 128 * NV, d % 4 == 0).
template <typename WT, int NV>
__global__ void __launch_bounds__(256)
decode_kernel(const float* __restrict__ target, const WT* __restrict__ w_decT,
              const float* __restrict__ b_dec, const float* __restrict__ b_pre,
              const int32_t* __restrict__ idx, const float* __restrict__ val, int B, int d, int F,
              int k, float* __restrict__ resid, float* __restrict__ recon_out,
              DecodeStats* __restrict__ stats, long long* __restrict__ last_activated,
              const long long* __restrict__ step_count) {
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const int warp_global = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
  const int warp_stride = gridDim.x * warps_per_block;

  float4 bias[NV];
#pragma unroll
  for (int c = 0; c < NV; ++c) {
    const int col = c * 128 + lane * 4;
    bias[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (col < d) {
      float4 a = *reinterpret_cast<const float4*>(b_dec + col);
      if (b_pre != nullptr) {
        const float4 p = *reinterpret_cast<const float4*>(b_pre + col);
        a.x += p.x; a.y += p.y; a.z += p.z; a.w += p.w;
      }
      bias[c] = a;
    }
  }
  const long long stamp = (last_activated != nullptr && step_count != nullptr) ? (*step_count + 1) : 0;

  float sse_local = 0.f;
  unsigned int l0_local = 0;

  for (int row = warp_global; row < B; row += warp_stride) {
    float4 acc[NV];
#pragma unroll
    for (int c = 0; c < NV; ++c) acc[c] = bias[c];
    const int32_t* irow = idx + static_cast<size_t>(row) * k;
    const float* vrow = val + static_cast<size_t>(row) * k;
    for (int j0 = 0; j0 < k; j0 += 32) {
      const int jj = j0 + lane;
      int32_t my_i = -1;
      float my_v = 0.f;
      if (jj < k) {
        my_i = irow[jj];
        my_v = vrow[jj];
      }
      const bool fired = (my_i >= 0) && (my_i < F) && (my_v > 0.f);
      if (fired && last_activated != nullptr) last_activated[my_i] = stamp;
      const uint32_t act_mask = __ballot_sync(0xffffffffu, fired);
      l0_local += (lane == 0) ? __popc(act_mask) : 0;
      // walk only the active features (relu zeroes the rest, model.py:116)
      uint32_t m = act_mask;
      while (m) {
        const int src = __ffs(m) - 1;
        m &= m - 1;
        const int32_t f = __shfl_sync(0xffffffffu, my_i, src);
        float a = __shfl_sync(0xffffffffu, my_v, src);
        // bf16 shadow => bf16 x bf16 product with fp32 accumulation (same as the fused K23 kernel)
        if (sizeof(WT) == 2) a = __bfloat162float(__float2bfloat16_rn(a));
        const WT* wrow = w_decT + static_cast<size_t>(f) * d;
#pragma unroll
        for (int c = 0; c < NV; ++c) {
          const int col = c * 128 + lane * 4;
          if (col < d) {
            const float4 w = load_w4<WT>(wrow + col);
            acc[c].x = fmaf(a, w.x, acc[c].x);
            acc[c].y = fmaf(a, w.y, acc[c].y);
            acc[c].z = fmaf(a, w.z, acc[c].z);
            acc[c].w = fmaf(a, w.w, acc[c].w);
          }
        }
      }
    }
    const float* trow = target + static_cast<size_t>(row) * d;
#pragma unroll
    for (int c = 0; c < NV; ++c) {
      const int col = c * 128 + lane * 4;
      if (col < d) {
        const float4 t = *reinterpret_cast<const float4*>(trow + col);
        if (recon_out != nullptr)
          *reinterpret_cast<float4*>(recon_out + static_cast<size_t>(row) * d + col) = acc[c];
        float4 r;
        r.x = acc[c].x - t.x; r.y = acc[c].y - t.y; r.z = acc[c].z - t.z; r.w = acc[c].w - t.w;
        if (resid != nullptr)
          *reinterpret_cast<float4*>(resid + static_cast<size_t>(row) * d + col) = r;
        sse_local = fmaf(r.x, r.x, sse_local);
        sse_local = fmaf(r.y, r.y, sse_local);
        sse_local = fmaf(r.z, r.z, sse_local);
        sse_local = fmaf(r.w, r.w, sse_local);
      }
    }
  }

  // block reduction -> one double atomic + one integer atomic per block
  __shared__ float s_sse[8];
  __shared__ unsigned int s_l0[8];
  const float wsum = warp_sum(sse_local);
  if (lane == 0) {
    s_sse[threadIdx.x >> 5] = wsum;
    s_l0[threadIdx.x >> 5] = l0_local;
  }
  __syncthreads();
  if (threadIdx.x == 0 && stats != nullptr) {
    double t = 0.0;
    unsigned long long c = 0;
    for (int w = 0; w < warps_per_block; ++w) {
      t += static_cast<double>(s_sse[w]);
      c += s_l0[w];
    }
    atomicAdd(&stats->sse, t);
    atomicAdd(&stats->l0_count, c);
  }
}

// Generic scalar fallback for d % 4 != 0 (still a CUDA kernel; slow, used only for odd shapes).
template <typename WT>
__global__ void __launch_bounds__(256)
decode_kernel_generic(const float* __restrict__ target, const WT* __restrict__ w_decT,
                      const float* __restrict__ b_dec, const float* __restrict__ b_pre,
                      const int32_t* __restrict__ idx, const float* __restrict__ val, int B, int d,
                      int F, int k, float* __restrict__ resid, float* __restrict__ recon_out,
                      DecodeStats* __restrict__ stats, long long* __restrict__ last_activated,
                      const long long* __restrict__ step_count) {
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const int warp_global = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
  const int warp_stride = gridDim.x * warps_per_block;
  const long long stamp = (last_activated != nullptr && step_count != nullptr) ? (*step_count + 1) : 0;
  float sse_local = 0.f;
  unsigned int l0_local = 0;
  for (int row = warp_global; row < B; row += warp_stride) {
    const int32_t* irow = idx + static_cast<size_t>(row) * k;
    const float* vrow = val + static_cast<size_t>(row) * k;
    for (int j = lane; j < k; j += 32) {
      const int32_t f = irow[j];
      if (f >= 0 && f < F && vrow[j] > 0.f) {
        ++l0_local;
        if (last_activated != nullptr) last_activated[f] = stamp;
      }
    }
    for (int col = lane; col < d; col += 32) {
      float acc = b_dec[col] + (b_pre != nullptr ? b_pre[col] : 0.f);
      for (int j = 0; j < k; ++j) {
        const int32_t f = irow[j];
        const float a = vrow[j];
        if (f >= 0 && f < F && a > 0.f)
          acc = fmaf(sizeof(WT) == 2 ? __bfloat162float(__float2bfloat16_rn(a)) : a,
                     static_cast<float>(w_decT[static_cast<size_t>(f) * d + col]), acc);
      }
      const float r = acc - target[static_cast<size_t>(row) * d + col];
      if (recon_out != nullptr) recon_out[static_cast<size_t>(row) * d + col] = acc;
      if (resid != nullptr) resid[static_cast<size_t>(row) * d + col] = r;
      sse_local = fmaf(r, r, sse_local);
    }
  }
  const float wsum = warp_sum(sse_local);
  unsigned int l0w = l0_local;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) l0w += __shfl_xor_sync(0xffffffffu, l0w, o);
  if (lane == 0 && stats != nullptr) {
    atomicAdd(&stats->sse, static_cast<double>(wsum));
    atomicAdd(&stats->l0_count, static_cast<unsigned long long>(l0w));
  }
}

template <typename WT>
static int launch_decode(const float* target, const void* w, const float* b_dec,
                         const float* b_pre, const int32_t* idx, const float* val, int B, int d,
                         int F, int k, float* resid, float* recon, void* stats,
                         long long* last_activated, const long long* step_count,
                         cudaStream_t stream) {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int threads = 256, wpb = threads / 32;
  int blocks = ceil_div(B, wpb);
  const int cap = sms * 8;
  if (blocks > cap) blocks = cap;
  const WT* wt = static_cast<const WT*>(w);
  DecodeStats* st = static_cast<DecodeStats*>(stats);
#define WSAE_DECODE_CASE(NVV)                                                                  \
  decode_kernel<WT, NVV><<<blocks, threads, 0, stream>>>(target, wt, b_dec, b_pre, idx, val, B, \
                                                         d, F, k, resid, recon, st,             \
                                                         last_activated, step_count)
  if (d % 4 != 0 || d > 128 * 16) {
    decode_kernel_generic<WT><<<blocks, threads, 0, stream>>>(target, wt, b_dec, b_pre, idx, val, B,
                                                             d, F, k, resid, recon, st,
                                                             last_activated, step_count);
  } else {
    const int nv = ceil_div(d, 128);
    if (nv <= 1) WSAE_DECODE_CASE(1);
    else if (nv <= 2) WSAE_DECODE_CASE(2);
    else if (nv <= 3) WSAE_DECODE_CASE(3);
    else if (nv <= 4) WSAE_DECODE_CASE(4);
    else if (nv <= 6) WSAE_DECODE_CASE(6);
    else if (nv <= 8) WSAE_DECODE_CASE(8);
    else if (nv <= 10) WSAE_DECODE_CASE(10);
    else if (nv <= 12) WSAE_DECODE_CASE(12);
    else WSAE_DECODE_CASE(16);
  }
#undef WSAE_DECODE_CASE
  return static_cast<int>(cudaGetLastError());
}

}  // namespace wsae

using namespace wsae;

extern "C" int wsae_decode_mse(const float* target, const void* w_decT, int w_is_bf16,
                               const float* b_dec, const float* b_pre, const int32_t* idx,
                               const float* val, int B, int d, int F, int k, float* resid,
                               float* recon, void* stats, long long* last_activated,
                               const long long* step_count, cudaStream_t stream) {
  if (!target || !w_decT || !b_dec || !idx || !val) return kBadArg;
  if (B <= 0 || d <= 0 || F <= 0 || k <= 0) return kBadArg;
  if (w_is_bf16)
    return launch_decode<__nv_bfloat16>(target, w_decT, b_dec, b_pre, idx, val, B, d, F, k, resid,
                                        recon, stats, last_activated, step_count, stream);
  return launch_decode<float>(target, w_decT, b_dec, b_pre, idx, val, B, d, F, k, resid, recon,
                              stats, last_activated, step_count, stream);
}
