// wsae_backward.cu — K3: sparse backward of the TopK-SAE loss (what autograd derives from
// /root/reference/src/whisper_sae/sae/model.py:108-145 and runs at sae/training.py:184).
//
// With  r = recon - target,  s = grad_out * 2 / (B_total * d),  g = s * r,  h_j = relu(v_j):
//   db_dec            = sum_b g                                   (mse_loss_backward + Linear bias)
//   dW_decT[i_j, :]  += h_j * g                                   (decoder Linear weight grad, sparse)
//   dv_j              = [v_j > 0] * (g . W_decT[i_j, :])          (scatter/relu/topk backward)
//   dW_enc[i_j, :]   += dv_j * xc,   xc = x - b_pre               (encoder Linear weight grad, sparse)
//   db_enc[i_j]      += dv_j
//   db_pre            = db_dec - db_enc . W_enc                   (wsae_bpre_grad below)
//   dx                = dpre . W_enc - g                          (wsae_input_grad, only if x needs grad)
// Nothing of shape [B, F] is ever materialised.  This file is the fp32 "scatter" implementation
// (red.global.add); the bf16 production path replaces the two weight-gradient scatters by the
// tensor-core kernel in wsae_wgrad_gemm.cu and keeps only the dv / bias parts from here.
#include "wsae_common.cuh"

namespace wsae {

template <typename WT>
__device__ __forceinline__ float4 bw_load_w4(const WT* p);
template <>
__device__ __forceinline__ float4 bw_load_w4<float>(const float* p) {
  return __ldg(reinterpret_cast<const float4*>(p));
}
template <>
__device__ __forceinline__ float4 bw_load_w4<__nv_bfloat16>(const __nv_bfloat16* p) {
  return bf16x4_to_float4(__ldg(reinterpret_cast<const uint2*>(p)));
}

// One warp per activation row.  scatter_w: 1 = also scatter-add dW_enc / dW_decT (fp32 path).
template <typename WT, int NV>
__global__ void __launch_bounds__(256)
backward_sparse_kernel(const float* __restrict__ resid, const float* __restrict__ x,
                       const float* __restrict__ b_pre, const WT* __restrict__ w_decT,
                       const int32_t* __restrict__ idx, const float* __restrict__ val,
                       const float* __restrict__ grad_out, float coef, int B, int d, int F, int k,
                       float* __restrict__ d_w_enc, float* __restrict__ d_w_decT,
                       float* __restrict__ d_b_enc, float* __restrict__ d_b_dec,
                       float* __restrict__ dpre_val, __nv_bfloat16* __restrict__ resid_bf16) {
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const int warp_global = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
  const int warp_stride = gridDim.x * warps_per_block;
  const float s = coef * (grad_out != nullptr ? *grad_out : 1.f);

  // Wide rows (NV > 8: d > 1024) cannot keep b_pre and the running db_dec sums in registers next to the
  // row's gradient and centred input (4 x NV float4 = 256 registers at NV = 16: 504 bytes of spills in
  // round 1): there b_pre is re-read (L1 resident) and db_dec accumulates in shared memory row by row.
  constexpr bool kLean = NV > 8;
  constexpr int NR = kLean ? 1 : NV;
  extern __shared__ float s_g[];  // [d] db_dec of this block
  for (int i = threadIdx.x; i < d; i += blockDim.x) s_g[i] = 0.f;
  __syncthreads();
  float4 bp[NR];
  float4 gsum[NR];
#pragma unroll
  for (int c = 0; c < NR; ++c) {
    const int col = c * 128 + lane * 4;
    bp[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    gsum[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!kLean && col < d && b_pre != nullptr) bp[c] = *reinterpret_cast<const float4*>(b_pre + col);
  }

  for (int row = warp_global; row < B; row += warp_stride) {
    float4 g[NV], xc[NV];
#pragma unroll
    for (int c = 0; c < NV; ++c) {
      const int col = c * 128 + lane * 4;
      g[c] = make_float4(0.f, 0.f, 0.f, 0.f);
      xc[c] = g[c];
      if (col < d) {
        const float4 r = *reinterpret_cast<const float4*>(resid + static_cast<size_t>(row) * d + col);
        g[c] = make_float4(s * r.x, s * r.y, s * r.z, s * r.w);
        if (resid_bf16 != nullptr) {
          __nv_bfloat162 lo = __floats2bfloat162_rn(r.x, r.y);
          __nv_bfloat162 hi = __floats2bfloat162_rn(r.z, r.w);
          uint2 o;
          o.x = *reinterpret_cast<uint32_t*>(&lo);
          o.y = *reinterpret_cast<uint32_t*>(&hi);
          *reinterpret_cast<uint2*>(resid_bf16 + static_cast<size_t>(row) * d + col) = o;
        }
        if (kLean) {
          atomicAdd(&s_g[col + 0], g[c].x);
          atomicAdd(&s_g[col + 1], g[c].y);
          atomicAdd(&s_g[col + 2], g[c].z);
          atomicAdd(&s_g[col + 3], g[c].w);
        } else {
          gsum[c].x += g[c].x; gsum[c].y += g[c].y; gsum[c].z += g[c].z; gsum[c].w += g[c].w;
        }
        if (d_w_enc != nullptr) {
          const float4 xv = *reinterpret_cast<const float4*>(x + static_cast<size_t>(row) * d + col);
          float4 b = kLean ? make_float4(0.f, 0.f, 0.f, 0.f) : bp[c];
          if (kLean && b_pre != nullptr) b = __ldg(reinterpret_cast<const float4*>(b_pre + col));
          xc[c] = make_float4(xv.x - b.x, xv.y - b.y, xv.z - b.z, xv.w - b.w);
        }
      }
    }
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
      float my_dv = 0.f;
      uint32_t m = __ballot_sync(0xffffffffu, fired);
      while (m) {
        const int src = __ffs(m) - 1;
        m &= m - 1;
        const int32_t f = __shfl_sync(0xffffffffu, my_i, src);
        const float h = __shfl_sync(0xffffffffu, my_v, src);
        const WT* wrow = w_decT + static_cast<size_t>(f) * d;
        float dot = 0.f;
#pragma unroll
        for (int c = 0; c < NV; ++c) {
          const int col = c * 128 + lane * 4;
          if (col < d) {
            const float4 w = bw_load_w4<WT>(wrow + col);
            dot = fmaf(g[c].x, w.x, dot);
            dot = fmaf(g[c].y, w.y, dot);
            dot = fmaf(g[c].z, w.z, dot);
            dot = fmaf(g[c].w, w.w, dot);
          }
        }
        dot = warp_sum(dot);  // dv for feature f of this row (v > 0 already known)
        if (lane == src) my_dv = dot;
        if (d_w_decT != nullptr) {
          float* drow = d_w_decT + static_cast<size_t>(f) * d;
#pragma unroll
          for (int c = 0; c < NV; ++c) {
            const int col = c * 128 + lane * 4;
            if (col < d) red_add_f32x4(drow + col, h * g[c].x, h * g[c].y, h * g[c].z, h * g[c].w);
          }
        }
        if (d_w_enc != nullptr) {
          float* erow = d_w_enc + static_cast<size_t>(f) * d;
#pragma unroll
          for (int c = 0; c < NV; ++c) {
            const int col = c * 128 + lane * 4;
            if (col < d)
              red_add_f32x4(erow + col, dot * xc[c].x, dot * xc[c].y, dot * xc[c].z, dot * xc[c].w);
          }
        }
      }
      if (jj < k) {
        if (dpre_val != nullptr) dpre_val[static_cast<size_t>(row) * k + jj] = my_dv;
        if (fired && d_b_enc != nullptr) atomicAdd(d_b_enc + my_i, my_dv);
      }
    }
  }

  // db_dec: per-block shared reduction, then one atomic per column per block
#pragma unroll
  for (int c = 0; c < (kLean ? 0 : NV); ++c) {
    const int col = c * 128 + lane * 4;
    if (col < d) {
      atomicAdd(&s_g[col + 0], gsum[c].x);
      atomicAdd(&s_g[col + 1], gsum[c].y);
      atomicAdd(&s_g[col + 2], gsum[c].z);
      atomicAdd(&s_g[col + 3], gsum[c].w);
    }
  }
  __syncthreads();
  if (d_b_dec != nullptr)
    for (int i = threadIdx.x; i < d; i += blockDim.x) atomicAdd(d_b_dec + i, s_g[i]);
}

// db_pre[c] = db_dec[c] - sum_f db_enc[f] * W_enc[f, c].  One block per chunk of 32 features;
// thread t owns columns {t, t + blockDim, ...}; features with a zero bias gradient (never selected
// in this batch) are skipped, so only fired rows of W_enc are read.
constexpr int kBpreFeat = 64;
// One block = 64 features x all columns: 8-feature blocks issued F/8 * d atomics onto the same d
// addresses (6.5 M at 1280 -> 40960: 2.2 ms, 1.5 % of HBM speed; 0.26 ms at 768 -> 6144).
__global__ void __launch_bounds__(256)
bpre_grad_kernel(const float* __restrict__ d_b_dec, const float* __restrict__ d_b_enc,
                 const float* __restrict__ w_enc, int F, int d, float* __restrict__ d_b_pre) {
  __shared__ float s_coef[kBpreFeat];
  __shared__ int s_any;
  const int f0 = blockIdx.x * kBpreFeat;
  if (threadIdx.x == 0) s_any = 0;
  __syncthreads();
  if (threadIdx.x < kBpreFeat) {
    const int f = f0 + threadIdx.x;
    const float c = f < F ? d_b_enc[f] : 0.f;
    s_coef[threadIdx.x] = c;
    if (c != 0.f) s_any = 1;
  }
  __syncthreads();
  if (s_any == 0 && blockIdx.x != 0) return;            // no fired feature in this block
  for (int col = threadIdx.x; col < d; col += blockDim.x) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;      // four independent chains; loads stay coalesced per feature row
    const float* wp = w_enc + static_cast<size_t>(f0) * d + col;
#pragma unroll 4
    for (int j = 0; j < kBpreFeat; j += 4) {
      const float c0 = s_coef[j], c1 = s_coef[j + 1], c2 = s_coef[j + 2], c3 = s_coef[j + 3];
      if (c0 != 0.f) a0 = fmaf(c0, __ldg(wp + static_cast<size_t>(j) * d), a0);
      if (c1 != 0.f) a1 = fmaf(c1, __ldg(wp + static_cast<size_t>(j + 1) * d), a1);
      if (c2 != 0.f) a2 = fmaf(c2, __ldg(wp + static_cast<size_t>(j + 2) * d), a2);
      if (c3 != 0.f) a3 = fmaf(c3, __ldg(wp + static_cast<size_t>(j + 3) * d), a3);
    }
    float out = -((a0 + a1) + (a2 + a3));
    if (blockIdx.x == 0) out += d_b_dec[col];
    if (out != 0.f) atomicAdd(d_b_pre + col, out);
  }
}

// Deterministic form of bpre_grad: chunk c of kBpreDetFeat features writes its partial GEMV row to
// ws[c][:] (features in index order), bpre_det_reduce_kernel subtracts the chunks in order.
constexpr int kBpreDetFeat = 256;
__global__ void __launch_bounds__(256)
bpre_det_partial_kernel(const float* __restrict__ d_b_enc, const float* __restrict__ w_enc, int F, int d,
                        float* __restrict__ ws) {
  __shared__ float s_coef[kBpreDetFeat];
  const int f0 = blockIdx.x * kBpreDetFeat;
  for (int i = threadIdx.x; i < kBpreDetFeat; i += blockDim.x) s_coef[i] = f0 + i < F ? d_b_enc[f0 + i] : 0.f;
  __syncthreads();
  for (int col = threadIdx.x; col < d; col += blockDim.x) {
    float acc = 0.f;
    for (int j = 0; j < kBpreDetFeat; ++j) {
      const float c = s_coef[j];
      if (c != 0.f) acc = fmaf(c, __ldg(w_enc + static_cast<size_t>(f0 + j) * d + col), acc);
    }
    ws[static_cast<size_t>(blockIdx.x) * d + col] = acc;
  }
}
__global__ void __launch_bounds__(256)
bpre_det_reduce_kernel(const float* __restrict__ d_b_dec, const float* __restrict__ ws, int nchunk, int d,
                       float* __restrict__ d_b_pre) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= d) return;
  float acc = 0.f;
  for (int c = 0; c < nchunk; ++c) acc += ws[static_cast<size_t>(c) * d + col];
  d_b_pre[col] = d_b_dec[col] - acc;
}

// dx[b, :] = sum_j dv_j * W_enc[i_j, :] - g[b, :]    (only when the input itself requires grad)
__global__ void __launch_bounds__(256)
input_grad_kernel(const float* __restrict__ resid, const float* __restrict__ w_enc,
                  const int32_t* __restrict__ idx, const float* __restrict__ dpre_val,
                  const float* __restrict__ grad_out, float coef, int B, int d, int F, int k,
                  int subtract_g, float* __restrict__ dx) {
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const int row = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
  if (row >= B) return;
  const float s = coef * (grad_out != nullptr ? *grad_out : 1.f);
  for (int col = lane; col < d; col += 32) {
    float acc = subtract_g ? -s * resid[static_cast<size_t>(row) * d + col] : 0.f;
    for (int j = 0; j < k; ++j) {
      const int32_t f = idx[static_cast<size_t>(row) * k + j];
      const float dv = dpre_val[static_cast<size_t>(row) * k + j];
      if (f >= 0 && f < F && dv != 0.f) acc = fmaf(dv, w_enc[static_cast<size_t>(f) * d + col], acc);
    }
    dx[static_cast<size_t>(row) * d + col] = acc;
  }
}

// out[idx[b,j], :] += vals[b,j] * (rows[b, :] - center)   — the encoder weight gradient
// dW_enc = dpre^T . (x - b_pre) as a sparse scatter, for the case where the input width differs
// from the decoder width (transcoders: transcoder.py:105-138), which K3's single-`d` form cannot do.
__global__ void __launch_bounds__(256)
scatter_rows_kernel(const float* __restrict__ rows, const float* __restrict__ center,
                    const int32_t* __restrict__ idx, const float* __restrict__ vals, int B, int dr,
                    int F, int k, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= B) return;
  for (int col = lane * 4; col < dr; col += 128) {
    float4 xv = *reinterpret_cast<const float4*>(rows + static_cast<size_t>(row) * dr + col);
    if (center != nullptr) {
      const float4 c = *reinterpret_cast<const float4*>(center + col);
      xv.x -= c.x; xv.y -= c.y; xv.z -= c.z; xv.w -= c.w;
    }
    for (int j = 0; j < k; ++j) {
      const int32_t f = idx[static_cast<size_t>(row) * k + j];
      const float v = vals[static_cast<size_t>(row) * k + j];
      if (f >= 0 && f < F && v != 0.f)
        red_add_f32x4(out + static_cast<size_t>(f) * dr + col, v * xv.x, v * xv.y, v * xv.z, v * xv.w);
    }
  }
}

template <typename WT>
static int launch_backward(const float* resid, const float* x, const float* b_pre, const void* w,
                           const int32_t* idx, const float* val, const float* grad_out, float coef,
                           int B, int d, int F, int k, float* d_w_enc, float* d_w_decT,
                           float* d_b_enc, float* d_b_dec, float* dpre_val, void* resid_bf16,
                           cudaStream_t stream) {
  if (d % 4 != 0 || d > 128 * 16) return kUnsupported;
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int threads = 256, wpb = threads / 32;
  int blocks = ceil_div(B, wpb);
  const int cap = sms * 4;
  if (blocks > cap) blocks = cap;
  const size_t smem = static_cast<size_t>(d) * sizeof(float);
  const WT* wt = static_cast<const WT*>(w);
#define WSAE_BWD_CASE(NVV)                                                                       \
  backward_sparse_kernel<WT, NVV><<<blocks, threads, smem, stream>>>(                            \
      resid, x, b_pre, wt, idx, val, grad_out, coef, B, d, F, k, d_w_enc, d_w_decT, d_b_enc,      \
      d_b_dec, dpre_val, static_cast<__nv_bfloat16*>(resid_bf16))
  const int nv = ceil_div(d, 128);
  if (nv <= 1) WSAE_BWD_CASE(1);
  else if (nv <= 2) WSAE_BWD_CASE(2);
  else if (nv <= 3) WSAE_BWD_CASE(3);
  else if (nv <= 4) WSAE_BWD_CASE(4);
  else if (nv <= 6) WSAE_BWD_CASE(6);
  else if (nv <= 8) WSAE_BWD_CASE(8);
  else if (nv <= 10) WSAE_BWD_CASE(10);
  else if (nv <= 12) WSAE_BWD_CASE(12);
  else WSAE_BWD_CASE(16);
#undef WSAE_BWD_CASE
  return static_cast<int>(cudaGetLastError());
}

}  // namespace wsae

using namespace wsae;

extern "C" int wsae_backward_sparse(const float* resid, const float* x, const float* b_pre,
                                    const void* w_decT, int w_is_bf16, const int32_t* idx,
                                    const float* val, const float* grad_out, float coef, int B,
                                    int d, int F, int k, float* d_w_enc, float* d_w_decT,
                                    float* d_b_enc, float* d_b_dec, float* dpre_val,
                                    void* resid_bf16, cudaStream_t stream) {
  if (!resid || !w_decT || !idx || !val) return kBadArg;
  if (d_w_enc != nullptr && x == nullptr) return kBadArg;
  if (B <= 0 || d <= 0 || F <= 0 || k <= 0) return kBadArg;
  if (w_is_bf16)
    return launch_backward<__nv_bfloat16>(resid, x, b_pre, w_decT, idx, val, grad_out, coef, B, d,
                                          F, k, d_w_enc, d_w_decT, d_b_enc, d_b_dec, dpre_val,
                                          resid_bf16, stream);
  return launch_backward<float>(resid, x, b_pre, w_decT, idx, val, grad_out, coef, B, d, F, k,
                                d_w_enc, d_w_decT, d_b_enc, d_b_dec, dpre_val, resid_bf16, stream);
}

extern "C" int wsae_scatter_rows(const float* rows, const float* center, const int32_t* idx,
                                 const float* vals, int B, int dr, int F, int k, float* out,
                                 cudaStream_t stream) {
  if (!rows || !idx || !vals || !out || B <= 0 || dr <= 0 || F <= 0 || k <= 0) return kBadArg;
  if (dr % 4 != 0) return kUnsupported;
  const int warps = 8;
  scatter_rows_kernel<<<ceil_div(B, warps), warps * 32, 0, stream>>>(rows, center, idx, vals, B, dr,
                                                                     F, k, out);
  return static_cast<int>(cudaGetLastError());
}

extern "C" int wsae_bpre_grad(const float* d_b_dec, const float* d_b_enc, const float* w_enc,
                              int F, int d, float* d_b_pre, cudaStream_t stream) {
  if (!d_b_dec || !d_b_enc || !w_enc || !d_b_pre || F <= 0 || d <= 0) return kBadArg;
  cudaError_t e = cudaMemsetAsync(d_b_pre, 0, static_cast<size_t>(d) * sizeof(float), stream);
  if (e != cudaSuccess) return static_cast<int>(e);
  const int threads = d >= 256 ? 256 : 128;
  bpre_grad_kernel<<<ceil_div(F, kBpreFeat), threads, 0, stream>>>(d_b_dec, d_b_enc, w_enc, F, d,
                                                                 d_b_pre);
  return static_cast<int>(cudaGetLastError());
}

// db_pre = db_dec - db_enc . W_enc in a fixed summation order; ws: ceil(F / 256) * d floats.
extern "C" int wsae_bpre_grad_det(const float* d_b_dec, const float* d_b_enc, const float* w_enc,
                                  int F, int d, float* d_b_pre, float* ws, cudaStream_t stream) {
  if (!d_b_dec || !d_b_enc || !w_enc || !d_b_pre || !ws || F <= 0 || d <= 0) return kBadArg;
  const int nchunk = ceil_div(F, kBpreDetFeat);
  bpre_det_partial_kernel<<<nchunk, 256, 0, stream>>>(d_b_enc, w_enc, F, d, ws);
  bpre_det_reduce_kernel<<<ceil_div(d, 256), 256, 0, stream>>>(d_b_dec, ws, nchunk, d, d_b_pre);
  return static_cast<int>(cudaGetLastError());
}

extern "C" int wsae_input_grad(const float* resid, const float* w_enc, const int32_t* idx,
                               const float* dpre_val, const float* grad_out, float coef, int B,
                               int d, int F, int k, int subtract_g, float* dx,
                               cudaStream_t stream) {
  if (!resid || !w_enc || !idx || !dpre_val || !dx) return kBadArg;
  const int warps = 8;
  input_grad_kernel<<<ceil_div(B, warps), warps * 32, 0, stream>>>(
      resid, w_enc, idx, dpre_val, grad_out, coef, B, d, F, k, subtract_g, dx);
  return static_cast<int>(cudaGetLastError());
}
