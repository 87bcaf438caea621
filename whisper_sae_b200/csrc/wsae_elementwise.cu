// wsae_elementwise.cu — K5 and friends: HBM-bound fused elementwise / reduction kernels.
//
//   wsae_renorm_decoder      model.py:91-96   W_dec <- F.normalize(W_dec, dim=0)   (rows of W_decT)
//   wsae_counters_update     model.py:174,183-195   step_count += 1 ; dead = (step - last) > thr
//   wsae_densify_hidden      model.py:115-116 zeros_like + scatter_(relu(topk_values))
//   wsae_cast_bf16           bf16 shadow of the decoder for the gather kernels
//   wsae_fused_adamw         training.py:187-198   clip (global-norm scale) + AdamW (+ renorm input)
//   wsae_sumsq               training.py:188-191   squared L2 norm of a gradient tensor (clip_grad_norm_)
#include "wsae_common.cuh"

namespace wsae {

// One warp per feature row of W_decT[F, d]: row /= max(||row||_2, eps).  Optionally also emits the
// bf16 shadow row so the refresh costs no extra pass.
__global__ void __launch_bounds__(256)
renorm_rows_kernel(float* __restrict__ w, int F, int d, float eps,
                   __nv_bfloat16* __restrict__ shadow) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= F) return;
  float* p = w + static_cast<size_t>(row) * d;
  float ss = 0.f;
  for (int c = lane; c < d; c += 32) {
    const float v = p[c];
    ss = fmaf(v, v, ss);
  }
  ss = warp_sum(ss);
  const float denom = fmaxf(sqrtf(ss), eps);
  for (int c = lane; c < d; c += 32) {
    const float v = p[c] / denom;
    p[c] = v;
    if (shadow != nullptr) shadow[static_cast<size_t>(row) * d + c] = __float2bfloat16_rn(v);
  }
}

// Single block: optional step_count bump, then count dead features against the *new* step.
// Optional mailbox (host-mapped pinned memory, 4 x int64): once the dead count is known the step's
// metrics are complete - {stats2[0] (SSE, f64 bits), stats2[1] (L0 count), dead count} are posted to
// the mailbox, then the sequence number *seq_src, behind a system-scope fence.  The host polls the
// sequence word instead of synchronising with the stream, so it reads the metrics while the rest of
// the step (weight gradients, optimizer) is still running and has the next step queued before the
// GPU goes idle.
__global__ void __launch_bounds__(1024)
counters_update_kernel(const long long* __restrict__ last_activated, long long* step_count, int F,
                       long long threshold, int bump, long long* __restrict__ dead_count,
                       const long long* __restrict__ stats2, const long long* __restrict__ seq_src,
                       long long* mailbox) {
  pdl_prologue();
  __shared__ long long s_step;
  __shared__ int s_part[32];
  if (threadIdx.x == 0) {
    long long sc = *step_count + (bump ? 1 : 0);
    if (bump) *step_count = sc;
    s_step = sc;
  }
  __syncthreads();
  const long long sc = s_step;
  int c = 0;
  for (int f = threadIdx.x; f < F; f += blockDim.x) c += ((sc - last_activated[f]) > threshold) ? 1 : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0 && (dead_count != nullptr || mailbox != nullptr)) {
    long long t = 0;
    for (int w = 0; w < (blockDim.x >> 5); ++w) t += s_part[w];
    if (dead_count != nullptr) *dead_count = t;
    if (mailbox != nullptr) {
      volatile long long* mb = mailbox;
      mb[0] = stats2 != nullptr ? stats2[0] : 0;
      mb[1] = stats2 != nullptr ? stats2[1] : 0;
      mb[2] = t;
      __threadfence_system();
      mb[3] = *seq_src;
    }
  }
}

// hidden[B, F] = 0 ; hidden[b, idx[b, j]] = relu(val[b, j]).  One block per row.
__global__ void __launch_bounds__(256)
densify_kernel(const int32_t* __restrict__ idx, const float* __restrict__ val, int B, int F, int k,
               float* __restrict__ hidden) {
  const int row = blockIdx.x;
  float* h = hidden + static_cast<size_t>(row) * F;
  for (int c = threadIdx.x; c < F; c += blockDim.x) h[c] = 0.f;
  __syncthreads();
  for (int j = threadIdx.x; j < k; j += blockDim.x) {
    const int32_t f = idx[static_cast<size_t>(row) * k + j];
    const float v = val[static_cast<size_t>(row) * k + j];
    if (f >= 0 && f < F) h[f] = fmaxf(v, 0.f);
  }
}

__global__ void __launch_bounds__(256)
cast_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, size_t n) {
  pdl_prologue();
  const size_t i = (static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    const float4 v = *reinterpret_cast<const float4*>(src + i);
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y);
    __nv_bfloat162 b = __floats2bfloat162_rn(v.z, v.w);
    uint2 o;
    o.x = *reinterpret_cast<uint32_t*>(&a);
    o.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(dst + i) = o;
  } else {
    for (size_t j = i; j < n; ++j) dst[j] = __float2bfloat16_rn(src[j]);
  }
}

// out[0] += sum(g^2)  (double accumulator; caller zeroes it).  Grid-stride, float4.
__global__ void __launch_bounds__(256)
sumsq_kernel(const float* __restrict__ g, size_t n, double* __restrict__ out) {
  pdl_prologue();
  float acc = 0.f;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x * 4;
  size_t i = (static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
  for (; i + 3 < n; i += stride) {
    const float4 v = *reinterpret_cast<const float4*>(g + i);
    acc = fmaf(v.x, v.x, acc);
    acc = fmaf(v.y, v.y, acc);
    acc = fmaf(v.z, v.z, acc);
    acc = fmaf(v.w, v.w, acc);
  }
  if (i < n && i + 3 >= n)
    for (size_t j = i; j < n; ++j) acc = fmaf(g[j], g[j], acc);
  __shared__ float s[8];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (blockDim.x >> 5); ++w) t += static_cast<double>(s[w]);
    atomicAdd(out, t);
  }
}

// Deterministic form of sumsq: every block writes its partial to ws[blockIdx.x] (float4 grid-stride
// sums in a fixed order), sumsq_ordered_kernel adds them up in index order.
__global__ void __launch_bounds__(256)
sumsq_partial_kernel(const float* __restrict__ g, size_t n, double* __restrict__ ws) {
  float acc = 0.f;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x * 4;
  size_t i = (static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
  for (; i + 3 < n; i += stride) {
    const float4 v = *reinterpret_cast<const float4*>(g + i);
    acc = fmaf(v.x, v.x, acc);
    acc = fmaf(v.y, v.y, acc);
    acc = fmaf(v.z, v.z, acc);
    acc = fmaf(v.w, v.w, acc);
  }
  if (i < n && i + 3 >= n)
    for (size_t j = i; j < n; ++j) acc = fmaf(g[j], g[j], acc);
  __shared__ float s[8];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (blockDim.x >> 5); ++w) t += static_cast<double>(s[w]);
    ws[blockIdx.x] = t;
  }
}
__global__ void sumsq_ordered_kernel(const double* __restrict__ ws, int nblocks, double* __restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double t = 0.0;
    for (int b = 0; b < nblocks; ++b) t += ws[b];
    *out += t;
  }
}

// AdamW with the clip_grad_norm_ scale folded in (torch.optim.AdamW semantics, amsgrad=False):
//   g   = grad * min(1, max_norm / (sqrt(sumsq) + 1e-6))
//   p  *= 1 - lr * wd ; m = b1*m + (1-b1)*g ; v = b2*v + (1-b2)*g*g
//   p  -= (lr / bc1) * m / (sqrt(v) / sqrt(bc2) + eps)
// hyper[] lives in device memory so a captured CUDA graph can be replayed with new lr / step:
//   hyper = {lr, beta1, beta2, eps, weight_decay, bias_corr1, bias_corr2_sqrt, max_norm}
__global__ void __launch_bounds__(256)
fused_adamw_kernel(float* __restrict__ p, const float* __restrict__ grad, float* __restrict__ m,
                   float* __restrict__ v, size_t n, const float* __restrict__ hyper,
                   const double* __restrict__ grad_sumsq) {
  const float lr = hyper[0], b1 = hyper[1], b2 = hyper[2], eps = hyper[3], wd = hyper[4];
  const float bc1 = hyper[5], bc2s = hyper[6], max_norm = hyper[7];
  float clip = 1.f;
  if (grad_sumsq != nullptr && max_norm > 0.f) {
    const float total = static_cast<float>(sqrt(*grad_sumsq));
    const float c = max_norm / (total + 1e-6f);
    clip = c < 1.f ? c : 1.f;
  }
  const float step_size = lr / bc1;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float g = grad[i] * clip;
    float pv = p[i];
    pv *= 1.f - lr * wd;
    const float mv = m[i] + (g - m[i]) * (1.f - b1);          // lerp form, as torch does
    const float vv = b2 * v[i] + (1.f - b2) * g * g;
    const float denom = sqrtf(vv) / bc2s + eps;
    pv -= step_size * (mv / denom);
    p[i] = pv;
    m[i] = mv;
    v[i] = vv;
  }
}

// ---- multi-tensor AdamW (+ decoder renorm) in ONE launch --------------------------------------
// Descriptor of one parameter tensor; row_len > 0 marks a [n / row_len, row_len] matrix whose rows
// are re-normalised to unit L2 norm right after the update (sae/training.py:193-198: optimizer.step()
// then normalize_decoder_weights(); sae/model.py:91-96), one warp per row, in the same pass.
struct AdamwTensor {
  float* p;
  const float* g;
  float* m;
  float* v;
  long long n;
  int row_len;
  int pad_;
};
constexpr int kAdamwMaxTensors = 8;
struct AdamwBatch {
  AdamwTensor t[kAdamwMaxTensors];
  long long unit_start[kAdamwMaxTensors + 1];   // prefix of work units (1024 elements or 8 rows)
  int count;
};

__device__ __forceinline__ void adamw_elem(float& pv, float g, float& mv, float& vv, float lr,
                                           float b1, float b2, float eps, float wd, float bc2s,
                                           float step_size) {
  pv *= 1.f - lr * wd;
  mv = mv + (g - mv) * (1.f - b1);
  vv = b2 * vv + (1.f - b2) * g * g;
  const float denom = sqrtf(vv) / bc2s + eps;
  pv -= step_size * (mv / denom);
}

// KEEP = float4 per lane of a decoder row held in registers between the update and the renorm
// (row_len <= 128 * KEEP); the launcher picks the smallest that fits: a 16-deep predicated loop costs
// a d = 384 row more than it saves (0.022 -> 0.030 ms at 384 -> 3072 when it was the only form).
template <int KEEP>
__global__ void __launch_bounds__(256)
adamw_multi_kernel(const AdamwBatch batch, const float* __restrict__ hyper,
                   const double* __restrict__ grad_sumsq, float renorm_eps) {
  pdl_prologue();
  const float lr = hyper[0], b1 = hyper[1], b2 = hyper[2], eps = hyper[3], wd = hyper[4];
  const float bc1 = hyper[5], bc2s = hyper[6], max_norm = hyper[7];
  float clip = 1.f;
  if (grad_sumsq != nullptr && max_norm > 0.f) {
    const float total = static_cast<float>(sqrt(*grad_sumsq));
    const float c = max_norm / (total + 1e-6f);
    clip = c < 1.f ? c : 1.f;
  }
  const float step_size = lr / bc1;
  const long long total_units = batch.unit_start[batch.count];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (long long u = blockIdx.x; u < total_units; u += gridDim.x) {
    int ti = 0;
    while (ti + 1 < batch.count && u >= batch.unit_start[ti + 1]) ++ti;
    const AdamwTensor& T = batch.t[ti];
    const long long lu = u - batch.unit_start[ti];
    if (T.row_len == 0) {
      const long long i = lu * 1024 + threadIdx.x * 4;
      if (i + 3 < T.n && ((reinterpret_cast<uintptr_t>(T.p) | reinterpret_cast<uintptr_t>(T.g) |
                           reinterpret_cast<uintptr_t>(T.m) | reinterpret_cast<uintptr_t>(T.v)) & 15u) == 0) {
        float4 pv = *reinterpret_cast<float4*>(T.p + i);
        const float4 g = *reinterpret_cast<const float4*>(T.g + i);
        float4 mv = *reinterpret_cast<float4*>(T.m + i);
        float4 vv = *reinterpret_cast<float4*>(T.v + i);
        adamw_elem(pv.x, g.x * clip, mv.x, vv.x, lr, b1, b2, eps, wd, bc2s, step_size);
        adamw_elem(pv.y, g.y * clip, mv.y, vv.y, lr, b1, b2, eps, wd, bc2s, step_size);
        adamw_elem(pv.z, g.z * clip, mv.z, vv.z, lr, b1, b2, eps, wd, bc2s, step_size);
        adamw_elem(pv.w, g.w * clip, mv.w, vv.w, lr, b1, b2, eps, wd, bc2s, step_size);
        *reinterpret_cast<float4*>(T.p + i) = pv;
        *reinterpret_cast<float4*>(T.m + i) = mv;
        *reinterpret_cast<float4*>(T.v + i) = vv;
      } else {
        for (long long j = i; j < T.n && j < i + 4; ++j) {
          float pv = T.p[j], mv = T.m[j], vv = T.v[j];
          adamw_elem(pv, T.g[j] * clip, mv, vv, lr, b1, b2, eps, wd, bc2s, step_size);
          T.p[j] = pv; T.m[j] = mv; T.v[j] = vv;
        }
      }
    } else {
      const long long rows = T.n / T.row_len;
      const long long row = lu * 8 + warp;
      if (row < rows) {
        const long long base = row * T.row_len;
        const bool project = (T.pad_ & 1) != 0;
        constexpr int kKeep = KEEP;
        const int nv = T.row_len >> 2;
        const bool vec = (T.row_len & 3) == 0 && nv <= 32 * kKeep &&
                         ((reinterpret_cast<uintptr_t>(T.p) | reinterpret_cast<uintptr_t>(T.g) |
                           reinterpret_cast<uintptr_t>(T.m) | reinterpret_cast<uintptr_t>(T.v)) & 15u) == 0;
        // optional gradient projection (north_star; NOT in the reference, off by default): remove
        // the component of the row's gradient along the (unit-norm) decoder row, g -= (g.w) w, so
        // the step does not spend itself on a length change the renormalisation undoes anyway
        float gw = 0.f;
        if (project) {
          float ww = 0.f;
          for (int c = lane; c < T.row_len; c += 32) {
            const float pv = T.p[base + c];
            gw = fmaf(T.g[base + c], pv, gw);
            ww = fmaf(pv, pv, ww);
          }
          gw = warp_sum(gw) / fmaxf(warp_sum(ww), 1e-30f);
        }
        float ss = 0.f;
        if (vec) {
          // ONE pass over the row: 16-byte accesses, updated weights stay in registers until the
          // row norm is known (the scalar form re-read them: 0.68 ms -> see DESIGN for large-v3)
          float4 keep[kKeep];
#pragma unroll
          for (int i = 0; i < kKeep; ++i) {
            const int c4 = lane + i * 32;
            if (c4 < nv) {
              float4 pv = *reinterpret_cast<const float4*>(T.p + base + c4 * 4);
              float4 g = *reinterpret_cast<const float4*>(T.g + base + c4 * 4);
              float4 mv = *reinterpret_cast<const float4*>(T.m + base + c4 * 4);
              float4 vv = *reinterpret_cast<const float4*>(T.v + base + c4 * 4);
              if (project) { g.x -= gw * pv.x; g.y -= gw * pv.y; g.z -= gw * pv.z; g.w -= gw * pv.w; }
              adamw_elem(pv.x, g.x * clip, mv.x, vv.x, lr, b1, b2, eps, wd, bc2s, step_size);
              adamw_elem(pv.y, g.y * clip, mv.y, vv.y, lr, b1, b2, eps, wd, bc2s, step_size);
              adamw_elem(pv.z, g.z * clip, mv.z, vv.z, lr, b1, b2, eps, wd, bc2s, step_size);
              adamw_elem(pv.w, g.w * clip, mv.w, vv.w, lr, b1, b2, eps, wd, bc2s, step_size);
              *reinterpret_cast<float4*>(T.m + base + c4 * 4) = mv;
              *reinterpret_cast<float4*>(T.v + base + c4 * 4) = vv;
              keep[i] = pv;
              ss = fmaf(pv.x, pv.x, ss); ss = fmaf(pv.y, pv.y, ss);
              ss = fmaf(pv.z, pv.z, ss); ss = fmaf(pv.w, pv.w, ss);
            }
          }
          ss = warp_sum(ss);
          const float nrm = fmaxf(sqrtf(ss), renorm_eps);     // x / max(||x||, eps): F.normalize
#pragma unroll
          for (int i = 0; i < kKeep; ++i) {
            const int c4 = lane + i * 32;
            if (c4 < nv) {
              float4 pv = keep[i];
              pv.x /= nrm; pv.y /= nrm; pv.z /= nrm; pv.w /= nrm;
              *reinterpret_cast<float4*>(T.p + base + c4 * 4) = pv;
            }
          }
        } else {
          // pass 1: update, store moments, keep the un-normalised weight in place
          for (int c = lane; c < T.row_len; c += 32) {
            float pv = T.p[base + c], mv = T.m[base + c], vv = T.v[base + c];
            float g = T.g[base + c];
            if (project) g -= gw * pv;
            adamw_elem(pv, g * clip, mv, vv, lr, b1, b2, eps, wd, bc2s, step_size);
            T.p[base + c] = pv; T.m[base + c] = mv; T.v[base + c] = vv;
            ss = fmaf(pv, pv, ss);
          }
          ss = warp_sum(ss);
          const float nrm = fmaxf(sqrtf(ss), renorm_eps);
          // pass 2 (same lane touches the same addresses: L1-resident)
          for (int c = lane; c < T.row_len; c += 32) T.p[base + c] = T.p[base + c] / nrm;
        }
      }
    }
  }
}

}  // namespace wsae

using namespace wsae;

// Host descriptor mirrored in include/wsae.h (wsae_adamw_tensor_t).
struct wsae_adamw_tensor_host {
  float* p;
  const float* g;
  float* m;
  float* v;
  long long n;
  int row_len;
  int reserved;
};

extern "C" int wsae_adamw_multi(const wsae_adamw_tensor_host* tensors, int count, const float* hyper,
                                const double* grad_sumsq, float renorm_eps, cudaStream_t stream) {
  if (!tensors || !hyper || count <= 0 || count > kAdamwMaxTensors) return kBadArg;
  AdamwBatch b;
  long long units = 0;
  for (int i = 0; i < count; ++i) {
    const wsae_adamw_tensor_host& h = tensors[i];
    if (!h.p || !h.g || !h.m || !h.v || h.n <= 0 || h.row_len < 0) return kBadArg;
    if (h.row_len > 0 && h.n % h.row_len != 0) return kBadArg;
    b.t[i] = AdamwTensor{h.p, h.g, h.m, h.v, h.n, h.row_len, h.reserved};
    b.unit_start[i] = units;
    units += h.row_len == 0 ? (h.n + 1023) / 1024 : (h.n / h.row_len + 7) / 8;
  }
  b.unit_start[count] = units;
  b.count = count;
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long cap = static_cast<long long>(sms) * 16;
  const unsigned blocks = static_cast<unsigned>(units < cap ? units : cap);
  int max_row = 0;
  for (int i = 0; i < count; ++i) max_row = tensors[i].row_len > max_row ? tensors[i].row_len : max_row;
  if (max_row <= 512)
    launch_pdl(adamw_multi_kernel<4>, blocks, 256, 0, stream, b, hyper, grad_sumsq, renorm_eps);
  else if (max_row <= 1024)
    launch_pdl(adamw_multi_kernel<8>, blocks, 256, 0, stream, b, hyper, grad_sumsq, renorm_eps);
  else
    launch_pdl(adamw_multi_kernel<16>, blocks, 256, 0, stream, b, hyper, grad_sumsq, renorm_eps);
  return static_cast<int>(cudaGetLastError());
}

extern "C" int wsae_renorm_decoder(float* w_decT, int F, int d, float eps, void* bf16_shadow,
                                   cudaStream_t stream) {
  if (!w_decT || F <= 0 || d <= 0) return kBadArg;
  const int warps = 8;
  renorm_rows_kernel<<<ceil_div(F, warps), warps * 32, 0, stream>>>(
      w_decT, F, d, eps, static_cast<__nv_bfloat16*>(bf16_shadow));
  return static_cast<int>(cudaGetLastError());
}

extern "C" int wsae_counters_update(const long long* last_activated, long long* step_count, int F,
                                    long long threshold, int bump, long long* dead_count,
                                    cudaStream_t stream) {
  if (!last_activated || !step_count || F <= 0) return kBadArg;
  launch_pdl(counters_update_kernel, 1, 1024, 0, stream, last_activated, step_count, F, threshold,
             bump, dead_count, nullptr, nullptr, nullptr);
  return static_cast<int>(cudaGetLastError());
}

// wsae_counters_update + the step's metrics posted to a host mailbox (see the kernel's comment).
// stats2: device {sse f64, l0 u64} (nullable: zeros are posted); seq: device int64; mailbox: 4 x
// int64 of page-locked host memory, device-accessible at the same address (UVA).
extern "C" int wsae_counters_update_post(const long long* last_activated, long long* step_count,
                                         int F, long long threshold, int bump,
                                         long long* dead_count, const long long* stats2,
                                         const long long* seq, long long* mailbox,
                                         cudaStream_t stream) {
  if (!last_activated || !step_count || F <= 0 || !seq || !mailbox) return kBadArg;
  launch_pdl(counters_update_kernel, 1, 1024, 0, stream, last_activated, step_count, F, threshold,
             bump, dead_count, stats2, seq, mailbox);
  return static_cast<int>(cudaGetLastError());
}

extern "C" int wsae_densify_hidden(const int32_t* idx, const float* val, int B, int F, int k,
                                   float* hidden, cudaStream_t stream) {
  if (!idx || !val || !hidden || B <= 0 || F <= 0 || k <= 0) return kBadArg;
  densify_kernel<<<B, 256, 0, stream>>>(idx, val, B, F, k, hidden);
  return static_cast<int>(cudaGetLastError());
}

extern "C" int wsae_cast_bf16(const float* src, void* dst, long long n, cudaStream_t stream) {
  if (!src || !dst || n <= 0) return kBadArg;
  const size_t groups = (static_cast<size_t>(n) + 3) / 4;
  const unsigned blocks = static_cast<unsigned>((groups + 255) / 256);
  launch_pdl(cast_bf16_kernel, blocks, 256, 0, stream, src, static_cast<__nv_bfloat16*>(dst),
             static_cast<size_t>(n));
  return static_cast<int>(cudaGetLastError());
}

extern "C" int wsae_sumsq(const float* g, long long n, double* out, cudaStream_t stream) {
  if (!g || !out || n <= 0) return kBadArg;
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  size_t want = (static_cast<size_t>(n) / 4 + 255) / 256;
  if (want < 1) want = 1;
  const size_t cap = static_cast<size_t>(sms) * 8;
  const unsigned blocks = static_cast<unsigned>(want < cap ? want : cap);
  launch_pdl(sumsq_kernel, blocks, 256, 0, stream, g, static_cast<size_t>(n), out);
  return static_cast<int>(cudaGetLastError());
}

extern "C" int wsae_fused_adamw(float* p, const float* grad, float* m, float* v, long long n,
                                const float* hyper, const double* grad_sumsq,
                                cudaStream_t stream) {
  if (!p || !grad || !m || !v || !hyper || n <= 0) return kBadArg;
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  size_t want = (static_cast<size_t>(n) + 255) / 256;
  const size_t cap = static_cast<size_t>(sms) * 16;
  const unsigned blocks = static_cast<unsigned>(want < cap ? want : cap);
  fused_adamw_kernel<<<blocks, 256, 0, stream>>>(p, grad, m, v, static_cast<size_t>(n), hyper,
                                                 grad_sumsq);
  return static_cast<int>(cudaGetLastError());
}

// *out += sum(g^2), bit-reproducible: per-block partials in ws (>= wsae_sumsq_det_blocks() doubles),
// summed in block order by a second kernel (wsae_sumsq adds them with atomics in arrival order).
extern "C" int wsae_sumsq_det_blocks(void) { return 1024; }
extern "C" int wsae_sumsq_det(const float* g, long long n, double* ws, double* out, cudaStream_t stream) {
  if (!g || !ws || !out || n <= 0) return kBadArg;
  size_t want = (static_cast<size_t>(n) / 4 + 255) / 256;
  if (want < 1) want = 1;
  const unsigned blocks = static_cast<unsigned>(want < 1024 ? want : 1024);   // fixed cap: same split on every device
  sumsq_partial_kernel<<<blocks, 256, 0, stream>>>(g, static_cast<size_t>(n), ws);
  sumsq_ordered_kernel<<<1, 32, 0, stream>>>(ws, static_cast<int>(blocks), out);
  return static_cast<int>(cudaGetLastError());
}
