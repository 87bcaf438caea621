// wsae_feature_topk.cu — per-feature top-K activating examples, kept on the device.
//
// Consumer of the hot path's sparse code (idx, val): replaces the Python heap loop of the
// reference's analysis/feature_viz.py
//   :94-158   TopKTracker.update   (for every sample, position and active feature: heappush /
//             heapreplace on a per-feature min-heap of (activation, example))
//   :425-484  collect_top_activations (encode every batch, feed the tracker)
// with four kernels per batch of entries (row, feature, value):
//   count  - entries with value > 0 are counted (total_activations, :120) and those that can still
//            enter their feature's list (list not full, or value above its K-th value) are counted
//            per feature;
//   scan   - exclusive prefix sum over the F per-feature counts;
//   fill   - the same entries are written as 64-bit sort keys into their feature's segment;
//   merge  - one warp per feature with candidates merges them into the feature's K-entry list.
// State per feature (caller-owned): top_val[K] descending (-inf = empty), top_sample[K] (int64),
// top_pos[K] (int32), top_count.
//
// Ordering rule = the heap's (:150-153): an entry replaces the current K-th only if STRICTLY
// larger, so among equal values the earlier arrival stays.  Arrival order within a batch is the
// row order (a feature occurs at most once per row).  The sort key encodes that:
//   key = order-preserving image of the value << 32 | tag,   tag(existing slot s) = 2^31 | (K-1-s)
//                                                            tag(batch row r)     = 2^31-1 - r
// so existing entries beat every new entry of equal value, and earlier rows beat later ones.
// Keys are unique, which makes the result independent of the order the fill kernel's atomics land.
//
// Byte/index work, HBM- and atomic-bound: 2 coalesced passes over the (feature, value) arrays
// (8 bytes per entry each), F-sized counters in L2, 8 bytes per surviving candidate.
#include "wsae_common.cuh"

namespace wsae {

constexpr int kTrackMaxK = 32;       // one list slot per lane of the merging warp
constexpr int kTrackPer = 8;         // candidates per lane per merge round (256 per round)

__device__ __forceinline__ bool ftk_qualifies(float v, int f, int K, const float* __restrict__ top_val,
                                              const int32_t* __restrict__ top_count) {
  return top_count[f] < K || v > top_val[static_cast<size_t>(f) * K + (K - 1)];
}

__global__ void __launch_bounds__(256)
ftk_count_kernel(const int32_t* __restrict__ feat, const float* __restrict__ val, long long n, int F,
                 int K, const float* __restrict__ top_val, const int32_t* __restrict__ top_count,
                 int32_t* __restrict__ cand_count, unsigned long long* __restrict__ total) {
  unsigned int positives = 0;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int f = feat[i];
    const float v = val[i];
    if (f < 0 || f >= F || !(v > 0.f)) continue;       // padded slots carry idx = -1; NaN never fires
    ++positives;
    if (ftk_qualifies(v, f, K, top_val, top_count)) atomicAdd(&cand_count[f], 1);
  }
  positives = __reduce_add_sync(0xffffffffu, positives);
  if ((threadIdx.x & 31) == 0 && positives) atomicAdd(total, static_cast<unsigned long long>(positives));
}

// Exclusive scan of cand_count[0..F) into cand_off[0..F]; one block.
__global__ void __launch_bounds__(1024)
ftk_scan_kernel(const int32_t* __restrict__ cand_count, int F, int32_t* __restrict__ cand_off) {
  __shared__ int warp_sums[32];
  __shared__ int carry;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < F; base += blockDim.x) {
    const int i = base + threadIdx.x;
    const int c = i < F ? cand_count[i] : 0;
    int incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      int w = warp_sums[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += t;
      }
      warp_sums[lane] = w;       // inclusive over warps
    }
    __syncthreads();
    const int before = carry + (warp ? warp_sums[warp - 1] : 0) + incl - c;
    if (i < F) cand_off[i] = before;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry = before + c;
    __syncthreads();
  }
  if (threadIdx.x == 0) cand_off[F] = carry;
}

__global__ void __launch_bounds__(256)
ftk_fill_kernel(const int32_t* __restrict__ feat, const float* __restrict__ val,
                const int32_t* __restrict__ rows, long long n, int k, int F, int K,
                const float* __restrict__ top_val, const int32_t* __restrict__ top_count,
                const int32_t* __restrict__ cand_off, int32_t* __restrict__ cand_fill,
                unsigned long long* __restrict__ keys) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int f = feat[i];
    const float v = val[i];
    if (f < 0 || f >= F || !(v > 0.f)) continue;
    if (!ftk_qualifies(v, f, K, top_val, top_count)) continue;
    const uint32_t row = rows ? static_cast<uint32_t>(rows[i]) : static_cast<uint32_t>(i / k);
    const int p = cand_off[f] + atomicAdd(&cand_fill[f], 1);
    keys[p] = (static_cast<unsigned long long>(f2key(v)) << 32) | (0x7fffffffu - row);
  }
}

__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long m) {
  const uint32_t hi = static_cast<uint32_t>(m >> 32);
  const uint32_t whi = __reduce_max_sync(0xffffffffu, hi);
  const uint32_t lo = (hi == whi) ? static_cast<uint32_t>(m) : 0u;
  const uint32_t wlo = __reduce_max_sync(0xffffffffu, lo);
  return (static_cast<unsigned long long>(whi) << 32) | wlo;
}

// One warp per feature.  Lane s holds list slot s (key 0 = empty; every valid key is > 0 because
// the value part of a positive float's key has its top bit set).
__global__ void __launch_bounds__(256)
ftk_merge_kernel(int F, int K, const int32_t* __restrict__ cand_count,
                 const int32_t* __restrict__ cand_off, const unsigned long long* __restrict__ keys,
                 const long long* __restrict__ sample_ids, long long sample_base,
                 const int32_t* __restrict__ pos_ids, float* __restrict__ top_val,
                 long long* __restrict__ top_sample, int32_t* __restrict__ top_pos,
                 int32_t* __restrict__ top_count) {
  const int lane = threadIdx.x & 31;
  const int f = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (f >= F) return;
  const int c = cand_count[f];
  if (c == 0) return;
  const size_t lbase = static_cast<size_t>(f) * K;
  const int have = top_count[f];
  unsigned long long cur = 0;
  if (lane < have)
    cur = (static_cast<unsigned long long>(f2key(top_val[lbase + lane])) << 32) |
          (0x80000000u | static_cast<uint32_t>(K - 1 - lane));
  const unsigned long long* ck_src = keys + cand_off[f];
  for (int base = 0; base < c; base += 32 * kTrackPer) {
    unsigned long long ck[kTrackPer];
#pragma unroll
    for (int t = 0; t < kTrackPer; ++t) {
      const int j = base + t * 32 + lane;
      ck[t] = j < c ? ck_src[j] : 0ull;
    }
    unsigned long long next = 0;
    for (int r = 0; r < K; ++r) {
      unsigned long long m = cur;
#pragma unroll
      for (int t = 0; t < kTrackPer; ++t) m = ck[t] > m ? ck[t] : m;
      const unsigned long long wm = warp_max_u64(m);
      if (wm == 0ull) break;
      if (cur == wm) {
        cur = 0;
      } else {
#pragma unroll
        for (int t = 0; t < kTrackPer; ++t) ck[t] = (ck[t] == wm) ? 0ull : ck[t];
      }
      if (lane == r) next = wm;
    }
    cur = next;
  }
  // payloads first (they may come from list slots other lanes are about to overwrite)
  long long s_out = -1;
  int p_out = -1;
  float v_out = __uint_as_float(0xff800000u);
  if (cur != 0ull) {
    const uint32_t tag = static_cast<uint32_t>(cur);
    v_out = key2f(static_cast<uint32_t>(cur >> 32));
    if (tag & 0x80000000u) {
      const int slot = K - 1 - static_cast<int>(tag & 0x7fffffffu);
      s_out = top_sample[lbase + slot];
      p_out = top_pos[lbase + slot];
    } else {
      const uint32_t row = 0x7fffffffu - tag;
      s_out = sample_ids ? sample_ids[row] : sample_base + static_cast<long long>(row);
      p_out = pos_ids ? pos_ids[row] : 0;
    }
  }
  const uint32_t filled = __ballot_sync(0xffffffffu, cur != 0ull);
  __syncwarp();
  if (lane < K) {
    top_val[lbase + lane] = v_out;
    top_sample[lbase + lane] = s_out;
    top_pos[lbase + lane] = p_out;
  }
  if (lane == 0) top_count[f] = __popc(filled);
}

}  // namespace wsae

using namespace wsae;

// See include/wsae.h for the contract.
extern "C" int wsae_feature_topk_workspace(long long n, int F, unsigned long long* bytes) {
  if (n < 0 || F <= 0 || !bytes) return kBadArg;
  // cand_count[F] | cand_fill[F] | cand_off[F+1] (int32, padded to 8 bytes) | keys[n] (u64)
  const unsigned long long ints = (3ull * static_cast<unsigned long long>(F) + 1ull + 1ull) & ~1ull;
  *bytes = ints * 4ull + static_cast<unsigned long long>(n) * 8ull;
  return kOk;
}

extern "C" int wsae_feature_topk_update(const int32_t* feat, const float* val, const int32_t* rows,
                                        long long n, int k, const long long* sample_ids,
                                        long long sample_base, const int32_t* pos_ids, int F, int K,
                                        float* top_val, long long* top_sample, int32_t* top_pos,
                                        int32_t* top_count, unsigned long long* total, void* ws,
                                        unsigned long long ws_bytes, cudaStream_t stream) {
  if (n < 0 || F <= 0 || K <= 0 || k <= 0) return kBadArg;
  if (K > kTrackMaxK) return kUnsupported;
  if (!top_val || !top_sample || !top_pos || !top_count || !total || !ws) return kBadArg;
  if (n == 0) return kOk;
  if (!feat || !val) return kBadArg;
  if (n > 0x7fffffffLL) return kUnsupported;      // row tags and segment offsets are 31-bit
  unsigned long long need = 0;
  wsae_feature_topk_workspace(n, F, &need);
  if (ws_bytes < need) return kBadArg;
  int32_t* cand_count = static_cast<int32_t*>(ws);
  int32_t* cand_fill = cand_count + F;
  int32_t* cand_off = cand_fill + F;
  const unsigned long long ints = (3ull * static_cast<unsigned long long>(F) + 2ull) & ~1ull;
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(cand_count + ints);
  cudaError_t e = cudaMemsetAsync(cand_count, 0, sizeof(int32_t) * 2 * static_cast<size_t>(F), stream);
  if (e != cudaSuccess) return static_cast<int>(e);
  int dev = 0, num_sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  const long long want = (n + 255) / 256;
  const int grid = static_cast<int>(want < 8LL * num_sms ? want : 8LL * num_sms);
  ftk_count_kernel<<<grid, 256, 0, stream>>>(feat, val, n, F, K, top_val, top_count, cand_count, total);
  ftk_scan_kernel<<<1, 1024, 0, stream>>>(cand_count, F, cand_off);
  ftk_fill_kernel<<<grid, 256, 0, stream>>>(feat, val, rows, n, k, F, K, top_val, top_count, cand_off,
                                            cand_fill, keys);
  ftk_merge_kernel<<<ceil_div(F, 8), 256, 0, stream>>>(F, K, cand_count, cand_off, keys, sample_ids,
                                                       sample_base, pos_ids, top_val, top_sample,
                                                       top_pos, top_count);
  return static_cast<int>(cudaGetLastError());
}
