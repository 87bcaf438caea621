// wsae_wgrad_gemm.cu — K4: the two weight-gradient GEMMs of the TopK-SAE backward on tcgen05.
//
// Autograd of the reference's two Linear layers (model.py:111 encoder, :129 decoder) produces
//   dW_enc [F,d]  = dpre^T   . xc      (dpre  [B,F]: k nonzeros per row, values dv)
//   dW_decT[F,d]  = hidden^T . g       (hidden[B,F]: same pattern,       values relu(v))
// as dense [F,B]x[B,d] products over >= 99 %-zero operands (training.py:184).  Both are
//   OUT[F, d] (+)= alpha * S^T . R,   S = k-sparse [B,F] given as bucketed entries, R = bf16 [B,d]
// and are computed here as a tensor-core GEMM with M = 128 features, N = d, K = batch rows:
//   * R tiles ([64 rows] x [n_tile cols], MN-major for the MMA) arrive by TMA as n_tile/64
//     swizzled 64x64 blocks on one mbarrier;
//   * the sparse operand is expanded on the fly: for every (feature tile, 64-row chunk) the builder
//     warps zero a 128x64 bf16 tile in shared memory and scatter that cell's few entries into it,
//     in the SWIZZLE_128B K-major layout the MMA expects — S is never dense in HBM;
//   * fp32 accumulators live in TMEM; split-K CTAs combine with red.global.add.f32;
//   * the CTAs of a thread-block cluster work on different feature tiles of the SAME row chunks, so
//     they need the same R tiles: each CTA fetches 1/c of every R tile and TMA-multicasts it into
//     all c CTAs' shared memory (L2 -> SM traffic / c; without it every one of the F/128 feature
//     tiles streams all of R from L2: 24 x 58 MB per launch at the bench shape).  A stage is
//     reusable when the MMAs of ALL c CTAs have read it: the commit is multicast to every CTA's
//     `empty` barrier, which counts c arrivals.
// The bucketing (entries grouped by feature tile, then by row chunk) is three small kernels below.
#include <cuda.h>
#include <cudaTypedefs.h>

#include "wsae_common.cuh"

namespace wsae {

constexpr int kWgFeat = 128;   // features per MMA tile (M, TMEM lanes)
constexpr int kWgRows = 64;    // batch rows per pipeline stage (K per stage)
constexpr int kWgATile = kWgFeat * kWgRows * 2;   // 16 KB
constexpr int kWgRBlock = kWgRows * 64 * 2;       // one 64x64 bf16 block: 8 KB
constexpr int kWgMaxNB = 8;                       // n_tile <= 512 columns (TMEM)

// ------------------------------------------------------------------------------------------------
// bucketing: entries of (idx,val) grouped by (feature tile, row chunk)
// ------------------------------------------------------------------------------------------------
// One block per 64-row chunk: count the chunk's active entries (value > 0) per feature tile in
// shared memory, scan, and write them grouped by feature tile into the chunk's own fixed segment
// [chunk * 64 * k, ...) of the entry arrays.  No global scan: a chunk holds at most 64 * k entries.
//   offsets[chunk * (n_ft + 1) + ft] = absolute index of the first entry of cell (chunk, ft)
//   meta = row_local (0..63) | f_local << 8 ; va = dv ; vb = relu(val)
__global__ void __launch_bounds__(256)
bucket_chunk_kernel(const int32_t* __restrict__ idx, const float* __restrict__ val,
                    const float* __restrict__ dpre, int B, int F, int k, int n_ft,
                    int* __restrict__ offsets, uint32_t* __restrict__ ent_meta,
                    float* __restrict__ ent_a, float* __restrict__ ent_b) {
  extern __shared__ int s_cell[];  // [n_ft] counts -> cursors
  pdl_prologue();
  const int chunk = blockIdx.x;
  for (int i = threadIdx.x; i < n_ft; i += blockDim.x) s_cell[i] = 0;
  __syncthreads();
  const int row0 = chunk * kWgRows;
  const int nrow = min(kWgRows, B - row0);
  const size_t g0 = static_cast<size_t>(row0) * k;
  for (int e = threadIdx.x; e < nrow * k; e += blockDim.x) {
    const int32_t f = idx[g0 + e];
    const float v = val[g0 + e];
    if (f >= 0 && f < F && v > 0.f) atomicAdd(&s_cell[f / kWgFeat], 1);
  }
  __syncthreads();
  if (threadIdx.x < 32) {   // exclusive scan over the feature tiles, 32 at a time
    const int lane = threadIdx.x;
    int carry = static_cast<int>(g0);
    int* orow = offsets + static_cast<size_t>(chunk) * (n_ft + 1);
    for (int base = 0; base < n_ft; base += 32) {
      const int i = base + lane;
      const int c = i < n_ft ? s_cell[i] : 0;
      int x = c;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
      }
      if (i < n_ft) {
        s_cell[i] = carry + x - c;
        orow[i] = carry + x - c;
      }
      carry += __shfl_sync(0xffffffffu, x, 31);
    }
    if (lane == 0) orow[n_ft] = carry;
  }
  __syncthreads();
  // grouped entries are staged in shared memory and copied out with coalesced stores (the cursor
  // positions are scattered over the chunk's 64 * k slots: direct global stores touched one sector
  // per lane)
  uint32_t* s_meta = reinterpret_cast<uint32_t*>(s_cell + n_ft);
  float* s_a = reinterpret_cast<float*>(s_meta + kWgRows * k);
  float* s_b = s_a + kWgRows * k;
  for (int e = threadIdx.x; e < nrow * k; e += blockDim.x) {
    const int32_t f = idx[g0 + e];
    const float v = val[g0 + e];
    if (f >= 0 && f < F && v > 0.f) {
      const int ft = f / kWgFeat;
      const int pos = atomicAdd(&s_cell[ft], 1) - static_cast<int>(g0);
      s_meta[pos] = static_cast<uint32_t>(e / k) | (static_cast<uint32_t>(f - ft * kWgFeat) << 8);
      s_a[pos] = dpre[g0 + e];
      s_b[pos] = v;
    }
  }
  __syncthreads();
  const int total = offsets[static_cast<size_t>(chunk) * (n_ft + 1) + n_ft] - static_cast<int>(g0);
  for (int e = threadIdx.x; e < total; e += blockDim.x) {
    ent_meta[g0 + e] = s_meta[e];
    ent_a[g0 + e] = s_a[e];
    ent_b[g0 + e] = s_b[e];
  }
}

// ------------------------------------------------------------------------------------------------
// the GEMM
// ------------------------------------------------------------------------------------------------
// MN-major bf16 operand stored as 64x64 blocks (rows = K index, 128 B each, SWIZZLE_128B), blocks
// along N spaced `block_bytes` apart: LBO = block stride (N direction), SBO = 8 rows = 1024 B.
__device__ __forceinline__ uint64_t umma_desc_sw128_mnmajor(uint32_t smem_addr, uint32_t block_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((block_bytes >> 4) & 0x3FFFu) << 16;   // LBO
  d |= static_cast<uint64_t>(1024 >> 4) << 32;                      // SBO
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
// kind::f16, A = bf16 K-major, B = bf16 MN-major, D = fp32
__host__ __device__ constexpr uint32_t umma_idesc_bf16_bmn(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | (static_cast<uint32_t>(N >> 3) << 17) |
         (static_cast<uint32_t>(M >> 4) << 24);
}

__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

template <int STAGES>
__global__ void __launch_bounds__(256, 1)
wgrad_gemm_kernel(const __grid_constant__ CUtensorMap tmap_r, int F, int d, int n_chunks, int n_ft,
                  int n_nt, int nb_tile, int ksplit, const int* __restrict__ offsets,
                  const uint32_t* __restrict__ ent_meta, const float* __restrict__ ent_val,
                  const float* __restrict__ grad_out, float alpha, float* __restrict__ out,
                  float* __restrict__ det_ws) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int stage_bytes = kWgATile + nb_tile * kWgRBlock;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * stage_bytes);
  uint64_t* r_full = bars;               // [STAGES]  TMA landed
  uint64_t* a_full = bars + STAGES;      // [STAGES]  sparse tile built (128 arrivals)
  uint64_t* empty = bars + 2 * STAGES;   // [STAGES]  MMAs that read the stage have retired
  uint64_t* acc_full = bars + 3 * STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * STAGES + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // cluster: consecutive blockIdx.x; its CTAs share (n tile, k split) and take adjacent feature tiles
  const int csize = static_cast<int>(cluster_nctarank());
  const int crank = static_cast<int>(cluster_ctarank());
  const uint16_t cmask = static_cast<uint16_t>((1u << csize) - 1u);
  // work item -> (feature tile, n tile, k split)
  int item = blockIdx.x / csize;
  const int ks = item % ksplit;
  item /= ksplit;
  const int nt = item % n_nt;
  const int ft = (item / n_nt) * csize + crank;
  const int per = ceil_div(n_chunks, ksplit);
  const int c0 = ks * per;
  const int c1 = min(n_chunks, c0 + per);
  const int nchunk = max(0, c1 - c0);
  const int nb0 = nt * nb_tile;                       // first 64-column block of this n tile
  const int nb_total = ceil_div(d, 64);
  const int nb_here = min(nb_tile, nb_total - nb0);   // blocks that exist
  const int n_cols = nb_here * 64;
  const uint32_t tmem_cols = n_cols > 256 ? 512u : (n_cols > 128 ? 256u : (n_cols > 64 ? 128u : 64u));

  pdl_launch_dependents();
  if (warp == 0 && lane == 0) tma_prefetch_desc(&tmap_r);
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&r_full[s], 1);
      mbar_init(&a_full[s], 128);
      mbar_init(&empty[s], static_cast<uint32_t>(csize));   // one commit arrival per CTA of the cluster
    }
    mbar_init(acc_full, 1);
    mbar_fence_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  if (csize > 1) cluster_sync_all();   // peers' barriers are initialised before anything is multicast
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();   // barriers, TMEM and the descriptor are set up; global memory only from here on

  if (warp == 0) {
    // ===================== TMA producer: R tiles =====================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int c = 0; c < nchunk; ++c) {
        mbar_wait(&empty[stage], phase ^ 1u);
        mbar_arrive_expect_tx(&r_full[stage], static_cast<uint32_t>(nb_here * kWgRBlock));
        uint8_t* dst = smem + stage * stage_bytes + kWgATile;
        if (csize == 1) {
          for (int b = 0; b < nb_here; ++b)   // one swizzled 64x64 block per bulk copy
            tma_load_2d(dst + b * kWgRBlock, &tmap_r, &r_full[stage], (nb0 + b) * 64, (c0 + c) * kWgRows);
        } else {
          // this CTA's share of the blocks, delivered to every CTA of the cluster (each one's r_full
          // expects the whole tile: the other blocks arrive from its peers)
          for (int b = crank; b < nb_here; b += csize)
            tma_load_2d_multicast(dst + b * kWgRBlock, &tmap_r, &r_full[stage], (nb0 + b) * 64,
                                  (c0 + c) * kWgRows, cmask);
        }
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const int n1 = min(256, n_cols);
    const int n2 = n_cols - n1;
    const uint32_t idesc1 = umma_idesc_bf16_bmn(kWgFeat, n1);
    const uint32_t idesc2 = umma_idesc_bf16_bmn(kWgFeat, n2 > 0 ? n2 : 16);
    int stage = 0;
    uint32_t phase = 0;
    for (int c = 0; c < nchunk; ++c) {
      mbar_wait(&r_full[stage], phase);
      mbar_wait(&a_full[stage], phase);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sa = smem_u32(smem + stage * stage_bytes);
        const uint32_t sr = sa + kWgATile;
        const uint64_t da = umma_desc_sw128_kmajor(sa);
        const uint64_t db1 = umma_desc_sw128_mnmajor(sr, kWgRBlock);
        const uint64_t db2 = umma_desc_sw128_mnmajor(sr + 4 * kWgRBlock, kWgRBlock);
#pragma unroll
        for (int kk = 0; kk < kWgRows / 16; ++kk) {
          const uint32_t acc = (c | kk) != 0 ? 1u : 0u;
          // A: +32 B per 16 K elements inside the swizzled 128 B row; R: +16 rows * 128 B
          umma_bf16(tmem_base, da + static_cast<uint64_t>(2 * kk),
                    db1 + static_cast<uint64_t>((kk * 16 * 128) >> 4), idesc1, acc);
          if (n2 > 0)
            umma_bf16(tmem_base + 256, da + static_cast<uint64_t>(2 * kk),
                      db2 + static_cast<uint64_t>((kk * 16 * 128) >> 4), idesc2, acc);
        }
        if (csize == 1) umma_commit(&empty[stage]);
        else umma_commit_multicast(&empty[stage], cmask);
        if (c == nchunk - 1) umma_commit(acc_full);
      }
      __syncwarp();
      if (++stage == STAGES) {
        stage = 0;
        phase ^= 1u;
      }
    }
  } else if (warp >= 4) {
    // ===================== sparse-tile builders, then epilogue =====================
    // The cell's entries come from global memory through two dependent loads (offsets -> entries,
    // ~2 L2 latencies).  They are software-pipelined two chunks deep - the offsets of chunk c+2 and
    // the entries of chunk c+1 are in flight while chunk c's tile is built - otherwise that latency
    // chain (~1500 cycles per chunk) paces the whole GEMM (tensor pipe 56 % active in round 1).
    const int t = threadIdx.x - 128;  // 0..127
    int stage = 0;
    uint32_t phase = 0;
    constexpr int kPre = 2;           // entries per thread held in registers (cells hold ~k*64*128/F)
    auto cell_of = [&](int c) { return static_cast<size_t>(c0 + c) * (n_ft + 1) + ft; };
    int2 off_next = make_int2(0, 0), off_next2 = make_int2(0, 0);
    if (nchunk > 0) off_next = make_int2(offsets[cell_of(0)], offsets[cell_of(0) + 1]);
    if (nchunk > 1) off_next2 = make_int2(offsets[cell_of(1)], offsets[cell_of(1) + 1]);
    uint32_t m_next[kPre];
    float v_next[kPre];
#pragma unroll
    for (int j = 0; j < kPre; ++j) {
      const int e = off_next.x + t + j * 128;
      m_next[j] = (nchunk > 0 && e < off_next.y) ? ent_meta[e] : 0u;
      v_next[j] = (nchunk > 0 && e < off_next.y) ? ent_val[e] : 0.f;
    }
    for (int c = 0; c < nchunk; ++c) {
      // rotate the pipeline registers: this chunk's data was loaded one iteration ago
      const int2 off = off_next;
      uint32_t m_cur[kPre];
      float v_cur[kPre];
#pragma unroll
      for (int j = 0; j < kPre; ++j) {
        m_cur[j] = m_next[j];
        v_cur[j] = v_next[j];
      }
      off_next = off_next2;
      if (c + 2 < nchunk) off_next2 = make_int2(offsets[cell_of(c + 2)], offsets[cell_of(c + 2) + 1]);
#pragma unroll
      for (int j = 0; j < kPre; ++j) {
        const int e = off_next.x + t + j * 128;
        const bool ok = (c + 1 < nchunk) && e < off_next.y;
        m_next[j] = ok ? ent_meta[e] : 0u;
        v_next[j] = ok ? ent_val[e] : 0.f;
      }

      mbar_wait(&empty[stage], phase ^ 1u);
      uint8_t* a_tile = smem + stage * stage_bytes;
      // zero 16 KB: consecutive threads clear consecutive 16-byte words (conflict-free)
      uint4* tile4 = reinterpret_cast<uint4*>(a_tile);
      const uint4 z = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
      for (int i = 0; i < 8; ++i) tile4[i * 128 + t] = z;
      named_bar_sync(1, 128);
      auto put = [&](uint32_t m, float v) {
        const uint32_t row = m & 0xFFu, fl = m >> 8;
        const uint32_t o = fl * 128u + ((((row >> 3) ^ (fl & 7u)) & 7u) << 4) + ((row & 7u) << 1);
        *reinterpret_cast<__nv_bfloat16*>(a_tile + o) = __float2bfloat16_rn(v);
      };
#pragma unroll
      for (int j = 0; j < kPre; ++j)
        if (off.x + t + j * 128 < off.y) put(m_cur[j], v_cur[j]);
      for (int e = off.x + t + kPre * 128; e < off.y; e += 128) put(ent_meta[e], ent_val[e]);   // crowded cell
      fence_proxy_async_smem();   // generic-proxy writes -> visible to the tensor core (async proxy)
      mbar_arrive(&a_full[stage]);
      if (++stage == STAGES) {
        stage = 0;
        phase ^= 1u;
      }
    }
    if (nchunk > 0) {
      mbar_wait(acc_full, 0);
      tc_fence_after();
      const float s = alpha * (grad_out != nullptr ? *grad_out : 1.f);
      const int q = warp - 4;
      const int f = ft * kWgFeat + q * 32 + lane;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
      for (int cb = 0; cb < n_cols; cb += 16) {
        uint32_t r[16];
        tmem_ld16(taddr + cb, r);
        tmem_ld_wait16(r);
        const int col0 = nb0 * 64 + cb;
        if (f < F && det_ws != nullptr) {
          // deterministic mode: this (feature tile, k split) partial goes to its own slab with plain
          // stores; wgrad_reduce_kernel adds the slabs to `out` in k-split order
          float* wrow = det_ws + (static_cast<size_t>(ks) * F + f) * d + col0;
#pragma unroll
          for (int j = 0; j < 16; j += 4)
            if (col0 + j + 3 < d)
              *reinterpret_cast<float4*>(wrow + j) =
                  make_float4(s * __uint_as_float(r[j]), s * __uint_as_float(r[j + 1]),
                              s * __uint_as_float(r[j + 2]), s * __uint_as_float(r[j + 3]));
            else
              for (int jj = j; jj < j + 4; ++jj)
                if (col0 + jj < d) wrow[jj] = s * __uint_as_float(r[jj]);
        } else if (f < F) {
          float* orow = out + static_cast<size_t>(f) * d + col0;
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            if (col0 + j + 3 < d) {
              const float a = s * __uint_as_float(r[j]), b = s * __uint_as_float(r[j + 1]);
              const float cc = s * __uint_as_float(r[j + 2]), dd = s * __uint_as_float(r[j + 3]);
              if (a != 0.f || b != 0.f || cc != 0.f || dd != 0.f) red_add_f32x4(orow + j, a, b, cc, dd);
            } else {
              for (int jj = j; jj < j + 4; ++jj)
                if (col0 + jj < d) atomicAdd(orow + jj, s * __uint_as_float(r[jj]));
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (csize > 1) cluster_sync_all();   // no CTA leaves while peers may still signal its barriers
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// out[i] += ws[0][i] + ws[1][i] + ... + ws[nslab-1][i], slabs of n floats, in that order.
__global__ void __launch_bounds__(256)
wgrad_reduce_kernel(const float* __restrict__ ws, int nslab, size_t n, float* __restrict__ out) {
  const size_t i = (static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    float4 acc = *reinterpret_cast<const float4*>(out + i);
    for (int sl = 0; sl < nslab; ++sl) {
      const float4 v = *reinterpret_cast<const float4*>(ws + static_cast<size_t>(sl) * n + i);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    *reinterpret_cast<float4*>(out + i) = acc;
  } else {
    for (size_t j = i; j < n; ++j) {
      float acc = out[j];
      for (int sl = 0; sl < nslab; ++sl) acc += ws[static_cast<size_t>(sl) * n + j];
      out[j] = acc;
    }
  }
}

static PFN_cuTensorMapEncodeTiled_v12000 wg_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return static_cast<PFN_cuTensorMapEncodeTiled_v12000>(nullptr);
    return reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }();
  return fn;
}

}  // namespace wsae

using namespace wsae;

static int g_wgrad_cluster = 2;   // upper bound on the cluster size (wsae_debug_wgrad_cluster: experiments)
extern "C" int wsae_debug_wgrad_cluster(int c) {
  if (c != 1 && c != 2 && c != 4 && c != 8) return kBadArg;
  g_wgrad_cluster = c;
  return kOk;
}

extern "C" int wsae_bucket_cells(int B, int F, int* n_chunks, int* n_ft) {
  if (B <= 0 || F <= 0) return kBadArg;
  if (n_chunks) *n_chunks = ceil_div(B, kWgRows);
  if (n_ft) *n_ft = ceil_div(F, kWgFeat);
  return kOk;
}

extern "C" int wsae_bucket_by_tile(const int32_t* idx, const float* val, const float* dpre, int B,
                                   int F, int k, int* offsets, uint32_t* ent_meta, float* ent_a,
                                   float* ent_b, cudaStream_t stream) {
  if (!idx || !val || !dpre || !offsets || !ent_meta || !ent_a || !ent_b) return kBadArg;
  if (B <= 0 || F <= 0 || k <= 0) return kBadArg;
  const int n_chunks = ceil_div(B, kWgRows), n_ft = ceil_div(F, kWgFeat);
  const size_t smem = static_cast<size_t>(n_ft) * sizeof(int) + static_cast<size_t>(kWgRows) * k * 12;
  if (smem > 200 * 1024) return kUnsupported;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(bucket_chunk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem));
    if (e != cudaSuccess) return static_cast<int>(e);
  }
  launch_pdl(bucket_chunk_kernel, n_chunks, 256, smem, stream, idx, val, dpre, B, F, k, n_ft, offsets,
             ent_meta, ent_a, ent_b);
  return static_cast<int>(cudaGetLastError());
}

// out[F,d] += alpha * (*grad_out) * S^T . R, S from the bucketed entries with values ent_val,
// R = bf16 [B rows, >= ceil(d/64)*64 columns] with row pitch r_pitch_elems.
// plan_only: just report the k-split count (workspace sizing), launch nothing.
static int wgrad_gemm_impl(const void* r_bf16, int r_pitch_elems, int B, int F, int d,
                           const int* offsets, const uint32_t* ent_meta, const float* ent_val,
                           const float* grad_out, float alpha, float* out, float* det_ws,
                           unsigned long long det_ws_bytes, int* plan_only, cudaStream_t stream) {
  if (!plan_only && (!r_bf16 || !offsets || !ent_meta || !ent_val || !out)) return kBadArg;
  if (B <= 0 || F <= 0 || d <= 0) return kBadArg;
  const int nb_total = ceil_div(d, 64);
  if (!plan_only && (r_pitch_elems < d || (r_pitch_elems % 8) != 0)) return kBadArg;
  if (!plan_only && (reinterpret_cast<uintptr_t>(r_bf16) & 15u) != 0) return kBadArg;
  const int n_chunks = ceil_div(B, kWgRows), n_ft = ceil_div(F, kWgFeat);
  const int n_nt = ceil_div(nb_total, kWgMaxNB);
  const int nb_tile = ceil_div(nb_total, n_nt);

  CUtensorMap tm;
  if (!plan_only) {
  auto fn = wg_encode_fn();
  if (!fn) return kNoDriver;
  // [B rows, d columns] bf16, row pitch r_pitch_elems; box = 64 rows x 64 columns; columns >= d and
  // rows >= B of a box are zero-filled by TMA
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(d), static_cast<cuuint64_t>(B)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(r_pitch_elems) * 2};
  cuuint32_t box[2] = {64, static_cast<cuuint32_t>(kWgRows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult cr = fn(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(r_bf16), gdim, gstride,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) return static_cast<int>(1000 + cr);
  }

  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int stage_bytes = kWgATile + nb_tile * kWgRBlock;
  const int stages = (3 * stage_bytes + 1024 + 256 <= 227 * 1024) ? 3 : 2;
  const int smem = stages * stage_bytes + 1024 + 256;
  if (smem > 227 * 1024) return kUnsupported;
  auto kern = stages == 3 ? wgrad_gemm_kernel<3> : wgrad_gemm_kernel<2>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return static_cast<int>(e);

  // cluster size: CTAs of a cluster share their R tiles by TMA multicast.  Largest of {g_wgrad_cluster,
  // 2, 1} that divides the feature-tile count and that the device can keep resident on (almost) all SMs.
  int csize = 1, resident = sms;
  for (int c = g_wgrad_cluster; c >= 2; c >>= 1) {
    if (n_ft % c != 0) continue;
    cudaLaunchConfig_t q = {};
    q.gridDim = dim3(static_cast<unsigned>(sms / c * c));
    q.blockDim = dim3(256);
    q.dynamicSmemBytes = static_cast<size_t>(smem);
    cudaLaunchAttribute qa[1];
    qa[0].id = cudaLaunchAttributeClusterDimension;
    qa[0].val.clusterDim.x = static_cast<unsigned>(c);
    qa[0].val.clusterDim.y = 1;
    qa[0].val.clusterDim.z = 1;
    q.attrs = qa;
    q.numAttrs = 1;
    int nclusters = 0;
    if (cudaOccupancyMaxActiveClusters(&nclusters, kern, &q) != cudaSuccess) {
      cudaGetLastError();
      continue;
    }
    if (nclusters * c >= (sms * 7) / 8) {   // keep >= 7/8 of the SMs busy
      csize = c;
      resident = nclusters * c;
      break;
    }
  }
  // split-K factor: fill the resident CTA slots, over up to three waves when one wave would leave
  // SMs idle (768 -> 6144: 96 (feature, n) tiles on 148 SMs = 65 % in one wave, 97 % as 3 x 96 CTAs)
  const int base_items = n_ft * n_nt;
  int ksplit = 1;
  double best_eff = 0.0;
  for (int waves = 1; waves <= 3; ++waves) {
    int ks = (resident * waves) / base_items;
    if (ks < 1) ks = 1;
    if (ks > n_chunks) ks = n_chunks;
    const int g = base_items * ks;
    const int w = (g + resident - 1) / resident;
    const double eff = static_cast<double>(g) / (static_cast<double>(w) * resident);
    if (eff > best_eff + 0.02) {     // prefer fewer waves (less split-K reduction traffic) on near ties
      best_eff = eff;
      ksplit = ks;
    }
  }
  const int per_split = ceil_div(n_chunks, ksplit);
  const int nslab = ceil_div(n_chunks, per_split);     // k splits that own at least one row chunk
  if (plan_only) {
    *plan_only = nslab;
    return kOk;
  }
  // deterministic mode: one partial slab per k split, then an ordered reduction.  A single split
  // already writes every output element from exactly one CTA (a RED onto the caller's value).
  const bool det = det_ws != nullptr && nslab > 1;
  if (det && det_ws_bytes < static_cast<unsigned long long>(nslab) * F * d * sizeof(float)) return kBadArg;
  if (det && d % 4 != 0) return kUnsupported;
  const int grid = n_ft * n_nt * ksplit;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(grid));
  cfg.blockDim = dim3(256);
  cfg.dynamicSmemBytes = static_cast<size_t>(smem);
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = static_cast<unsigned>(csize);
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = pdl_enabled();
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  e = cudaLaunchKernelEx(&cfg, kern, tm, F, d, n_chunks, n_ft, n_nt, nb_tile, ksplit, offsets, ent_meta,
                         ent_val, grad_out, alpha, out, det ? det_ws : static_cast<float*>(nullptr));
  if (e != cudaSuccess) return static_cast<int>(e);
  if (det) {
    const size_t n = static_cast<size_t>(F) * d;
    const unsigned blocks = static_cast<unsigned>((n / 4 + 255) / 256 + 1);
    wgrad_reduce_kernel<<<blocks, 256, 0, stream>>>(det_ws, nslab, n, out);
  }
  return static_cast<int>(cudaGetLastError());
}

extern "C" int wsae_wgrad_gemm(const void* r_bf16, int r_pitch_elems, int B, int F, int d,
                               const int* offsets, const uint32_t* ent_meta, const float* ent_val,
                               const float* grad_out, float alpha, float* out,
                               cudaStream_t stream) {
  return wgrad_gemm_impl(r_bf16, r_pitch_elems, B, F, d, offsets, ent_meta, ent_val, grad_out, alpha, out,
                         nullptr, 0, nullptr, stream);
}

// Deterministic form: the split-K partial sums go to `ws` (wsae_wgrad_gemm_workspace bytes) with
// plain stores and are added to `out` in k-split order by a second kernel, instead of red.global.add
// in arrival order.  Same result up to the summation order; bit-reproducible from run to run.
extern "C" int wsae_wgrad_gemm_workspace(int B, int F, int d, unsigned long long* bytes) {
  if (!bytes || B <= 0 || F <= 0 || d <= 0) return kBadArg;
  int nslab = 0;
  const int rc = wgrad_gemm_impl(nullptr, 0, B, F, d, nullptr, nullptr, nullptr, nullptr, 0.f, nullptr,
                                 nullptr, 0, &nslab, nullptr);
  if (rc) return rc;
  *bytes = nslab > 1 ? static_cast<unsigned long long>(nslab) * F * d * sizeof(float) : 0ull;
  return kOk;
}
extern "C" int wsae_wgrad_gemm_det(const void* r_bf16, int r_pitch_elems, int B, int F, int d,
                                   const int* offsets, const uint32_t* ent_meta, const float* ent_val,
                                   const float* grad_out, float alpha, float* out, float* ws,
                                   unsigned long long ws_bytes, cudaStream_t stream) {
  return wgrad_gemm_impl(r_bf16, r_pitch_elems, B, F, d, offsets, ent_meta, ent_val, grad_out, alpha, out,
                         ws, ws_bytes, nullptr, stream);
}
