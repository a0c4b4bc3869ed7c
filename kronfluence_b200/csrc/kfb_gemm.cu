// The tensor-core engine of libkfb: a persistent, warp-specialised, batched NT GEMM for sm_100a.
//
//   D[b] (128 x BLOCK_N fp32 tile per CTA in TMEM)  =  sum_k  A[b][m,k] * B[b][n,k]
//
//   warp 0 (one lane)  TMA producer: cp.async.bulk.tensor.3d loads of the hi/lo operand planes (bf16; scaled FP16 for
//                      KFB_PREC_STRICT) into a ring of 128B/64B-swizzled shared-memory stages (mbarrier complete_tx)
//   warp 1 (one lane)  MMA issuer: tcgen05.mma.kind::f16, K=16; cta_group::1 (M=128) or cta_group::2 CTA pairs (M=256,
//                      each CTA stages its own A rows and half of B); with two planes three MMAs per k-step
//                      (lo*hi + hi*lo + hi*hi) into the same fp32 accumulator; tcgen05.commit releases smem stages /
//                      publishes the accumulator
//   warp 2             TMEM allocator (2 accumulator stages: the epilogue of tile i overlaps the MMAs of tile i+1)
//   warps 4-7 (4-11)   epilogue: tcgen05.ld 32x32b (thread == accumulator row), then one of
//                        STORE   fp32 and/or bf16 hi/lo planes (TMA bulk stores for plain planes), optional
//                                transpose / square / elementwise factor / accumulate / split-K reductions /
//                                SYRK mirroring / per-example column groups
//                        ROWDOT  out[b,m] = sum_n D[m,n] g[m,n]   (fused pairwise-score epilogue:
//                                module/linear.py:112-122 "qio,bi,bo->qb" without the [T,d_out]
//                                intermediate), looping over all n-tiles of a (b, m-tile) unit
//                        REGACC  fp32 register accumulation across TMEM passes: either over a chunk
//                                of the batch with squaring, out[m,n] += sum_b D[b][m,n]^2 (Lambda
//                                sweep, tracker/factor.py:218-226), or over k-chunks of one long
//                                contraction (see kMaxPassK / kStrictPassK), finished by the STORE options;
//                                256-wide tiles take two epilogue warpgroups (setmaxnreg re-balances registers)
//   clusters           MC = 2: two CTA pairs on adjacent m-tiles fetch their common B tile once (TMA multicast)
//
// A second, trivially simple SIMT implementation of the same contract exists for debugging the
// host logic (kfb_set_gemm_backend(1)); it is never selected implicitly.
#include "kfb_gemm.cuh"

#include <atomic>
#include <mutex>

namespace kfb {

// Kernel-level epilogue families.  The public KFB_EPI_STORE maps to EPI_STORE when the contraction
// fits one TMEM accumulation pass and to EPI_REGACC (k-chunk mode) when it does not; KFB_EPI_SQACC
// maps to EPI_REGACC (batch mode).
enum : int { EPI_STORE = 0, EPI_ROWDOT = 1, EPI_REGACC = 2 };
// REGACC_BATCH_DOT: the registers hold an elementwise FACTOR instead of sums: out[b] += alpha * sum_{m,n} D_b[m,n]^2 factor[m,n]
// for every batch entry of the chunk (the self-influence contraction: the factor tile is read once per unit, not once
// per example and chunk, whose L2 latency used to be 16 serial round trips per tile)
// REGACC_BATCH_MUL: the same resident factor, but every batch entry is STORED: out[b][m,n] = alpha * D_b[m,n] factor[m,n]
// as operand planes (query preconditioning of S > 1 layers: P~_q = scale (G~_q) o Lambda^-1 straight into the store)
enum : int { REGACC_BATCH = 0, REGACC_KCHUNK = 1, REGACC_BATCH_DOT = 2, REGACC_BATCH_MUL = 3 };

// The tensor core adds each MMA's partial products into the fp32 TMEM accumulator with truncation,
// so a long accumulation chain acquires a relative bias of ~(K/16)*nsplit_mmas*2^-25.  One TMEM pass
// is therefore limited to kMaxPassK contraction elements (~8e-6 relative with the 3-MMA split);
// longer contractions are cut into passes that are summed in fp32 registers (round-to-nearest).
static const int kMaxPassK = 2048;
static const int kAtomicPassK = 1024;
static const int kStrictPassK = 128;
static std::atomic<int> g_strict_pass_k{kStrictPassK};  // debug knob (kfb_set_strict_pass_k): pass-length experiments

struct GemmParams {
  int M, N, K, batch;
  int a_batched, b_batched;
  int m_blocks, n_blocks, k_blocks;
  int kb_hint;                 // host only: k-blocks of 64 per unit (pass), to keep multicast off short passes
  int mc, m_units;             // CTA pairs per cluster sharing one B tile (TMA multicast); m_units = ceil(m_blocks / mc)
  int k_splits, kb_per_split;  // contraction split ACROSS CTAs (atomics)
  int k_chunks, kb_per_chunk;  // contraction passes INSIDE a unit (register accumulation)
  int inner;                   // REGACC_BATCH: batches per chunk
  int n_splits, nb_per_split;  // ROWDOT: the n-tiles of one row block spread over several units (atomics)
  int regacc_mode;
  int max_pass_k;              // longest TMEM accumulation chain (contraction elements) of this launch
  long long num_units;
  // outputs
  float* out_f32;
  long long ldo, out_bs;
  __nv_bfloat16* out_hi;
  __nv_bfloat16* out_lo;
  long long ldo_s, out_bs_s;
  const float* mul;
  long long ldmul;
  const float* factor;     // REGACC_BATCH_DOT / _MUL: the [M, N] factor kept in registers across the batch entries of a unit
  long long ldfactor;
  int transpose_out, square, accumulate, use_atomic, vec_ok, zero_pad;
  int reduce_sq;  // STORE: reduce D^2 (o mul) to one scalar per batch entry instead of storing
  float alpha;
  const float* g;
  long long ldg, g_bs;  // ROWDOT factor rows; g_bs = offset between the factors of consecutive batch entries
  int g_vec4;
  int row_group;        // ROWDOT: rows [j*row_group, (j+1)*row_group) are summed into out[b][j] (tokens of an example)
  int f32_vec4, mul_vec4;  // out_f32 rows / mul rows are 16-byte aligned
  int batch_fastest;       // STORE: consecutive units share the (m, n) tile (and so the `mul` tile) across the batch
  int tma_store;           // STORE to plain (untransposed, unfactored) planes: chunks leave through TMA bulk stores
  int f16;                 // strict operands: FP16 planes scaled by strict_scale(*absmax); the epilogue undoes it
  const float* a_absmax;
  const float* b_absmax;
  int col_group;           // STORE to planes: column n lands in batch entry n / col_group at column n % col_group
  unsigned int* sync_ctr;  // ROWDOT on clusters: one counter per group of `sync_group` clusters that stream the same
  int sync_group;          //   query's P tiles (adjacent m-units): they start every unit together (see the producer)
  int max_groups;          // host only: cap on resident CTA groups (the idle-SM side launch of the fused pairwise kernel)
  int symmetric;           // STORE: A == B (SYRK): only tiles with n_blk >= m_blk are computed, off-diagonal ones are
  long long tri_tiles;     //        also written transposed; tri_tiles = m_blocks (m_blocks + 1) / 2
};

// Counters of the cluster-group rendezvous of the fused pairwise kernel (zeroed before every launch that uses them).
__device__ unsigned int g_rowdot_sync[64];

struct Tile {
  int b, m_blk, n_blk, kb0, kb1;
};

template <int EPI>
__device__ __forceinline__ int unit_inner_count(const GemmParams& p, long long unit) {
  if (EPI == EPI_ROWDOT) {
    const int ns = (int)((unit / p.m_units) % p.n_splits);
    const int nb0 = ns * p.nb_per_split;
    return (min(p.n_blocks, nb0 + p.nb_per_split) - nb0) * p.k_chunks;
  }
  if (EPI == EPI_REGACC) {
    if (p.regacc_mode != REGACC_KCHUNK) {
      const long long chunk = unit / ((long long)p.m_units * p.n_blocks);
      const long long rem = (long long)p.batch - chunk * p.inner;
      return rem < p.inner ? (int)rem : p.inner;
    }
    const int ks = (int)((unit / ((long long)p.m_units * p.n_blocks)) % p.k_splits);
    const int kb0 = ks * p.kb_per_split;
    const int kb1 = min(p.k_blocks, kb0 + p.kb_per_split);
    return (kb1 - kb0 + p.kb_per_chunk - 1) / p.kb_per_chunk;
  }
  return 1;
}

// `pair` is the CTA pair's index inside its cluster: the pairs of a cluster work on adjacent m-tiles of the same unit
// (same n-tile, batch entry and k-range, i.e. the same B operand, which they fetch once and multicast).
template <int EPI>
__device__ __forceinline__ Tile decode_tile(const GemmParams& p, long long unit, int j, int pair = 0) {
  Tile t;
  if (EPI == EPI_STORE && p.batch_fastest) {
    // the Lambda^-1 tile of an (m, n) block is read by every batch entry: schedule those units back to back so that
    // it is fetched from HBM once and served from L2 to the rest
    t.b = (int)(unit % p.batch);
    long long r = unit / p.batch;
    t.m_blk = (int)(r % p.m_blocks);
    r /= p.m_blocks;
    t.n_blk = (int)(r % p.n_blocks);
    const int ks = (int)(r / p.n_blocks);
    t.kb0 = ks * p.kb_per_split;
    t.kb1 = min(p.k_blocks, t.kb0 + p.kb_per_split);
    return t;
  }
  if (EPI == EPI_STORE && p.symmetric) {
    // upper-triangular tile index r = n (n + 1) / 2 + m  (m <= n), then k-split, then batch
    const long long tri = unit % p.tri_tiles;
    long long r = unit / p.tri_tiles;
    int n = (int)((sqrtf(8.f * (float)tri + 1.f) - 1.f) * 0.5f);
    while ((long long)(n + 1) * (n + 2) / 2 <= tri) ++n;
    while ((long long)n * (n + 1) / 2 > tri) --n;
    t.n_blk = n;
    t.m_blk = (int)(tri - (long long)n * (n + 1) / 2);
    const int ks = (int)(r % p.k_splits);
    t.b = (int)(r / p.k_splits);
    t.kb0 = ks * p.kb_per_split;
    t.kb1 = min(p.k_blocks, t.kb0 + p.kb_per_split);
    return t;
  }
  t.m_blk = (int)(unit % p.m_units) * p.mc + pair;
  long long r = unit / p.m_units;
  if (EPI == EPI_ROWDOT) {
    t.b = (int)(r / p.n_splits);
    const int jn = j / p.k_chunks;
    t.n_blk = (int)(r % p.n_splits) * p.nb_per_split + jn;
    const int kc = j - jn * p.k_chunks;
    t.kb0 = kc * p.kb_per_chunk;
    t.kb1 = min(p.k_blocks, t.kb0 + p.kb_per_chunk);
  } else if (EPI == EPI_REGACC && p.regacc_mode != REGACC_KCHUNK) {
    t.n_blk = (int)(r % p.n_blocks);
    const long long chunk = r / p.n_blocks;
    t.b = (int)(chunk * p.inner + j);
    t.kb0 = 0;
    t.kb1 = p.k_blocks;
  } else {
    t.n_blk = (int)(r % p.n_blocks);
    r /= p.n_blocks;
    const int ks = (int)(r % p.k_splits);
    t.b = (int)(r / p.k_splits);
    const int split0 = ks * p.kb_per_split;
    const int split1 = min(p.k_blocks, split0 + p.kb_per_split);
    if (EPI == EPI_REGACC) {
      t.kb0 = split0 + j * p.kb_per_chunk;
      t.kb1 = min(split1, t.kb0 + p.kb_per_chunk);
    } else {
      t.kb0 = split0;
      t.kb1 = split1;
    }
  }
  return t;
}

// Store of a 32 x 32 accumulator chunk (rows = the warp's 32 lanes, columns col0..col0+31) through every STORE
// option.  Bias layers make d_in+1 odd, so fp32 factors such as the covariance [d,d], Lambda [d_out,d_in+1] and
// Lambda^-1 have rows that are not 16-byte aligned, and a lane that walks its own accumulator row through them
// touches 32 different cache lines per instruction.  Three paths, chosen warp-uniformly:
//   registers  no `mul` factor and no row-major fp32 target: every layout left is contiguous per LANE (bf16 planes
//              as 16-byte vectors, transposed fp32 as consecutive lanes), so the values never leave registers;
//   vector     the chunk is parked in shared memory ([32][33] floats per epilogue warp) and re-read so that a group of
//              lanes shares a row: 4 lanes x 8 columns for bf16 planes (16-byte stores, two full sectors per row), 8
//              lanes x 4 columns for aligned fp32 rows (float4 read-modify-write / atomics); the Lambda^-1 factor is
//              read in the same mapping, and with HOIST all loads of the chunk are issued before the first use;
//   rows       whatever is left (unaligned fp32 rows, reduce_sq, mixed targets): lane = column, one coalesced
//              128-byte line per row, loads hoisted in groups of rows.
// Must be called by all 32 lanes.  Returns the lane's share of sum((alpha D)^2 * mul) when p.reduce_sq is set.
__device__ __forceinline__ void put_f32(const GemmParams& p, float* o, float v, float old) {
  if (p.use_atomic) atomicAdd(o, v);
  else *o = v + old;
}

// Offset of element (row, col) of a non-transposed plane output; with column groups (flat token index -> per-example
// [rows, S] matrices) the column selects the batch entry.  Groups are multiples of 8, so an aligned 8-column vector
// never straddles two of them.
__device__ __forceinline__ long long plane_index(const GemmParams& p, long long row, int col) {
  if (p.col_group > 0) return (long long)(col / p.col_group) * p.out_bs_s + row * p.ldo_s + (col % p.col_group);
  return row * p.ldo_s + col;
}

// TMA64 (one epilogue warpgroup): plain plane outputs leave as [32 rows x 64 columns] boxes, i.e. whole 128-byte lines
// per row — two consecutive 32-column chunks share one 128B-swizzled staging tile per plane (`half` says which half this
// chunk fills, `flush` that the box is complete or has no second chunk inside the matrix).  A B200 SM retires TMA store
// requests at a fixed rate per box ROW, so 64-byte rows (32-column boxes) capped the per-sample-gradient formation at
// ~2.5 TB/s of plane writes; full lines double the bytes per request.
template <int OFF, int N, bool HOIST, bool TMA64 = false>
__device__ __forceinline__ float store_chunk(const GemmParams& p, const CUtensorMap* tm_o_hi, const CUtensorMap* tm_o_lo,
                                             float alpha, int b, long long row, int col0, const float (&acc)[N],
                                             float* st, int lane, bool mirror = false, int half = 0, bool flush = true) {
  // mirror: the transposed copy of an off-diagonal tile of a symmetric product (fp32 target, no factor)
  const bool transpose_out = p.transpose_out != 0 || mirror;
  const bool direct_f32 = p.out_f32 != nullptr && !transpose_out && !p.reduce_sq;
  const bool rmw = p.accumulate && !p.use_atomic;
  const bool full = col0 + 32 <= p.N;
  float* ob = p.out_f32 != nullptr ? p.out_f32 + (long long)b * p.out_bs : nullptr;
  __nv_bfloat16* oh = p.out_hi != nullptr ? p.out_hi + (long long)b * p.out_bs_s : nullptr;
  __nv_bfloat16* ol = (oh != nullptr && p.out_lo != nullptr) ? p.out_lo + (long long)b * p.out_bs_s : nullptr;

  // ---- TMA: plain planes.  The lane's 32 values are split, parked as a dense [32][32] bf16 tile per plane in the
  // warp's staging buffer and leave as two asynchronous bulk tensor stores: fully coalesced lines, no LSU work per
  // row, rows/columns outside the matrix clipped by the tensor map (whose inner extent is the padded ld, so the
  // zero columns of the last tile double as the operand padding).
  if (p.tma_store && !mirror && TMA64) {
    uint8_t* sh = reinterpret_cast<uint8_t*>(st);  // [32 rows][128 bytes], 16-byte chunks XOR-swizzled by (row & 7)
    uint8_t* sl = sh + 4096;
    if (half == 0) {
      if (lane == 0) tma_store_wait_read();  // the previous box's stores have finished reading the staging tiles
      __syncwarp();
    }
#pragma unroll
    for (int grp = 0; grp < 4; ++grp) {
      __nv_bfloat16 h[8], l[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        float v = alpha * acc[OFF + grp * 8 + e];
        if (p.square) v *= v;
        split_bf16(v, h[e], l[e]);
      }
      const int at = lane * 128 + (((half * 4 + grp) ^ (lane & 7)) << 4);
      *reinterpret_cast<uint4*>(sh + at) = *reinterpret_cast<uint4*>(h);
      if (p.out_lo != nullptr) *reinterpret_cast<uint4*>(sl + at) = *reinterpret_cast<uint4*>(l);
    }
    if (flush) {
      if (half == 0) {  // no second chunk inside the matrix: the right half of the box is operand padding (zeros)
#pragma unroll
        for (int grp = 4; grp < 8; ++grp) {
          const int at = lane * 128 + ((grp ^ (lane & 7)) << 4);
          *reinterpret_cast<uint4*>(sh + at) = make_uint4(0u, 0u, 0u, 0u);
          if (p.out_lo != nullptr) *reinterpret_cast<uint4*>(sl + at) = make_uint4(0u, 0u, 0u, 0u);
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        const int row0 = (int)(row - lane);
        tma_store_3d(tm_o_hi, smem_u32(sh), col0 - half * 32, row0, b);
        if (p.out_lo != nullptr) tma_store_3d(tm_o_lo, smem_u32(sl), col0 - half * 32, row0, b);
        tma_store_commit();
      }
    }
    return 0.f;
  }
  if (p.tma_store && !mirror) {
    __nv_bfloat16* sh = reinterpret_cast<__nv_bfloat16*>(st);
    __nv_bfloat16* sl = sh + 32 * 32;
    if (lane == 0) tma_store_wait_read();  // the previous chunk's stores have finished reading the staging tile
    __syncwarp();
#pragma unroll
    for (int grp = 0; grp < 4; ++grp) {
      __nv_bfloat16 h[8], l[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        float v = alpha * acc[OFF + grp * 8 + e];
        if (p.square) v *= v;
        split_bf16(v, h[e], l[e]);
      }
      *reinterpret_cast<uint4*>(sh + lane * 32 + grp * 8) = *reinterpret_cast<uint4*>(h);
      if (p.out_lo != nullptr) *reinterpret_cast<uint4*>(sl + lane * 32 + grp * 8) = *reinterpret_cast<uint4*>(l);
    }
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
      const int row0 = (int)(row - lane);
      tma_store_3d(tm_o_hi, smem_u32(sh), col0, row0, b);
      if (p.out_lo != nullptr) tma_store_3d(tm_o_lo, smem_u32(sl), col0, row0, b);
      tma_store_commit();
    }
    return 0.f;
  }

  // ---- registers ----
  if (p.mul == nullptr && !direct_f32 && !p.reduce_sq) {
    if (row >= p.M) return 0.f;
    if (ob != nullptr) {  // transposed fp32: consecutive lanes = consecutive addresses
      // accumulation goes through fire-and-forget L2 reductions (every element has exactly one writer per launch
      // unless the contraction is split, so the result is the same as load-add-store without its round trip)
      const bool red = p.use_atomic || p.accumulate;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        if (col0 + i < p.N) {
          float v = alpha * acc[OFF + i];
          if (p.square) v *= v;
          float* o = ob + (long long)(col0 + i) * p.ldo + row;
          if (red) atomicAdd(o, v);
          else *o = v;
        }
      }
    }
    if (oh != nullptr) {
      if (!transpose_out && p.vec_ok && full) {
#pragma unroll
        for (int grp = 0; grp < 4; ++grp) {
          __nv_bfloat16 h[8], l[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            float v = alpha * acc[OFF + grp * 8 + e];
            if (p.square) v *= v;
            split_bf16(v, h[e], l[e]);
          }
          const long long at = plane_index(p, row, col0 + grp * 8);
          *reinterpret_cast<uint4*>(oh + at) = *reinterpret_cast<uint4*>(h);
          if (ol) *reinterpret_cast<uint4*>(ol + at) = *reinterpret_cast<uint4*>(l);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          if (col0 + i < p.N) {
            float v = alpha * acc[OFF + i];
            if (p.square) v *= v;
            const long long idx = transpose_out ? (long long)(col0 + i) * p.ldo_s + row : plane_index(p, row, col0 + i);
            __nv_bfloat16 h, l;
            split_bf16(v, h, l);
            oh[idx] = h;
            if (ol) ol[idx] = l;
          }
        }
      }
    }
    return 0.f;
  }

  // ---- park the chunk ----
  const long long row0 = row - lane;  // first row of this warp's 32-row slab
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 32; ++i) st[lane * 33 + i] = alpha * acc[OFF + i];
  __syncwarp();

  // ---- vector: bf16 planes, 4 lanes x 8 columns per row, 8 rows per pass ----
  if (oh != nullptr && ob == nullptr && !transpose_out && p.vec_ok && full) {
    const int qr = lane >> 2, cg = (lane & 3) * 8;
    constexpr int PASSES = HOIST ? 4 : 1;  // passes whose factor loads are in flight together
    for (int it0 = 0; it0 < 4; it0 += PASSES) {
      float m[PASSES][8];
#pragma unroll
      for (int u = 0; u < PASSES; ++u) {
        const long long rr = row0 + (it0 + u) * 8 + qr;
        const float* mp = p.mul + rr * p.ldmul + col0 + cg;
        if (rr < p.M && p.mul_vec4) {
          const float4 m0 = __ldg(reinterpret_cast<const float4*>(mp));
          const float4 m1 = __ldg(reinterpret_cast<const float4*>(mp) + 1);
          m[u][0] = m0.x; m[u][1] = m0.y; m[u][2] = m0.z; m[u][3] = m0.w;
          m[u][4] = m1.x; m[u][5] = m1.y; m[u][6] = m1.z; m[u][7] = m1.w;
        } else {
#pragma unroll
          for (int e = 0; e < 8; ++e) m[u][e] = rr < p.M ? __ldg(mp + e) : 0.f;
        }
      }
#pragma unroll
      for (int u = 0; u < PASSES; ++u) {
        const int r = (it0 + u) * 8 + qr;
        const long long rr = row0 + r;
        if (rr >= p.M) continue;
        __nv_bfloat16 h[8], l[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          float v = st[r * 33 + cg + e] * m[u][e];
          if (p.square) v *= v;
          split_bf16(v, h[e], l[e]);
        }
        const long long at = plane_index(p, rr, col0 + cg);
        *reinterpret_cast<uint4*>(oh + at) = *reinterpret_cast<uint4*>(h);
        if (ol) *reinterpret_cast<uint4*>(ol + at) = *reinterpret_cast<uint4*>(l);
      }
    }
    __syncwarp();
    return 0.f;
  }

  // ---- vector: aligned fp32 rows, 8 lanes x 4 columns per row, 4 rows per pass ----
  if (direct_f32 && oh == nullptr && p.f32_vec4 && full && (p.mul == nullptr || p.mul_vec4)) {
    const int qr = lane >> 3, cg = (lane & 7) * 4;
    constexpr int PASSES = HOIST ? 8 : 2;
    for (int it0 = 0; it0 < 8; it0 += PASSES) {
      float4 m[PASSES], old[PASSES];
#pragma unroll
      for (int u = 0; u < PASSES; ++u) {
        const long long rr = row0 + (it0 + u) * 4 + qr;
        const bool ok = rr < p.M;
        m[u] = (ok && p.mul != nullptr) ? __ldg(reinterpret_cast<const float4*>(p.mul + rr * p.ldmul + col0 + cg))
                                        : make_float4(1.f, 1.f, 1.f, 1.f);
        old[u] = (ok && rmw) ? *reinterpret_cast<const float4*>(ob + rr * p.ldo + col0 + cg)
                             : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < PASSES; ++u) {
        const int r = (it0 + u) * 4 + qr;
        const long long rr = row0 + r;
        if (rr >= p.M) continue;
        float4 v = make_float4(st[r * 33 + cg] * m[u].x, st[r * 33 + cg + 1] * m[u].y, st[r * 33 + cg + 2] * m[u].z,
                               st[r * 33 + cg + 3] * m[u].w);
        if (p.square) { v.x *= v.x; v.y *= v.y; v.z *= v.z; v.w *= v.w; }
        float4* o = reinterpret_cast<float4*>(ob + rr * p.ldo + col0 + cg);
        if (p.use_atomic) atomicAdd(o, v);
        else *o = make_float4(v.x + old[u].x, v.y + old[u].y, v.z + old[u].z, v.w + old[u].w);
      }
    }
    __syncwarp();
    return 0.f;
  }

  // ---- rows: lane = column ----
  float part = 0.f;
  const int col = col0 + lane;
  const bool col_ok = col < p.N;
  constexpr int G = HOIST ? 16 : 8;
  for (int r0 = 0; r0 < 32; r0 += G) {
    if (row0 + r0 >= p.M) break;  // warp-uniform
    float m[G], old[G];
#pragma unroll
    for (int j = 0; j < G; ++j) {
      const long long rr = row0 + r0 + j;
      const bool ok = col_ok && rr < p.M;
      m[j] = (ok && p.mul != nullptr) ? __ldg(p.mul + rr * p.ldmul + col) : 1.f;
      old[j] = (ok && rmw && direct_f32) ? ob[rr * p.ldo + col] : 0.f;
    }
#pragma unroll
    for (int j = 0; j < G; ++j) {
      const long long rr = row0 + r0 + j;
      const bool ok = col_ok && rr < p.M;
      float v = st[(r0 + j) * 33 + lane];  // alpha * D[rr][col]
      if (p.reduce_sq) {
        if (ok) part = fmaf(v * v, m[j], part);
        continue;
      }
      v *= m[j];
      if (p.square) v *= v;
      st[(r0 + j) * 33 + lane] = v;
      if (direct_f32 && ok) put_f32(p, ob + rr * p.ldo + col, v, old[j]);
    }
  }
  __syncwarp();
  if (p.reduce_sq) return part;
  // lanes return to their own row for the lane-contiguous layouts
  if (row < p.M) {
    if (ob != nullptr && transpose_out) {
#pragma unroll 8
      for (int i = 0; i < 32; ++i) {
        if (col0 + i < p.N) {
          float* o = ob + (long long)(col0 + i) * p.ldo + row;
          put_f32(p, o, st[lane * 33 + i], rmw ? *o : 0.f);
        }
      }
    }
    if (oh != nullptr) {
      if (!transpose_out && p.vec_ok && full) {
#pragma unroll
        for (int grp = 0; grp < 4; ++grp) {
          __nv_bfloat16 h[8], l[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) split_bf16(st[lane * 33 + grp * 8 + e], h[e], l[e]);
          const long long at = plane_index(p, row, col0 + grp * 8);
          *reinterpret_cast<uint4*>(oh + at) = *reinterpret_cast<uint4*>(h);
          if (ol) *reinterpret_cast<uint4*>(ol + at) = *reinterpret_cast<uint4*>(l);
        }
      } else {
#pragma unroll 8
        for (int i = 0; i < 32; ++i) {
          if (col0 + i < p.N) {
            const long long idx = transpose_out ? (long long)(col0 + i) * p.ldo_s + row : plane_index(p, row, col0 + i);
            __nv_bfloat16 h, l;
            split_bf16(st[lane * 33 + i], h, l);
            oh[idx] = h;
            if (ol) ol[idx] = l;
          }
        }
      }
    }
  }
  __syncwarp();
  return 0.f;
}

// Width of the MMA issued for tile `t`: the last n-tile of a ragged N (every bias layer: d_in + 1 = 769, 4097, ...) only
// needs its valid columns, rounded up to whole 32-column epilogue chunks (so that every column of a chunk that is read
// back was written by this tile: zeros from the zero-filled operand rows, never stale TMEM).  A 32-wide MMA costs about a
// quarter of a 256-wide one (it is bound by the shared-memory read of A).  Not with B-tile multicast (fixed quarters).
template <int BLOCK_N, int MC>
__device__ __forceinline__ int tile_n_mma(const GemmParams& p, const Tile& t) {
  if (MC != 1) return BLOCK_N;
  const int valid = p.N - t.n_blk * BLOCK_N;
  if (valid >= BLOCK_N) return BLOCK_N;
  if (BLOCK_N > 128 && valid > 64) return valid > 128 ? BLOCK_N : 128;
  if (BLOCK_N > 64 && valid > 32) return valid > 64 ? BLOCK_N : 64;
  return BLOCK_N > 32 ? (valid > 32 ? BLOCK_N : 32) : BLOCK_N;
}

// keep the [N, ld) padding of non-transposed split outputs finite (zero): they may be re-read as part
// of a flattened contraction dimension.
__device__ __forceinline__ void store_zero_pad(const GemmParams& p, int b, long long row) {
  __nv_bfloat16* oh = p.out_hi + (long long)b * p.out_bs_s + row * p.ldo_s;
  __nv_bfloat16* ol = p.out_lo != nullptr ? p.out_lo + (long long)b * p.out_bs_s + row * p.ldo_s : nullptr;
  for (long long c = p.N; c < p.ldo_s; ++c) {
    oh[c] = __float2bfloat16_rn(0.f);
    if (ol) ol[c] = __float2bfloat16_rn(0.f);
  }
}

// ROWDOT: the 32 factors g[row][col0 .. col0+31] of this thread's accumulator row (zeros outside the matrix), issued
// as independent loads so that a whole chunk is in flight at once.
__device__ __forceinline__ void load_g_chunk(const GemmParams& p, const float* grow, int col0, bool row_ok,
                                             float (&dst)[32]) {
  if (row_ok && p.g_vec4 && col0 + 32 <= p.N) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 gv = __ldg(reinterpret_cast<const float4*>(grow + col0) + i);
      dst[4 * i + 0] = gv.x; dst[4 * i + 1] = gv.y; dst[4 * i + 2] = gv.z; dst[4 * i + 3] = gv.w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i) dst[i] = (row_ok && col0 + i < p.N) ? __ldg(grow + col0 + i) : 0.f;
  }
}

// CG = 1: one CTA per tile (M = 128).  CG = 2: a CTA PAIR per tile (tcgen05 cta_group::2, M = 256):
// each CTA stages its own 128 rows of A and HALF of the B tile, the pair's tensor cores read both B
// halves, so shared-memory and L2 operand traffic per MMA drop by a third.
template <int BLOCK_N, int BLOCK_K, int NSPLIT, int CG, int MC = 1, int EW = 4>
struct GemmCfg {
  static constexpr int THREADS = 128 + 32 * EW;  // 4 control warps + EW epilogue warps
  static constexpr int BLOCK_M = 128;          // accumulator rows per CTA
  static constexpr int TILE_M = BLOCK_M * CG;  // rows of one scheduled tile
  static constexpr int LOAD_N = BLOCK_N / CG;  // B rows staged by one CTA
  static constexpr int FETCH_N = LOAD_N / MC;  // B rows one CTA fetches itself (and multicasts to its MC - 1 siblings)
  static_assert(MC == 1 || (CG == 2 && MC == 2), "B-tile multicast is built for clusters of two CTA pairs");
  static constexpr int SWIZZLE = BLOCK_K * 2;  // bytes per smem row == swizzle span
  static constexpr int A_PLANE = BLOCK_M * BLOCK_K * 2;
  static constexpr int B_PLANE = LOAD_N * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = NSPLIT * (A_PLANE + B_PLANE);
  // per epilogue warp: one [32][33] fp32 staging tile; with one epilogue warpgroup 8 KB (1024-byte aligned): two
  // 128B-swizzled [32 rows][64 columns] bf16 tiles (hi, lo) for the 64-column TMA plane stores
  static constexpr int EPI_WARP_BYTES = EW == 4 ? 8192 : 32 * 33 * 4;
  static constexpr int EPI_STAGE_BYTES = EW * EPI_WARP_BYTES;
  static constexpr int BARRIER_BYTES = 1024;  // keeps the staging tiles behind it 1024-byte aligned
  static_assert(EW == 4 || EW == 8, "one or two epilogue warpgroups");
  static constexpr int SMEM_BUDGET = 227 * 1024 - 2048 - EPI_STAGE_BYTES;
  static constexpr int STAGES_RAW = SMEM_BUDGET / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int ACC_STAGES = 2;
  static constexpr int TMEM_COLS_RAW = ACC_STAGES * BLOCK_N;
  static constexpr int TMEM_COLS =
      TMEM_COLS_RAW <= 32 ? 32 : TMEM_COLS_RAW <= 64 ? 64 : TMEM_COLS_RAW <= 128 ? 128
                                 : TMEM_COLS_RAW <= 256 ? 256 : 512;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + BARRIER_BYTES + EPI_STAGE_BYTES;
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
  static_assert(STAGES >= 2, "need at least a double-buffered smem ring");
  static_assert(BLOCK_K == 64 || BLOCK_K == 32, "BLOCK_K must match a 128B or 64B swizzle span");
  static_assert(BLOCK_N % 16 == 0 && BLOCK_N >= 16 && BLOCK_N <= 256, "invalid UMMA N");
  static_assert(CG == 1 || CG == 2, "cta_group must be 1 or 2");
};

// EW = 8 (register-accumulating kernels with 256-wide tiles): two epilogue warpgroups, each thread of which keeps
// 128 accumulator columns of its row; the control warps hand most of their registers over (setmaxnreg).
template <int BLOCK_N, int BLOCK_K, int NSPLIT, int EPI, int CG, int MC, int EW>
__global__ void __launch_bounds__(128 + 32 * EW, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
               const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo,
               const __grid_constant__ CUtensorMap tm_o_hi, const __grid_constant__ CUtensorMap tm_o_lo,
               const GemmParams p) {
  using Cfg = GemmCfg<BLOCK_N, BLOCK_K, NSPLIT, CG, MC, EW>;
  static_assert(EW == 4 || EPI == EPI_REGACC, "only the register-accumulating epilogue splits the tile's columns");
  static_assert(NSPLIT == 1 || NSPLIT == 2, "1 (bf16) or 2 (hi/lo; bf16 or, for strict operands, fp16) operand planes");
  constexpr int BLOCK_M = Cfg::BLOCK_M;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int ACC_STAGES = Cfg::ACC_STAGES;
  const uint32_t IDESC = p.f16 ? make_idesc_bf16(Cfg::TILE_M, BLOCK_N, true) : make_idesc_bf16(Cfg::TILE_M, BLOCK_N);

  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment is required by the 128B swizzle atoms.
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + STAGES * Cfg::STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tmem_full_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + s); };
  auto tmem_empty_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + ACC_STAGES + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 2 * ACC_STAGES);
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_gen + STAGES * Cfg::STAGE_BYTES +
                                           8 * (2 * STAGES + 2 * ACC_STAGES));

  auto smem_a = [&](int s, int plane) {
    return smem_base + s * Cfg::STAGE_BYTES + plane * Cfg::A_PLANE;
  };
  auto smem_b = [&](int s, int plane) {
    return smem_base + s * Cfg::STAGE_BYTES + NSPLIT * Cfg::A_PLANE + plane * Cfg::B_PLANE;
  };

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cluster_rank = CG == 2 ? cluster_ctarank() : 0u;
  const uint32_t cta_rank = cluster_rank & 1u;         // rank inside the CTA pair; 0 = leader (issues the pair's MMAs)
  const int pair_id = (int)(cluster_rank >> 1);        // which pair of the cluster (MC == 2: 0 or 1)
  const uint32_t leader_rank = cluster_rank & ~1u;     // cluster rank of this pair's leader
  constexpr int CLUSTER = CG * MC;
  const long long unit0 = (long long)(blockIdx.x / CLUSTER);
  const long long unit_stride = (long long)(gridDim.x / CLUSTER);

  if (CG == 2) cluster_sync_all();  // both CTAs of the pair are resident before the paired TMEM allocation
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a_hi);
    tma_prefetch_desc(&tm_b_hi);
    if (NSPLIT >= 2) {
      tma_prefetch_desc(&tm_a_lo);
      tma_prefetch_desc(&tm_b_lo);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), CG);   // one producer arrival (+tx bytes) per CTA of the group
      mbar_init(empty_bar(s), MC);  // one tcgen05.commit per pair of the cluster (multicast to all its CTAs)
    }
    for (int s = 0; s < ACC_STAGES; ++s) {
      mbar_init(tmem_full_bar(s), 1);
      mbar_init(tmem_empty_bar(s), EW * CG);  // one arrival per epilogue warp of the group
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    if (CG == 2) {
      tmem_alloc_2sm(tmem_slot, Cfg::TMEM_COLS);
      tmem_relinquish_2sm();
    } else {
      tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
      tmem_relinquish();
    }
  }
  tcgen05_fence_before();
  if (CG == 2) cluster_sync_all();  // the peer's barriers must exist before anyone signals them
  else __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  // EW == 8: 384 threads x 168 registers = 64512 at launch, and that is the pool: the control warpgroup drops to 56
  // here, the two epilogue warpgroups rise to 224 at the top of their branch (128 x 56 + 256 x 224 = 64512)
  if (EW == 8 && warp < 4) setmaxnreg_dec<56>();
  if (warp == 0 && lane == 0) {
    // ===================================== TMA producer ======================================
    int stage = 0;
    uint32_t phase = 0;
    uint32_t sync_step = 0;
    // ROWDOT streams B (the query store: every byte used once per launch) past a small A (the train batch, re-read for
    // every query): without hints the stream evicts A from L2 and A comes back from HBM once per query (measured: +50 %
    // DRAM traffic on the target layer)
    const bool hint = EPI == EPI_ROWDOT && CG == 2;
    const uint64_t pol_a = hint ? l2_policy_evict_last() : 0ull;
    const uint64_t pol_b = hint ? l2_policy_evict_first() : 0ull;
    for (long long unit = unit0; unit < p.num_units; unit += unit_stride) {
      const int inner = unit_inner_count<EPI>(p, unit);
      for (int j = 0; j < inner; ++j) {
        if (EPI == EPI_ROWDOT && p.sync_group > 0 && j % p.k_chunks == 0) {
          // The `sync_group` clusters that work on the m-units of one query stream the SAME P tiles.  They only share
          // them through L2 while they run within its retention window (~100 us of stream) of each other, and nothing
          // keeps them there over thousands of units: measured at Q = 1024, every cluster ended up fetching its own copy
          // from HBM (4.2x the algorithmic traffic).  A rendezvous of the group's producers at every n-tile restores the
          // lock-step; the grid is a whole number of groups, so the members of a group always hold the same query.
          const unsigned grp = (unsigned)(blockIdx.x / CLUSTER) / (unsigned)p.sync_group;
          const unsigned target = (unsigned)(CLUSTER * p.sync_group) * (++sync_step);
          atomicAdd(p.sync_ctr + grp, 1u);
          const long long t0 = clock64();
          unsigned seen;
          do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(p.sync_ctr + grp) : "memory");
            if (seen >= target) break;
            __nanosleep(32);
            if (clock64() - t0 > KFB_WATCHDOG_CYCLES) __trap();
          } while (true);
        }
        const Tile t = decode_tile<EPI>(p, unit, j, pair_id);
        const int row_a = t.m_blk * Cfg::TILE_M + (int)cta_rank * BLOCK_M;
        // each CTA of a pair stages ITS half of the B rows the MMA reads (a narrower MMA on a ragged last n-tile reads
        // fewer rows: the second CTA's half starts right behind the first one's)
        const int row_b = t.n_blk * BLOCK_N + (int)cta_rank * (tile_n_mma<BLOCK_N, MC>(p, t) / CG) + pair_id * Cfg::FETCH_N;
        const int ba = p.a_batched ? t.b : 0;
        const int bb = p.b_batched ? t.b : 0;
        for (int kb = t.kb0; kb < t.kb1; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const int k0 = kb * BLOCK_K;
          if (CG == 2) {
            // both CTAs credit the LEADER's full barrier: its MMA thread consumes the pair's stage
            if (cta_rank == 0) mbar_arrive_expect_tx(full_bar(stage), Cfg::STAGE_BYTES);
            else mbar_arrive_expect_tx_cluster(full_bar(stage), leader_rank, Cfg::STAGE_BYTES);
            if (hint) {
              tma_load_3d_2sm_hint(&tm_a_hi, full_bar(stage), smem_a(stage, 0), k0, row_a, ba, pol_a);
              if (NSPLIT == 2) tma_load_3d_2sm_hint(&tm_a_lo, full_bar(stage), smem_a(stage, 1), k0, row_a, ba, pol_a);
            } else {
              tma_load_3d_2sm(&tm_a_hi, full_bar(stage), smem_a(stage, 0), k0, row_a, ba);
              if (NSPLIT == 2) tma_load_3d_2sm(&tm_a_lo, full_bar(stage), smem_a(stage, 1), k0, row_a, ba);
            }
            if (MC == 2) {
              // this CTA fetches its quarter of the B tile and multicasts it to the same-ranked CTA of both pairs
              const uint32_t off = (uint32_t)(pair_id * Cfg::FETCH_N * Cfg::SWIZZLE);
              const uint16_t mask = (uint16_t)(0x5u << cta_rank);
              if (hint) {
                tma_load_3d_2sm_mc_hint(&tm_b_hi, full_bar(stage), smem_b(stage, 0) + off, k0, row_b, bb, mask, pol_b);
                if (NSPLIT == 2) tma_load_3d_2sm_mc_hint(&tm_b_lo, full_bar(stage), smem_b(stage, 1) + off, k0, row_b, bb, mask, pol_b);
              } else {
                tma_load_3d_2sm_mc(&tm_b_hi, full_bar(stage), smem_b(stage, 0) + off, k0, row_b, bb, mask);
                if (NSPLIT == 2) tma_load_3d_2sm_mc(&tm_b_lo, full_bar(stage), smem_b(stage, 1) + off, k0, row_b, bb, mask);
              }
            } else if (hint) {
              tma_load_3d_2sm_hint(&tm_b_hi, full_bar(stage), smem_b(stage, 0), k0, row_b, bb, pol_b);
              if (NSPLIT == 2) tma_load_3d_2sm_hint(&tm_b_lo, full_bar(stage), smem_b(stage, 1), k0, row_b, bb, pol_b);
            } else {
              tma_load_3d_2sm(&tm_b_hi, full_bar(stage), smem_b(stage, 0), k0, row_b, bb);
              if (NSPLIT == 2) tma_load_3d_2sm(&tm_b_lo, full_bar(stage), smem_b(stage, 1), k0, row_b, bb);
            }
          } else {
            mbar_arrive_expect_tx(full_bar(stage), Cfg::STAGE_BYTES);
            tma_load_3d(&tm_a_hi, full_bar(stage), smem_a(stage, 0), k0, row_a, ba);
            tma_load_3d(&tm_b_hi, full_bar(stage), smem_b(stage, 0), k0, row_b, bb);
            if (NSPLIT >= 2) {
              tma_load_3d(&tm_a_lo, full_bar(stage), smem_a(stage, 1), k0, row_a, ba);
              tma_load_3d(&tm_b_lo, full_bar(stage), smem_b(stage, 1), k0, row_b, bb);
            }
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1 && lane == 0 && cta_rank == 0) {
    // ====================================== MMA issuer =======================================
    int stage = 0;
    uint32_t phase = 0;
    uint32_t it = 0;
    for (long long unit = unit0; unit < p.num_units; unit += unit_stride) {
      const int inner = unit_inner_count<EPI>(p, unit);
      for (int j = 0; j < inner; ++j, ++it) {
        const Tile t = decode_tile<EPI>(p, unit, j, pair_id);
        const uint32_t as = it % ACC_STAGES;
        const uint32_t aphase = (it / ACC_STAGES) & 1u;
        mbar_wait(tmem_empty_bar(as), aphase ^ 1u);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + as * BLOCK_N;
        const int n_mma = tile_n_mma<BLOCK_N, MC>(p, t);
        const uint32_t idesc = n_mma == BLOCK_N ? IDESC : make_idesc_bf16(Cfg::TILE_M, n_mma, p.f16 != 0);
        for (int kb = t.kb0; kb < t.kb1; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tcgen05_fence_after();
          const uint64_t a_hi = make_kmajor_desc<Cfg::SWIZZLE>(smem_a(stage, 0));
          const uint64_t b_hi = make_kmajor_desc<Cfg::SWIZZLE>(smem_b(stage, 0));
          const uint64_t a_lo = make_kmajor_desc<Cfg::SWIZZLE>(smem_a(stage, NSPLIT >= 2 ? 1 : 0));
          const uint64_t b_lo = make_kmajor_desc<Cfg::SWIZZLE>(smem_b(stage, NSPLIT >= 2 ? 1 : 0));
#pragma unroll
          for (int k = 0; k < BLOCK_K / 16; ++k) {
            // advancing 16 bf16 (32 bytes) along K inside the swizzle span: +2 in 16-byte units
            const uint64_t koff = (uint64_t)(2 * k);
            const uint32_t acc = (kb > t.kb0 || k > 0) ? 1u : 0u;
            if (CG == 2) {
              if (NSPLIT == 2) {
                umma_bf16_2sm(d_tmem, a_lo + koff, b_hi + koff, idesc, acc);
                umma_bf16_2sm(d_tmem, a_hi + koff, b_lo + koff, idesc, 1u);
                umma_bf16_2sm(d_tmem, a_hi + koff, b_hi + koff, idesc, 1u);
              } else {
                umma_bf16_2sm(d_tmem, a_hi + koff, b_hi + koff, idesc, acc);
              }
            } else if (NSPLIT == 2) {
              umma_bf16(d_tmem, a_lo + koff, b_hi + koff, idesc, acc);
              umma_bf16(d_tmem, a_hi + koff, b_lo + koff, idesc, 1u);
              umma_bf16(d_tmem, a_hi + koff, b_hi + koff, idesc, 1u);
            } else {
              umma_bf16(d_tmem, a_hi + koff, b_hi + koff, idesc, acc);
            }
          }
          // smem stage reusable (in both CTAs) once these MMAs retire; accumulator published at the end
          if (CG == 2) {
            // the stage is free once BOTH pairs of the cluster have consumed it (their siblings write into it)
            umma_commit_2sm(empty_bar(stage), (uint16_t)((1u << CLUSTER) - 1u));
            if (kb == t.kb1 - 1) umma_commit_2sm(tmem_full_bar(as), (uint16_t)(3u << leader_rank));
          } else {
            umma_commit(empty_bar(stage));
            if (kb == t.kb1 - 1) umma_commit(tmem_full_bar(as));
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp >= 4) {
    // ======================================= epilogue ========================================
    if (EW == 8) setmaxnreg_inc<224>();
    const int quarter = warp & 3;              // TMEM lane quarter this warp may access
    const int lane_row = quarter * 32 + lane;  // accumulator row owned by this thread
    float* st = reinterpret_cast<float*>(smem_gen + STAGES * Cfg::STAGE_BYTES + Cfg::BARRIER_BYTES) + (warp - 4) * (Cfg::EPI_WARP_BYTES / 4);
    constexpr bool TMA64 = EW == 4;  // 64-column TMA plane stores (see store_chunk)
    constexpr int NACC = EPI == EPI_REGACC ? BLOCK_N / (EW / 4) : 1;  // accumulator columns kept per thread
    const int col_off = EW == 8 ? ((warp - 4) >> 2) * NACC : 0;          // ... starting at this column of the tile
    // strict operands were scaled by powers of two: undo both scales together with alpha (exact)
    const float alpha = p.f16 ? p.alpha / strict_scale(__ldg(p.a_absmax)) / strict_scale(__ldg(p.b_absmax)) : p.alpha;
    uint32_t it = 0;
    for (long long unit = unit0; unit < p.num_units; unit += unit_stride) {
      const int inner = unit_inner_count<EPI>(p, unit);
      float rowdot = 0.f;
      float racc[NACC];
#pragma unroll
      for (int i = 0; i < NACC; ++i) racc[i] = 0.f;
      Tile t = decode_tile<EPI>(p, unit, 0, pair_id);
      if (EPI == EPI_REGACC && p.regacc_mode >= REGACC_BATCH_DOT) {
        // this thread's slice of the factor row (all batch entries of the unit share the (m, n) tile)
        const long long frow = (long long)t.m_blk * Cfg::TILE_M + (long long)cta_rank * BLOCK_M + lane_row;
        const int fcol0 = t.n_blk * BLOCK_N + col_off;
#pragma unroll
        for (int i = 0; i < NACC; ++i) {
          const bool ok = frow < p.M && fcol0 + i < p.N;
          racc[i] = !ok ? 0.f : (p.factor != nullptr ? __ldg(p.factor + frow * p.ldfactor + fcol0 + i) : 1.f);
        }
      }
      for (int j = 0; j < inner; ++j, ++it) {
        t = decode_tile<EPI>(p, unit, j, pair_id);
        const uint32_t as = it % ACC_STAGES;
        const uint32_t aphase = (it / ACC_STAGES) & 1u;
        const uint32_t taddr = tmem_base + as * BLOCK_N + ((uint32_t)(quarter * 32) << 16);
        const long long row = (long long)t.m_blk * Cfg::TILE_M + (long long)cta_rank * BLOCK_M + lane_row;
        const bool row_ok = row < p.M;
        const int n0 = t.n_blk * BLOCK_N;
        if (EPI == EPI_ROWDOT) {
          // software pipeline: the factors of chunk 0 are requested BEFORE waiting for the accumulator and those of
          // chunk c+1 while chunk c is reduced, so their L2 latency never sits on the pass's critical path
          const float* grow = p.g + (long long)t.b * p.g_bs + (row_ok ? row : 0) * p.ldg;
          float gb[2][32];
          load_g_chunk(p, grow, n0, row_ok, gb[0]);
          mbar_wait(tmem_full_bar(as), aphase);
          tcgen05_fence_after();
          // ... and so are the accumulator chunks: chunk c+1 is on its way from TMEM while chunk c is reduced
          uint32_t v[2][32];
          tmem_ld32(taddr, v[0]);
#pragma unroll
          for (int c = 0; c < BLOCK_N / 32; ++c) {
            const int col0 = n0 + c * 32;
            if (col0 >= p.N) break;  // warp-uniform
            const bool more = c + 1 < BLOCK_N / 32 && col0 + 32 < p.N;
            tmem_ld_wait();
            if (more) {
              tmem_ld32(taddr + (c + 1) * 32, v[(c + 1) & 1]);
              load_g_chunk(p, grow, col0 + 32, row_ok, gb[(c + 1) & 1]);
            }
#pragma unroll
            for (int i = 0; i < 32; ++i) rowdot = fmaf(__uint_as_float(v[c & 1][i]), gb[c & 1][i], rowdot);
          }
          tmem_ld_wait();
        } else if (EPI == EPI_REGACC) {
          // A pass is short here (strict rotations drain TMEM every 128 contraction elements), so the drain must not
          // serialise on the TMEM latency: chunk c+1 is requested before chunk c is added.
          mbar_wait(tmem_full_bar(as), aphase);
          tcgen05_fence_after();
          uint32_t v[2][32];
          tmem_ld32(taddr + col_off, v[0]);
#pragma unroll
          for (int c = 0; c < NACC / 32; ++c) {
            tmem_ld_wait();
            if (c + 1 < NACC / 32) tmem_ld32(taddr + col_off + (c + 1) * 32, v[(c + 1) & 1]);
            if (p.regacc_mode == REGACC_BATCH) {
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                const float x = __uint_as_float(v[c & 1][i]);
                racc[(EPI == EPI_REGACC ? c * 32 + i : 0)] += x * x;
              }
            } else if (p.regacc_mode == REGACC_BATCH_DOT) {
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                const float x = __uint_as_float(v[c & 1][i]);
                rowdot = fmaf(x * x, racc[(EPI == EPI_REGACC ? c * 32 + i : 0)], rowdot);
              }
            } else if (p.regacc_mode == REGACC_BATCH_MUL) {
              const int col0 = n0 + col_off + c * 32;
              if (col0 < p.N) {  // warp-uniform
                float y[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) y[i] = __uint_as_float(v[c & 1][i]) * racc[(EPI == EPI_REGACC ? c * 32 + i : 0)];
                const bool last_of_box = (c & 1) || c + 1 == NACC / 32 || col0 + 32 >= p.N;
                store_chunk<0, 32, false, TMA64>(p, &tm_o_hi, &tm_o_lo, alpha, t.b, row, col0, y, st, lane, false, c & 1, last_of_box);
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) racc[(EPI == EPI_REGACC ? c * 32 + i : 0)] += __uint_as_float(v[c & 1][i]);
            }
          }
        } else {
          mbar_wait(tmem_full_bar(as), aphase);
          tcgen05_fence_after();
#pragma unroll
          for (int c = 0; c < BLOCK_N / 32; ++c) {
            const int col0 = n0 + c * 32;
            if (col0 >= p.N) break;  // warp-uniform
            uint32_t v[32];
            tmem_ld32(taddr + c * 32, v);
            tmem_ld_wait();
            {
              float x[32];
#pragma unroll
              for (int i = 0; i < 32; ++i) x[i] = __uint_as_float(v[i]);
              const bool last_of_box = (c & 1) || c + 1 == BLOCK_N / 32 || col0 + 32 >= p.N;
              rowdot += store_chunk<0, 32, true, TMA64>(p, &tm_o_hi, &tm_o_lo, alpha, t.b, row, col0, x, st, lane, false, c & 1, last_of_box);
              if (p.symmetric && t.n_blk != t.m_blk) store_chunk<0, 32, true>(p, &tm_o_hi, &tm_o_lo, alpha, t.b, row, col0, x, st, lane, true);
            }
          }
        }
        if (EPI == EPI_REGACC && p.regacc_mode == REGACC_BATCH_MUL && row_ok && p.zero_pad && !p.tma_store && col_off == 0 &&
            t.n_blk == p.n_blocks - 1)
          store_zero_pad(p, t.b, row);
        if (EPI == EPI_REGACC && p.regacc_mode == REGACC_BATCH_DOT) {
          float part = rowdot;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
          if (lane == 0 && part != 0.f) atomicAdd(p.out_f32 + (long long)t.b * p.out_bs, alpha * part);
          rowdot = 0.f;
        }
        if (EPI == EPI_STORE && p.reduce_sq) {
          // one scalar per batch entry: warp-reduce, then one atomic per warp
          float part = rowdot;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
          // store_chunk squared alpha-scaled values: sum (alpha D)^2 mul; the contract is alpha * sum D^2 mul
          if (lane == 0 && part != 0.f) atomicAdd(p.out_f32 + (long long)t.b * p.out_bs, part / alpha);
          rowdot = 0.f;
        }
        if (EPI == EPI_STORE && p.zero_pad && !p.tma_store && !p.reduce_sq && row_ok && t.n_blk == p.n_blocks - 1)
          store_zero_pad(p, t.b, row);
        // hand the accumulator stage back to the (leader's) MMA thread: one arrival per warp
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (CG == 2 && cta_rank != 0) mbar_arrive_cluster(tmem_empty_bar(as), leader_rank);
          else mbar_arrive(tmem_empty_bar(as));
        }
      }
      // ---- per-unit finalisation ----
      if (EPI == EPI_ROWDOT) {
        const long long row = (long long)t.m_blk * Cfg::TILE_M + (long long)cta_rank * BLOCK_M + lane_row;
        if (p.row_group > 1) {
          // sum the rows of one example: segmented warp reduction (groups are contiguous lane ranges), then one
          // atomic per group and warp
          const long long grp = row < p.M ? row / p.row_group : -1;
          const unsigned peers = __match_any_sync(0xffffffffu, grp);
          float val = alpha * rowdot;
#pragma unroll
          for (int off = 1; off < 32; off <<= 1) {
            const float other = __shfl_down_sync(0xffffffffu, val, off);
            if (lane + off < 32 && ((peers >> (lane + off)) & 1u)) val += other;
          }
          if (grp >= 0 && lane == __ffs(peers) - 1) atomicAdd(p.out_f32 + (long long)t.b * p.out_bs + grp, val);
        } else if (row < p.M) {
          float* o = p.out_f32 + (long long)t.b * p.out_bs + row;
          const float val = alpha * rowdot;
          if (p.use_atomic) atomicAdd(o, val);
          else if (p.accumulate) *o += val;
          else *o = val;
        }
      } else if (EPI == EPI_REGACC && p.regacc_mode < REGACC_BATCH_DOT) {
        const long long row = (long long)t.m_blk * Cfg::TILE_M + (long long)cta_rank * BLOCK_M + lane_row;
        const int n0 = t.n_blk * BLOCK_N + col_off;
        const int ob = p.regacc_mode == REGACC_BATCH ? 0 : t.b;
        // warp-uniform conditions: store_chunk is a warp-cooperative call
        if (NACC >= 32 && n0 < p.N) store_chunk<0, NACC, false, TMA64>(p, &tm_o_hi, &tm_o_lo, alpha, ob, row, n0, racc, st, lane, false, 0, NACC < 64 || n0 + 32 >= p.N);
        if (NACC >= 64 && n0 + 32 < p.N) store_chunk<(NACC >= 64 ? 32 : 0), NACC, false, TMA64>(p, &tm_o_hi, &tm_o_lo, alpha, ob, row, n0 + 32, racc, st, lane, false, 1, true);
        if (NACC >= 128 && n0 + 64 < p.N) store_chunk<(NACC >= 128 ? 64 : 0), NACC, false, TMA64>(p, &tm_o_hi, &tm_o_lo, alpha, ob, row, n0 + 64, racc, st, lane, false, 0, n0 + 96 >= p.N);
        if (NACC >= 128 && n0 + 96 < p.N) store_chunk<(NACC >= 128 ? 96 : 0), NACC, false, TMA64>(p, &tm_o_hi, &tm_o_lo, alpha, ob, row, n0 + 96, racc, st, lane, false, 1, true);
        if (row < p.M && p.zero_pad && !p.tma_store && t.n_blk == p.n_blocks - 1) store_zero_pad(p, ob, row);
      }
    }
  }

  if (warp >= 4 && lane == 0 && p.tma_store) tma_store_wait_all();  // staging tiles are read / outputs written
  tcgen05_fence_before();
  if (CG == 2) cluster_sync_all();  // the peer may still be reading operands / signalling our barriers
  else __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    if (CG == 2) tmem_dealloc_2sm(tmem_base, Cfg::TMEM_COLS);
    else tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// -------------------------------------------------------------------------------------------------
// SIMT debug implementation of the same contract (kfb_set_gemm_backend(1)).  One thread per output.
// -------------------------------------------------------------------------------------------------
struct SimtOperand {
  const __nv_bfloat16* hi;
  const __nv_bfloat16* lo;
  long long ld, bs;
};

__device__ __forceinline__ float simt_plane(const __nv_bfloat16* p, long long i, int f16) {
  return f16 ? __half2float(reinterpret_cast<const __half*>(p)[i]) : __bfloat162float(p[i]);
}

__device__ __forceinline__ float simt_dot(const SimtOperand& A, const SimtOperand& B, long long ab,
                                          long long bb, long long m, long long n, int K, int f16) {
  const __nv_bfloat16* ah = A.hi + ab * A.bs + m * A.ld;
  const __nv_bfloat16* bh = B.hi + bb * B.bs + n * B.ld;
  const __nv_bfloat16* al = A.lo ? A.lo + ab * A.bs + m * A.ld : nullptr;
  const __nv_bfloat16* bl = B.lo ? B.lo + bb * B.bs + n * B.ld : nullptr;
  float acc = 0.f;
  for (int k = 0; k < K; ++k) {
    const float a_h = simt_plane(ah, k, f16), b_h = simt_plane(bh, k, f16);
    const float a_l = al ? simt_plane(al, k, f16) : 0.f;
    const float b_l = bl ? simt_plane(bl, k, f16) : 0.f;
    acc += a_l * b_h + a_h * b_l + a_h * b_h;
  }
  return acc;
}

__global__ void gemm_simt_kernel(SimtOperand A, SimtOperand B, GemmParams p, int epi) {
  const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (p.f16) p.alpha = p.alpha / strict_scale(*p.a_absmax) / strict_scale(*p.b_absmax);
  if (epi == KFB_EPI_STORE) {
    const long long total = (long long)p.batch * p.M * p.N;
    if (tid >= total) return;
    const long long n = tid % p.N, m = (tid / p.N) % p.M, b = tid / ((long long)p.N * p.M);
    if (p.reduce_sq) {
      const float d = simt_dot(A, B, p.a_batched ? b : 0, p.b_batched ? b : 0, m, n, p.K, p.f16);
      atomicAdd(p.out_f32 + b * p.out_bs, p.alpha * d * d * (p.mul ? p.mul[m * p.ldmul + n] : 1.f));
      return;
    }
    float val = p.alpha * simt_dot(A, B, p.a_batched ? b : 0, p.b_batched ? b : 0, m, n, p.K, p.f16);
    if (p.mul) val *= p.mul[m * p.ldmul + n];
    if (p.square) val *= val;
    if (p.out_f32) {
      float* o = p.out_f32 + b * p.out_bs + (p.transpose_out ? n * p.ldo + m : m * p.ldo + n);
      if (p.accumulate) *o += val;
      else *o = val;
    }
    if (p.out_hi) {
      const long long idx = p.col_group > 0 ? (n / p.col_group) * p.out_bs_s + m * p.ldo_s + (n % p.col_group)
                                            : b * p.out_bs_s + (p.transpose_out ? n * p.ldo_s + m : m * p.ldo_s + n);
      __nv_bfloat16 h, l;
      split_bf16(val, h, l);
      p.out_hi[idx] = h;
      if (p.out_lo) p.out_lo[idx] = l;
      if (p.zero_pad && !p.transpose_out && n == p.N - 1) {
        for (long long c = p.N; c < p.ldo_s; ++c) {
          p.out_hi[b * p.out_bs_s + m * p.ldo_s + c] = __float2bfloat16_rn(0.f);
          if (p.out_lo) p.out_lo[b * p.out_bs_s + m * p.ldo_s + c] = __float2bfloat16_rn(0.f);
        }
      }
    }
  } else if (epi == KFB_EPI_ROWDOT) {
    const long long total = (long long)p.batch * p.M;
    if (tid >= total) return;
    const long long m = tid % p.M, b = tid / p.M;
    float sum = 0.f;
    for (int n = 0; n < p.N; ++n)
      sum += simt_dot(A, B, p.a_batched ? b : 0, p.b_batched ? b : 0, m, n, p.K, p.f16) * p.g[b * p.g_bs + m * p.ldg + n];
    if (p.row_group > 1) {
      atomicAdd(p.out_f32 + b * p.out_bs + m / p.row_group, p.alpha * sum);
      return;
    }
    float* o = p.out_f32 + b * p.out_bs + m;
    if (p.accumulate) *o += p.alpha * sum;
    else *o = p.alpha * sum;
  } else {
    const long long total = (long long)p.M * p.N;
    if (tid >= total) return;
    const long long n = tid % p.N, m = tid / p.N;
    float sum = 0.f;
    for (int b = 0; b < p.batch; ++b) {
      const float d = simt_dot(A, B, p.a_batched ? b : 0, p.b_batched ? b : 0, m, n, p.K, p.f16);
      sum += d * d;
    }
    p.out_f32[m * p.ldo + n] += p.alpha * sum;
  }
}

// -------------------------------------------------------------------------------------------------
// Host side
// -------------------------------------------------------------------------------------------------
static std::atomic<long long> g_launches{0};
static std::atomic<int> g_backend{0};
static std::atomic<int> g_cta_pairs{1};
static std::atomic<int> g_tma_store{1};
static std::atomic<int> g_multicast{1};
static std::atomic<int> g_wide_regacc{1};
void count_launch(int n) { g_launches += n; }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) ==
            cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  });
  return fn;
}

static int make_tmap(CUtensorMap* tm, const void* base, long long cols, long long rows,
                     long long batch, long long ld, long long bs, int box_k, int box_rows, bool swizzled = true) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) {
    set_error("cuTensorMapEncodeTiled is unavailable (no CUDA driver?)");
    return KFB_ERR_NO_DEVICE;
  }
  KFB_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "operand base must be 16-byte aligned");
  KFB_REQUIRE(ld % 8 == 0, "operand leading dimension (%lld) must be a multiple of 8", ld);
  KFB_REQUIRE(batch == 1 || bs % 8 == 0, "operand batch stride (%lld) must be a multiple of 8", bs);
  cuuint64_t gdim[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)batch};
  long long bs_bytes = (batch > 1 ? bs : rows * ld) * 2;
  if (bs_bytes < 16) bs_bytes = 16;
  cuuint64_t gstride[2] = {(cuuint64_t)(ld * 2), (cuuint64_t)bs_bytes};
  cuuint32_t box[3] = {(cuuint32_t)box_k, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  const CUtensorMapSwizzle sw = !swizzled ? CU_TENSOR_MAP_SWIZZLE_NONE
                                : (box_k == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B);
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstride,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): cols=%lld rows=%lld batch=%lld ld=%lld bs=%lld",
              (int)r, cols, rows, batch, ld, bs);
    return KFB_ERR_CUDA;
  }
  return KFB_OK;
}

static const int kRetryWithoutMulticast = 4242;
static std::atomic<int> g_resident_clusters{0};  // 4-CTA clusters of the fused pairwise kernel the device holds at once
static std::atomic<int> g_idle_fill{-1};
static std::atomic<int> g_group_sync{1};         // cluster-group rendezvous of the fused pairwise kernel

// Fork / join helper for the idle-SM side launch: one non-blocking stream and two events per device.
struct SideStream {
  cudaStream_t stream = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
};
static int side_stream(SideStream** out) {
  static SideStream table[16];
  static std::mutex mutex;
  int dev = 0;
  KFB_CUDA_TRY(cudaGetDevice(&dev));
  KFB_REQUIRE(dev >= 0 && dev < 16, "side_stream: device index out of range");
  std::lock_guard<std::mutex> lock(mutex);
  SideStream& s = table[dev];
  if (s.stream == nullptr) {
    KFB_CUDA_TRY(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    KFB_CUDA_TRY(cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming));
    KFB_CUDA_TRY(cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming));
  }
  *out = &s;
  return KFB_OK;
}

static inline kfb_split batch_slice(const kfb_split& s, long long b0, long long nb) {
  kfb_split v = s;
  if (s.batch > 1) {
    v.hi = static_cast<char*>(s.hi) + b0 * s.batch_stride * 2;
    v.lo = s.lo ? static_cast<char*>(s.lo) + b0 * s.batch_stride * 2 : nullptr;
    v.batch = nb;
  }
  return v;
}  // internal: launch_tc<..., MC = 2> declined, use the pair kernel

template <int BLOCK_N, int BLOCK_K, int NSPLIT, int EPI, int CG = 1, int MC = 1, int EW = 4>
static int launch_tc(const kfb_split& A, const kfb_split& B, GemmParams p, cudaStream_t stream) {
  using Cfg = GemmCfg<BLOCK_N, BLOCK_K, NSPLIT, CG, MC, EW>;
  p.m_blocks = (int)ceil_div_ll(p.M, Cfg::TILE_M);
  p.mc = MC;
  p.m_units = (int)ceil_div_ll(p.m_blocks, MC);
  p.n_blocks = (int)ceil_div_ll(p.N, BLOCK_N);
  p.k_blocks = (int)ceil_div_ll(p.K, BLOCK_K);
  const int max_pass_kb = (p.max_pass_k > 0 ? p.max_pass_k : kMaxPassK) / BLOCK_K > 0
                              ? (p.max_pass_k > 0 ? p.max_pass_k : kMaxPassK) / BLOCK_K : 1;
  if (EPI != EPI_STORE || Cfg::TILE_M != BLOCK_N || p.m_blocks != p.n_blocks) p.symmetric = 0;
  p.tri_tiles = (long long)p.m_blocks * (p.m_blocks + 1) / 2;
  const long long tiles = p.symmetric ? p.tri_tiles : (long long)p.m_units * p.n_blocks;
  if (EPI == EPI_ROWDOT) {
    p.k_splits = 1;
    p.kb_per_split = p.k_blocks;
    p.k_chunks = (int)ceil_div_ll(p.k_blocks, max_pass_kb);
    p.kb_per_chunk = (int)ceil_div_ll(p.k_blocks, p.k_chunks);
    p.k_chunks = (int)ceil_div_ll(p.k_blocks, p.kb_per_chunk);
    // few row blocks (self-influence: one "query" per batch): spread each block's n-tiles over the idle SMs
    const long long row_units = (long long)p.m_units * p.batch, groups = sm_count() / (CG * MC);
    p.n_splits = 1;
    if (row_units * 2 <= groups && p.n_blocks > 1 && (p.batch == 1 || p.out_bs >= p.M))
      p.n_splits = (int)(groups / row_units < p.n_blocks ? groups / row_units : p.n_blocks);
    p.nb_per_split = (int)ceil_div_ll(p.n_blocks, p.n_splits);
    p.n_splits = (int)ceil_div_ll(p.n_blocks, p.nb_per_split);
    p.use_atomic = p.n_splits > 1 ? 1 : 0;
    if (p.use_atomic && !p.accumulate && row_units > 0) {
      if (p.batch == 1) KFB_CUDA_TRY(cudaMemsetAsync(p.out_f32, 0, (size_t)p.M * 4, stream));
      else KFB_CUDA_TRY(cudaMemset2DAsync(p.out_f32, (size_t)p.out_bs * 4, 0, (size_t)p.M * 4, (size_t)p.batch, stream));
    }
    p.num_units = row_units * p.n_splits;
  } else if (EPI == EPI_REGACC && p.regacc_mode != REGACC_KCHUNK) {
    // cut the batch into chunks so that there are enough units to fill the machine
    long long chunks = ceil_div_ll(2LL * sm_count(), tiles);
    if (chunks > p.batch) chunks = p.batch;
    if (chunks < 1) chunks = 1;
    p.inner = (int)ceil_div_ll(p.batch, chunks);
    chunks = ceil_div_ll(p.batch, p.inner);
    p.k_splits = 1;
    p.kb_per_split = p.k_blocks;
    p.k_chunks = 1;
    p.kb_per_chunk = p.k_blocks;
    p.num_units = tiles * chunks;
  } else {
    if (p.k_splits < 1) p.k_splits = 1;
    if (p.k_splits > p.k_blocks) p.k_splits = p.k_blocks;
    p.kb_per_split = (int)ceil_div_ll(p.k_blocks, p.k_splits);
    p.k_splits = (int)ceil_div_ll(p.k_blocks, p.kb_per_split);
    p.use_atomic = p.k_splits > 1 ? 1 : 0;
    if (EPI == EPI_REGACC) {
      p.k_chunks = (int)ceil_div_ll(p.kb_per_split, max_pass_kb);
      p.kb_per_chunk = (int)ceil_div_ll(p.kb_per_split, p.k_chunks);
    } else {
      p.k_chunks = 1;
      p.kb_per_chunk = p.kb_per_split;
    }
    p.num_units = tiles * p.k_splits * p.batch;
  }
  if (p.num_units == 0) return KFB_OK;

  CUtensorMap ta_hi, ta_lo, tb_hi, tb_lo;
  KFB_TRY(make_tmap(&ta_hi, A.hi, A.cols, A.rows, A.batch, A.ld, A.batch_stride, BLOCK_K, 128));
  KFB_TRY(make_tmap(&tb_hi, B.hi, B.cols, B.rows, B.batch, B.ld, B.batch_stride, BLOCK_K, Cfg::FETCH_N));
  if (NSPLIT >= 2) {
    KFB_TRY(make_tmap(&ta_lo, A.lo, A.cols, A.rows, A.batch, A.ld, A.batch_stride, BLOCK_K, 128));
    KFB_TRY(make_tmap(&tb_lo, B.lo, B.cols, B.rows, B.batch, B.ld, B.batch_stride, BLOCK_K, Cfg::FETCH_N));
  } else {
    ta_lo = ta_hi;
    tb_lo = tb_hi;
  }
  // output planes through TMA bulk stores (32 x 32 boxes) when the STORE target is plain planes
  CUtensorMap to_hi = ta_hi, to_lo = ta_hi;
  if (EPI == EPI_ROWDOT) p.tma_store = 0;
  if (p.tma_store) {
    // one epilogue warpgroup: [32 rows x 64 columns] boxes out of 128B-swizzled staging tiles; two: plain 32 x 32 boxes
    const int box_cols = EW == 4 ? 64 : 32;
    KFB_TRY(make_tmap(&to_hi, p.out_hi, p.ldo_s, p.M, p.batch, p.ldo_s, p.out_bs_s, box_cols, 32, EW == 4));
    if (p.out_lo != nullptr)
      KFB_TRY(make_tmap(&to_lo, p.out_lo, p.ldo_s, p.M, p.batch, p.ldo_s, p.out_bs_s, box_cols, 32, EW == 4));
  }
  auto kernel = gemm_tc_kernel<BLOCK_N, BLOCK_K, NSPLIT, EPI, CG, MC, EW>;
  static bool attr_set = false;
  if (!attr_set) {
    KFB_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      Cfg::SMEM_BYTES));
    attr_set = true;
  }
  constexpr int CLUSTER = CG * MC;
  cudaLaunchConfig_t cfg{};
  cfg.blockDim = dim3(Cfg::THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CLUSTER;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = CLUSTER > 1 ? 1 : 0;
  // CTA groups resident at once: one CTA per SM; clusters of 4 may not tile every GPC, so ask the runtime
  long long groups = sm_count() / CLUSTER;
  if (p.max_groups > 0 && groups > p.max_groups) groups = p.max_groups;
  if (MC > 1) {
    static int max_clusters = -1;
    if (max_clusters < 0) {
      cfg.gridDim = dim3((unsigned)(groups * CLUSTER));
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, kernel, &cfg) != cudaSuccess || n <= 0) {
        cudaGetLastError();
        n = 0;
      }
      max_clusters = n;
    }
    // no (or hardly any) cluster of this size can be resident on this device / partition: the caller falls back to
    // the plain CTA-pair kernel
    if (max_clusters * CLUSTER * 4 < sm_count() * 3) return kRetryWithoutMulticast;
    if (groups > max_clusters) groups = max_clusters;
    g_resident_clusters.store(max_clusters);
    // cluster-group rendezvous of the fused pairwise kernel (see the producer): the grid becomes a whole number of
    // groups of m_units clusters, if that idles at most one cluster and every cluster gets several units
    p.sync_group = 0;
    if (EPI == EPI_ROWDOT && g_group_sync.load() != 0 && p.n_splits == 1 && p.m_units >= 2 && p.m_units <= 8 &&
        p.max_groups == 0) {
      const long long whole = (groups / p.m_units) * p.m_units;
      if (whole >= p.m_units && groups - whole <= 1 && whole / p.m_units <= 64 && p.num_units >= 4 * whole) {
        unsigned int* ctr = nullptr;  // per device: not cached (a process may drive several GPUs)
        KFB_CUDA_TRY(cudaGetSymbolAddress(reinterpret_cast<void**>(&ctr), g_rowdot_sync));
        KFB_CUDA_TRY(cudaMemsetAsync(ctr, 0, sizeof(unsigned int) * 64, stream));
        p.sync_ctr = ctr;
        p.sync_group = p.m_units;
        groups = whole;
      }
    }
  }
  const long long grid = (p.num_units < groups ? p.num_units : groups) * CLUSTER;
  cfg.gridDim = dim3((unsigned)grid);
  KFB_CUDA_TRY(cudaLaunchKernelEx(&cfg, kernel, ta_hi, ta_lo, tb_hi, tb_lo, to_hi, to_lo, p));
  count_launch();
  KFB_CUDA_TRY(cudaGetLastError());
  return KFB_OK;
}

static inline int pick_bn(int N, int cap) {
  int bn = N > 128 ? 256 : (N > 64 ? 128 : 64);
  return bn > cap ? cap : bn;
}

template <int EPI>
static int dispatch_tc(const kfb_split& A, const kfb_split& B, const GemmParams& p, int nsplit,
                       cudaStream_t stream) {
  // Tile width: as wide as N allows (wider tiles = more reuse of the A stage per MMA), except for
  // REGACC whose per-thread register accumulators limit it to 128.
  const bool wide_regacc = EPI == EPI_REGACC && nsplit == 2 && g_cta_pairs.load() != 0 && g_wide_regacc.load() != 0 &&
                           p.M > 128 && p.N > 128;
  const int bn = pick_bn(p.N, EPI == EPI_REGACC && !wide_regacc ? 128 : 256);
  // CTA pairs (M = 256 tiles) whenever the problem has at least two full 128-row tiles to pair up
  const bool pair = g_cta_pairs.load() != 0 && bn == 256 && p.M > 128;
  // Clusters of two CTA pairs on adjacent m-tiles fetch their common B tile once (TMA multicast): a quarter less
  // L2->SM operand traffic, which is what bounds the single-MMA (bf16) mode of the fused pairwise kernel.
  const bool mcast = g_multicast.load() != 0 && g_cta_pairs.load() != 0 && p.M > 256 && !p.symmetric && !p.batch_fastest;
  if constexpr (EPI == EPI_ROWDOT) {
    if (mcast && bn == 256) {
      // Clusters of four CTAs do not tile the GPCs: 33 of them are resident on a B200, 132 of 148 SMs.  When there are
      // many queries, the last few percent of them go to a concurrent launch of the plain CTA-pair kernel on a side
      // stream, capped at the idle SMs (launched second, so the cluster kernel is placed first).
      const int resident = g_resident_clusters.load();
      const int idle_pairs = resident > 0 ? (sm_count() - 4 * resident) / 2 : 0;
      long long side_q = 0;
      // Measured (target layer, Q = 1024): single-MMA (bf16) mode +2.4 % at 8-9 % of the queries; the 3-MMA mode runs
      // against the 1 kW power cap, where more active SMs only lower the clock (+-1 %), so it stays off there unless
      // kfb_set_idle_fill() asks for it.
      const int pct = g_idle_fill.load();
      if (idle_pairs >= 2 && p.batch >= 128 && p.row_group <= 1 && (long long)p.batch * ((p.M + 511) / 512) >= 4LL * resident &&
          (nsplit == 1 || pct > 0)) {
        // a pair of the side launch retires m-tiles at ~0.8 of the rate of a cluster's pair (no multicast: more
        // L2 -> SM operand traffic per tile), and both launches share the L2
        const double rate = 0.8;
        const double frac = pct >= 0 ? pct / 100.0 : idle_pairs * rate / (2.0 * resident + idle_pairs * rate);
        side_q = (long long)(p.batch * frac + 0.5);
        if (side_q < 8 || side_q * 2 > p.batch) side_q = 0;
      }
      if (side_q > 0) {
        const long long main_q = p.batch - side_q;
        GemmParams pm = p;
        pm.batch = (int)main_q;
        const kfb_split Am = batch_slice(A, 0, main_q), Bm = batch_slice(B, 0, main_q);
        SideStream* side = nullptr;
        KFB_TRY(side_stream(&side));
        KFB_CUDA_TRY(cudaEventRecord(side->fork, stream));
        KFB_CUDA_TRY(cudaStreamWaitEvent(side->stream, side->fork, 0));
        const int rc = nsplit == 2 ? launch_tc<256, 64, 2, EPI, 2, 2>(Am, Bm, pm, stream)
                                   : launch_tc<256, 64, 1, EPI, 2, 2>(Am, Bm, pm, stream);
        if (rc != KFB_OK && rc != kRetryWithoutMulticast) return rc;
        if (rc == KFB_OK) {
          GemmParams ps = p;
          ps.batch = (int)side_q;
          ps.out_f32 = p.out_f32 + main_q * p.out_bs;
          ps.g = p.g + main_q * p.g_bs;
          ps.max_groups = idle_pairs;
          const kfb_split As = batch_slice(A, main_q, side_q), Bs = batch_slice(B, main_q, side_q);
          KFB_TRY(nsplit == 2 ? (launch_tc<256, 64, 2, EPI, 2>(As, Bs, ps, side->stream))
                              : (launch_tc<256, 64, 1, EPI, 2>(As, Bs, ps, side->stream)));
          KFB_CUDA_TRY(cudaEventRecord(side->join, side->stream));
          KFB_CUDA_TRY(cudaStreamWaitEvent(stream, side->join, 0));
          return KFB_OK;
        }
        // no cluster launch on this device: everything goes to the pair kernel below
      } else {
        const int rc = nsplit == 2 ? launch_tc<256, 64, 2, EPI, 2, 2>(A, B, p, stream)
                                   : launch_tc<256, 64, 1, EPI, 2, 2>(A, B, p, stream);
        if (rc != kRetryWithoutMulticast) return rc;
      }
    }
  }
  if constexpr (EPI == EPI_STORE) {
    // long passes only (the flat pairwise GEMM: 16 k-blocks per pass; measured +8 % on the ResNet-9 conv layer at
    // Q = 1000, while the K = S per-sample-gradient GEMMs with their frequent epilogues lose a little)
    if (mcast && bn == 256 && p.kb_hint >= 16) {
      const int rc = nsplit == 2 ? launch_tc<256, 64, 2, EPI, 2, 2>(A, B, p, stream)
                                 : launch_tc<256, 64, 1, EPI, 2, 2>(A, B, p, stream);
      if (rc != kRetryWithoutMulticast) return rc;
    }
  }
  // (measured: NOT for the register-accumulating kernels — with a TMEM drain every two k-blocks, coupling the two
  // pairs' stage releases costs more than the saved traffic: Lambda sweep 3.09 -> 3.49 ms on the BERT FFN layer)
  if constexpr (EPI == EPI_REGACC) {
    // 256-wide register-accumulating tiles on CTA pairs: two epilogue warpgroups of 128 accumulator columns each
    if (wide_regacc && bn == 256) return launch_tc<256, 64, 2, EPI, 2, 1, 8>(A, B, p, stream);
  }
  if (nsplit == 2) {
    if (EPI != EPI_REGACC && bn == 256 && pair) return launch_tc<256, 64, 2, EPI, 2>(A, B, p, stream);
    if (EPI != EPI_REGACC && bn == 256) return launch_tc<256, 32, 2, EPI>(A, B, p, stream);
    // register-accumulating passes (128 accumulator columns per thread): CTA pairs halve the B-operand traffic
    if (EPI == EPI_REGACC && bn == 128 && g_cta_pairs.load() != 0 && p.M > 128) return launch_tc<128, 64, 2, EPI, 2>(A, B, p, stream);
    if (bn == 128) return launch_tc<128, 64, 2, EPI>(A, B, p, stream);
    return launch_tc<64, 64, 2, EPI>(A, B, p, stream);
  }
  if (EPI != EPI_REGACC && bn == 256 && pair) return launch_tc<256, 64, 1, EPI, 2>(A, B, p, stream);
  if (EPI != EPI_REGACC && bn == 256) return launch_tc<256, 64, 1, EPI>(A, B, p, stream);
  if (bn == 128) return launch_tc<128, 64, 1, EPI>(A, B, p, stream);
  return launch_tc<64, 64, 1, EPI>(A, B, p, stream);
}

static int launch_simt(const kfb_split& A, const kfb_split& B, GemmParams p, int epi, int nsplit,
                       cudaStream_t stream) {
  SimtOperand a{(const __nv_bfloat16*)A.hi, nsplit >= 2 ? (const __nv_bfloat16*)A.lo : nullptr, A.ld, A.batch_stride};
  SimtOperand b{(const __nv_bfloat16*)B.hi, nsplit >= 2 ? (const __nv_bfloat16*)B.lo : nullptr, B.ld, B.batch_stride};
  long long total = epi == KFB_EPI_STORE    ? (long long)p.batch * p.M * p.N
                    : epi == KFB_EPI_ROWDOT ? (long long)p.batch * p.M
                                            : (long long)p.M * p.N;
  if (total == 0) return KFB_OK;
  gemm_simt_kernel<<<(unsigned)ceil_div_ll(total, 128), 128, 0, stream>>>(a, b, p, epi);
  count_launch();
  KFB_CUDA_TRY(cudaGetLastError());
  return KFB_OK;
}

int gemm_nt(const kfb_split& A, const kfb_split& B, const kfb_epilogue& epi, int precision,
            int k_splits, cudaStream_t stream) {
  KFB_REQUIRE(A.hi != nullptr && B.hi != nullptr, "gemm_nt: null operand");
  KFB_REQUIRE(A.cols == B.cols, "gemm_nt: contraction lengths differ (%lld vs %lld)",
              (long long)A.cols, (long long)B.cols);
  KFB_REQUIRE(A.batch >= 1 && B.batch >= 1 && (A.batch == B.batch || A.batch == 1 || B.batch == 1),
              "gemm_nt: incompatible batch counts %lld / %lld", (long long)A.batch, (long long)B.batch);
  KFB_REQUIRE(A.rows < (1LL << 31) && B.rows < (1LL << 31) && A.cols < (1LL << 31),
              "gemm_nt: dimension too large");
  const bool strict = precision == KFB_PREC_STRICT;
  const int nsplit = precision == KFB_PREC_BF16 ? 1 : 2;
  KFB_REQUIRE(nsplit == 1 || (A.lo != nullptr && B.lo != nullptr), "gemm_nt: missing lo planes");
  KFB_REQUIRE(!strict || (A.absmax != nullptr && B.absmax != nullptr),
              "gemm_nt: KFB_PREC_STRICT operands carry an absmax word (build them with KFB_PREC_STRICT)");
  KFB_REQUIRE(strict || (A.absmax == nullptr && B.absmax == nullptr),
              "gemm_nt: a KFB_PREC_STRICT operand (scaled FP16 planes) was passed to a bf16-plane contraction; build the "
              "operand with the precision of the stage that consumes it");
  GemmParams p{};
  p.f16 = strict ? 1 : 0;
  p.a_absmax = A.absmax;
  p.b_absmax = B.absmax;
  p.M = (int)A.rows;
  p.N = (int)B.rows;
  p.K = (int)A.cols;
  p.batch = (int)(A.batch > B.batch ? A.batch : B.batch);
  p.a_batched = A.batch > 1;
  p.b_batched = B.batch > 1;
  p.out_f32 = epi.out_f32;
  p.ldo = epi.ldo;
  p.out_bs = epi.out_batch_stride;
  p.out_hi = (__nv_bfloat16*)epi.out_split.hi;
  p.out_lo = (__nv_bfloat16*)epi.out_split.lo;
  p.ldo_s = epi.out_split.ld;
  p.out_bs_s = epi.out_split.batch_stride;
  p.mul = epi.mul;
  p.ldmul = epi.ldmul;
  p.transpose_out = epi.transpose_out;
  p.square = epi.square;
  p.accumulate = epi.accumulate;
  p.alpha = epi.alpha;
  p.g = epi.g;
  p.ldg = epi.ldg;
  p.g_bs = epi.g_batch_stride;
  p.row_group = epi.row_group > 1 ? (int)epi.row_group : 1;
  p.reduce_sq = epi.kind == KFB_EPI_STORE ? epi.reduce_sq : 0;
  p.col_group = epi.kind == KFB_EPI_STORE ? (int)epi.col_group : 0;
  p.k_splits = 1;
  p.k_chunks = 1;
  // strict mode: drain TMEM every kStrictPassK contraction elements (24 truncating accumulations per pass)
  // (bf16 mode: one MMA per k-step and operands that are themselves only good to 2^-9, so a pass may run 4x longer)
  p.max_pass_k = strict ? g_strict_pass_k.load() : (nsplit == 1 ? 4 * kMaxPassK : kMaxPassK);
  if (p.M == 0 || p.N == 0 || p.batch == 0) return KFB_OK;
  KFB_REQUIRE(p.K > 0, "gemm_nt: empty contraction");

  if (epi.kind == KFB_EPI_STORE) {
    KFB_REQUIRE(p.out_f32 != nullptr || p.out_hi != nullptr, "gemm_nt: STORE without an output");
    p.vec_ok = 1;
    p.f32_vec4 = (p.out_f32 && !((reinterpret_cast<uintptr_t>(p.out_f32) & 15) || p.ldo % 4 || p.out_bs % 4)) ? 1 : 0;
    p.mul_vec4 = (p.mul && !((reinterpret_cast<uintptr_t>(p.mul) & 15) || p.ldmul % 4)) ? 1 : 0;
    if (p.out_hi && ((reinterpret_cast<uintptr_t>(p.out_hi) & 15) || p.ldo_s % 8 || p.out_bs_s % 8 ||
                     (p.out_lo && (reinterpret_cast<uintptr_t>(p.out_lo) & 15))))
      p.vec_ok = 0;
    p.zero_pad = (p.out_hi != nullptr && !p.transpose_out) ? 1 : 0;
    p.tma_store = (p.out_hi != nullptr && p.out_f32 == nullptr && p.mul == nullptr && !p.transpose_out && !p.reduce_sq &&
                   p.col_group == 0 && p.vec_ok && p.ldo_s >= 64 && g_tma_store.load() != 0) ? 1 : 0;
    if (p.col_group > 0) {
      KFB_REQUIRE(p.out_hi != nullptr && p.out_f32 == nullptr && p.mul == nullptr && !p.transpose_out && p.batch == 1 &&
                      p.col_group % 8 == 0 && p.N % p.col_group == 0 && p.ldo_s == p.col_group,
                  "gemm_nt: col_group needs a plane-only, non-transposed target with ld == col_group (multiple of 8)");
      p.zero_pad = 0;
    }
    p.batch_fastest = (p.mul != nullptr && p.batch > 1) ? 1 : 0;
    if (p.reduce_sq) {
      KFB_REQUIRE(p.out_f32 != nullptr && p.out_hi == nullptr, "gemm_nt: reduce_sq needs out_f32 only");
      KFB_REQUIRE(p.K <= p.max_pass_k, "gemm_nt: reduce_sq needs the contraction to fit one TMEM pass (K <= %d)",
                  p.max_pass_k);
    }
  } else if (epi.kind == KFB_EPI_ROWDOT) {
    KFB_REQUIRE(p.out_f32 != nullptr && p.g != nullptr, "gemm_nt: ROWDOT needs out_f32 and g");
    KFB_REQUIRE(A.batch == 1 || A.batch == p.batch, "gemm_nt: ROWDOT batch mismatch");
    p.g_vec4 = ((reinterpret_cast<uintptr_t>(p.g) & 15) == 0 && p.ldg % 4 == 0 && p.g_bs % 4 == 0) ? 1 : 0;
    KFB_REQUIRE(p.row_group == 1 || p.accumulate, "gemm_nt: ROWDOT row groups add into the output (accumulate = 1)");
  } else if (epi.kind == KFB_EPI_SQACC) {
    KFB_REQUIRE(p.out_f32 != nullptr, "gemm_nt: SQACC needs out_f32");
    // sum over the batch of squares: register accumulation, atomically added to the target
    p.regacc_mode = REGACC_BATCH;
    p.use_atomic = 1;
    p.out_hi = nullptr;
    p.out_lo = nullptr;
    p.mul = nullptr;
    p.square = 0;
    p.transpose_out = 0;
    p.out_bs = 0;
  } else {
    KFB_REQUIRE(false, "gemm_nt: unknown epilogue %d", epi.kind);
  }

  if (g_backend.load() == 1) return launch_simt(A, B, p, epi.kind, nsplit, stream);
  if (epi.kind == KFB_EPI_STORE && p.reduce_sq && p.batch > 1) {
    // self-influence: one scalar per batch entry.  The factor tile lives in registers across the batch entries of a
    // unit (register-accumulating kernel, batch mode) instead of being re-read from L2 for every example and chunk.
    p.regacc_mode = REGACC_BATCH_DOT;
    p.factor = p.mul;
    p.ldfactor = p.ldmul;
    p.tma_store = 0;
    p.zero_pad = 0;
    p.batch_fastest = 0;
    return dispatch_tc<EPI_REGACC>(A, B, p, nsplit, stream);
  }
  if (epi.kind == KFB_EPI_STORE && p.mul != nullptr && p.batch > 1 && p.out_hi != nullptr && p.out_f32 == nullptr &&
      !p.transpose_out && !p.square && p.col_group == 0 && p.K <= p.max_pass_k && !strict) {
    // an elementwise factor shared by a batch of short products stored as planes (query preconditioning of S > 1
    // layers): the factor tile stays in registers across the batch entries of a unit and the planes leave through TMA
    // stores, instead of one L2 round trip per 16 rows of every chunk and 16-byte scattered stores
    p.regacc_mode = REGACC_BATCH_MUL;
    p.factor = p.mul;
    p.ldfactor = p.ldmul;
    p.mul = nullptr;
    p.mul_vec4 = 0;
    p.batch_fastest = 0;
    p.tma_store = (p.vec_ok && p.ldo_s >= 64 && g_tma_store.load() != 0) ? 1 : 0;
    return dispatch_tc<EPI_REGACC>(A, B, p, nsplit, stream);
  }
  if (epi.kind == KFB_EPI_ROWDOT) return dispatch_tc<EPI_ROWDOT>(A, B, p, nsplit, stream);
  if (epi.kind == KFB_EPI_SQACC) return dispatch_tc<EPI_REGACC>(A, B, p, nsplit, stream);

  // STORE: one TMEM pass if the contraction is short.  A longer contraction is cut into passes of at most
  // max_pass_k elements: when the target is a plain accumulated fp32 matrix the passes are separate units of the
  // wide (256 x 256, CTA-pair) STORE kernel combined by fp32 atomic adds in L2 (the same mechanism that splits the
  // contraction across CTAs when there are too few output tiles to fill the machine); otherwise they are summed in
  // registers by the REGACC kernel (128-wide tiles).
  const bool long_k = p.K > p.max_pass_k;
  const bool splittable = p.out_hi == nullptr && p.mul == nullptr && !p.square && p.accumulate && !p.reduce_sq;
  if (epi.symmetric) {
    KFB_REQUIRE(A.hi == B.hi && A.rows == B.rows && A.ld == B.ld && p.batch == 1 && splittable && !p.transpose_out,
                "gemm_nt: symmetric needs A == B and a plain accumulated fp32 target");
    p.symmetric = 1;
  }
  if (splittable) {
    // Atomically combined passes are cheap (one 256 x 256 fp32 reduction per pair and pass), so in the 3-MMA mode
    // they are kept to kAtomicPassK elements: sums of squares (covariance diagonals) see the accumulator's
    // truncation bias coherently, and halving the chain halves it.
    const long long passes = ceil_div_ll(p.K, strict ? g_strict_pass_k.load() : (nsplit == 2 ? kAtomicPassK : p.max_pass_k));
    if (k_splits == 0) {
      const int bn = pick_bn(p.N, 256);
      long long tiles = ceil_div_ll(p.M, bn == 256 && p.M > 128 ? 256 : 128) * ceil_div_ll(p.N, bn) * p.batch;
      if (p.symmetric) tiles = (tiles + 1) / 2;
      long long want = ceil_div_ll(sm_count(), tiles);
      const long long max_want = strict ? ceil_div_ll(p.K, 1024) : passes;
      if (want > max_want) want = max_want;  // never make a split shorter than one full pass
      k_splits = want < 1 ? 1 : (int)want;
    }
    if (!strict && k_splits < passes) k_splits = (int)passes;
    p.k_splits = k_splits;
    p.kb_hint = (int)(ceil_div_ll(p.K, 64) / (k_splits > 0 ? k_splits : 1));
    if (!strict) return dispatch_tc<EPI_STORE>(A, B, p, nsplit, stream);
  }
  if (long_k) {
    p.regacc_mode = REGACC_KCHUNK;
    p.symmetric = 0;
    return dispatch_tc<EPI_REGACC>(A, B, p, nsplit, stream);
  }
  p.kb_hint = (int)ceil_div_ll(p.K, 64);
  return dispatch_tc<EPI_STORE>(A, B, p, nsplit, stream);
}

}  // namespace kfb

extern "C" {

int kfb_gemm_nt(const kfb_split* A, const kfb_split* B, const kfb_epilogue* epi, int precision,
                void* stream) {
  if (A == nullptr || B == nullptr || epi == nullptr) {
    kfb::set_error("kfb_gemm_nt: null argument");
    return KFB_ERR_INVALID;
  }
  return kfb::gemm_nt(*A, *B, *epi, precision, 0, static_cast<cudaStream_t>(stream));
}

int kfb_set_gemm_backend(int backend) {
  if (backend != 0 && backend != 1) {
    kfb::set_error("kfb_set_gemm_backend: backend must be 0 (tcgen05) or 1 (simt debug)");
    return KFB_ERR_INVALID;
  }
  kfb::g_backend.store(backend);
  return KFB_OK;
}

int64_t kfb_launch_count(void) { return kfb::g_launches.load(); }

int kfb_set_cta_pairs(int enable) {
  kfb::g_cta_pairs.store(enable ? 1 : 0);
  return KFB_OK;
}

int kfb_set_multicast(int enable) {
  kfb::g_multicast.store(enable < 0 ? 0 : enable);
  return KFB_OK;
}

int kfb_set_group_sync(int enable) {
  kfb::g_group_sync.store(enable ? 1 : 0);
  return KFB_OK;
}

int kfb_set_idle_fill(int percent) {
  kfb::g_idle_fill.store(percent < 0 ? -1 : (percent > 50 ? 50 : percent));
  return KFB_OK;
}

int kfb_set_wide_regacc(int enable) {
  kfb::g_wide_regacc.store(enable ? 1 : 0);
  return KFB_OK;
}

int kfb_set_strict_pass_k(int k) {
  if (k < 64 || k % 64 != 0) {
    kfb::set_error("kfb_set_strict_pass_k: the pass length must be a positive multiple of 64");
    return KFB_ERR_INVALID;
  }
  kfb::g_strict_pass_k.store(k);
  return KFB_OK;
}

int kfb_set_tma_store(int enable) {
  kfb::g_tma_store.store(enable ? 1 : 0);
  return KFB_OK;
}

}  // extern "C"
