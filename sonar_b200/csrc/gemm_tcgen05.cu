// bf16 x bf16 -> fp32-accumulate GEMM on the 5th-gen tensor cores (tcgen05 + TMEM),
// fed by TMA, warp-specialised, persistent.  C = epi(A[M,K] · W[N,K]^T + bias[N]).
//
// This is the contraction behind every `F.linear` of the SONAR text encoder layer
// (reference wiring: sonar/models/sonar_text/factory.py:130-153 -- q/k/v/out
// projections and the 1024->8192->1024 ReLU FFN; nn.Linear weight layout [out,in]).
//
// Tile shape per CTA pair (cta_group::2): 256(M) x 256(N) x 64(K) per pipeline stage,
//   CTA r loads A rows [128r,128r+128) and W rows [128r,128r+128) of the tile;
//   one tcgen05.mma.cta_group::2 (M=256,N=256,K=16) x4 per stage, issued by ONE thread
//   of the leader CTA; each CTA's TMEM holds its 128 rows x 256 fp32 columns, double
//   buffered (2 x 256 = all 512 columns) so the epilogue of tile i overlaps the
//   mainloop of tile i+1.  cta_group::1 variant: 128 x 256 x 64 per CTA.
// Warp roles (384 threads): w0 TMA producer, w1 MMA issuer, w2 TMEM alloc/dealloc, w3 idle,
//   w4-7 and w8-11 two epilogue warpgroups (TMEM lane quarter = warp % 4) that take alternate column chunks of every
//   accumulator tile (TMEM -> regs -> bias/ReLU/residual -> swizzled smem -> TMA store): with K = 1024 a 256 x 256 tile is
//   only ~4 us of MMA, which ONE warpgroup's epilogue does not fit under.
// LayerNorm folding (LnFold, sonar_b200_internal.h): a consumer GEMM scales its accumulator rows by the LayerNorm
// statistics of its input; the residual-stream GEMMs emit those statistics and the bf16 copy of the stream.
// Operand smem layout: K-major, 128-byte rows, SWIZZLE_128B (TMA writes it, UMMA reads it).

#include "common.cuh"
#include "sonar_b200_internal.h"

#include <math_constants.h>

namespace sb {

template <int kCtaGroup, int kEpiGroups = 2>
struct GemmCfg {
  static constexpr int BLOCK_M = 128;                 // rows per CTA
  static constexpr int BLOCK_N = 256;                 // UMMA N
  static constexpr int BLOCK_K = 64;                  // 128 bytes of bf16 = one swizzle atom
  static constexpr int LOAD_N = BLOCK_N / kCtaGroup;  // W rows each CTA stages
  static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
  static constexpr int B_BYTES = LOAD_N * BLOCK_K * 2;
  // 32 KB (cta_group::2) / 48 KB (cta_group::1) per mainloop stage; the staging buffers of the second epilogue warpgroup
  // cost one stage
  static constexpr int STAGES = (kCtaGroup == 2) ? (kEpiGroups == 2 ? 5 : 6) : (kEpiGroups == 2 ? 3 : 4);
  static constexpr int CD_STAGE_BYTES = 128 * 128;  // 128 rows x 128 B
  static constexpr int EPI_GROUPS = kEpiGroups;     // epilogue warpgroups (4 warps each), alternating column chunks
  static constexpr int CD_STAGES = 2;               // staging buffers per epilogue warpgroup
  static constexpr int BAR_BYTES = 256;
  static constexpr int SMEM_BYTES =
      STAGES * (A_BYTES + B_BYTES) + EPI_GROUPS * CD_STAGES * CD_STAGE_BYTES + BAR_BYTES + 1024 /*align slack*/;
  static constexpr int TMEM_COLS = 512;
  static constexpr int THREADS = 128 + 128 * EPI_GROUPS;
};

// Tile scheduler shared by the three warp roles (each role walks an identical copy).
// Default: tiles are dealt round-robin with n fastest, so the clusters running concurrently share A rows
// through L2 and every weight tile stays L2-resident.
// Sweep (top-k epilogue): a work item is (m-block, n-chunk); the cluster walks all n tiles of the chunk so the
// running top-k / log-sum-exp of a row lives in the epilogue thread's registers for the whole item.  Items of one
// m-block sit next to each other, and the same chunk of different m-blocks runs concurrently on different
// clusters, so W streams through L2 once.
template <bool kSweep>
struct TileSched {
  int num_m_tiles, num_n_tiles, cluster_id, num_clusters, n_chunks, tiles_per_chunk;
  int k_splits = 1;  // split-K (accumulate epilogue, few tiles): item = split * tiles + tile, `chunk` returns the split
  int i = 0, j = 0;
  __device__ __forceinline__ bool next(int& m_blk, int& n_blk, int& chunk, bool& first, bool& last) {
    if constexpr (!kSweep) {
      const int item = cluster_id + i * num_clusters;
      ++i;
      const int tiles = num_m_tiles * num_n_tiles;
      if (item >= tiles * k_splits) return false;
      const int tile = (k_splits == 1) ? item : item % tiles;
      chunk = (k_splits == 1) ? 0 : item / tiles;
      m_blk = tile / num_n_tiles;
      n_blk = tile % num_n_tiles;
      first = last = true;
      return true;
    } else {
      for (;;) {
        const int item = cluster_id + i * num_clusters;
        if (item >= num_m_tiles * n_chunks) return false;
        m_blk = item / n_chunks;
        chunk = item % n_chunks;
        const int n_begin = chunk * tiles_per_chunk;
        const int cnt = min(tiles_per_chunk, num_n_tiles - n_begin);
        if (j < cnt) {
          n_blk = n_begin + j;
          first = (j == 0);
          last = (j == cnt - 1);
          ++j;
          return true;
        }
        j = 0;
        ++i;
      }
    }
  }
};

// insert (v, idx) into a descending list kept in registers; equal values keep the earlier entry first
template <int KC>
__device__ __forceinline__ void topk_insert(float (&tv)[KC], int (&ti)[KC], float v, int idx) {
  float cv = v;
  int ci = idx;
#pragma unroll
  for (int p = 0; p < KC; ++p) {
    const bool gt = cv > tv[p];
    const float ov = tv[p];
    const int oi = ti[p];
    tv[p] = gt ? cv : ov;
    ti[p] = gt ? ci : oi;
    cv = gt ? ov : cv;
    ci = gt ? oi : ci;
  }
}

template <int kCtaGroup, int kEpi, typename OutT, int kEpiGroups>
__global__ void __launch_bounds__(GemmCfg<kCtaGroup, kEpiGroups>::THREADS, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                         const __grid_constant__ CUtensorMap tm_c, const float* __restrict__ bias,
                         const OutT* residual, long long ldr, int M, int N, int K, float* __restrict__ cand_val,
                         int* __restrict__ cand_idx, float* __restrict__ lse_part, int n_chunks, const LnFold lf,
                         const ColFilter cf, int k_splits, int* __restrict__ splitk_flags) {
  constexpr bool kSweep = (kEpi == EPI_TOPK);
  using Cfg = GemmCfg<kCtaGroup, kEpiGroups>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem_a + Cfg::STAGES * Cfg::A_BYTES;
  uint8_t* smem_cd = smem_b + Cfg::STAGES * Cfg::B_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_cd + Cfg::EPI_GROUPS * Cfg::CD_STAGES * Cfg::CD_STAGE_BYTES);
  uint64_t* empty_bar = full_bar + Cfg::STAGES;
  uint64_t* tmem_full_bar = empty_bar + Cfg::STAGES;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = (kCtaGroup == 2) ? cluster_ctarank() : 0u;
  const bool is_leader = (cta_rank == 0);

  if (kCtaGroup == 2) cluster_sync_all();  // both CTAs resident before the paired TMEM allocation

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_b);
    tma_prefetch_desc(&tm_c);
  }
  if (warp_idx == 1 && lane == 0) {
    for (int i = 0; i < Cfg::STAGES; ++i) {
      mbar_init(&full_bar[i], 1);   // leader's arrive.expect_tx (+ TMA bytes of both CTAs)
      mbar_init(&empty_bar[i], 1);  // one tcgen05.commit (multicast to both CTAs)
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full_bar[i], 1);               // one tcgen05.commit
      mbar_init(&tmem_empty_bar[i], 4 * Cfg::EPI_GROUPS * kCtaGroup);  // one arrive per epilogue warp of every CTA
    }
    fence_mbar_init();
  }
  if (warp_idx == 2) tmem_alloc<kCtaGroup>(tmem_ptr_smem, Cfg::TMEM_COLS);
  tc_fence_before();
  if (kCtaGroup == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const int tile_m = Cfg::BLOCK_M * kCtaGroup;
  const int num_m_tiles = (M + tile_m - 1) / tile_m;
  const int num_n_tiles = (N + Cfg::BLOCK_N - 1) / Cfg::BLOCK_N;  // N tail only with the top-k epilogue
  const int num_kb = K / Cfg::BLOCK_K;
  const int cluster_id = blockIdx.x / kCtaGroup;
  const int num_clusters = gridDim.x / kCtaGroup;
  TileSched<kSweep> sched{num_m_tiles, num_n_tiles, cluster_id, num_clusters, n_chunks,
                          (num_n_tiles + n_chunks - 1) / n_chunks, (kEpi == EPI_BIAS_ACCUM) ? k_splits : 1};
  // split-K: split s of a tile runs k-blocks [s * num_kb / k_splits, (s + 1) * num_kb / k_splits)
  auto kb_begin = [&](int split) { return (kEpi == EPI_BIAS_ACCUM && k_splits > 1) ? split * num_kb / k_splits : 0; };
  auto kb_end = [&](int split) { return (kEpi == EPI_BIAS_ACCUM && k_splits > 1) ? (split + 1) * num_kb / k_splits : num_kb; };
  int m_blk, n_blk, chunk;
  bool first_in_item, last_in_item;

  if (warp_idx == 0) {
    // ===================== TMA producer (one thread) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      while (sched.next(m_blk, n_blk, chunk, first_in_item, last_in_item)) {
        const int m0 = m_blk * tile_m + int(cta_rank) * Cfg::BLOCK_M;
        const int n0 = n_blk * Cfg::BLOCK_N + int(cta_rank) * Cfg::LOAD_N;
        const int kb1 = kb_end(chunk);
        for (int kb = kb_begin(chunk); kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (is_leader) mbar_arrive_expect_tx(&full_bar[stage], (Cfg::A_BYTES + Cfg::B_BYTES) * kCtaGroup);
          if (kCtaGroup == 2) {
            tma_load_2d_cta2(smem_a + stage * Cfg::A_BYTES, &tm_a, &full_bar[stage], kb * Cfg::BLOCK_K, m0);
            tma_load_2d_cta2(smem_b + stage * Cfg::B_BYTES, &tm_b, &full_bar[stage], kb * Cfg::BLOCK_K, n0);
          } else {
            tma_load_2d(smem_a + stage * Cfg::A_BYTES, &tm_a, &full_bar[stage], kb * Cfg::BLOCK_K, m0);
            tma_load_2d(smem_b + stage * Cfg::B_BYTES, &tm_b, &full_bar[stage], kb * Cfg::BLOCK_K, n0);
          }
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp_idx == 1) {
    // ===================== MMA issuer (leader CTA; whole warp walks the loop, one elected lane issues) ===========
    // The loop state is kept warp-uniform so the compiler can hold descriptors in uniform registers: the
    // issue path must stay well under the 512 tensor-core cycles one k-block takes.
    if (is_leader) {
      constexpr uint32_t idesc = umma_idesc_bf16_f32(Cfg::BLOCK_M * kCtaGroup, Cfg::BLOCK_N);
      constexpr uint32_t desc_hi = (1024u >> 4) | (1u << 14) | (2u << 29);  // SBO=1024 B, version 1, SWIZZLE_128B
      const uint32_t a_lo0 = (smem_u32(smem_a) >> 4) & 0x3FFFu;
      const uint32_t b_lo0 = (smem_u32(smem_b) >> 4) & 0x3FFFu;
      int stage = 0;
      uint32_t phase = 0;
      for (uint32_t iter = 0; sched.next(m_blk, n_blk, chunk, first_in_item, last_in_item); ++iter) {
        const uint32_t acc = iter & 1, acc_phase = (iter >> 1) & 1;
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);  // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * Cfg::BLOCK_N;
        const int kb0 = kb_begin(chunk), kb1 = kb_end(chunk);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);  // TMA bytes of both CTAs have landed
          tc_fence_after();
          if (elect_one()) {
            const uint32_t a_lo = a_lo0 + uint32_t(stage) * (Cfg::A_BYTES >> 4);
            const uint32_t b_lo = b_lo0 + uint32_t(stage) * (Cfg::B_BYTES >> 4);
#pragma unroll
            for (int k = 0; k < Cfg::BLOCK_K / 16; ++k) {
              // +32 bytes (= 16 bf16) along K inside the 128B swizzle atom -> +2 in the >>4 address field
              const uint64_t a_desc = (uint64_t(desc_hi) << 32) | uint64_t(a_lo + 2 * k);
              const uint64_t b_desc = (uint64_t(desc_hi) << 32) | uint64_t(b_lo + 2 * k);
              umma_bf16<kCtaGroup>(d_tmem, a_desc, b_desc, idesc, ((kb - kb0) | k) != 0 ? 1u : 0u);
            }
            umma_commit<kCtaGroup>(&empty_bar[stage]);  // smem slot reusable once these MMAs retire
            if (kb == kb1 - 1) umma_commit<kCtaGroup>(&tmem_full_bar[acc]);
          }
          __syncwarp();
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp_idx >= 4) {
    // ===================== epilogue: 2 warpgroups x 4 warps (each group spans the 128 TMEM lanes) =====================
    const int wg = (warp_idx - 4) >> 2;   // epilogue warpgroup: takes the column chunks c with c % 2 == wg
    const int ew = warp_idx & 3;          // TMEM lane quarter
    const int row_in_tile = ew * 32 + lane;
    constexpr int CHUNK_COLS = 128 / int(sizeof(OutT));  // one 128-byte smem row per output row
    constexpr int NUM_CHUNKS = Cfg::BLOCK_N / CHUNK_COLS;
    constexpr int SUBS = CHUNK_COLS / 32;
    static_assert(NUM_CHUNKS % Cfg::EPI_GROUPS == 0, "column chunks must split evenly over the epilogue warpgroups");
    uint8_t* smem_cd_wg = smem_cd + wg * Cfg::CD_STAGES * Cfg::CD_STAGE_BYTES;
    const uint32_t bar_id = 1 + wg;  // named barrier of this warpgroup
    int cd_stage = 0;
    constexpr int KC = kTopkCandidates;
    [[maybe_unused]] float tv[KC];
    [[maybe_unused]] int ti[KC];
    [[maybe_unused]] float run_max = -CUDART_INF_F, run_sum = 0.f;  // online log-sum-exp of the row (optional)
    [[maybe_unused]] int q_n = 0;  // top-k sweep: candidates pending in this thread's shared-memory queue
    // ---- LayerNorm folding, consumer side: the (mean, M2) partials of this thread's input row are fetched ONE TILE AHEAD
    // (a lookahead copy of the scheduler names the next tile) so their latency hides under the current tile's epilogue ----
    [[maybe_unused]] bool fold_in = false;
    [[maybe_unused]] float2 part_next[8];
    [[maybe_unused]] TileSched<kSweep> sched_ahead = sched;
    [[maybe_unused]] auto fetch_parts = [&]() {
      int mb, nb, ch;
      bool f0, f1;
      if (!sched_ahead.next(mb, nb, ch, f0, f1)) return;
      const int r = mb * tile_m + int(cta_rank) * Cfg::BLOCK_M + row_in_tile;
      if (r < M) {
        const float2* sp = reinterpret_cast<const float2*>(lf.stats_in) + (long long)r * lf.chunks;
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (i < lf.chunks) part_next[i] = sp[i];
      }
    };
    if constexpr (kEpi == EPI_BIAS || kEpi == EPI_BIAS_RELU) {
      fold_in = lf.stats_in != nullptr;
      if (fold_in) fetch_parts();
    }
    for (uint32_t iter = 0; sched.next(m_blk, n_blk, chunk, first_in_item, last_in_item); ++iter) {
      const int m0 = m_blk * tile_m + int(cta_rank) * Cfg::BLOCK_M;
      const int n0 = n_blk * Cfg::BLOCK_N;
      const uint32_t acc = iter & 1, acc_phase = (iter >> 1) & 1;
      const int grow = m0 + row_in_tile;
      [[maybe_unused]] float ln_mean = 0.f, ln_rstd = 1.f;
      if constexpr (kEpi == EPI_BIAS || kEpi == EPI_BIAS_RELU) {
        if (fold_in) {
          if (grow < M) {  // Chan merge of `chunks` partials of kLnPartCols columns each
            float msum = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i)
              if (i < lf.chunks) msum += part_next[i].x;
            ln_mean = msum / float(lf.chunks);
            float m2 = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i)
              if (i < lf.chunks) { const float dm = part_next[i].x - ln_mean; m2 += part_next[i].y + float(kLnPartCols) * dm * dm; }
            ln_rstd = 1.0f / sqrtf(m2 / float(kLnPartCols * lf.chunks) + lf.eps);
          }
          fetch_parts();  // for the next tile
        }
      }
      // ---- producer side: the first residual chunk of this thread's row is fetched while the MMAs still run ----
      [[maybe_unused]] float4 rnext[8];
      [[maybe_unused]] float st_n = 0.f, st_mean = 0.f, st_m2 = 0.f;
      if constexpr (kEpi == EPI_BIAS_RESIDUAL_STATS) {
#pragma unroll
        for (int q = 0; q < 8; ++q) rnext[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (grow < M) {
          const float4* rp = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(residual) + (long long)grow * ldr +
                                                             n0 + wg * CHUNK_COLS);
#pragma unroll
          for (int q = 0; q < 8; ++q) rnext[q] = rp[q];
        }
      }
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      if constexpr (kEpi == EPI_TOPK) {
        // ---- running per-row top-KC over the whole sweep of n tiles (no C matrix is ever written) ----
        if (first_in_item) {
#pragma unroll
          for (int p = 0; p < KC; ++p) { tv[p] = -CUDART_INF_F; ti[p] = -1; }
          run_max = -CUDART_INF_F;
          run_sum = 0.f;
        }
        // each epilogue warpgroup sweeps ITS 32-column sub-chunks (c % 2 == wg) and keeps its own list: a row ends up with
        // EPI_GROUPS lists per n-chunk, merged by the caller's next kernel.
        // Candidates are not inserted where they are found: a thread PUSHES (value, column) of every element above its
        // current 16th value into a small queue in shared memory (the idle output staging buffer) and the warp drains all
        // 32 queues together when one fills up.  Inserting in place costs the whole warp ~80 instructions whenever ANY lane
        // qualifies (32 rows with independent insertion events -> ~2500 instructions per tile per warp on random logits);
        // drained together, a round of insertions serves every lane that has one pending.  Same elements, same order per
        // row, same result bit for bit.
        const bool cf_on = cf.thr != nullptr && grow < M;
        constexpr int kQ = 16;  // queue entries per thread; drained once any lane holds more than kQ - 8
        uint2* queue = reinterpret_cast<uint2*>(smem_cd_wg) + row_in_tile;  // entry e of this thread at queue[e * 128]
        auto drain = [&]() {
          const int rounds = __reduce_max_sync(0xffffffffu, q_n);
          for (int r = 0; r < rounds; ++r) {
            if (r < q_n) {
              const uint2 e = queue[r * 128];
              const float x = __uint_as_float(e.x);
              if (x > tv[KC - 1]) topk_insert<KC>(tv, ti, x, int(e.y));
            }
          }
          q_n = 0;
        };
#pragma unroll 1
        for (int c = wg; c < Cfg::BLOCK_N / 32; c += Cfg::EPI_GROUPS) {
          uint32_t v[32];
          tmem_ld_32x32(tmem_base + (uint32_t(ew * 32) << 16) + acc * Cfg::BLOCK_N + c * 32, v);
          tmem_ld_wait();
          const int gcol = n0 + c * 32;
          if (gcol + 32 > N) {  // ragged last tile: columns >= N are zero-filled operands, not candidates
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (gcol + j >= N) v[j] = __float_as_uint(-CUDART_INF_F);
          }
          // maxima of the four 8-column groups: shared by the candidate pre-filter and the log-sum-exp
          float gm[4];
#pragma unroll
          for (int g8 = 0; g8 < 4; ++g8) {
            float mx = __uint_as_float(v[g8 * 8]);
#pragma unroll
            for (int j = 1; j < 8; ++j) mx = fmaxf(mx, __uint_as_float(v[g8 * 8 + j]));
            gm[g8] = mx;
          }
          if (lse_part != nullptr) {  // online log-sum-exp over every column of the row (fp32, like log_softmax)
            const float cm = fmaxf(fmaxf(gm[0], gm[1]), fmaxf(gm[2], gm[3]));
            if (cm > run_max) {
              run_sum *= __expf(run_max - cm);  // exp(-inf) = 0 on the first chunk
              run_max = cm;
            }
            if (run_max > -CUDART_INF_F) {
              float cs = 0.f;
#pragma unroll
              for (int j = 0; j < 32; ++j) cs += __expf(__uint_as_float(v[j]) - run_max);
              run_sum += cs;
            }
          }
          const float thr = tv[KC - 1];  // (stale between drains: a few extra pushes, rejected when drained)
          float t8[4] = {CUDART_INF_F, CUDART_INF_F, CUDART_INF_F, CUDART_INF_F};  // column filter off / rows beyond M: never hit
          if (cf_on) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(cf.thr8 + (gcol >> 3)));
            t8[0] = t.x; t8[1] = t.y; t8[2] = t.z; t8[3] = t.w;
          }
#pragma unroll
          for (int g8 = 0; g8 < 4; ++g8) {
            if (gm[g8] > t8[g8]) {  // column filter: some element of these 8 is above its COLUMN's threshold (rare)
              const float4 t0 = *reinterpret_cast<const float4*>(cf.thr + gcol + g8 * 8);  // (not __ldg: columns get closed)
              const float4 t1 = *reinterpret_cast<const float4*>(cf.thr + gcol + g8 * 8 + 4);
              const float th[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float x = __uint_as_float(v[g8 * 8 + j]);
                if (x > th[j]) {
                  const int col = gcol + g8 * 8 + j;
                  const int slot = atomicAdd(cf.cnt + col, 1);
                  if (slot < cf.cap) cf.buf[(long long)col * cf.cap + slot] = make_uint2(__float_as_uint(x), uint32_t(grow));
                  else cf.thr[col] = CUDART_INF_F;  // full: the caller redoes this column; stop collecting for it
                }
              }
            }
            if (gm[g8] > thr) {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float x = __uint_as_float(v[g8 * 8 + j]);
                if (x > thr) {
                  queue[q_n * 128] = make_uint2(__float_as_uint(x), uint32_t(gcol + g8 * 8 + j));
                  ++q_n;
                }
              }
            }
            if (__any_sync(0xffffffffu, q_n > kQ - 8)) drain();
          }
        }
        if (last_in_item) drain();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (kCtaGroup == 2) mbar_arrive_cluster(&tmem_empty_bar[acc], 0);
          else mbar_arrive(&tmem_empty_bar[acc]);
        }
        if (last_in_item && grow < M) {
          const long long slot = ((long long)grow * n_chunks + chunk) * Cfg::EPI_GROUPS + wg;
#pragma unroll
          for (int p = 0; p < KC; ++p) {
            cand_val[slot * KC + p] = tv[p];
            cand_idx[slot * KC + p] = ti[p];
          }
          if (lse_part != nullptr) {
            lse_part[slot * 2] = run_max;
            lse_part[slot * 2 + 1] = run_sum;
          }
        }
        continue;
      }
#pragma unroll 1
      for (int c = wg; c < NUM_CHUNKS; c += Cfg::EPI_GROUPS) {
        if (ew == 0 && lane == 0) tma_store_wait_read<Cfg::CD_STAGES - 1>();  // staging buffer free again
        named_bar_sync(bar_id, 128);
        uint8_t* cd_row = smem_cd_wg + cd_stage * Cfg::CD_STAGE_BYTES + row_in_tile * 128;
#pragma unroll
        for (int s = 0; s < SUBS; ++s) {
          uint32_t v[32];
          const int col_in_tile = c * CHUNK_COLS + s * 32;
          tmem_ld_32x32(tmem_base + (uint32_t(ew * 32) << 16) + acc * Cfg::BLOCK_N + col_in_tile, v);
          tmem_ld_wait();
          const int gcol = n0 + col_in_tile;
          float f[32];
          if (fold_in) {  // rstd * (x.W'^T - mean * c) + b'   (LnFold)
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + gcol + j));
              const float4 c4 = __ldg(reinterpret_cast<const float4*>(lf.colsum + gcol + j));
              f[j + 0] = fmaf(ln_rstd, fmaf(-ln_mean, c4.x, __uint_as_float(v[j + 0])), b4.x);
              f[j + 1] = fmaf(ln_rstd, fmaf(-ln_mean, c4.y, __uint_as_float(v[j + 1])), b4.y);
              f[j + 2] = fmaf(ln_rstd, fmaf(-ln_mean, c4.z, __uint_as_float(v[j + 2])), b4.z);
              f[j + 3] = fmaf(ln_rstd, fmaf(-ln_mean, c4.w, __uint_as_float(v[j + 3])), b4.w);
            }
          } else {
            // split-K: the bias belongs to split 0, the later splits add their bare partial products
            const bool with_bias = !(kEpi == EPI_BIAS_ACCUM && chunk > 0);
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 b4 = with_bias ? __ldg(reinterpret_cast<const float4*>(bias + gcol + j)) : make_float4(0.f, 0.f, 0.f, 0.f);
              f[j + 0] = __uint_as_float(v[j + 0]) + b4.x;
              f[j + 1] = __uint_as_float(v[j + 1]) + b4.y;
              f[j + 2] = __uint_as_float(v[j + 2]) + b4.z;
              f[j + 3] = __uint_as_float(v[j + 3]) + b4.w;
            }
          }
          if constexpr (kEpi == EPI_BIAS_RESIDUAL_STATS) {
            // x_new = x + (acc + bias): the residual chunk was fetched one chunk ahead; fetch the next one now
            float4 rcur[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) rcur[q] = rnext[q];
            if (c + Cfg::EPI_GROUPS < NUM_CHUNKS && grow < M) {
              const float4* rp = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(residual) +
                                                                 (long long)grow * ldr + gcol + Cfg::EPI_GROUPS * CHUNK_COLS);
#pragma unroll
              for (int q = 0; q < 8; ++q) rnext[q] = rp[q];
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              f[4 * q + 0] += rcur[q].x; f[4 * q + 1] += rcur[q].y; f[4 * q + 2] += rcur[q].z; f[4 * q + 3] += rcur[q].w;
            }
            // running (mean, M2) of the row over this tile's columns: two-pass inside the chunk, Chan merge across chunks
            float cs = 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) cs += f[j];
            const float cm = cs * (1.0f / 32.0f);
            float cq = 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) { const float dv = f[j] - cm; cq = fmaf(dv, dv, cq); }
            const float n_new = st_n + 32.f;
            const float dlt = cm - st_mean;
            st_mean = fmaf(dlt, 32.f / n_new, st_mean);
            st_m2 += cq + dlt * dlt * (st_n * 32.f / n_new);
            st_n = n_new;
            if (grow < M) {  // bf16 copy of the new residual stream: the A operand of the next (LayerNorm-folded) GEMM
              uint4* hp = reinterpret_cast<uint4*>(lf.h_out + (long long)grow * lf.ldh + gcol);
#pragma unroll
              for (int q = 0; q < 4; ++q)
                hp[q] = make_uint4(pack_bf16x2(f[8 * q], f[8 * q + 1]), pack_bf16x2(f[8 * q + 2], f[8 * q + 3]),
                                   pack_bf16x2(f[8 * q + 4], f[8 * q + 5]), pack_bf16x2(f[8 * q + 6], f[8 * q + 7]));
              if (c + Cfg::EPI_GROUPS >= NUM_CHUNKS)  // this warpgroup's share of the tile: kLnPartCols columns
                reinterpret_cast<float2*>(lf.stats_out)[((long long)grow * num_n_tiles + n_blk) * Cfg::EPI_GROUPS + wg] =
                    make_float2(st_mean, st_m2);
            }
          }
          if constexpr (kEpi == EPI_BIAS_RELU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.0f);
          }
          if constexpr (kEpi == EPI_BIAS_SILU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = silu_fast(f[j]);
          }
          if constexpr (kEpi == EPI_BIAS_RESIDUAL) {
            if (grow < M) {
              const OutT* rp = residual + (long long)grow * ldr + gcol;
              if constexpr (sizeof(OutT) == 4) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                  const float4 r4 = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(rp) + j);
                  f[j + 0] += r4.x; f[j + 1] += r4.y; f[j + 2] += r4.z; f[j + 3] += r4.w;
                }
              } else {
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                  const uint4 r8 = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(rp) + j);
                  const uint32_t w[4] = {r8.x, r8.y, r8.z, r8.w};
#pragma unroll
                  for (int q = 0; q < 4; ++q) {
                    const __nv_bfloat162 p = *reinterpret_cast<const __nv_bfloat162*>(&w[q]);
                    f[j + 2 * q] += __low2float(p);
                    f[j + 2 * q + 1] += __high2float(p);
                  }
                }
              }
            }
          }
          // 128B-swizzled staging row: logical 16B chunk j lives at physical chunk j ^ (row % 8)
          if constexpr (sizeof(OutT) == 4) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const int phys = q ^ (row_in_tile & 7);
              *reinterpret_cast<float4*>(cd_row + phys * 16) =
                  make_float4(f[4 * q], f[4 * q + 1], f[4 * q + 2], f[4 * q + 3]);
            }
          } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int phys = (s * 4 + q) ^ (row_in_tile & 7);
              *reinterpret_cast<uint4*>(cd_row + phys * 16) =
                  make_uint4(pack_bf16x2(f[8 * q], f[8 * q + 1]), pack_bf16x2(f[8 * q + 2], f[8 * q + 3]),
                             pack_bf16x2(f[8 * q + 4], f[8 * q + 5]), pack_bf16x2(f[8 * q + 6], f[8 * q + 7]));
            }
          }
        }
        if (c + Cfg::EPI_GROUPS >= NUM_CHUNKS) {
          // this warp's TMEM reads of the accumulator are complete -> hand it back to the MMA issuer
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (kCtaGroup == 2) mbar_arrive_cluster(&tmem_empty_bar[acc], 0);
            else mbar_arrive(&tmem_empty_bar[acc]);
          }
        }
        fence_proxy_async_smem();
        named_bar_sync(bar_id, 128);
        if (ew == 0 && lane == 0) {
          if constexpr (kEpi == EPI_BIAS_ACCUM) {
            // Ordered split-K: the splits of a tile add into C one after the other (x + p0, + p1, + p2 -- the same sum on
            // every run).  Split s waits for the counter its predecessor leaves after ITS adds have completed; the
            // predecessor is a lower-numbered item, so it is already running on another resident cluster or finished.
            int* flag = nullptr;
            if (k_splits > 1) {
              flag = splitk_flags + ((long long)(m_blk * num_n_tiles + n_blk) * kCtaGroup + int(cta_rank)) * Cfg::EPI_GROUPS + wg;
              if (c == wg && chunk > 0) {
                while (ld_acquire_gpu(flag) != chunk) __nanosleep(64);
                fence_proxy_async_all();
              }
            }
            tma_reduce_add_2d(&tm_c, smem_cd_wg + cd_stage * Cfg::CD_STAGE_BYTES, n0 + c * CHUNK_COLS, m0);
            tma_store_commit();
            if (k_splits > 1 && c + Cfg::EPI_GROUPS >= NUM_CHUNKS) {
              tma_store_wait_all<0>();  // this split's adds have been performed
              fence_proxy_async_all();
              __threadfence();
              st_release_gpu(flag, chunk == k_splits - 1 ? 0 : chunk + 1);  // the last split re-arms the counter
            }
          } else {
            tma_store_2d(&tm_c, smem_cd_wg + cd_stage * Cfg::CD_STAGE_BYTES, n0 + c * CHUNK_COLS, m0);
            tma_store_commit();
          }
        }
        cd_stage ^= 1;
      }
    }
    if (ew == 0 && lane == 0) tma_store_wait_all<0>();
  }

  // ===================== teardown =====================
  tc_fence_before();
  if (kCtaGroup == 2) cluster_sync_all(); else __syncthreads();
  if (warp_idx == 2) tmem_dealloc<kCtaGroup>(tmem_base, Cfg::TMEM_COLS);
}

// ----------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------
typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                        CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                        CUtensorMapFloatOOBfill);

static PFN_tmapEncodeTiled get_encode_fn() {
  static PFN_tmapEncodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess || !p) {
      set_last_error("cuTensorMapEncodeTiled driver entry point not available");
      return nullptr;
    }
    fn = reinterpret_cast<PFN_tmapEncodeTiled>(p);
  }
  return fn;
}

// 2-D row-major tensor [rows, cols] with leading dimension ld (elements); box = [box_rows, box_cols];
// inner box extent must be exactly 128 bytes (SWIZZLE_128B).
int make_tmap_2d(CUtensorMap* out, const void* ptr, int elem_bytes, long long rows, long long cols, long long ld,
                 int box_rows, int box_cols) {
  PFN_tmapEncodeTiled enc = get_encode_fn();
  if (!enc) return -3;
  CUtensorMapDataType dt = (elem_bytes == 2) ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * (cuuint64_t)elem_bytes};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  if (box_cols * elem_bytes != 128) {
    set_last_error("make_tmap_2d: inner box must span 128 bytes");
    return -1;
  }
  CUresult r = enc(out, dt, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%lld cols=%lld ld=%lld elem=%d)", (int)r,
                   rows, cols, ld, elem_bytes);
    return -3;
  }
  return 0;
}

template <int kCtaGroup, int kEpi, typename OutT, int kEpiGroups = 2>
static int launch_inst(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const float* bias,
                       const void* residual, long long ldr, int M, int N, int K, int num_sms, cudaStream_t stream,
                       float* cand_val = nullptr, int* cand_idx = nullptr, float* lse_part = nullptr,
                       int n_chunks = 1, const LnFold& lf = LnFold(), const ColFilter& cf = ColFilter(), int k_splits = 1,
                       int* splitk_flags = nullptr) {
  using Cfg = GemmCfg<kCtaGroup, kEpiGroups>;
  auto kern = gemm_bf16_tcgen05_kernel<kCtaGroup, kEpi, OutT, kEpiGroups>;
  static bool attr_set[64] = {};
  if (first_use_on_device(attr_set)) {
    SB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
  }
  const int tile_m = Cfg::BLOCK_M * kCtaGroup;
  const long long num_m_tiles = (M + tile_m - 1) / tile_m;
  const long long num_tiles = num_m_tiles * ((N + Cfg::BLOCK_N - 1) / Cfg::BLOCK_N);
  long long clusters = num_sms / kCtaGroup;
  if (clusters > num_tiles * k_splits) clusters = num_tiles * k_splits;
  if (kEpi == EPI_TOPK && clusters > num_m_tiles * n_chunks) clusters = num_m_tiles * n_chunks;  // whole items
  if (clusters < 1) clusters = 1;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(clusters * kCtaGroup), 1, 1);
  cfg.blockDim = dim3(Cfg::THREADS, 1, 1);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCtaGroup;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  SB_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, ta, tb, tc, bias, reinterpret_cast<const OutT*>(residual), ldr, M, N, K,
                                   cand_val, cand_idx, lse_part, n_chunks, lf, cf, k_splits, splitk_flags));
  return 0;
}

// Number of n-chunks the top-k sweep is split into so that (m-blocks x chunks) fills the clusters.
int gemm_topk_chunks(int M, int N, int cta_group, int num_sms) {
  const int cg = (cta_group == 1) ? 1 : 2;
  const int clusters = (num_sms > 0 ? num_sms : 148) / cg;
  const int num_m_tiles = (M + 128 * cg - 1) / (128 * cg);
  const int num_n_tiles = (N + 255) / 256;
  int want = clusters / num_m_tiles;
  if (want < 1) want = 1;
  if (want > num_n_tiles) want = num_n_tiles;
  const int tpc = (num_n_tiles + want - 1) / want;
  return (num_n_tiles + tpc - 1) / tpc;  // every chunk non-empty
}

// Per-row top-kTopkCandidates of A[M,K] . W[N,K]^T (bf16 operands, fp32 accumulate) without materialising
// the product.  Outputs are per (row, list) with kTopkLists(n_chunks) = 2 * n_chunks lists per row (one per n-chunk and
// epilogue warpgroup; a list covers a disjoint subset of the columns): cand_val / cand_idx [M, lists, kTopkCandidates]
// sorted by value descending, and (optional) lse_part [M, lists, 2] = (max, sum exp(v - max)) over the list's columns.
int gemm_bf16_topk(const __nv_bfloat16* A, long long lda, const __nv_bfloat16* W, long long ldw, int M, int N, int K,
                   float* cand_val, int* cand_idx, float* lse_part, int n_chunks, int cta_group, int num_sms,
                   cudaStream_t stream, const ColFilter& cf) {
  if (M <= 0 || N <= 0) return 0;
  if (cf.thr != nullptr && (!cf.thr8 || !cf.cnt || !cf.buf || cf.cap <= 0)) {
    set_last_error("gemm_bf16_topk: column filter needs cnt, buf and a positive capacity");
    return -1;
  }
  if (K % 64 != 0 || K <= 0) {
    set_last_error("gemm_bf16_topk: K must be a positive multiple of 64 (got %d)", K);
    return -1;
  }
  const int num_n_tiles = (N + 255) / 256;
  if (n_chunks < 1 || n_chunks > num_n_tiles ||
      ((num_n_tiles + n_chunks - 1) / n_chunks) * (n_chunks - 1) >= num_n_tiles) {
    set_last_error("gemm_bf16_topk: invalid n_chunks=%d for %d n-tiles", n_chunks, num_n_tiles);
    return -1;
  }
  const int cg = (cta_group == 1) ? 1 : 2;
  CUtensorMap ta, tb;
  int rc;
  if ((rc = make_tmap_2d(&ta, A, 2, M, K, lda, 128, 64))) return rc;
  if ((rc = make_tmap_2d(&tb, W, 2, N, K, ldw, 256 / cg, 64))) return rc;
  const int sms = num_sms > 0 ? num_sms : 148;
  if (cg == 2)
    return launch_inst<2, EPI_TOPK, float>(ta, tb, ta /*unused*/, nullptr, nullptr, 0, M, N, K, sms, stream, cand_val,
                                           cand_idx, lse_part, n_chunks, LnFold(), cf);
  return launch_inst<1, EPI_TOPK, float>(ta, tb, ta /*unused*/, nullptr, nullptr, 0, M, N, K, sms, stream, cand_val,
                                         cand_idx, lse_part, n_chunks, LnFold(), cf);
}

int gemm_bf16(const GemmArgs& g, cudaStream_t stream) {
  if (g.M <= 0) return 0;
  if (g.allow_skinny && gemm_skinny_eligible(g) && !g.lf.stats_in && g.epi != EPI_BIAS_RESIDUAL_STATS)
    return gemm_skinny(g, stream);
  if (g.N % 256 != 0 || g.K % 64 != 0 || g.K <= 0 || g.N <= 0) {
    set_last_error("gemm_bf16: need N %% 256 == 0 and K %% 64 == 0 (got M=%d N=%d K=%d)", g.M, g.N, g.K);
    return -1;
  }
  if (g.epi == EPI_BIAS_RESIDUAL && !g.residual) {
    set_last_error("gemm_bf16: residual epilogue without residual pointer");
    return -1;
  }
  if (!g.bias) {
    set_last_error("gemm_bf16: bias pointer is required");
    return -1;
  }
  const int cg = (g.cta_group == 1) ? 1 : 2;
  const int out_bytes = g.out_fp32 ? 4 : 2;
  CUtensorMap ta, tb, tc;
  int rc;
  if ((rc = make_tmap_2d(&ta, g.A, 2, g.M, g.K, g.lda, 128, 64))) return rc;
  if ((rc = make_tmap_2d(&tb, g.W, 2, g.N, g.K, g.ldw, 256 / cg, 64))) return rc;
  if ((rc = make_tmap_2d(&tc, g.C, out_bytes, g.M, g.N, g.ldc, 128, 128 / out_bytes))) return rc;
  const int sms = g.num_sms > 0 ? g.num_sms : 148;
  int epi = g.epi;
  if (epi == EPI_BIAS_RESIDUAL && g.out_fp32 && g.residual == g.C && g.ldr == g.ldc) epi = EPI_BIAS_ACCUM;
  if (epi == EPI_BIAS_ACCUM && !g.out_fp32) {
    set_last_error("gemm_bf16: accumulate epilogue needs fp32 output");
    return -1;
  }

  // ---- LayerNorm folding (LnFold) ----
  if (g.lf.stats_in != nullptr) {  // consumer
    if ((epi != EPI_BIAS && epi != EPI_BIAS_RELU) || !g.lf.colsum || g.lf.chunks < 1 || g.lf.chunks > 8 ||
        g.K != kLnPartCols * g.lf.chunks) {
      set_last_error("gemm_bf16: folded LayerNorm input needs a bias / bias+ReLU epilogue and K = 128 * chunks <= 1024 "
                     "(K=%d chunks=%d)", g.K, g.lf.chunks);
      return -1;
    }
  }
  if (epi == EPI_BIAS_RESIDUAL_STATS) {  // producer
    if (!g.out_fp32 || !g.residual || !g.lf.h_out || !g.lf.stats_out || g.lf.ldh < g.N) {
      set_last_error("gemm_bf16: the residual+statistics epilogue needs fp32 C, a residual, h_out and stats_out");
      return -1;
    }
  } else if (g.lf.h_out != nullptr || g.lf.stats_out != nullptr) {
    set_last_error("gemm_bf16: h_out / stats_out are outputs of EPI_BIAS_RESIDUAL_STATS only");
    return -1;
  }

  // A/B variant: ONE epilogue warpgroup + 6 mainloop stages (the round-1 kernel) for the text encoder's four GEMMs
  if (g.epi_groups == 1) {
    if (cg != 2 || (epi != EPI_BIAS && epi != EPI_BIAS_RELU && epi != EPI_BIAS_ACCUM) || g.lf.stats_in != nullptr) {
      set_last_error("gemm_bf16: epi_groups = 1 is only built for the paired-CTA bias / bias+ReLU / accumulate epilogues");
      return -1;
    }
    if (epi == EPI_BIAS_ACCUM)
      return launch_inst<2, EPI_BIAS_ACCUM, float, 1>(ta, tb, tc, g.bias, g.residual, g.ldr, g.M, g.N, g.K, sms, stream);
    if (g.out_fp32) { set_last_error("gemm_bf16: epi_groups = 1 needs bf16 output for bias / bias+ReLU"); return -1; }
    if (epi == EPI_BIAS)
      return launch_inst<2, EPI_BIAS, __nv_bfloat16, 1>(ta, tb, tc, g.bias, g.residual, g.ldr, g.M, g.N, g.K, sms, stream);
    return launch_inst<2, EPI_BIAS_RELU, __nv_bfloat16, 1>(ta, tb, tc, g.bias, g.residual, g.ldr, g.M, g.N, g.K, sms, stream);
  }

  // Ordered split-K for the accumulate epilogue when the tiles do not fill the machine (the decoder's FFN output
  // projection at 2 560 rows: 40 tile pairs on 74 SM pairs): pick the split count with the fewest k-blocks on the
  // critical path, ~8 k-blocks charged per item for its epilogue and hand-over.
  if (epi == EPI_BIAS_ACCUM && cg == 2 && g.splitk_flags != nullptr) {
    const long long tiles = (long long)((g.M + 255) / 256) * (g.N / 256);
    const long long clusters = sms / 2;
    const int num_kb = g.K / 64;
    int best = 1;
    double best_cost = 1e30;
    for (int ks = 1; ks <= 4; ++ks) {
      if (num_kb / ks < 8) break;
      const double cost = double((tiles * ks + clusters - 1) / clusters) * (double(num_kb) / ks + 8.0);
      if (cost < best_cost * 0.95) { best_cost = cost; best = ks; }  // a later candidate must win by 5 %
    }
    if (best > 1 && tiles * 2 * 2 <= g.splitk_flags_len)
      return launch_inst<2, EPI_BIAS_ACCUM, float>(ta, tb, tc, g.bias, g.residual, g.ldr, g.M, g.N, g.K, sms, stream, nullptr,
                                                   nullptr, nullptr, 1, LnFold(), ColFilter(), best, g.splitk_flags);
  }

#define SB_DISPATCH(CG, EPI, T)                                                                                   \
  return launch_inst<CG, EPI, T>(ta, tb, tc, g.bias, g.residual, g.ldr, g.M, g.N, g.K, sms, stream, nullptr, nullptr, \
                                 nullptr, 1, g.lf)
#define SB_DISPATCH_EPI(CG, T)                                              \
  switch (epi) {                                                            \
    case EPI_BIAS: SB_DISPATCH(CG, EPI_BIAS, T);                            \
    case EPI_BIAS_RELU: SB_DISPATCH(CG, EPI_BIAS_RELU, T);                  \
    case EPI_BIAS_SILU: SB_DISPATCH(CG, EPI_BIAS_SILU, T);                  \
    case EPI_BIAS_RESIDUAL: SB_DISPATCH(CG, EPI_BIAS_RESIDUAL, T);          \
    case EPI_BIAS_ACCUM: SB_DISPATCH(CG, EPI_BIAS_ACCUM, float);            \
    case EPI_BIAS_RESIDUAL_STATS: SB_DISPATCH(CG, EPI_BIAS_RESIDUAL_STATS, float); \
    default: set_last_error("gemm_bf16: bad epilogue %d", epi); return -1;  \
  }
  if (cg == 2) {
    if (g.out_fp32) { SB_DISPATCH_EPI(2, float) } else { SB_DISPATCH_EPI(2, __nv_bfloat16) }
  } else {
    if (g.out_fp32) { SB_DISPATCH_EPI(1, float) } else { SB_DISPATCH_EPI(1, __nv_bfloat16) }
  }
#undef SB_DISPATCH
#undef SB_DISPATCH_EPI
  return -1;
}

}  // namespace sb
