// Transformer-XL relative-position self-attention of the w2v-BERT Conformer blocks on tcgen05 (BASELINE.json config 3,
// SURVEY §8 row a11 / App. B.2; reference wiring sonar/models/sonar_speech/factory.py:64-71, parameter names
// sdpa.u_bias / sdpa.v_bias / sdpa.r_proj in sonar_speech/handler.py:81-83):
//
//     score(i, j) = ((q_i + u) . k_j + (q_i + v) . p[c - 1 - i + j]) / 8,      p = r_proj(R) [2S-1 -> Npad rows],  c = S_center
//
// over PACKED utterances (row = cu[b] + t), keys >= len get probability exactly 0.
//
// Work decomposition.  An ITEM is (utterance b, 128-query tile, head h); a UNIT is one 128-key tile of it:
//     S1[128 x 128] = Qu . K^T                      Qu = bf16(q + u), Qv = bf16(q + v): prepared once per layer by
//     B [128 x 256] = Qv . Pw^T                     relpos_qprep_kernel, so no bias vector is added inside this kernel
//       Pw = the 256 rows of p the tile can reach: row n0 + cB, n0 = c - 1 - (q0 + 127) + j0, and the score of (query il,
//       key jl) of the tile uses column cB = 127 - il + jl -- the Transformer-XL "shift": every query row reads B at its
//       own offset.  TMEM loads are warp-uniform in the column address, so the 32-column window a warp fetches covers its
//       32 rows' needs (63 columns) and each thread then shifts its row by s = 31 - lane with a 5-stage barrel shifter of
//       register selects (16, 8, 4, 2, 1) -- no shared-memory round trip, no bank conflicts.
//     P = exp2((S1 + shift(B) - m_ref) / 8 * log2 e) one thread per query row; m_ref is a lazily raised reference (see the
//                                                   softmax section), so the scores are read once and exponentiated at once
//     O[128 x 64]  = P . V                          P stays in tensor memory (tcgen05.st + A-from-TMEM MMA)
// Tensor memory holds 256 columns per softmax group: S1 (128) | B half (128).  B is produced in two halves (columns
// 0-127, then 128-255 into the same TMEM columns): a chunk of 32 keys takes its position term from the low half when
// jl <= il and from the high half otherwise, so pass A completes the score chunks c < warp, pass B the chunks c >= warp (the
// diagonal chunk keeps its 32 low-half columns in registers in between); the probabilities overwrite the chunk's own score
// columns and O later reuses the B columns.  Two softmax warpgroups per CTA (one persistent CTA per SM) work on different items, so one group's
// exponentials overlap the other's MMAs; each group has its own operand slots and its own TMA producer thread.
//
// Warps: 0 = producer of group 0, 3 = producer of group 1, 1 = MMA issuer (event driven over both groups), 2 = TMEM allocator,
// 4-7 / 8-11 = softmax groups 0 / 1 (TMEM lane quarter = warp % 4).

#include "common.cuh"
#include "sonar_b200_internal.h"

#include <math_constants.h>

namespace sb {
namespace {

constexpr int kTile = 128 * 64 * 2;              // one [128 x 64] bf16 operand tile = 16 KB
constexpr int kGroupBytes = 6 * kTile;            // per softmax group: Qu | Qv | K | P_lo | P_hi | V  (96 KB)
constexpr int kOffQu = 0, kOffQv = kTile, kOffK = 2 * kTile, kOffPlo = 3 * kTile, kOffPhi = 4 * kTile, kOffV = 5 * kTile;
constexpr int kBarBytes = 512;
constexpr int kMaxBatch = 2047;                   // cu_seqlens and the query-tile prefix are staged in shared memory
constexpr int kSmemBytes = 2 * kGroupBytes + kBarBytes + 2 * (kMaxBatch + 1) * 4 + 1024;
constexpr int kThreads = 384;
static_assert(kSmemBytes <= 232448, "shared memory budget");

// barriers of one softmax group
enum { Q_FULL = 0, Q_EMPTY, KP_FULL, KP_EMPTY, V_FULL, V_EMPTY, AB_FULL, A_DONE, BHI_FULL, P_READY, O_FULL, O_FREE, kNumBars };

__device__ __forceinline__ uint64_t desc_mnmajor_sw128(uint32_t smem_addr) {  // V as the MN-major B operand (see attention_tc.cu)
  uint64_t lo = ((smem_addr >> 4) & 0x3FFFu) | (uint64_t((128u * 128u) >> 4) << 16);
  uint64_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
  return lo | (hi << 32);
}

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
      "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),
      "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}

// The unit sequence of one softmax group: items first + i * stride; an item = (query tile, head) of one utterance, expanded
// into its key tiles.  tile_cu[b] = number of query tiles before utterance b.
struct RelStream {
  const int32_t* cu;
  const int32_t* tile_cu;
  int B, H, num_items, item, stride;
  int h, tok0, len, q0, nt, kt;
  bool valid;
  __device__ __forceinline__ void load_item() {
    valid = item < num_items;
    if (!valid) return;
    const int tile = item / H;
    h = item - tile * H;
    int lo = 0, hi = B - 1;  // last b with tile_cu[b] <= tile
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (tile_cu[mid] <= tile) lo = mid; else hi = mid - 1;
    }
    tok0 = cu[lo];
    len = cu[lo + 1] - tok0;
    q0 = (tile - tile_cu[lo]) * 128;
    nt = (len + 127) >> 7;
    kt = 0;
  }
  __device__ __forceinline__ void init(const int32_t* cu_, const int32_t* tile_cu_, int B_, int H_, int first, int stride_) {
    cu = cu_; tile_cu = tile_cu_; B = B_; H = H_; num_items = tile_cu_[B_] * H_; item = first; stride = stride_;
    load_item();
  }
  __device__ __forceinline__ void advance() {
    if (++kt == nt) {
      item += stride;
      load_item();
    }
  }
};

// lo[k] = in[k + s], k = 0..31, s in [0, 31], where in = lo | hi (64 values): five select stages (16, 8, 4, 2, 1), in place.
// (Two separate 32-register arrays with compile-time indices only: nothing here may end up in local memory.)
template <int SH>
__device__ __forceinline__ void barrel_stage(uint32_t (&lo)[32], uint32_t (&hi)[32], bool on) {
  // after this stage positions [0, 64 - sum of shifts so far) are meaningful; ascending k reads only not-yet-written slots
#pragma unroll
  for (int k = 0; k < 32; ++k) {
    const uint32_t src = (k + SH < 32) ? lo[(k + SH) & 31] : hi[(k + SH - 32) & 31];
    lo[k] = on ? src : lo[k];
  }
#pragma unroll
  for (int k = 0; k + SH < 32; ++k) hi[k] = on ? hi[k + SH] : hi[k];
}
__device__ __forceinline__ void barrel_shift(uint32_t (&lo)[32], uint32_t (&hi)[32], int s) {
  barrel_stage<16>(lo, hi, s & 16);
  barrel_stage<8>(lo, hi, s & 8);
  barrel_stage<4>(lo, hi, s & 4);
  barrel_stage<2>(lo, hi, s & 2);
  barrel_stage<1>(lo, hi, s & 1);
}

__global__ void __launch_bounds__(kThreads, 1)
attention_relpos_tc_kernel(const __grid_constant__ CUtensorMap tm_qu, const __grid_constant__ CUtensorMap tm_qv,
                           const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_p,
                           const int32_t* cu_g, int B, int H, int S_center, __nv_bfloat16* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * kGroupBytes);  // [2][kNumBars]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * kNumBars);
  int32_t* cu = reinterpret_cast<int32_t*>(smem + 2 * kGroupBytes + kBarBytes);
  int32_t* tile_cu = cu + (kMaxBatch + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int D = H * 64;
  for (int i = threadIdx.x; i <= B; i += kThreads) cu[i] = cu_g[i];
  if (warp == 1 && lane == 0) {
    tma_prefetch_desc(&tm_qu);
    tma_prefetch_desc(&tm_qv);
    tma_prefetch_desc(&tm_qkv);
    tma_prefetch_desc(&tm_p);
    for (int g = 0; g < 2; ++g) {
      uint64_t* b = bars + g * kNumBars;
      mbar_init(&b[Q_FULL], 1);
      mbar_init(&b[Q_EMPTY], 1);
      mbar_init(&b[KP_FULL], 1);
      mbar_init(&b[KP_EMPTY], 1);
      mbar_init(&b[V_FULL], 1);
      mbar_init(&b[V_EMPTY], 1);
      mbar_init(&b[AB_FULL], 1);
      mbar_init(&b[A_DONE], 4);
      mbar_init(&b[BHI_FULL], 1);
      mbar_init(&b[P_READY], 4);
      mbar_init(&b[O_FULL], 1);
      mbar_init(&b[O_FREE], 4);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<1>(tmem_ptr_smem, 512);
  __syncthreads();
  if (threadIdx.x == 0) {  // query-tile prefix over the utterances (B <= kMaxBatch)
    int acc = 0;
    for (int b = 0; b < B; ++b) {
      tile_cu[b] = acc;
      acc += (cu[b + 1] - cu[b] + 127) >> 7;
    }
    tile_cu[B] = acc;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  if ((warp == 0 || warp == 3) && lane == 0) {
    // ============================ TMA producer of group g ============================
    const int g = (warp == 0) ? 0 : 1;
    uint64_t* bar = bars + g * kNumBars;
    uint8_t* base = smem + g * kGroupBytes;
    RelStream s;
    s.init(cu, tile_cu, B, H, blockIdx.x + g * gridDim.x, 2 * gridDim.x);
    uint32_t n = 0, ni = 0;  // units / items loaded so far
    while (s.valid) {
      const int col = s.h * 64;
      if (s.kt == 0) {  // the item's biased queries
        mbar_wait(&bar[Q_EMPTY], (ni & 1) ^ 1);
        mbar_arrive_expect_tx(&bar[Q_FULL], 2 * kTile);
        tma_load_2d(base + kOffQu, &tm_qu, &bar[Q_FULL], col, s.tok0 + s.q0);
        tma_load_2d(base + kOffQv, &tm_qv, &bar[Q_FULL], col, s.tok0 + s.q0);
        ++ni;
      }
      const int j0 = s.kt * 128;
      const int n0 = S_center - 1 - (s.q0 + 127) + j0;  // p row of window column 0 (may be negative: zero filled)
      mbar_wait(&bar[KP_EMPTY], (n & 1) ^ 1);
      mbar_arrive_expect_tx(&bar[KP_FULL], 3 * kTile);
      tma_load_2d(base + kOffK, &tm_qkv, &bar[KP_FULL], D + col, s.tok0 + j0);
      tma_load_2d(base + kOffPlo, &tm_p, &bar[KP_FULL], col, n0);
      tma_load_2d(base + kOffPhi, &tm_p, &bar[KP_FULL], col, n0 + 128);
      mbar_wait(&bar[V_EMPTY], (n & 1) ^ 1);
      mbar_arrive_expect_tx(&bar[V_FULL], kTile);
      tma_load_2d(base + kOffV, &tm_qkv, &bar[V_FULL], 2 * D + col, s.tok0 + j0);
      ++n;
      s.advance();
    }
  } else if (warp == 1 && lane == 0) {
    // ============================ MMA issuer: event driven over both groups ============================
    constexpr uint32_t idesc_s = umma_idesc_bf16_f32(128, 128);
    constexpr uint32_t idesc_o = umma_idesc_bf16_f32(128, 64) | (1u << 16);  // B operand MN-major
    RelStream s0, s1;
    s0.init(cu, tile_cu, B, H, blockIdx.x, 2 * gridDim.x);
    s1.init(cu, tile_cu, B, H, blockIdx.x + gridDim.x, 2 * gridDim.x);
    int ph0 = s0.valid ? 0 : 3, ph1 = s1.valid ? 0 : 3;  // next phase of each group (3 = stream exhausted)
    uint32_t n0u = 0, n1u = 0, ni0 = 0, ni1 = 0;          // units / items issued per group
    long long t_idle = clock64();
    while (ph0 != 3 || ph1 != 3) {
      bool progressed = false;
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        const int ph = g ? ph1 : ph0;
        if (ph == 3) continue;
        const uint32_t n = g ? n1u : n0u;
        const uint32_t ni = g ? ni1 : ni0;
        const int kt = g ? s1.kt : s0.kt;
        const int nt = g ? s1.nt : s0.nt;
        const int len = g ? s1.len : s0.len;
        uint64_t* bar = bars + g * kNumBars;
        const uint32_t sbase = smem_u32(smem + g * kGroupBytes);
        const uint32_t tS = tmem_base + g * 256, tB = tS + 128;
        if (ph == 0) {  // S1 = Qu K^T, B_lo = Qv P_lo^T
          if (kt == 0 && !mbar_test_wait(&bar[Q_FULL], ni & 1)) continue;
          if (!mbar_test_wait(&bar[KP_FULL], n & 1)) continue;
          if (n > 0 && !mbar_test_wait(&bar[O_FREE], (n - 1) & 1)) continue;
          tc_fence_after();
          const uint64_t qu = umma_desc_kmajor_sw128(sbase + kOffQu), qv = umma_desc_kmajor_sw128(sbase + kOffQv);
          const uint64_t kd = umma_desc_kmajor_sw128(sbase + kOffK), pl = umma_desc_kmajor_sw128(sbase + kOffPlo);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16<1>(tS, qu + uint64_t(2 * k), kd + uint64_t(2 * k), idesc_s, k != 0);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16<1>(tB, qv + uint64_t(2 * k), pl + uint64_t(2 * k), idesc_s, k != 0);
          umma_commit<1>(&bar[AB_FULL]);
          if (g) ph1 = 1; else ph0 = 1;
          progressed = true;
        } else if (ph == 1) {  // B_hi = Qv P_hi^T over the same TMEM columns, once pass A has consumed B_lo
          if (!mbar_test_wait(&bar[A_DONE], n & 1)) continue;
          tc_fence_after();
          const uint64_t qv = umma_desc_kmajor_sw128(sbase + kOffQv), phd = umma_desc_kmajor_sw128(sbase + kOffPhi);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16<1>(tB, qv + uint64_t(2 * k), phd + uint64_t(2 * k), idesc_s, k != 0);
          umma_commit<1>(&bar[BHI_FULL]);
          umma_commit<1>(&bar[KP_EMPTY]);                    // K, P_lo, P_hi are dead
          if (kt == nt - 1) umma_commit<1>(&bar[Q_EMPTY]);   // last key tile of the item: Qu, Qv are dead too
          if (g) ph1 = 2; else ph0 = 2;
          progressed = true;
        } else {  // O = P V  (P over the score columns, O into B+0..63)
          if (!mbar_test_wait(&bar[P_READY], n & 1)) continue;
          if (!mbar_test_wait(&bar[V_FULL], n & 1)) continue;
          tc_fence_after();
          const int kv_valid = min(128, len - kt * 128);
          const int ksteps = (kv_valid + 15) >> 4;
          for (int k = 0; k < ksteps; ++k)  // 16 keys per k-step: P of key chunk c lives at S columns 32c .. 32c+15
            umma_bf16_ts(tB, tS + 32 * (k >> 1) + 8 * (k & 1), desc_mnmajor_sw128(sbase + kOffV + k * 2048), idesc_o, k != 0);
          umma_commit<1>(&bar[O_FULL]);
          umma_commit<1>(&bar[V_EMPTY]);
          if (g) {
            ++n1u;
            if (kt == nt - 1) ++ni1;
            s1.advance();
            ph1 = s1.valid ? 0 : 3;
          } else {
            ++n0u;
            if (kt == nt - 1) ++ni0;
            s0.advance();
            ph0 = s0.valid ? 0 : 3;
          }
          progressed = true;
        }
      }
      if (progressed) {
        t_idle = clock64();
      } else if (clock64() - t_idle > SB_MBAR_TIMEOUT_CYCLES) {
        printf("sonar_b200: rel-pos attention MMA issuer stuck block=%d phases=(%d,%d)\n", blockIdx.x, ph0, ph1);
        __trap();
      }
    }
  } else if (warp >= 4) {
    // ============================ softmax + epilogue: one thread per query row ============================
    const int g = (warp - 4) >> 2;
    const int wq = warp & 3;  // TMEM lane quarter = 32-row block of the query tile
    const int row = wq * 32 + lane;
    const int shift = 31 - lane;
    const uint32_t lane_base = uint32_t(wq * 32) << 16;
    const uint32_t tS = tmem_base + g * 256 + lane_base, tB = tS + 128;
    uint64_t* bar = bars + g * kNumBars;
    const float sl2 = 0.125f * 1.4426950408889634f;  // 1/sqrt(64) * log2(e)
    RelStream u;
    u.init(cu, tile_cu, B, H, blockIdx.x + g * gridDim.x, 2 * gridDim.x);
    uint32_t n = 0;
    // Softmax state of the query row across the key tiles of an item.  The exponentials are taken against a REFERENCE m_ref
    // that is only raised when a score exceeds it by more than kTau (then everything accumulated so far is rescaled): the
    // probabilities stay below 2^8 relative to the reference, well inside bf16 / fp32 range, and the common case needs no
    // second pass over the scores and no per-tile rescaling of the accumulator.
    constexpr float kTau = 8.0f / (0.125f * 1.4426950408889634f);
    float m_ref = -CUDART_INF_F, l_run = 0.f;
    float o_acc[64];
    while (u.valid) {
      const int kv_valid = min(128, u.len - u.kt * 128);
      const int nch = (kv_valid + 31) >> 5;  // 32-key chunks holding valid keys
      if (u.kt == 0) {
        m_ref = -CUDART_INF_F;
        l_run = 0.f;
#pragma unroll
        for (int j = 0; j < 64; ++j) o_acc[j] = 0.f;
      }
      float sum = 0.f;       // this key tile's share of the row sum (relative to m_ref)
      uint32_t done = 0;     // chunks of this tile whose P is already in tensor memory (warp-uniform)

      // sc = the complete scores of chunk c -> probabilities (bf16 pairs) over the chunk's own S columns
      auto finalize = [&](int c, uint32_t (&sc)[32]) {
        const int lim = kv_valid - c * 32;
        if (lim < 32) {  // the chunk reaches past the utterance: those keys get probability exactly 0
#pragma unroll
          for (int k = 0; k < 32; ++k)
            if (k >= lim) sc[k] = __float_as_uint(-CUDART_INF_F);
        }
        float m0 = __uint_as_float(sc[0]), m1 = __uint_as_float(sc[1]);
#pragma unroll
        for (int k = 2; k < 32; k += 2) {
          m0 = fmaxf(m0, __uint_as_float(sc[k]));
          m1 = fmaxf(m1, __uint_as_float(sc[k + 1]));
        }
        const float cm = fmaxf(m0, m1);
        const bool bump = cm > m_ref + kTau;  // always on the item's first chunk (m_ref = -inf)
        if (__any_sync(0xffffffffu, bump)) {   // warp-uniform branch; lanes that keep their reference scale by exactly 1
          const float m_new = bump ? cm : m_ref;
          const float f = ex2f((m_ref - m_new) * sl2);
          l_run *= f;
          sum *= f;
#pragma unroll
          for (int j = 0; j < 64; ++j) o_acc[j] *= f;
          if (done) tmem_st_wait();  // (their tcgen05.st must have landed before they are read back)
          for (int cc = 0; cc < 4; ++cc) {  // probabilities of this tile already written against the old reference
            if (!(done & (1u << cc))) continue;
            uint32_t pk[16];
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                : "=r"(pk[0]), "=r"(pk[1]), "=r"(pk[2]), "=r"(pk[3]), "=r"(pk[4]), "=r"(pk[5]), "=r"(pk[6]), "=r"(pk[7]),
                  "=r"(pk[8]), "=r"(pk[9]), "=r"(pk[10]), "=r"(pk[11]), "=r"(pk[12]), "=r"(pk[13]), "=r"(pk[14]), "=r"(pk[15])
                : "r"(tS + cc * 32)
                : "memory");
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const __nv_bfloat162 h2 = *reinterpret_cast<const __nv_bfloat162*>(&pk[j]);
              pk[j] = pack_bf16x2(__low2float(h2) * f, __high2float(h2) * f);
            }
            tmem_st_32x16(tS + cc * 32, pk);
          }
          m_ref = m_new;
        }
        const float mxs = m_ref * sl2;
        uint32_t pk[16];
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int k = 0; k < 32; k += 2) {
          const float p0 = ex2f(fmaf(__uint_as_float(sc[k]), sl2, -mxs));
          const float p1 = ex2f(fmaf(__uint_as_float(sc[k + 1]), sl2, -mxs));
          s0 += p0;
          s1 += p1;
          pk[k >> 1] = pack_bf16x2(p0, p1);
        }
        sum += s0 + s1;
        tmem_st_32x16(tS + c * 32, pk);  // over the first 16 of the chunk's 32 score columns (its scores are in registers)
        done |= 1u << c;
      };
      // chunk c gets its position terms from window columns [colB, colB + 64) of the resident half of B
      auto fold = [&](int c, int colB) {
        uint32_t lo[32], hi[32], sc[32];
        tmem_ld_32x32(tB + colB, lo);
        tmem_ld_32x32(tB + colB + 32, hi);
        tmem_ld_32x32(tS + c * 32, sc);
        tmem_ld_wait();
        barrel_shift(lo, hi, shift);
#pragma unroll
        for (int k = 0; k < 32; ++k) sc[k] = __float_as_uint(__uint_as_float(sc[k]) + __uint_as_float(lo[k]));
        finalize(c, sc);
      };

      mbar_wait(&bar[AB_FULL], n & 1);
      tc_fence_after();
      // ---- pass A: low half of B (window columns 0..127): chunks c < wq complete; the diagonal chunk c == wq needs
      //      columns 96..127 of this half (kept in registers) and 128..158 of the high half ----
      for (int c = 0; c < nch && c < wq; ++c) fold(c, 96 - 32 * (wq - c));
      uint32_t dg[32];
      const bool has_diag = wq < nch;
      if (has_diag) {
        tmem_ld_32x32(tB + 96, dg);
        tmem_ld_wait();
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar[A_DONE]);
      // ---- pass B: high half of B, now in the same TMEM columns: the diagonal chunk and the chunks c > wq ----
      mbar_wait(&bar[BHI_FULL], n & 1);
      tc_fence_after();
      if (has_diag) {
        uint32_t hi[32], sc[32];
        tmem_ld_32x32(tB, hi);
        tmem_ld_32x32(tS + wq * 32, sc);
        tmem_ld_wait();
        barrel_shift(dg, hi, shift);
#pragma unroll
        for (int k = 0; k < 32; ++k) sc[k] = __float_as_uint(__uint_as_float(sc[k]) + __uint_as_float(dg[k]));
        finalize(wq, sc);
      }
      for (int c = wq + 1; c < nch; ++c) fold(c, 32 * (c - wq) - 32);
      l_run += sum;
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar[P_READY]);
      // ---- O tile of this key tile (relative to the same reference) -> register accumulator ----
      mbar_wait(&bar[O_FULL], n & 1);
      tc_fence_after();
      uint32_t o[2][32];
      tmem_ld_32x32(tB, o[0]);
      tmem_ld_32x32(tB + 32, o[1]);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar[O_FREE]);
#pragma unroll
      for (int j = 0; j < 64; ++j) o_acc[j] += __uint_as_float(o[j >> 5][j & 31]);
      if (u.kt == u.nt - 1 && u.q0 + row < u.len) {
        const float inv = 1.0f / l_run;
        uint4* dst = reinterpret_cast<uint4*>(out + (long long)(u.tok0 + u.q0 + row) * D + u.h * 64);
#pragma unroll
        for (int q = 0; q < 8; ++q)
          dst[q] = make_uint4(pack_bf16x2(o_acc[8 * q] * inv, o_acc[8 * q + 1] * inv),
                              pack_bf16x2(o_acc[8 * q + 2] * inv, o_acc[8 * q + 3] * inv),
                              pack_bf16x2(o_acc[8 * q + 4] * inv, o_acc[8 * q + 5] * inv),
                              pack_bf16x2(o_acc[8 * q + 6] * inv, o_acc[8 * q + 7] * inv));
      }
      ++n;
      u.advance();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<1>(tmem_base, 512);
}

// qu = bf16(q + u), qv = bf16(q + v): q = first D columns of the packed qkv rows; u, v fp32 [D]  (8 elements per thread)
__global__ void __launch_bounds__(256)
relpos_qprep_kernel(const __nv_bfloat16* __restrict__ qkv, const float* __restrict__ u_bias, const float* __restrict__ v_bias,
                    long long T, int D, __nv_bfloat16* __restrict__ qu, __nv_bfloat16* __restrict__ qv) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int per_row = D / 8;
  if (i >= T * per_row) return;
  const long long t = i / per_row;
  const int c = int(i - t * per_row) * 8;
  const uint4 q8 = *reinterpret_cast<const uint4*>(qkv + t * 3 * D + c);
  const uint32_t w[4] = {q8.x, q8.y, q8.z, q8.w};
  uint32_t a[4], b[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const __nv_bfloat162 p = *reinterpret_cast<const __nv_bfloat162*>(&w[e]);
    const float x0 = __low2float(p), x1 = __high2float(p);
    a[e] = pack_bf16x2(x0 + u_bias[c + 2 * e], x1 + u_bias[c + 2 * e + 1]);
    b[e] = pack_bf16x2(x0 + v_bias[c + 2 * e], x1 + v_bias[c + 2 * e + 1]);
  }
  *reinterpret_cast<uint4*>(qu + t * D + c) = make_uint4(a[0], a[1], a[2], a[3]);
  *reinterpret_cast<uint4*>(qv + t * D + c) = make_uint4(b[0], b[1], b[2], b[3]);
}

}  // namespace

// qkv [T, 3D] bf16 packed rows (q | k | v), p [Npad, D] bf16 = r_proj(relative-position table), u_bias / v_bias fp32 [D],
// qu / qv [T, D] bf16 scratch, out [T, D] bf16.  S_center = the batch's maximum length (row c-1-i+j of p <-> offset i-j).
int attention_relpos_tc(const __nv_bfloat16* qkv, const __nv_bfloat16* p, const float* u_bias, const float* v_bias,
                        const int32_t* cu_seqlens, int B, int H, long long total_tokens, int Npad, int S_center,
                        __nv_bfloat16* qu, __nv_bfloat16* qv, __nv_bfloat16* out, int num_sms, cudaStream_t stream) {
  if (B <= 0 || total_tokens <= 0) return 0;
  if (B > kMaxBatch) {
    set_last_error("attention_relpos_tc: at most %d utterances per batch (got %d)", kMaxBatch, B);
    return -1;
  }
  const int D = H * 64;
  const long long nthr = total_tokens * (D / 8);
  relpos_qprep_kernel<<<(unsigned)((nthr + 255) / 256), 256, 0, stream>>>(qkv, u_bias, v_bias, total_tokens, D, qu, qv);
  SB_CUDA_CHECK(cudaGetLastError());
  CUtensorMap tm_qu, tm_qv, tm_qkv, tm_p;
  int rc;
  if ((rc = make_tmap_2d(&tm_qu, qu, 2, total_tokens, D, D, 128, 64))) return rc;
  if ((rc = make_tmap_2d(&tm_qv, qv, 2, total_tokens, D, D, 128, 64))) return rc;
  if ((rc = make_tmap_2d(&tm_qkv, qkv, 2, total_tokens, 3ll * D, 3ll * D, 128, 64))) return rc;
  if ((rc = make_tmap_2d(&tm_p, p, 2, Npad, D, D, 128, 64))) return rc;
  static bool attr_set[64] = {};
  if (first_use_on_device(attr_set)) {
    SB_CUDA_CHECK(cudaFuncSetAttribute(attention_relpos_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
  }
  const int grid = num_sms > 0 ? num_sms : 148;
  attention_relpos_tc_kernel<<<(unsigned)grid, kThreads, kSmemBytes, stream>>>(tm_qu, tm_qv, tm_qkv, tm_p, cu_seqlens, B, H,
                                                                               S_center, out);
  SB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // namespace sb
