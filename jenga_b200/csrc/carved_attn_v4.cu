// Carved attention forward, generation 4 — two softmax streams over one q block.
//
// Generation 2 (carved_attn.cu) is tensor-pipe bound at 70 % of peak because its QK^T MMAs are
// N=64 (48 cycles instead of 32, tools/umma_probe.cu); generation 3 issues N=128 MMAs only but
// runs all softmax warps in lockstep on one tile and leaves the tensor pipe idle 40 % of the time.
// This generation keeps every MMA at N=128 AND keeps two tiles in flight in opposite phases:
//
//   one CTA per SM, 512 TMEM columns:  S0 [0,128)  S1 [128,256)  O0 [256,384)  O1 [384,512)
//   * the live key blocks of the q block are dealt alternately to two streams (even / odd);
//     each stream has its own score buffer S_s, its own accumulator O_s and its own softmax
//     warpgroup (one thread per query row, the whole 128-key row in registers: no shuffles, no
//     shared-memory exchange per tile) with its own running max m_s and row sum l_s
//     — "split-KV" inside the CTA.
//   * one tcgen05 issuer, fixed order  QK(0) QK(1) | PV(0) QK(2) | PV(1) QK(3) | ...
//     While warpgroup 0 works on S(j), the tensor pipe runs PV(j-1) and QK(j+1) of the other
//     stream; P(j) (16-bit, written over S(j)) is handed over in two 64-key halves so PV(j)
//     starts as soon as the first half is ready.  QK(j+2) is issued right behind PV(j): the
//     tcgen05 pipe executes in issue order, so it overwrites S_s only after PV(j) consumed P(j).
//   * S_FULL(j) is a tcgen05.commit by the same issuing thread, so when a warpgroup sees it,
//     PV(j-2) — the previous accumulation into its O_s — has retired: the (rare, lazy) rescale
//     of O_s needs no further wait.
//   * epilogue: O = (O0 2^(m0-m) + O1 2^(m1-m)) / (l0 2^(m0-m) + l1 2^(m1-m)), m = max(m0, m1);
//     warpgroup s writes output columns [64 s, 64 s + 64).
//   * 227 KB of shared memory: Q tile + 3-deep rings of 32 KB K and V tiles.
//
// O/l does not depend on the softmax reference point or on the order partial sums are combined,
// so this is the reference's online softmax (attention_block_triton_diffres.py:121-135) up to
// fp32 rounding.  Same C-ABI, same KernelParams as generation 2 (tests run every generation).
#include "carved_attn_common.cuh"

#ifndef JENGA_POLY_EVERY
#define JENGA_POLY_EVERY 3
#endif
#ifndef JENGA_V4_STAGES
#define JENGA_V4_STAGES 3
#endif

namespace jenga {

namespace {

using attn::BlockWalker;
using attn::KernelParams;

constexpr int kBlock = 128;
constexpr int kHeadDim = 128;
constexpr int kStages = JENGA_V4_STAGES;           // K ring and V ring depth
constexpr int kHalfTileBytes = kBlock * 64 * 2;    // 16 KB: 128 rows x 64 d
constexpr int kTileBytes = 2 * kHalfTileBytes;     // 32 KB
constexpr int kMaxMaskWords = 256;
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kColS = 0, kColO = 256;         // + 128 * stream
constexpr int kPolyEvery = JENGA_POLY_EVERY;
constexpr int kThreads = 64 + 256;                 // TMA warp, MMA warp, 2 x 4 softmax warps

constexpr int kOffQ = 0;                           // Q tile; reused for the final (m, l) exchange
constexpr int kOffK = kTileBytes;
constexpr int kOffV = kOffK + kStages * kTileBytes;
constexpr int kOffBars = kOffV + kStages * kTileBytes;
constexpr int kOffMask = kOffBars + 256;
constexpr int kSmemBytes = 1024 + kOffMask + kMaxMaskWords * 4;
static_assert(kSmemBytes <= 232448, "exceeds the 227 KB per-CTA shared memory limit");

enum BarId {
  Q_FULL = 0, Q_READY,
  K_FULL0, K_EMPTY0 = K_FULL0 + kStages,
  V_FULL0 = K_EMPTY0 + kStages, V_EMPTY0 = V_FULL0 + kStages,
  S_FULL0 = V_EMPTY0 + kStages, S_FULL1,
  P_FULL00, P_FULL01, P_FULL10, P_FULL11,   // [stream][key half]
  ALL_DONE,
  NUM_BARS
};
static_assert(NUM_BARS * 8 + 8 <= 256, "barrier block overflow");

template <bool kBF16>
__global__ void __launch_bounds__(kThreads, 1)
carved_attn_v4_kernel(const __grid_constant__ CUtensorMap tm_q,
                      const __grid_constant__ CUtensorMap tm_k,
                      const __grid_constant__ CUtensorMap tm_v, const KernelParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint8_t* sQ = smem + kOffQ;
  uint8_t* sK = smem + kOffK;
  uint8_t* sV = smem + kOffV;
  float* s_exch = reinterpret_cast<float*>(smem + kOffQ);  // [2 streams][m, l][128 rows]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBars);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NUM_BARS);
  uint32_t* s_mask = reinterpret_cast<uint32_t*>(smem + kOffMask);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // ---- which tile is this CTA?  head-major; within a head the (long) dense blocks first ----
  const int per_bh = p.nq_sparse + p.nq_dense;
  const int bh = blockIdx.x / per_bh;
  const int local = blockIdx.x - bh * per_bh;
  const bool dense = local < p.nq_dense;
  const int qb = dense ? p.nq_sparse + local : local - p.nq_dense;
  const int b = bh / p.heads;
  const int h = bh - b * p.heads;
  const long long q_row0 = static_cast<long long>(qb) * kBlock;
  const long long seqlen_over = p.seqlen_dev ? static_cast<long long>(__ldg(p.seqlen_dev)) : -1;
  const long long q_limit_sparse = seqlen_over >= 0 ? seqlen_over : p.q_limit_sparse;
  const long long kv_limit =
      dense ? p.kv_limit_dense : (seqlen_over >= 0 ? seqlen_over : p.kv_limit_sparse);
  const bool skip_all = (!dense && q_row0 >= q_limit_sparse);  // ref :59-61

  const int nwords = p.mask_words;
  for (int w = threadIdx.x; w < nwords; w += kThreads) {
    uint32_t bits;
    if (skip_all) {
      bits = 0;
    } else if (dense) {
      const int rem = p.nb_kv - w * 32;
      bits = rem >= 32 ? 0xffffffffu : (rem > 0 ? ((1u << rem) - 1u) : 0u);
    } else {
      bits = p.mask_bits[(static_cast<size_t>(bh) * p.nq_sparse + qb) * nwords + w];
      const int rem = p.nb_kv - w * 32;
      if (rem < 32) bits &= rem > 0 ? ((1u << rem) - 1u) : 0u;
    }
    s_mask[w] = bits;
  }

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
#pragma unroll
    for (int i = 0; i < NUM_BARS; ++i)
      mbar_init(&bars[i], i == Q_READY ? 256u : ((i >= P_FULL00 && i <= P_FULL11) ? 128u : 1u));
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  int n_tiles = 0;
  for (int w = 0; w < nwords; ++w) n_tiles += __popc(s_mask[w]);

  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (n_tiles > 0 && elect_one()) {
      mbar_arrive_expect_tx(&bars[Q_FULL], kTileBytes);
      tma_load_4d(sQ, &tm_q, &bars[Q_FULL], 0, static_cast<int>(q_row0), h, b);
      tma_load_4d(sQ + kHalfTileBytes, &tm_q, &bars[Q_FULL], 64, static_cast<int>(q_row0), h, b);
      BlockWalker it(s_mask, nwords);
      int slot = 0;
      uint32_t ph = 0;
      for (int blk = it.next(); blk >= 0; blk = it.next()) {
        const int row0 = blk * kBlock;
        uint8_t* kd = sK + slot * kTileBytes;
        uint8_t* vd = sV + slot * kTileBytes;
        mbar_wait(&bars[K_EMPTY0 + slot], ph ^ 1, p.err_flag);
        mbar_arrive_expect_tx(&bars[K_FULL0 + slot], kTileBytes);
        tma_load_4d(kd, &tm_k, &bars[K_FULL0 + slot], 0, row0, h, b);
        tma_load_4d(kd + kHalfTileBytes, &tm_k, &bars[K_FULL0 + slot], 64, row0, h, b);
        mbar_wait(&bars[V_EMPTY0 + slot], ph ^ 1, p.err_flag);
        mbar_arrive_expect_tx(&bars[V_FULL0 + slot], kTileBytes);
        tma_load_4d(vd, &tm_v, &bars[V_FULL0 + slot], 0, row0, h, b);
        tma_load_4d(vd + kHalfTileBytes, &tm_v, &bars[V_FULL0 + slot], 64, row0, h, b);
        if (++slot == kStages) {
          slot = 0;
          ph ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // =============================== tcgen05 issuer ===============================
    if (n_tiles > 0 && elect_one()) {
      constexpr uint32_t idesc_qk = umma_idesc_f16(kBF16, /*b_mn_major=*/false, 128, kBlock);
      constexpr uint32_t idesc_pv = umma_idesc_f16(kBF16, /*b_mn_major=*/true, 128, kHeadDim);
      // K-major SW128 operands (Q, K): 8-row groups 1024 B apart (SBO), d halves 16 KB apart.
      const uint64_t q_desc = umma_smem_desc(smem_u32(sQ), 16, 1024, UMMA_LAYOUT_SW128);
      int qk_slot = 0, pv_slot = 0;
      uint32_t qk_ph = 0, pv_ph = 0;
      auto issue_qk = [&](int j) {  // S[j&1] = Q~ K(j)^T
        mbar_wait(&bars[K_FULL0 + qk_slot], qk_ph, p.err_flag);
        tc_fence_after();
        const uint64_t k_desc =
            umma_smem_desc(smem_u32(sK + qk_slot * kTileBytes), 16, 1024, UMMA_LAYOUT_SW128);
        const uint32_t s_col = tmem_base + kColS + (j & 1) * 128;
#pragma unroll
        for (int kk = 0; kk < kHeadDim / 16; ++kk) {
          const uint64_t off = static_cast<uint64_t>(((kk & 3) * 32 + (kk >> 2) * kHalfTileBytes) >> 4);
          umma_ss(s_col, q_desc + off, k_desc + off, idesc_qk, kk > 0 ? 1u : 0u);
        }
        umma_commit(&bars[K_EMPTY0 + qk_slot]);
        umma_commit(&bars[S_FULL0 + (j & 1)]);
        if (++qk_slot == kStages) {
          qk_slot = 0;
          qk_ph ^= 1;
        }
      };
      auto issue_pv = [&](int j) {  // O[j&1] += P(j) V(j); P(j) = 64 packed columns over S[j&1]
        const int s = j & 1;
        const uint32_t par = (j >> 1) & 1;
        mbar_wait(&bars[V_FULL0 + pv_slot], pv_ph, p.err_flag);
        // MN-major SW128 operand V: 64-wide d chunks one 16 KB box apart (LBO), 8-key groups
        // 1024 B apart (SBO); 16 keys = 2048 B inside each box.
        const uint64_t v_desc =
            umma_smem_desc(smem_u32(sV + pv_slot * kTileBytes), kHalfTileBytes, 1024, UMMA_LAYOUT_SW128);
        const uint32_t p_col = tmem_base + kColS + s * 128;
        const uint32_t o_col = tmem_base + kColO + s * 128;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          mbar_wait(&bars[P_FULL00 + 2 * s + hh], par, p.err_flag);
          tc_fence_after();
#pragma unroll
          for (int kk = hh * 4; kk < hh * 4 + 4; ++kk) {
            const uint64_t off = static_cast<uint64_t>((kk * 16 * 128) >> 4);
            umma_ts(o_col, p_col + kk * 8, v_desc + off, idesc_pv, (j > 1 || kk > 0) ? 1u : 0u);
          }
        }
        umma_commit(&bars[V_EMPTY0 + pv_slot]);
        if (++pv_slot == kStages) {
          pv_slot = 0;
          pv_ph ^= 1;
        }
      };
      mbar_wait(&bars[Q_READY], 0, p.err_flag);
      tc_fence_after();
      issue_qk(0);
      if (n_tiles > 1) issue_qk(1);
      for (int j = 0; j < n_tiles; ++j) {
        issue_pv(j);
        if (j + 2 < n_tiles) issue_qk(j + 2);
      }
      umma_commit(&bars[ALL_DONE]);
    }
  } else {
    // =============================== softmax / epilogue ===============================
    const int quad = warp & 3;                 // TMEM lane quadrant this warp may touch
    const int strm = (warp - 2) >> 2;          // 0: even tiles, 1: odd tiles
    const int row = quad * 32 + lane;
    const int st = threadIdx.x - 64;           // 0..255
    const uint32_t lane_base = static_cast<uint32_t>(quad * 32) << 16;
    const long long q_row = q_row0 + row;
    const float c = dense ? p.qk_scale : 1.0f;  // see generation 2
    const uint32_t tmem_S = tmem_base + kColS + strm * 128 + lane_base;
    const uint32_t tmem_O = tmem_base + kColO + strm * 128 + lane_base;

    float m_used = -INFINITY, l_sum = 0.f;

    if (n_tiles > 0) {
      mbar_wait(&bars[Q_FULL], 0, p.err_flag);
      if (!dense) {
        // ref :87-88  q = (q * qk_scale).to(dtype) — elementwise, so swizzle-agnostic
        uint4* q4 = reinterpret_cast<uint4*>(sQ);
#pragma unroll 4
        for (int i = 0; i < kTileBytes / 16 / 256; ++i) {
          uint4 v = q4[st + i * 256];
          uint32_t* e = reinterpret_cast<uint32_t*>(&v);
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const float2 f = unpack2<kBF16>(e[t]);
            e[t] = pack2<kBF16>(f.x * p.qk_scale, f.y * p.qk_scale);
          }
          q4[st + i * 256] = v;
        }
        fence_proxy_async_smem();
      }
      mbar_arrive(&bars[Q_READY]);

      BlockWalker it(s_mask, nwords);
      if (strm == 1) it.next();  // odd stream starts at the second live block
      int u = 0;                 // index of the tile inside this stream
      for (int blk = it.next(); blk >= 0; it.next(), blk = it.next(), ++u) {
        // ref :113-114 — text_amp added in log2 units to text key blocks (sparse class, c == 1)
        const float amp = (!dense && blk >= p.text_block_start) ? p.text_amp : 0.f;
        mbar_wait(&bars[S_FULL0 + strm], u & 1, p.err_flag);
        tc_fence_after();
        float s[kBlock];
        {
          uint32_t* su = reinterpret_cast<uint32_t*>(s);
          tmem_ld32(tmem_S, su);
          tmem_ld32(tmem_S + 32, su + 32);
          tmem_ld32(tmem_S + 64, su + 64);
          tmem_ld32(tmem_S + 96, su + 96);
          tmem_ld_wait();
        }
        const long long col0 = static_cast<long long>(blk) * kBlock;
        if (col0 + kBlock > kv_limit) {  // rare: only the tile that straddles seqlen
          asm volatile("" ::: "memory");
#pragma unroll
          for (int i = 0; i < kBlock; ++i)
            if (col0 + i >= kv_limit) s[i] = -INFINITY;  // ref :117-118
        }
        float mx0 = fmaxf(s[0], s[1]), mx1 = fmaxf(s[2], s[3]);
        float mx2 = fmaxf(s[4], s[5]), mx3 = fmaxf(s[6], s[7]);
#pragma unroll
        for (int i = 8; i < kBlock; i += 8) {
          mx0 = fmaxf(mx0, fmaxf(s[i], s[i + 1]));
          mx1 = fmaxf(mx1, fmaxf(s[i + 2], s[i + 3]));
          mx2 = fmaxf(mx2, fmaxf(s[i + 4], s[i + 5]));
          mx3 = fmaxf(mx3, fmaxf(s[i + 6], s[i + 7]));
        }
        const float mx = fmaf(fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)), c, amp);
        const float m_cand = fmaxf(m_used, mx);
        // Lazy rescale (see generation 2): move the reference point only when the max grew by
        // more than 2^8.  S_FULL(j) implies PV(j-2) retired, so O_s is quiescent here.
        const bool need = (m_cand - m_used) > 8.0f;  // false for NaN (-inf - -inf)
        if (__any_sync(0xffffffffu, need)) {
          if (u > 0) {
            const float alpha = (m_cand == -INFINITY) ? 1.0f : fast_exp2(m_used - m_cand);
#pragma unroll 1
            for (int cc = 0; cc < 128; cc += 16) {
              uint32_t o[16];
              tmem_ld16(tmem_O + cc, o);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
              tmem_st16(tmem_O + cc, o);
            }
            tmem_st_wait();
            l_sum *= alpha;
          }
          m_used = m_cand;
        }
        const float off = amp - ((m_used == -INFINITY) ? 0.f : m_used);
        const f32x2 c2 = f2_pack(c, c), off2 = f2_pack(off, off);
        f32x2 sum2 = f2_pack(0.f, 0.f);
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
#pragma unroll
          for (int cc = hh * 64; cc < hh * 64 + 64; cc += 32) {
            uint32_t pk[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const f32x2 x = f2_fma(f2_pack(s[cc + 2 * i], s[cc + 2 * i + 1]), c2, off2);
              float x0, x1, p0, p1;
              f2_unpack(x, x0, x1);
              f32x2 pp;
              if (kPolyEvery > 0 && (i % kPolyEvery) == kPolyEvery - 1) {
                pp = f2_exp2_poly(f2_pack(fmaxf(x0, -125.f), fmaxf(x1, -125.f)));
                f2_unpack(pp, p0, p1);
              } else {
                p0 = fast_exp2(x0);
                p1 = fast_exp2(x1);
                pp = f2_pack(p0, p1);
              }
              sum2 = f2_add(sum2, pp);       // l accumulates the unrounded p (ref :131)
              pk[i] = pack2<kBF16>(p0, p1);  // P rounded to the input dtype (ref :128)
            }
            tmem_st16(tmem_S + (cc >> 1), pk);
          }
          tmem_st_wait();
          tc_fence_before();
          mbar_arrive(&bars[P_FULL00 + 2 * strm + hh]);
        }
        float sum0, sum1;
        f2_unpack(sum2, sum0, sum1);
        l_sum += sum0 + sum1;
      }
      mbar_wait(&bars[ALL_DONE], 0, p.err_flag);  // every MMA retired; the Q tile is dead
      tc_fence_after();
      s_exch[(strm * 2 + 0) * 128 + row] = m_used;
      s_exch[(strm * 2 + 1) * 128 + row] = l_sum;
      named_bar_sync(1, 256);
    }

    // ---- combine the two streams and store O / l (ref :135-136); rows past the limit are 0 ----
    const bool in_tensor = q_row < p.q_rows;
    const bool zero_row = (n_tiles == 0) || (!dense && q_row >= q_limit_sparse);
    float a0 = 0.f, a1 = 0.f, inv_l = 0.f;
    if (n_tiles > 0) {
      const float m0 = s_exch[0 * 128 + row], l0 = s_exch[1 * 128 + row];
      const float m1 = n_tiles > 1 ? s_exch[2 * 128 + row] : -INFINITY;
      const float l1 = n_tiles > 1 ? s_exch[3 * 128 + row] : 0.f;
      const float m = fmaxf(m0, m1);
      a0 = (m0 == -INFINITY) ? 0.f : fast_exp2(m0 - m);
      a1 = (m1 == -INFINITY) ? 0.f : fast_exp2(m1 - m);
      inv_l = 1.0f / (l0 * a0 + l1 * a1);
      a0 *= inv_l;
      a1 *= inv_l;
    }
    if (zero_row) a0 = a1 = 0.f;
    const int col_base = strm * 64;  // this warpgroup's output columns
    long long o_off = b * p.o_stride_b + q_row * p.o_stride_s + static_cast<long long>(h) * p.o_stride_h;
    uint16_t* obase = reinterpret_cast<uint16_t*>(p.out);
    int n_dst = 1;  // > 1 only for replicated (text) rows under Ulysses
    if (p.sp_world > 0) {
      // Fused Ulysses exchange (ref xdit_ring_atten.py:206-219): see generation 2.
      const long long n_img = p.sp_rows * p.sp_world;
      const int gh = p.sp_rank * p.heads + h;
      if (q_row < n_img) {
        const int owner = static_cast<int>(q_row / p.sp_rows);
        obase = reinterpret_cast<uint16_t*>(p.peer_out[owner]);
        o_off = ((q_row - owner * p.sp_rows) * p.sp_heads_total + gh) * kHeadDim;
      } else {
        n_dst = p.sp_world;
        o_off = ((p.sp_rows + (q_row - n_img)) * p.sp_heads_total + gh) * kHeadDim;
      }
    }
    const uint32_t tO0 = tmem_base + kColO + lane_base + col_base;
    const uint32_t tO1 = tO0 + 128;
#pragma unroll 1
    for (int cc = 0; cc < 64; cc += 32) {
      uint32_t o0[32], o1[32];
      if (n_tiles > 0) {
        tmem_ld32(tO0 + cc, o0);
        if (n_tiles > 1) tmem_ld32(tO1 + cc, o1);
        tmem_ld_wait();
      }
      if (n_tiles <= 1) {
#pragma unroll
        for (int i = 0; i < 32; ++i) o1[i] = 0;
      }
      if (n_tiles == 0) {
#pragma unroll
        for (int i = 0; i < 32; ++i) o0[i] = 0;
      }
      if (in_tensor) {
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          uint4 v;
          uint32_t* e = reinterpret_cast<uint32_t*>(&v);
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            float x = 0.f, y = 0.f;
            if (!zero_row) {
              x = fmaf(__uint_as_float(o0[i + 2 * t]), a0, __uint_as_float(o1[i + 2 * t]) * a1);
              y = fmaf(__uint_as_float(o0[i + 2 * t + 1]), a0, __uint_as_float(o1[i + 2 * t + 1]) * a1);
            }
            e[t] = pack2<kBF16>(x, y);
          }
          const long long col = col_base + cc + i;
          if (n_dst > 1) {
            for (int r = 0; r < n_dst; ++r)
              *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.peer_out[r]) + o_off + col) = v;
          } else if (!p.out_f32) {
            *reinterpret_cast<uint4*>(obase + o_off + col) = v;
          } else {  // wan/…:530-532: the 16-bit result is widened back to the query dtype
            float* orow32 = reinterpret_cast<float*>(p.out) + o_off;
            float2 t0 = unpack2<kBF16>(e[0]), t1 = unpack2<kBF16>(e[1]);
            const float4 f0 = make_float4(t0.x, t0.y, t1.x, t1.y);
            t0 = unpack2<kBF16>(e[2]);
            t1 = unpack2<kBF16>(e[3]);
            const float4 f1 = make_float4(t0.x, t0.y, t1.x, t1.y);
            *reinterpret_cast<float4*>(orow32 + col) = f0;
            *reinterpret_cast<float4*>(orow32 + col + 4) = f1;
          }
        }
      }
    }
  }

  // ---- teardown ----
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<kTmemCols>(tmem_base);
}

}  // namespace

int launch_carved_attn_v4(const CUtensorMap& tm_q, const CUtensorMap& tm_k, const CUtensorMap& tm_v,
                          const attn::KernelParams& p, unsigned grid, bool bf16, cudaStream_t stream) {
  auto kern = bf16 ? carved_attn_v4_kernel<true> : carved_attn_v4_kernel<false>;
  cudaError_t ce = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
  if (ce != cudaSuccess) return set_cuda_error(ce, "cudaFuncSetAttribute(carved_attn_v4)");
  kern<<<grid, kThreads, kSmemBytes, stream>>>(tm_q, tm_k, tm_v, p);
  ce = cudaGetLastError();
  if (ce != cudaSuccess) return set_cuda_error(ce, "carved_attn_v4 launch");
  return JENGA_OK;
}

}  // namespace jenga
