// Carved attention forward, generation 6 (experimental, JENGA_ATTN_KERNEL=v6) — three key
// blocks in flight over ONE accumulator, every tcgen05.mma at N=128.
//
// Why it exists.  Generation 2 (carved_attn.cu, the default) pipelines on 64-key half tiles, so
// its QK^T MMAs are N=64 — 48 cycles instead of 32 (tools/umma_probe.cu) — and its tensor pipe is
// ~90 % busy delivering ~70 % of peak.  Full-tile (N=128) designs with two tiles in flight were
// built and measured (generations 3-5, removed; numbers in profiles/README.md): the per-tile chain
//     QK^T -> commit -> softmax -> hand P over -> PV -> (next QK^T)
// has ~600 cycles of fixed latency plus ~2000 cycles of softmax for a lone warp, against 1024
// tensor cycles per tile, so two tiles in flight (also FlashAttention-4's layout) leave the
// tensor pipe ~60 % active.  A third tile does not fit in TMEM next to per-tile accumulators
// (3 x (S + O) = 768 columns), so this generation gives all tiles ONE accumulator:
//
//   one CTA per SM, 512 TMEM columns:  S0 [0,128)  S1 [128,256)  S2 [256,384)  O [384,512)
//   tile j uses S[j % 3] and softmax warpgroup j % 3 (one thread per query row);
//   issue order  QK(0) QK(1) QK(2) | PV(0) QK(3) | PV(1) QK(4) | ...  (TMA loads in the same order)
//
// Sharing O means all three warpgroups must exponentiate against the SAME reference point
// without meeting on the critical path.  The reference of tile t is a pure function of values
// every thread can already see when S(t) arrives:
//     m_ref(0) = m_ref(1) = m_ref(2) = rowmax(tile 0)
//     m_ref(t) = lazy(m_ref(t-1), rowmax(tile t-3)),  lazy(m, x) = max(m,x) if it exceeds m by 2^8
// rowmax(tile t-3) is published to shared memory before P(t-3) is handed over, PV(t-3) waits for
// that hand-over and QK(t) is issued behind PV(t-3) — so by the time S(t) exists the value is
// visible.  Each thread replays the recurrence for the tiles the other warpgroups own, rescales
// its own partial row sum, and the owner of tile t rescales O (after PV(t-1) retired) when the
// reference moves at t — rare after the first tiles.  p = 2^(s - m_ref) may exceed 1 while a
// row maximum is still growing; fp32/bf16 hold that easily (a guard flags growth > 2^100 within
// three tiles).  O/l does not depend on the reference point: this is the reference's online
// softmax (attention_block_triton_diffres.py:121-135) up to fp32 rounding.
// With no max-before-exp dependency the scores stream through registers 32 columns at a time.
//
// Measured (HY-720p, drop 0.7): 1485 cycles per tile, tensor pipe 69.7 % active under ncu — the
// same as generation 2 — and 45.0 ms vs 42.5 ms in the power-capped back-to-back bench, so
// generation 2 stays the default.  Parity tests cover both (tests/test_attn_gpu.py).
//
// Same C-ABI, same KernelParams, same results (up to fp32 rounding) as generation 2.
#include "carved_attn_common.cuh"

#ifndef JENGA_POLY_EVERY
#define JENGA_POLY_EVERY 3
#endif
#ifndef JENGA_V6_KSTAGES
#define JENGA_V6_KSTAGES 3
#endif
#ifndef JENGA_V6_VSTAGES
#define JENGA_V6_VSTAGES 2
#endif

namespace jenga {

namespace {

using attn::BlockWalker;
using attn::KernelParams;

constexpr int kBlock = 128;
constexpr int kHeadDim = 128;
constexpr int kStagesK = JENGA_V6_KSTAGES;         // K ring depth
constexpr int kStagesV = JENGA_V6_VSTAGES;         // V ring depth
constexpr int kHalfTileBytes = kBlock * 64 * 2;    // 16 KB: 128 rows x 64 d
constexpr int kTileBytes = 2 * kHalfTileBytes;     // 32 KB
constexpr int kMaxMaskWords = 256;
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kColS = 0, kColO = 384;         // S[b] at kColS + 128 b
constexpr int kPolyEvery = JENGA_POLY_EVERY;
constexpr int kSoftmaxThreads = 384;               // 3 warpgroups x 128 rows
constexpr int kThreads = 64 + kSoftmaxThreads;     // + TMA warp, MMA warp
constexpr int JENGA_DEV_RANGE = 0x7002;            // a row maximum jumped by more than 2^100 in one tile

constexpr int kOffQ = 0;                           // Q tile; reused for the final (m, l) exchange
constexpr int kOffK = kTileBytes;
constexpr int kOffV = kOffK + kStagesK * kTileBytes;
constexpr int kOffBars = kOffV + kStagesV * kTileBytes;
constexpr int kOffMask = kOffBars + 256;
constexpr int kPubSlots = 8;
constexpr int kOffPub = kOffMask + kMaxMaskWords * 4;   // float [kPubSlots][128 rows]: row maxima of recent tiles
constexpr int kSmemBytes = 1024 + kOffPub + kPubSlots * 128 * 4;
static_assert(kSmemBytes <= 232448, "exceeds the 227 KB per-CTA shared memory limit");

enum BarId {
  Q_FULL = 0, Q_READY,
  K_FULL0, K_EMPTY0 = K_FULL0 + kStagesK,
  V_FULL0 = K_EMPTY0 + kStagesK, V_EMPTY0 = V_FULL0 + kStagesV,
  S_FULL0 = V_EMPTY0 + kStagesV, S_FULL1, S_FULL2,
  P_FULL00, P_FULL01, P_FULL10, P_FULL11, P_FULL20, P_FULL21,   // [S buffer][key half]
  PV_DONE0, PV_DONE1, PV_DONE2,   // PV of the tile in S buffer b retired
  M0_READY, ALL_DONE,
  NUM_BARS
};
static_assert(NUM_BARS * 8 + 8 <= 256, "barrier block overflow");

template <bool kBF16>
__global__ void __launch_bounds__(kThreads, 1)
carved_attn_v6_kernel(const __grid_constant__ CUtensorMap tm_q,
                      const __grid_constant__ CUtensorMap tm_k,
                      const __grid_constant__ CUtensorMap tm_v, const KernelParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint8_t* sQ = smem + kOffQ;
  uint8_t* sK = smem + kOffK;
  uint8_t* sV = smem + kOffV;
  float* s_exch = reinterpret_cast<float*>(smem + kOffQ);  // final exchange: l [3][128]
  float* s_pub = reinterpret_cast<float*>(smem + kOffPub);
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
      mbar_init(&bars[i], i == Q_READY ? static_cast<uint32_t>(kSoftmaxThreads)
                                          : (((i >= P_FULL00 && i <= P_FULL21) || i == M0_READY) ? 128u : 1u));
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
      // Loads are issued in the order the tensor pipe consumes them — K0 K1 K2 | V0 K3 | V1 K4 |
      // ... — so a wait for a free V slot never holds back a K tile that is needed earlier.
      BlockWalker itK(s_mask, nwords), itV(s_mask, nwords);
      int ks = 0, vs = 0;
      uint32_t kph = 0, vph = 0;
      auto load_k = [&](int blk) {
        uint8_t* kd = sK + ks * kTileBytes;
        mbar_wait(&bars[K_EMPTY0 + ks], kph ^ 1, p.err_flag);
        mbar_arrive_expect_tx(&bars[K_FULL0 + ks], kTileBytes);
        tma_load_4d(kd, &tm_k, &bars[K_FULL0 + ks], 0, blk * kBlock, h, b);
        tma_load_4d(kd + kHalfTileBytes, &tm_k, &bars[K_FULL0 + ks], 64, blk * kBlock, h, b);
        if (++ks == kStagesK) {
          ks = 0;
          kph ^= 1;
        }
      };
      auto load_v = [&](int blk) {
        uint8_t* vd = sV + vs * kTileBytes;
        mbar_wait(&bars[V_EMPTY0 + vs], vph ^ 1, p.err_flag);
        mbar_arrive_expect_tx(&bars[V_FULL0 + vs], kTileBytes);
        tma_load_4d(vd, &tm_v, &bars[V_FULL0 + vs], 0, blk * kBlock, h, b);
        tma_load_4d(vd + kHalfTileBytes, &tm_v, &bars[V_FULL0 + vs], 64, blk * kBlock, h, b);
        if (++vs == kStagesV) {
          vs = 0;
          vph ^= 1;
        }
      };
      for (int i = 0; i < 3; ++i) {
        const int blk = itK.next();
        if (blk >= 0) load_k(blk);
      }
      for (int blkv = itV.next(); blkv >= 0; blkv = itV.next()) {
        load_v(blkv);
        const int blk = itK.next();
        if (blk >= 0) load_k(blk);
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
      int qk_buf = 0, pv_buf = 0;      // S buffer of the next QK / PV (tile index mod 3)
      uint32_t pv_par = 0;             // phase parity of P_FULL[pv_buf][*]
      auto issue_qk = [&]() {  // S[j % 3] = Q~ K(j)^T
        mbar_wait(&bars[K_FULL0 + qk_slot], qk_ph, p.err_flag);
        tc_fence_after();
        const uint64_t k_desc =
            umma_smem_desc(smem_u32(sK + qk_slot * kTileBytes), 16, 1024, UMMA_LAYOUT_SW128);
        const uint32_t s_col = tmem_base + kColS + qk_buf * 128;
#pragma unroll
        for (int kk = 0; kk < kHeadDim / 16; ++kk) {
          const uint64_t off = static_cast<uint64_t>(((kk & 3) * 32 + (kk >> 2) * kHalfTileBytes) >> 4);
          umma_ss(s_col, q_desc + off, k_desc + off, idesc_qk, kk > 0 ? 1u : 0u);
        }
        umma_commit(&bars[K_EMPTY0 + qk_slot]);
        umma_commit(&bars[S_FULL0 + qk_buf]);
        if (++qk_slot == kStagesK) {
          qk_slot = 0;
          qk_ph ^= 1;
        }
        if (++qk_buf == 3) qk_buf = 0;
      };
      auto issue_pv = [&](bool first) {  // O += P(j) V(j); P(j) = 64 packed columns over S[j % 3]
        mbar_wait(&bars[V_FULL0 + pv_slot], pv_ph, p.err_flag);
        // MN-major SW128 operand V: 64-wide d chunks one 16 KB box apart (LBO), 8-key groups
        // 1024 B apart (SBO); 16 keys = 2048 B inside each box.
        const uint64_t v_desc =
            umma_smem_desc(smem_u32(sV + pv_slot * kTileBytes), kHalfTileBytes, 1024, UMMA_LAYOUT_SW128);
        const uint32_t p_col = tmem_base + kColS + pv_buf * 128;
        const uint32_t o_col = tmem_base + kColO;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          mbar_wait(&bars[P_FULL00 + 2 * pv_buf + hh], pv_par, p.err_flag);
          tc_fence_after();
#pragma unroll
          for (int kk = hh * 4; kk < hh * 4 + 4; ++kk) {
            const uint64_t off = static_cast<uint64_t>((kk * 16 * 128) >> 4);
            umma_ts(o_col, p_col + kk * 8, v_desc + off, idesc_pv, (!first || kk > 0) ? 1u : 0u);
          }
        }
        umma_commit(&bars[V_EMPTY0 + pv_slot]);
        umma_commit(&bars[PV_DONE0 + pv_buf]);
        if (++pv_slot == kStagesV) {
          pv_slot = 0;
          pv_ph ^= 1;
        }
        if (++pv_buf == 3) {
          pv_buf = 0;
          pv_par ^= 1;
        }
      };
      mbar_wait(&bars[Q_READY], 0, p.err_flag);
      tc_fence_after();
      issue_qk();
      if (n_tiles > 1) issue_qk();
      if (n_tiles > 2) issue_qk();
      for (int j = 0; j < n_tiles; ++j) {
        issue_pv(j == 0);
        if (j + 3 < n_tiles) issue_qk();  // overwrites S[j % 3] after PV(j): in-order pipe
      }
      umma_commit(&bars[ALL_DONE]);
    }
  } else {
    // =============================== softmax / epilogue ===============================
    const int quad = warp & 3;                 // TMEM lane quadrant this warp may touch
    const int wg = (warp - 2) >> 2;            // 0..2: owns tiles j with j % 3 == wg, S buffer wg
    const int row = quad * 32 + lane;
    const int st = threadIdx.x - 64;           // 0..383
    const uint32_t lane_base = static_cast<uint32_t>(quad * 32) << 16;
    const long long q_row = q_row0 + row;
    const float c = dense ? p.qk_scale : 1.0f;  // see generation 2
    const uint32_t tmem_S = tmem_base + kColS + wg * 128 + lane_base;
    const uint32_t tmem_O = tmem_base + kColO + lane_base;
    const float* pub = s_pub + row;            // pub[(t & 7) * 128] = scaled row maximum of tile t

    float m_ref = -INFINITY, l_sum = 0.f;
    // lazy(m, x): move the reference only when the maximum grew by more than 2^8
    auto lazy = [](float m, float x) {
      const float cand = fmaxf(m, x);
      return (cand - m) > 8.0f ? cand : m;     // false for NaN (-inf - -inf)
    };

    if (n_tiles > 0) {
      mbar_wait(&bars[Q_FULL], 0, p.err_flag);
      if (!dense) {
        // ref :87-88  q = (q * qk_scale).to(dtype) — elementwise, so swizzle-agnostic
        uint4* q4 = reinterpret_cast<uint4*>(sQ);
        for (int i = st; i < kTileBytes / 16; i += kSoftmaxThreads) {
          uint4 v = q4[i];
          uint32_t* e = reinterpret_cast<uint32_t*>(&v);
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const float2 f = unpack2<kBF16>(e[t]);
            e[t] = pack2<kBF16>(f.x * p.qk_scale, f.y * p.qk_scale);
          }
          q4[i] = v;
        }
        fence_proxy_async_smem();
      }
      mbar_arrive(&bars[Q_READY]);

      BlockWalker it(s_mask, nwords);
      for (int i = 0; i < wg; ++i) it.next();  // my first live block is the wg-th
      int j = wg;                              // global tile index
      uint32_t par = 0;                        // phase parity of S_FULL[wg]
      for (int blk = it.next(); blk >= 0; it.next(), it.next(), blk = it.next(), j += 3, par ^= 1) {
        // ref :113-114 — text_amp added in log2 units to text key blocks (sparse class, c == 1)
        const float amp = (!dense && blk >= p.text_block_start) ? p.text_amp : 0.f;
        const long long col0 = static_cast<long long>(blk) * kBlock;
        const bool straddles = col0 + kBlock > kv_limit;  // rare: the tile that crosses seqlen
        mbar_wait(&bars[S_FULL0 + wg], par, p.err_flag);
        tc_fence_after();

        // ---- reference point of this tile ----
        if (j == 0) {
          // no history: one extra pass over S(0) for the true row maximum
          float mx = -INFINITY;
#pragma unroll 1
          for (int cc = 0; cc < kBlock; cc += 32) {
            uint32_t su[32];
            tmem_ld32(tmem_S + cc, su);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float v = __uint_as_float(su[i]);
              mx = fmaxf(mx, (straddles && col0 + cc + i >= kv_limit) ? -INFINITY : v);
            }
          }
          m_ref = fmaf(mx, c, amp);
          s_pub[row] = m_ref;                 // slot 0
          mbar_arrive(&bars[M0_READY]);       // release: tiles 1 and 2 start from this reference
        } else if (j < 3) {
          mbar_wait(&bars[M0_READY], 0, p.err_flag);
          m_ref = pub[0];
        } else {
          // replay m_ref(j-2), m_ref(j-1), m_ref(j) from m_ref(j-3) (mine) and the published maxima
          float m = m_ref;
          if (j >= 5) m = lazy(m, pub[((j - 5) & 7) * 128]);
          if (j >= 4) m = lazy(m, pub[((j - 4) & 7) * 128]);
          const float m_before = m;           // m_ref(j-1)
          m = lazy(m, pub[((j - 3) & 7) * 128]);
          if (m != m_ref) l_sum *= fast_exp2(m_ref - m);
          m_ref = m;
          // the reference moved AT this tile: O (relative to m_ref(j-1)) is rescaled by this
          // tile's owner, after PV(j-1) retired and before P(j) is handed over
          const bool moved = m != m_before;
          if (__any_sync(0xffffffffu, moved)) {
            // PV(j-1) lives in buffer (wg+2)%3; its phase is mine (wg >= 1) or my previous one (wg == 0).
            // S_FULL(j) implies PV(j-3) retired, so that barrier is exactly one phase behind at most.
            mbar_wait(&bars[PV_DONE0 + (wg + 2) % 3], wg ? par : par ^ 1u, p.err_flag);
            tc_fence_after();
            const float alpha = moved ? fast_exp2(m_before - m) : 1.0f;
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
          }
        }
        const float off = amp - ((m_ref == -INFINITY) ? 0.f : m_ref);
        const f32x2 c2 = f2_pack(c, c), off2 = f2_pack(off, off);
        f32x2 sum2 = f2_pack(0.f, 0.f);
        float mx0 = -INFINITY, mx1 = -INFINITY;

        // ---- stream the 128 scores through registers, 32 columns at a time ----
        auto chunk = [&](uint32_t (&su)[32], int cc) {
          float* sf = reinterpret_cast<float*>(su);
          if (straddles) {
            asm volatile("" ::: "memory");  // keep this a branch, not 32 predicated selects
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (col0 + cc + i >= kv_limit) sf[i] = -INFINITY;  // ref :117-118
          }
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            mx0 = fmaxf(mx0, fmaxf(sf[i], sf[i + 1]));
            mx1 = fmaxf(mx1, fmaxf(sf[i + 2], sf[i + 3]));
          }
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const f32x2 x = f2_fma(f2_pack(sf[2 * i], sf[2 * i + 1]), c2, off2);
            float x0, x1, p0, p1;
            f2_unpack(x, x0, x1);
            f32x2 pp;
            if (kPolyEvery > 0 && (i % kPolyEvery) == kPolyEvery - 1) {
              pp = f2_exp2_poly(f2_pack(fmaxf(x0, -125.f), fmaxf(x1, -125.f)));  // x <= 100 (guard below)
              f2_unpack(pp, p0, p1);
            } else {
              p0 = fast_exp2(x0);
              p1 = fast_exp2(x1);
              pp = f2_pack(p0, p1);
            }
            sum2 = f2_add(sum2, pp);       // l accumulates the unrounded p (ref :131)
            pk[i] = pack2<kBF16>(p0, p1);  // P rounded to the input dtype (ref :128)
          }
          tmem_st16(tmem_S + (cc >> 1), pk);  // P over columns already consumed
        };
        uint32_t sa[32], sb[32];
        tmem_ld32(tmem_S, sa);
        tmem_ld_wait();
        tmem_ld32(tmem_S + 32, sb);
        chunk(sa, 0);
        tmem_ld_wait();
        tmem_ld32(tmem_S + 64, sa);
        chunk(sb, 32);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&bars[P_FULL00 + 2 * wg + 0]);
        tmem_ld_wait();
        tmem_ld32(tmem_S + 96, sb);
        chunk(sa, 64);
        tmem_ld_wait();
        chunk(sb, 96);
        const float m_loc = fmaf(fmaxf(mx0, mx1), c, amp);
        s_pub[(j & 7) * 128 + row] = m_loc;   // published before the hand-over below
        if (m_loc - m_ref > 100.f && p.err_flag) atomicExch(p.err_flag, JENGA_DEV_RANGE);
        float sum0, sum1;
        f2_unpack(sum2, sum0, sum1);
        l_sum += sum0 + sum1;
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&bars[P_FULL00 + 2 * wg + 1]);
      }
      mbar_wait(&bars[ALL_DONE], 0, p.err_flag);  // every MMA retired; the Q tile is dead
      tc_fence_after();
      // bring my partial row sum to the final reference m_ref(n_tiles - 1)
      if (j - 3 >= 0 && j - 3 < n_tiles) {        // I processed at least one tile; the last was j - 3
        float m = m_ref;
        for (int t = j - 2; t < n_tiles; ++t)
          if (t >= 3) m = lazy(m, pub[((t - 3) & 7) * 128]);
        if (m != m_ref) l_sum *= fast_exp2(m_ref - m);
      }
      s_exch[wg * 128 + row] = l_sum;
      named_bar_sync(1, kSoftmaxThreads);
    }

    // ---- store O / l (ref :135-136); rows past the limit are zeros (:156) ----
    const bool in_tensor = q_row < p.q_rows;
    const bool zero_row = (n_tiles == 0) || (!dense && q_row >= q_limit_sparse);
    float inv_l = 0.f;
    if (n_tiles > 0 && !zero_row) inv_l = 1.0f / (s_exch[row] + s_exch[128 + row] + s_exch[256 + row]);
    const int col_base = wg * 64;  // warpgroups 0 and 1 write 64 output columns each
    long long o_off = b * p.o_stride_b + q_row * p.o_stride_s + static_cast<long long>(h) * p.o_stride_h;
    uint16_t* obase = reinterpret_cast<uint16_t*>(p.out);
    int n_dst = 1;  // > 1 only for replicated (text) rows under Ulysses
    if (p.sp_world > 0) {
      // Fused Ulysses exchange (ref xdit_ring_atten.py:206-219): see generation 2.
      const long long n_img = p.sp_rows * p.sp_world;
      const int gh = p.sp_head_base + h;
      if (q_row < n_img) {
        const int owner = static_cast<int>(q_row / p.sp_rows);
        obase = reinterpret_cast<uint16_t*>(p.peer_out[owner]);
        o_off = ((q_row - owner * p.sp_rows) * p.sp_heads_total + gh) * kHeadDim;
      } else {
        n_dst = p.sp_world;
        o_off = ((p.sp_rows + (q_row - n_img)) * p.sp_heads_total + gh) * kHeadDim;
      }
    }
    if (wg < 2) {
#pragma unroll 1
      for (int cc = 0; cc < 64; cc += 32) {
        uint32_t o[32];
        if (n_tiles > 0) {
          tmem_ld32(tmem_O + col_base + cc, o);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = 0;
        }
        if (in_tensor) {
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            uint4 v;
            uint32_t* e = reinterpret_cast<uint32_t*>(&v);
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const float x = zero_row ? 0.f : __uint_as_float(o[i + 2 * t]) * inv_l;
              const float y = zero_row ? 0.f : __uint_as_float(o[i + 2 * t + 1]) * inv_l;
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
  }

  // ---- teardown ----
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<kTmemCols>(tmem_base);
}

}  // namespace

int launch_carved_attn_v6(const CUtensorMap& tm_q, const CUtensorMap& tm_k, const CUtensorMap& tm_v,
                          const attn::KernelParams& p, unsigned grid, bool bf16, cudaStream_t stream) {
  auto kern = bf16 ? carved_attn_v6_kernel<true> : carved_attn_v6_kernel<false>;
  cudaError_t ce = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
  if (ce != cudaSuccess) return set_cuda_error(ce, "cudaFuncSetAttribute(carved_attn_v6)");
  kern<<<grid, kThreads, kSmemBytes, stream>>>(tm_q, tm_k, tm_v, p);
  ce = cudaGetLastError();
  if (ce != cudaSuccess) return set_cuda_error(ce, "carved_attn_v6 launch");
  return JENGA_OK;
}

}  // namespace jenga
