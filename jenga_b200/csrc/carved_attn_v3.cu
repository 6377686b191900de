// Carved attention forward, generation 3 — every tcgen05.mma at N=128.
//
// Measured on B200 (tools/umma_probe.cu): an M128 N64 K16 MMA takes 48 cycles with A in shared
// memory and 45 with A in TMEM (ideal 32), while N=128 runs at exactly 64 cycles either way.
// Generation 2 (carved_attn.cu) pipelines on 64-key half tiles, so its QK^T MMAs are N=64 and
// the tensor pipe is 89 % busy delivering 69 % of peak.  This generation keeps full 128-key
// tiles and gets its overlap from TMEM capacity instead:
//
//   one CTA per SM, 512 TMEM columns:  S0 [0,128)  S1 [128,256)  O [256,384)  Q~ [384,448)
//   * Q~ (pre-scaled, 16-bit packed) lives in TMEM, so S = Q~ K^T is a TS-MMA that reads only K
//     from shared memory; O += P V reads P from TMEM (aliased over S) and V from shared memory.
//   * S is double-buffered per key block: the tensor pipe runs PV(j-1) and QK(j+1) while the
//     softmax warps work on S(j).
//   * 227 KB of shared memory hold a 3-deep ring for K and for V (32 KB tiles).
//   * 8 softmax warps: the two warps of a TMEM lane quadrant split each row's 128 columns.
//     They agree on the softmax reference point WITHOUT a per-tile rendezvous: each publishes
//     its running half-row max to shared memory, and the offset used for tile j is a function
//     of both halves' maxima up to tile j-2 — values that are guaranteed published (tile j's
//     scores exist only after PV(j-2) was issued, i.e. after every softmax thread finished tile
//     j-2).  O/l does not depend on the reference point, so this is the reference's online
//     softmax up to fp32 rounding; exponents stay far below overflow because attention logits
//     move by a few units between tiles (a guard flags > 2^100).
//
// Same C-ABI, same KernelParams, same results as generation 2 (tests run both).
#include "carved_attn_common.cuh"

#ifndef JENGA_POLY_EVERY
#define JENGA_POLY_EVERY 3
#endif

namespace jenga {

namespace {

using attn::BlockWalker;
using attn::KernelParams;

constexpr int kBlock = 128;
constexpr int kHeadDim = 128;
// kSplit softmax warps share one TMEM lane quadrant; each thread owns 128/kSplit columns of its
// row.  kSplit = 2: 8 softmax warps (320 threads); kSplit = 4: 16 softmax warps (576 threads),
// twice the warps per scheduler to hide the TMEM / MUFU / mbarrier latencies.
constexpr int kStages = 3;                     // K ring and V ring depth
constexpr int kHalfTileBytes = kBlock * 64 * 2;  // 16 KB: 128 rows x 64 d
constexpr int kTileBytes = 2 * kHalfTileBytes;   // 32 KB
constexpr int kMaxMaskWords = 256;
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kColS0 = 0, kColO = 256, kColQ = 384;
constexpr int kPolyEvery = JENGA_POLY_EVERY;
#ifndef JENGA_SPLIT_PIPES
#define JENGA_SPLIT_PIPES 1
#endif
constexpr bool kSplitPipes = JENGA_SPLIT_PIPES != 0;

// shared memory (offsets from the 1024-aligned base).  The Q staging tile is dead once Q~ is in
// TMEM; its space is reused for the max-exchange slots and the final row-sum exchange.
constexpr int kOffQ = 0;
constexpr int kOffPub = 0;                          // float [4 slots][kSplit][128 rows] (<= 8 KB)
constexpr int kOffLsum = 4 * 4 * 128 * 4;           // float [kSplit][128]
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
  S_FULL0 = V_EMPTY0 + kStages, S_FULL1, P_FULL0, P_FULL1, PV_DONE,
  NUM_BARS
};
static_assert(NUM_BARS * 8 + 8 <= 256, "barrier block overflow");

constexpr int JENGA_DEV_RANGE = 0x7002;  // score jumped by more than 2^100 within two tiles

template <bool kBF16, int kSplit>
__global__ void __launch_bounds__(64 + 128 * kSplit, 1)
carved_attn_v3_kernel(const __grid_constant__ CUtensorMap tm_q,
                      const __grid_constant__ CUtensorMap tm_k,
                      const __grid_constant__ CUtensorMap tm_v, const KernelParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint8_t* sQ = smem + kOffQ;
  uint8_t* sK = smem + kOffK;
  uint8_t* sV = smem + kOffV;
  float* s_pub = reinterpret_cast<float*>(smem + kOffPub);
  float* s_lsum = reinterpret_cast<float*>(smem + kOffLsum);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBars);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NUM_BARS);
  uint32_t* s_mask = reinterpret_cast<uint32_t*>(smem + kOffMask);

  constexpr int kThreads = 64 + 128 * kSplit;
  constexpr int kSoftmaxThreads = 128 * kSplit;
  constexpr int kCols = 128 / kSplit;      // S columns (keys) per softmax thread
  constexpr int kOCols = 128 / kSplit;     // O columns per softmax thread
  static_assert(kSplit == 2 || kSplit == 4, "");
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

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
  const bool skip_all = (!dense && q_row0 >= q_limit_sparse);

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
      mbar_init(&bars[i], (i == Q_READY || i == P_FULL0 || i == P_FULL1) ? kSoftmaxThreads : 1u);
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
      int j = 0, slot = 0;
      uint32_t ph = 0;
      for (int blk = it.next(); blk >= 0; blk = it.next(), ++j) {
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
      auto issue_qk = [&](int j) {  // S[j&1] = Q~ K(j)^T, A from TMEM
        const int slot = j % kStages;
        mbar_wait(&bars[K_FULL0 + slot], (j / kStages) & 1, p.err_flag);
        tc_fence_after();
        const uint64_t k_desc = umma_smem_desc(smem_u32(sK + slot * kTileBytes), 16, 1024, UMMA_LAYOUT_SW128);
        const uint32_t s_col = tmem_base + kColS0 + (j & 1) * 128;
#pragma unroll
        for (int kk = 0; kk < kHeadDim / 16; ++kk) {
          const uint64_t off = static_cast<uint64_t>(((kk & 3) * 32 + (kk >> 2) * kHalfTileBytes) >> 4);
          umma_ts(s_col, tmem_base + kColQ + kk * 8, k_desc + off, idesc_qk, kk > 0 ? 1u : 0u);
        }
        umma_commit(&bars[K_EMPTY0 + slot]);
        umma_commit(&bars[S_FULL0 + (j & 1)]);
      };
      auto issue_pv = [&](int j) {  // O += P(j) V(j); P(j) sits over S[j&1], slice by slice
        const int slot = j % kStages;
        mbar_wait(&bars[V_FULL0 + slot], (j / kStages) & 1, p.err_flag);
        mbar_wait(&bars[P_FULL0 + (j & 1)], (j >> 1) & 1, p.err_flag);
        tc_fence_after();
        const uint64_t v_desc =
            umma_smem_desc(smem_u32(sV + slot * kTileBytes), kHalfTileBytes, 1024, UMMA_LAYOUT_SW128);
        const uint32_t p_col = tmem_base + kColS0 + (j & 1) * 128;
#pragma unroll
        for (int kk = 0; kk < kBlock / 16; ++kk) {
          const uint64_t off = static_cast<uint64_t>((kk * 16 * 128) >> 4);
          // keys of k-step kk belong to column slice kk*16/kCols; its P sits at the start of
          // that slice's own S columns
          const uint32_t a = p_col + ((kk * 16) / kCols) * kCols + (((kk * 16) % kCols) >> 1);
          umma_ts(tmem_base + kColO, a, v_desc + off, idesc_pv, (j > 0 || kk > 0) ? 1u : 0u);
        }
        umma_commit(&bars[V_EMPTY0 + slot]);
        umma_commit(&bars[PV_DONE]);
      };
      mbar_wait(&bars[Q_READY], 0, p.err_flag);
      tc_fence_after();
      issue_qk(0);
      if (n_tiles > 1) issue_qk(1);
      for (int j = 0; j < n_tiles; ++j) {
        issue_pv(j);
        if (j + 2 < n_tiles) issue_qk(j + 2);  // overwrites S[j&1] after PV(j): in-order pipe
      }
    }
  } else {
    // =============================== softmax / epilogue ===============================
    const int quad = warp & 3;             // TMEM lane quadrant of this warp
    const int ch = (warp - 2) >> 2;        // column slice: keys [ch*kCols, (ch+1)*kCols)
    const int row = quad * 32 + lane;
    const uint32_t lane_base = static_cast<uint32_t>(quad * 32) << 16;
    const long long q_row = q_row0 + row;
    const float c = dense ? p.qk_scale : 1.0f;

    float m_used = -INFINITY, l_sum = 0.f;
    float run0 = -INFINITY, run1 = -INFINITY, run2 = -INFINITY;  // my running max: <=j, <=j-1, <=j-2
    float m_first = -INFINITY;

    if (n_tiles > 0) {
      // ---- Q~ -> TMEM (ref :87-88: q = (q * qk_scale).to(dtype) for the sparse class) ----
      mbar_wait(&bars[Q_FULL], 0, p.err_flag);
      {
        // this thread moves kCols/2 packed columns (= kCols d-elements) of its row
        constexpr int kChunks = kCols / 8;  // 16-byte chunks
        uint32_t qreg[kCols / 2];
        const int d0 = ch * kCols;          // first d element
        const uint8_t* qsrc = sQ + (d0 >> 6) * kHalfTileBytes + row * 128;
        const int chunk0 = (d0 & 63) >> 3;
#pragma unroll
        for (int cchunk = 0; cchunk < kChunks; ++cchunk) {
          const uint4 v = *reinterpret_cast<const uint4*>(qsrc + (((chunk0 + cchunk) ^ (row & 7)) << 4));
          uint32_t e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            if (!dense) {
              const float2 f = unpack2<kBF16>(e[t]);
              e[t] = pack2<kBF16>(f.x * p.qk_scale, f.y * p.qk_scale);
            }
            qreg[cchunk * 4 + t] = e[t];
          }
        }
        if constexpr (kSplit == 2) {
          tmem_st32(tmem_base + kColQ + lane_base + ch * 32, qreg);
        } else {
          tmem_st16(tmem_base + kColQ + lane_base + ch * 16, reinterpret_cast<const uint32_t(&)[16]>(qreg));
        }
        tmem_st_wait();
      }
      tc_fence_before();
      mbar_arrive(&bars[Q_READY]);

      // (Issuing the TMEM load of S(j+1) before the exponential phase of tile j was tried and was
      // slower — 785 vs 1050 TF/s at HY-720p; the loop below is the straight version.)
      auto load_scores = [&](float (&dst)[kCols], int j) {
        const int buf = j & 1;
        mbar_wait(&bars[S_FULL0 + buf], (j >> 1) & 1, p.err_flag);
        tc_fence_after();
        uint32_t* su = reinterpret_cast<uint32_t*>(dst);
        const uint32_t t = tmem_base + kColS0 + buf * 128 + ch * kCols + lane_base;
#pragma unroll
        for (int cc = 0; cc < kCols; cc += 32) tmem_ld32(t + cc, su + cc);
      };
      auto process = [&](float (&s)[kCols], int j, int blk) {
        const int buf = j & 1;
        const uint32_t tmem_S = tmem_base + kColS0 + buf * 128 + ch * kCols + lane_base;
        const float amp = (!dense && blk >= p.text_block_start) ? p.text_amp : 0.f;
        // partner maxima up to tile j-2: issue the shared-memory reads before the TMEM wait
        float other = -INFINITY;
        if (j >= 2) {
#pragma unroll
          for (int o = 1; o < kSplit; ++o)
            other = fmaxf(other, s_pub[(((j - 2) & 3) * kSplit + ((ch + o) % kSplit)) * 128 + row]);
        }
        load_scores(s, j);
        tmem_ld_wait();  // S(j) has landed in s[]
        const long long col0 = static_cast<long long>(blk) * kBlock + ch * kCols;
        if (col0 + kCols > kv_limit) {
          asm volatile("" ::: "memory");
#pragma unroll
          for (int i = 0; i < kCols; ++i)
            if (col0 + i >= kv_limit) s[i] = -INFINITY;  // ref :117-118
        }
        float mx0 = fmaxf(s[0], s[1]), mx1 = fmaxf(s[2], s[3]);
#pragma unroll
        for (int i = 4; i < kCols; i += 4) {
          mx0 = fmaxf(mx0, fmaxf(s[i], s[i + 1]));
          mx1 = fmaxf(mx1, fmaxf(s[i + 2], s[i + 3]));
        }
        const float m_loc = fmaf(fmaxf(mx0, mx1), c, amp);
        run2 = run1;
        run1 = run0;
        run0 = fmaxf(run0, m_loc);
        s_pub[((j & 3) * kSplit + ch) * 128 + row] = run0;
        // reference point for this tile: every column slice's maximum up to tile j-2
        float m_both;
        if (j == 0) {
          named_bar_sync(1 + quad, 32 * kSplit);  // the only rendezvous: first tile needs the true max
          m_first = run0;
#pragma unroll
          for (int o = 1; o < kSplit; ++o)
            m_first = fmaxf(m_first, s_pub[(0 * kSplit + ((ch + o) % kSplit)) * 128 + row]);
          m_both = m_first;
        } else if (j == 1) {
          m_both = m_first;
        } else {
          m_both = fmaxf(run2, other);
        }
        const float m_cand = fmaxf(m_used, m_both);
        const bool need = (m_cand - m_used) > 8.0f;
        if (__any_sync(0xffffffffu, need)) {  // same rows in the partner warps -> same decision
          if (j > 0) {
            mbar_wait(&bars[PV_DONE], (j - 1) & 1, p.err_flag);  // every PV so far has retired
            tc_fence_after();
            const float alpha = (m_cand == -INFINITY) ? 1.0f : fast_exp2(m_used - m_cand);
#pragma unroll 1
            for (int cc = 0; cc < kOCols; cc += 16) {
              uint32_t o[16];
              tmem_ld16(tmem_base + kColO + lane_base + ch * kOCols + cc, o);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
              tmem_st16(tmem_base + kColO + lane_base + ch * kOCols + cc, o);
            }
            tmem_st_wait();
            l_sum *= alpha;
          }
          m_used = m_cand;
        }
        if (m_loc - m_used > 100.f && p.err_flag) atomicExch(p.err_flag, JENGA_DEV_RANGE);
        const float off = amp - ((m_used == -INFINITY) ? 0.f : m_used);
        const f32x2 c2 = f2_pack(c, c), off2 = f2_pack(off, off);
        f32x2 sum2 = f2_pack(0.f, 0.f);
#pragma unroll
        for (int cc = 0; cc < kCols; cc += 32) {
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
            sum2 = f2_add(sum2, pp);
            pk[i] = pack2<kBF16>(p0, p1);
          }
          tmem_st16(tmem_S + (cc >> 1), pk);  // P over the start of my own S columns
        }
        float sum0, sum1;
        f2_unpack(sum2, sum0, sum1);
        l_sum += sum0 + sum1;
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&bars[P_FULL0 + buf]);
      };

      BlockWalker it(s_mask, nwords);
      float sA[kCols];
      int j = 0;
      for (int blk = it.next(); blk >= 0; blk = it.next(), ++j) process(sA, j, blk);
      mbar_wait(&bars[PV_DONE], (n_tiles - 1) & 1, p.err_flag);
      tc_fence_after();
      // full row sum = sum of the two column halves
      s_lsum[ch * 128 + row] = l_sum;
      named_bar_sync(1 + quad, 32 * kSplit);
#pragma unroll
      for (int o = 1; o < kSplit; ++o) l_sum += s_lsum[((ch + o) % kSplit) * 128 + row];
    }

    // ---- epilogue: my kOCols columns of O / l ----
    const bool in_tensor = q_row < p.q_rows;
    const bool zero_row = (n_tiles == 0) || (!dense && q_row >= q_limit_sparse);
    const float inv_l = zero_row ? 0.f : 1.0f / l_sum;
    const long long o_off = b * p.o_stride_b + q_row * p.o_stride_s +
                            static_cast<long long>(h) * p.o_stride_h + ch * kOCols;
    uint16_t* orow = reinterpret_cast<uint16_t*>(p.out) + o_off;
    float* orow32 = reinterpret_cast<float*>(p.out) + o_off;
#pragma unroll 1
    for (int cc = 0; cc < kOCols; cc += 32) {
      uint32_t o[32];
      if (n_tiles > 0) {
        tmem_ld32(tmem_base + kColO + lane_base + ch * kOCols + cc, o);
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
            const float a = zero_row ? 0.f : __uint_as_float(o[i + 2 * t]) * inv_l;
            const float bb = zero_row ? 0.f : __uint_as_float(o[i + 2 * t + 1]) * inv_l;
            e[t] = pack2<kBF16>(a, bb);
          }
          if (!p.out_f32) {
            *reinterpret_cast<uint4*>(orow + cc + i) = v;
          } else {
            float2 t0 = unpack2<kBF16>(e[0]), t1 = unpack2<kBF16>(e[1]);
            const float4 f0 = make_float4(t0.x, t0.y, t1.x, t1.y);
            t0 = unpack2<kBF16>(e[2]);
            t1 = unpack2<kBF16>(e[3]);
            const float4 f1 = make_float4(t0.x, t0.y, t1.x, t1.y);
            *reinterpret_cast<float4*>(orow32 + cc + i) = f0;
            *reinterpret_cast<float4*>(orow32 + cc + i + 4) = f1;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<kTmemCols>(tmem_base);
}

}  // namespace

int launch_carved_attn_v3(const CUtensorMap& tm_q, const CUtensorMap& tm_k, const CUtensorMap& tm_v,
                          const attn::KernelParams& p, unsigned grid, bool bf16, int split, cudaStream_t stream) {
  auto kern = split == 4 ? (bf16 ? carved_attn_v3_kernel<true, 4> : carved_attn_v3_kernel<false, 4>)
                         : (bf16 ? carved_attn_v3_kernel<true, 2> : carved_attn_v3_kernel<false, 2>);
  const int threads = 64 + 128 * (split == 4 ? 4 : 2);
  cudaError_t ce = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
  if (ce != cudaSuccess) return set_cuda_error(ce, "cudaFuncSetAttribute(carved_attn_v3)");
  kern<<<grid, threads, kSmemBytes, stream>>>(tm_q, tm_k, tm_v, p);
  ce = cudaGetLastError();
  if (ce != cudaSuccess) return set_cuda_error(ce, "carved_attn_v3 launch");
  return JENGA_OK;
}

}  // namespace jenga
