// Carved attention, generation 7: TWO query blocks ("streams") per CTA, one CTA per SM, and ONE
// tcgen05 issuer that interleaves the P.V MMAs of one stream with the Q.K^T MMAs of the other.
//
// Why: generation 2 (carved_attn.cu) processes each key block as two 64-key halves so that S is
// double-buffered inside 256 TMEM columns; its Q.K^T MMAs are therefore M128 N64 K16, which read
// 4 KB of A + 2 KB of B from shared memory for 32 cycles of math — more than the 128 B/clk
// operand port, 48 cycles each (tools/umma_probe).  The probe also shows that the port limit is an
// AVERAGE over neighbouring MMAs: the issue order  P Q Q P Q Q …  (P = M128 N128 TS P.V step, 4 KB
// for 64 cycles) runs 12 MMAs in 560 cycles, against 619-640 for the bursts  PPPP QQQQQQQQ  that
// a per-stream issue order produces, and P Q P Q … runs at the ideal rate.  Within ONE q block the
// P.V and Q.K^T MMAs that are adjacent in time depend on each other (S_h(j+1) overwrites P_h(j)),
// so the interleave partner has to come from an independent q block: two gen-2 pipelines, each
// with its own Q / K_a K_b V_a V_b shared-memory slots, 256 TMEM columns, TMA producer warp and
// four softmax warps, share one issuing thread that keeps them in anti-phase.
//
// Everything a stream does (loads in consumption order, scaled-Q re-rounding, lazy rescale,
// partial polynomial exp2, masks, epilogue incl. the Ulysses peer stores) is generation 2's code
// on a per-stream context; results are bit-identical to generation 2.
#include <type_traits>

#include "carved_attn_common.cuh"

#ifndef JENGA_POLY_EVERY
#define JENGA_POLY_EVERY 3
#endif

namespace jenga {

namespace {

using attn::BlockWalker;
using attn::KernelParams;

constexpr int kBlock = 128;
constexpr int kHalf = 64;
constexpr int kHeadDim = 128;
constexpr int kThreads = 384;                    // 12 warps: 0,1 TMA | 2 MMA | 3 - | 4-7 softmax s0 | 8-11 softmax s1
constexpr int kQHalfBytes = kBlock * 64 * 2;     // 16 KB
constexpr int kQTileBytes = 2 * kQHalfBytes;     // 32 KB
constexpr int kKVBoxBytes = kHalf * 64 * 2;      // 8 KB
constexpr int kKVSlotBytes = 2 * kKVBoxBytes;    // 16 KB
constexpr int kMaxMaskWords = 256;
constexpr uint32_t kTmemCols = 512;              // stream s: [256 s, 256 s + 256): S_a S_b O
constexpr int kPolyEvery = JENGA_POLY_EVERY;

// per-stream shared memory
constexpr int kOffQ = 0;
constexpr int kOffK = kOffQ + kQTileBytes;
constexpr int kOffV = kOffK + 2 * kKVSlotBytes;
constexpr int kStreamBytes = kOffV + 2 * kKVSlotBytes;   // 96 KB
constexpr int kOffBars = 2 * kStreamBytes;               // 2 x 16 barriers
constexpr int kOffMask = kOffBars + 2 * 128 + 16;
constexpr int kSmemBytes = 1024 + kOffMask + 2 * kMaxMaskWords * 4;

enum BarId {
  Q_FULL = 0, Q_READY,
  K_FULL0, K_FULL1, K_EMPTY0, K_EMPTY1,
  V_FULL0, V_FULL1, V_EMPTY0, V_EMPTY1,
  S_FULL0, S_FULL1, P_FULL0, P_FULL1,
  NUM_BARS
};
static_assert(NUM_BARS <= 16, "barrier block overflow");

__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}

struct Stream {
  uint8_t* sQ;
  uint8_t* sK;
  uint8_t* sV;
  uint64_t* bars;
  uint32_t* mask;
  uint32_t tmem;        // base column of this stream's 256 columns
  int qb, h, b;
  bool active, dense, skip_all;
  int n_tiles;
  long long q_row0, kv_limit, q_limit_sparse;
};

// unit u of a stream's tensor-pipe program:  0: QK_a(0)  1: QK_b(0)  then for tile j = (u-2)/4:
// 2+4j: PV_a(j)   3+4j: QK_a(j+1)   4+4j: PV_b(j)   5+4j: QK_b(j+1)     (QK of tile n_tiles: none)
struct Unit {
  bool is_pv;
  int hh, tile;
};
__device__ __forceinline__ Unit decode(int u) {
  if (u < 2) return {false, u, 0};
  const int k = u - 2, j = k >> 2, r = k & 3;
  return {(r & 1) == 0, r >> 1, (r & 1) ? j + 1 : j};
}

template <bool kBF16>
__global__ void __launch_bounds__(kThreads, 1)
carved_attn_v7_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                      const __grid_constant__ CUtensorMap tm_v, const KernelParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint64_t* bars_all = reinterpret_cast<uint64_t*>(smem + kOffBars);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kOffBars + 2 * 128);
  uint32_t* mask_all = reinterpret_cast<uint32_t*>(smem + kOffMask);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // ---- which pair of q blocks?  head-major; inside a head the (long) dense blocks come first ----
  const int per_bh = p.nq_sparse + p.nq_dense;
  const int pairs = (per_bh + 1) >> 1;
  const int bh = blockIdx.x / pairs;
  const int pair = blockIdx.x - bh * pairs;
  const long long seqlen_over = p.seqlen_dev ? static_cast<long long>(__ldg(p.seqlen_dev)) : -1;
  const int nwords = p.mask_words;

  Stream st[2];
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    Stream& x = st[s];
    x.sQ = smem + s * kStreamBytes + kOffQ;
    x.sK = smem + s * kStreamBytes + kOffK;
    x.sV = smem + s * kStreamBytes + kOffV;
    x.bars = bars_all + s * 16;
    x.mask = mask_all + s * kMaxMaskWords;
    const int local = pair * 2 + s;
    x.active = local < per_bh;
    x.dense = x.active && local < p.nq_dense;
    x.qb = x.dense ? p.nq_sparse + local : local - p.nq_dense;
    x.b = bh / p.heads;
    x.h = bh - x.b * p.heads;
    x.q_row0 = static_cast<long long>(x.qb) * kBlock;
    x.q_limit_sparse = seqlen_over >= 0 ? seqlen_over : p.q_limit_sparse;
    x.kv_limit = x.dense ? p.kv_limit_dense : (seqlen_over >= 0 ? seqlen_over : p.kv_limit_sparse);
    x.skip_all = x.active && !x.dense && x.q_row0 >= x.q_limit_sparse;   // ref :59-61
  }

  // ---- stage both row masks ----
  for (int i = threadIdx.x; i < 2 * nwords; i += kThreads) {
    const int s = i / nwords, w = i - s * nwords;
    const Stream& x = st[s];
    uint32_t bits = 0;
    if (x.active && !x.skip_all) {
      const int rem = p.nb_kv - w * 32;
      const uint32_t lim = rem >= 32 ? 0xffffffffu : (rem > 0 ? ((1u << rem) - 1u) : 0u);
      bits = x.dense ? lim
                     : (p.mask_bits[(static_cast<size_t>(bh) * p.nq_sparse + x.qb) * nwords + w] & lim);
    }
    x.mask[w] = bits;
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
#pragma unroll
    for (int s = 0; s < 2; ++s)
#pragma unroll
      for (int i = 0; i < NUM_BARS; ++i)
        mbar_init(&bars_all[s * 16 + i], (i == Q_READY || i == P_FULL0 || i == P_FULL1) ? 128u : 1u);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    st[s].tmem = tmem_base + 256 * s;
    int n = 0;
    for (int w = 0; w < nwords; ++w) n += __popc(st[s].mask[w]);
    st[s].n_tiles = n;
  }

  if (warp < 2) {
    // =============================== TMA producer of stream `warp` ===============================
    const Stream& x = st[warp];
    if (x.n_tiles > 0 && elect_one()) {
      uint64_t* bars = x.bars;
      mbar_arrive_expect_tx(&bars[Q_FULL], kQTileBytes);
      tma_load_4d(x.sQ, &tm_q, &bars[Q_FULL], 0, static_cast<int>(x.q_row0), x.h, x.b);
      tma_load_4d(x.sQ + kQHalfBytes, &tm_q, &bars[Q_FULL], 64, static_cast<int>(x.q_row0), x.h, x.b);
      // consumption order: K_a(0) K_b(0) | V_a(0) K_a(1) | V_b(0) K_b(1) | …
      auto load_k = [&](int blk, int hh, uint32_t par) {
        const int row0 = blk * kBlock + hh * kHalf;
        uint8_t* dst = x.sK + hh * kKVBoxBytes;
        mbar_wait(&bars[K_EMPTY0 + hh], par, p.err_flag);
        mbar_arrive_expect_tx(&bars[K_FULL0 + hh], kKVSlotBytes);
        tma_load_4d(dst, &tm_k, &bars[K_FULL0 + hh], 0, row0, x.h, x.b);
        tma_load_4d(dst + 2 * kKVBoxBytes, &tm_k, &bars[K_FULL0 + hh], 64, row0, x.h, x.b);
      };
      auto load_v = [&](int blk, int hh, uint32_t par) {
        const int row0 = blk * kBlock + hh * kHalf;
        uint8_t* dst = x.sV + hh * kKVSlotBytes;
        mbar_wait(&bars[V_EMPTY0 + hh], par, p.err_flag);
        mbar_arrive_expect_tx(&bars[V_FULL0 + hh], kKVSlotBytes);
        tma_load_4d(dst, &tm_v, &bars[V_FULL0 + hh], 0, row0, x.h, x.b);
        tma_load_4d(dst + kKVBoxBytes, &tm_v, &bars[V_FULL0 + hh], 64, row0, x.h, x.b);
      };
      BlockWalker it(x.mask, nwords);
      int blk = it.next();
      load_k(blk, 0, 1);
      load_k(blk, 1, 1);
      for (int j = 0; blk >= 0; ++j) {
        const uint32_t par = (j & 1) ^ 1;
        const int nxt = it.next();
        load_v(blk, 0, par);
        if (nxt >= 0) load_k(nxt, 0, par ^ 1);
        load_v(blk, 1, par);
        if (nxt >= 0) load_k(nxt, 1, par ^ 1);
        blk = nxt;
      }
    }
  } else if (warp == 2) {
    // =============================== tcgen05 issuer (both streams) ===============================
    // A lone thread has to keep a 560-cycle group of 12 MMAs ahead of the tensor pipe while two busy
    // softmax warps share its scheduler, so the issue path is straight-line code: descriptors live
    // in registers, every offset is a compile-time constant, and the two streams follow a FIXED
    // anti-phase schedule (stream Y runs one unit behind stream X):
    //   per tile j:  [X.PV_a(j) | Y.QK_b(j)]  [Y.PV_a(j) | X.QK_a(j+1)]  [X.PV_b(j) | Y.QK_a(j+1)]  [Y.PV_b(j) | X.QK_b(j+1)]
    // each bracket issued as P Q Q P Q Q P Q Q P Q Q.  Inside a stream this is generation 2's order.
    const int n0 = st[0].n_tiles, n1 = st[1].n_tiles;
    if ((n0 > 0 || n1 > 0) && elect_one()) {
      constexpr uint32_t idesc_qk = umma_idesc_f16(kBF16, false, 128, kHalf);
      constexpr uint32_t idesc_pv = umma_idesc_f16(kBF16, true, 128, kHeadDim);
      struct SD {
        uint64_t qd, kd[2], vd[2];
        uint32_t tm;
        uint64_t* bars;
      };
      const uint32_t smem0 = smem_u32(smem);
      auto make = [&](int s) {
        SD d;
        d.qd = umma_smem_desc(smem0 + s * kStreamBytes + kOffQ, 16, 1024, UMMA_LAYOUT_SW128);
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          d.kd[hh] = umma_smem_desc(smem0 + s * kStreamBytes + kOffK + hh * kKVBoxBytes, 16, 1024, UMMA_LAYOUT_SW128);
          d.vd[hh] = umma_smem_desc(smem0 + s * kStreamBytes + kOffV + hh * kKVSlotBytes, kKVBoxBytes, 1024, UMMA_LAYOUT_SW128);
        }
        d.tm = tmem_base + 256u * static_cast<uint32_t>(s);
        d.bars = bars_all + s * 16;
        return d;
      };
      const SD X = make(0), Y = make(1);

      // one Q.K^T k-step / one P.V k-step, all offsets compile-time
      auto q_step = [&](const SD& d, auto HH, auto KK) {
        constexpr int hh = decltype(HH)::value, kk = decltype(KK)::value;
        constexpr uint64_t qoff = static_cast<uint64_t>(((kk & 3) * 32 + (kk >> 2) * kQHalfBytes) >> 4);
        constexpr uint64_t koff = static_cast<uint64_t>(((kk & 3) * 32 + (kk >> 2) * 2 * kKVBoxBytes) >> 4);
        umma_ss(d.tm + hh * kHalf, d.qd + qoff, d.kd[hh] + koff, idesc_qk, kk > 0 ? 1u : 0u);
      };
      auto p_step = [&](const SD& d, auto HH, auto KK, uint32_t acc0) {
        constexpr int hh = decltype(HH)::value, kk = decltype(KK)::value;
        constexpr uint64_t off = static_cast<uint64_t>((kk * 16 * 128) >> 4);
        umma_ts(d.tm + 128, d.tm + hh * kHalf + kk * 8, d.vd[hh] + off, idesc_pv, kk > 0 ? 1u : acc0);
      };
      using I0 = std::integral_constant<int, 0>;
      using I1 = std::integral_constant<int, 1>;
      using I2 = std::integral_constant<int, 2>;
      using I3 = std::integral_constant<int, 3>;
      using I4 = std::integral_constant<int, 4>;
      using I5 = std::integral_constant<int, 5>;
      using I6 = std::integral_constant<int, 6>;
      using I7 = std::integral_constant<int, 7>;
      auto qk_alone = [&](const SD& d, auto HH, int tile) {
        constexpr int hh = decltype(HH)::value;
        mbar_wait(&d.bars[K_FULL0 + hh], tile & 1, p.err_flag);
        tc_fence_after();
        q_step(d, HH, I0{}); q_step(d, HH, I1{}); q_step(d, HH, I2{}); q_step(d, HH, I3{});
        q_step(d, HH, I4{}); q_step(d, HH, I5{}); q_step(d, HH, I6{}); q_step(d, HH, I7{});
        umma_commit(&d.bars[K_EMPTY0 + hh]);
        umma_commit(&d.bars[S_FULL0 + hh]);
      };
      auto pv_alone = [&](const SD& d, auto HH, int tile) {
        constexpr int hh = decltype(HH)::value;
        mbar_wait(&d.bars[V_FULL0 + hh], tile & 1, p.err_flag);
        mbar_wait(&d.bars[P_FULL0 + hh], tile & 1, p.err_flag);
        tc_fence_after();
        const uint32_t acc0 = (tile == 0 && hh == 0) ? 0u : 1u;
        p_step(d, HH, I0{}, acc0); p_step(d, HH, I1{}, acc0); p_step(d, HH, I2{}, acc0); p_step(d, HH, I3{}, acc0);
        umma_commit(&d.bars[V_EMPTY0 + hh]);
      };
      // [P.V of half HP, tile tp, of stream dp]  interleaved with  [Q.K^T of half HQ, tile tq, of stream dq]
      auto pair = [&](const SD& dp, auto HP, int tp, bool p_valid, const SD& dq, auto HQ, int tq, bool q_valid) {
        constexpr int hp = decltype(HP)::value, hq = decltype(HQ)::value;
        if (p_valid && q_valid) {
          mbar_wait(&dq.bars[K_FULL0 + hq], tq & 1, p.err_flag);
          mbar_wait(&dp.bars[V_FULL0 + hp], tp & 1, p.err_flag);
          mbar_wait(&dp.bars[P_FULL0 + hp], tp & 1, p.err_flag);
          tc_fence_after();
          const uint32_t acc0 = (tp == 0 && hp == 0) ? 0u : 1u;
#if defined(JENGA_V7_ORDER) && JENGA_V7_ORDER == 1   // experiment: bursts instead of the interleave
          p_step(dp, HP, I0{}, acc0); p_step(dp, HP, I1{}, acc0); p_step(dp, HP, I2{}, acc0); p_step(dp, HP, I3{}, acc0);
          q_step(dq, HQ, I0{}); q_step(dq, HQ, I1{}); q_step(dq, HQ, I2{}); q_step(dq, HQ, I3{});
          q_step(dq, HQ, I4{}); q_step(dq, HQ, I5{}); q_step(dq, HQ, I6{}); q_step(dq, HQ, I7{});
#elif defined(JENGA_V7_ORDER) && JENGA_V7_ORDER == 2   // experiment: Q Q P
          q_step(dq, HQ, I0{}); q_step(dq, HQ, I1{}); p_step(dp, HP, I0{}, acc0);
          q_step(dq, HQ, I2{}); q_step(dq, HQ, I3{}); p_step(dp, HP, I1{}, acc0);
          q_step(dq, HQ, I4{}); q_step(dq, HQ, I5{}); p_step(dp, HP, I2{}, acc0);
          q_step(dq, HQ, I6{}); q_step(dq, HQ, I7{}); p_step(dp, HP, I3{}, acc0);
#else
          p_step(dp, HP, I0{}, acc0); q_step(dq, HQ, I0{}); q_step(dq, HQ, I1{});
          p_step(dp, HP, I1{}, acc0); q_step(dq, HQ, I2{}); q_step(dq, HQ, I3{});
          p_step(dp, HP, I2{}, acc0); q_step(dq, HQ, I4{}); q_step(dq, HQ, I5{});
          p_step(dp, HP, I3{}, acc0); q_step(dq, HQ, I6{}); q_step(dq, HQ, I7{});
#endif
          umma_commit(&dp.bars[V_EMPTY0 + hp]);
          umma_commit(&dq.bars[K_EMPTY0 + hq]);
          umma_commit(&dq.bars[S_FULL0 + hq]);
        } else if (p_valid) {
          pv_alone(dp, HP, tp);
        } else if (q_valid) {
          qk_alone(dq, HQ, tq);
        }
      };

      if (n0 > 0) mbar_wait(&X.bars[Q_READY], 0, p.err_flag);
      if (n1 > 0) mbar_wait(&Y.bars[Q_READY], 0, p.err_flag);
      tc_fence_after();
      if (n0 > 0) {
        qk_alone(X, I0{}, 0);
        qk_alone(X, I1{}, 0);
      }
      if (n1 > 0) qk_alone(Y, I0{}, 0);
      const int nmax = n0 > n1 ? n0 : n1;
#pragma unroll 1
      for (int j = 0; j < nmax; ++j) {
        pair(X, I0{}, j, j < n0, Y, I1{}, j, j < n1);              // X.PV_a(j)  | Y.QK_b(j)
        pair(Y, I0{}, j, j < n1, X, I0{}, j + 1, j + 1 < n0);      // Y.PV_a(j)  | X.QK_a(j+1)
        pair(X, I1{}, j, j < n0, Y, I0{}, j + 1, j + 1 < n1);      // X.PV_b(j)  | Y.QK_a(j+1)
        pair(Y, I1{}, j, j < n1, X, I1{}, j + 1, j + 1 < n0);      // Y.PV_b(j)  | X.QK_b(j+1)
      }
    }
  } else if (warp >= 4) {
    // =============================== softmax / epilogue of stream (warp-4)/4 ===============================
    const Stream& x = st[(warp - 4) >> 2];
    if (x.active) {
      uint64_t* bars = x.bars;
      const bool dense = x.dense;
      const int n_tiles = x.n_tiles;
      const int quad = warp & 3;
      const int row = quad * 32 + lane;
      const int st_idx = ((warp - 4) & 3) * 32 + lane;   // 0..127 inside the stream
      const uint32_t lane_base = static_cast<uint32_t>(quad * 32) << 16;
      const long long q_row = x.q_row0 + row;
      const uint32_t tmem_O = x.tmem + 128;

      float m_used = -INFINITY;
      float l_sum = 0.f;
      const float c = dense ? p.qk_scale : 1.0f;

      if (n_tiles > 0) {
        mbar_wait(&bars[Q_FULL], 0, p.err_flag);
        if (!dense) {
          uint4* q4 = reinterpret_cast<uint4*>(x.sQ);
#pragma unroll 4
          for (int i = 0; i < kQTileBytes / 16 / 128; ++i) {
            uint4 v = q4[st_idx + i * 128];
            uint32_t* e = reinterpret_cast<uint32_t*>(&v);
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const float2 f = unpack2<kBF16>(e[t]);
              e[t] = pack2<kBF16>(f.x * p.qk_scale, f.y * p.qk_scale);
            }
            q4[st_idx + i * 128] = v;
          }
          fence_proxy_async_smem();
        }
        mbar_arrive(&bars[Q_READY]);

        BlockWalker it(x.mask, nwords);
        int j = 0;
        for (int blk = it.next(); blk >= 0; blk = it.next(), ++j) {
          const float amp = (!dense && blk >= p.text_block_start) ? p.text_amp : 0.f;
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            const uint32_t tmem_S = x.tmem + hh * kHalf + lane_base;
            mbar_wait(&bars[S_FULL0 + hh], j & 1, p.err_flag);
            tc_fence_after();
            float s[kHalf];
            {
              uint32_t* su = reinterpret_cast<uint32_t*>(s);
              tmem_ld32(tmem_S, su);
              tmem_ld32(tmem_S + 32, su + 32);
              tmem_ld_wait();
            }
            const long long col0 = static_cast<long long>(blk) * kBlock + hh * kHalf;
            if (col0 + kHalf > x.kv_limit) {
              asm volatile("" ::: "memory");
#pragma unroll
              for (int i = 0; i < kHalf; ++i)
                if (col0 + i >= x.kv_limit) s[i] = -INFINITY;
            }
#if defined(JENGA_V7_FAKE) && (JENGA_V7_FAKE & 2)
            // TIMING EXPERIMENT ONLY: the running max is taken from the first half tile only
            float mx0 = s[0], mx1 = s[1];
            if (j == 0 && hh == 0) {
#pragma unroll
              for (int i = 4; i < kHalf; i += 4) {
                mx0 = fmaxf(mx0, fmaxf(s[i], s[i + 1]));
                mx1 = fmaxf(mx1, fmaxf(s[i + 2], s[i + 3]));
              }
            } else {
              mx0 = mx1 = -INFINITY;
            }
#else
            float mx0 = fmaxf(s[0], s[1]), mx1 = fmaxf(s[2], s[3]);
#pragma unroll
            for (int i = 4; i < kHalf; i += 4) {
              mx0 = fmaxf(mx0, fmaxf(s[i], s[i + 1]));
              mx1 = fmaxf(mx1, fmaxf(s[i + 2], s[i + 3]));
            }
#endif
            const float mx = fmaf(fmaxf(mx0, mx1), c, amp);
            const float m_cand = fmaxf(m_used, mx);
            const bool need = (m_cand - m_used) > 8.0f;
            if (__any_sync(0xffffffffu, need)) {
              if (j > 0 || hh > 0) {
                if (hh == 0) mbar_wait(&bars[V_EMPTY1], (j - 1) & 1, p.err_flag);
                else mbar_wait(&bars[V_EMPTY0], j & 1, p.err_flag);
                tc_fence_after();
                const float alpha = (m_cand == -INFINITY) ? 1.0f : fast_exp2(m_used - m_cand);
#pragma unroll 1
                for (int cc = 0; cc < 128; cc += 16) {
                  uint32_t o[16];
                  tmem_ld16(tmem_O + lane_base + cc, o);
                  tmem_ld_wait();
#pragma unroll
                  for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                  tmem_st16(tmem_O + lane_base + cc, o);
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
            for (int cc = 0; cc < kHalf; cc += 32) {
              uint32_t pk[16];
#if defined(JENGA_V7_FAKE) && (JENGA_V7_FAKE & 1)
              // TIMING EXPERIMENT ONLY (wrong results): half of the exponentials are skipped
              if (cc == 32) {
#pragma unroll
                for (int i = 0; i < 16; ++i) pk[i] = 0x3c003c00u;
                tmem_st16(tmem_S + (cc >> 1), pk);
                continue;
              }
#endif
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const f32x2 xx = f2_fma(f2_pack(s[cc + 2 * i], s[cc + 2 * i + 1]), c2, off2);
                float x0, x1, p0, p1;
                f2_unpack(xx, x0, x1);
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
              tmem_st16(tmem_S + (cc >> 1), pk);
            }
            float sum0, sum1;
            f2_unpack(sum2, sum0, sum1);
            l_sum += sum0 + sum1;
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(&bars[P_FULL0 + hh]);
          }
        }
        mbar_wait(&bars[V_EMPTY1], (n_tiles - 1) & 1, p.err_flag);
        tc_fence_after();
      }

      // ---- epilogue ----
      const int h = x.h, b = x.b;
      const bool in_tensor = q_row < p.q_rows;
      const bool zero_row = (n_tiles == 0) || (!dense && q_row >= x.q_limit_sparse);
      const float inv_l = zero_row ? 0.f : 1.0f / l_sum;
      if (p.lse_out && dense && in_tensor) {
        p.lse_out[(static_cast<long long>(b) * p.heads + h) * p.q_rows + q_row] =
            zero_row ? -INFINITY : (m_used + log2f(l_sum)) * 0.6931471805599453f;
      }
      long long o_off = b * p.o_stride_b + q_row * p.o_stride_s + static_cast<long long>(h) * p.o_stride_h;
      uint16_t* obase = reinterpret_cast<uint16_t*>(p.out);
      int n_dst = 1;
      if (p.sp_world > 0) {
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
#pragma unroll 1
      for (int cc = 0; cc < 128; cc += 32) {
        uint32_t o[32];
        if (n_tiles > 0) {
          tmem_ld32(tmem_O + lane_base + cc, o);
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
            if (n_dst > 1) {
              for (int r = 0; r < n_dst; ++r)
                *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.peer_out[r]) + o_off + cc + i) = v;
            } else if (!p.out_f32) {
              *reinterpret_cast<uint4*>(obase + o_off + cc + i) = v;
            } else {
              float* orow32 = reinterpret_cast<float*>(p.out) + o_off;
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
  }

  // ---- teardown ----
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<kTmemCols>(tmem_base);
}

}  // namespace

int launch_carved_attn_v7(const CUtensorMap& tm_q, const CUtensorMap& tm_k, const CUtensorMap& tm_v,
                          const attn::KernelParams& p, int batch_heads, bool bf16, cudaStream_t stream) {
  auto kern = bf16 ? carved_attn_v7_kernel<true> : carved_attn_v7_kernel<false>;
  cudaError_t ce = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
  if (ce != cudaSuccess) return set_cuda_error(ce, "cudaFuncSetAttribute(carved_attn_v7)");
  const long long pairs = (static_cast<long long>(p.nq_sparse) + p.nq_dense + 1) / 2;
  const long long grid = pairs * batch_heads;
  if (grid > 0x7fffffffll) return set_error(JENGA_E_UNSUPPORTED, "grid too large");
  kern<<<static_cast<unsigned>(grid), kThreads, kSmemBytes, stream>>>(tm_q, tm_k, tm_v, p);
  ce = cudaGetLastError();
  if (ce != cudaSuccess) return set_cuda_error(ce, "carved_attn_v7 launch");
  return JENGA_OK;
}

}  // namespace jenga
