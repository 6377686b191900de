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
    const int n0 = st[0].n_tiles, n1 = st[1].n_tiles;
    if ((n0 > 0 || n1 > 0) && elect_one()) {
      constexpr uint32_t idesc_qk = umma_idesc_f16(kBF16, false, 128, kHalf);
      constexpr uint32_t idesc_pv = umma_idesc_f16(kBF16, true, 128, kHeadDim);
      // stream-indexed state as arithmetic on `s`, so that nothing is a dynamically indexed array
      const uint32_t smem0 = smem_u32(smem);
      auto q_desc = [&](int s) { return umma_smem_desc(smem0 + s * kStreamBytes + kOffQ, 16, 1024, UMMA_LAYOUT_SW128); };
      auto k_desc = [&](int s, int hh) {
        return umma_smem_desc(smem0 + s * kStreamBytes + kOffK + hh * kKVBoxBytes, 16, 1024, UMMA_LAYOUT_SW128);
      };
      auto v_desc = [&](int s, int hh) {
        return umma_smem_desc(smem0 + s * kStreamBytes + kOffV + hh * kKVSlotBytes, kKVBoxBytes, 1024, UMMA_LAYOUT_SW128);
      };
      auto bar = [&](int s, int id) { return bars_all + s * 16 + id; };
      auto tmem_of = [&](int s) { return tmem_base + 256u * static_cast<uint32_t>(s); };
      auto q_mma = [&](uint64_t qd, uint64_t kd, uint32_t d_tmem, int kk) {
        const uint64_t qoff = static_cast<uint64_t>(((kk & 3) * 32 + (kk >> 2) * kQHalfBytes) >> 4);
        const uint64_t koff = static_cast<uint64_t>(((kk & 3) * 32 + (kk >> 2) * 2 * kKVBoxBytes) >> 4);
        umma_ss(d_tmem, qd + qoff, kd + koff, idesc_qk, kk > 0 ? 1u : 0u);
      };
      auto p_mma = [&](uint64_t vd, uint32_t o_tmem, uint32_t p_tmem, int kk, bool first_of_stream) {
        const uint64_t off = static_cast<uint64_t>((kk * 16 * 128) >> 4);
        umma_ts(o_tmem, p_tmem + kk * 8, vd + off, idesc_pv, (!first_of_stream || kk > 0) ? 1u : 0u);
      };

      int u[2] = {0, 0};
      const int total0 = n0 > 0 ? 2 + 4 * n0 : 0, total1 = n1 > 0 ? 2 + 4 * n1 : 0;
      bool done0 = total0 == 0, done1 = total1 == 0;
      if (!done0) mbar_wait(bar(0, Q_READY), 0, p.err_flag);
      if (!done1) mbar_wait(bar(1, Q_READY), 0, p.err_flag);
      tc_fence_after();
      int u0 = 0, u1 = 0;
      (void)u;
      auto ready = [&](int s, const Unit& un) -> bool {
        const uint32_t par = un.tile & 1;
        if (un.is_pv) return mbar_test(bar(s, V_FULL0 + un.hh), par) && mbar_test(bar(s, P_FULL0 + un.hh), par);
        return mbar_test(bar(s, K_FULL0 + un.hh), par);
      };
      auto issue_alone = [&](int s, const Unit& un) {
        const uint32_t tm = tmem_of(s);
        if (un.is_pv) {
          const uint64_t vd = v_desc(s, un.hh);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) p_mma(vd, tm + 128, tm + un.hh * kHalf, kk, un.tile == 0 && un.hh == 0);
          umma_commit(bar(s, V_EMPTY0 + un.hh));
        } else {
          const uint64_t qd = q_desc(s), kd = k_desc(s, un.hh);
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) q_mma(qd, kd, tm + un.hh * kHalf, kk);
          umma_commit(bar(s, K_EMPTY0 + un.hh));
          umma_commit(bar(s, S_FULL0 + un.hh));
        }
      };
      long long spin_t0 = 0;
      int spins = 0;
      for (;;) {
        // skip the Q.K^T units that belong to the non-existent tile n_tiles; detect the end
        while (!done0) {
          if (u0 >= total0) { done0 = true; break; }
          const Unit un = decode(u0);
          if (!un.is_pv && un.tile >= n0) { ++u0; continue; }
          break;
        }
        while (!done1) {
          if (u1 >= total1) { done1 = true; break; }
          const Unit un = decode(u1);
          if (!un.is_pv && un.tile >= n1) { ++u1; continue; }
          break;
        }
        if (done0 && done1) break;
        const Unit a0 = decode(u0), a1 = decode(u1);
        const bool r0 = !done0 && ready(0, a0);
        const bool r1 = !done1 && ready(1, a1);
        if (r0 && r1 && a0.is_pv != a1.is_pv) {
          // the point of this kernel:  P Q Q P Q Q P Q Q P Q Q  across the two streams
          tc_fence_after();
          const int sp = a0.is_pv ? 0 : 1, sq = sp ^ 1;
          const Unit up = a0.is_pv ? a0 : a1, uq = a0.is_pv ? a1 : a0;
          const bool first = up.tile == 0 && up.hh == 0;
          const uint32_t tp = tmem_of(sp), tq = tmem_of(sq);
          const uint64_t vd = v_desc(sp, up.hh), qd = q_desc(sq), kd = k_desc(sq, uq.hh);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            p_mma(vd, tp + 128, tp + up.hh * kHalf, i, first);
            q_mma(qd, kd, tq + uq.hh * kHalf, 2 * i);
            q_mma(qd, kd, tq + uq.hh * kHalf, 2 * i + 1);
          }
          umma_commit(bar(sp, V_EMPTY0 + up.hh));
          umma_commit(bar(sq, K_EMPTY0 + uq.hh));
          umma_commit(bar(sq, S_FULL0 + uq.hh));
          ++u0;
          ++u1;
          spins = 0;
        } else if (r0 || r1) {
          // one stream only (or both at the same kind: advance the one that is behind, which puts
          // the two streams in anti-phase for the steps that follow)
          tc_fence_after();
          const int s = (r0 && r1) ? (u0 <= u1 ? 0 : 1) : (r0 ? 0 : 1);
          issue_alone(s, s == 0 ? a0 : a1);
          if (s == 0) ++u0; else ++u1;
          spins = 0;
        } else {
          __nanosleep(40);   // nothing ready: do not steal issue slots from the softmax warps of this scheduler
          if (++spins == 1) spin_t0 = clock64();
          if ((spins & 1023) == 0 && (clock64() - spin_t0) > (1ll << 31)) {
            if (p.err_flag) atomicExch(p.err_flag, JENGA_DEV_WATCHDOG);
            __threadfence_system();
            __trap();
          }
        }
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
            float mx0 = fmaxf(s[0], s[1]), mx1 = fmaxf(s[2], s[3]);
#pragma unroll
            for (int i = 4; i < kHalf; i += 4) {
              mx0 = fmaxf(mx0, fmaxf(s[i], s[i + 1]));
              mx1 = fmaxf(mx1, fmaxf(s[i + 2], s[i + 3]));
            }
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
