// Adaptive graph conv with the channels on the TMEM lanes (k_tc_agcnt): AdaptiveGraphConvolution.forward with T = 1,
// models/a_gcn/a_gcn.py:52-69, the dense-mix half (the mixing rows come from k_tc_attn / k_attn_small):
//
//   z[:, w] = sum_i sum_v M_i[v, w] * (W_i x[:, v]) + R x[:, w] + bias ,  ReLU ,   M_i = softmax_i + A_i + graph_attn_i  per skeleton and frame
//
// Same orientation as k_tc_gcnt (tc_gcnt.cuh): Y_i^T = W_i X^T, 128 output channels on the accumulator lanes, the tokens of a
// tile along the columns, so a thread of the epilogue owns one output channel and mixes vertices in registers.  The mixing
// matrix is dense and differs per skeleton, but it is the SAME for every channel: the 3 * V * V coefficients of a skeleton are
// read from shared memory with 16-byte broadcast loads (one load feeds four FMAs of every lane) instead of sitting in
// per-lane registers, and the operand Y_i[v] of all those FMAs is a register.  One FMA per (partition, v, w, channel) and a
// quarter of a shared-memory instruction -- against one FMA per loaded shared-memory word in the token-major k_tc_agcn.
//
// STATUS: compiles for sm_100a without spills; NOT yet run on hardware (the round's GPU budget was spent before its first launch).
// It is therefore not part of the default build: COSK_WITH_AGCNT=1 (lib.py -> -DCOSK_WITH_AGCNT) compiles it in, COSK_AGCN_T=1
// selects it at run time, and tests/test_gpu_parity.py::test_channel_major_adaptive_graph_conv (COSK_TEST_UNVERIFIED=1) checks it
// against the step oracle and the token-major kernel.
//
// Work item = (frame, token tile): the tile's mixing rows are one contiguous block of the attention kernel's scratch
// ([token][3][VP] floats) and arrive with a single bulk copy (double buffered); the X tile stays resident for the item
// (N = 128 columns per MMA), weight slabs stream through a two-stage ring; four accumulators of 128 columns rotate in
// TMEM, one per (128-channel chunk, part), parts 0-2 = partitions, part 3 = gcn_residual (folded conv or identity).
//   warp 0  TMA producer (weights)   warp 2  TMA producer (X, mixing rows)   warp 1  MMA issuer + TMEM owner
//   warps 4-11  epilogue: warp % 4 = TMEM sub-partition (32 channels), warps 4-7 take the first skeletons of the tile,
//               warps 8-11 the rest, so two warps per scheduler hide each other's shared-memory latency
#pragma once
#include "tc_gcnt.cuh"

namespace cosk {

enum : unsigned int {
  kDbgAgcntCEmpty = 0x39000000u,
  kDbgAgcntCFull = 0x3a000000u,
};

namespace ptx {
// 1D bulk copy global -> shared, completion counted in bytes on `bar` (size and addresses multiples of 16)
__device__ __forceinline__ void bulk_load_1d(uint32_t smem_dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_dst),
               "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
}  // namespace ptx

template <int V, int NKB>
struct TcAgcntCfg {
  static constexpr int kSkel = kTileRows / V;
  static constexpr int kVP = (V + 3) / 4 * 4;            // == dense_vp
  static constexpr int kLd = 3 * kVP;                    // floats per token row of the scratch
  static constexpr int kCoefBytes = kSkel * V * kLd * 4; // one tile's mixing rows (42000 B for V = 25)
  static constexpr int kCoefStride = (kCoefBytes + 1023) / 1024 * 1024;
  static constexpr int kXBytes = NKB * 2 * kABytes;      // per K-block: hi plane, lo plane
  static constexpr int kWStageBytes = 2 * kABytes;
  static constexpr int kWStages = 2;
  static constexpr int kXOff = 0;
  static constexpr int kWOff = kXBytes;
  static constexpr int kCoefOff = kWOff + kWStages * kWStageBytes;
  static constexpr int kBarOff = kCoefOff + 2 * kCoefStride;
  static constexpr int kSmemBytes = kBarOff + 256 + 1024;
  static constexpr int kAccCols = 128, kAccBufs = 4;
  static constexpr int kTmemCols = 512;
  static constexpr int kSkelA = (kSkel + 1) / 2;  // skeletons of the first epilogue group
  static_assert(kCoefBytes % 16 == 0, "bulk copy granularity");
  static_assert(kSmemBytes <= kSmemLimit, "shared memory budget");
};

template <int V, int NKB, int COUT>
__global__ void __launch_bounds__(384, 1) k_tc_agcnt(const __grid_constant__ TcGcntArgs ta) {
  using Cfg = TcAgcntCfg<V, NKB>;
  constexpr int S = Cfg::kSkel, VP = Cfg::kVP, LD = Cfg::kLd;
  constexpr int kChunks = COUT / 128;
  constexpr int P = 4;
  const TcGcnArgs &a = ta.g;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t *xfull = reinterpret_cast<uint64_t *>(smem + Cfg::kBarOff);  // [NKB <= 2]
  uint64_t *xempty = xfull + 2;
  uint64_t *wfull = xempty + 2;
  uint64_t *wempty = wfull + Cfg::kWStages;
  uint64_t *tfull = wempty + Cfg::kWStages;
  uint64_t *tempty = tfull + Cfg::kAccBufs;
  uint64_t *cfull = tempty + Cfg::kAccBufs;
  uint64_t *cempty = cfull + 2;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(cempty + 2);
  const uint32_t smem_base = ptx::smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cta = (int)blockIdx.x, ncta = (int)gridDim.x;

  pdl_trigger();
  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&xfull[s], 1);
      ptx::mbar_init(&xempty[s], 1);
      ptx::mbar_init(&cfull[s], 1);
      ptx::mbar_init(&cempty[s], 8);  // the eight epilogue warps
    }
    for (int s = 0; s < Cfg::kWStages; ++s) {
      ptx::mbar_init(&wfull[s], 1);
      ptx::mbar_init(&wempty[s], 1);
    }
    for (int s = 0; s < Cfg::kAccBufs; ++s) {
      ptx::mbar_init(&tfull[s], 1);
      ptx::mbar_init(&tempty[s], 8);
    }
    ptx::fence_barrier_init();
    ptx::prefetch_tmap(&a.tm_x);
    ptx::prefetch_tmap(&a.tm_w);
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, Cfg::kTmemCols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // prologue above touched only static data; activations and mixing rows of the previous kernels from here on

  constexpr int n_pass = kChunks * P;
  const int n_items = a.n_tiles * a.n_frames;

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");  // register trade as in k_tc_gcnt: 128 * 56 + 256 * 224 = 384 * 168
  if (warp == 0) {
    if (lane == 0) {
      // ---- TMA producer, weights: one slab per (item, chunk, part, K-block) --------------------------------------------
      PipeState ws;
      bool ok = true;
      for (int vt = cta; ok && vt < n_items; vt += ncta)
        for (int pass = 0; ok && pass < n_pass; ++pass)
          for (int kb = 0; kb < NKB; ++kb) {
            ok = ptx::mbar_wait(&wempty[ws.stage], ws.phase ^ 1, a.dbg, kDbgGcntWEmpty | (unsigned)(pass * 16 + kb));
            if (!ok) break;
            const uint32_t wst = smem_base + Cfg::kWOff + ws.stage * Cfg::kWStageBytes;
            ptx::mbar_arrive_expect_tx(&wfull[ws.stage], Cfg::kWStageBytes);
            ptx::tma_load_2d_hint(wst, &a.tm_w, &wfull[ws.stage], kb * kBK, pass * 128, ptx::kEvictLast);
            ptx::tma_load_2d_hint(wst + kABytes, &a.tm_w, &wfull[ws.stage], kb * kBK, P * COUT + pass * 128, ptx::kEvictLast);
            ws.advance<Cfg::kWStages>();
          }
    }
  } else if (warp == 2) {
    if (lane == 0) {
      // ---- TMA producer, per item: the tile's mixing rows (one bulk copy) and the K-blocks of the X tile -------------
      bool ok = true;
      uint32_t it = 0;
      for (int vt = cta; ok && vt < n_items; vt += ncta, ++it) {
        const int fr = vt / a.n_tiles, tile = vt - fr * a.n_tiles;
        const int row = gcn_x_row(a, fr) + tile * a.tile_tokens;
        const uint32_t cb = it & 1;
        ok = ptx::mbar_wait(&cempty[cb], ((it >> 1) & 1) ^ 1, a.dbg, kDbgAgcntCEmpty | (it & 0xffff));
        if (!ok) break;
        ptx::mbar_arrive_expect_tx(&cfull[cb], Cfg::kCoefBytes);
        ptx::bulk_load_1d(smem_base + Cfg::kCoefOff + cb * Cfg::kCoefStride, a.dense + (long long)tile * a.tile_tokens * LD,
                          Cfg::kCoefBytes, &cfull[cb]);
        for (int kb = 0; kb < NKB; ++kb) {
          ok = ptx::mbar_wait(&xempty[kb], (it & 1) ^ 1, a.dbg, kDbgGcntXEmpty | (it & 0xffff));
          if (!ok) break;
          const uint32_t xs = smem_base + Cfg::kXOff + kb * 2 * kABytes;
          ptx::mbar_arrive_expect_tx(&xfull[kb], 2 * kABytes);
          ptx::tma_load_2d_hint(xs, &a.tm_x, &xfull[kb], kb * kBK, row, ptx::kEvictFirst);
          ptx::tma_load_2d_hint(xs + kABytes, &a.tm_x, &xfull[kb], kb * kBK, row + a.t_alloc, ptx::kEvictFirst);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ---- MMA issuer: D[128 channels x 128 tokens] (+)= W_part[128 x 64] * X[128 x 64]^T, three split products -------
      constexpr uint32_t idesc = ptx::umma_idesc_bf16(128, 128);
      PipeState ws;
      uint32_t pidx = 0, it = 0;
      bool ok = true;
      for (int vt = cta; ok && vt < n_items; vt += ncta, ++it)
        for (int pass = 0; ok && pass < n_pass; ++pass, ++pidx) {
          const uint32_t buf = pidx % Cfg::kAccBufs, use = pidx / Cfg::kAccBufs;
          ok = ptx::mbar_wait(&tempty[buf], (use & 1) ^ 1, a.dbg, kDbgGcntTEmpty | (pidx & 0xffff));
          if (!ok) break;
          const uint32_t d = tmem_base + buf * Cfg::kAccCols;
          for (int kb = 0; kb < NKB; ++kb) {
            ok = ptx::mbar_wait(&xfull[kb], it & 1, a.dbg, kDbgGcntXFull | (it & 0xffff));
            if (!ok) break;
            ok = ptx::mbar_wait(&wfull[ws.stage], ws.phase, a.dbg, kDbgGcntWFull | (unsigned)(pass * 16 + kb));
            if (!ok) break;
            ptx::tc_fence_after();
            const uint32_t xs = smem_base + Cfg::kXOff + kb * 2 * kABytes;
            const uint32_t wst = smem_base + Cfg::kWOff + ws.stage * Cfg::kWStageBytes;
            const uint32_t wh = ptx::umma_desc_lo(wst), wl = ptx::umma_desc_lo(wst + kABytes);
            const uint32_t xh = ptx::umma_desc_lo(xs), xl = ptx::umma_desc_lo(xs + kABytes);
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k) {
              ptx::umma_bf16_lo(d, wh + 2 * k, xh + 2 * k, idesc, (kb == 0 && k == 0) ? 0u : 1u);
              ptx::umma_bf16_lo(d, wl + 2 * k, xh + 2 * k, idesc, 1u);
              ptx::umma_bf16_lo(d, wh + 2 * k, xl + 2 * k, idesc, 1u);
            }
            ptx::umma_commit(&wempty[ws.stage]);
            if (pass == n_pass - 1) ptx::umma_commit(&xempty[kb]);  // the item's last reader of this K-block
            ws.advance<Cfg::kWStages>();
          }
          if (ok) ptx::umma_commit(&tfull[buf]);
        }
    }
  }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    // ---- epilogue: thread = output channel; group 0 mixes skeletons [0, kSkelA), group 1 the rest -----------------------
    const int grp = (warp - 4) >> 2, q = warp & 3;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    const bool odd = (lane & 1) != 0;
    uint32_t pidx = 0, it = 0;
    bool ok = true;
    auto run = [&](auto grp_) {
      constexpr int G = decltype(grp_)::value;
      constexpr int S0 = G == 0 ? 0 : Cfg::kSkelA, S1 = G == 0 ? Cfg::kSkelA : S;  // skeletons [S0, S1) of the tile
      constexpr int NS = S1 - S0, NZ = NS * V, T0 = S0 * V;                        // tokens [T0, T0 + NZ) of the tile
      for (int vt = cta; ok && vt < n_items; vt += ncta, ++it) {
        const int fr = vt / a.n_tiles, tile = vt - fr * a.n_tiles;
        const long long tok0 = (long long)tile * a.tile_tokens + T0;
        const EpiArgs epi = gcn_epi(a, fr);
        const uint32_t cb = it & 1;
        ok = ptx::mbar_wait(&cfull[cb], (it >> 1) & 1, a.dbg, kDbgAgcntCFull | (it & 0xffff));
        if (!ok) break;
        const float *coef = reinterpret_cast<const float *>(smem + Cfg::kCoefOff + cb * Cfg::kCoefStride);
        for (int chunk = 0; ok && chunk < kChunks; ++chunk) {
          const int ch = chunk * 128 + q * 32 + lane;
          const float bias = __ldg(epi.bias + ch);
          float z[NZ];
          for (int part = 0; part < P; ++part, ++pidx) {
            const uint32_t buf = pidx % Cfg::kAccBufs, use = pidx / Cfg::kAccBufs;
            ok = ptx::mbar_wait(&tfull[buf], use & 1, a.dbg, kDbgGcntTFull | (pidx & 0xffff));
            if (!ok) break;
            ptx::tc_fence_after();
            const uint32_t taddr = lane_base + buf * Cfg::kAccCols;
            static_for<NS>([&](auto s_) {
              constexpr int sl = decltype(s_)::value, s = S0 + sl;
              constexpr int col0 = s * V < kTileRows - 32 ? s * V : kTileRows - 32;
              constexpr int off = s * V - col0;
              uint32_t y[32];
              ptx::tmem_ld_32x32(taddr + col0, y);
              ptx::tmem_ld_wait();
              if (part < 3) {
                // z[w] (+)= sum_v M_part[v][w] * Y_part[v]: the row of output vertex w holds its V coefficients contiguously
                const float *cs = coef + (s * V) * LD + part * VP;
                static_for<V>([&](auto w_) {
                  constexpr int w = decltype(w_)::value;
                  float acc = part == 0 ? bias : z[sl * V + w];
                  const float4 *c4 = reinterpret_cast<const float4 *>(cs + w * LD);
                  static_for<VP / 4>([&](auto j_) {
                    constexpr int j = decltype(j_)::value;
                    const float4 c = c4[j];
                    acc = fmaf(c.x, __uint_as_float(y[off + 4 * j]), acc);
                    if constexpr (4 * j + 1 < V) acc = fmaf(c.y, __uint_as_float(y[off + 4 * j + 1]), acc);
                    if constexpr (4 * j + 2 < V) acc = fmaf(c.z, __uint_as_float(y[off + 4 * j + 2]), acc);
                    if constexpr (4 * j + 3 < V) acc = fmaf(c.w, __uint_as_float(y[off + 4 * j + 3]), acc);
                  });
                  z[sl * V + w] = acc;
                });
              } else {
                static_for<V>([&](auto w_) {
                  constexpr int w = decltype(w_)::value;
                  z[sl * V + w] += __uint_as_float(y[off + w]);
                });
              }
            });
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&tempty[buf]);  // this warp's columns are read: (with the other seven) the MMA may overwrite
          }
          if (!ok) break;
          // ReLU, split, store: lane pairs trade tokens so that every store carries two channels (as k_tc_gcnt)
          long long left = tile < a.n_tiles ? a.n_tokens - tok0 : 0;
          const int n_valid = left > NZ ? NZ : (left < 0 ? 0 : (int)left);
          const uint32_t sel_keep = odd ? 0x7632u : 0x5410u, sel_send = odd ? 0x5410u : 0x7632u;
          const uint32_t sel_hi = odd ? 0x1054u : 0x5410u, sel_lo = odd ? 0x3276u : 0x7632u;
          __nv_bfloat16 *ph = epi.y_hi + (tok0 + (odd ? 1 : 0)) * COUT + (ch & ~1);
          __nv_bfloat16 *pl = epi.y_lo + (tok0 + (odd ? 1 : 0)) * COUT + (ch & ~1);
          static_for<NZ / 2>([&](auto t_) {
            constexpr int t = 2 * decltype(t_)::value;
            const float x0 = fmaxf(z[t], epi.floor), x1 = fmaxf(z[t + 1], epi.floor);
            const uint32_t h2 = pack_bf16x2(x0, x1);
            const uint32_t l2 = pack_bf16x2(x0 - bf16_lo_as_float(h2), x1 - bf16_hi_as_float(h2));
            const uint32_t keep = __byte_perm(h2, l2, sel_keep);
            const uint32_t recv = __shfl_xor_sync(0xffffffffu, __byte_perm(h2, l2, sel_send), 1);
            if (t + (odd ? 1 : 0) < n_valid) {
              *reinterpret_cast<uint32_t *>(ph + t * COUT) = __byte_perm(keep, recv, sel_hi);
              *reinterpret_cast<uint32_t *>(pl + t * COUT) = __byte_perm(keep, recv, sel_lo);
            }
          });
          if constexpr (NZ % 2 == 1) {
            if (NZ - 1 < n_valid) {
              const float x = fmaxf(z[NZ - 1], epi.floor);
              const __nv_bfloat16 h = __float2bfloat16_rn(x);
              st_bf16(epi.y_hi + (tok0 + NZ - 1) * COUT + ch, h);
              st_bf16(epi.y_lo + (tok0 + NZ - 1) * COUT + ch, __float2bfloat16_rn(x - __bfloat162float(h)));
            }
          }
        }
        if (!ok) break;
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&cempty[cb]);  // mixing rows consumed
      }
    };
    if (grp == 0) run(std::integral_constant<int, 0>{});
    else run(std::integral_constant<int, 1>{});
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

}  // namespace cosk
