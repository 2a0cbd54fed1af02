// Graph conv of a CoST-GCN block as "mix, then GEMM with the A operand in tensor memory".
//
//   z[w] = sum_i W_i (sum_v A_i[v,w] x[v]) + R x[w] + b        (GraphConvolution.forward, models/base.py:260-270)
//
// k_tc_gcn (tc_kernels.cuh) multiplies first and combines vertices afterwards, which makes the accumulator four
// column groups wide; all of them have to come back out of tensor memory (64 B/clk) and cross shared memory once
// more, and that epilogue is what bounds it.  Here the order is the reference's own: the sparse adjacency
// combination is applied to the INPUT rows by CUDA cores reading the TMA-staged tile, the mixed rows X'_i are
// re-split into bf16 hi/lo and written straight into tensor memory (tcgen05.st, thread = token row = TMEM lane),
// and the GEMM  Z = [X'_0 | X'_1 | X'_2 (| X)] . [W_0 | W_1 | W_2 (| R)]^T  reads its A operand from there
// (tcgen05.mma with A in TMEM).  The accumulator is COUT columns (2*COUT in the stacked split-precision form), so
// the epilogue drains a quarter of what k_tc_gcn drains and is the temporal conv's plain row epilogue; the mixed
// operand never touches shared memory.
//
// Roles (512 threads): warp 0 input-tile producer, warp 2 weight-slab producer (independent rings), warp 1 MMA issuer +
// TMEM owner, warps 4-7 epilogue (TMEM -> bias / ReLU -> split -> 128-bit stores), warps 8-15 mix (two warps per
// TMEM lane quarter, each half of the 64 channels of a K-block).
// The mix works in two passes per K-block so that every input element is unpacked once instead of once per adjacency
// entry: the raw hi / lo words of the own row go to registers and, unchanged, into the first A slot ("plain x": it carries
// the gcn_residual weights -- identity or folded 1x1 conv -- plus W_0 when every self link is exactly 1); x = hi + lo is
// then written back IN PLACE over the staged tile as fp32 rows, from which partitions 1 and 2 gather with one FMA per
// adjacency entry and element (first version: 15 k warp-instructions per tile, this one ~6 k; profiles/r2a, r2c).
// TMEM: [accumulator buffers][ring of 4 A slots x 64 columns: 32 columns hi + 32 columns lo of one (K-block, part)].
#pragma once
#include "tc_kernels.cuh"

namespace cosk {

constexpr int kPartSrcMax = 4;  // most sources of one output vertex inside partition 1 or 2

enum : unsigned int {
  kDbgMixXFull = 0x08000000u,
  kDbgMixAEmpty = 0x09000000u,
  kDbgMmaAFull = 0x0a000000u,
};

struct TcGcnpArgs {
  CUtensorMap tm_x;  // block input ring [kOutSlots*2*t_alloc rows][cin], box {64, 128}
  CUtensorMap tm_w;  // [2*cout rows: hi rows, then lo rows][n_parts*cin], box {64, min(2*cout, 256)}; K = part*cin + c, parts in
                     // A-slot order: plain x (gcn_residual weights [+ W_0 if every self link is 1]), [W_0 on a0*x], W_1, W_2
  int x_row;         // first row of the hi plane of the input slot
  int t_alloc;
  int cin;      // multiple of 64
  int n_parts;  // 3: self links are exactly 1 (W_0 folded into the plain-x part); 4: separate a0*x part
  int V;
  int n_tiles, tile_tokens;
  long long n_tokens;
  const int *mix_ptr;  // CSR over (partition * V + output vertex); partition 0 must be diagonal (self links)
  const int *mix_src;
  const float *mix_val;
  EpiArgs epi;  // no residual rows: gcn_residual rides in the GEMM (plain-x part)
  unsigned long long *trace;  // optional phase timers of CTA 0 (COSK_TRACE=1), SM clock cycles: [32..37] mix warp {wait input
                              // K-block, compute, wait A slot, store + signal, total, slots}; [40..44] MMA thread {wait accumulator,
                              // wait A slot, wait weights, issue, total}; [48..50] epilogue warp {wait accumulator, work, total};
                              // [52..54] producer {wait input stage, wait weight stage, total}
  unsigned int *dbg;
};

template <int COUT, bool STACKED>
struct TcGcnpCfg {
  static constexpr int kSlabBytes = 2 * COUT * kBK * 2;  // [COUT hi rows; COUT lo rows] x 64 K of one (part, K-block)
  static constexpr int kXStages = COUT == 256 ? 2 : 3;
  static constexpr int kWStages = COUT == 64 ? 4 : (COUT == 128 ? 3 : 2);
  static constexpr int kXOff = 0;
  static constexpr int kWOff = kXStages * 2 * kABytes;
  static constexpr int kBarOff = kWOff + kWStages * kSlabBytes;
  static constexpr int kBiasOff = kBarOff + 512;
  static constexpr int kSmemBytes = kBiasOff + COUT * 4 + 1024;
  static constexpr int kAccCols = STACKED ? 2 * COUT : COUT;
  static constexpr int kASlots = 4;
  static constexpr int kSlotCols = 64;
  static constexpr int kAccBufs = (2 * kAccCols + kASlots * kSlotCols <= 512) ? 2 : 1;
  static constexpr int kAOffCols = kAccBufs * kAccCols;
  static constexpr int kTmemCols = 512;
  static_assert(!STACKED || 2 * COUT <= 256, "stacked operand rows");
  static_assert(kAOffCols + kASlots * kSlotCols <= 512, "TMEM columns");
  static_assert(kSmemBytes <= kSmemLimit, "shared memory budget");
};

// 8 fp32 values -> 4 packed hi words + 4 packed lo words (even channel in the low half, as in memory)
__device__ __forceinline__ void split8(const float (&v)[8], uint32_t *hi, uint32_t *lo) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint32_t h = pack_bf16x2(v[2 * j], v[2 * j + 1]);
    hi[j] = h;
    lo[j] = pack_bf16x2(v[2 * j] - bf16_lo_as_float(h), v[2 * j + 1] - bf16_hi_as_float(h));
  }
}

// acc += coef * (hi + lo) for the 8 channels of one 16-byte chunk of each plane
__device__ __forceinline__ void fma_chunk(float (&acc)[8], float coef, const uint4 &h, const uint4 &l) {
  const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    acc[2 * j] = fmaf(coef, bf16_lo_as_float(hw[j]), acc[2 * j]);
    acc[2 * j] = fmaf(coef, bf16_lo_as_float(lw[j]), acc[2 * j]);
    acc[2 * j + 1] = fmaf(coef, bf16_hi_as_float(hw[j]), acc[2 * j + 1]);
    acc[2 * j + 1] = fmaf(coef, bf16_hi_as_float(lw[j]), acc[2 * j + 1]);
  }
}

// byte offset of the 16-byte chunk `c` of row `r` inside a TMA SWIZZLE_128B tile with 128-byte rows
__device__ __forceinline__ uint32_t sw128_off(int r, int c) { return (uint32_t)(r * 128 + ((c ^ (r & 7)) << 4)); }

template <int COUT, bool STACKED>
__device__ __forceinline__ void gcnp_body(const TcGcnpArgs &a, uint8_t *smem_raw, const int cta, const int ncta) {
  using Cfg = TcGcnpCfg<COUT, STACKED>;
  uint8_t *smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t *xfull = reinterpret_cast<uint64_t *>(smem + Cfg::kBarOff);
  uint64_t *xempty = xfull + Cfg::kXStages;
  uint64_t *wfull = xempty + Cfg::kXStages;
  uint64_t *wempty = wfull + Cfg::kWStages;
  uint64_t *afull = wempty + Cfg::kWStages;
  uint64_t *aempty = afull + Cfg::kASlots;
  uint64_t *tfull = aempty + Cfg::kASlots;
  uint64_t *tempty = tfull + 2;
  uint64_t *mixbar = tempty + 2;  // two rendezvous points of the eight mix warps per K-block
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(mixbar + 2);
  float *bias_s = reinterpret_cast<float *>(smem + Cfg::kBiasOff);
  const uint32_t smem_base = ptx::smem_u32(smem);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_trigger();
  if (threadIdx.x == 0) {
    ptx::mbar_init(&mixbar[0], 8);
    ptx::mbar_init(&mixbar[1], 8);
    for (int s = 0; s < Cfg::kXStages; ++s) {
      ptx::mbar_init(&xfull[s], 1);
      ptx::mbar_init(&xempty[s], 8);  // the eight mix warps
    }
    for (int s = 0; s < Cfg::kWStages; ++s) {
      ptx::mbar_init(&wfull[s], 1);
      ptx::mbar_init(&wempty[s], 1);
    }
    for (int s = 0; s < Cfg::kASlots; ++s) {
      ptx::mbar_init(&afull[s], 8);
      ptx::mbar_init(&aempty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&tfull[s], 1);
      ptx::mbar_init(&tempty[s], 4);
    }
    ptx::fence_barrier_init();
    ptx::prefetch_tmap(&a.tm_x);
    ptx::prefetch_tmap(&a.tm_w);
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, Cfg::kTmemCols);
    ptx::tmem_relinquish();
  }
  for (int i = threadIdx.x; i < COUT; i += blockDim.x) bias_s[i] = a.epi.bias[i];
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // prologue above touched only static data; activations of the previous kernel from here on
  const int nkb = a.cin / kBK;
  const int P = a.n_parts;

  if (warp == 0) {
    // ---- input-tile producer: the K-blocks of every tile's input rows ------------------------------------
    if (lane == 0) {
      PipeState px;
      bool ok = true;
      const bool tr = a.trace != nullptr && cta == 0;
      unsigned long long tw = 0;
      const long long tstart = tr ? clock64() : 0;
      for (int tile = cta; ok && tile < a.n_tiles; tile += ncta) {
        const int row = a.x_row + tile * a.tile_tokens;
        for (int kc = 0; ok && kc < nkb; ++kc) {
          const long long w0 = tr ? clock64() : 0;
          ok = ptx::mbar_wait(&xempty[px.stage], px.phase ^ 1, a.dbg, kDbgProdEmpty | (unsigned)kc);
          if (!ok) break;
          if (tr) tw += clock64() - w0;
          const uint32_t sx = smem_base + Cfg::kXOff + px.stage * 2 * kABytes;
          ptx::mbar_arrive_expect_tx(&xfull[px.stage], 2 * kABytes);
          ptx::tma_load_2d_hint(sx, &a.tm_x, &xfull[px.stage], kc * kBK, row, ptx::kEvictFirst);
          ptx::tma_load_2d_hint(sx + kABytes, &a.tm_x, &xfull[px.stage], kc * kBK, row + a.t_alloc, ptx::kEvictFirst);
          px.advance<Cfg::kXStages>();
        }
      }
      if (tr) {
        a.trace[52] = tw;
        a.trace[54] = clock64() - tstart;
      }
    }
  } else if (warp == 2) {
    // ---- weight producer: one slab per (K-block, part), in the order the GEMM eats them --------------------
    if (lane == 0) {
      PipeState pw;
      bool ok = true;
      const bool tr = a.trace != nullptr && cta == 0;
      unsigned long long tw = 0;
      for (int tile = cta; ok && tile < a.n_tiles; tile += ncta) {
        for (int kc = 0; ok && kc < nkb; ++kc) {
          for (int p = 0; p < P; ++p) {
            const long long w1 = tr ? clock64() : 0;
            ok = ptx::mbar_wait(&wempty[pw.stage], pw.phase ^ 1, a.dbg, kDbgProdEmpty | 0x800000u | (unsigned)(kc * 4 + p));
            if (!ok) break;
            if (tr) tw += clock64() - w1;
            const uint32_t sw = smem_base + Cfg::kWOff + pw.stage * Cfg::kSlabBytes;
            ptx::mbar_arrive_expect_tx(&wfull[pw.stage], Cfg::kSlabBytes);
            const int c0 = p * a.cin + kc * kBK;
            if (COUT <= 128) {
              ptx::tma_load_2d_hint(sw, &a.tm_w, &wfull[pw.stage], c0, 0, ptx::kEvictLast);
            } else {  // 512 rows: two boxes of 256 (hi plane, lo plane)
              ptx::tma_load_2d_hint(sw, &a.tm_w, &wfull[pw.stage], c0, 0, ptx::kEvictLast);
              ptx::tma_load_2d_hint(sw + Cfg::kSlabBytes / 2, &a.tm_w, &wfull[pw.stage], c0, COUT, ptx::kEvictLast);
            }
            pw.advance<Cfg::kWStages>();
          }
        }
      }
      if (tr) a.trace[53] = tw;
    }
  } else if (warp == 1) {
    if (lane == 0) {
      PipeState pw, pa;
      bool ok = true;
      int it = 0;
      constexpr uint32_t idesc = ptx::umma_idesc_bf16(kTileRows, STACKED ? 2 * COUT : COUT);
      const bool tr = a.trace != nullptr && cta == 0;
      unsigned long long tw[4] = {0, 0, 0, 0};
      const long long tstart = tr ? clock64() : 0;
      for (int tile = cta; ok && tile < a.n_tiles; tile += ncta, ++it) {
        const int acc = it % Cfg::kAccBufs;
        const int use = it / Cfg::kAccBufs;
        const long long w0 = tr ? clock64() : 0;
        ok = ptx::mbar_wait(&tempty[acc], (use & 1) ^ 1, a.dbg, kDbgMmaTmemEmpty | (unsigned)it);
        if (!ok) break;
        if (tr) tw[0] += clock64() - w0;
        ptx::tc_fence_after();
        const uint32_t d = tmem_base + acc * Cfg::kAccCols;
        bool first = true;
        for (int kc = 0; ok && kc < nkb; ++kc) {
          for (int p = 0; p < P; ++p) {
            long long w1 = tr ? clock64() : 0;
            ok = ptx::mbar_wait(&afull[pa.stage], pa.phase, a.dbg, kDbgMmaAFull | (unsigned)(kc * 4 + p));
            if (!ok) break;
            if (tr) { const long long n_ = clock64(); tw[1] += n_ - w1; w1 = n_; }
            ok = ptx::mbar_wait(&wfull[pw.stage], pw.phase, a.dbg, kDbgMmaFull | (unsigned)(kc * 4 + p));
            if (!ok) break;
            if (tr) { const long long n_ = clock64(); tw[2] += n_ - w1; w1 = n_; }
            ptx::tc_fence_after();
            const uint32_t ta = tmem_base + Cfg::kAOffCols + pa.stage * Cfg::kSlotCols;
            const uint32_t sw = smem_base + Cfg::kWOff + pw.stage * Cfg::kSlabBytes;
            const uint32_t bh = ptx::umma_desc_lo(sw), bl = ptx::umma_desc_lo(sw + Cfg::kSlabBytes / 2);
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k) {  // one K-step: 8 TMEM columns of A, 32 bytes of every B row
              if (STACKED) {
                ptx::umma_bf16_ts(d, ta + 8 * k, bh + 2 * k, idesc, (first && k == 0) ? 0u : 1u);
                ptx::umma_bf16_ts(d, ta + 32 + 8 * k, bh + 2 * k, idesc, 1u);
              } else {
                ptx::umma_bf16_ts(d, ta + 8 * k, bh + 2 * k, idesc, (first && k == 0) ? 0u : 1u);
                ptx::umma_bf16_ts(d, ta + 32 + 8 * k, bh + 2 * k, idesc, 1u);
                ptx::umma_bf16_ts(d, ta + 8 * k, bl + 2 * k, idesc, 1u);
              }
            }
            first = false;
            ptx::umma_commit(&aempty[pa.stage]);
            ptx::umma_commit(&wempty[pw.stage]);
            pa.advance<Cfg::kASlots>();
            pw.advance<Cfg::kWStages>();
            if (tr) tw[3] += clock64() - w1;
          }
        }
        if (ok) ptx::umma_commit(&tfull[acc]);
      }
      if (tr) {
        for (int i = 0; i < 4; ++i) a.trace[40 + i] = tw[i];
        a.trace[44] = clock64() - tstart;
      }
    }
  } else if (warp >= 4 && warp < 8) {
    const int q = warp & 3;
    bool ok = true;
    int it = 0;
    const bool tr = a.trace != nullptr && cta == 0 && q == 0 && lane == 0;
    unsigned long long tw[2] = {0, 0};
    const long long tstart = tr ? clock64() : 0;
    for (int tile = cta; ok && tile < a.n_tiles; tile += ncta, ++it) {
      const int acc = it % Cfg::kAccBufs;
      const int use = it / Cfg::kAccBufs;
      long long w0 = tr ? clock64() : 0;
      ok = ptx::mbar_wait(&tfull[acc], use & 1, a.dbg, kDbgEpiTmemFull | (unsigned)it);
      if (!ok) break;
      if (tr) { const long long n_ = clock64(); tw[0] += n_ - w0; w0 = n_; }
      ptx::tc_fence_after();
      const int row = q * 32 + lane;
      const long long tok = (long long)tile * a.tile_tokens + row;
      const bool valid = row < a.tile_tokens && tok < a.n_tokens;
      epilogue_rows<COUT, STACKED>(tmem_base + ((uint32_t)(q * 32) << 16) + acc * Cfg::kAccCols, bias_s, a.epi, tok, valid);
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tempty[acc]);
      if (tr) tw[1] += clock64() - w0;
    }
    if (tr) {
      a.trace[48] = tw[0];
      a.trace[49] = tw[1];
      a.trace[50] = clock64() - tstart;
    }
  } else if (warp >= 8) {
    // ---- mix: thread = token row (TMEM lane), this warp's half of the K-block's 64 channels -------------
    const int q = warp & 3, h = (warp - 8) >> 2;
    const int row = q * 32 + lane;
    // CSR of this row (tile invariant: tiles are skeleton aligned).  Sources as byte offset of the fp32 row (its
    // swizzle key is bits 8..10); unused entries carry a zero coefficient and point at the own row; loops run to
    // the warp's maximum count.
    uint32_t s_off[2][kPartSrcMax];
    float s_cf[2][kPartSrcMax];
    int nmax[2] = {0, 0};
    float d0 = 0.f;
    {
      const int wv = row % a.V, sk0 = row - wv;
      const bool live = row < a.tile_tokens;
      if (live) {
        const int e0 = a.mix_ptr[wv];
        if (a.mix_ptr[wv + 1] > e0) d0 = a.mix_val[e0];
      }
#pragma unroll
      for (int p = 0; p < 2; ++p) {
        const int eb = a.mix_ptr[(p + 1) * a.V + wv], ee = a.mix_ptr[(p + 1) * a.V + wv + 1];
        const int n = live ? min(ee - eb, kPartSrcMax) : 0;
#pragma unroll
        for (int e = 0; e < kPartSrcMax; ++e) {
          const bool on = e < n;
          s_off[p][e] = (uint32_t)(on ? sk0 + a.mix_src[eb + e] : row) * 256u;
          s_cf[p][e] = on ? a.mix_val[eb + e] : 0.f;
        }
        int m = n;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
        nmax[p] = m;
      }
    }
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16) + Cfg::kAOffCols + 16 * h;
    PipeState px, pa;
    bool ok = true;
    uint32_t nk = 0;  // K-blocks mixed so far: parity of the two rendezvous barriers
    const bool tr = a.trace != nullptr && cta == 0 && warp == 8 && lane == 0;
    unsigned long long tw[4] = {0, 0, 0, 0};
    unsigned long long nslots = 0;
    long long tc = 0;
    const long long tstart = tr ? clock64() : 0;
    auto lap = [&](int i) {
      if (tr) {
        const long long n_ = clock64();
        tw[i] += n_ - tc;
        tc = n_;
      }
    };
    // one finished part -> its A slot in tensor memory
    auto put = [&](const uint32_t (&oh)[16], const uint32_t (&ol)[16], unsigned code) {
      lap(1);
      ok = ptx::mbar_wait(&aempty[pa.stage], pa.phase ^ 1, a.dbg, kDbgMixAEmpty | code);
      if (!ok) return;
      lap(2);
      ptx::tc_fence_after();
      const uint32_t ta = lane_addr + pa.stage * Cfg::kSlotCols;
      ptx::tmem_st_32x16(ta, oh);
      ptx::tmem_st_32x16(ta + 32, ol);
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&afull[pa.stage]);
      pa.advance<Cfg::kASlots>();
      lap(3);
      ++nslots;
    };
    // rendezvous of the mix warps through an mbarrier (bounded wait: a failed pipeline cannot park a warp forever)
    auto mix_sync = [&](int which) {
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&mixbar[which]);
      ok = ptx::mbar_wait(&mixbar[which], nk & 1u, a.dbg, kDbgMixXFull | 0x800000u | (unsigned)which);
    };
    for (int tile = cta; ok && tile < a.n_tiles; tile += ncta) {
      for (int kc = 0; ok && kc < nkb; ++kc, ++nk) {
        if (tr) tc = clock64();
        ok = ptx::mbar_wait(&xfull[px.stage], px.phase, a.dbg, kDbgMixXFull | (unsigned)kc);
        if (!ok) break;
        lap(0);
        uint8_t *xs = smem + Cfg::kXOff + px.stage * 2 * kABytes;
        // pass 1: own row, raw words: 4 chunks of 8 channels per plane
        uint32_t rh[16], rl[16];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t off = sw128_off(row, 4 * h + j);
          const uint4 hv = *reinterpret_cast<const uint4 *>(xs + off);
          const uint4 lv = *reinterpret_cast<const uint4 *>(xs + kABytes + off);
          rh[4 * j] = hv.x; rh[4 * j + 1] = hv.y; rh[4 * j + 2] = hv.z; rh[4 * j + 3] = hv.w;
          rl[4 * j] = lv.x; rl[4 * j + 1] = lv.y; rl[4 * j + 2] = lv.z; rl[4 * j + 3] = lv.w;
        }
        mix_sync(0);  // every raw word of the staged K-block is in registers: it may be overwritten
        if (!ok) break;
        // pass 2: x = hi + lo as fp32, in place: row r at bytes [256 r, 256 r + 256), 16-byte chunks XOR-swizzled with r & 7
#pragma unroll
        for (int b = 0; b < 8; ++b) {
          const float4 f4 = make_float4(bf16_lo_as_float(rh[2 * b]) + bf16_lo_as_float(rl[2 * b]), bf16_hi_as_float(rh[2 * b]) + bf16_hi_as_float(rl[2 * b]),
                                        bf16_lo_as_float(rh[2 * b + 1]) + bf16_lo_as_float(rl[2 * b + 1]),
                                        bf16_hi_as_float(rh[2 * b + 1]) + bf16_hi_as_float(rl[2 * b + 1]));
          *reinterpret_cast<float4 *>(xs + row * 256 + (((8 * h + b) ^ (row & 7)) << 4)) = f4;
        }
        put(rh, rl, 0);  // part "plain x": the words pass through unchanged
        if (!ok) break;
        if (P == 4) {  // self links with a coefficient other than 1: a0 * x from the own registers
          uint32_t oh[16], ol[16];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float a8[8];
#pragma unroll
            for (int w = 0; w < 4; ++w) {
              a8[2 * w] = d0 * (bf16_lo_as_float(rh[4 * j + w]) + bf16_lo_as_float(rl[4 * j + w]));
              a8[2 * w + 1] = d0 * (bf16_hi_as_float(rh[4 * j + w]) + bf16_hi_as_float(rl[4 * j + w]));
            }
            split8(a8, oh + 4 * j, ol + 4 * j);
          }
          put(oh, ol, 1);
          if (!ok) break;
        }
        mix_sync(1);  // fp32 rows complete
        if (!ok) break;
#pragma unroll
        for (int pp = 0; pp < 2; ++pp) {
          const int n = nmax[pp];
          uint32_t oh[16], ol[16];
#pragma unroll
          for (int jj = 0; jj < 2; ++jj) {  // 16 channels at a time: four 16-byte loads in flight per adjacency entry
            float acc[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) acc[e] = 0.f;
#pragma unroll
            for (int e = 0; e < kPartSrcMax; ++e) {
              if (e < n) {
                const uint32_t off = s_off[pp][e];
                const uint8_t *base = xs + off;
                const uint32_t key = (off >> 8) & 7u;
                float4 v[4];
#pragma unroll
                for (int b = 0; b < 4; ++b) v[b] = *reinterpret_cast<const float4 *>(base + (((uint32_t)(8 * h + 4 * jj + b) ^ key) << 4));
                const float cf = s_cf[pp][e];
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                  acc[4 * b] = fmaf(cf, v[b].x, acc[4 * b]);
                  acc[4 * b + 1] = fmaf(cf, v[b].y, acc[4 * b + 1]);
                  acc[4 * b + 2] = fmaf(cf, v[b].z, acc[4 * b + 2]);
                  acc[4 * b + 3] = fmaf(cf, v[b].w, acc[4 * b + 3]);
                }
              }
            }
            float lo8[8], hi8[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              lo8[e] = acc[e];
              hi8[e] = acc[8 + e];
            }
            split8(lo8, oh + 8 * jj, ol + 8 * jj);
            split8(hi8, oh + 8 * jj + 4, ol + 8 * jj + 4);
          }
          put(oh, ol, 2 + pp);
          if (!ok) break;
        }
        if (!ok) break;
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&xempty[px.stage]);  // this warp is done with the staged K-block
        px.advance<Cfg::kXStages>();
      }
    }
    if (tr) {
      for (int i = 0; i < 4; ++i) a.trace[32 + i] = tw[i];
      a.trace[36] = clock64() - tstart;
      a.trace[37] = nslots;
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

template <int COUT, bool STACKED>
__global__ void __launch_bounds__(512, 1) k_tc_gcnp(const __grid_constant__ TcGcnpArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  gcnp_body<COUT, STACKED>(a, smem_raw, (int)blockIdx.x, (int)gridDim.x);
}

}  // namespace cosk
