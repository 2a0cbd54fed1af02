// tcgen05 / TMEM / TMA tile kernels for the two dense contractions of a CoST-GCN block step.
//
// Orientation: D[128 tokens x COUT] += A[128 tokens x 64] * B[COUT x 64]^T per K-block, i.e. the
// token rows of 5 skeletons (125 of 128 rows valid for V = 25) are the UMMA M dimension and the
// BN-folded weights are the N x K operand.  State and weights are split-bf16 (hi/lo planes); every
// K-block issues the three products hi*hi + lo*hi + hi*lo into one fp32 TMEM accumulator.
//
//   k_tc_tcn<COUT>      9-tap temporal conv over the ring (+ folded strided residual conv as extra
//                       K-blocks) + bias + identity residual + ReLU.     A and B arrive by TMA.
//   k_tc_gcn<COUT>      graph conv: the adjacency mix x*A_i is applied by CUDA cores while staging
//                       the A operand (sparse CSR rows held in registers), then 3 (+1) K-blocks per
//                       64 input channels; bias + identity residual + ReLU; result -> ring slot.
//
// Both are persistent (one CTA per SM loops over token tiles) and warp specialised:
//   warp 0  TMA producer (activations)     warp 1  MMA issuer + TMEM owner
//   warp 2  TMA producer (weights, gcn)    warps 4-7  epilogue (TMEM -> registers -> HBM)
//   warps 8-11 (gcn only) adjacency-mix producers of the A operand
#pragma once
#include "common.cuh"
#include "ptx.cuh"

namespace cosk {

constexpr int kBK = 64;                      // K elements per block = one 128-byte swizzle row
constexpr int kABytes = kTileRows * kBK * 2; // 16 KB per plane
constexpr int kSmemLimit = 232448;           // 227 KB opt-in maximum per CTA

// error codes written to the debug word when a bounded wait expires: role << 24 | stage info
enum : unsigned int {
  kDbgProdEmpty = 0x01000000u,
  kDbgMmaFull = 0x02000000u,
  kDbgMmaTmemEmpty = 0x03000000u,
  kDbgEpiTmemFull = 0x04000000u,
  kDbgWProdEmpty = 0x05000000u,
  kDbgMixXFull = 0x06000000u,
  kDbgMixAEmpty = 0x07000000u,
  kDbgMmaAFull = 0x08000000u,
  kDbgMmaXFull = 0x09000000u,
  kDbgXProdEmpty = 0x0a000000u,
};

struct PipeState {
  int stage = 0;
  uint32_t phase = 0;
  template <int N>
  __device__ __forceinline__ void advance() {
    if (++stage == N) {
      stage = 0;
      phase ^= 1;
    }
  }
};

// The three split-precision products of one K-block (4 UMMA K-steps each).
template <int COUT>
__device__ __forceinline__ void issue_kblock(uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo,
                                             bool first) {
  constexpr uint32_t idesc = ptx::umma_idesc_bf16(kTileRows, COUT);
  const uint32_t a_sel[3] = {a_hi, a_lo, a_hi};
  const uint32_t b_sel[3] = {b_hi, b_hi, b_lo};
#pragma unroll
  for (int p = 0; p < 3; ++p) {
#pragma unroll
    for (int k = 0; k < kBK / 16; ++k) {
      const uint64_t da = ptx::umma_desc_sw128(a_sel[p] + k * 32);
      const uint64_t db = ptx::umma_desc_sw128(b_sel[p] + k * 32);
      ptx::umma_bf16(tmem_d, da, db, idesc, (first && p == 0 && k == 0) ? 0u : 1u);
    }
  }
}

struct EpiArgs {
  const float *bias;                // [COUT]
  const __nv_bfloat16 *r_hi, *r_lo; // identity residual rows (nullptr: none)
  int cs_r;
  __nv_bfloat16 *y_hi, *y_lo;
  int cs_out;
};

// One epilogue warp: 32 accumulator rows, COUT fp32 columns each -> bias, residual, ReLU, split,
// 128-bit stores of the row's channel vector.
template <int COUT>
__device__ __forceinline__ void epilogue_rows(uint32_t taddr_row0, const float *bias_s, const EpiArgs &e, long long tok,
                                              bool valid) {
#pragma unroll 1
  for (int c0 = 0; c0 < COUT; c0 += 32) {
    uint32_t r[32];
    ptx::tmem_ld_32x32(taddr_row0 + c0, r);
    ptx::tmem_ld_wait();
    if (valid) {
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) + bias_s[c0 + j];
      if (e.r_hi != nullptr) {
        const uint4 *ph = reinterpret_cast<const uint4 *>(e.r_hi + tok * e.cs_r + c0);
        const uint4 *pl = reinterpret_cast<const uint4 *>(e.r_lo + tok * e.cs_r + c0);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint4 h = ptx::ldg_v4(ph + q), l = ptx::ldg_v4(pl + q);
          const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            v[q * 8 + w * 2] += bf16_lo_as_float(hw[w]) + bf16_lo_as_float(lw[w]);
            v[q * 8 + w * 2 + 1] += bf16_hi_as_float(hw[w]) + bf16_hi_as_float(lw[w]);
          }
        }
      }
      uint32_t oh[16], ol[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float x0 = fmaxf(v[2 * j], 0.f), x1 = fmaxf(v[2 * j + 1], 0.f);
        const uint32_t h = pack_bf16x2(x0, x1);
        oh[j] = h;
        ol[j] = pack_bf16x2(x0 - bf16_lo_as_float(h), x1 - bf16_hi_as_float(h));
      }
      uint4 *qh = reinterpret_cast<uint4 *>(e.y_hi + tok * e.cs_out + c0);
      uint4 *ql = reinterpret_cast<uint4 *>(e.y_lo + tok * e.cs_out + c0);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        ptx::stg_v4(qh + q, make_uint4(oh[4 * q], oh[4 * q + 1], oh[4 * q + 2], oh[4 * q + 3]));
        ptx::stg_v4(ql + q, make_uint4(ol[4 * q], ol[4 * q + 1], ol[4 * q + 2], ol[4 * q + 3]));
      }
    }
  }
}

// =============================================================================================
// temporal conv
// =============================================================================================
struct TcTcnArgs {
  CUtensorMap tm_ring;  // [slots*2*t_alloc rows][C]   box {64, 128}
  CUtensorMap tm_res;   // previous block's output ring [kOutSlots*2*t_alloc rows][Cr], box {64, 128}
  CUtensorMap tm_w;     // [2*COUT rows (hi then lo)][9*C + Cr]  box {64, COUT}
  int tap_row[kTaps];   // first row of the hi plane of each tap's slot (oldest .. newest)
  int res_row;          // first row of the hi plane of the residual slot
  int t_alloc;          // rows per plane
  int kb_per_tap;       // C / 64
  int kb_res;           // Cr / 64 when res_kind == 2 else 0
  int n_tiles, tile_tokens;
  long long n_tokens;
  EpiArgs epi;
  unsigned int *dbg;
};

template <int COUT>
struct TcTcnCfg {
  static constexpr int kBBytes = COUT * kBK * 2;
  static constexpr int kStageBytes = 2 * kABytes + 2 * kBBytes;
  static constexpr int kStages = COUT == 64 ? 4 : (COUT == 128 ? 3 : 2);
  static constexpr int kBarOff = kStages * kStageBytes;
  static constexpr int kBiasOff = kBarOff + 256;
  static constexpr int kSmemBytes = kBiasOff + COUT * 4 + 1024;  // + slack for manual 1024-B alignment
  static constexpr int kTmemCols = 2 * COUT;                     // double-buffered accumulator
  static_assert(kSmemBytes <= kSmemLimit, "shared memory budget");
  static_assert(kTmemCols <= 512 && (kTmemCols & (kTmemCols - 1)) == 0, "TMEM columns");
};

template <int COUT>
__global__ void __launch_bounds__(256, 1) k_tc_tcn(const __grid_constant__ TcTcnArgs a) {
  using Cfg = TcTcnCfg<COUT>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment in the shared window: align on the shared address
  uint8_t *smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t *full = reinterpret_cast<uint64_t *>(smem + Cfg::kBarOff);
  uint64_t *empty = full + Cfg::kStages;
  uint64_t *tfull = empty + Cfg::kStages;
  uint64_t *tempty = tfull + 2;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tempty + 2);
  float *bias_s = reinterpret_cast<float *>(smem + Cfg::kBiasOff);
  const uint32_t smem_base = ptx::smem_u32(smem);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) {
      ptx::mbar_init(&full[s], 1);
      ptx::mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&tfull[s], 1);
      ptx::mbar_init(&tempty[s], 4);
    }
    ptx::fence_barrier_init();
    ptx::prefetch_tmap(&a.tm_ring);
    ptx::prefetch_tmap(&a.tm_w);
    if (a.kb_res) ptx::prefetch_tmap(&a.tm_res);
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
  const int nkb = kTaps * a.kb_per_tap + a.kb_res;

  if (warp == 0) {
    if (lane == 0) {
      PipeState ps;
      bool ok = true;
      for (int tile = blockIdx.x; ok && tile < a.n_tiles; tile += gridDim.x) {
        const int tok0 = tile * a.tile_tokens;
        for (int kb = 0; kb < nkb; ++kb) {
          ok = ptx::mbar_wait(&empty[ps.stage], ps.phase ^ 1, a.dbg, kDbgProdEmpty | (unsigned)kb);
          if (!ok) break;
          const uint32_t st = smem_base + ps.stage * Cfg::kStageBytes;
          ptx::mbar_arrive_expect_tx(&full[ps.stage], Cfg::kStageBytes);
          const CUtensorMap *tm;
          int c0, row;
          if (kb < kTaps * a.kb_per_tap) {
            const int tap = kb / a.kb_per_tap;
            tm = &a.tm_ring;
            c0 = (kb - tap * a.kb_per_tap) * kBK;
            row = a.tap_row[tap] + tok0;
          } else {
            tm = &a.tm_res;
            c0 = (kb - kTaps * a.kb_per_tap) * kBK;
            row = a.res_row + tok0;
          }
          ptx::tma_load_2d(st, tm, &full[ps.stage], c0, row);
          ptx::tma_load_2d(st + kABytes, tm, &full[ps.stage], c0, row + a.t_alloc);
          ptx::tma_load_2d(st + 2 * kABytes, &a.tm_w, &full[ps.stage], kb * kBK, 0);
          ptx::tma_load_2d(st + 2 * kABytes + Cfg::kBBytes, &a.tm_w, &full[ps.stage], kb * kBK, COUT);
          ps.advance<Cfg::kStages>();
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      PipeState ps;
      bool ok = true;
      int it = 0;
      for (int tile = blockIdx.x; ok && tile < a.n_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        ok = ptx::mbar_wait(&tempty[acc], ((it >> 1) & 1) ^ 1, a.dbg, kDbgMmaTmemEmpty | (unsigned)it);
        if (!ok) break;
        ptx::tc_fence_after();
        const uint32_t d = tmem_base + acc * COUT;
        for (int kb = 0; kb < nkb; ++kb) {
          ok = ptx::mbar_wait(&full[ps.stage], ps.phase, a.dbg, kDbgMmaFull | (unsigned)kb);
          if (!ok) break;
          ptx::tc_fence_after();
          const uint32_t st = smem_base + ps.stage * Cfg::kStageBytes;
          issue_kblock<COUT>(d, st, st + kABytes, st + 2 * kABytes, st + 2 * kABytes + Cfg::kBBytes, kb == 0);
          ptx::umma_commit(&empty[ps.stage]);
          ps.advance<Cfg::kStages>();
        }
        if (ok) ptx::umma_commit(&tfull[acc]);
      }
    }
  } else if (warp >= 4) {
    const int q = warp & 3;
    bool ok = true;
    int it = 0;
    for (int tile = blockIdx.x; ok && tile < a.n_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      ok = ptx::mbar_wait(&tfull[acc], (it >> 1) & 1, a.dbg, kDbgEpiTmemFull | (unsigned)it);
      if (!ok) break;
      ptx::tc_fence_after();
      const int row = q * 32 + lane;
      const long long tok = (long long)tile * a.tile_tokens + row;
      const bool valid = row < a.tile_tokens && tok < a.n_tokens;
      epilogue_rows<COUT>(tmem_base + ((uint32_t)(q * 32) << 16) + acc * COUT, bias_s, a.epi, tok, valid);
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tempty[acc]);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

// =============================================================================================
// graph conv
// =============================================================================================
constexpr int kMixMaxNz = 4;  // non-zeros per (partition, output vertex) the register CSR can hold

struct TcGcnArgs {
  CUtensorMap tm_x;  // block input ring [kOutSlots*2*t_alloc rows][CIN], box {64, 128}
  CUtensorMap tm_w;  // [2*COUT rows][(3 + res_conv)*CIN], box {64, COUT}
  int x_row;         // first row of the hi plane of the input slot
  int t_alloc;
  int cin;           // multiple of 64
  int res_conv;      // 1: a 4th K-block per chunk multiplies the raw input with the folded gcn_residual conv
  int V;
  int n_tiles, tile_tokens;
  long long n_tokens;
  const int *mix_ptr;  // CSR over (partition * V + output vertex)
  const int *mix_src;
  const float *mix_val;
  EpiArgs epi;  // r_hi/r_lo = input rows when cin == cout (identity gcn_residual)
  unsigned int *dbg;
};

template <int COUT>
struct TcGcnCfg {
  static constexpr int kBBytes = COUT * kBK * 2;
  static constexpr int kXStages = COUT == 256 ? 1 : 2;  // raw input chunks (hi+lo)
  static constexpr int kAStages = 2;                     // mixed operand (hi+lo)
  static constexpr int kBStages = COUT == 64 ? 4 : 2;    // weights (hi+lo)
  static constexpr int kXOff = 0;
  static constexpr int kAOff = kXOff + kXStages * 2 * kABytes;
  static constexpr int kBOff = kAOff + kAStages * 2 * kABytes;
  static constexpr int kBarOff = kBOff + kBStages * 2 * kBBytes;
  static constexpr int kBiasOff = kBarOff + 256;
  static constexpr int kSmemBytes = kBiasOff + COUT * 4 + 1024;
  static constexpr int kTmemCols = 2 * COUT;
  static_assert(kSmemBytes <= kSmemLimit, "shared memory budget");
};

// byte offset of the 16-byte chunk (row, chunk) inside a 128-byte-row SWIZZLE_128B tile
__device__ __forceinline__ uint32_t sw128_off(int row, int chunk) { return row * 128 + ((chunk ^ (row & 7)) << 4); }

template <int COUT>
__global__ void __launch_bounds__(384, 1) k_tc_gcn(const __grid_constant__ TcGcnArgs a) {
  using Cfg = TcGcnCfg<COUT>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment in the shared window: align on the shared address
  uint8_t *smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t *xfull = reinterpret_cast<uint64_t *>(smem + Cfg::kBarOff);
  uint64_t *xempty = xfull + Cfg::kXStages;
  uint64_t *afull = xempty + Cfg::kXStages;
  uint64_t *aempty = afull + Cfg::kAStages;
  uint64_t *bfull = aempty + Cfg::kAStages;
  uint64_t *bempty = bfull + Cfg::kBStages;
  uint64_t *tfull = bempty + Cfg::kBStages;
  uint64_t *tempty = tfull + 2;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tempty + 2);
  float *bias_s = reinterpret_cast<float *>(smem + Cfg::kBiasOff);
  const uint32_t smem_base = ptx::smem_u32(smem);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int parts = 3 + a.res_conv;
  const int nchunk = a.cin / kBK;
  if (threadIdx.x == 0) {
    for (int s = 0; s < Cfg::kXStages; ++s) {
      ptx::mbar_init(&xfull[s], 1);
      ptx::mbar_init(&xempty[s], 128 + a.res_conv);  // 128 mix threads (+ the MMA commit of part 3)
    }
    for (int s = 0; s < Cfg::kAStages; ++s) {
      ptx::mbar_init(&afull[s], 128);
      ptx::mbar_init(&aempty[s], 1);
    }
    for (int s = 0; s < Cfg::kBStages; ++s) {
      ptx::mbar_init(&bfull[s], 1);
      ptx::mbar_init(&bempty[s], 1);
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

  if (warp == 0) {
    // ---- raw input chunks -------------------------------------------------------------------
    if (lane == 0) {
      PipeState ps;
      bool ok = true;
      for (int tile = blockIdx.x; ok && tile < a.n_tiles; tile += gridDim.x) {
        const int row = a.x_row + tile * a.tile_tokens;
        for (int kc = 0; kc < nchunk; ++kc) {
          ok = ptx::mbar_wait(&xempty[ps.stage], ps.phase ^ 1, a.dbg, kDbgXProdEmpty | (unsigned)kc);
          if (!ok) break;
          const uint32_t st = smem_base + Cfg::kXOff + ps.stage * 2 * kABytes;
          ptx::mbar_arrive_expect_tx(&xfull[ps.stage], 2 * kABytes);
          ptx::tma_load_2d(st, &a.tm_x, &xfull[ps.stage], kc * kBK, row);
          ptx::tma_load_2d(st + kABytes, &a.tm_x, &xfull[ps.stage], kc * kBK, row + a.t_alloc);
          ps.advance<Cfg::kXStages>();
        }
      }
    }
  } else if (warp == 2) {
    // ---- weights ----------------------------------------------------------------------------
    if (lane == 0) {
      PipeState ps;
      bool ok = true;
      for (int tile = blockIdx.x; ok && tile < a.n_tiles; tile += gridDim.x) {
        for (int kc = 0; ok && kc < nchunk; ++kc) {
          for (int p = 0; p < parts; ++p) {
            ok = ptx::mbar_wait(&bempty[ps.stage], ps.phase ^ 1, a.dbg, kDbgWProdEmpty | (unsigned)(kc * 4 + p));
            if (!ok) break;
            const uint32_t st = smem_base + Cfg::kBOff + ps.stage * 2 * Cfg::kBBytes;
            ptx::mbar_arrive_expect_tx(&bfull[ps.stage], 2 * Cfg::kBBytes);
            const int k0 = p * a.cin + kc * kBK;
            ptx::tma_load_2d(st, &a.tm_w, &bfull[ps.stage], k0, 0);
            ptx::tma_load_2d(st + Cfg::kBBytes, &a.tm_w, &bfull[ps.stage], k0, COUT);
            ps.advance<Cfg::kBStages>();
          }
        }
      }
    }
  } else if (warp == 1) {
    // ---- MMA issuer -------------------------------------------------------------------------
    if (lane == 0) {
      PipeState px, pa, pb;
      bool ok = true;
      int it = 0;
      for (int tile = blockIdx.x; ok && tile < a.n_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        ok = ptx::mbar_wait(&tempty[acc], ((it >> 1) & 1) ^ 1, a.dbg, kDbgMmaTmemEmpty | (unsigned)it);
        if (!ok) break;
        ptx::tc_fence_after();
        const uint32_t d = tmem_base + acc * COUT;
        for (int kc = 0; ok && kc < nchunk; ++kc) {
          for (int p = 0; p < parts; ++p) {
            ok = ptx::mbar_wait(&bfull[pb.stage], pb.phase, a.dbg, kDbgMmaFull | (unsigned)(kc * 4 + p));
            if (!ok) break;
            const uint32_t sb = smem_base + Cfg::kBOff + pb.stage * 2 * Cfg::kBBytes;
            uint32_t sa;
            if (p < 3) {
              ok = ptx::mbar_wait(&afull[pa.stage], pa.phase, a.dbg, kDbgMmaAFull | (unsigned)(kc * 4 + p));
              if (!ok) break;
              sa = smem_base + Cfg::kAOff + pa.stage * 2 * kABytes;
            } else {
              ok = ptx::mbar_wait(&xfull[px.stage], px.phase, a.dbg, kDbgMmaXFull | (unsigned)kc);
              if (!ok) break;
              sa = smem_base + Cfg::kXOff + px.stage * 2 * kABytes;
            }
            ptx::tc_fence_after();
            issue_kblock<COUT>(d, sa, sa + kABytes, sb, sb + Cfg::kBBytes, kc == 0 && p == 0);
            ptx::umma_commit(&bempty[pb.stage]);
            pb.advance<Cfg::kBStages>();
            if (p < 3) {
              ptx::umma_commit(&aempty[pa.stage]);
              pa.advance<Cfg::kAStages>();
            } else {
              ptx::umma_commit(&xempty[px.stage]);
            }
          }
          px.advance<Cfg::kXStages>();
        }
        if (ok) ptx::umma_commit(&tfull[acc]);
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // ---- epilogue ---------------------------------------------------------------------------
    const int q = warp & 3;
    bool ok = true;
    int it = 0;
    for (int tile = blockIdx.x; ok && tile < a.n_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      ok = ptx::mbar_wait(&tfull[acc], (it >> 1) & 1, a.dbg, kDbgEpiTmemFull | (unsigned)it);
      if (!ok) break;
      ptx::tc_fence_after();
      const int row = q * 32 + lane;
      const long long tok = (long long)tile * a.tile_tokens + row;
      const bool valid = row < a.tile_tokens && tok < a.n_tokens;
      epilogue_rows<COUT>(tmem_base + ((uint32_t)(q * 32) << 16) + acc * COUT, bias_s, a.epi, tok, valid);
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tempty[acc]);
    }
  } else if (warp >= 8) {
    // ---- adjacency mix: A_part[row w] = sum_v coef(part, v, w) * X[row of v]  ------------------
    const int row = threadIdx.x - 256;  // 0..127, fixed for the whole kernel
    const bool live = row < a.tile_tokens;
    const int wv = row % a.V;
    const int sk0 = row - wv;
    int cnt[3], src[3][kMixMaxNz];
    float coef[3][kMixMaxNz];
#pragma unroll
    for (int p = 0; p < 3; ++p) {
      const int e0 = a.mix_ptr[p * a.V + wv];
      cnt[p] = live ? a.mix_ptr[p * a.V + wv + 1] - e0 : 0;
#pragma unroll
      for (int j = 0; j < kMixMaxNz; ++j) {
        const bool on = j < cnt[p];
        src[p][j] = on ? sk0 + a.mix_src[e0 + j] : 0;
        coef[p][j] = on ? a.mix_val[e0 + j] : 0.f;
      }
    }
    PipeState px, pa;
    bool ok = true;
    for (int tile = blockIdx.x; ok && tile < a.n_tiles; tile += gridDim.x) {
      for (int kc = 0; ok && kc < nchunk; ++kc) {
        ok = ptx::mbar_wait(&xfull[px.stage], px.phase, a.dbg, kDbgMixXFull | (unsigned)kc);
        if (!ok) break;
        const uint8_t *xh = smem + Cfg::kXOff + px.stage * 2 * kABytes;
        const uint8_t *xl = xh + kABytes;
#pragma unroll
        for (int p = 0; p < 3; ++p) {
          ok = ptx::mbar_wait(&aempty[pa.stage], pa.phase ^ 1, a.dbg, kDbgMixAEmpty | (unsigned)(kc * 4 + p));
          if (!ok) break;
          uint8_t *ah = smem + Cfg::kAOff + pa.stage * 2 * kABytes;
          uint8_t *al = ah + kABytes;
#pragma unroll 2
          for (int ch = 0; ch < 8; ++ch) {
            float m[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) m[j] = 0.f;
#pragma unroll
            for (int e = 0; e < kMixMaxNz; ++e) {
              if (e < cnt[p]) {
                const uint32_t off = sw128_off(src[p][e], ch);
                const uint4 h = *reinterpret_cast<const uint4 *>(xh + off);
                const uint4 l = *reinterpret_cast<const uint4 *>(xl + off);
                const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
                const float cf = coef[p][e];
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                  m[2 * w] = fmaf(cf, bf16_lo_as_float(hw[w]) + bf16_lo_as_float(lw[w]), m[2 * w]);
                  m[2 * w + 1] = fmaf(cf, bf16_hi_as_float(hw[w]) + bf16_hi_as_float(lw[w]), m[2 * w + 1]);
                }
              }
            }
            uint32_t oh[4], ol[4];
#pragma unroll
            for (int w = 0; w < 4; ++w) {
              const uint32_t h = pack_bf16x2(m[2 * w], m[2 * w + 1]);
              oh[w] = h;
              ol[w] = pack_bf16x2(m[2 * w] - bf16_lo_as_float(h), m[2 * w + 1] - bf16_hi_as_float(h));
            }
            const uint32_t off = sw128_off(row, ch);
            *reinterpret_cast<uint4 *>(ah + off) = make_uint4(oh[0], oh[1], oh[2], oh[3]);
            *reinterpret_cast<uint4 *>(al + off) = make_uint4(ol[0], ol[1], ol[2], ol[3]);
          }
          ptx::fence_proxy_async_smem();
          ptx::mbar_arrive(&afull[pa.stage]);
          pa.advance<Cfg::kAStages>();
        }
        if (!ok) break;
        ptx::mbar_arrive(&xempty[px.stage]);  // this thread no longer reads the raw chunk
        px.advance<Cfg::kXStages>();
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

}  // namespace cosk
