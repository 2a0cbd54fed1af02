// fp32 CUDA-core kernels: input staging, generic graph-conv / temporal-conv tiles for any channel
// count, pooling + FC, block read-back.  The graph/temporal kernels here serve (a) channel counts
// the tcgen05 tiles do not cover (the 3-channel first layer, the reference's 2/4-channel test
// blocks) and (b) as the on-device checker of the tcgen05 kernels (COSK_PATH_SIMT).
#pragma once
#include "common.cuh"

namespace cosk {

// ---------------------------------------------------------------------------------------------
// input: one frame (N, C, V, S) fp32 -> token-major split-bf16 rows, with data_bn folded in.
// Reference: reshape1 / data_bn / reshape2, models/base.py:73-82 (feature f = s*V*C + v*C + c).
// ---------------------------------------------------------------------------------------------
// Time-batched launch: blockIdx.y = frame f of the chunk, read at x + f * x_frame_stride and written to slot ow.slot(f)
// of the input ring (hi / lo then point at slot 0; a single-frame launch passes the slot's own pointers and a unit walk).
__global__ void k_input(const float *__restrict__ x, long long nc_stride, int C, int V, int S,
                        const float *__restrict__ scale, const float *__restrict__ shift,
                        __nv_bfloat16 *__restrict__ hi, __nv_bfloat16 *__restrict__ lo, int cs, long long n_tokens,
                        long long x_frame_stride, RingWalk ow, long long slot_elems) {
  pdl_trigger();
  pdl_wait();
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_tokens * cs) return;
  {
    const int f = (int)blockIdx.y;
    x += (long long)f * x_frame_stride;
    const long long o = (long long)ow.slot(f) * slot_elems;
    hi += o;
    lo += o;
  }
  long long tok = idx / cs;
  int c = (int)(idx % cs);
  float val = 0.f;
  if (c < C) {
    long long skel = tok / V;
    int v = (int)(tok % V);
    long long n = skel / S;
    int s = (int)(skel % S);
    val = x[(n * C + c) * nc_stride + (long long)v * S + s];
    if (scale != nullptr) {
      int f = (s * V + v) * C + c;
      val = val * scale[f] + shift[f];
    }
  }
  __nv_bfloat16 h, l;
  split_bf16(val, h, l);
  hi[idx] = h;
  lo[idx] = l;
}

// ---------------------------------------------------------------------------------------------
// Shared 128x64 register-tiled fp32 GEMM step: acc[8][4] += As[k][row] * Bs[k][col], k < 16.
// 256 threads: ty = tid / 16 owns rows ty*8..+8, tx = tid % 16 owns cols tx*4..+4.
// ---------------------------------------------------------------------------------------------
constexpr int kSimtK = 16;
constexpr int kSimtN = 64;

__device__ __forceinline__ void simt_mma_chunk(float (&acc)[8][4], const float (*As)[kTileRows],
                                               const float (*Bs)[kSimtN], int ty, int tx) {
#pragma unroll
  for (int k = 0; k < kSimtK; ++k) {
    float a[8], b[4];
    const float4 a0 = *reinterpret_cast<const float4 *>(&As[k][ty * 8]);
    const float4 a1 = *reinterpret_cast<const float4 *>(&As[k][ty * 8 + 4]);
    const float4 b0 = *reinterpret_cast<const float4 *>(&Bs[k][tx * 4]);
    a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w;
    a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
    b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
  }
}

// Load a [16 channels][128 rows] fp32 chunk of a split-bf16 plane pair into smem (transposed).
__device__ __forceinline__ void simt_load_rows(float (*dst)[kTileRows], const __nv_bfloat16 *__restrict__ hi,
                                               const __nv_bfloat16 *__restrict__ lo, int cs, int c0, int c_valid,
                                               long long tok0, int rows_valid) {
  const int row = threadIdx.x >> 1;
  const int half = (threadIdx.x & 1) * 8;
  const bool ok = row < rows_valid;
  const long long base = (tok0 + row) * cs;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = c0 + half + j;
    float v = 0.f;
    if (ok && c < c_valid) v = join_bf16(hi[base + c], lo[base + c]);
    dst[half + j][row] = v;
  }
}

// Load a [16][64] chunk of a k-major fp32 weight matrix W[k][n_total].
__device__ __forceinline__ void simt_load_w(float (*dst)[kSimtN], const float *__restrict__ w, int ld, int k0,
                                            int k_valid, int n0, int n_total) {
  const int k = threadIdx.x >> 4;
  const int n4 = (threadIdx.x & 15) * 4;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int n = n0 + n4 + j;
    float v = 0.f;
    if (k < k_valid && n < n_total) v = w[(long long)(k0 + k) * ld + n];
    dst[k][n4 + j] = v;
  }
}

struct GcnArgs {
  const __nv_bfloat16 *x_hi, *x_lo;  // block input rows
  int cs_in, cin;
  __nv_bfloat16 *y_hi, *y_lo;  // temporal-ring slot
  int cs_out, cout;
  const float *w;     // [3*cin (+cin)][cout]  k-major, BN folded
  const float *bias;  // [cout]
  int res_conv;       // 1: rows 3*cin.. of w are the folded gcn_residual 1x1 conv (cin != cout)
  int res_identity;   // 1: add x (cin == cout)
  const int *mix_ptr;  // CSR over (partition, output vertex): sources and coefficients of A*graph_attn
  const int *mix_src;
  const float *mix_val;
  int V;
  long long n_tokens;
  int tile_tokens;
  // adaptive graph conv: per-token dense mixing rows written by k_agcn_attn replace the CSR above;
  // row of output token w, partition p, source vertex v: dense[w*dense_ld + p*dense_vp + v]
  const float *dense;
  int dense_ld, dense_vp;
  // time-batched launch of k_gcn_small (blockIdx.y = frame): input slot in_walk.slot(f), ring slot out_walk.slot(f);
  // x_* / y_* then point at slot 0 of their rings
  RingWalk in_walk, out_walk;
  long long in_slot_elems = 0, out_slot_elems = 0;
};

// Graph convolution of one frame: z = sum_i W_i (x A_i) ; BN ; + gcn_residual(x) ; ReLU
// (GraphConvolution.forward, models/base.py:260-270), one 128-token x 64-channel tile per CTA.
__global__ void __launch_bounds__(256) k_gcn_simt(GcnArgs a) {
  __shared__ __align__(16) float Xs[kSimtK][kTileRows];
  __shared__ __align__(16) float As[kSimtK][kTileRows];
  __shared__ __align__(16) float Bs[kSimtK][kSimtN];
  pdl_trigger();
  pdl_wait();
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  const long long tok0 = (long long)blockIdx.x * a.tile_tokens;
  const int n0 = blockIdx.y * kSimtN;
  long long remain = a.n_tokens - tok0;
  const int rows_valid = (int)(remain < a.tile_tokens ? remain : a.tile_tokens);
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int parts = 3 + a.res_conv;
  for (int c0 = 0; c0 < a.cin; c0 += kSimtK) {
    __syncthreads();
    simt_load_rows(Xs, a.x_hi, a.x_lo, a.cs_in, c0, a.cin, tok0, rows_valid);
    __syncthreads();
    const int kv = min(kSimtK, a.cin - c0);
    for (int part = 0; part < parts; ++part) {
      if (part < 3) {
        // adjacency mix of this partition: As[k][w] = sum_v A_eff[part][v][w] * Xs[k][v]
        const int row = threadIdx.x >> 1;
        const int half = (threadIdx.x & 1) * 8;
        float m[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) m[j] = 0.f;
        if (row < rows_valid) {
          const int wv = row % a.V;
          const int sk0 = row - wv;
          if (a.dense != nullptr) {
            const float *dm = a.dense + (tok0 + row) * a.dense_ld + part * a.dense_vp;
            for (int v = 0; v < a.V; ++v) {
              const float coef = dm[v];
#pragma unroll
              for (int j = 0; j < 8; ++j) m[j] = fmaf(coef, Xs[half + j][sk0 + v], m[j]);
            }
          } else {
            const int e1 = a.mix_ptr[part * a.V + wv + 1];
            for (int e = a.mix_ptr[part * a.V + wv]; e < e1; ++e) {
              const int src = sk0 + a.mix_src[e];
              const float coef = a.mix_val[e];
#pragma unroll
              for (int j = 0; j < 8; ++j) m[j] = fmaf(coef, Xs[half + j][src], m[j]);
            }
          }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) As[half + j][row] = m[j];
      }
      simt_load_w(Bs, a.w, a.cout, part * a.cin + c0, kv, n0, a.cout);
      __syncthreads();
      simt_mma_chunk(acc, part < 3 ? As : Xs, Bs, ty, tx);
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = ty * 8 + i;
    if (row >= rows_valid) continue;
    const long long tok = tok0 + row;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= a.cout) continue;
      float v = acc[i][j] + a.bias[n];
      if (a.res_identity) v += join_bf16(a.x_hi[tok * a.cs_in + n], a.x_lo[tok * a.cs_in + n]);
      v = fmaxf(v, 0.f);
      __nv_bfloat16 h, l;
      split_bf16(v, h, l);
      a.y_hi[tok * a.cs_out + n] = h;
      a.y_lo[tok * a.cs_out + n] = l;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Graph conv for a narrow input (cin <= 8: the 3-channel first layer).  K = (3 + res_conv) * cin <= 32
// is far below a tensor-core tile, and the kernel is bound by its output write, so each thread simply
// owns one token row x 32 output channels: mixed inputs and weights sit in shared memory.
// ---------------------------------------------------------------------------------------------
constexpr int kSmallKMax = 32;

__global__ void __launch_bounds__(256) k_gcn_small(GcnArgs a) {
  {
    const int f = (int)blockIdx.y;
    const long long i = (long long)a.in_walk.slot(f) * a.in_slot_elems, o = (long long)a.out_walk.slot(f) * a.out_slot_elems;
    a.x_hi += i;
    a.x_lo += i;
    a.y_hi += o;
    a.y_lo += o;
  }
  __shared__ float xs[kTileRows][8 + 1];             // normalised input rows
  __shared__ __align__(16) float xa[kSmallKMax][kTileRows + 4];  // k-major [part-major mixed inputs | raw input]
  extern __shared__ __align__(16) float w_s[];       // [K][cout] k-major weights, then bias[cout]
  const int K = (3 + a.res_conv) * a.cin;
  const long long tok0 = (long long)blockIdx.x * a.tile_tokens;
  long long remain = a.n_tokens - tok0;
  const int rows_valid = (int)(remain < a.tile_tokens ? remain : a.tile_tokens);
  float *bias_s = w_s + K * a.cout;
  pdl_trigger();
  for (int i = threadIdx.x; i < K * a.cout; i += blockDim.x) w_s[i] = a.w[i];
  for (int i = threadIdx.x; i < a.cout; i += blockDim.x) bias_s[i] = a.bias[i];
  pdl_wait();  // weights above are static; the input rows below come from the previous kernel
  for (int i = threadIdx.x; i < kTileRows * a.cin; i += blockDim.x) {
    const int r = i / a.cin, c = i - r * a.cin;
    float v = 0.f;
    if (r < rows_valid) v = join_bf16(a.x_hi[(tok0 + r) * a.cs_in + c], a.x_lo[(tok0 + r) * a.cs_in + c]);
    xs[r][c] = v;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kTileRows * K; i += blockDim.x) {
    const int k = i / kTileRows, r = i - k * kTileRows;
    const int part = k / a.cin, c = k - part * a.cin;
    float v = 0.f;
    if (r < rows_valid) {
      if (part < 3) {
        const int wv = r % a.V, sk0 = r - wv;
        if (a.dense != nullptr) {
          const float *dm = a.dense + (tok0 + r) * a.dense_ld + part * a.dense_vp;
          for (int sv = 0; sv < a.V; ++sv) v = fmaf(dm[sv], xs[sk0 + sv][c], v);
        } else {
          const int e1 = a.mix_ptr[part * a.V + wv + 1];
          for (int e = a.mix_ptr[part * a.V + wv]; e < e1; ++e) v = fmaf(a.mix_val[e], xs[sk0 + a.mix_src[e]][c], v);
        }
      } else {
        v = xs[r][c];
      }
    }
    xa[k][r] = v;
  }
  __syncthreads();
  // register tile: 4 token rows x 8 output channels per thread; 32 row groups x 8 channel groups per pass of 64 channels
  const int rg = threadIdx.x >> 3, cg = threadIdx.x & 7;
  for (int n0 = cg * 8; n0 < a.cout; n0 += 64) {
    float acc[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = bias_s[n0 + j];
    for (int k = 0; k < K; ++k) {
      const float4 xv = *reinterpret_cast<const float4 *>(&xa[k][rg * 4]);
      const float4 w0 = *reinterpret_cast<const float4 *>(&w_s[k * a.cout + n0]);
      const float4 w1 = *reinterpret_cast<const float4 *>(&w_s[k * a.cout + n0 + 4]);
      const float xr[4] = {xv.x, xv.y, xv.z, xv.w};
      const float wr[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(xr[i], wr[j], acc[i][j]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = rg * 4 + i;
      if (r >= rows_valid) continue;
      const long long tok = tok0 + r;
      uint32_t oh[4], ol[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float x0 = fmaxf(acc[i][2 * j], 0.f), x1 = fmaxf(acc[i][2 * j + 1], 0.f);
        const uint32_t h = pack_bf16x2(x0, x1);
        oh[j] = h;
        ol[j] = pack_bf16x2(x0 - bf16_lo_as_float(h), x1 - bf16_hi_as_float(h));
      }
      *reinterpret_cast<uint4 *>(a.y_hi + tok * a.cs_out + n0) = make_uint4(oh[0], oh[1], oh[2], oh[3]);
      *reinterpret_cast<uint4 *>(a.y_lo + tok * a.cs_out + n0) = make_uint4(ol[0], ol[1], ol[2], ol[3]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Adaptive graph conv, attention half (AdaptiveGraphConvolution.forward with T = 1, models/a_gcn/a_gcn.py:52-63):
// per skeleton and partition i
//     S_i[v][w] = sum_c theta_i[c][v] * phi_i[c][w] / inter_c ,  theta_i = a_conv_i(x), phi_i = b_conv_i(x)
//     M_i[v][w] = softmax over v of S_i[.][w]  +  (A + graph_attn)_i[v][w]
// and the mixing row of every output token w is written to the dense scratch the graph-conv kernels read.
// fp32 CUDA-core version (any channel count): one token tile per CTA.
// ---------------------------------------------------------------------------------------------
struct AttnArgs {
  const __nv_bfloat16 *x_hi, *x_lo;
  int cs_in, cin;
  const float *w;     // [cin][6*inter_c] k-major: columns theta_0 | phi_0 | theta_1 | phi_1 | theta_2 | phi_2
  const float *bias;  // [6*inter_c]
  int inter_c;
  const float *adj;   // [3][V][V] static term A + graph_attn
  int V;
  long long n_tokens;
  int tile_tokens;
  float *dense;
  int dense_ld, dense_vp;
};

constexpr int kAttnMaxV = 32;
constexpr int kAttnMaxInter = 64;                                      // inter_c of a 256-channel block
constexpr int kAttnMaxSmem = 2 * kAttnMaxInter * kTileRows * 4;        // dynamic shared memory of k_agcn_attn

__global__ void __launch_bounds__(256) k_agcn_attn(AttnArgs a) {
  __shared__ __align__(16) float Xs[kSimtK][kTileRows];
  __shared__ __align__(16) float Bs[kSimtK][kSimtN];
  extern __shared__ __align__(16) float tp[];  // [2*inter_c][kTileRows]: theta rows, then phi rows, of one partition
  pdl_trigger();
  pdl_wait();
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  const long long tok0 = (long long)blockIdx.x * a.tile_tokens;
  long long remain = a.n_tokens - tok0;
  const int rows_valid = (int)(remain < a.tile_tokens ? remain : a.tile_tokens);
  const int ic = a.inter_c, ld = 6 * ic;
  for (int part = 0; part < 3; ++part) {
    for (int n0 = 0; n0 < 2 * ic; n0 += kSimtN) {
      float acc[8][4];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
      for (int c0 = 0; c0 < a.cin; c0 += kSimtK) {
        __syncthreads();
        simt_load_rows(Xs, a.x_hi, a.x_lo, a.cs_in, c0, a.cin, tok0, rows_valid);
        simt_load_w(Bs, a.w, ld, c0, min(kSimtK, a.cin - c0), part * 2 * ic + n0, (part + 1) * 2 * ic);
        __syncthreads();
        simt_mma_chunk(acc, Xs, Bs, ty, tx);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int n = n0 + tx * 4 + j;
        if (n >= 2 * ic) continue;
        const float b = a.bias[part * 2 * ic + n];
#pragma unroll
        for (int i = 0; i < 8; ++i) tp[n * kTileRows + ty * 8 + i] = acc[i][j] + b;
      }
    }
    __syncthreads();
    if (threadIdx.x < rows_valid) {
      const int row = threadIdx.x;
      const int wv = row % a.V, sk0 = row - wv;
      float sv[kAttnMaxV];
      float mx = -INFINITY;
      for (int v = 0; v < a.V; ++v) {
        float d = 0.f;
        for (int c = 0; c < ic; ++c) d = fmaf(tp[c * kTileRows + sk0 + v], tp[(ic + c) * kTileRows + row], d);
        d = d / (float)ic;
        sv[v] = d;
        mx = fmaxf(mx, d);
      }
      float sum = 0.f;
      for (int v = 0; v < a.V; ++v) {
        sv[v] = expf(sv[v] - mx);
        sum += sv[v];
      }
      float *dst = a.dense + (tok0 + row) * a.dense_ld + part * a.dense_vp;
      for (int v = 0; v < a.V; ++v) dst[v] = sv[v] / sum + a.adj[((size_t)part * a.V + v) * a.V + wv];
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// Spatial self-attention unit (GcnUnitAttention with only_attention, models/s_tr/s_tr.py:417-476), two kernels
// in front of the output conv (which is k_tcn_simt / k_tc_tcn with a single tap):
//   k_sa_qkv:  qkv = Wqkv (data_bn(x)) + b for one token tile x 64 columns, fp32 -> scratch [token][2*dk + dv]
//   k_sa_attn: per token row i and head h:  w_j = softmax_j <q_i, k_j>,  out_i = sum_j w_j v_j  over the V vertices
//              of the row's skeleton -> split-bf16 rows [token][dv] (the output conv's operand)
// ---------------------------------------------------------------------------------------------
struct SaQkvArgs {
  const __nv_bfloat16 *x_hi, *x_lo;
  int cs_in, cin;
  const float *in_scale, *in_shift;  // [cin][V]
  const float *w;                    // [cin][nq] k-major, nq = 2*dk + dv
  const float *bias;                 // [nq]
  int nq, V;
  long long n_tokens;
  int tile_tokens;
  float *qkv;  // [token][nq]
};

__global__ void __launch_bounds__(256) k_sa_qkv(SaQkvArgs a) {
  __shared__ __align__(16) float Xs[kSimtK][kTileRows];
  __shared__ __align__(16) float Bs[kSimtK][kSimtN];
  pdl_trigger();
  pdl_wait();
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  const long long tok0 = (long long)blockIdx.x * a.tile_tokens;
  const int n0 = blockIdx.y * kSimtN;
  long long remain = a.n_tokens - tok0;
  const int rows_valid = (int)(remain < a.tile_tokens ? remain : a.tile_tokens);
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int c0 = 0; c0 < a.cin; c0 += kSimtK) {
    __syncthreads();
    {  // rows with the per-(channel, vertex) affine of data_bn applied
      const int row = threadIdx.x >> 1;
      const int half = (threadIdx.x & 1) * 8;
      const bool ok = row < rows_valid;
      const long long base = (tok0 + row) * a.cs_in;
      const int wv = row % a.V;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = c0 + half + j;
        float v = 0.f;
        if (ok && c < a.cin) v = join_bf16(a.x_hi[base + c], a.x_lo[base + c]) * a.in_scale[c * a.V + wv] + a.in_shift[c * a.V + wv];
        Xs[half + j][row] = v;
      }
    }
    simt_load_w(Bs, a.w, a.nq, c0, min(kSimtK, a.cin - c0), n0, a.nq);
    __syncthreads();
    simt_mma_chunk(acc, Xs, Bs, ty, tx);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = ty * 8 + i;
    if (row >= rows_valid) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n < a.nq) a.qkv[(tok0 + row) * a.nq + n] = acc[i][j] + a.bias[n];
    }
  }
}

// data_bn of the attention unit as a pre-pass (tensor-core path): x'[c] = x[c] * scale[c][v] + shift[c][v] on the
// split-bf16 rows, 8 channels per thread
struct SaAffineArgs {
  const __nv_bfloat16 *x_hi, *x_lo;
  __nv_bfloat16 *y_hi, *y_lo;
  int cs, c, V;  // same row stride in and out
  const float *scale, *shift;
  long long n_tokens;
};

__global__ void __launch_bounds__(256) k_sa_affine(SaAffineArgs a) {
  pdl_trigger();
  pdl_wait();
  const int per_row = a.c / 8;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= a.n_tokens * per_row) return;
  const long long tok = idx / per_row;
  const int c0 = (int)(idx - tok * per_row) * 8;
  const int wv = (int)(tok % a.V);
  const uint4 h = *reinterpret_cast<const uint4 *>(a.x_hi + tok * a.cs + c0);
  const uint4 l = *reinterpret_cast<const uint4 *>(a.x_lo + tok * a.cs + c0);
  const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
  uint32_t oh[4], ol[4];
#pragma unroll
  for (int w = 0; w < 4; ++w) {
    const int c = c0 + 2 * w;
    const float x0 = (bf16_lo_as_float(hw[w]) + bf16_lo_as_float(lw[w])) * a.scale[c * a.V + wv] + a.shift[c * a.V + wv];
    const float x1 = (bf16_hi_as_float(hw[w]) + bf16_hi_as_float(lw[w])) * a.scale[(c + 1) * a.V + wv] + a.shift[(c + 1) * a.V + wv];
    oh[w] = pack_bf16x2(x0, x1);
    ol[w] = pack_bf16x2(x0 - bf16_lo_as_float(oh[w]), x1 - bf16_hi_as_float(oh[w]));
  }
  *reinterpret_cast<uint4 *>(a.y_hi + tok * a.cs + c0) = make_uint4(oh[0], oh[1], oh[2], oh[3]);
  *reinterpret_cast<uint4 *>(a.y_lo + tok * a.cs + c0) = make_uint4(ol[0], ol[1], ol[2], ol[3]);
}

struct SaAttnArgs {
  const float *qkv;                  // fp32 rows [token][nq] (CUDA-core qkv conv) ...
  const __nv_bfloat16 *q_hi, *q_lo;  // ... or split-bf16 rows [token][cs_q] (tensor-core qkv conv); nullptr selects fp32
  int cs_q;
  int dk, dv, V;  // 8 heads: dkh = dk/8, dvh = dv/8
  long long n_tokens;
  int tile_tokens;
  __nv_bfloat16 *y_hi, *y_lo;  // [token][cs_out]
  int cs_out;
};

constexpr int kSaHeads = 8;    // Nh of GcnUnitAttention, never overridden by the reference (models/s_tr/s_tr.py:311)
constexpr int kSaMaxDkh = 8;   // dk / heads of a 256-channel unit
constexpr int kSaMaxDvh = 32;  // dv / heads

// N consecutive columns of a token's q | k | v row as fp32: 16-byte loads of the hi and lo planes where N allows,
// fp32 loads when the row comes from the CUDA-core qkv conv
template <int N>
__device__ __forceinline__ void sa_load_cols(const SaAttnArgs &a, long long tok, int col, float (&out)[N]) {
  if (a.q_hi == nullptr) {
    const float *p = a.qkv + tok * (2 * a.dk + a.dv) + col;
#pragma unroll
    for (int i = 0; i < N; ++i) out[i] = p[i];
    return;
  }
  const __nv_bfloat16 *ph = a.q_hi + tok * a.cs_q + col, *pl = a.q_lo + tok * a.cs_q + col;
  if (N % 8 == 0) {
#pragma unroll
    for (int g = 0; g < N / 8; ++g) {
      const uint4 h = *reinterpret_cast<const uint4 *>(ph + 8 * g), l = *reinterpret_cast<const uint4 *>(pl + 8 * g);
      const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
      for (int w = 0; w < 4; ++w) {
        out[8 * g + 2 * w] = bf16_lo_as_float(hw[w]) + bf16_lo_as_float(lw[w]);
        out[8 * g + 2 * w + 1] = bf16_hi_as_float(hw[w]) + bf16_hi_as_float(lw[w]);
      }
    }
  } else {  // 2 or 4 columns: 4-byte loads
#pragma unroll
    for (int w = 0; w < N / 2; ++w) {
      const uint32_t h = *reinterpret_cast<const uint32_t *>(ph + 2 * w), l = *reinterpret_cast<const uint32_t *>(pl + 2 * w);
      out[2 * w] = bf16_lo_as_float(h) + bf16_lo_as_float(l);
      out[2 * w + 1] = bf16_hi_as_float(h) + bf16_hi_as_float(l);
    }
  }
}

// One token tile per CTA, 256 threads = 2 heads x 128 token rows; per head: k and v of the tile staged in shared
// memory as fp32 (each element converted once), then for row i:  w_j = softmax_j <q_i, k_j> over the V vertices of its
// skeleton and out_i = sum_j w_j v_j, from registers and 16-byte broadcast loads; the dvh outputs leave as packed
// split-bf16 rows.  DVH = dv / 8, dkh = DVH / 4.
template <int DVH>
__global__ void __launch_bounds__(256) k_sa_attn(SaAttnArgs a) {
  constexpr int DKH = DVH / 4;
  constexpr int KP = DKH < 4 ? 4 : DKH;  // padded to whole float4s
  __shared__ __align__(16) float ks[2][kTileRows][KP];
  __shared__ __align__(16) float vs[2][kTileRows][DVH + 4];
  pdl_trigger();
  pdl_wait();
  const int row = threadIdx.x & 127, hh = threadIdx.x >> 7;
  const long long tok0 = (long long)blockIdx.x * a.tile_tokens;
  long long remain = a.n_tokens - tok0;
  const int rows_valid = (int)(remain < a.tile_tokens ? remain : a.tile_tokens);
  const bool valid = row < rows_valid;
  const long long tok = tok0 + row;
  const int sk0 = valid ? row - row % a.V : 0;
  for (int h = hh; h < kSaHeads; h += 2) {
    __syncthreads();
    if (valid) {
      float kk[DKH], vv[DVH];
      sa_load_cols<DKH>(a, tok, a.dk + h * DKH, kk);
      sa_load_cols<DVH>(a, tok, 2 * a.dk + h * DVH, vv);
#pragma unroll
      for (int d = 0; d < KP; ++d) ks[hh][row][d] = d < DKH ? kk[d] : 0.f;
#pragma unroll
      for (int d4 = 0; d4 < DVH / 4; ++d4)
        *reinterpret_cast<float4 *>(&vs[hh][row][4 * d4]) = make_float4(vv[4 * d4], vv[4 * d4 + 1], vv[4 * d4 + 2], vv[4 * d4 + 3]);
    }
    __syncthreads();
    if (!valid) continue;
    float q[KP];
    {
      float qq[DKH];
      sa_load_cols<DKH>(a, tok, h * DKH, qq);
#pragma unroll
      for (int d = 0; d < KP; ++d) q[d] = d < DKH ? qq[d] : 0.f;
    }
    float w[kAttnMaxV];
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < kAttnMaxV; ++j) {
      if (j < a.V) {
        float sdot = 0.f;
#pragma unroll
        for (int d4 = 0; d4 < KP / 4; ++d4) {
          const float4 k4 = *reinterpret_cast<const float4 *>(&ks[hh][sk0 + j][4 * d4]);
          sdot = fmaf(q[4 * d4], k4.x, sdot);
          sdot = fmaf(q[4 * d4 + 1], k4.y, sdot);
          sdot = fmaf(q[4 * d4 + 2], k4.z, sdot);
          sdot = fmaf(q[4 * d4 + 3], k4.w, sdot);
        }
        w[j] = sdot;
        mx = fmaxf(mx, sdot);
      } else {
        w[j] = 0.f;
      }
    }
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < kAttnMaxV; ++j)
      if (j < a.V) {
        w[j] = expf(w[j] - mx);
        sum += w[j];
      }
    const float inv = 1.0f / sum;
    float o[DVH];
#pragma unroll
    for (int d = 0; d < DVH; ++d) o[d] = 0.f;
#pragma unroll
    for (int j = 0; j < kAttnMaxV; ++j)
      if (j < a.V) {
        const float wj = w[j] * inv;
#pragma unroll
        for (int d4 = 0; d4 < DVH / 4; ++d4) {
          const float4 v4 = *reinterpret_cast<const float4 *>(&vs[hh][sk0 + j][4 * d4]);
          o[4 * d4] = fmaf(wj, v4.x, o[4 * d4]);
          o[4 * d4 + 1] = fmaf(wj, v4.y, o[4 * d4 + 1]);
          o[4 * d4 + 2] = fmaf(wj, v4.z, o[4 * d4 + 2]);
          o[4 * d4 + 3] = fmaf(wj, v4.w, o[4 * d4 + 3]);
        }
      }
#pragma unroll
    for (int g = 0; g < DVH / 8; ++g) {
      uint32_t oh[4], ol[4];
#pragma unroll
      for (int w2 = 0; w2 < 4; ++w2) {
        const float x0 = o[8 * g + 2 * w2], x1 = o[8 * g + 2 * w2 + 1];
        oh[w2] = pack_bf16x2(x0, x1);
        ol[w2] = pack_bf16x2(x0 - bf16_lo_as_float(oh[w2]), x1 - bf16_hi_as_float(oh[w2]));
      }
      *reinterpret_cast<uint4 *>(a.y_hi + tok * a.cs_out + h * DVH + 8 * g) = make_uint4(oh[0], oh[1], oh[2], oh[3]);
      *reinterpret_cast<uint4 *>(a.y_lo + tok * a.cs_out + h * DVH + 8 * g) = make_uint4(ol[0], ol[1], ol[2], ol[3]);
    }
  }
}

// any head width (the reference's small test blocks): scalar version of the above
__global__ void __launch_bounds__(128) k_sa_attn_any(SaAttnArgs a) {
  __shared__ float ks[kTileRows][kSaMaxDkh + 1];
  __shared__ float vs[kTileRows][kSaMaxDvh + 1];
  pdl_trigger();
  pdl_wait();
  const int row = threadIdx.x;
  const long long tok0 = (long long)blockIdx.x * a.tile_tokens;
  long long remain = a.n_tokens - tok0;
  const int rows_valid = (int)(remain < a.tile_tokens ? remain : a.tile_tokens);
  const bool valid = row < rows_valid;
  const int nq = 2 * a.dk + a.dv, dkh = a.dk / kSaHeads, dvh = a.dv / kSaHeads;
  const int sk0 = valid ? row - row % a.V : 0;
  const float *mine = a.qkv + (tok0 + row) * nq;
  for (int h = 0; h < kSaHeads; ++h) {
    __syncthreads();
    if (valid) {
      for (int d = 0; d < dkh; ++d) ks[row][d] = mine[a.dk + h * dkh + d];
      for (int d = 0; d < dvh; ++d) vs[row][d] = mine[2 * a.dk + h * dvh + d];
    }
    __syncthreads();
    if (!valid) continue;
    float q[kSaMaxDkh];
#pragma unroll
    for (int d = 0; d < kSaMaxDkh; ++d) q[d] = d < dkh ? mine[h * dkh + d] : 0.f;
    float w[kAttnMaxV];
    float mx = -INFINITY;
    for (int j = 0; j < a.V; ++j) {
      float sdot = 0.f;
#pragma unroll
      for (int d = 0; d < kSaMaxDkh; ++d)
        if (d < dkh) sdot = fmaf(q[d], ks[sk0 + j][d], sdot);
      w[j] = sdot;
      mx = fmaxf(mx, sdot);
    }
    float sum = 0.f;
    for (int j = 0; j < a.V; ++j) {
      w[j] = expf(w[j] - mx);
      sum += w[j];
    }
    const float inv = 1.0f / sum;
    for (int d = 0; d < dvh; ++d) {
      float o = 0.f;
      for (int j = 0; j < a.V; ++j) o = fmaf(w[j] * inv, vs[sk0 + j][d], o);
      __nv_bfloat16 hh, ll;
      split_bf16(o, hh, ll);
      a.y_hi[(tok0 + row) * a.cs_out + h * dvh + d] = hh;
      a.y_lo[(tok0 + row) * a.cs_out + h * dvh + d] = ll;
    }
  }
}

struct TcnArgs {
  const __nv_bfloat16 *tap_hi[kTaps], *tap_lo[kTaps];  // oldest .. newest ring slots
  int n_taps;                                           // kTaps; 1 when the kernel serves as the attention unit's output conv
  int cs, c;                                            // row stride / channels of the ring
  const float *w;                                       // [9*c][c] k-major (k = tap*c + ci), BN folded
  const __nv_bfloat16 *r_hi, *r_lo;                     // block input of 4 executions ago
  int cs_r, cr, res_kind;
  const float *w_r;   // [cr][c] k-major folded residual conv (res_kind == 2)
  const float *bias;  // [c]
  __nv_bfloat16 *y_hi, *y_lo;
  int cs_out;
  long long n_tokens;
  int tile_tokens;
};

// 9-tap temporal convolution over the ring + BN + delayed residual + ReLU
// (co.Conv2d step + BatchNorm2d, models/base.py:307-334; residual wiring :412-446).
__global__ void __launch_bounds__(256) k_tcn_simt(TcnArgs a) {
  __shared__ __align__(16) float As[kSimtK][kTileRows];
  __shared__ __align__(16) float Bs[kSimtK][kSimtN];
  pdl_trigger();
  pdl_wait();
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  const long long tok0 = (long long)blockIdx.x * a.tile_tokens;
  const int n0 = blockIdx.y * kSimtN;
  long long remain = a.n_tokens - tok0;
  const int rows_valid = (int)(remain < a.tile_tokens ? remain : a.tile_tokens);
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int segs = a.n_taps + (a.res_kind == 2 ? 1 : 0);
  for (int seg = 0; seg < segs; ++seg) {
    const bool is_res = seg == a.n_taps;
    const __nv_bfloat16 *hi = is_res ? a.r_hi : a.tap_hi[seg];
    const __nv_bfloat16 *lo = is_res ? a.r_lo : a.tap_lo[seg];
    const int cs = is_res ? a.cs_r : a.cs;
    const int cc = is_res ? a.cr : a.c;
    const float *w = is_res ? a.w_r : a.w + (long long)seg * a.c * a.c;
    for (int c0 = 0; c0 < cc; c0 += kSimtK) {
      simt_load_rows(As, hi, lo, cs, c0, cc, tok0, rows_valid);
      simt_load_w(Bs, w, a.c, c0, min(kSimtK, cc - c0), n0, a.c);
      __syncthreads();
      simt_mma_chunk(acc, As, Bs, ty, tx);
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = ty * 8 + i;
    if (row >= rows_valid) continue;
    const long long tok = tok0 + row;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= a.c) continue;
      float v = acc[i][j] + a.bias[n];
      if (a.res_kind == 1) v += join_bf16(a.r_hi[tok * a.cs_r + n], a.r_lo[tok * a.cs_r + n]);
      v = fmaxf(v, 0.f);
      __nv_bfloat16 h, l;
      split_bf16(v, h, l);
      a.y_hi[tok * a.cs_out + n] = h;
      a.y_lo[tok * a.cs_out + n] = l;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// head: spatial mean over V and S, sliding temporal mean over the last P pooled vectors, FC.
// Reference: spatial_pool / co.AvgPool1d / co.Linear, models/base.py:84,97,99.
// One CTA per stream; window sum kept in fp64 (adding and removing the same fp32 values is then
// exact, so the running sum cannot drift); class dot products reduced with warp shuffles.
// ---------------------------------------------------------------------------------------------
struct HeadArgs {
  const __nv_bfloat16 *y_hi, *y_lo;
  int cs, c, V, S;
  float *ring;   // [P][N][c] pooled vectors
  double *sum;   // [N][c]     running window sum
  int slot, P;
  long long n_streams;
  int emit;
  const float *w, *b;  // [classes][c], [classes]
  int classes;
  float *out;  // [N][classes]
  // time-batched launch: n_frames pooled frames in sequence (slot in_walk.slot(f) of the last block's output ring; y_* then
  // point at slot 0), window slot (slot + f) % P, emission from frame first_emit on, emission j written at out + j * out_stride
  int n_frames = 1;
  RingWalk in_walk;
  long long in_slot_elems = 0;
  int first_emit = 0;
  long long out_stride = 0;
};

constexpr int kHeadStreams = 4;  // streams per CTA (256 threads each): the FC weight rows are fetched once for all of them

__global__ void __launch_bounds__(256 * kHeadStreams) k_head(HeadArgs a) {
  extern __shared__ float head_s[];
  pdl_trigger();
  pdl_wait();
  const int g = threadIdx.x >> 8, tid = threadIdx.x & 255;
  float *part = head_s + g * 2 * a.cs;                    // [streams][2][cs] partial spatial means
  float *mean_s = head_s + kHeadStreams * 2 * a.cs;       // [streams][c] window means
  const long long n0 = (long long)blockIdx.x * kHeadStreams;
  const long long n = n0 + g;
  const bool live = n < a.n_streams;
  const int half = tid >> 7, t128 = tid & 127;
  for (int f = 0; f < a.n_frames; ++f) {
    const bool emit = a.n_frames == 1 ? a.emit != 0 : f >= a.first_emit;
    const int wslot = (a.slot + f) % a.P;
    const __nv_bfloat16 *y_hi = a.y_hi, *y_lo = a.y_lo;
    if (a.n_frames > 1) {
      const long long o = (long long)a.in_walk.slot(f) * a.in_slot_elems;
      y_hi += o;
      y_lo += o;
    }
    if (live && y_hi == nullptr) {
      // end-of-sequence flush of the pooling window (pad_end): a zero vector enters (co.AvgPool1d's end padding)
      for (int c = t128 + 128 * half; c < 2 * a.cs; c += 256) part[c] = 0.f;
    } else if (live) {
      const long long tok0 = n * a.S * a.V;
      // two thread halves split the skeletons of the stream; each thread owns a bf16x2 channel pair
      for (int cp = t128; cp < a.cs / 2; cp += 128) {
        float t0 = 0.f, t1 = 0.f;
        for (int s = half; s < a.S; s += 2) {
          float p0 = 0.f, p1 = 0.f;
          const long long base = (tok0 + (long long)s * a.V) * a.cs + 2 * cp;
          // rows in batches of 13 (26 loads in flight per thread: the kernel is bound by HBM latency x bytes in flight);
          // the sums run in vertex order whatever the batch size
          constexpr int kVB = 13;
          for (int v0 = 0; v0 < a.V; v0 += kVB) {
            uint32_t h[kVB], l[kVB];
#pragma unroll
            for (int j = 0; j < kVB; ++j) {
              const bool on = v0 + j < a.V;
              h[j] = on ? __ldg(reinterpret_cast<const unsigned int *>(y_hi + base + (long long)(v0 + j) * a.cs)) : 0u;
              l[j] = on ? __ldg(reinterpret_cast<const unsigned int *>(y_lo + base + (long long)(v0 + j) * a.cs)) : 0u;
            }
#pragma unroll
            for (int j = 0; j < kVB; ++j)
              if (v0 + j < a.V) {
                p0 += bf16_lo_as_float(h[j]) + bf16_lo_as_float(l[j]);
                p1 += bf16_hi_as_float(h[j]) + bf16_hi_as_float(l[j]);
              }
          }
          t0 += p0 / (float)a.V;
          t1 += p1 / (float)a.V;
        }
        part[half * a.cs + 2 * cp] = t0;
        part[half * a.cs + 2 * cp + 1] = t1;
      }
    }
    __syncthreads();
    if (live) {
      for (int c = tid; c < a.c; c += 256) {
        const float h = (part[c] + part[a.cs + c]) / (float)a.S;
        const long long ri = ((long long)wslot * a.n_streams + n) * a.c + c;
        const float old = a.ring[ri];
        a.ring[ri] = h;
        const double s2 = a.sum[n * a.c + c] + ((double)h - (double)old);
        a.sum[n * a.c + c] = s2;
        mean_s[g * a.c + c] = (float)(s2 / (double)a.P);
      }
    }
    __syncthreads();
    if (!emit) continue;
    float *out = a.out + (a.n_frames == 1 ? 0 : (long long)(f - a.first_emit) * a.out_stride);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int k = warp; k < a.classes; k += nw) {
      float p[kHeadStreams];
#pragma unroll
      for (int j = 0; j < kHeadStreams; ++j) p[j] = 0.f;
      for (int c = lane; c < a.c; c += 32) {
        const float wv = a.w[(long long)k * a.c + c];
#pragma unroll
        for (int j = 0; j < kHeadStreams; ++j) p[j] = fmaf(wv, mean_s[j * a.c + c], p[j]);
      }
#pragma unroll
      for (int j = 0; j < kHeadStreams; ++j) {
        float v = p[j];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0 && n0 + j < a.n_streams) out[(n0 + j) * a.classes + k] = v + a.b[k];
      }
    }
    __syncthreads();  // mean_s / part are rewritten by the next frame
  }
}

// token-major split-bf16 rows -> fp32 (B, C, V)
__global__ void k_read_block(const __nv_bfloat16 *__restrict__ hi, const __nv_bfloat16 *__restrict__ lo, int cs, int C,
                             int V, long long n_tokens, float *__restrict__ dst) {
  pdl_trigger();
  pdl_wait();
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_tokens * C) return;
  const int v = (int)(idx % V);
  const long long t2 = idx / V;
  const int c = (int)(t2 % C);
  const long long skel = t2 / C;
  const long long i = (skel * V + v) * cs + c;
  dst[idx] = join_bf16(hi[i], lo[i]);
}

}  // namespace cosk
